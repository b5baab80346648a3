#!/bin/bash
# round 2, last call: the pair rule as the default (deep variant 9, streams_per_warp_for over warp pairs):
# smoke, the whole GPU suite, racecheck/synccheck of the two-warp group kernels, the default bench line, and the
# stream-count sweep under the default rule (no forcing).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final2_smoke.txt
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final2_pytest.txt
timeout 150 python bench.py > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err
run() {  # n
  bytes=$(( (8<<30) / $1 / 256 * 256 ))
  timeout 100 python bench.py --blobs $1 --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu --no-probes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('default rule streams=%6d kernel=%-32s kernel_ms=%9.2f hash_GBps=%7.1f' % ($1, d['config'].get('kernel'), r['kernel_ms_per_step'], r['achieved']/2))"
}
{ run 296; run 512; run 592; run 1024; run 2048; run 4096; run 4736; run 8192; } > gpurun_out/r02_final2_sweep.txt 2>&1
for tool in racecheck synccheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fips or two_warps_per_group" > gpurun_out/r02_final2_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02_final2_$tool.txt
done
tail -3 gpurun_out/r02_final2_pytest.txt; cat gpurun_out/r02_final2_sweep.txt; tail -2 gpurun_out/r02_final2_racecheck.txt
