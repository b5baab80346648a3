#!/bin/bash
# round 2, closing call: DRAM bytes of one launch of the dominant kernel on the final build (roofline.traffic), the
# regression test for equal-sized long batches, and the 1024-stream point of the default-rule sweep again.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sha256_deep -s 4 -c 1 --csv \
    --log-file gpurun_out/r02_dram_default.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-probes > gpurun_out/r02_dram_default.log 2>&1
sed -n '/^python - <<.PY.$/,/^PY$/p' tools/prof_r02.sh | sed '1d;$d' > /tmp/traffic_extract.py
python /tmp/traffic_extract.py > gpurun_out/r02_traffic_extract.txt 2>&1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "equal_sized_long or skewed or two_warps" > gpurun_out/r02_final3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final3_pytest.txt
run() {
  bytes=$(( (8<<30) / $1 / 256 * 256 ))
  timeout 60 python bench.py --blobs $1 --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu --no-probes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('default rule streams=%6d kernel=%-32s kernel_ms=%9.2f hash_GBps=%7.1f' % ($1, d['config'].get('kernel'), r['kernel_ms_per_step'], r['achieved']/2))"
}
{ run 1024; run 700; } > gpurun_out/r02_final3_sweep.txt 2>&1
cat gpurun_out/r02_traffic_extract.txt | tail -8; tail -2 gpurun_out/r02_final3_pytest.txt; cat gpurun_out/r02_final3_sweep.txt
