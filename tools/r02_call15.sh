#!/bin/bash
# round 2, final validation on the final defaults (wide 21, deep 8): smoke, the whole GPU suite, the default bench line,
# racecheck / memcheck over the kernel edge cases, launch list + DRAM bytes of the default workload, ncu of the two-warp kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests.txt
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
K="fips or every_length or boundaries or ragged or edge_sizes or reference_fixture or device_generator or skewed or two_warps"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" | tail -6
done > gpurun_out/r02_sanitizer.txt 2>&1
bash tools/prof_r02.sh > gpurun_out/r02_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_deep -s 3 -c 1 -o gpurun_out/r02_prof_deep2 -f \
    python bench.py --blobs 256 --blob-bytes 1048576 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_prof_deep2.log 2>&1
