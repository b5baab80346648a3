#!/bin/bash
# round 2, final single-GPU call: the whole GPU suite on the final defaults (wide variant 21, streams-per-warp rule), the
# default bench line, the saturation probe, compute-sanitizer over the kernel edge-case tests (incl. the cp.async wide
# kernel: racecheck), ncu of the wide kernel as shipped, soak
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests.txt
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
timeout 300 python bench.py --workload saturate_151552x112KiB --steps 3 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_saturate.json 2> gpurun_out/r02_saturate.err
K="fips or every_length or boundaries or ragged or edge_sizes or reference_fixture or device_generator or skewed"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" | tail -6
done > gpurun_out/r02_sanitizer.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/r02_prof_wide -f \
    python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_prof_wide.log 2>&1
SECONDS=40 THREADS=12 timeout 200 python tools/soak.py > gpurun_out/r02_soak_hbm.txt 2>&1
SECONDS=30 THREADS=12 CAS_DIR=/dev/shm/dm_soak_r02 timeout 200 python tools/soak.py > gpurun_out/r02_soak_disk.txt 2>&1
