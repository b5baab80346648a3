#!/bin/bash
# round 2, GPU call 4: rolled round loops A/B, ncu of the new default deep kernel, tiny-body and bulk-index probes,
# BASELINE configs[1] as a bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 9,7 9,15 9,13 9,12 9,8; do timeout 200 python tools/check_variant.py $v | tail -2; done > gpurun_out/r02_variants_rolled.txt 2>&1
# ncu: the deep kernel as shipped (256 x 1 MiB), full set with source-level stall samples
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_deep -s 3 -c 1 -o gpurun_out/r02_prof_deep -f \
    python bench.py --blobs 256 --blob-bytes 1048576 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_prof_deep.log 2>&1
DM_KERNEL_VARIANT=9,15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_deep -s 3 -c 1 -o gpurun_out/r02_prof_deep_rolled -f \
    python bench.py --blobs 256 --blob-bytes 1048576 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_prof_deep_rolled.log 2>&1
timeout 300 python tools/tiny_blob_probe.py > gpurun_out/r02_tiny_blob.txt 2>&1
timeout 300 python bench.py --workload saturate_151552x112KiB --steps 3 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/r02_saturate.json 2> gpurun_out/r02_saturate.err
timeout 900 python bench.py --workload llama3_8b_shards --steps 1 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_llama3_8b_shards.json 2> gpurun_out/r02_bench_llama3.err
SECONDS=40 THREADS=12 timeout 200 python tools/soak.py > gpurun_out/r02_soak_hbm.txt 2>&1
SECONDS=30 THREADS=12 CAS_DIR=/dev/shm/dm_soak_r02 timeout 200 python tools/soak.py > gpurun_out/r02_soak_disk.txt 2>&1
