# ncu passes on the exact default bench command (B200_PROFILING.md recipe); outputs under gpurun_out/
mkdir -p gpurun_out
# 1. launch list with device time per launch (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-serve > gpurun_out/r01_launches_default.log 2>&1
# 2. DRAM bytes of the dominant kernel at the full workload size (1 pass: two metrics)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sha256_deep -s 4 -c 1 --csv \
    --log-file gpurun_out/r01_dram_default.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01_dram_default.log 2>&1
# 3. full set on the group kernel (S=8, 4096 streams) for the record
ncu --set full --clock-control none --import-source on -k regex:sha256_group -s 3 -c 1 -o gpurun_out/prof_group_s8 -f \
    python bench.py --blobs 4096 --blob-bytes 262144 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_group_s8.log 2>&1
ls -la gpurun_out | tail -8
