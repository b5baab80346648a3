# NOTE: written when DM_KERNEL_VARIANT took a single fma digit (0 = ptxas choice, 1 = adds on the FMA pipe); still valid: 0 and 1 are style-0 variants.
mkdir -p gpurun_out
# launch list (cold-cache, serialised) of the default bench command, short
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_default.csv python bench.py --blobs 256 --blob-bytes 4194304 --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_default.log 2>&1
# full capture: wide kernel (variant 0 and 1), deep kernel (variant 0)
DM_KERNEL_VARIANT=0 ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/prof_wide_v0 -f python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_wide_v0.log 2>&1
DM_KERNEL_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/prof_wide_v1 -f python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_wide_v1.log 2>&1
DM_KERNEL_VARIANT=0 ncu --set full --clock-control none --import-source on -k regex:sha256_deep -s 3 -c 1 -o gpurun_out/prof_deep_v0 -f python bench.py --blobs 256 --blob-bytes 1048576 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_deep_v0.log 2>&1
ls -la gpurun_out/
