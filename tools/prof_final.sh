# ncu --set full of the three shipped kernel shapes at the end of round 1 (current code, default variants)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/prof_wide_shipped -f python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/prof_wide_shipped.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sha256_deep -s 3 -c 1 -o gpurun_out/prof_deep_shipped -f python bench.py --blobs 256 --blob-bytes 1048576 --steps 1 --warmup 3 --no-e2e --no-cpu --no-probes > gpurun_out/prof_deep_shipped.log 2>&1
ls -la gpurun_out/*shipped*
