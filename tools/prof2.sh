mkdir -p gpurun_out
for v in 7 10; do
DM_KERNEL_VARIANT=$v,0 ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/prof_wide_v$v -f python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_wide_v$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
