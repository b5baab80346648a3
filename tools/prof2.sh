mkdir -p gpurun_out
for v in 9 13; do   # fma + 4*style: 9 = shipped (fma 1, style 2), 13 = fma 1, style 3 (these were "v7"/"v10" in the 3-based numbering of the round-1 profile file names)
DM_KERNEL_VARIANT=$v,0 ncu --set full --clock-control none --import-source on -k regex:sha256_wide -s 3 -c 1 -o gpurun_out/prof_wide_v$v -f python bench.py --blobs 151552 --blob-bytes 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_wide_v$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
