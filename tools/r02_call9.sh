#!/bin/bash
# round 2, GPU call 9: wide kernel with cp.async-staged lines (style 5 = variant 21) vs the shipped variant 9; streams-per-warp
# crossovers with the variant-7 round; size sweep after the skewed-batch split; tiny-body probe; the parity suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 9,7 21,7; do timeout 200 python tools/check_variant.py $v | tail -4; done > gpurun_out/r02_wide_style5_parity.txt 2>&1
VARIANTS="9 21" EXTRA="--no-probes" timeout 400 bash tools/ab_wide.sh > gpurun_out/r02_ab_wide.txt 2>&1
# crossovers: each stream count with the kernel shapes on either side of the current threshold (8 GiB per run)
for pair in "512:1 2 4" "640:1 2 4" "768:1 2 4" "1024:1 2 4" "2048:2 4 8" "3072:4 8" "4096:4 8" "6144:8 16" "8192:8 16" "12288:16 32" "16384:16 32"; do
  n=${pair%%:*}; bytes=$(( (8<<30) / n / 256 * 256 ))
  for spw in ${pair#*:}; do
    DM_FORCE_SPW=$spw timeout 120 python bench.py --blobs $n --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu --no-probes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('streams=%6d spw=%2d kernel_ms=%9.2f hash_GBps=%7.1f' % ($n, $spw, r['kernel_ms_per_step'], r['achieved']/2))"
  done
done > gpurun_out/r02_streams_per_warp_sweep.txt 2>&1
timeout 600 python bench.py --workload size_sweep --steps 2 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_size_sweep.json 2> gpurun_out/r02_bench_size_sweep.err
timeout 300 python tools/tiny_blob_probe.py > gpurun_out/r02_tiny_blob.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests.txt
