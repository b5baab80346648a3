#!/bin/bash
# One GPU call that decides whether the short-chain round (deep variant 4) becomes the default.
#   gpurun --timeout 600 -- 'bash tools/ab_deep.sh > gpurun_out/ab_deep.txt 2>&1'
# 1. latencies the SASS timing model assumes (tools/ubench/chainlat.cu)
# 2. parity of variant 4 on every kernel shape + its deep-kernel time next to the default's
# 3. the headline workload (256 x 64 MiB) and the 2048/4096-stream group regimes, default vs variant 4
set -u
cd "$(dirname "$0")/.."
echo "== chainlat"; (cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chainlat chainlat.cu && timeout 60 ./chainlat)
for v in 9,0 9,4; do
  echo "== check_variant $v"; timeout 240 python tools/check_variant.py $v | tail -12
done
for v in 9,0 9,4; do
  echo "== bench hf_lfs_256x64MiB DM_KERNEL_VARIANT=$v"
  DM_KERNEL_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.2f %s  ms_per_step %.1f  kernel_ms %.1f' % (d['value'], d['unit'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_step', 0)))"
  for n in 2048 4096; do
    bytes=$(( (8<<30) / n / 256 * 256 ))
    DM_KERNEL_VARIANT=$v timeout 300 python bench.py --blobs $n --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('streams=$n  kernel_ms=%.2f  hash_GBps=%.1f' % (r.get('kernel_ms_per_step', 0), r['achieved']/2))"
  done
done
