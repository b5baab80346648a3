#!/bin/bash
# round 2: two-warp group kernels (deep variant 9) against the one-warp ones (8), per stream count and streams per warp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/check_variant.py 21,9 | tail -12 > gpurun_out/r02_group2_parity.txt 2>&1
run() {  # variant n spw
  bytes=$(( (8<<30) / $2 / 256 * 256 ))
  DM_KERNEL_VARIANT=$1 DM_FORCE_SPW=$3 timeout 120 python bench.py --blobs $2 --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu --no-probes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('variant $1 streams=%6d spw=%2d kernel_ms=%9.2f hash_GBps=%7.1f' % ($2, $3, r['kernel_ms_per_step'], r['achieved']/2))"
}
{
run 21,8 1024 2; run 21,9 1024 2; run 21,9 1024 4
run 21,8 2048 4; run 21,9 2048 4; run 21,9 2048 8
run 21,8 4096 16; run 21,9 4096 8; run 21,9 4096 16
run 21,8 8192 16; run 21,9 8192 16
run 21,8 512 1; run 21,10 512 1; run 21,9 512 2
run 21,8 592 1; run 21,10 592 1; run 21,9 592 2
} > gpurun_out/r02_group2_sweep.txt 2>&1
