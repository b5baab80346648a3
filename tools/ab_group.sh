# kernel choice by stream count: time each streams-per-warp setting on N blobs of equal size (hash-and-cache)
for n in ${COUNTS:-256 512 1024 2048 4096 8192 16384 32768}; do
  bytes=$(( (8<<30) / n / 256 * 256 ))
  for spw in 1 2 4 8 16 32; do
    DM_FORCE_SPW=$spw python bench.py --blobs $n --blob-bytes $bytes --steps 2 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('streams=%6d spw=%2d kernel_ms=%9.2f hash_GBps=%7.1f' % ($n, $spw, r['kernel_ms_per_step'], r['achieved']/2))"
  done
done
