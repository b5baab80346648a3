# BASELINE configs[4] in miniature on one GPU: 8 GiB per size class, blob size x4 per row, kernel chosen by the engine.
for spec in "8192 1048576" "2048 4194304" "512 16777216" "128 67108864" "32 268435456"; do
  set -- $spec
  timeout 300 python bench.py --blobs $1 --blob-bytes $2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('blob %4d MiB x %5d  kernel=%-26s value %7.1f GB/s  kernel-only %7.1f GB/s hashed  frac_hbm(2B/B) %.4f' % ($2>>20, $1, d['config']['kernel'], d['value'], r['achieved']/2, r['frac']))"
done
