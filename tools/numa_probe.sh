W="--blobs 4096 --blob-bytes 4194304 --e2e-concurrency 1024 --ring-mib 4096 --e2e-threads 12 --steps 2 --no-cpu --no-serve --no-probes"
for extra in "" "--numa-bind" "" "--numa-bind"; do
  timeout 200 python bench.py $W $extra 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[$extra] e2e %.2f GB/s  numa=%s' % (d['e2e']['value'], d['host']['numa_bind']))"
done
