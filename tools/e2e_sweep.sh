python tools/h2d_probe.py
run() { python bench.py --steps 2 --no-cpu "$@" 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('$*', '-> value %.1f e2e %.2f launches %.0f kms %.0f ringwaits %.0f' % (d['value'], e['value'], e['launches_per_step'], e['kernel_ms_sum_per_step'], e['ring_waits_per_step']))"; }
run
run --ring-mib 4096
run --slab-kib 256
run --e2e-threads 8
run --e2e-threads 32
run --blobs 4096 --blob-bytes 4194304 --e2e-concurrency 1024 --ring-mib 4096
run --blobs 4096 --blob-bytes 4194304 --e2e-concurrency 1024 --ring-mib 4096 --e2e-zero-copy
run --blobs 16384 --blob-bytes 1048576 --e2e-concurrency 4096 --ring-mib 4096 --slab-kib 256
