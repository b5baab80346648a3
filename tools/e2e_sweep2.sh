run() { python bench.py --steps 2 --no-cpu "$@" 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('$*', '-> value %.1f e2e %.2f launches %.0f kms %.0f ringwaits %.0f' % (d['value'], e['value'], e['launches_per_step'], e['kernel_ms_sum_per_step'], e['ring_waits_per_step']))"; }
for t in 4 6 8 10 12; do run --e2e-threads $t; done
W="--blobs 4096 --blob-bytes 4194304 --e2e-concurrency 1024 --ring-mib 4096"
for t in 4 8 12; do run $W --e2e-threads $t; run $W --e2e-threads $t --e2e-zero-copy; done
