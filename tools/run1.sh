mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 1 2; do
  echo "== variant $v deep"; DM_KERNEL_VARIANT=$v python bench.py --blobs 256 --blob-bytes 16777216 --steps 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['achieved'])"
  echo "== variant $v wide"; DM_KERNEL_VARIANT=$v python bench.py --workload saturate_151552x112KiB --steps 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['achieved'])"
done
echo "== e2e default"; python bench.py --steps 3 --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; python -c "import json; d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e'])"; tail -3 gpurun_out/bench_e2e.err
echo "== e2e zero-copy"; python bench.py --steps 3 --no-cpu --e2e-zero-copy > gpurun_out/bench_e2e_zc.json 2>&1; python -c "import json; d=json.loads(open('gpurun_out/bench_e2e_zc.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e'])"
