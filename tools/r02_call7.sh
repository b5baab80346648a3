#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests.txt
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
for t in 4 16 48 96; do
  echo "== e2e threads $t"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-probes --no-serve --e2e-threads $t 2>&1 | tail -1
done > gpurun_out/r02_e2e_sweep.txt 2>&1
timeout 300 python tools/tiny_blob_probe.py > gpurun_out/r02_tiny_blob.txt 2>&1
timeout 600 python bench.py --workload size_sweep --steps 2 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_size_sweep.json 2> gpurun_out/r02_bench_size_sweep.err
timeout 900 python bench.py --workload ollama_7b_manifest --steps 1 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_ollama_n1.json 2> gpurun_out/r02_bench_ollama_n1.err
bash tools/prof_r02.sh > gpurun_out/r02_prof.log 2>&1
