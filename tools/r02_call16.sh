#!/bin/bash
# round 2: BASELINE configs[1] and [4] (scaled) again on the final build (two-warp deep kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --workload llama3_8b_shards --steps 1 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_llama3_8b_shards.json 2> gpurun_out/r02_bench_llama3.err
timeout 600 python bench.py --workload size_sweep --steps 2 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_size_sweep.json 2> gpurun_out/r02_bench_size_sweep.err
