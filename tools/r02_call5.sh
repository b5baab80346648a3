#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=64 timeout 120 python tools/dbg_arena.py > gpurun_out/r02_dbg_arena.txt 2>&1
N=256 timeout 120 python tools/dbg_arena.py >> gpurun_out/r02_dbg_arena.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests.txt
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
