#!/bin/bash
# round 2, 8-GPU call: BASELINE configs[4] at its stated GPU count, scaled and saying so (classes 1 MiB .. 1 GiB, 12 GiB per GPU
# per step instead of 125 GB; the 64 GiB top class is one 15-minute chain on any number of GPUs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n8_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --workload size_sweep --steps 2 --warmup 3 --no-e2e --no-serve --no-probes > gpurun_out/r02_bench_size_sweep_n8.json 2> gpurun_out/r02_bench_size_sweep_n8.err
