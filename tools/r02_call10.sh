#!/bin/bash
# round 2, 4-GPU call: BASELINE configs[3] at its stated GPU count - Ollama manifest + 7B layer blobs, digest verify routed
# by dm_shard_of over 4 B200 (value leg only: a 50 s chain per step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n4_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --workload ollama_7b_manifest --steps 1 --warmup 3 --no-e2e --no-serve --no-probes > gpurun_out/r02_bench_ollama_n4.json 2> gpurun_out/r02_bench_ollama_n4.err
