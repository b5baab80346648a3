#!/bin/bash
# round 2, 2-GPU call: the in-process multi-GPU routing test, the weak-scaling line with the digest-routed fixed-set probe,
# and BASELINE configs[3] (Ollama manifest + 7B layers) routed by dm_shard_of over the ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "engines_on_every_gpu" -rA > gpurun_out/r02_multigpu_test.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload ollama_7b_manifest --steps 1 --warmup 3 --no-serve --no-probes > gpurun_out/r02_bench_ollama_n2.json 2> gpurun_out/r02_bench_ollama_n2.err
