#!/bin/bash
# round 2: wide style 6 (one copy of the compression code per loop, 78 registers) against the shipped style 5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/check_variant.py 25,7 | tail -10 > gpurun_out/r02_wide_style6_parity.txt 2>&1
VARIANTS="21 25 21 25" EXTRA="--no-probes" timeout 600 bash tools/ab_wide.sh > gpurun_out/r02_ab_wide6.txt 2>&1
DM_KERNEL_VARIANT=25,7 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_gputests_wide25.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests_wide25.txt
