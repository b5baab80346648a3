#!/bin/bash
# round 2: the two-warp deep kernel (variant 8: schedule warp ahead of the round warp) against variant 7
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 21,7 21,8 21,7 21,8; do timeout 200 python tools/check_variant.py $v | tail -12; done > gpurun_out/r02_deep2.txt 2>&1
DM_KERNEL_VARIANT=21,8 timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fips or every_length or boundaries" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" | tail -5 > gpurun_out/r02_deep2_racecheck.txt 2>&1
DM_KERNEL_VARIANT=21,8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_gputests_deep8.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputests_deep8.txt
for v in 21,7 21,8; do
  DM_KERNEL_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-probes --no-serve 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v value %.2f ms/step %.1f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done > gpurun_out/r02_deep2_bench.txt 2>&1
