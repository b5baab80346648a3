#!/bin/bash
# round 2: the bulk dm_ingest_device path (flat index, re-ingest in place, chunked launches): the whole GPU suite on
# the new engine, then the saturation workload through dm_ingest_device with one launch and with the default chunking.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 140 python -m pytest tests -x -q -m gpu > gpurun_out/r02_bulk_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_bulk_pytest.txt
for c in 1 0; do
  name=$([ $c = 0 ] && echo default || echo one_launch)
  env $([ $c = 0 ] && echo DM_X=1 || echo DM_INGEST_CHUNKS=$c) timeout 45 python bench.py --workload saturate_151552x112KiB --steps 5 --warmup 3 --no-e2e --no-cpu --no-probes \
      > gpurun_out/r02_saturate_$name.json 2> gpurun_out/r02_saturate_$name.err
done
tail -3 gpurun_out/r02_bulk_pytest.txt
for name in one_launch default; do python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02_saturate_{sys.argv[1]}.json").read().strip().split("\n")[-1])
    print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "kernel_ms", round(d["roofline"]["kernel_ms_per_step"], 2), "launches", d["gpu_launches"])
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
done
