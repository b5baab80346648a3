#!/bin/bash
# Round-2 ncu passes on the exact default bench command (B200_PROFILING.md recipe).  Run under gpurun; outputs under
# gpurun_out/, summaries are copied into profiles/ by hand (tools/ncu_extract.py for the .ncu-rep files).
#   1. launch list with device time per launch (cold-cache, serialised: compare SHARES, not absolutes)
#   2. DRAM bytes of ONE launch of the dominant kernel at the full workload -> roofline_traffic.json entry
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-serve --no-probes > gpurun_out/r02_launches_default.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sha256_deep -s 4 -c 1 --csv \
    --log-file gpurun_out/r02_dram_default.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-probes > gpurun_out/r02_dram_default.log 2>&1
python - <<'PY'
import csv, json, re
rows = [r for r in csv.reader(open("gpurun_out/r02_dram_default.csv")) if len(r) > 10]
hdr = next(r for r in rows if "Metric Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
vals = {}
for r in rows:
    if r is hdr or len(r) <= ix["Metric Value"]:
        continue
    name = r[ix["Metric Name"]]
    if name.startswith("dram__bytes"):
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]].lower()
        vals[name] = v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        kernel = r[ix["Kernel Name"]]
import sys
sys.path.insert(0, ".")
import demodel_b200
lib = demodel_b200.load()
variant = f"{lib.dm_default_kernel_variant(0)},{lib.dm_default_kernel_variant(1)}"
tot = int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"])
json.dump({"hf_lfs_256x64MiB": {"bytes": tot, "kernel_variant": variant, "family": "wide" if "wide" in kernel else "deep", "kernel": re.sub(r"\(.*", "", kernel),
           "source": "profiles/r02_dram_default.csv: dram__bytes_read.sum + dram__bytes_write.sum of one launch at the full default "
                     f"workload ({int(vals['dram__bytes_read.sum'])} + {int(vals['dram__bytes_write.sum'])} B); algorithmic 2 x 17179869184 B"}},
          open("gpurun_out/roofline_traffic.json", "w"), indent=1)
print(open("gpurun_out/roofline_traffic.json").read())
PY
