#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep into small, committed summaries under profiles/.

usage: python tools/ncu_extract.py <report.ncu-rep> <out_prefix>
Writes <out_prefix>_metrics.csv (selected raw metrics) and <out_prefix>_stalls.csv
(per-opcode warp-stall samples from the source page).
"""
import collections
import csv
import re
import subprocess
import sys

WANT = re.compile(
    r"^(gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|occupancy_limit_registers|waves_per_multiprocessor)"
    r"|sm__cycles_elapsed\.avg|smsp__cycles_active\.avg|smsp__inst_executed\.sum"
    r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|dram__bytes_(read|write)\.sum|lts__t_bytes\.sum|l1tex__t_bytes\.sum"
    r"|sm__pipe_(alu|fma|fmaheavy|fmalite)_cycles_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__inst_executed_pipe_(alu|fma|lsu|adu|uniform|xu)\.avg\.pct_of_peak_sustained_active"
    r"|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)"
    r"|smsp__warps_eligible\.avg\.per_cycle_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio"
    r"|sm__pipe_tensor.*cycles_active.*pct_of_peak_sustained_active)$")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for i, h in enumerate(hdr):
            if WANT.match(h):
                res.append((name, h, units[i], r[i]))
    return res


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by = collections.defaultdict(collections.Counter)
    insts = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        insts[op] += int(r[ix["Instructions Executed"]] or 0)
        for k in keys:
            by[op][k] += int(r[ix[k]] or 0)
    return keys, by, insts


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    with open(prefix + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        w.writerows(raw(rep))
    keys, by, insts = stalls(rep)
    with open(prefix + "_stalls.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["opcode", "warp_instructions_executed"] + keys)
        for op, _ in insts.most_common():
            w.writerow([op, insts[op]] + [by[op][k] for k in keys])


if __name__ == "__main__":
    main()
