#!/usr/bin/env python
"""Reduce an `ncu --set full` capture of the sweep kernel to the counters bench.py reports (profiles/kernel_counters.json):

    python profiles/ncu_counters.py <report.ncu-rep> <cells per launch> <workload key> [fast]

fp64_thread_instr_per_cell = (DADD + DMUL + DFMA thread instructions) / cells; dram_bytes_per_launch = dram read + write."""
import csv
import io
import json
import os
import subprocess
import sys

rep, cells, key = sys.argv[1], float(sys.argv[2]), sys.argv[3]
fast = len(sys.argv) > 4 and sys.argv[4] == "fast"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))
f = lambda k: float(d[k].replace(",", ""))
cyc = f("sm__cycles_elapsed.max")
ops = sum(f(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") for o in ("dadd", "dmul", "dfma")) * cyc
def bytes_of(k):
    v, unit = f(k), u[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
rec = {"fp64_thread_instr_per_cell": ops / cells, "fp64_peak_thread_instr_per_s": 16.7e12,
       "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
       "kernel_ms_under_ncu": f("gpu__time_duration.sum") * {"msecond": 1, "ms": 1, "usecond": 1e-3, "us": 1e-3, "second": 1e3, "s": 1e3}[u["gpu__time_duration.sum"].lower()],
       "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "registers": f("launch__registers_per_thread"), "source": os.path.basename(rep)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_counters.json")
allc = json.load(open(path)) if os.path.exists(path) else {}
if fast:
    allc.setdefault(key, {})["fast"] = rec
else:
    keep = allc.get(key, {}).get("fast")
    allc[key] = rec
    if keep:
        allc[key]["fast"] = keep
json.dump(allc, open(path, "w"), indent=1)
print(key, "fast" if fast else "strict", json.dumps(rec))
