#!/usr/bin/env python
"""Record the single-GPU state / dt hashes bench.py's in-run correctness check compares with (tests/golden/bench_hashes.json).

    python profiles/make_bench_golden.py <workload> <total_steps>[,<total_steps>...]

Runs the workload on ONE GPU exactly like bench.py (same tree, same replicated initial condition) and hashes the state after
each requested number of level steps (warmup + steps of a bench run) with bench.canonical_check.  The state after n steps does
not depend on how the steps are chunked into rgpu_level_steps calls, nor on the number of GPUs: an N-GPU bench run must
reproduce these hashes on every rank."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ramses_b200.hydro import HydroGPU  # noqa: E402
from ramses_b200.tree import build_uniform_tree, fill_state  # noqa: E402

workload = sys.argv[1]
totals = sorted(int(t) for t in sys.argv[2].split(","))
w = bench.WORKLOADS[workload]
level = w["level"]
a = build_uniform_tree(3, level, coarse=(1, 1, 1), myid=1, ncpu=1, order="lattice", boxlen=0.5)
a.gamma, a.courant_factor, a.slope_type, a.riemann = bench.GAMMA, 0.8, w["slope_type"], w["riemann"]
base = bench.sedov_ic(0.5, 1, level) if w["ic"] == "sedov" else bench.smooth_ic((1, 1, 1))
fill_state(a, level, lambda x, y, z: base(np.mod(x, 1.0), np.mod(y, 1.0), np.mod(z, 1.0)))
h = HydroGPU(a, device=0)
h.bind_level(level)
h.upload_state(level)
gold = json.load(open(bench.GOLDEN_HASHES)) if os.path.exists(bench.GOLDEN_HASHES) else {}
done, hist = 0, []
for t in totals:
    dts, _ = h.level_steps(level, t - done)
    hist.append(dts)
    done = t
    h.download_state(level)
    c = bench.canonical_check(a, level, (1, 1, 1), 0, np.concatenate(hist))
    gold[f"{workload}:{t}"] = c
    print(workload, t, c["state_sha1"], c["dt_sha1"], flush=True)
h.finalize()
json.dump(gold, open(bench.GOLDEN_HASHES, "w"), indent=1, sort_keys=True)
