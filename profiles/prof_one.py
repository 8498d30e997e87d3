#!/usr/bin/env python
"""Minimal driver for ncu captures of the dense sweep: python profiles/prof_one.py <ic> <riemann> <level> <nsteps> [variant]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ic, riemann, level, nsteps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
if len(sys.argv) > 5:
    os.environ["RGPU_SWEEP"] = sys.argv[5]
import bench
from ramses_b200.hydro import HydroGPU
from ramses_b200.tree import build_uniform_tree, fill_state
a = build_uniform_tree(3, level, coarse=(1, 1, 1), myid=1, ncpu=1, order="creation", boxlen=0.5)
fill_state(a, level, bench.sedov_ic(0.5, 1, level) if ic == "sedov" else bench.smooth_ic((1, 1, 1)))
a.gamma, a.courant_factor, a.slope_type, a.riemann = 1.4, 0.8, 1, riemann
a.fast = os.environ.get("RGPU_FAST_MODE", "0") == "1"
h = HydroGPU(a, device=0)
h.bind_level(level)
h.upload_state(level)
dts, _ = h.level_steps(level, nsteps)
print("dt", dts[-1], "ms/step", h.level_info(level).last_steps_ms / nsteps)
h.finalize()
