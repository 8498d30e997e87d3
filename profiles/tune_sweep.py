#!/usr/bin/env python
"""Time the variants of the 3-D dense sweep kernel in ONE process on one GPU (tuning runs; results copied into profiles/).

    python profiles/tune_sweep.py [--level 8] [--steps 10] [--variants old,1211,...] [--workloads sedov:hllc,smooth:hllc,...]

For every (initial condition, Riemann solver) the tree is built once; for every variant (RGPU_SWEEP, read at bind time) the level
is re-bound, K level steps are timed with CUDA events on the launching stream (rgpu_level_steps, last_steps_ms) and the state
after the timed steps is hashed: every variant must give the SAME bits (they are all bit-identical to the oracle by
construction; a variant that differs is reported as MISMATCH).  Needs a library built with -DSWEEP3_TUNING_VARIANTS.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="old,1211,1210,1212,811,821,820,1611,1610")
    ap.add_argument("--workloads", default="sedov:hllc,smooth:hllc,sedov:exact,smooth:exact,smooth:llf")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench
    from ramses_b200.hydro import HydroGPU
    from ramses_b200.tree import build_uniform_tree, fill_state
    level = args.level
    n = 1 << level
    t0 = time.time()
    a = build_uniform_tree(3, level, coarse=(1, 1, 1), myid=1, ncpu=1, order="creation", boxlen=0.5)
    print(f"# tree level {level} built in {time.time() - t0:.1f} s", flush=True)
    results = []
    ics = {}
    for wl in args.workloads.split(","):
        ic, riemann = wl.split(":")
        if ic not in ics:
            fill_state(a, level, bench.sedov_ic(0.5, 1, level) if ic == "sedov" else bench.smooth_ic((1, 1, 1)))
            ics[ic] = a.uold.copy()
        a.gamma, a.courant_factor, a.slope_type, a.riemann = 1.4, 0.8, 1, riemann
        ref_hash = None
        for v in args.variants.split(","):
            fast = v.endswith("f")                 # e.g. 1212f: the FAST arithmetic build of variant 1212
            os.environ["RGPU_SWEEP"] = v.rstrip("f")
            a.fast = fast
            a.uold[:, :] = ics[ic]
            h = HydroGPU(a, device=0)
            try:
                h.bind_level(level)
                h.upload_state(level)
                h.level_steps(level, 3)
                dts, _ = h.level_steps(level, args.steps)
                ms = h.level_info(level).last_steps_ms / args.steps
                h.set_timing(True)
                ks = []
                for _ in range(3):
                    h.level_steps(level, 1)
                    ks.append(h.level_info(level).last_sweep_ms)
                h.set_timing(False)
                h.download_state(level)
                hs = hashlib.sha1(a.uold.tobytes()).hexdigest()[:12]
                if ref_hash is None:
                    ref_hash = hs
                    ref_state = a.uold.copy()
                if fast:
                    act = a.uold[:, a.ncoarse:]
                    ref_act = ref_state[:, a.ncoarse:]
                    extra = {"max_rel_diff_vs_first": float(max(np.abs(act[k] - ref_act[k]).max() / np.abs(ref_act[k]).max() if np.abs(ref_act[k]).max() > 0 else 0.0 for k in range(5)))}
                else:
                    extra = {}
                rec = {"ic": ic, "riemann": riemann, "variant": v, "grid": n, "ms_per_step": ms, "sweep_kernel_ms": float(np.mean(ks)),
                       "cell_updates_per_s": n ** 3 / (ms * 1e-3), "state_hash": hs, "same_bits_as_first": hs == ref_hash,
                       "dt_last": float(dts[-1]), **extra}
            except Exception as e:  # a variant that cannot launch (registers / shared memory) is reported, not fatal
                rec = {"ic": ic, "riemann": riemann, "variant": v, "error": repr(e)}
            finally:
                h.finalize()
            results.append(rec)
            print(json.dumps(rec), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
