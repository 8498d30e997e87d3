"""The reference's own golden files with `godunov_fine` running on the GPU.

The drop-in boundary of this repo is `call godunov_fine(ilevel)` (amr/amr_step.f90:388): everything else of the reference stays on
the host.  Here the "host code" is the oracle's AMR driver (oracle/amr.py, oracle/amr_mhd.py: flagging, refinement, time-step
control, set_unew / set_uold, upload_fine, boundaries -- the parts of RAMSES this repo does not replace) and ONLY its
`c_godunov_fine` is swapped for the library's Level-0 call `rgpu_godunov_fine` (host arrays in and out, tree and communicator
lists re-bound after every regrid exactly like the Fortran shim of INTEGRATION.md does after build_comm).  The four golden files
the reference holds for this path must still come out within the reference's own tolerance (tests/visu/visu_ramses.py:497,
3e-13):  sod-tube (1-D AMR hydro), implosion (2-D AMR hydro), imhd-tube (1-D AMR MHD), orszag-tang (2-D AMR MHD)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 3.0e-13


def attach_gpu_godunov(r, mhd=False, riemann="hllc", riemann2d="llf", slope_type=1, interpol_type=2, interpol_var=0):
    """replace r.c_godunov_fine by the GPU drop-in; returns (HydroGPU, counters)"""
    from ramses_b200.hydro import AmrCommons, HydroGPU
    m = r.m
    nvar = 8 if mhd else r.nvar
    a = AmrCommons(r.ndim, nvar, m.ncoarse, m.ngridmax, m.nx, m.ny, m.nz, (m.icoarse_min, m.icoarse_max), (m.jcoarse_min, m.jcoarse_max),
                   (m.kcoarse_min, m.kcoarse_max), nlevelmax=r.nlevelmax, boxlen=r.p.boxlen, mhd=mhd)
    # the caller's arrays, not copies: son(1:ncell), father(1:ngridmax), uold/unew(1:ncell,1:nvar[+3]) are views of the driver's
    a.son, a.father = r.son[1:], r.father[1:]
    a.uold, a.unew = r.uold.reshape(-1, m.ncell), r.unew.reshape(-1, m.ncell)
    a.nbor[:, :] = r.nbor[:, 1:]
    a.boundary_type = [m.boundary_type[b] for b in range(m.nboundary)]
    pp = r.pm if mhd else r.p
    a.gamma, a.courant_factor = pp.gamma, pp.courant_factor
    a.slope_type, a.riemann, a.nvector = slope_type, riemann, r.nvector
    if mhd:
        a.slope_mag_type, a.riemann2d = -1, riemann2d
    h = HydroGPU(a, amr_mode=True, interpol_type=interpol_type, interpol_var=interpol_var)
    st = dict(dirty=True, sig={}, calls=0, binds=0, syncs=0)

    def mark(fn):
        def w(*args, **kw):
            st["dirty"] = True
            return fn(*args, **kw)
        return w
    for name in ("make_grid_fine", "kill_grid", "make_grid_coarse"):
        setattr(r, name, mark(getattr(r, name)))

    def sync():
        a.nbor[:, :] = r.nbor[:, 1:]
        h.bind_tree()
        for l in range(1, r.nlevelmax + 1):
            act = np.array(r.active[l], dtype=np.int32)
            bnd = [np.array(r.bound[b][l], dtype=np.int32) for b in range(m.nboundary)]
            a.active[l], a.boundary[l] = act, bnd
            if len(act) == 0:
                continue
            sig = (act.tobytes(), a.father[act - 1].tobytes(), a.nbor[:, act - 1].tobytes(), tuple(x.tobytes() for x in bnd))
            if st["sig"].get(l) != sig:
                h.bind_level(l)
                st["sig"][l] = sig
                st["binds"] += 1
        st["dirty"] = False
        st["syncs"] += 1

    def c_godunov_fine(l):
        if len(r.active[l]) == 0:
            return
        if st["dirty"]:
            sync()
        a.dtnew[l] = r.dtnew[l]
        h.godunov_fine(l)
        st["calls"] += 1
    r.c_godunov_fine = c_godunov_fine
    return h, st


def _check(sums, ref, keys):
    for key in keys:
        err = abs(sums[key] - ref[key]) / max(min(abs(sums[key]), abs(ref[key])), 1e-300)
        assert err <= TOL, (key, sums[key], ref[key], err)


def test_sod_tube_golden_with_gpu_godunov_fine(orc):
    """tests/hydro/sod-tube: 43 coarse / 688 fine steps, levels 3..10"""
    from oracle.amr import AmrRun, check_sums
    from test_oracle_golden import SOD
    r = AmrRun(1, 3, 10, (1, 1, 0, 0, 0, 0), 1.0, nsubcycle=[1, 1, 1, 2], nexpand=1, ngridmax=2000, riemann="hllc",
               slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
               interpol_type=2, interpol_var=0, regions=SOD, tout=[0.245])
    h, st = attach_gpu_godunov(r, riemann="hllc", slope_type=2)
    snap = r.run()
    h.finalize()
    ref = json.load(open(os.path.join(GOLD, "sod_tube_ref.json")))
    sums = check_sums(snap["rows"], 1)
    sums["time"] = snap["t"]
    _check(sums, ref, ("ncells", "level", "x", "density", "pressure", "velocity_x", "time"))
    assert snap["nstep_coarse"] == 43 and snap["nstep"] == 688
    assert st["calls"] >= 688


def test_imhd_tube_golden_with_gpu_godunov_fine(orc):
    """tests/mhd/imhd-tube: 259 coarse / 16576 fine steps, levels 5..15, hlld, slope_type=0"""
    from oracle.amr_mhd import MhdAmrRun, check_sums_mhd
    from test_oracle_golden import IMHD
    r = MhdAmrRun(5, 15, (2, 2, 0, 0, 0, 0), 3.5, nsubcycle=[1, 1, 1, 1], riemann="hlld", slope_type=0, gamma=1.6666667,
                  courant_factor=0.8, err_grad_d=0.01, err_grad_u=0.05, err_grad_p=0.05, interpol_type=2, regions=IMHD,
                  tout=[0.4], ngridmax=10000)
    h, st = attach_gpu_godunov(r, mhd=True, riemann="hlld", slope_type=0)
    snap = r.run()
    h.finalize()
    ref = json.load(open(os.path.join(GOLD, "imhd_tube_ref.json")))
    sums = check_sums_mhd(snap["rows"])
    sums["time"] = snap["t"]
    _check(sums, ref, [k for k in ref if k in sums])
    assert sums["ncells"] == 437 and snap["nstep_coarse"] == 259
    assert st["calls"] > 10000


def test_implosion_golden_with_gpu_godunov_fine(orc):
    """tests/hydro/implosion: 1049 coarse / 8392 fine steps, levels 5..8, four reflexive walls"""
    from conftest import IMPL, IMPL_BOUND
    from oracle.amr import FastAmrRun, check_sums
    # ngridmax: the run peaks at 6 739 octs (incl. boundary octs); a small array keeps the per-call host<->device traffic small
    r = FastAmrRun(2, 5, 8, (1, 1, 1, 1, 0, 0), 1.0, nsubcycle=[2] * 10, nexpand=[4], ngridmax=12000, riemann="hllc",
                   slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
                   interpol_type=2, interpol_var=0, regions=IMPL, tout=[0.0, 5.0], bound_regions=IMPL_BOUND)
    h, st = attach_gpu_godunov(r, riemann="hllc", slope_type=2)
    snap = r.run()
    h.finalize()
    ref = json.load(open(os.path.join(GOLD, "implosion_ref.json")))
    sums = check_sums(snap["rows"], 2)
    sums["time"] = snap["t"]
    _check(sums, ref, ("ncells", "level", "dx", "x", "y", "density", "pressure", "velocity_x", "velocity_y", "time"))
    assert snap["nstep_coarse"] == 1049 and snap["nstep"] == 8392
    assert [snap["grids"][l] for l in range(1, 9)] == [1, 4, 16, 64, 256, 781, 1288, 2706]


def test_orszag_tang_golden_with_gpu_godunov_fine(orc):
    """tests/mhd/orszag-tang: levels 5..9, hlld / hlld, moncen, periodic"""
    from oracle.amr_mhd import MhdAmrRun2D, check_sums_cols
    r = MhdAmrRun2D(5, 9, 1.0, nsubcycle=[1], riemann="hlld", riemann2d="hlld", slope_type=2, gamma=1.6666667,
                    courant_factor=0.8, err_grad_p=0.1, interpol_type=2, tout=[0.5], nexpand=1, ngridmax=40000)   # peaks at 33 585 octs
    h, st = attach_gpu_godunov(r, mhd=True, riemann="hlld", riemann2d="hlld", slope_type=2)
    snap = r.run()
    h.finalize()
    ref = json.load(open(os.path.join(GOLD, "orszag_tang_ref.json")))
    sums = check_sums_cols(snap["rows"])
    sums["time"] = snap["t"]
    for key in ("ncells", "level", "dx", "x", "y", "z", "density", "pressure", "velocity_x", "velocity_y", "velocity_z",
                "B_x_left", "B_x_right", "B_y_left", "B_y_right", "B_z_left", "B_z_right", "time"):
        den = min(abs(sums[key]), abs(ref[key]))
        err = abs(sums[key] - ref[key]) / den if den > 0 else abs(sums[key] - ref[key])
        assert err <= TOL, (key, sums[key], ref[key], err)
    assert snap["nstep_coarse"] == 174 and snap["nstep"] == 1236
    assert [snap["grids"][l] for l in range(5, 10)] == [256, 1024, 4018, 11252, 16720]
    assert r.divb_max() < 5e-14
