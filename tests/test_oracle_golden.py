"""The oracle reproduces the reference's own golden sums (tests/hydro/sod-tube/sod-tube-ref.dat, tolerance 3e-13 as in
tests/visu/visu_ramses.py:497) for the 1-D AMR Sod tube: levelmin=3, levelmax=10, hllc, moncen, reflexive walls,
sub-cycling nsubcycle=3*1,2, interpol_type=2, err_grad_{d,u,p}=0.05 -- every floating-point routine of the C oracle on
the path (ctoprim, uslope, trace1d, cmpflxm, riemann_hllc, godfine1 gather + interpol_hydro + coarse refluxing, cmpdt /
courant_fine, set_unew / set_uold, upload_fine, make_boundary_hydro, condinit) is exercised by oracle/amr.py."""
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SOD = [dict(type="square", x_center=0.25, length_x=0.5, d=1.0, p=1.0),
       dict(type="square", x_center=0.75, length_x=0.5, d=0.125, p=0.1)]       # tests/hydro/sod-tube/sod-tube.nml


@pytest.fixture(scope="module")
def sod_run(orc):
    from oracle.amr import AmrRun
    r = AmrRun(1, 3, 10, (1, 1, 0, 0, 0, 0), 1.0, nsubcycle=[1, 1, 1, 2], nexpand=1, ngridmax=2000, riemann="hllc",
               slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
               interpol_type=2, interpol_var=0, regions=SOD, tout=[0.245])
    snap = r.run()
    return r, snap


def test_sod_tube_golden_sums(sod_run):
    from oracle.amr import check_sums
    r, snap = sod_run
    ref = json.load(open(os.path.join(GOLD, "sod_tube_ref.json")))
    sums = check_sums(snap["rows"], 1)
    sums["time"] = snap["t"]
    tol = 3.0e-13                                 # var_tol["all"], tests/visu/visu_ramses.py:497
    for key in ("ncells", "level", "x", "density", "pressure", "velocity_x", "time"):
        err = abs(sums[key] - ref[key]) / min(abs(sums[key]), abs(ref[key]))
        assert err <= tol, (key, sums[key], ref[key], err)


def test_sod_tube_mesh_structure_and_step_counts(sod_run):
    """doc/wiki/Start.md:158-187: initial and final mesh structure, 43 main steps, 688 fine steps of the same run."""
    r, snap = sod_run
    assert [r.initial_grids[l] for l in range(1, 9)] == [1, 2, 4, 8, 8, 8, 8, 8]
    assert [snap["grids"][l] for l in range(1, 11)] == [1, 2, 4, 8, 16, 27, 37, 17, 16, 13]
    assert snap["nstep_coarse"] == 43 and snap["nstep"] == 688
    assert abs(snap["t"] - 2.45047e-01) < 5e-7


# ---------------------------------------------------------------------------------------------------------------------
# ideal MHD: tests/mhd/imhd-tube (NDIM=1, AMR levels 5..15, hlld, slope_type=0, zero-gradient ends, interpol_type=2)
IMHD = [dict(type="square", x_center=0.75, length_x=1.5, d=1.0, u=0.0, v=0.0, w=0.0, p=1.0, A=1.0, B=1.0, C=0.0),
        dict(type="square", x_center=2.5, length_x=2.0, d=0.2, u=0.0, v=0.0, w=0.0, p=0.2, A=1.0, B=-0.989992, C=0.141120)]


@pytest.fixture(scope="module")
def imhd_run(orc):
    from oracle.amr_mhd import MhdAmrRun
    r = MhdAmrRun(5, 15, (2, 2, 0, 0, 0, 0), 3.5, nsubcycle=[1, 1, 1, 1], riemann="hlld", slope_type=0, gamma=1.6666667,
                  courant_factor=0.8, err_grad_d=0.01, err_grad_u=0.05, err_grad_p=0.05, interpol_type=2, regions=IMHD,
                  tout=[0.4], ngridmax=10000)
    return r, r.run()


def test_imhd_tube_golden_sums(imhd_run):
    """The MHD oracle (find_mhd_flux, hlld, find_speed_fast, ctoprim, trace1d, cmpflxm, godfine1 with interpol_hydro /
    interpol_mag ghost prolongation, flux reset, coarse refluxing of the 11 variables, cmpdt / courant_fine, upload_fine,
    make_boundary_hydro, hydro_refine, condinit of the MHD build) reproduces every sum of imhd-tube-ref.dat at the
    reference's tolerance of 3e-13 after 259 coarse / 16576 fine steps on an 11-level AMR hierarchy."""
    from oracle.amr_mhd import check_sums_mhd
    r, snap = imhd_run
    ref = json.load(open(os.path.join(GOLD, "imhd_tube_ref.json")))
    sums = check_sums_mhd(snap["rows"])
    sums["time"] = snap["t"]
    tol = 3.0e-13
    checked = 0
    for key, val in ref.items():
        if key not in sums:
            continue                              # boxlen, unit_*, y, z: header constants of the snapshot
        err = abs(sums[key] - val) / max(min(abs(sums[key]), abs(val)), 1e-300)
        assert err <= tol, (key, sums[key], val, err)
        checked += 1
    assert checked >= 15
    assert sums["ncells"] == 437 and snap["nstep_coarse"] == 259


# ---------------------------------------------------------------------------------------------------------------------
# 2-D hydro: tests/hydro/implosion (NDIM=2, AMR levels 5..8, hllc, moncen, four reflexive walls whose y-regions include the
# corner cells, nsubcycle=10*2, nexpand=4 (first level only), interpol_type=2, t=5: 1049 coarse / 8392 fine steps)
from conftest import IMPL, IMPL_BOUND      # the run itself is the session fixture `implosion_run` (tests/conftest.py)


def test_implosion_golden_sums(implosion_run):
    """tests/hydro/implosion/implosion-ref.dat with the reference's own tolerance (3e-13): pins the NDIM=2 paths of the
    oracle -- trace2d, 2-D slopes, cmpflxm in two directions, 2-D interpol_hydro / refluxing / flux masking, corner
    boundary octs, cmpdt -- through 8392 fine steps of a flow that amplifies any 1-ulp slip."""
    from oracle.amr import check_sums
    r, snap = implosion_run
    ref = json.load(open(os.path.join(GOLD, "implosion_ref.json")))
    sums = check_sums(snap["rows"], 2)
    sums["time"] = snap["t"]
    tol = 3.0e-13
    for key in ("ncells", "level", "dx", "x", "y", "density", "pressure", "velocity_x", "velocity_y", "time"):
        err = abs(sums[key] - ref[key]) / min(abs(sums[key]), abs(ref[key]))
        assert err <= tol, (key, sums[key], ref[key], err)
    assert snap["nstep_coarse"] == 1049 and snap["nstep"] == 8392
    assert [snap["grids"][l] for l in range(1, 9)] == [1, 4, 16, 64, 256, 781, 1288, 2706]


def test_fast_amr_driver_equals_python_driver(orc):
    """oracle/ramses_oracle_amr.c (flag / scan passes in C) against the pure-Python passes of AmrRun: identical state,
    tree and time on the sod-tube golden case."""
    import numpy as np
    from oracle.amr import AmrRun, FastAmrRun
    out = []
    for cls in (AmrRun, FastAmrRun):
        r = cls(1, 3, 10, (1, 1, 0, 0, 0, 0), 1.0, nsubcycle=[1, 1, 1, 2], nexpand=1, ngridmax=2000, riemann="hllc",
                slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
                interpol_type=2, interpol_var=0, regions=SOD, tout=[0.245])
        snap = r.run()
        out.append((snap["t"], snap["rows"], r.uold.copy(), r.son.copy(), r.nbor.copy()))
    a, b = out
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


def test_amr_flux_threads_do_not_change_results(orc):
    """orc_set_amr_threads: the flux phase of a group of batches runs concurrently, every update of unew (incl. the coarse
    refluxes) stays serial and in the reference's order -> bit-identical to the serial routine."""
    import numpy as np
    from oracle.amr import FastAmrRun
    out = []
    for nt in (1, 4):
        r = FastAmrRun(2, 5, 8, (1, 1, 1, 1, 0, 0), 1.0, nsubcycle=[2] * 10, nexpand=[4], ngridmax=20000, riemann="hllc",
                       slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
                       interpol_type=2, interpol_var=0, regions=IMPL, tout=[0.0, 0.03], bound_regions=IMPL_BOUND, nthreads=nt)
        snap = r.run()
        out.append((snap["t"], snap["nstep"], r.uold.copy(), r.son.copy()))
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3])


# ---------------------------------------------------------------------------------------------------------------------
# 2-D ideal MHD: tests/mhd/orszag-tang (NDIM=2, AMR levels 5..9, hlld / hlld, moncen, periodic, nsubcycle=1*1,2, err_grad_p=0.1,
# interpol_type=2, the patch's condinit.f90, t=0.5: 174 coarse / 1236 fine steps, 100 066 leaf cells)
def test_orszag_tang_golden_sums(orszag_run):
    """tests/mhd/orszag-tang/orszag-tang-ref.dat at the reference's tolerance (3e-13; we get <= 2e-15 on every sum): pins the
    NDIM=2 MHD paths -- trace2d, the hlld 1-D solver and the hlld 2-D (corner EMF) solver of cmp_mag_flx, the constrained-transport
    update, divergence-free prolongation (interpol_mag), EMF refluxing at coarse-fine edges, face-centred restriction, cmpdt,
    hydro_refine -- and, through test_oracle_mhd.py::test_unsplit_2d_equals_z_invariant_3d, the x/y/E_z paths of the NDIM=3
    restatement the GPU kernels are compared with."""
    from oracle.amr_mhd import check_sums_cols
    r, snap = orszag_run
    ref = json.load(open(os.path.join(GOLD, "orszag_tang_ref.json")))
    sums = check_sums_cols(snap["rows"])
    sums["time"] = snap["t"]
    tol = 3.0e-13
    for key in ("ncells", "level", "dx", "x", "y", "z", "density", "pressure", "velocity_x", "velocity_y", "velocity_z",
                "B_x_left", "B_x_right", "B_y_left", "B_y_right", "B_z_left", "B_z_right", "time"):
        den = min(abs(sums[key]), abs(ref[key]))
        err = abs(sums[key] - ref[key]) / den if den > 0 else abs(sums[key] - ref[key])
        assert err <= tol, (key, sums[key], ref[key], err)
    assert snap["nstep_coarse"] == 174 and snap["nstep"] == 1236
    assert [snap["grids"][l] for l in range(5, 10)] == [256, 1024, 4018, 11252, 16720]
    assert r.divb_max() < 2e-14                      # |sum of face differences| per leaf cell: div B = 0 to round-off
