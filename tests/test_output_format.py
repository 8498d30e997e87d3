"""ramses_b200/output.py writes the reference's snapshot format; the files are read back with the REFERENCE's own reader and
checked with the REFERENCE's own check_solution (tests/visu/visu_ramses.py) against the golden sums: state -> reference file
format -> reference reader -> reference checker.  Needs the reference tree (only present in the build container): skipped on
machines without it."""
import json
import os
import sys

import numpy as np
import pytest

REF_VISU = "/root/reference/tests/visu"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_VISU, "visu_ramses.py")),
                                reason="reference tree not available")


def _write_from_run(r, tmp, iout, mhd=False):
    """the oracle AMR driver's tree / state (1-based arrays with an unused element 0) -> output_NNNNN/"""
    from ramses_b200.output import write_snapshot
    m = r.m
    nb = m.nboundary
    L = r.nlevelmax
    nvs = r.nvar
    return write_snapshot(
        str(tmp), iout, ndim=r.ndim, nvar=8 if mhd else r.nvar, levelmin=r.levelmin, nlevelmax=L, ngridmax=r.ngridmax,
        ncoarse=r.ncoarse, nxyz=(m.nx, m.ny, m.nz), coarse_min=(m.icoarse_min, m.jcoarse_min, m.kcoarse_min),
        coarse_max=(m.icoarse_max, m.jcoarse_max, m.kcoarse_max), boxlen=r.p.boxlen, gamma=(r.pm.gamma if mhd else r.p.gamma),
        smallr=r.p.smallr, son=r.son[1:], father=r.father[1:], nbor=r.nbor[:, 1:], xg=r.xg[:, 1:],
        active=[r.active[l] for l in range(1, L + 1)], boundary=[[r.bound[b][l] for l in range(1, L + 1)] for b in range(nb)],
        uold=r.uold.reshape(nvs, r.ncell), t=r.t, dtold=[r.dtold[l] for l in range(1, L + 1)],
        dtnew=[r.dtnew[l] for l in range(1, L + 1)], nstep=r.nstep, nstep_coarse=r.nstep_coarse, tout=r.tout, mhd=mhd)


def _check_with_reference(tmp, iout, test_name, ref_json):
    sys.path.insert(0, REF_VISU)
    try:
        import visu_ramses
    finally:
        sys.path.remove(REF_VISU)
    ref = json.load(open(os.path.join(GOLD, ref_json)))
    cwd = os.getcwd()
    os.chdir(str(tmp))
    try:
        with open(test_name + "-ref.dat", "w") as f:
            for k in sorted(ref):
                f.write("%s : %.16e\n" % (k, ref[k]))
        data = visu_ramses.load_snapshot(iout)
        visu_ramses.check_solution(data["data"], test_name)       # what tests/*/plot-*.py end with; prints PASSED
    finally:
        os.chdir(cwd)
    return data


def test_sod_tube_snapshot_through_reference_reader(orc, tmp_path, capsys):
    from oracle.amr import AmrRun
    from test_oracle_golden import SOD
    r = AmrRun(1, 3, 10, (1, 1, 0, 0, 0, 0), 1.0, nsubcycle=[1, 1, 1, 2], nexpand=1, ngridmax=2000, riemann="hllc",
               slope_type=2, gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05,
               interpol_type=2, interpol_var=0, regions=SOD, tout=[0.245])
    snap = r.run()
    # the driver stops after the output step: the state at the output time is the current one
    _write_from_run(r, tmp_path, 2)
    data = _check_with_reference(tmp_path, 2, "sod-tube", "sod_tube_ref.json")
    assert "PASSED" in capsys.readouterr().out              # the reference's own verdict on the reference's own golden file
    assert data["data"]["ncells"] == 142 and abs(data["data"]["time"] - snap["t"]) < 1e-14
    rows = snap["rows"]
    x = np.array([q[1][0] for q in rows])
    order_ref, order_ours = np.argsort(data["data"]["x"]), np.argsort(x)
    assert np.array_equal(np.sort(data["data"]["x"]), np.sort(x))
    assert np.array_equal(data["data"]["density"][order_ref], np.array([q[2] for q in rows])[order_ours])
    assert np.array_equal(data["data"]["pressure"][order_ref], np.array([q[4] for q in rows])[order_ours])


def test_orszag_tang_snapshot_through_reference_reader(orc, tmp_path):
    """a short NDIM=2 MHD AMR run: the eleven output fields survive the file format bit for bit"""
    from oracle.amr_mhd import MhdAmrRun2D
    r = MhdAmrRun2D(4, 6, 1.0, nsubcycle=[1], riemann="hlld", riemann2d="hlld", slope_type=2, gamma=1.6666667, courant_factor=0.8,
                    err_grad_p=0.1, interpol_type=2, tout=[0.1], nexpand=1, ngridmax=20000)
    snap = r.run()
    _write_from_run(r, tmp_path, 2, mhd=True)
    sys.path.insert(0, REF_VISU)
    try:
        import visu_ramses
    finally:
        sys.path.remove(REF_VISU)
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        data = visu_ramses.load_snapshot(2)
    finally:
        os.chdir(cwd)
    ours = snap["rows"]
    assert data["data"]["ncells"] == len(ours["level"])
    key_ref = np.lexsort((data["data"]["y"], data["data"]["x"]))
    key_our = np.lexsort((ours["y"], ours["x"]))
    for k in ("level", "x", "y", "dx", "density", "velocity_x", "velocity_y", "velocity_z", "pressure", "B_x_left", "B_y_left",
              "B_z_left", "B_x_right", "B_y_right", "B_z_right"):
        assert np.array_equal(np.asarray(data["data"][k])[key_ref], np.asarray(ours[k])[key_our]), k


def test_implosion_and_orszag_tang_through_reference_checker(implosion_run, orszag_run, tmp_path, capsys):
    """the two long golden runs (session fixtures shared with test_oracle_golden.py): write the final state in the reference's
    format and let the REFERENCE's check_solution compare with the REFERENCE's golden file -- it prints PASSED for orszag-tang
    (all sums to 2e-15) and for implosion (within its 3e-13 tolerance)."""
    r, _ = orszag_run
    d1 = tmp_path / "ot"
    d1.mkdir()
    _write_from_run(r, d1, 2, mhd=True)
    _check_with_reference(d1, 2, "orszag-tang", "orszag_tang_ref.json")
    assert "PASSED" in capsys.readouterr().out
    r, _ = implosion_run
    d2 = tmp_path / "impl"
    d2.mkdir()
    _write_from_run(r, d2, 2)
    _check_with_reference(d2, 2, "implosion", "implosion_ref.json")
    assert "PASSED" in capsys.readouterr().out


def test_snapshot_from_host_mirror_nested_tree(tmp_path):
    """the product's host mirror (AmrCommons from ramses_b200.tree.build_nested_tree, three levels) -> snapshot -> reference
    reader: every leaf cell comes back at its position with its value."""
    from ramses_b200.output import snapshot_from_commons
    from ramses_b200.tree import build_nested_tree, fill_state
    a = build_nested_tree(3, 5, half_width=2)
    a.gamma = 1.4

    def fn(x, y, z):
        u = np.zeros((5, len(x)))
        u[0] = 1.0 + x + 2 * y + 4 * z
        u[1], u[2], u[3] = 0.1 * u[0], -0.2 * u[0], 0.3 * u[0]
        u[4] = 2.5 + 0.5 * u[0] * (0.01 + 0.04 + 0.09) + x * y
        return u
    for l in range(1, 6):
        fill_state(a, l, fn)
    snapshot_from_commons(a, str(tmp_path), 1, t=0.125, levelmin=3)
    sys.path.insert(0, REF_VISU)
    try:
        import visu_ramses
    finally:
        sys.path.remove(REF_VISU)
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        data = visu_ramses.load_snapshot(1)["data"]
    finally:
        os.chdir(cwd)
    nleaf = sum(int((a.son[a.ncoarse + ind * a.ngridmax + a.active[l].astype(np.int64) - 1] == 0).sum())
                for l in range(1, 6) for ind in range(8))
    assert data["ncells"] == nleaf and data["time"] == 0.125
    x, y, z = data["x"], data["y"], data["z"]
    assert np.array_equal(data["density"], 1.0 + x + 2 * y + 4 * z)
    assert np.allclose(data["velocity_y"], -0.2, rtol=0, atol=1e-16)
    assert np.allclose(data["pressure"], 0.4 * (2.5 + x * y), rtol=1e-14, atol=0)
    assert set(np.unique(data["level"])) == {3.0, 4.0, 5.0}


def test_snapshot_restart_round_trip(orc, tmp_path):
    """write_snapshot -> read_snapshot (the restart side, amr/init_amr.f90 + hydro/init_hydro.f90:57-250): the tree arrays and the
    lists come back identically, the conserved state to round-off (the file holds primitive variables, like the reference's own
    restart), and a second write of the restored state reproduces the amr file byte for byte."""
    from oracle.amr import AmrRun
    from ramses_b200.output import read_snapshot, write_snapshot
    from test_oracle_golden import SOD
    r = AmrRun(1, 3, 8, (1, 1, 0, 0, 0, 0), 1.0, nsubcycle=[1, 1, 1, 2], nexpand=1, ngridmax=500, riemann="hllc", slope_type=2,
               gamma=1.4, courant_factor=0.8, err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05, interpol_type=2, interpol_var=0,
               regions=SOD, tout=[0.1])
    r.run()
    d1 = tmp_path / "a"
    d1.mkdir()
    _write_from_run(r, d1, 3)
    s = read_snapshot(str(d1), 3, smallr=r.p.smallr)
    assert s["t"] == r.t and s["nstep"] == r.nstep and s["ndim"] == 1 and s["nboundary"] == 2
    used = np.concatenate([np.asarray(r.active[l], dtype=np.int64) for l in range(1, 9)] +
                          [np.asarray(r.bound[b][l], dtype=np.int64) for b in range(2) for l in range(1, 9)])
    assert np.array_equal(s["son"][:r.ncoarse], r.son[1:r.ncoarse + 1])
    for ind in range(2):
        c = r.ncoarse + ind * r.ngridmax + used
        assert np.array_equal(s["son"][c - 1], r.son[c])
    assert np.array_equal(s["father"][used - 1], r.father[used]) and np.array_equal(s["nbor"][:, used - 1], r.nbor[:, used])
    assert np.array_equal(s["xg"][:, used - 1], r.xg[:, used])
    for l in range(1, 9):
        assert np.array_equal(s["active"][l - 1], np.asarray(r.active[l], dtype=np.int32))
    U = r.uold.reshape(3, r.ncell)
    act = np.concatenate([np.asarray(r.active[l], dtype=np.int64) for l in range(1, 9)])
    for ind in range(2):
        c = r.ncoarse + ind * r.ngridmax + act - 1
        assert np.array_equal(s["uold"][0, c], U[0, c])
        assert np.allclose(s["uold"][1:, c], U[1:, c], rtol=4e-16, atol=1e-300)
    d2 = tmp_path / "b"
    d2.mkdir()
    write_snapshot(str(d2), 3, ndim=1, nvar=3, levelmin=3, nlevelmax=8, ngridmax=s["ngridmax"], ncoarse=s["ncoarse"], nxyz=s["nxyz"],
                   coarse_min=(1, 0, 0), coarse_max=(1, 0, 0), boxlen=s["boxlen"], gamma=s["gamma"], smallr=r.p.smallr, son=s["son"],
                   father=s["father"], nbor=s["nbor"], xg=s["xg"], active=s["active"], boundary=s["boundary"], uold=s["uold"],
                   t=s["t"], dtold=s["dtold"], dtnew=s["dtnew"], nstep=s["nstep"], nstep_coarse=s["nstep_coarse"], tout=s["tout"],
                   flag1=s["flag1"], cpu_map=s["cpu_map"])
    f = "output_00003/amr_00003.out00001"
    assert open(d1 / f, "rb").read() == open(d2 / f, "rb").read()
