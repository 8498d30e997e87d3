"""CPU tests of the oracle (oracle/ramses_oracle.c) against the reference's own fixtures and invariants."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import Case, SEDOV3D_REGIONS, SOD_REGIONS, smooth_state

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_indices3cube_match_reference_tables(orc):
    """Generated neighbour tables == getindices3cube (amr/nbors_utils.f90:305-358)."""
    tab = json.load(open(os.path.join(GOLD, "indices3cube.json")))
    L = orc.lib()
    n = 0
    for key, ref_l in tab["lll"].items():
        ndim, ind = [int(x) for x in key.split(",")]
        lll = (C.c_int * 27)()
        mmm = (C.c_int * 27)()
        L.orc_getindices3cube(ndim, ind, lll, mmm)
        assert list(lll)[:3 ** ndim] == ref_l
        assert list(mmm)[:3 ** ndim] == tab["mmm"][key]
        n += 1
    assert n == 2 + 4 + 8


@pytest.mark.parametrize("riemann", ["exact", "hllc", "acoustic", "hll", "llf"])
def test_sod_tube_uniform_vs_analytic(orc, riemann):
    """tube1d.nml on a uniform 1024-cell grid vs the reference's exact solution (tests/hydro/sod-tube/sod-tube-ana.dat).
    Discretisation-level pin of every Riemann solver (first-order convergent at the discontinuities)."""
    ana = np.array(json.load(open(os.path.join(GOLD, "sod_tube_ana.json")))["rows"])
    c = Case(1, 10, riemann=riemann, slope_type=2, bound=(1, 1, 0, 0, 0, 0), boxlen=1.0)
    c.init_regions(SOD_REGIONS)
    u, t = c.u.copy(), 0.0
    while t < 0.245:
        u, dts = c.oracle_steps(1, u)
        t += dts[0]
    d = c.dense(u)
    rho = d[0, 0, 0]
    vel = d[1, 0, 0] / rho
    p = (c.p.gamma - 1) * (d[2, 0, 0] - 0.5 * rho * vel ** 2)
    x = (np.arange(1024) + 0.5) / 1024
    assert np.allclose(x, ana[:, 1], atol=6e-5)   # the file prints 4 significant digits
    # the analytic file is at t=0.245, the run overshoots by < 1 dt: compare in L1
    tol = {"exact": 2.5e-3, "hllc": 2.5e-3, "acoustic": 2.5e-3, "hll": 3e-3, "llf": 4e-3}[riemann]
    assert np.abs(rho - ana[:, 3]).mean() < tol
    assert np.abs(vel - ana[:, 2]).mean() < 2 * tol
    assert np.abs(p - ana[:, 4]).mean() < tol
    # conservation (reflexive walls)
    assert abs(rho.mean() - 0.5625) < 1e-14


def test_sod_tube_convergence(orc):
    """L1 error against the analytic solution drops ~linearly with resolution (shock-dominated)."""
    ana = np.array(json.load(open(os.path.join(GOLD, "sod_tube_ana.json")))["rows"])
    errs = []
    for lev in (7, 8, 9):
        c = Case(1, lev, riemann="hllc", slope_type=2, bound=(1, 1, 0, 0, 0, 0))
        c.init_regions(SOD_REGIONS)
        u, t = c.u.copy(), 0.0
        while t < 0.245:
            u, dts = c.oracle_steps(1, u)
            t += dts[0]
        n = 1 << lev
        x = (np.arange(n) + 0.5) / n
        errs.append(np.abs(c.dense(u)[0, 0, 0] - np.interp(x, ana[:, 1], ana[:, 3])).mean())
    assert errs[1] < 0.7 * errs[0] and errs[2] < 0.7 * errs[1]


@pytest.mark.parametrize("riemann", ["llf", "hllc", "exact", "acoustic", "hll"])
def test_sedov3d_conservation_and_symmetry(orc, riemann):
    """Periodic sedov3d: mass/energy to round-off (mcons/econs ~1e-16, doc/wiki/Start.md:174-187) and the
    x<->y<->z permutation symmetry of a corner blast."""
    c = Case(3, 4, riemann=riemann, slope_type=1, boxlen=0.5)
    c.init_regions(SEDOV3D_REGIONS)
    d0 = c.dense()
    u, dts = c.oracle_steps(8)
    d = c.dense(u)
    assert abs(d[0].sum() - d0[0].sum()) <= 1e-14 * d0[0].sum()
    assert abs(d[4].sum() - d0[4].sum()) <= 1e-13 * d0[4].sum()
    assert np.abs(d[0] - d[0].transpose(0, 2, 1)).max() < 1e-13
    assert np.abs(d[1] - d[2].transpose(0, 2, 1)).max() < 1e-11 * np.abs(d[1]).max()
    assert np.abs(d[1] - d[3].transpose(2, 1, 0)).max() < 1e-11 * np.abs(d[1]).max()


def test_oct_order_independence(orc):
    """Results do not depend on the igrid numbering (nvector only batches octs)."""
    outs = []
    for order in (0, 1, 2):
        c = Case(3, 3, riemann="hllc", slope_type=2, order=order, seed=9)
        c.init_dense(smooth_state(3, 8))
        u, _ = c.oracle_steps(3)
        outs.append(c.dense(u))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_ndim_embedding(orc):
    """A y,z-independent 3-D state evolves like the 1-D run (slope/trace/riemann formulas differ in form per NDIM
    build but agree to round-off)."""
    n = 16
    d1 = smooth_state(1, n)
    c1 = Case(1, 4, riemann="hllc", slope_type=2)
    c1.init_dense(d1)
    c3 = Case(3, 4, riemann="hllc", slope_type=2)
    d3 = np.zeros((5, n, n, n))
    d3[0] = d1[0, 0, 0][None, None, :]
    d3[1] = d1[1, 0, 0][None, None, :]
    d3[4] = d1[2, 0, 0][None, None, :]
    c3.init_dense(d3)
    dt1, _ = c1.oracle_courant()
    dt = 0.3 * dt1
    u1 = c1.dense(c1.oracle_godunov(dt))
    u3 = c3.dense(c3.oracle_godunov(dt))
    assert np.allclose(u3[0][3, 5], u1[0][0, 0], rtol=1e-13, atol=1e-15)
    assert np.allclose(u3[4][3, 5], u1[2][0, 0], rtol=1e-13, atol=1e-15)
    assert np.abs(u3[2]).max() < 1e-14 and np.abs(u3[3]).max() < 1e-14


@pytest.mark.parametrize("riemann", ["hllc", "llf", "hll", "acoustic", "exact"])
@pytest.mark.parametrize("slope_type", [1, 2])
def test_2d_run_equals_z_invariant_3d_run_bitwise(orc, riemann, slope_type):
    """Several level steps of a 2-D run == the same steps of the z-invariant 3-D run (w = 0), BIT FOR BIT, for every solver:
    the NDIM=3 branches of ctoprim / uslope / trace3d / cmpflxm / riemann_* / godfine1 reduce to the NDIM=2 branches plus exact
    zeros.  The NDIM=2 branches reproduce the reference's implosion golden file (test_oracle_golden.py), so this carries that
    pin to the NDIM=3 code the GPU kernels are compared with (y<->z and x<->z by the permutation-symmetry test)."""
    n = 16
    d2 = smooth_state(2, n)
    rough = np.random.default_rng(7).standard_normal(d2[0].shape)
    d2[0] *= 1 + 0.3 * (rough > 1.0)                         # a few density jumps so that the limiters switch
    c2 = Case(2, 4, riemann=riemann, slope_type=slope_type)
    c2.init_dense(d2)
    c3 = Case(3, 4, riemann=riemann, slope_type=slope_type)
    d3 = np.zeros((5, n, n, n))
    d3[0], d3[1], d3[2], d3[4] = d2[0][0][None], d2[1][0][None], d2[2][0][None], d2[3][0][None]
    c3.init_dense(d3)
    u2, u3 = c2.u, c3.u
    L = orc.lib()
    for _ in range(4):
        dt2, _ = c2.oracle_courant(u2)
        dt = 0.9 * dt2                                        # cmpdt sums ndim sound speeds: use the 2-D step in both runs
        u2 = c2.oracle_godunov(dt, u2, nthreads=1)
        u3 = c3.oracle_godunov(dt, u3, nthreads=1)
    a2, a3 = c2.dense(u2), c3.dense(u3)
    for k in (0, 7, 15):
        assert np.array_equal(a3[0][k], a2[0][0]) and np.array_equal(a3[1][k], a2[1][0])
        assert np.array_equal(a3[2][k], a2[2][0]) and np.array_equal(a3[4][k], a2[3][0])
    assert np.abs(a3[3]).max() == 0.0


def test_riemann_solvers_agree_on_uniform_state(orc):
    """Equal left and right states: every solver returns the exact Euler flux."""
    p = orc.make_params(ndim=3, nvector=4)
    ql = np.zeros((5, 4)); qr = np.zeros((5, 4))
    ql[0] = [1.0, 0.5, 2.0, 1e-3]; ql[1] = [0.3, -0.2, 0.0, 1.5]; ql[2] = [1.0, 0.1, 5.0, 1e-2]; ql[3] = 0.1; ql[4] = -0.2
    qr[:] = ql
    entho = 1 / (p.gamma - 1)
    etot = ql[2] * entho + 0.5 * ql[0] * (ql[1] ** 2 + ql[3] ** 2 + ql[4] ** 2)
    exact = np.array([ql[0] * ql[1], ql[0] * ql[1] ** 2 + ql[2], ql[1] * (etot + ql[2]), ql[0] * ql[1] * ql[3], ql[0] * ql[1] * ql[4]])
    for nm in ("llf", "hll", "hllc", "acoustic", "approx"):
        fg = np.zeros((6, 4))
        getattr(orc.lib(), "orc_riemann_" + nm)(C.byref(p), orc.dptr(ql), orc.dptr(qr), orc.dptr(fg), 4)
        assert np.allclose(fg[:5], exact, rtol=1e-12, atol=1e-14), nm


def test_cmpdt_matches_formula(orc):
    p = orc.make_params(ndim=3, nvector=8, courant_factor=0.8)
    rs = np.random.RandomState(1)
    uu = np.zeros((5, 8)); uu[0] = 1 + rs.rand(8); uu[1:4] = rs.randn(3, 8) * 0.3; uu[4] = 3 + rs.rand(8)
    ref = uu.copy()
    dt = C.c_double()
    dx = 1 / 64
    orc.lib().orc_cmpdt(C.byref(p), orc.dptr(uu), None, dx, C.byref(dt), 8)
    v = ref[1:4] / ref[0]
    pr = (p.gamma - 1) * (ref[4] - 0.5 * ref[0] * (v ** 2).sum(0))
    ws = 3 * np.sqrt(p.gamma * pr / ref[0]) + np.abs(v).sum(0)
    g = 1e-4
    expect = (dx / ws * (np.sqrt(1 + 2 * 0.8 * g) - 1) / g).min()
    assert abs(dt.value - expect) < 1e-14 * expect


def test_oracle_difmag_is_conservative_and_only_acts_in_compressions():
    """cmpdivu + consup (hydro/uplmde.f90:702,769): flux += dt*difmag*min(0, div u)*dU -- conservative, and a no-op where
    the velocity field expands everywhere."""
    from helpers import Case, smooth_state
    n = 8
    u0 = smooth_state(3, n)
    u0[1] += u0[0] * 0.8 * np.sin(2 * np.pi * (np.arange(8) + 0.5) / 8)[None, None, :]   # compressive x-velocity: div u < 0 around x = 0.5
    outs = []
    for dm in (0.0, 0.4):
        c = Case(3, 3, riemann="hllc", slope_type=1)
        c.p.difmag = dm
        c.init_dense(u0)
        dt, _ = c.oracle_courant()
        outs.append(c.dense(c.oracle_godunov(dt)))
    assert np.abs(outs[1] - outs[0]).max() > 1e-6
    for iv in range(5):
        assert abs(outs[1][iv].sum() - u0[iv].sum()) <= 1e-12 * max(np.abs(u0[iv]).sum(), 1.0)
    # pure expansion: v = +x (periodic box: restrict the check to the interior where div u > 0 at every vertex)
    x = (np.arange(n) + 0.5) / n - 0.5
    ue = np.zeros((5, n, n, n))
    ue[0] = 1.0
    ue[1] = x[None, None, :]; ue[2] = x[None, :, None]; ue[3] = x[:, None, None]
    ue[4] = 2.5 + 0.5 * (ue[1] ** 2 + ue[2] ** 2 + ue[3] ** 2)
    res = []
    for dm in (0.0, 0.4):
        c = Case(3, 3, riemann="hllc", slope_type=1)
        c.p.difmag = dm
        c.init_dense(ue)
        res.append(c.dense(c.oracle_godunov(1e-3)))
    assert np.array_equal(res[0][:, 2:-2, 2:-2, 2:-2], res[1][:, 2:-2, 2:-2, 2:-2])


def test_oracle_passive_scalars_keep_uniform_concentration_and_are_conserved():
    """NVAR = NDIM+2+2: passive scalars (q = u/rho in ctoprim, advective trace, upwinded with the mass flux by every solver).
    A uniform concentration stays uniform to round-off, the scalar mass is conserved, and the hydro variables do not notice."""
    from helpers import Case, smooth_state
    n = 8
    u5 = smooth_state(3, n)
    for riemann in ("llf", "hll", "hllc", "acoustic", "exact"):
        c5 = Case(3, 3, riemann=riemann, slope_type=1)
        c5.init_dense(u5)
        ref5, dts5 = c5.oracle_steps(3)
        c7 = Case(3, 3, riemann=riemann, slope_type=1, nvar=7)
        u7 = np.zeros((7, n, n, n))
        u7[:5] = u5
        u7[5] = 0.3 * u5[0]                                               # uniform concentration 0.3
        u7[6] = u5[0] * (0.5 + 0.4 * np.sin(2 * np.pi * (np.arange(n) + 0.5) / n))[None, :, None]
        c7.init_dense(u7)
        ref7, dts7 = c7.oracle_steps(3)
        d5, d7 = c5.dense(ref5), c7.dense(ref7)
        assert np.array_equal(dts5, dts7) and np.array_equal(d5, d7[:5])
        assert np.abs(d7[5] / d7[0] - 0.3).max() < 1e-14
        assert abs(d7[6].sum() - u7[6].sum()) < 1e-12 * u7[6].sum()
        assert np.abs(d7[6] - u7[6]).max() > 1e-3


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("ivar,itype", [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (2, 4)])
def test_interpol_hydro_variants(orc, ndim, ivar, itype):
    """interpol_hydro (hydro/interpol_hydro.f90:268) for every interpol_var / interpol_type: the children average back to the
    father for mass and momentum (interpol_var 2 through its explicit momentum correction :393-415); total energy for
    interpol_var=0, internal energy for interpol_var=1; a uniform neighbourhood is reproduced exactly; the limited types keep
    every child inside the range of the father's neighbourhood for the limited variables."""
    import ctypes as C
    L = orc.lib()
    L.orc_set_interpol.argtypes = [C.c_int, C.c_int]
    L.orc_interpol_hydro.argtypes = [C.POINTER(orc.Params), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    p = orc.make_params(ndim=ndim)
    nvar, T, nn = ndim + 2, 1 << ndim, 2 * ndim + 1
    rng = np.random.default_rng(10 * ndim + ivar + itype)
    try:
        L.orc_set_interpol(itype, ivar)
        for trial in range(20):
            u1 = np.zeros((nn, nvar))
            u1[:, 0] = 1 + 0.5 * rng.random(nn)
            vel = 0.4 * rng.standard_normal((nn, ndim))
            u1[:, 1:1 + ndim] = u1[:, :1] * vel
            eint = 1 + 0.5 * rng.random(nn)
            u1[:, ndim + 1] = eint + 0.5 * u1[:, 0] * (vel ** 2).sum(axis=1)
            if trial == 0:
                u1[:] = u1[0]
            u2 = np.zeros((T, nvar))
            L.orc_interpol_hydro(C.byref(p), orc.dptr(np.ascontiguousarray(u1)), orc.dptr(u2))
            if trial == 0:
                assert np.allclose(u2, u1[0], rtol=0, atol=1e-15)
            assert abs(u2[:, 0].mean() - u1[0, 0]) < 1e-14
            assert np.abs(u2[:, 1:1 + ndim].mean(axis=0) - u1[0, 1:1 + ndim]).max() < 1e-14
            ek = lambda u: 0.5 * (u[..., 1:1 + ndim] ** 2).sum(axis=-1) / u[..., 0]
            if ivar == 0:
                assert abs(u2[:, ndim + 1].mean() - u1[0, ndim + 1]) < 1e-14
            else:
                assert abs((u2[:, ndim + 1] - ek(u2)).mean() - (u1[0, ndim + 1] - ek(u1[0]))) < 1e-14
            if itype in (1, 2):
                assert u2[:, 0].min() >= u1[:, 0].min() - 1e-14 and u2[:, 0].max() <= u1[:, 0].max() + 1e-14
    finally:
        L.orc_set_interpol(1, 0)


def test_pressure_fix_restatement(orc):
    """pressure_fix (hydro/godunov_fine.f90:71-90,164-168,203-227,294-437,737-903): divu and enew ride along the sweep.
    (1) switched on with beta_fix=0 on a smooth flow the conserved state is bit-identical to the plain run (the energy switch
    only fires for e_cons < 0); (2) divu = -div(u) dt from the Riemann face velocities, enew = the internal energy advanced
    non-conservatively; (3) in a cold Mach-1e4 compressive flow the conservative internal energy is pure truncation noise while
    the run with beta_fix=0.5 keeps the pressure positive and the gas on its adiabat outside the steepening front."""
    import ctypes as C
    L = orc.lib()
    L.orc_set_pressure_fix.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double]
    n, ndim = 32, 2
    c = Case(ndim, 5, riemann="hllc", slope_type=1)
    d0 = smooth_state(ndim, n)
    c.init_dense(d0)
    plain, dts = c.oracle_steps(3, nthreads=1)
    divu, enew = np.zeros(c.mesh.s.ncell), np.zeros(c.mesh.s.ncell)
    try:
        L.orc_set_pressure_fix(orc.dptr(divu), orc.dptr(enew), 0.0)
        fixed, dts2 = c.oracle_steps(3, nthreads=1)
        assert np.array_equal(dts, dts2) and np.array_equal(plain, fixed)
        dv_of = lambda arr: c.mesh.level_to_dense(np.concatenate([arr, np.zeros(3 * c.mesh.s.ncell)]), c.level, 4)[0][0]
        b = c.dense(fixed)
        eint = b[3][0] - 0.5 * (b[1][0] ** 2 + b[2][0] ** 2) / b[0][0]
        en = dv_of(enew)
        assert np.abs(en - eint).max() < 2e-3 * np.abs(eint).max()
        x = (np.arange(n) + 0.5) / n
        X, Y = np.meshgrid(x, x, indexing="xy")
        for mode in ("x", "y"):                                       # divu of one step of a pure velocity wave, per direction
            uu = 0.1 * np.sin(2 * np.pi * X) if mode == "x" else 0 * X
            vv = 0.1 * np.sin(2 * np.pi * Y) if mode == "y" else 0 * X
            wave = np.zeros((4, 1, n, n))
            wave[0, 0], wave[1, 0], wave[2, 0], wave[3, 0] = 1, uu, vv, 1 / 0.4 + 0.5 * (uu * uu + vv * vv)
            c.init_dense(wave)
            _, dtw = c.oracle_steps(1, nthreads=1)
            div = 0.2 * np.pi * (np.cos(2 * np.pi * X) if mode == "x" else np.cos(2 * np.pi * Y))
            assert np.abs(dv_of(divu) + div * dtw[0]).max() < 0.06 * np.abs(div * dtw[0]).max()     # minmod clips the extrema of u
        # cold hypersonic compressive flow (Mach ~ 1e4, before shock formation): the conservative internal energy is the
        # difference of two O(1) numbers -> truncation noise, negative in places; with the fix the gas stays on its adiabat
        p0, g = 1e-8, 1.4
        uu, vv = 1 + 0.2 * np.sin(2 * np.pi * X), 0.5 + 0 * X
        cold = np.zeros((4, 1, n, n))
        cold[0, 0], cold[1, 0], cold[2, 0] = 1.0, uu, vv
        cold[3, 0] = p0 / (g - 1) + 0.5 * (uu * uu + vv * vv)
        c.init_dense(cold)
        L.orc_set_pressure_fix(None, None, 0.0)
        raw, _ = c.oracle_steps(10, nthreads=1)
        r = c.dense(raw)
        p_raw = (g - 1) * (r[3][0] - 0.5 * (r[1][0] ** 2 + r[2][0] ** 2) / r[0][0])
        L.orc_set_pressure_fix(orc.dptr(divu), orc.dptr(enew), 0.5)
        fx, _ = c.oracle_steps(10, nthreads=1)
        f = c.dense(fx)
        p_fix = (g - 1) * (f[3][0] - 0.5 * (f[1][0] ** 2 + f[2][0] ** 2) / f[0][0])
        s_fix = p_fix / f[0][0] ** g
        assert np.abs(f[0] - r[0]).max() < 2e-3                                          # same flow (the pressure is dynamically irrelevant)
        s_raw = p_raw / r[0][0] ** g
        on_adiabat = lambda s_: np.mean(np.abs(s_ / p0 - 1) < 0.05)
        assert on_adiabat(s_raw) < 0.1                                  # raw: garbage almost everywhere
        assert p_fix.min() > 0 and s_fix.min() > 0.9 * p0 and on_adiabat(s_fix) > 0.45   # fixed: adiabatic in the expanding half
    finally:
        L.orc_set_pressure_fix(None, None, 0.0)


def test_gravity_restatement(orc):
    """poisson (umuscl.f90:932-938, godunov_fine.f90:237-289,637-647, godunov_utils.f90:99-111).  No golden file of the reference
    exercises gravity without cooling, so the restatement is pinned by properties ("parity unpinned" for this option):
    (1) f = 0 reproduces the plain run to rounding; (2) a uniform gas in a uniform field g: the predictor shifts every face
    velocity alike (no flux differences), set_uold adds the half-step kick, so after one level step v = g*dt/2, density and
    internal energy unchanged; (3) the Courant step shrinks to dx/ws*(sqrt(1+2*CFL*gt)-1)/gt with gt = |g| dx / ws^2;
    (4) an isothermal atmosphere rho ~ exp(-g z / c^2) in hydrostatic balance stays put: velocities remain O(truncation error)
    instead of O(g t) (the kick cancels the pressure-gradient flux)."""
    import ctypes as C
    L = orc.lib()
    L.orc_set_gravity.argtypes = [C.POINTER(C.c_double)]
    n, ndim = 32, 2
    c = Case(ndim, 5, riemann="hllc", slope_type=1)
    nc = c.mesh.s.ncell
    c.init_dense(smooth_state(ndim, n))
    plain, dts = c.oracle_steps(2, nthreads=1)
    try:
        f = np.zeros((ndim, nc))
        L.orc_set_gravity(orc.dptr(f))
        same, dts2 = c.oracle_steps(2, nthreads=1)
        # add_gravity_source_terms rewrites momentum as d*(m/d) and E as (E-ekin)+ekin even for f = 0: last-bit differences only
        assert np.allclose(plain, same, rtol=1e-13, atol=1e-15) and np.allclose(dts, dts2, rtol=1e-13)
        # (2), (3): uniform gas, uniform field
        g = np.array([0.3, -0.2])
        f[:] = g[:, None]
        uni = np.zeros((4, 1, n, n))
        uni[0], uni[3] = 2.0, 1.0 / 0.4
        c.init_dense(uni)
        out, dt1 = c.oracle_steps(1, nthreads=1)
        b = c.dense(out)
        dx, ws = 1.0 / n, 2 * np.sqrt(1.4 * 1.0 / 2.0)
        gt = np.abs(g).sum() * dx / ws ** 2
        assert abs(dt1[0] - dx / ws * (np.sqrt(1 + 2 * 0.8 * max(gt, 1e-4)) - 1) / max(gt, 1e-4)) < 1e-15
        assert np.abs(b[0] - 2.0).max() == 0
        for k in range(ndim):
            assert np.abs(b[1 + k][0] / b[0][0] - g[k] * dt1[0] / 2).max() < 1e-16
        eint = b[3][0] - 0.5 * (b[1][0] ** 2 + b[2][0] ** 2) / b[0][0]
        assert np.abs(eint - 2.5).max() < 1e-14
        # (4): hydrostatic isothermal atmosphere along y between reflecting-free periodic images is not available on this periodic
        # box, so use a field that is itself periodic: g_y = -g0 sin(2 pi y), rho = exp(g0 cos(2 pi y)/(2 pi c2)), P = c2 rho
        g0, c2 = 2.0, 1.0
        y = (np.arange(n) + 0.5) / n
        rho = np.exp(g0 * np.cos(2 * np.pi * y) / (2 * np.pi * c2))
        atm = np.zeros((4, 1, n, n))
        atm[0, 0] = rho[:, None]
        atm[3, 0] = c2 * rho[:, None] / 0.4
        c.init_dense(atm)
        fd = np.zeros((4, 1, n, n))
        fd[1, 0] = (-g0 * np.sin(2 * np.pi * y))[:, None]
        fbuf = np.zeros(4 * nc)
        c.mesh.dense_to_level(fd, fbuf, c.level, 4)
        f[:] = fbuf.reshape(4, nc)[:2]
        # the path holds the first half of the kick (set_uold); synchro_hydro_fine (pm/synchro_hydro_fine.f90, outside the path)
        # applies the second half after the new force is known -- done here in numpy so that the balance can be observed
        u, t = c.u.copy(), 0.0
        for _ in range(20):
            u, dt1 = c.oracle_steps(1, u=u, nthreads=1)
            U = u.reshape(4, nc)
            d = np.maximum(U[0], 1e-10)
            vx, vy = U[1] / d, U[2] / d
            eprim = U[3] - 0.5 * d * (vx * vx + vy * vy)
            vy = vy + f[1] * 0.5 * dt1[0]
            U[2] = d * vy
            U[3] = eprim + 0.5 * d * (vx * vx + vy * vy)
            t += dt1[0]
        bb = c.dense(u)
        vmax = np.abs(bb[2][0] / bb[0][0]).max()
        assert vmax < 0.02 * g0 * t, (vmax, g0 * t)                  # free fall would give g0*t
        c.init_dense(atm)                                            # without the field the same atmosphere does accelerate
        L.orc_set_gravity(None)
        free, dtf = c.oracle_steps(20, nthreads=1)
        fb = c.dense(free)
        assert np.abs(fb[2][0] / fb[0][0]).max() > 10 * vmax
    finally:
        L.orc_set_gravity(None)


def test_imposed_boundary_supersonic_inflow(orc):
    """bound_type=3 (hydro/hydro_boundary.f90:229-252, default boundana): ghost cells hold boundary_var; a supersonic uniform
    inflow through the left face of a 1-D tube with an outflow right face keeps the uniform state exactly, and with a denser
    inflow the mass of the domain grows by (rho u)_in * dt per step until the front arrives at the other end."""
    import ctypes as C
    L = orc.lib()
    L.orc_set_boundary_var.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
    n = 64
    c = Case(1, 6, riemann="hllc", slope_type=1, bound=(3, 2, 0, 0, 0, 0))
    rho, u, p = 1.0, 3.0, 1.0                                     # Mach 2.5
    cons = np.array([rho, rho * u, p / 0.4 + 0.5 * rho * u * u])
    L.orc_set_boundary_var(0, orc.dptr(cons), 3)
    d = np.zeros((3, 1, 1, n))
    d[:, 0, 0, :] = cons[:, None]
    c.init_dense(d)
    out, dts = c.oracle_steps(5, nthreads=1)
    assert np.array_equal(c.dense(out), d)                         # uniform supersonic flow is a fixed point
    cons2 = np.array([2.0, 2.0 * u, p / 0.4 + 0.5 * 2.0 * u * u])
    L.orc_set_boundary_var(0, orc.dptr(cons2), 3)
    c.init_dense(d)
    out, dts = c.oracle_steps(5, nthreads=1)
    a = c.dense(out)
    dx = 1.0 / n
    assert abs((a[0].sum() - d[0].sum()) * dx - (2.0 * u - rho * u) * dts.sum()) < 1e-12     # in: 2u, out: rho u (front not there yet)
    assert a[0][0, 0, -8:].max() == rho and a[0][0, 0, 0] > 1.5
