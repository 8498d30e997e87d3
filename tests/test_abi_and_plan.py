"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol of include/ramses_gpu.h,
fails loudly without a GPU (no fallback), and the host-only level planner (rgpu_plan_level) maps oct trees to
dense boxes correctly.  No GPU compute in here."""
import ctypes as C
import re

import numpy as np
import pytest

from helpers import Case


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_header_symbol():
    from ramses_b200 import lib
    L = lib.load()
    syms = lib.exported_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ramses_gpu.h but not exported"
    assert L.rgpu_abi_version() == 4
    # every entry point documents the reference interface it replaces
    txt = open(lib.HEADER).read()
    assert len(re.findall(r"\.f90:\d+", txt)) >= 15


@pytest.mark.skipif(_have_gpu(), reason="needs a box without GPU")
def test_no_cpu_fallback():
    """Without a CUDA device every entry point fails loudly; nothing is computed on the host."""
    from ramses_b200 import lib
    from ramses_b200.hydro import HydroGPU
    c = Case(3, 2)
    with pytest.raises(lib.RgpuError) as e:
        HydroGPU(c.amr_commons())
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    L = lib.load()
    assert L.rgpu_set_unew(2) == -1          # not initialised
    assert L.rgpu_godunov_fine(2, 1.0, None, None) == -1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under ramses_b200/ may reference it."""
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ramses_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ramses_oracle" not in txt, f


@pytest.mark.parametrize("ndim,level,bound", [(3, 4, (0,) * 6), (3, 3, (1, 1, 2, 2, 1, 1)), (2, 5, (1, 1, 0, 0, 0, 0)),
                                               (1, 7, (1, 1, 0, 0, 0, 0)), (1, 6, (0,) * 6)])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_plan_level_dense_box(ndim, level, bound, order):
    from ramses_b200.hydro import plan_level
    c = Case(ndim, level, bound=bound, order=order, seed=4)
    a = c.amr_commons()
    info, slots = plan_level(a, level)
    assert info.dense == 1
    n = 1 << level
    for d in range(ndim):
        has_b = bound[2 * d] > 0
        assert info.wrap[d] == (0 if has_b else 1)
        assert info.own_hi[d] - info.own_lo[d] == n
        assert info.ncell_box[d] == n + (4 if has_b else 0)
        assert info.own_lo[d] == (2 if has_b else 0)
    # slot numbering = lattice order of the oracle's own oct positions
    m = c.mesh
    nocts = [info.ncell_box[d] // 2 if d < ndim else 1 for d in range(3)]
    assert info.nslot == nocts[0] * nocts[1] * nocts[2] == (slots > 0).sum()
    lo = [(m.s.icoarse_min, m.s.jcoarse_min, m.s.kcoarse_min)[d] * (1 << (level - 1)) - (1 if bound[2 * d] else 0) for d in range(3)]
    for ig in list(m.active(level)[:50]) + [g for b in range(m.s.nboundary) for g in m.bound(b, level)[:20]]:
        pos = m.oct_pos(level, ig)
        s = (pos[0] - lo[0]) + nocts[0] * ((pos[1] - lo[1] if ndim > 1 else 0) + nocts[1] * (pos[2] - lo[2] if ndim > 2 else 0))
        assert slots[s] == ig


def test_plan_level_rejects_partial_level():
    """A level whose active octs do not fill a box is bound as non-dense (AMR path), not silently mis-tiled."""
    from ramses_b200.hydro import plan_level
    c = Case(3, 3)
    a = c.amr_commons()
    a.active[3] = a.active[3][:-5]
    info, _ = plan_level(a, 3)
    assert info.dense == 0


def test_tree_builder_matches_oracle_builder():
    """ramses_b200.tree (bench / multi-GPU plumbing) fabricates the same tree as the oracle's builder."""
    from oracle import orc
    from ramses_b200.tree import build_uniform_tree
    a = build_uniform_tree(3, 4, order="creation")
    m = orc.Mesh(3, 4, order=0)
    n = m.s.ngrid_used
    assert np.array_equal(a.father[:n] % 1, 0 * a.father[:n])
    # father / son / nbor identical up to the different ngridmax stride
    def split(c, ngm, nco):
        c = np.asarray(c, dtype=np.int64)
        ind = np.where(c > nco, (c - nco - 1) // ngm, -1)
        g = np.where(c > nco, (c - nco - 1) % ngm + 1, c)
        return ind, g
    fa, fm = split(a.father[:n], a.ngridmax, a.ncoarse), split(m.father()[1:n + 1], m.ngridmax, m.ncoarse)
    assert np.array_equal(fa[0], fm[0]) and np.array_equal(fa[1], fm[1])
    for j in range(6):
        na, nm = split(a.nbor[j, :n], a.ngridmax, a.ncoarse), split(m.nbor()[j, 1:n + 1], m.ngridmax, m.ncoarse)
        assert np.array_equal(na[0], nm[0]) and np.array_equal(na[1], nm[1])
    assert np.array_equal(a.active[4], m.active(4))


def test_nested_tree_fabricator_invariants():
    """ramses_b200.tree.build_nested_tree: son/father/nbor are mutually consistent, the refined cubes are nested 2:1 and the
    oracle's neighbour walk (get3cubefather) finds an existing father cell for every neighbour of every oct."""
    import ctypes as C
    from oracle import orc
    from ramses_b200.tree import build_nested_tree, leaf_cells
    levelmin, levelmax, hw = 4, 6, 3
    a = build_nested_tree(levelmin, levelmax, half_width=hw)
    assert [len(a.active[l]) for l in range(1, levelmax + 1)] == [1, 8, 64, 512, (2 * hw) ** 3, (2 * hw) ** 3]
    for l in range(2, levelmax + 1):
        ig = a.active[l].astype(np.int64)
        f = a.father[ig - 1].astype(np.int64)
        assert (a.son[f - 1] == ig).all()                                   # son(father(igrid)) == igrid
        for d in range(3):                                                   # nbor(nbor) symmetry through the sons
            left, right = a.nbor[2 * d, ig - 1], a.nbor[2 * d + 1, ig - 1]
            assert (left > 0).all() and (right > 0).all()
    leaves = sum(len(leaf_cells(a, l)) for l in range(1, levelmax + 1))
    assert leaves == (8 * 512 - (2 * hw) ** 3) + 7 * (2 * hw) ** 3 + 8 * (2 * hw) ** 3
    # the oracle's tree walk agrees with the fabricated arrays
    L = orc.lib()
    L.orc_mesh_new.restype = C.POINTER(orc.MeshS)
    L.orc_mesh_new.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
    m = L.orc_mesh_new(3, (C.c_int * 6)(0, 0, 0, 0, 0, 0), a.ngridmax, levelmax)
    s = m.contents
    np.ctypeslib.as_array(s.son, shape=(s.ncell + 1,))[1:] = a.son
    np.ctypeslib.as_array(s.father, shape=(s.ngridmax + 1,))[1:] = a.father
    np.ctypeslib.as_array(s.nbor, shape=(6, s.ngridmax + 1))[:, 1:] = a.nbor
    nfc = (C.c_int * 27)()
    for l in range(2, levelmax + 1):
        for g in a.active[l][:: max(1, len(a.active[l]) // 50)]:
            L.orc_get3cubefather(m, int(a.father[g - 1]), l, nfc, None)
            assert all(c > 0 for c in nfc)
            assert nfc[13] == a.father[g - 1]


def test_xbound_tree_matches_the_oracle_mesh_plan():
    """ramses_b200.tree.build_uniform_tree(xbound=...) (3x1x1 coarse grid, boundary octs at the x ends, namelist/tube_mhd.nml
    BOUNDARY_PARAMS) is planned as the same dense box as the oracle's mesh with the same boundary types."""
    from helpers import MhdCase
    from ramses_b200.hydro import plan_level
    from ramses_b200.tree import build_uniform_tree
    a = build_uniform_tree(3, 4, mhd=True, xbound=(2, 2), boxlen=2.0)
    ia, _ = plan_level(a, 4)
    b = MhdCase(4, bound=(2, 2, 0, 0, 0, 0), boxlen=2.0).amr_commons()
    ib, _ = plan_level(b, 4)
    assert ia.dense == 1 and ib.dense == 1
    for f in ("ncell_box", "own_lo", "own_hi", "wrap"):
        assert list(getattr(ia, f)) == list(getattr(ib, f))
    assert ia.nslot == ib.nslot and a.boundary_type == b.boundary_type == [11, 12]
    assert a.uold.shape[0] == 11 and (a.icoarse_min, a.icoarse_max) == (1, 1)


def test_host_amr_step_mirror_call_order():
    """ramses_b200.hydro.amr_step issues the per-level calls in the order of amr/amr_step.f90 (recursion, sub-cycling 1,2,2)."""
    from ramses_b200.hydro import amr_step

    class Rec:
        def __init__(self, a):
            self.a, self.log = a, []

        def __getattr__(self, name):
            def f(l):
                self.log.append((name, l))
                return 0.125 / l if name == "courant_fine" else None
            return f

    class A:
        pass
    a = A()
    a.active = {3: [1], 4: [1], 5: [1]}
    a.nlevelmax, a.boxlen, a.smallc, a.dtnew = 5, 1.0, 1e-10, {}
    h = Rec(a)
    dtnew, dtold = {l: 0.0 for l in range(8)}, {l: 0.0 for l in range(8)}
    amr_step(h, 3, 1, 3, [1, 1, 1, 2, 2, 2, 2], dtnew, dtold)
    god = [l for n, l in h.log if n == "godunov_fine_dev"]
    assert god == [5, 5, 4, 5, 5, 4, 3]                               # post-order: 2 fine steps per coarser step
    per = [n for n, l in h.log if l == 3]
    assert per == ["courant_fine", "set_unew", "godunov_fine_dev", "set_uold", "upload_fine", "make_boundary_hydro"]
    assert dtnew[4] <= dtnew[3] / 2 + 1e-18 and dtnew[5] <= dtnew[4] / 2 + 1e-18


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with the contract's keys,
    needs no GPU, and times the oracle port on the host cores."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "sedov3d_256_exact"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1
    j = json.loads(line[0])
    assert j["impl"] == "reference" and j["metric"] == "cell_updates_per_s" and j["unit"] == "cell-updates/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["value"] > 1e5
    assert j["config"]["workload"] == "sedov3d_256_exact"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0


@pytest.mark.parametrize("order,ncpu,expect", [("lattice", 1, True), ("lattice", 8, True), ("creation", 1, False), ("random", 8, False)])
def test_level0_pipeline_plan_host_only(order, ncpu, expect):
    """plan of the Level-0 slab pipeline (host only, rgpu_plan_level): a spatially coherent oct numbering gives a plan -- also for
    a rank of a 2x2x2 decomposition, whose owned octs inside a slab are equally spaced runs (pitched copies) --, the reference's
    creation order and a random numbering scatter every slab over the igrid window (no plan: serial order)."""
    from ramses_b200.hydro import plan_level
    from ramses_b200.tree import build_uniform_tree, coarse_dims_for_ranks
    coarse = coarse_dims_for_ranks(3, ncpu)
    a = build_uniform_tree(3, 6, coarse=coarse, myid=1, ncpu=ncpu, order=order, boxlen=1.0)
    info, slots = plan_level(a, 6)
    assert info.dense == 1
    assert (info.pipeline_slabs >= 3) == expect, info.pipeline_slabs


def test_params_struct_mirrors_agree():
    import os
    """struct rgpu_params: the C header, the ctypes mirror (ramses_b200/lib.py) and the ISO_C_BINDING type shown to the Fortran
    maintainer (INTEGRATION.md) list the same fields, in the same order, with the same types."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = open(os.path.join(root, "include", "ramses_gpu.h")).read()
    body = h[h.index("typedef struct rgpu_params {") + len("typedef struct rgpu_params {"):h.index("} rgpu_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for stmt in body.split(";"):
        m = re.match(r"\s*(int|double)\s+(.*)", stmt.strip().replace("\n", " "))
        if m:
            c_fields += [(m.group(1), n.strip()) for n in m.group(2).split(",")]
    from ramses_b200 import lib as _l
    import ctypes as C
    py_fields = [("int" if t is C.c_int else "double", n) for n, t in _l.Params._fields_]
    assert c_fields == py_fields
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    blk = doc[doc.index("type, bind(C) :: rgpu_params"):doc.index("end type")]
    f_fields = []
    for line in blk.splitlines()[1:]:
        line = line.split("!")[0]
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\))\s*::\s*(.*)", line)
        if m:
            f_fields += [("int" if "integer" in m.group(1) else "double", n.strip()) for n in m.group(2).split(",") if n.strip()]
    assert f_fields == c_fields
    assert f"RGPU_ABI_VERSION ({_l.load().rgpu_abi_version()})" in doc


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md (the binding a maintainer adds) mentions every entry point of include/ramses_gpu.h."""
    import os
    from ramses_b200 import lib as _l
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    assert [s for s in _l.exported_symbols() if s not in doc] == []
