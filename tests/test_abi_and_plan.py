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
    assert L.rgpu_abi_version() == 2
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
