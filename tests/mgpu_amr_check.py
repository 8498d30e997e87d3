#!/usr/bin/env python
"""Multi-GPU AMR parity check, run under torchrun:   torchrun --nproc-per-node N tests/mgpu_amr_check.py [ndim]

Every rank holds the same refined oct tree (built by the oracle's AMR driver); octs are owned by x-slabs, all other octs of a
level are ghost ("reception") octs.  One coarse step with sub-cycling runs in amr_step order through the C-ABI in AMR mode:
packed NCCL forward exchange of uold (make_virtual_fine), reverse accumulation of the coarse refluxes (make_virtual_reverse),
NCCL min/sum all-reduce in courant_fine.  Rank 0 repeats the step single-rank and compares the owned cells of every rank
(identical up to the summation order of refluxes arriving from different ranks: <= 1e-13 relative)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ramses_b200 import lib as _l                      # noqa: E402
from ramses_b200.hydro import HydroGPU, amr_step       # noqa: E402
from test_gpu_amr import commons_from_run, BOUND_2D_WALLS_X_OUTFLOW_Y   # noqa: E402


def build_run(ndim):
    from oracle.amr import AmrRun
    if ndim == 3:
        regs = [dict(type="square", x_center=0.5, y_center=0.5, z_center=0.5, length_x=10, length_y=10, length_z=10, exp_region=10, d=1.0, p=0.1),
                dict(type="square", x_center=0.45, y_center=0.45, z_center=0.55, length_x=0.4, length_y=0.3, length_z=0.3, exp_region=2, d=1.5, p=2.0)]
        r = AmrRun(3, 3, 5, (0,) * 6, 1.0, nsubcycle=[2, 2], ngridmax=40000, riemann="hllc", slope_type=1, err_grad_d=0.05,
                   err_grad_u=0.05, err_grad_p=0.05, interpol_type=1, regions=regs, tout=[1e9])
    else:
        regs = [dict(type="square", x_center=0.5, y_center=0.5, length_x=10, length_y=10, exp_region=10, d=1.0, p=0.1),
                dict(type="square", x_center=0.45, y_center=0.4, length_x=0.4, length_y=0.25, exp_region=2, d=2.0, u=0.3, v=-0.2, p=1.0)]
        r = AmrRun(2, 3, 6, (1, 1, 2, 2, 0, 0), 1.0, nsubcycle=[1, 2, 2], ngridmax=40000, riemann="hllc", slope_type=2,
                   err_grad_d=0.05, err_grad_u=0.05, err_grad_p=0.05, interpol_type=2, regions=regs, tout=[1e9],
                   bound_regions=BOUND_2D_WALLS_X_OUTFLOW_Y)      # y regions cover the corner cells (implosion.nml:17-24)
    r.flag_coarse(); r.init_refine(); r.init_refine_2()
    for i in range(r.levelmin, r.nlevelmax + 1):
        if i > r.levelmin:
            r.make_boundary_hydro(i)
        r.refine_fine(i)
    for l in range(1, r.nlevelmax + 1):
        r.make_boundary_hydro(l)
    return r


def main():
    ndim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    empty_top = len(sys.argv) > 2 and sys.argv[2] == "empty"     # the last rank owns NO oct of the finest level
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    r = build_run(ndim)
    L = r.nlevelmax
    itype = 2 if ndim == 2 else 1
    ntot = {l: len(r.active[l]) for l in range(1, L + 2)}
    owner = {}
    for l in range(1, L + 1):
        ig = np.array(r.active[l], dtype=np.int64)
        frac = r.xg[0, ig] - r.m.icoarse_min
        owner[l] = np.minimum((frac * world).astype(np.int64), world - 1)
        if empty_top and l == L:
            owner[l] = np.minimum(owner[l], max(world - 2, 0))

    def make_commons(ncpu, myid):
        a = commons_from_run(r, "hllc", r.p.slope_type)
        a.ncpu, a.myid = ncpu, myid
        if ncpu > 1:
            for l in range(1, L + 1):
                ig = np.array(r.active[l], dtype=np.int32)
                mine = owner[l] == myid - 1
                a.active[l] = np.ascontiguousarray(ig[mine])
                a.reception[l] = [np.sort(ig[owner[l] == c]) if c != myid - 1 else np.zeros(0, np.int32) for c in range(ncpu)]
                a.emission[l] = [np.sort(ig[mine]) if c != myid - 1 else np.zeros(0, np.int32) for c in range(ncpu)]
        return a

    a = make_commons(world, rank + 1)
    h = HydroGPU(a, device=lr, amr_mode=True, interpol_type=itype)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        _l.check(h.L.rgpu_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    buf = (C.c_ubyte * 128)(*uid.cpu().tolist())
    _l.check(h.L.rgpu_comm_init(world, rank, buf))
    for l in range(1, L + 1):
        if ntot[l]:
            h.bind_level(l)          # a rank without active octs at a level is bound too: it takes part in the exchanges
    tot = h.level_totals()           # numbtot(1,:) through the library (NCCL sum of the per-rank active counts)
    assert all(tot[l] == ntot[l] for l in range(1, L + 1)), (tot, ntot)
    if empty_top and rank == world - 1:
        assert len(a.active[L]) == 0
    a._u_start = a.uold.copy()
    h.upload_state(0)
    dtnew = {l: r.dtnew[l] for l in range(0, L + 2)}
    dtold = {l: r.dtold[l] for l in range(0, L + 2)}
    amr_step(h, r.levelmin, 1, r.levelmin, r.nsubcycle, dtnew, dtold, multi_rank=True)     # the library's host mirror of amr_step
    h.download_state(0)
    # the same coarse step through rgpu_amr_steps (time steps on the device) from the same initial state: identical bits
    u_host_driven = a.uold.copy()
    a.uold[:, :] = a._u_start
    h.upload_state(0)
    nsl = [1] + [r.nsubcycle.get(l, 2) for l in range(1, L + 2)]        # by level, element 0 unused
    dts_dev = h.amr_steps(r.levelmin, nsl, 1)
    h.download_state(0)
    same_dev = bool(np.array_equal(a.uold, u_host_driven)) and dts_dev[0] == dtnew[r.levelmin]
    a.uold[:, :] = u_host_driven
    launches = sum(h.level_info(l).kernel_launches for l in range(1, L + 1) if ntot[l])
    h.finalize()
    # owned cells of every rank -> everybody
    mask = np.zeros(a.ncell, dtype=np.float64)
    for l in range(r.levelmin, L + 1):
        for ind in range(r.T):
            mask[r.ncoarse + ind * r.ngridmax + a.active[l].astype(np.int64) - 1] = 1.0
    mk = torch.from_numpy(mask).cuda()
    contrib = torch.from_numpy(a.uold.copy()).cuda() * mk[None, :]
    dist.all_reduce(contrib)
    dist.all_reduce(mk)
    ok = True
    if rank == 0:
        a1 = make_commons(1, 1)
        h1 = HydroGPU(a1, device=lr, amr_mode=True, interpol_type=itype)
        for l in range(1, L + 1):
            if ntot[l]:
                h1.bind_level(l)
        h1.upload_state(0)
        d1 = {l: r.dtnew[l] for l in range(0, L + 2)}
        d2 = {l: r.dtold[l] for l in range(0, L + 2)}
        amr_step(h1, r.levelmin, 1, r.levelmin, r.nsubcycle, d1, d2, multi_rank=False)
        h1.download_state(0)
        h1.finalize()
        sel = mk.cpu().numpy() > 0
        got = contrib.cpu().numpy()[:, sel]
        ref = a1.uold[:, sel]
        scale = np.abs(ref).max(axis=1, keepdims=True)
        err = float((np.abs(got - ref) / scale).max())
        nid = float((got == ref).mean())
        same_dt = d1[r.levelmin] == dtnew[r.levelmin]
        once = bool((mk.cpu().numpy()[sel] == 1).all())
        print(f"mgpu_amr_check world={world} ndim={ndim} octs/level={[ntot[l] for l in range(1, L + 1)]}: cells={int(sel.sum())} "
              f"max rel diff={err:.3e} identical fraction={nid:.6f} dt identical={same_dt} each cell owned once={once} launches/rank={launches}")
        ok = err <= 1e-13 and same_dt and once
    flag = torch.tensor([1 if (ok and same_dev) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"rgpu_amr_steps == host-driven amr_step on every rank: {bool(flag.item())}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
