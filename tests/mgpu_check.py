#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun:   torchrun --nproc-per-node N tests/mgpu_check.py [level] [steps] [riemann]

Every rank owns one coarse cell of an (nx,ny,nz) periodic coarse grid refined to `level`; K fused level steps run with
the packed NCCL ghost-oct exchange (rgpu_make_virtual_fine inside rgpu_level_steps) and a NCCL min-allreduce of dt.
Rank 0 then repeats the SAME global problem on one GPU (ncpu=1, same coarse grid) and compares cell by cell: the
decomposed run must be bit-identical (dt and conserved state)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ramses_b200 import lib as _l                      # noqa: E402
from ramses_b200.hydro import HydroGPU                 # noqa: E402
from ramses_b200.tree import build_uniform_tree, coarse_dims_for_ranks, fill_state  # noqa: E402


def state_fn(coarse):
    def fn(x, y, z):
        tw = 2 * np.pi
        xs, ys, zs = x / coarse[0], y / coarse[1], z / coarse[2]
        rho = 1 + 0.2 * np.sin(tw * xs) * np.cos(tw * ys)
        vx, vy, vz = 0.3 * np.sin(tw * ys), 0.3 * np.sin(tw * zs), 0.3 * np.sin(tw * xs)
        p = 1 + 0.1 * np.cos(tw * (xs + ys + zs))
        u = np.zeros((5, len(x)))
        u[0] = rho
        u[1], u[2], u[3] = rho * vx, rho * vy, rho * vz
        u[4] = p / 0.4 + 0.5 * rho * (vx ** 2 + vy ** 2 + vz ** 2)
        return u
    return fn


def mhd_state_fn(coarse):
    """Smooth periodic MHD state; B_x depends on (y,z) only etc.: divergence-free and the two copies of a face agree."""
    def fn(x, y, z):
        tw = 2 * np.pi
        xs, ys, zs = x / coarse[0], y / coarse[1], z / coarse[2]
        rho = 1 + 0.2 * np.sin(tw * xs) * np.cos(tw * ys)
        vx, vy, vz = 0.3 * np.sin(tw * ys), 0.2 * np.cos(tw * zs), 0.1 * np.sin(tw * xs)
        p = 1 + 0.1 * np.cos(tw * zs) * np.sin(tw * (xs + ys))
        bx = 0.5 + 0.2 * np.sin(tw * ys) * np.cos(tw * zs)
        by = 0.3 + 0.2 * np.sin(tw * zs + 1.0) * np.cos(tw * xs)
        bz = 0.2 + 0.2 * np.cos(tw * xs) * np.sin(tw * ys + 0.3)
        u = np.zeros((11, len(x)))
        u[0] = rho
        u[1], u[2], u[3] = rho * vx, rho * vy, rho * vz
        u[5], u[6], u[7] = bx, by, bz
        u[8], u[9], u[10] = bx, by, bz
        u[4] = p * 1.5 + 0.5 * rho * (vx ** 2 + vy ** 2 + vz ** 2) + 0.5 * (bx ** 2 + by ** 2 + bz ** 2)
        return u
    return fn


def keys_and_state(a, l):
    """(global cell key, state) of the active cells."""
    ig = a.active[l].astype(np.int64)
    p = a._pos[l][ig - a._igrid0[l]]
    ext = (np.array([a.nx, a.ny, a.nz], dtype=np.int64) << l)
    ks, us = [], []
    for ind in range(8):
        b = np.array([(ind >> d) & 1 for d in range(3)])
        c = 2 * p + b[None, :]
        ks.append(c[:, 0] + ext[0] * (c[:, 1] + ext[1] * c[:, 2]))
        us.append(a.uold[:, a.ncoarse + ind * a.ngridmax + ig - 1])
    return np.concatenate(ks), np.concatenate(us, axis=1)


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    riemann = sys.argv[3] if len(sys.argv) > 3 else "hllc"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    coarse = coarse_dims_for_ranks(3, world)
    mhd = riemann.startswith("mhd:")          # e.g. mhd:hlld:hlld = MHD build, riemann='hlld', riemann2d='hlld'

    def configure(t):
        if mhd:
            _, t.riemann, t.riemann2d = riemann.split(":")
            t.gamma = 5.0 / 3.0
        else:
            t.riemann = riemann
        t.slope_type, t.courant_factor = 1, 0.8
        fill_state(t, level, mhd_state_fn(coarse) if mhd else state_fn(coarse))
    a = build_uniform_tree(3, level, coarse=coarse, myid=rank + 1, ncpu=world, order="random", seed=rank + 5, mhd=mhd)
    configure(a)
    h = HydroGPU(a, device=lr)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        _l.check(h.L.rgpu_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    buf = (C.c_ubyte * 128)(*uid.cpu().tolist())
    _l.check(h.L.rgpu_comm_init(world, rank, buf))
    h.bind_level(level)
    h.upload_state(level)
    dts, _ = h.level_steps(level, steps)
    h.download_state(level)
    launches = h.level_info(level).kernel_launches
    h.finalize()
    k, u = keys_and_state(a, level)
    # gather everything on rank 0
    kt = torch.from_numpy(k).cuda()
    ut = torch.from_numpy(np.ascontiguousarray(u)).cuda()
    ks = [torch.zeros_like(kt) for _ in range(world)]
    us = [torch.zeros_like(ut) for _ in range(world)]
    dist.all_gather(ks, kt)
    dist.all_gather(us, ut)
    ok = True
    if rank == 0:
        kk = torch.cat(ks).cpu().numpy()
        uu = torch.cat(us, dim=1).cpu().numpy()
        order = np.argsort(kk)
        kk, uu = kk[order], uu[:, order]
        g = build_uniform_tree(3, level, coarse=coarse, myid=1, ncpu=1, order="creation", mhd=mhd)
        configure(g)
        hg = HydroGPU(g, device=lr)
        hg.bind_level(level)
        hg.upload_state(level)
        dts1, _ = hg.level_steps(level, steps)
        hg.download_state(level)
        hg.finalize()
        k1, u1 = keys_and_state(g, level)
        o1 = np.argsort(k1)
        k1, u1 = k1[o1], u1[:, o1]
        same_dt = np.array_equal(dts, dts1)
        same_u = np.array_equal(kk, k1) and np.array_equal(uu, u1)
        maxrel = float(np.abs(uu - u1).max() / np.abs(u1).max())
        print(f"mgpu_check world={world} level={level} steps={steps} riemann={riemann}: dt identical={same_dt} "
              f"state identical={same_u} maxrel={maxrel:.3e} launches/rank={launches}")
        ok = same_dt and same_u
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
