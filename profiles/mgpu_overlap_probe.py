#!/usr/bin/env python
"""torchrun --nproc-per-node N profiles/mgpu_overlap_probe.py <level> : ms per level step of the fused multi-GPU path with the
exchange overlap off / on and different numbers of SMs left to the exchange stream (RGPU_OVERLAP, RGPU_OVERLAP_RESERVE)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ramses_b200 import lib as _l  # noqa: E402
from ramses_b200.hydro import HydroGPU  # noqa: E402
from ramses_b200.tree import build_uniform_tree, coarse_dims_for_ranks, fill_state  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
coarse = coarse_dims_for_ranks(3, world)
a = build_uniform_tree(3, level, coarse=coarse, myid=rank + 1, ncpu=world, order="lattice", boxlen=0.5 * coarse[0])
a.gamma, a.courant_factor, a.slope_type, a.riemann = 1.4, 0.8, 1, "hllc"
base = bench.sedov_ic(0.5, 1, level)
fill_state(a, level, lambda x, y, z: base(np.mod(x, 1.0), np.mod(y, 1.0), np.mod(z, 1.0)))
u0 = a.uold.copy()
for ovl, res in ((0, 0), (1, 0), (1, 2), (1, 4), (1, 8), (1, 16), (0, 0), (1, 4)):
    os.environ["RGPU_OVERLAP"] = str(ovl)
    os.environ["RGPU_OVERLAP_RESERVE"] = str(res)
    a.uold[:, :] = u0
    h = HydroGPU(a, device=lr)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        _l.check(h.L.rgpu_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    buf = (C.c_ubyte * 128)(*uid.cpu().tolist())
    _l.check(h.L.rgpu_comm_init(world, rank, buf))
    h.bind_level(level)
    h.upload_state(level)
    h.level_steps(level, 5)
    ms = []
    for _ in range(3):
        h.synchronize(); dist.barrier(); torch.cuda.synchronize()
        h.level_steps(level, 10)
        t = torch.tensor([h.level_info(level).last_steps_ms / 10], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(t.item())
    h.finalize()
    if rank == 0:
        print(f"world={world} level={level} overlap={ovl} reserve={res}: ms/step {[round(m, 3) for m in ms]}", flush=True)
dist.destroy_process_group()
