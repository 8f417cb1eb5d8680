#!/usr/bin/env python
"""2-rank debug of the fused halo push: after one step, compare every halo row with its owner's row."""
import os
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppnp_b200 import dist as pd, _lib

n, raw, scale, F = 204_800, 3_000_000, 18, 16
indptr, cols, bounds = pd.rmat_shard(n, raw, scale, 0, dev, rank, world, batch=1 << 20)
dinv = pd.global_dinv(indptr, bounds, rank, world, dev)
topo = pd.build_shard_topology(indptr, cols, bounds, rank)
prop = pd.FusedPushPropagation(topo, dinv)
H, Z, S = prop.alloc(F, 3)
for b in (H, Z, S):
    b.fill_(float("nan"))
H[: topo.n_local].normal_()
prop._push_input(H)
torch.cuda.synchronize(); dist.barrier()


def check(buf, tag):
    # owner rows via NCCL all_gather (padded), compare with my halo region
    width = max(bounds[r + 1] - bounds[r] for r in range(world))
    mine = torch.zeros(width, F, device=dev); mine[: topo.n_local] = buf[: topo.n_local]
    allb = torch.empty(world * width, F, device=dev)
    dist.all_gather_into_tensor(allb, mine)
    hc = topo.halo_cols
    b = torch.tensor(bounds, device=dev)
    owner = torch.searchsorted(b, hc, right=True) - 1
    exp = allb[owner * width + (hc - b[owner])]
    got = buf[topo.n_local: topo.n_local + topo.n_halo]
    bad = ~(torch.isclose(got, exp, rtol=0, atol=0, equal_nan=False).all(dim=1))
    nb = int(bad.sum())
    msg = f"[rank {rank}] {tag}: halo rows {topo.n_halo}, mismatched {nb}"
    if nb:
        # are the bad rows split rows (fix-up) at their owner?  gather fix rows of all ranks
        fr = prop.sub.plan.fix_row.to(torch.int64) + bounds[rank]
        cnt = torch.tensor([fr.numel()], device=dev); cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        mx = int(max(c.item() for c in cnts))
        pad = torch.full((mx,), -1, dtype=torch.int64, device=dev); pad[: fr.numel()] = fr
        allfr = torch.empty(world * mx, dtype=torch.int64, device=dev); dist.all_gather_into_tensor(allfr, pad)
        isfix = torch.isin(hc[bad], allfr)
        nan_rows = int(torch.isnan(got[bad]).any(dim=1).sum())
        msg += f"; of those split-at-owner {int(isfix.sum())}, still NaN (never written) {nan_rows}; first bad global ids {hc[bad][:8].tolist()}"
    else:
        dist.all_gather([torch.zeros(1, device=dev) for _ in range(world)], torch.zeros(1, device=dev))
        dummy = torch.empty(world * 1, dtype=torch.int64, device=dev); dist.all_gather_into_tensor(dummy, torch.zeros(1, dtype=torch.int64, device=dev))
    print(msg, flush=True)


check(H, "input push")
prop._step(H, H, Z, 0.1, _lib.EPI_Z2Y, True, True)
prop._barrier(Z)
torch.cuda.synchronize(); dist.barrier()
check(Z, "after fused step 1")
prop._step(Z, H, S, 0.1, _lib.EPI_Y, False, True)
prop._barrier(S)
torch.cuda.synchronize(); dist.barrier()
check(S, "after fused step 2")
dist.destroy_process_group()
