#!/usr/bin/env python
"""2-rank probe (torchrun): NCCL all_to_all / send-recv bandwidth vs peer-memory (symmetric memory)
gathers over NVLink.  Diagnostic for the halo exchange design."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows, F = 16_000_000, 16            # 1 GB
send = torch.randn(rows, F, device=dev)
recv = torch.empty_like(send)
per = rows // world
ms = timed(lambda: dist.all_to_all_single(recv, send, [per] * world, [per] * world))
if rank == 0:
    print(f"nccl all_to_all_single: {(world - 1) * per * F * 4 / 1e9 / (ms * 1e-3):.1f} GB/s per rank per direction ({ms:.2f} ms)", flush=True)
peer = (rank + 1) % world
src = (rank - 1) % world
def sr():
    ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, src)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
ms = timed(sr)
if rank == 0:
    print(f"nccl send/recv 1 GB: {rows * F * 4 / 1e9 / (ms * 1e-3):.1f} GB/s ({ms:.2f} ms)", flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((rows, F), dtype=torch.float32, device=dev)
    t.copy_(send)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    pbuf = hdl.get_buffer(peer, (rows, F), torch.float32)
    hdl.barrier()
    ms = timed(lambda: recv.copy_(pbuf))
    if rank == 0:
        print(f"symm_mem peer contiguous copy (pull): {rows * F * 4 / 1e9 / (ms * 1e-3):.1f} GB/s ({ms:.2f} ms)", flush=True)
    idx = torch.randperm(rows, device=dev)[: rows // 2].sort().values
    out = torch.empty(rows // 2, F, device=dev)
    ms = timed(lambda: torch.index_select(pbuf, 0, idx, out=out))
    if rank == 0:
        print(f"symm_mem peer row gather (64 B rows, sorted ids, 50% of rows): {rows // 2 * F * 4 / 1e9 / (ms * 1e-3):.1f} GB/s ({ms:.2f} ms)", flush=True)
    t256 = symm_mem.empty((rows // 4, 64), dtype=torch.float32, device=dev)
    h2 = symm_mem.rendezvous(t256, dist.group.WORLD)
    p2 = h2.get_buffer(peer, (rows // 4, 64), torch.float32)
    h2.barrier()
    idx2 = torch.randperm(rows // 4, device=dev)[: rows // 8].sort().values
    out2 = torch.empty(rows // 8, 64, device=dev)
    ms = timed(lambda: torch.index_select(p2, 0, idx2, out=out2))
    if rank == 0:
        print(f"symm_mem peer row gather (256 B rows): {rows // 8 * 64 * 4 / 1e9 / (ms * 1e-3):.1f} GB/s ({ms:.2f} ms)", flush=True)
    ms = timed(lambda: hdl.barrier())
    if rank == 0:
        print(f"symm_mem barrier: {ms * 1e3:.1f} us", flush=True)
    hdl.barrier()
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] symmetric memory probe failed: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
