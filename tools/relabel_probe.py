#!/usr/bin/env python
"""DESIGN TOOL: does STORING the rows of Z in degree order (hot rows adjacent in memory, whole 128-byte lines useful)
speed the F = 16 / F = 64 step up?  Same graph, ids relabelled by descending degree, same kernels, K = 10."""
import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P
from ppnp_b200.synth import rmat_adjacency


def relabel(ip, idx, new_of_old):
    """CSR of the same graph with vertex v renamed new_of_old[v] (rows and columns), canonical again."""
    n = ip.numel() - 1
    deg = ip[1:] - ip[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=ip.device), deg)
    keys = (new_of_old[rows] << 32) | new_of_old[idx.long()]
    del rows
    keys = torch.sort(keys).values
    r = keys >> 32
    cols = (keys & 0xFFFFFFFF).to(torch.int32)
    cnt = torch.bincount(r, minlength=n)
    out = torch.zeros(n + 1, dtype=torch.int64, device=ip.device)
    out[1:] = torch.cumsum(cnt, 0)
    return out, cols


def timed(graph, F, K=10, reps=3):
    n = graph.n
    H = torch.randn(n, F, device="cuda")
    Z, S = torch.empty_like(H), torch.empty_like(H)
    P.appnp_propagate(graph, H, K, 0.1, out=Z, scratch=S)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        P.appnp_propagate(graph, H, K, 0.1, out=Z, scratch=S)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / K


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "16m"
    n, raw, scale, F = {"16m": (16_000_000, 220_000_000, 24, 16), "2m": (2_000_000, 26_400_000, 21, 64)}[wl]
    dev = torch.device("cuda:0")
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    deg = ip[1:] - ip[:-1]
    labels = {"natural ids": None}
    order = torch.sort(deg, descending=True, stable=True).indices
    by_deg = torch.empty(n, dtype=torch.int64, device=dev); by_deg[order] = torch.arange(n, device=dev)
    labels["ids by descending degree"] = by_deg
    if os.environ.get("PROBE_RANDOM"):
        labels["random ids"] = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    # hot rows by degree; cold rows grouped by their highest-degree neighbour (in that neighbour's degree rank), so that
    # the cold columns a hub row gathers are CONSECUTIVE rows of Z
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    cand = (deg[idx.long()] << 32) | (0xFFFFFFFF - idx.long())
    best = torch.full((n,), -1, dtype=torch.int64, device=dev)
    best.scatter_reduce_(0, rows, cand, reduce="amax", include_self=True)
    del rows, cand
    hub = torch.where(best >= 0, 0xFFFFFFFF - (best & 0xFFFFFFFF), torch.arange(n, device=dev))
    hub_rank = torch.where(best >= 0, by_deg[hub], torch.full_like(hub, n))
    for T in (16, 64, 256):
        hot = deg >= T
        k1 = torch.where(hot, torch.zeros_like(hub_rank), hub_rank + 1)          # hot block first, then by the hub's rank
        k2 = by_deg                                                                 # inside a group: by own degree rank
        o = torch.sort(k1 * n + k2).indices                                          # (n + 1) * n < 2^63
        f = torch.empty(n, dtype=torch.int64, device=dev); f[o] = torch.arange(n, device=dev)
        labels[f"hot (deg >= {T}) by degree, cold grouped by their best hub"] = f
    for name, f in labels.items():
        a_ip, a_idx = (ip, idx) if f is None else relabel(ip, idx, f)
        ahat = P.csr_normalize(a_ip.to(torch.int32) if int(a_ip[-1]) < 2**31 else a_ip, a_idx)
        for kw in (dict(order="degree", idx16=True), dict(order="degree", idx16=True, rows_below=64)):
            g = P.PropagationGraph(ahat, **kw)
            print(f"{wl} F={F} | {name} | {kw}: {timed(g, F):.3f} ms / step", flush=True)
            del g
        del ahat
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
