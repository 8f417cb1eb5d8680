"""Graph standardisation (csrc/standardize.cu) on the config-4 R-MAT recipe:
raw one-directional draws with duplicates and self loops in, canonical symmetric loop-free largest
component out.  Device-resident ms (CUDA events), end to end with host arrays, algorithmic bytes against
the HBM peak, and the oracle's numpy restatement on the host cores as the CPU baseline (bounded sample).
One JSON line.  Prepared at the end of round 1; not yet run on a GPU."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_b200 as P  # noqa: E402
import ppnp_oracle as oracle  # noqa: E402  (CPU baseline / checker only)
from ppnp_b200 import _lib  # noqa: E402


def raw_graph(n, raw, scale, dev):
    """Raw directed draws (src, dst) of the R-MAT stream as an unsorted-inside-rows CSR with duplicates and loops."""
    keys = torch.empty(2 * raw, dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.ppnp_rmat_keys(0, scale, n, 0, raw, _lib.ptr(keys), _lib.current_stream()), "ppnp_rmat_keys")
    keys = keys[0::2]                       # even entries: (src << 32 | dst) as drawn; invalid draws are -1
    keys = keys[keys >= 0]
    src, dst = keys >> 32, keys & 0xffffffff
    loops = torch.randint(0, n, (n // 100,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    dup = src.numel() // 50
    src, dst = torch.cat([src, loops, src[:dup]]), torch.cat([dst, loops, dst[:dup]])
    order = torch.sort(src, stable=True).indices
    src, dst = src[order], dst[order]
    ip = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ip[1:] = torch.cumsum(torch.bincount(src, minlength=n), 0)
    return ip, dst.to(torch.int32)


def main():
    dev = torch.device("cuda:0")
    n, raw, scale = 2_000_000, 26_400_000, 21
    ip, idx = raw_graph(n, raw, scale, dev)
    nnz = idx.numel()
    for _ in range(2):
        out = P.graph_standardize(ip, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = P.graph_standardize(ip, idx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # end to end: host arrays in, host arrays out
    iph, idxh = ip.cpu().pin_memory(), idx.cpu().pin_memory()
    t0 = time.perf_counter()
    o = P.graph_standardize(iph.to(dev, non_blocking=True), idxh.to(dev, non_blocking=True))
    res = [t.cpu() for t in o]
    ms_e2e = (time.perf_counter() - t0) * 1e3
    n_keys = 2 * nnz
    # keys: 1 write + radix sort (8 passes of 8 bits, read + write each) + unique (read + write); + the compaction
    algo_bytes = 8 * n_keys * (1 + 16 + 2) + 12 * int(out[1].numel()) + 40 * n
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # CPU baseline on a bounded sample: the same recipe at 1/8 of the size
    ns, raws = n // 8, raw // 8
    ips, idxs = raw_graph(ns, raws, scale - 3, dev)
    t0 = time.perf_counter()
    want = oracle.standardize(ips.cpu().numpy(), idxs.cpu().numpy())
    cpu_s = time.perf_counter() - t0
    got = P.graph_standardize(ips, idxs)
    ok = all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(got, want))
    print(json.dumps({
        "what": "graph_standardize", "n": n, "stored_entries_in": nnz, "n_out": int(out[2].numel()), "nnz_out": int(out[1].numel()),
        "ms": ms, "entries_per_s": nnz / (ms * 1e-3), "e2e_ms": ms_e2e,
        "roofline": {"bound": "hbm", "algorithmic_bytes": algo_bytes, "achieved_GBps": algo_bytes / 1e9 / (ms * 1e-3),
                     "peak_GBps": peak, "frac": algo_bytes / 1e9 / (ms * 1e-3) / peak},
        "cpu_baseline": {"kind": "port", "cores": 1, "sample": f"numpy restatement on the same recipe at n = {ns}, {idxs.numel()} entries",
                         "seconds": cpu_s, "entries_per_s": idxs.numel() / cpu_s, "matches_gpu": bool(ok)}}))


if __name__ == "__main__":
    main()
