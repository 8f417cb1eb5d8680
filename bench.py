#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: APPNP K=10 propagate throughput in edge*feature/s and the
fraction of the HBM roofline, beside the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input: APPNP K=10 forward on H
plus K=10 backward on the upstream gradient G (the same operator, A_hat is symmetric) = 20
fused SpMM+teleport launches.  One edge*feature = one multiply-add of a stored non-zero of A_hat
(self loops included) with one feature column for one propagation step.

Workloads (BASELINE.json configs):
  rmat2m   (N=1 default)  config 4: R-MAT n=2 000 000, ~50 M directed non-zeros, F=64
  rmat100m (N>1 default)  config 5: R-MAT n=100 000 000, ~2 B non-zeros, F=16, rows partitioned
                          over the N GPUs, halo exchange over NCCL each iteration (strong scaling)
  tiny                    a 20 k-node graph for plumbing checks
  pubmed_exact            config 2: exact PPNP on a PubMed-shape graph (n = 19 717): Pi built on the GPU, then
                          a step = Pi[idx] @ H forward + its adjoint through the tcgen05 bf16 gather-GEMM
  pubmed_batch            config 3: batch-main.py's per-batch gather/propagate on the compact top-k Pi, a step
                          = one sweep over batches of --batch-size random rows
Every line carries "parity": the results of the timed configuration checked against the CPU oracle after the
timed region (full-size fp64 C oracle on one GPU; on N GPUs the adjointness identity at full size plus the C
oracle on a smaller graph through the same partitioned code path).
--order auto picks the processing order of the edge stream by measurement before the warm-up (degree
order or an L2-blocked carved order, same results either way; config.order names the one used); the
default is the measured degree order.
One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same pass
through the public API with pinned HOST buffers, H2D/D2H copies inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "APPNP K=10 propagate edge*feature/s & % HBM roofline at 1/2/4/8 B200 vs CPU ref"
UNIT = "edge*feature/s"
ALPHA, KSTEPS = 0.1, 10

PUBMED = {"n": 19717, "nnz_a": 88648, "C": 3, "rows": 940, "k": 128}     # PubMed shape

WORKLOADS = {
    #            n            raw draws      scale  F
    "rmat2m": (2_000_000, 26_400_000, 21, 64),
    "rmat100m": (100_000_000, 1_050_000_000, 27, 16),
    "rmat16m": (16_000_000, 220_000_000, 24, 16),
    "rmatl2": (250_000, 13_200_000, 18, 64),     # same recipe, Z (64 MB) fits the L2: gather-rate probe
    "tiny": (20_000, 300_000, 15, 64),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(0.05)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if key in d:
            return d[key]["bytes_per_launch"], d[key]["source"]
    return None, None


def n1_reference(wl):
    """T_1 of this workload through the same code path (committed measurement, profiles/scaling.json)."""
    p = os.path.join(ROOT, "profiles", "scaling.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(wl)
    return None


def algorithmic_bytes_per_pass(n, nnz, F, value_free=True):
    """per SpMM+axpy step 4(n+1) [indptr] + 4 nnz [indices] (+ 4 nnz [values]) +
    12 n F [read Z once, read H, write Z'].  A pass = 2 x K steps; the value-free iteration still
    reads the stored values in the first step of each propagation."""
    per = 4 * (n + 1) + 4 * nnz + 12 * n * F
    steps = 2 * KSTEPS
    if value_free:
        return steps * per + 2 * 4 * nnz
    return steps * (per + 4 * nnz)


# ------------------------------------------------------------------------------ CPU reference arm
def host_graph(workload):
    """The same recipe on the host through the C oracle (checker / baseline only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnp_oracle as oracle
    n, raw, scale, F = WORKLOADS[workload]
    ip, idx = oracle.rmat_graph(n, raw, scale, seed=0)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    return oracle, n, F, oip, oidx, oval


def cpu_port_rate(oracle, n, F, oip, oidx, oval, k_steps, repeats):
    """edge*feature/s of the multi-threaded fp32 C port (oracle/ppnp_oracle.c) over `k_steps`
    propagation steps of the same graph, best of `repeats`."""
    import numpy as np
    H = np.random.RandomState(1).randn(n, F).astype(np.float32)
    val32 = oval.astype(np.float32)
    best = float("inf")
    for _ in range(repeats):
        t = time.perf_counter()
        oracle.c_appnp_f32(oip, oidx, val32, H, k_steps, ALPHA)
        best = min(best, time.perf_counter() - t)
    return len(oidx) * F * k_steps / best, best


def cpu_torch_sparse_rate(n, F, oip, oidx, oval, k_steps):
    """edge*feature/s of the literal recurrence with torch.sparse_csr_tensor @ dense on the host cores
    (the "torch sparse path" north_star mentions; SURVEY.md section 6 measured 4.84e9 on 8 cores)."""
    import warnings
    import numpy as np
    import torch
    warnings.filterwarnings("ignore", message=".*[Ss]parse.*")
    A = torch.sparse_csr_tensor(torch.from_numpy(np.ascontiguousarray(oip, dtype=np.int64)),
                                torch.from_numpy(np.ascontiguousarray(oidx, dtype=np.int64)),
                                torch.from_numpy(np.ascontiguousarray(oval, dtype=np.float32)), size=(n, n))
    H = torch.from_numpy(np.random.RandomState(1).randn(n, F).astype(np.float32))
    Z = H
    t = time.perf_counter()
    for _ in range(k_steps):
        Z = (1 - ALPHA) * (A @ Z) + ALPHA * H
    dt = time.perf_counter() - t
    return len(oidx) * F * k_steps / dt, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores (the oracle
    port: the reference is pure Python and has no K-step propagation to run).
    Loads nothing of the product (no CUDA library, no GPU work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; this arm uses every host core
    ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    __import__("__graft_entry__").build_oracle()
    ours_wl = args.workload or ("rmat2m" if args.gpus == 1 else "rmat100m")
    wl, same = ours_wl, True
    if ours_wl == "rmat100m":
        # building the 2 B-edge graph on the host takes longer than the whole bench may run: same recipe and
        # feature width at 1/6.25 of the nodes (config-5 scale model), said so in the line
        wl, same = "rmat16m", False
    if ours_wl in ("pubmed_exact", "pubmed_batch"):
        return run_reference_pubmed(args, ours_wl, ncpu)
    oracle, n, F, oip, oidx, oval = host_graph(wl)
    cores = oracle.clib().oracle_num_threads()
    k_sample = 2
    warm = args.warmup if args.warmup is not None else 1
    for _ in range(max(1, warm)):
        cpu_port_rate(oracle, n, F, oip, oidx, oval, 1, 1)
    steps = args.steps or 3
    rates, t_total = [], 0.0
    for _ in range(steps):
        r, t = cpu_port_rate(oracle, n, F, oip, oidx, oval, k_sample, 1)
        rates.append(r)
        t_total += t
    value = sum(rates) / len(rates)
    sample = (f"each bench step = {k_sample} of the 20 propagation steps of one pass on {wl} (nnz(A_hat)={len(oidx)}, F={F}), "
              f"fp32 OpenMP C port (oracle/ppnp_oracle.c), {cores} threads")
    ts_rate, ts_threads = cpu_torch_sparse_rate(n, F, oip, oidx, oval, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "n": n, "nnz_a_hat": int(len(oidx)), "F": F, "K": KSTEPS, "alpha": ALPHA,
                   "pass": f"{k_sample} propagation steps per bench step (a sample of the K=10 forward + K=10 backward pass; "
                           "ms_per_step is the sample's own time, nothing is extrapolated)",
                   "same_workload": same, "gpu_arm_workload": ours_wl},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "torch_sparse_csr": {"value": ts_rate, "threads": ts_threads, "sample": "1 step"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not same:
        line["same_workload"] = False
    print(json.dumps(line))


def run_reference_pubmed(args, wl, ncpu):
    """CPU arm of configs 2/3: the reference's own dense lines (model.py:63 / batch-main.py:140-146) with numpy /
    torch on the host cores, on the same synthetic PubMed-shape Pi (built by the fp64 oracle)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnp_oracle as oracle
    import torch
    torch.set_num_threads(ncpu)
    n, C = PUBMED["n"], PUBMED["C"]
    rng = np.random.RandomState(0)
    # a bounded sample: Pi rows for the rows the step touches come from a random dense fp32 matrix of the same
    # shape (the arithmetic and the bytes are what is timed; the values do not matter for the rate)
    m = n if wl == "pubmed_exact" else args.batch_size
    Pi = torch.from_numpy(rng.rand(n, n).astype(np.float32))
    if wl == "pubmed_batch":
        Pi[Pi < 1.0 - PUBMED["k"] / n] = 0
    H = torch.from_numpy(rng.randn(n, C).astype(np.float32))
    idx = torch.from_numpy(rng.permutation(n)[:m].astype(np.int64))
    steps = args.steps or 5
    warm = args.warmup if args.warmup is not None else 1

    def step():
        if wl == "pubmed_exact":
            return Pi @ H                       # model.py:65 over all rows, like the GPU arm
        sub = Pi[idx]
        sel = (sub > 0).any(dim=0)              # batch-main.py:141
        return sub[:, sel] @ H[sel]             # batch-main.py:142-146
    for _ in range(max(1, warm)):
        step()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / steps
    work = m * n * C if wl == "pubmed_exact" else int((Pi[idx] > 0).sum()) * C
    value = work / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "n": n, "C": C, "rows": m, "edge": "one stored entry of the dense Pi rows the step reads"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncpu, "kind": "port",
                             "sample": "the reference's dense torch lines on the host, random Pi of the same shape"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- parity / helpers
def parity_single(ip, idx, H, G, Z, dH, sample_rows=4096):
    """Full-size check of the timed results against the fp64 C oracle on the host (checker only)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnp_oracle as oracle
    t0 = time.perf_counter()
    oip, oidx, oval, _ = oracle.c_a_hat(ip.cpu().numpy().astype(np.int64), idx.cpu().numpy(), None, "sym")
    rows = np.random.RandomState(7).choice(H.shape[0], size=min(sample_rows, H.shape[0]), replace=False)
    out = {"oracle": "oracle/ppnp_oracle.c (fp64), K=%d alpha=%g, all %d rows" % (KSTEPS, ALPHA, H.shape[0]), "tol": 1e-5}
    ok = True
    for name, X, Y in (("fwd", H, Z), ("bwd", G, dH)):
        ref = oracle.c_appnp_f64(oip, oidx, oval, X.cpu().numpy().astype(np.float64), KSTEPS, ALPHA)
        got = Y.cpu().numpy().astype(np.float64)
        rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        rel_s = float(np.linalg.norm(got[rows] - ref[rows]) / np.linalg.norm(ref[rows]))
        mx = float(np.abs(got - ref).max() / np.abs(ref).max())
        agree = float((got.argmax(1) == ref.argmax(1)).mean())
        out[name] = {"rel_fro": rel, "rel_fro_%d_sampled_rows" % len(rows): rel_s, "max_abs_over_max": mx, "argmax_agreement": agree}
        ok = ok and rel < 1e-5 and rel_s < 1e-5
        del ref, got
    # adjointness: <P(H), G> == <H, P(G)>
    lhs = float((Z.double() * G.double()).sum())
    rhs = float((H.double() * dH.double()).sum())
    out["adjointness_rel"] = abs(lhs - rhs) / max(abs(lhs), 1e-30)
    out["ok"] = bool(ok and out["adjointness_rel"] < 1e-5)
    out["seconds"] = round(time.perf_counter() - t0, 1)
    return out


def parity_small_partitioned(make_prop, alloc_of, dev, rank, world, F, K, alpha):
    """The same partitioned code path (same transport, same kernels) on a graph the fp64 C oracle finishes in a
    second: R-MAT n = 204 800 (tests/test_gpu_dist.py uses the same one), every rank checks its own rows."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ppnp_b200.dist import auto_stripes, build_shard_topology, global_dinv, rmat_shard, stripe_relabel
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnp_oracle as oracle       # checker only
    n, raw, scale = 204_800, 3_000_000, 18
    indptr, cols, bounds, relabel = rmat_shard(n, raw, scale, 0, dev, rank, world, batch=1 << 20, return_relabel=True)
    dinv = global_dinv(indptr, bounds, rank, world, dev)
    topo = build_shard_topology(indptr, cols, bounds, rank)
    pr = make_prop(topo, dinv, None)
    H, Z, S = alloc_of(pr, 3)
    new_of_old = relabel.cpu().numpy()         # generator ids -> the partitioner's ids (stripes, then degree order inside a block)
    mine = np.argsort(new_of_old)[bounds[rank]: bounds[rank + 1]]
    Hg = np.random.RandomState(0).randn(n, F).astype(np.float32)
    for b in (H, Z, S):
        b.zero_()
    H[: topo.n_local] = torch.from_numpy(Hg[mine]).to(dev)
    out = pr.propagate(H, Z, S, K, alpha).cpu().numpy()
    ip, idx = oracle.rmat_graph(n, raw, scale, seed=0)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    ref = oracle.c_appnp_f64(oip, oidx, oval, Hg.astype(np.float64), K, alpha)[mine]
    err = torch.tensor([float(np.linalg.norm(out - ref) / np.linalg.norm(ref))], device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    return {"n": n, "F": F, "K": K, "rel_fro_max_over_ranks": float(err), "tol": 1e-5,
            "oracle": "oracle/ppnp_oracle.c fp64, same recipe as the timed graph"}


def sub_bench(argv, gpu_index):
    """Run this script once more in a fresh process on one GPU and return its JSON line (or the error)."""
    import subprocess
    # a clean single-process environment: everything torch.distributed.run exported to this rank must go (with
    # TORCHELASTIC_USE_AGENT_STORE left set the child would wait for the parent job's store forever)
    env = {k: v for k, v in os.environ.items()
           if not (k.startswith(("TORCHELASTIC_", "GROUP_", "ROLE_", "TORCH_NCCL_", "NCCL_ASYNC")) or
                   k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "MASTER_PORT", "MASTER_ADDR", "OMP_NUM_THREADS"))}
    env["MASTER_ADDR"] = "127.0.0.1"
    env["MASTER_PORT"] = str(29600 + (os.getpid() % 300))
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    env["CUDA_VISIBLE_DEVICES"] = (vis.split(",")[gpu_index] if vis else str(gpu_index))
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=300)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                return {"ms_per_step": d.get("ms_per_step"), "value": d.get("value"), "unit": d.get("unit"),
                        "workload": d.get("config", {}).get("workload"), "n_gpus": d.get("n_gpus"),
                        "how": "python bench.py " + " ".join(argv) + " (same box, same run)"}
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def _timed_steps(fn, steps, warmup, local):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return sum(per) / steps, per, clocks


def run_pubmed(args, wl, dev, local):
    """BASELINE.json configs 2 and 3 on one B200: synthetic PubMed-shape graph, Pi built on the
    GPU (Chebyshev-accelerated multi-RHS iteration), then the timed step:
      pubmed_exact  logits = Pi @ H over all n rows through the tcgen05 bf16 gather-GEMM (model.py:65), flush-free
                    because Pi (0.78 GB) is larger than the L2;
      pubmed_batch  one sweep of 16 batches of --batch-size random rows: support union + propagate on the compact
                    top-k Pi (batch-main.py:113-117, 140-146)."""
    import numpy as np
    import torch
    import ppnp_b200 as P
    from ppnp_b200.synth import powerlaw_adjacency
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnp_oracle as oracle   # checker only
    import scipy.sparse as sp
    n, C = PUBMED["n"], PUBMED["C"]
    steps = args.steps if args.steps is not None else 20
    warmup = args.warmup if args.warmup is not None else 3
    peak, peak_src = measured_peaks()
    ip, idx = powerlaw_adjacency(n, PUBMED["nnz_a"], seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    Kc = P.ppr_cheb_steps_for_tol(ALPHA, 1e-7)
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Pi = P.ppr_dense(ahat, ALPHA, K=Kc, method="chebyshev")       # warm-up / result
    a_.record()
    Pi = P.ppr_dense(ahat, ALPHA, K=Kc, method="chebyshev")
    b_.record()
    torch.cuda.synchronize()
    build_ms = a_.elapsed_time(b_)
    build = {"method": "chebyshev", "K": Kc, "ms": build_ms, "algorithmic_bytes": 3 * n * n * 4 * Kc,
             "frac_of_hbm_peak": 3 * n * n * 4 * Kc / 1e9 / (build_ms * 1e-3) / peak}
    # oracle: rows of Pi = alpha (I - (1-alpha) A_hat)^-1 (helpers.py:68-71) from the fp64 restatement, KAT-2 of SURVEY 8c
    # (H = unit vectors reproduces columns of Pi; Pi is symmetric), iterated to round-off
    adj = sp.csr_matrix((np.ones(int(ip[-1]), np.float32), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))
    A64 = oracle.calc_A_hat(adj, "sym")
    probe = np.random.RandomState(5).choice(n, 16, replace=False)
    E = np.zeros((n, len(probe)))
    E[probe, np.arange(len(probe))] = 1.0
    cols_ref = oracle.appnp(A64, E, ALPHA, 400)                 # (1-alpha)^400 ~ 5e-19
    got = Pi[torch.from_numpy(probe).to(dev)].cpu().numpy().astype(np.float64).T
    pi_err = float(np.linalg.norm(got - cols_ref) / np.linalg.norm(cols_ref))
    g = torch.Generator(device=dev).manual_seed(1)
    H = torch.randn(n, C, device=dev, generator=g)
    Hh = torch.empty((n, C), dtype=torch.float32, pin_memory=True).copy_(H)
    Hd = torch.empty_like(H)

    if wl == "pubmed_exact":
        Pb = P.to_bf16_padded(Pi)
        step = lambda: P.gather_gemm_bf16(Pb, H, None)
        ms, per, clocks = _timed_steps(step, steps, warmup, local)
        out = step()
        ref_rows = cols_ref.T @ H.cpu().numpy().astype(np.float64)          # oracle logits of the probed rows
        lg_err = float(np.linalg.norm(out[torch.from_numpy(probe).to(dev)].cpu().numpy() - ref_rows) / np.linalg.norm(ref_rows))
        out32 = P.gather_gemm(Pi, H, None)
        lg32 = float(np.linalg.norm(out32[torch.from_numpy(probe).to(dev)].cpu().numpy() - ref_rows) / np.linalg.norm(ref_rows))
        work = n * n * C
        algo = n * n * 2 + 2 * n * C * 4
        outh = torch.empty((n, C), dtype=torch.float32, pin_memory=True)

        def e2e_step():
            Hd.copy_(Hh, non_blocking=True)
            outh.copy_(P.gather_gemm_bf16(Pb, Hd, None), non_blocking=True)
        ms_e2e, _, _ = _timed_steps(e2e_step, max(3, steps // 2), 2, local)
        launches = 3 * steps      # H pack + tcgen05 gather-GEMM + slice reduction per apply
        cpu = None
        if not args.no_cpu_baseline:
            Pi_h = Pi.cpu().numpy()
            H_h = H.cpu().numpy()
            m = 4096
            t0 = time.perf_counter()
            for _ in range(3):
                Pi_h[:m] @ H_h
            dt = (time.perf_counter() - t0) / 3
            cpu = {"value": m * n * C / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"model.py:65 with numpy fp32 on the host: the first {m} rows of the same Pi @ H"}
        extra = {"pi_build": build, "fp32_path_rel_err_vs_oracle_rows": lg32}
        for N in (16, 64):
            HN = torch.randn(n, N, device=dev, generator=g)
            t, _, _ = _timed_steps(lambda: P.gather_gemm_bf16(Pb, HN, None), 10, 2, local)
            extra[f"apply_N{N}_ms"] = t
            extra[f"apply_N{N}_frac_of_hbm_peak"] = n * n * 2 / 1e9 / (t * 1e-3) / peak
        parity = {"pi_rows_vs_oracle_rel_fro": pi_err, "logits_bf16_vs_oracle_rel_fro": lg_err, "tol_pi": 1e-5, "tol_bf16": 1e-2,
                  "oracle": "oracle/ppnp_oracle.py fp64 (helpers.py:58-71 restated; 16 probed rows)",
                  "ok": bool(pi_err < 1e-5 and lg_err < 1e-2 and lg32 < 1e-5)}
        config = {"workload": wl, "n": n, "nnz_a": int(ip[-1]), "C": C, "rows": n, "alpha": ALPHA,
                  "edge": "one stored entry of the dense Pi", "pass": "logits = Pi @ H, bf16 operands, fp32 accumulate in TMEM",
                  "l2": "Pi (0.78 GB bf16) larger than L2"}
        kernel, dtype = "gather_gemm_tc_kernel", "bf16"
        h2d, d2h = n * C * 4, n * C * 4
    else:
        k, B, nb = PUBMED["k"], args.batch_size, 16
        P.topk_sparsify_(Pi, k)
        spp = P.dense_to_sparse_ppr(Pi)
        batches = [torch.randperm(n, device=dev, generator=g)[:B].sort().values for _ in range(nb)]

        def step():
            for ib in batches:
                sel = P.batch_support(spp, ib)
                P.batch_propagate(spp, ib, sel, H[sel])
        ms, per, clocks = _timed_steps(step, steps, warmup, local)
        kept = sum(int((spp.indptr[ib + 1] - spp.indptr[ib]).sum()) for ib in batches)
        work = kept * C
        algo = kept * 8 + sum(int(P.batch_support(spp, ib).sum()) for ib in batches) * C * 4 + nb * B * C * 4
        # parity: the literal batch lines (batch-main.py:140-146) in the oracle on the host, first batch
        ib = batches[0]
        sel = P.batch_support(spp, ib)
        ours = P.batch_propagate(spp, ib, sel, H[sel]).cpu().numpy()
        lit = oracle.batch_step(Pi.cpu().numpy(), ib.cpu().numpy(), H.cpu().numpy())
        lit_logits = lit[0] if isinstance(lit, tuple) else lit
        b_err = float(np.linalg.norm(ours - lit_logits) / np.linalg.norm(lit_logits))
        outh = torch.empty((B, C), dtype=torch.float32, pin_memory=True)

        def e2e_step():
            Hd.copy_(Hh, non_blocking=True)
            for ib_ in batches:
                s_ = P.batch_support(spp, ib_)
                outh.copy_(P.batch_propagate(spp, ib_, s_, Hd[s_]), non_blocking=True)
        ms_e2e, _, _ = _timed_steps(e2e_step, max(3, steps // 2), 2, local)
        launches = 3 * nb * steps
        cpu = None
        if not args.no_cpu_baseline:
            Pi_h, H_h = Pi.cpu(), H.cpu()
            t0 = time.perf_counter()
            for ib_ in batches[:4]:
                sub = Pi_h[ib_.cpu()]
                s_ = (sub > 0).any(dim=0)
                sub[:, s_] @ H_h[s_]
            dt = (time.perf_counter() - t0) / 4 * nb
            cpu = {"value": work / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": "batch-main.py:140-146 with torch on the host, 4 of the 16 batches"}
        extra = {"pi_build": build, "kept_entries": int(spp.indices.numel()), "batches_per_s": nb / (ms * 1e-3)}
        parity = {"pi_rows_vs_oracle_rel_fro": pi_err, "batch_logits_vs_oracle_rel_fro": b_err, "tol": 1e-5,
                  "oracle": "oracle/ppnp_oracle.py batch_step (batch-main.py:140-146 restated)", "ok": bool(pi_err < 1e-5 and b_err < 1e-5)}
        config = {"workload": wl, "n": n, "nnz_a": int(ip[-1]), "C": C, "k": k, "batch_size": B, "batches_per_step": nb,
                  "edge": "one kept entry of the top-k Pi rows of a batch", "pass": "support union + compact gather/propagate per batch",
                  "l2": "latency-bound: working set fits the L2"}
        kernel, dtype = "batch_propagate_kernel", "f32"
        h2d, d2h = n * C * 4, nb * B * C * 4
    achieved = algo / (ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": work / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_step": algo},
            "cpu_baseline": cpu,
            "e2e": {"value": work / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "ms_per_step_minmax": [min(per), max(per)], "parity": parity, "extra": extra}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = args.workload or ("rmat2m" if world == 1 else "rmat100m")
    if wl in ("pubmed_exact", "pubmed_batch"):
        if world != 1:
            raise SystemExit("the PubMed-shape workloads run on one GPU (replicas only: independent batches / row blocks)")
        __import__("__graft_entry__").build()
        return run_pubmed(args, wl, dev, local)
    # config 5 family: always through the partitioned code path, also at N=1, so that T_1 and T_P of the
    # scaling study come from the same code
    partitioned = world > 1 or wl in ("rmat100m", "rmat16m")
    if args.rows_below is None:
        args.rows_below = 64 if partitioned else 0
    if partitioned:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    __import__("__graft_entry__").build() if rank == 0 else None
    if partitioned:
        dist.barrier()
    import ppnp_b200 as P
    from ppnp_b200.synth import rmat_adjacency

    n, raw, scale, F = WORKLOADS[wl]
    steps = args.steps if args.steps is not None else 10
    warmup = args.warmup if args.warmup is not None else 3
    peak, peak_src = measured_peaks()

    if partitioned:
        from ppnp_b200 import dist as pd
        result = pd.bench_partitioned(wl, n, raw, scale, F, KSTEPS, ALPHA, steps, warmup, dev, rank, world,
                                       phases=args.phases, transport=args.transport, stripes=args.stripes,
                                       row_groups=args.row_groups, hub_degree=args.hub_degree, idx16=args.dist_idx16,
                                       carve=({"block_cols": args.carve_block_cols, "n_blocks": args.carve_blocks,
                                               "min_piece": args.carve_min_piece} if args.order == "carve" else None),
                                       rows_below=(args.rows_below or None), rows_order=args.rows_order,
                                       window=(args.window_key if args.order == "window" else None),
                                       degree_sort=not args.no_degree_sort, row_cost=args.row_cost,
                                       check_small=(None if args.no_parity else
                                                    (lambda mk, al: parity_small_partitioned(mk, al, dev, rank, world, F, KSTEPS, ALPHA))))
        if rank == 0:
            sampler_clocks = result.pop("clocks")
            nnz = result.pop("nnz")
            ms = result.pop("ms_per_step")
            work = 2 * KSTEPS * nnz * F
            value = work / (ms * 1e-3)
            bytes_pass = algorithmic_bytes_per_pass(n, nnz, F)
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": wl, "n": n, "nnz_a_hat": nnz, "F": F, "K": KSTEPS, "alpha": ALPHA,
                           "pass": "K=10 forward + K=10 backward", "partition": result.pop("partition"),
                           "l2": "inputs larger than L2"},
                "roofline": {"bound": "hbm", "achieved": bytes_pass / (ms * 1e-3) / 1e9, "peak": peak * world,
                             "unit": "GB/s", "frac": bytes_pass / (ms * 1e-3) / 1e9 / (peak * world), "traffic": None,
                             "peak_source": peak_src + f" x {world} GPUs", "kernel": "spmm_stream_kernel"},
                "e2e": result.pop("e2e"), "gpu_launches": result.pop("gpu_launches"), "clocks": sampler_clocks,
                "parity": result.pop("parity"),
                "extra": dict(result, n1_committed=n1_reference(wl)),
            }
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            if world > 1 and not args.no_extras:
                # T_1 of the SAME workload through the same code path, measured now on this rank's GPU (the other
                # ranks have left): the parallel efficiency below uses nothing but numbers of this run
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                t1 = sub_bench(["--gpus", "1", "--workload", wl, "--steps", "2", "--warmup", "1", "--no-cpu-baseline",
                                "--no-extras", "--no-parity"], local)
                line["extra"]["t1_live"] = t1
                if t1.get("ms_per_step"):
                    line["extra"]["efficiency"] = {"value": t1["ms_per_step"] / (world * ms), "formula": "T_1 / (N * T_N), strong scaling, "
                                                   "T_1 measured in this run on one GPU of this box (extra.t1_live)"}
            print(json.dumps(line))
        return

    # ---------------------------------------------------------------- single GPU: config 4
    t0 = time.perf_counter()
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    carve = None
    if args.order == "carve":
        carve = {"block_cols": args.carve_block_cols, "n_blocks": args.carve_blocks, "min_piece": args.carve_min_piece,
                 "wide_cta": not args.carve_narrow_cta, "interleave": args.carve_interleave}
        if args.carve_levels:      # "512x64x8,125000x16x16": block_cols x n_blocks x min_piece per level
            carve["levels"] = [tuple(int(v) for v in lv.split("x")) for lv in args.carve_levels.split(",")]
    order_tried = None
    if args.order == "auto":
        # processing orders that give the same results (parity-tested): time one propagation each and keep the
        # fastest.  Degree order is the measured default; the others stream the hub rows' cold columns in
        # L2-sized blocks.  A candidate that fails to build or run is skipped.
        bc = max(1, n // 16)
        cands = [("degree", dict(order="degree")),
                 ("carve-l2 %dx16 min 16" % bc, dict(order="carve", carve=dict(levels=[(bc, 16, 16)], wide_cta=False))),
                 ("carve-l2 %dx16 min 32" % bc, dict(order="carve", carve=dict(levels=[(bc, 16, 32)], wide_cta=False)))]
        order_tried, best = {}, None
        Hp = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        Zp, Sp = torch.empty_like(Hp), torch.empty_like(Hp)
        for name, kw in cands:
            try:
                gph = P.PropagationGraph(ahat, chunk_edges=args.chunk_edges, idx16=args.idx16, **kw)
                P.appnp_propagate(gph, Hp, KSTEPS, ALPHA, use_vals=args.use_vals, out=Zp, scratch=Sp)
                torch.cuda.synchronize()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                for _ in range(2):
                    P.appnp_propagate(gph, Hp, KSTEPS, ALPHA, use_vals=args.use_vals, out=Zp, scratch=Sp)
                b_.record()
                torch.cuda.synchronize()
                order_tried[name] = a_.elapsed_time(b_) / (2 * KSTEPS)
                if best is None or order_tried[name] < best[1]:
                    best = (name, order_tried[name], gph)
                del gph
            except Exception as e:  # noqa: BLE001
                if name == "degree":
                    raise
                order_tried[name] = "skipped: " + repr(e)[:200]
        del Hp, Zp, Sp
        graph = best[2]
        args.order = best[0]
        del best
        torch.cuda.empty_cache()
    else:
        tiled = None
        if args.tiled:
            tiled = {"slice_width": args.tiled_slice, "slack": args.tiled_slack, "fine_cols": args.tiled_fine_cols,
                     "rest": "rows" if args.rows_below else "stream"}
        graph = P.PropagationGraph(ahat, chunk_edges=args.chunk_edges, order=args.order, idx16=args.idx16, carve=carve, tiled=tiled,
                                   rows_below=(args.rows_below or None) if not args.tiled else None,
                                   window=({"key": args.window_key, "wide_cta": args.window_wide} if args.order == "window" else None))
    nnz = ahat.nnz
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    g = torch.Generator(device=dev).manual_seed(1)
    H = torch.randn(n, F, device=dev, generator=g)
    G = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    Z, dH, scratch = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)

    def one_pass():
        P.appnp_propagate(graph, H, KSTEPS, ALPHA, use_vals=args.use_vals, out=Z, scratch=scratch)
        P.appnp_propagate(graph, G, KSTEPS, ALPHA, use_vals=args.use_vals, out=dH, scratch=scratch)

    for _ in range(warmup):
        one_pass()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        one_pass()
        ev[i + 1].record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    ms = sum(per) / steps
    work = 2 * KSTEPS * nnz * F
    value = work / (ms * 1e-3)
    bytes_pass = algorithmic_bytes_per_pass(n, nnz, F, value_free=not args.use_vals)
    achieved = bytes_pass / (ms * 1e-3) / 1e9
    launches_per_pass = 2 * KSTEPS * (2 if (graph.plan is not None and graph.plan.n_fix > 0) else 1)
    traffic, traffic_src = measured_traffic(f"{wl}/{args.order}/{'stored-values' if args.use_vals else 'value-free'}")

    # ---- e2e: host buffers through the public API, copies inside the timed region
    e2e_steps = max(4, steps)       # as many passes as the device-timed region: fill and drain of the copy pipeline amortise alike
    Hh = torch.empty((n, F), dtype=torch.float32, pin_memory=True).copy_(H)
    Gh = torch.empty((n, F), dtype=torch.float32, pin_memory=True).copy_(G)
    Zh = torch.empty((n, F), dtype=torch.float32, pin_memory=True)
    dHh = torch.empty((n, F), dtype=torch.float32, pin_memory=True)
    Hd, Gd = torch.empty_like(H), torch.empty_like(G)

    # copy-in, compute and copy-out on three streams: the H2D of G rides under the forward propagation,
    # the D2H of Z under the backward one, and consecutive passes pipeline (PCIe is full duplex)
    s_in, s_out, cur = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    last = {"fwd": None, "bwd": None, "z": None, "dh": None}

    def one_pass_e2e():
        with torch.cuda.stream(s_in):
            if last["fwd"] is not None:
                s_in.wait_event(last["fwd"])          # Hd is free once the previous forward has consumed it
            Hd.copy_(Hh, non_blocking=True)
            eH = torch.cuda.Event(); eH.record(s_in)
            if last["bwd"] is not None:
                s_in.wait_event(last["bwd"])
            Gd.copy_(Gh, non_blocking=True)
            eG = torch.cuda.Event(); eG.record(s_in)
        cur.wait_event(eH)
        if last["z"] is not None:
            cur.wait_event(last["z"])                 # Z is free once its previous read-back has finished
        P.appnp_propagate(graph, Hd, KSTEPS, ALPHA, use_vals=args.use_vals, out=Z, scratch=scratch)
        f = torch.cuda.Event(); f.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(f)
            Zh.copy_(Z, non_blocking=True)
            zo = torch.cuda.Event(); zo.record(s_out)
        cur.wait_event(eG)
        if last["dh"] is not None:
            cur.wait_event(last["dh"])
        P.appnp_propagate(graph, Gd, KSTEPS, ALPHA, use_vals=args.use_vals, out=dH, scratch=scratch)
        b = torch.cuda.Event(); b.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(b)
            dHh.copy_(dH, non_blocking=True)
            do = torch.cuda.Event(); do.record(s_out)
        last.update(fwd=f, bwd=b, z=zo, dh=do)

    def drain():
        cur.wait_stream(s_in)
        cur.wait_stream(s_out)

    one_pass_e2e()
    drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    e0.record()
    for _ in range(e2e_steps):
        one_pass_e2e()
    drain()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / e2e_steps
    assert torch.equal(Zh, Z.cpu()) and torch.equal(dHh, dH.cpu())      # the results did reach the host

    # ---- parity of exactly what was timed: Z = P(H) and dH = P(G) of the last pass against the fp64 C oracle
    # (oracle/ppnp_oracle.c: calc_A_hat restated from helpers.py:58-63 + the K-step recurrence), every row
    parity = None
    if not args.no_parity:
        parity = parity_single(ip, idx, H, G, Z, dH)

    if graph.rows_part is not None:
        launches_per_pass = 2 * KSTEPS * ((0 if graph.plan is None else (2 if graph.plan.n_fix > 0 else 1)) + 1)
    tl = graph.tiled_for(F)
    if tl is not None:
        launches_per_pass = 2 * KSTEPS * (1 + (0 if tl[1] is None else (2 if tl[1].n_fix > 0 else 1)) + (0 if tl[3] is None else 1))

    extra = {}
    if wl == "rmat2m" and not args.no_extras:
        # T_1 of the multi-GPU workload (config 5) through the partitioned code path, measured in this run on this
        # box: the denominator of the parallel efficiency of the N > 1 lines
        del Hh, Gh, Zh, dHh, Hd, Gd
        torch.cuda.empty_cache()
        extra["rmat100m_n1"] = sub_bench(["--gpus", "1", "--workload", "rmat100m", "--steps", "2", "--warmup", "1",
                                          "--no-cpu-baseline", "--no-extras", "--no-parity"], local)

    # ---- CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ppnp_oracle as oracle
        oip = ahat.indptr.cpu().numpy().astype("int64")
        oidx = ahat.indices.cpu().numpy()
        oval = ahat.val32.cpu().numpy()
        cores = oracle.clib().oracle_num_threads()
        k_sample = 2 if n >= 1_000_000 else 10
        rate, secs = cpu_port_rate(oracle, n, F, oip, oidx, oval, k_sample, 2)
        ts_rate, ts_threads = cpu_torch_sparse_rate(n, F, oip, oidx, oval, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "torch_sparse_csr": {"value": ts_rate, "threads": ts_threads, "sample": "1 step"},
               "sample": f"{k_sample} of the 20 propagation steps of one pass, same graph and F, fp32 OpenMP C port "
                         f"(oracle/ppnp_oracle.c), best of 2 ({secs:.2f} s)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl, "n": n, "nnz_a": int(ip[-1]), "nnz_a_hat": nnz, "F": F, "K": KSTEPS, "alpha": ALPHA,
                   "pass": "K=10 forward + K=10 backward = 20 fused SpMM+teleport launches",
                   "form": "stored values" if args.use_vals else "value-free Y-space (stored values in step 1)",
                   "order": args.order, "chunk_edges": args.chunk_edges, "l2": "inputs larger than L2 (3 x 512 MB)",
                   "idx16": bool(args.idx16), "carve": None if graph.plan is None else graph.plan.carve,
                   "rows_below": args.rows_below or None,
                   "order_candidates_ms_per_launch": order_tried,
                   "graph_build_s": round(t_build, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "spmm_stream_kernel", "algorithmic_bytes_per_launch": bytes_pass / (2 * KSTEPS),
                     "launch_ms": ms / (2 * KSTEPS)},
        "cpu_baseline": cpu,
        "e2e": {"value": work / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n * F * 4, "d2h_bytes_per_step": 2 * n * F * 4},
        "gpu_launches": launches_per_pass * steps,
        "clocks": clocks,
        "ms_per_step_minmax": [min(per), max(per)],
        "parity": parity,
        "extra": extra,
    }
    if tl is not None:
        line["config"]["tiled"] = dict(tl[0].stats, slice_width=tl[2])
        line["roofline"]["kernel"] = "spmm_tiled_kernel + spmm_stream_kernel"
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["pubmed_exact", "pubmed_batch"])
    ap.add_argument("--order", default="degree", choices=["auto", "natural", "degree", "carve", "window"],
                    help="processing order of the edge stream; auto times degree order and two L2-blocked carves and keeps the fastest")
    ap.add_argument("--idx16", dest="idx16", action="store_true", default=True,
                    help="16-byte staging of a lane-transposed index stream (default; bit-identical results)")
    ap.add_argument("--no-idx16", dest="idx16", action="store_false", help="4-byte staging of the linear index stream")
    ap.add_argument("--carve-block-cols", type=int, default=512, help="--order carve: columns per L1-sized block")
    ap.add_argument("--carve-blocks", type=int, default=64, help="--order carve: number of hot column blocks")
    ap.add_argument("--carve-min-piece", type=int, default=4, help="--order carve: smallest (row, block) piece taken out of its row")
    ap.add_argument("--carve-narrow-cta", action="store_true", help="--order carve: keep the 256-thread CTAs")
    ap.add_argument("--carve-interleave", action="store_true", help="--order carve: alternate carved and residual chunk units")
    ap.add_argument("--carve-levels", default=None, help="--order carve: stacked block levels, e.g. 512x64x8,125000x16x16")
    ap.add_argument("--window-key", default="mid", choices=["first", "mid", "last"], help="--order window: column of a chunk whose rank orders the chunks")
    ap.add_argument("--window-wide", action="store_true", help="--order window: 1024-thread CTAs (64 consecutive chunks per SM)")
    ap.add_argument("--chunk-edges", type=int, default=256)
    ap.add_argument("--use-vals", action="store_true", help="stored-value form in every step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed results")
    ap.add_argument("--no-extras", action="store_true", help="skip the companion measurements (T_1 of config 5 in the N=1 run, live T_1 in the N>1 runs)")
    ap.add_argument("--tiled", dest="tiled", action="store_true", default=False,
                    help="hub rows through the shared-memory-resident kernel (csrc/appnp_tiled.cu)")
    ap.add_argument("--no-tiled", dest="tiled", action="store_false")
    ap.add_argument("--rows-below", type=int, default=None,
                    help="rows with fewer stored entries go through the one-lane-group-per-row kernel (csrc/appnp_rows.cu); 0 = off. "
                         "Default: 0 for config 4 (measured equal), 64 for the partitioned config-5 family (measured -3.7 %% per pass "
                         "at 8 GPUs: the rows kernel's epilogue pushes whole warps of finished rows at once)")
    ap.add_argument("--rows-order", default="dest", choices=["dest", "degree"],
                    help="multi-GPU fused transport with --rows-below: process those rows grouped by destination peer and halo slot "
                         "(contiguous peer writes) or by descending degree")
    ap.add_argument("--tiled-slice", type=int, default=64, choices=[16, 32, 64], help="--tiled: floats of the feature dimension per CTA")
    ap.add_argument("--tiled-slack", type=int, default=1)
    ap.add_argument("--tiled-fine-cols", type=int, default=256)
    ap.add_argument("--batch-size", type=int, default=128, help="pubmed_batch: rows per batch (batch-main.py:52 default)")
    ap.add_argument("--phases", default="one", choices=["peer", "two", "one"], help="multi-GPU: how a step is split")
    ap.add_argument("--row-groups", type=int, default=4, help="multi-GPU: kernels per step of the pipelined push")
    ap.add_argument("--stripes", type=int, default=0, help="multi-GPU: block-cyclic stripes per rank (0 = auto, ~4096-id stripes; 1 = plain contiguous blocks)")
    ap.add_argument("--dist-idx16", action="store_true", help="multi-GPU fused transport: 16-byte index staging (validated on one GPU only)")
    ap.add_argument("--hub-degree", type=int, default=64, help="multi-GPU --transport hybrid: rows of at least this degree are summed where their columns live")
    ap.add_argument("--no-degree-sort", action="store_true", help="partitioned path: keep the generator's order of the rows inside a block "
                    "(default: rows of a block are stored by descending degree, so that hot 64-byte rows share their 128-byte lines)")
    ap.add_argument("--row-cost", type=float, default=None, help="multi-GPU: weight of a row, in stored entries, added to its non-zeros when the "
                    "row blocks are cut (default: dist.auto_row_cost; 0 = cut by non-zeros alone)")
    ap.add_argument("--transport", default="auto", choices=["auto", "fused", "hybrid", "pipe", "pull", "push", "p2p"], help="multi-GPU: halo transport")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
