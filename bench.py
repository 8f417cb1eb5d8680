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
    port: the reference is pure Python and has no K-step propagation to run)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    __import__("__graft_entry__").build()
    wl = args.workload or "rmat2m"
    if wl == "rmat100m":
        wl = "rmat2m"  # host RAM/time: the CPU arm is timed on config 4
    oracle, n, F, oip, oidx, oval = host_graph(wl)
    cores = oracle.clib().oracle_num_threads()
    k_sample = 2
    rates = []
    for _ in range(max(1, args.warmup if args.warmup is not None else 1)):
        cpu_port_rate(oracle, n, F, oip, oidx, oval, 1, 1)
    steps = args.steps or 3
    t_total = 0.0
    for _ in range(steps):
        r, t = cpu_port_rate(oracle, n, F, oip, oidx, oval, k_sample, 1)
        rates.append(r)
        t_total += t
    value = sum(rates) / len(rates)
    sample = f"{k_sample} of the 20 propagation steps of one pass per bench step ({wl}, nnz(A_hat)={len(oidx)}, F={F})"
    ts_rate, ts_threads = cpu_torch_sparse_rate(n, F, oip, oidx, oval, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup if args.warmup is not None else 1,
        "ms_per_step": 1e3 * t_total / steps * (2 * KSTEPS / k_sample), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "n": n, "nnz_a_hat": int(len(oidx)), "F": F, "K": KSTEPS, "alpha": ALPHA,
                   "pass": "K=10 forward + K=10 backward (extrapolated from the sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "torch_sparse_csr": {"value": ts_rate, "threads": ts_threads, "sample": "1 step"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
    # config 5 family: always through the partitioned code path, also at N=1, so that T_1 and T_P of the
    # scaling study come from the same code
    partitioned = world > 1 or wl in ("rmat100m", "rmat16m")
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
                                               "min_piece": args.carve_min_piece} if args.order == "carve" else None))
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
                "extra": dict(result, n1_same_workload=n1_reference(wl)),
            }
            print(json.dumps(line))
        dist.barrier()
        dist.destroy_process_group()
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
        graph = P.PropagationGraph(ahat, chunk_edges=args.chunk_edges, order=args.order, idx16=args.idx16, carve=carve)
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
    launches_per_pass = 2 * KSTEPS * (2 if graph.plan.n_fix > 0 else 1)
    traffic, traffic_src = measured_traffic(f"{wl}/{args.order}/{'stored-values' if args.use_vals else 'value-free'}")

    # ---- e2e: host buffers through the public API, copies inside the timed region
    e2e_steps = max(2, min(steps, 3))
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
    e2e_steps = max(e2e_steps, 4)
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
                   "idx16": bool(args.idx16), "carve": graph.plan.carve,
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
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--order", default="degree", choices=["auto", "natural", "degree", "carve"],
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
    ap.add_argument("--chunk-edges", type=int, default=256)
    ap.add_argument("--use-vals", action="store_true", help="stored-value form in every step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--phases", default="one", choices=["peer", "two", "one"], help="multi-GPU: how a step is split")
    ap.add_argument("--row-groups", type=int, default=4, help="multi-GPU: kernels per step of the pipelined push")
    ap.add_argument("--stripes", type=int, default=0, help="multi-GPU: block-cyclic stripes per rank (0 = auto, ~4096-id stripes; 1 = plain contiguous blocks)")
    ap.add_argument("--dist-idx16", action="store_true", help="multi-GPU fused transport: 16-byte index staging (validated on one GPU only)")
    ap.add_argument("--hub-degree", type=int, default=64, help="multi-GPU --transport hybrid: rows of at least this degree are summed where their columns live")
    ap.add_argument("--transport", default="auto", choices=["auto", "fused", "hybrid", "pipe", "pull", "push", "p2p"], help="multi-GPU: halo transport")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
