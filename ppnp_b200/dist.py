"""Row-partitioned multi-GPU APPNP (BASELINE.json config 5; not in the reference, which is
single-process -- SURVEY.md section 2.3 / 8e).

Rank p owns a contiguous block of rows of A_hat, Z and H.  Row i of Z_{k+1} needs the Z_k rows of
i's neighbours only, so one exchange per iteration suffices:

  * the shard's column ids are remapped to [local rows | halo slots]; halo slots are the distinct
    remote rows the shard references, grouped by owner and sorted (``build_shard_topology``);
  * row-block boundaries are chosen by the non-zero prefix sum of a block-cyclically relabelled id
    space (``stripe_relabel``, ``balanced_row_blocks``): R-MAT rows are heavily skewed, equal row
    blocks would put 44 % of the edges on rank 0 of 8, and un-mixed nnz-balanced blocks make the
    last rank ship 11x the rows of the first;
  * the exchange itself has several implementations, all driving the same fused SpMM+teleport
    kernel (csrc/appnp_spmm.cu) over the extended [local | halo] buffer:
      - ``FusedPushPropagation`` (default on GPUs): the kernel's epilogue stores every finished row
        that peers reference straight into their halo slots over NVLink peer memory; one 4-byte
        all-reduce per step is the barrier;
      - ``PipelinedPushPropagation``: row groups, the rows of a finished group are pushed by a
        gather kernel on a side stream while the next group computes;
      - ``PartitionedPropagation``: per-owner / local-remote edge phases with the accumulate
        epilogue (PPNP_EPI_ACC) over peer pull, owner push, NCCL point-to-point or one all-to-all.
    profiles/r01_scaling.md has the measurements that ordered them.

The backward pass is the same operator on the upstream gradient (A_hat symmetric), so it uses
the same partition and the same lists.

Everything here is index bookkeeping plus torch.distributed / symmetric-memory calls; the
arithmetic is in the CUDA library.  Topology, lists and every orchestration are device-agnostic
and run on CPU with the gloo backend (tests/test_dist_cpu.py) with a numpy walker of the edge
stream standing in for the kernel; the package itself has no CPU implementation of the kernel.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------ rank barrier
_FLAGS = {}


def rank_barrier(handle, device, group=None):
    """Stream-ordered barrier across the ranks between two propagation steps: everything enqueued on
    the current stream before it (kernels that stored into peers' memory included) has completed on
    EVERY rank before anything enqueued after it starts: a 4-byte NCCL all-reduce (measured correct for
    the in-kernel peer stores).  The symmetric-memory signal barrier was faster but let a rank run ahead in
    the fused-push tests on 2 GPUs; it is not selectable any more (PPNP_DIST_BARRIER=symm raises)."""
    import os
    if os.environ.get("PPNP_DIST_BARRIER", "nccl") != "nccl":
        raise RuntimeError("PPNP_DIST_BARRIER: only the NCCL barrier is supported (the symmetric-memory signal barrier "
                           "let ranks run ahead of in-kernel peer stores and was removed)")
    key = (device, id(group))
    flag = _FLAGS.get(key)
    if flag is None:
        flag = torch.zeros(1, device=device)
        _FLAGS[key] = flag
    dist.all_reduce(flag, group=group)


# ------------------------------------------------------------------------------ partitioning
def balanced_row_blocks(weights, world):
    """Boundaries lo_0=0 <= ... <= lo_world=n such that every block carries ~1/world of the weight
    (weights = per-row non-zeros incl. the self loop).  Returns a python list of world+1 ints."""
    w = weights.to(torch.float64)
    prefix = torch.cumsum(w, 0)
    total = float(prefix[-1])
    targets = torch.arange(1, world, dtype=torch.float64, device=w.device) * (total / world)
    cuts = torch.searchsorted(prefix, targets, right=False) + 1
    bounds = [0] + [int(c) for c in cuts.tolist()] + [int(w.numel())]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


@dataclass
class ShardTopology:
    """Local view of rank ``rank``: remapped CSR of its rows of A + I and the halo description."""
    rank: int
    world: int
    bounds: List[int]                 # world + 1 row boundaries
    n_local: int
    indptr: torch.Tensor              # int64 [n_local + 1]
    indices: torch.Tensor             # int32 [nnz_local], local rows then halo slots (n_local + h)
    halo_cols: torch.Tensor           # int64 [n_halo] global ids, grouped by owner, ascending
    recv_counts: List[int]            # halo rows owned by each rank
    interior: torch.Tensor            # bool [n_local]: row references no halo slot

    @property
    def n_halo(self):
        return int(self.halo_cols.numel())


def build_shard_topology(indptr_local, cols_global, bounds, rank):
    """indptr_local/cols_global: CSR of this rank's rows of A + I with GLOBAL column ids (sorted)."""
    world = len(bounds) - 1
    lo, hi = bounds[rank], bounds[rank + 1]
    n_local = hi - lo
    dev = cols_global.device
    cg = cols_global.to(torch.int64)
    is_local = (cg >= lo) & (cg < hi)
    remote = cg[~is_local]
    halo_cols = torch.unique(remote, sorted=True)          # ascending global id == grouped by owner
    b = torch.tensor(bounds, dtype=torch.int64, device=dev)
    owner = torch.searchsorted(b, halo_cols, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world).tolist() if halo_cols.numel() else [0] * world
    remap = torch.where(is_local, cg - lo, n_local + torch.searchsorted(halo_cols, cg))
    ip = indptr_local.to(torch.int64)
    # interior rows: every column local
    row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), ip[1:] - ip[:-1])
    has_remote = torch.zeros(n_local, dtype=torch.bool, device=dev)
    has_remote[row_of[~is_local]] = True
    return ShardTopology(rank=rank, world=world, bounds=list(bounds), n_local=n_local, indptr=ip,
                         indices=remap.to(torch.int32), halo_cols=halo_cols, recv_counts=[int(c) for c in recv_counts],
                         interior=~has_remote)


# ------------------------------------------------------------------------------ halo exchange
class HaloExchange:
    """Per-step exchange of boundary rows.  Set-up trades the halo id lists once (all_to_all of
    counts, then of ids); ``start``/``finish`` move the rows: pack -> all_to_all_single -> the halo
    region of the destination buffer."""

    def __init__(self, topo: ShardTopology, group=None):
        self.topo, self.group = topo, group
        dev = topo.halo_cols.device
        world = topo.world
        recv_counts = torch.tensor(topo.recv_counts, dtype=torch.int64, device=dev)
        send_counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_counts, recv_counts, group=group)          # how many rows each peer wants from me
        self.recv_counts = topo.recv_counts
        self.send_counts = [int(c) for c in send_counts.tolist()]
        want = topo.halo_cols                                                   # ids I want, grouped by owner
        give = torch.empty(sum(self.send_counts), dtype=torch.int64, device=dev)
        dist.all_to_all_single(give, want, output_split_sizes=self.send_counts, input_split_sizes=self.recv_counts, group=group)
        lo = topo.bounds[topo.rank]
        self.send_idx = (give - lo).contiguous()                                # local rows to ship, grouped by destination
        if self.send_idx.numel():
            assert int(self.send_idx.min()) >= 0 and int(self.send_idx.max()) < topo.n_local
        self._send_buf = {}

    def bytes_per_step(self, F):
        return (sum(self.send_counts) + sum(self.recv_counts)) * F * 4

    def exchange(self, Zext, async_op=False):
        """Fill rows n_local.. of ``Zext`` ([n_local + n_halo, F]) with the owners' current rows 0..n_local-1."""
        t = self.topo
        F = Zext.shape[1]
        key = (F, Zext.device)
        buf = self._send_buf.get(key)
        if buf is None:
            buf = torch.empty((max(self.send_idx.numel(), 1), F), dtype=Zext.dtype, device=Zext.device)
            self._send_buf[key] = buf
        send = buf[: self.send_idx.numel()]
        if self.send_idx.numel():
            if Zext.is_cuda:
                from .ops import gather_rows
                gather_rows(Zext[: t.n_local], self.send_idx, send)
            else:
                torch.index_select(Zext[: t.n_local], 0, self.send_idx, out=send)
        recv = Zext[t.n_local:]
        return dist.all_to_all_single(recv, send, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts,
                                      group=self.group, async_op=async_op)


# ------------------------------------------------------------------------------ transports
class PeerPull:
    """Halo transport over peer memory: the Z buffers are symmetric-memory allocations, every rank maps
    its peers' buffers over NVLink and PULLS the rows it needs with one gather kernel per owner
    (csrc/rows.cu) -- de-duplicated rows, no pack on the sender, no collective; a device-side barrier
    per step orders the pulls after the owners' writes."""

    def __init__(self, topo: ShardTopology, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.topo, self.group = topo, (group if group is not None else dist.group.WORLD)
        dev = topo.halo_cols.device
        # ids I want from each owner, local to the owner, and where they land in my halo region
        self.want, self.offs = [], []
        o = 0
        for q in range(topo.world):
            c = topo.recv_counts[q]
            self.want.append((topo.halo_cols[o: o + c] - topo.bounds[q]).contiguous())
            self.offs.append(o)
            o += c
        ext = torch.tensor([topo.n_local + topo.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        self.rows_alloc = int(ext)
        self.handles = {}

    def alloc(self, F, count=3):
        """``count`` symmetric [rows_alloc, F] buffers (same size on every rank)."""
        dev = self.topo.halo_cols.device
        bufs = []
        for _ in range(count):
            t = self.symm_mem.empty((self.rows_alloc, F), dtype=torch.float32, device=dev)
            h = self.symm_mem.rendezvous(t, self.group)
            self.handles[t.data_ptr()] = (h, F)
            bufs.append(t)
        return bufs

    def step_barrier(self, buf):
        rank_barrier(self.handles[buf.data_ptr()][0], buf.device, self.group)

    def fetch(self, src, owner):
        """Pull owner's rows of ``src`` (the same symmetric buffer on every rank) into my halo region."""
        from .ops import gather_rows
        t = self.topo
        c = t.recv_counts[owner]
        if c == 0:
            return
        h, F = self.handles[src.data_ptr()]
        n_owner = t.bounds[owner + 1] - t.bounds[owner]
        peer = h.get_buffer(owner, (n_owner, F), torch.float32)
        gather_rows(peer, self.want[owner], src[t.n_local + self.offs[owner]: t.n_local + self.offs[owner] + c])


class PeerPush:
    """Halo transport over peer memory, owner-driven: every rank gathers the rows a peer needs from its
    local Z (fast, local HBM) and writes them CONTIGUOUSLY into that peer's halo slots through the
    symmetric-memory mapping (csrc/rows.cu with a peer destination): NVLink sees full-line posted
    writes instead of 64-byte read round trips.  One device-side barrier per step tells every rank
    that its halo has landed."""

    def __init__(self, topo: ShardTopology, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.topo, self.group = topo, (group if group is not None else dist.group.WORLD)
        self.hx = HaloExchange(topo, group)                     # trades the send lists once
        dev = topo.halo_cols.device
        P = topo.world
        mine = torch.tensor(topo.recv_counts + [topo.n_local], dtype=torch.int64, device=dev)
        allc = torch.empty(P * (P + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)
        allc = allc.view(P, P + 1).cpu()
        # where my rows land inside peer q's halo region: after the rows of all owners < me
        self.dst_off = [int(allc[q, P]) + int(allc[q, : topo.rank].sum()) for q in range(P)]
        self.soffs, o = [], 0
        for q in range(P):
            self.soffs.append(o)
            o += self.hx.send_counts[q]
        ext = torch.tensor([topo.n_local + topo.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        self.rows_alloc = int(ext)
        self.handles = {}

    def alloc(self, F, count=3):
        dev = self.topo.halo_cols.device
        bufs = []
        for _ in range(count):
            t = self.symm_mem.empty((self.rows_alloc, F), dtype=torch.float32, device=dev)
            h = self.symm_mem.rendezvous(t, self.group)
            self.handles[t.data_ptr()] = (h, F)
            bufs.append(t)
        return bufs

    def step_barrier(self, buf):
        pass                                                    # the barrier after the pushes orders everything

    def push_all(self, src):
        """Write my rows into every peer's halo slots of the same symmetric buffer, then barrier."""
        from .ops import gather_rows
        t = self.topo
        h, F = self.handles[src.data_ptr()]
        for d in range(1, t.world):
            q = (t.rank + d) % t.world                          # start with a different peer on every rank
            ns = self.hx.send_counts[q]
            if ns == 0:
                continue
            peer = h.get_buffer(q, (self.rows_alloc, F), torch.float32)
            ids = self.hx.send_idx[self.soffs[q]: self.soffs[q] + ns]
            gather_rows(src[: t.n_local], ids, peer[self.dst_off[q]: self.dst_off[q] + ns])
        rank_barrier(h, src.device, self.group)


class RoundSendRecv:
    """Halo transport over torch.distributed point-to-point (NCCL on GPUs, gloo in the CPU tests): the
    rows a peer needs are packed (csrc/rows.cu on CUDA) and sent, its rows for me are received straight
    into their halo slots."""

    def __init__(self, topo: ShardTopology, group=None):
        self.hx = HaloExchange(topo, group)
        self.topo, self.group = topo, group
        self.soffs, o = [], 0
        for q in range(topo.world):
            self.soffs.append(o)
            o += self.hx.send_counts[q]
        self.roffs, o = [], 0
        for q in range(topo.world):
            self.roffs.append(o)
            o += topo.recv_counts[q]
        self._buf = {}

    def alloc(self, F, count=3):
        dev = self.topo.halo_cols.device
        return [torch.zeros((self.topo.n_local + self.topo.n_halo, F), dtype=torch.float32, device=dev) for _ in range(count)]

    def step_barrier(self, buf):
        pass

    def exchange_round(self, src, dest, source):
        """Send my rows that ``dest`` needs, receive ``source``'s rows that I need."""
        t = self.topo
        F = src.shape[1]
        ops = []
        ns, nr = self.hx.send_counts[dest], t.recv_counts[source]
        if ns:
            key = (F, dest)
            buf = self._buf.get(key)
            if buf is None:
                buf = torch.empty((ns, F), dtype=torch.float32, device=src.device)
                self._buf[key] = buf
            ids = self.hx.send_idx[self.soffs[dest]: self.soffs[dest] + ns]
            if src.is_cuda:
                from .ops import gather_rows
                gather_rows(src[: t.n_local], ids, buf)
            else:
                torch.index_select(src[: t.n_local], 0, ids, out=buf)
            ops.append(dist.P2POp(dist.isend, buf, dest, self.group))
        if nr:
            ops.append(dist.P2POp(dist.irecv, src[t.n_local + self.roffs[source]: t.n_local + self.roffs[source] + nr], source, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


# ------------------------------------------------------------------------------ propagation
class PartitionedPropagation:
    """K-step APPNP over one shard.

    A step is split into PHASES by where an edge's column lives: phase 0 = columns owned by this
    rank (every row has its self loop there, so phase 0 writes every output row, teleport term
    included); phase t >= 1 = columns owned by rank (rank - t) mod P, ADDED to the rows phase 0 wrote
    (PPNP_EPI_ACC).  Transfers run on a side stream in the same order, so the rows of owner t+1 are in
    flight while the edges that point at owner t are processed -- the exchange hides behind the
    compute edge by edge, not just behind the few rows without remote neighbours.
    ``phases``: "peer" (one phase per owner), "two" (local, then all remote), "one" (no split)."""

    def __init__(self, topo: ShardTopology, deg_global_dinv, chunk_edges=256, phases="peer", transport="auto",
                 group=None, step_fn=None):
        """``step_fn(plan, Zin, T, Zout, alpha, epi, use_vals)`` replaces the CUDA launch; it exists so
        that tests can drive this orchestration on CPU tensors (gloo) with a numpy walker of the
        plan.  The package itself ships no CPU implementation: the default is the CUDA library."""
        from .plan import build_stream_plan
        self.topo, self._step_fn, self.phases = topo, step_fn, phases
        dev = topo.indices.device
        self.on_gpu = dev.type == "cuda"
        if transport == "auto":
            transport = "push" if (self.on_gpu and topo.world > 1) else "p2p"
        if transport == "push" and phases == "peer":
            phases = "two"                                   # pushes complete together: local phase, then remote phase
        self.phases = phases
        auto = transport == "push" and self.on_gpu
        try:
            self.transport = {"pull": PeerPull, "push": PeerPush}.get(transport, RoundSendRecv)(topo, group)
            if auto:
                self.transport.alloc(4, 1)                   # probe: symmetric memory must be usable on this box
        except Exception as e:  # noqa: BLE001  (peer mapping unavailable: NCCL point-to-point still works)
            if not auto:
                raise
            import warnings
            warnings.warn(f"ppnp_b200.dist: peer-memory transport unavailable ({type(e).__name__}: {e}); using NCCL p2p")
            transport = "p2p"
            self.transport = RoundSendRecv(topo, group)
        self.transport_name = transport
        P, rank, n_local = topo.world, topo.rank, topo.n_local
        ip = topo.indptr
        deg = ip[1:] - ip[:-1]
        lo = topo.bounds[rank]
        # stored values of the first step: dinv_i * dinv_j with GLOBAL degrees
        dinv_ext = torch.cat([deg_global_dinv[lo: lo + n_local], deg_global_dinv[topo.halo_cols]])
        row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), deg)
        cols = topo.indices.to(torch.int64)
        vals = dinv_ext[row_of] * dinv_ext[cols]
        # phase of every edge
        b = torch.tensor(topo.bounds, dtype=torch.int64, device=dev)
        halo_owner = torch.searchsorted(b, topo.halo_cols, right=True) - 1
        if phases == "one" or P == 1:
            self.rounds = []                              # (phase id, owner) pairs after phase 0
            edge_phase = torch.zeros_like(cols)
            n_ph = 1
        else:
            if phases == "two":
                owner_phase = torch.ones(P, dtype=torch.int64, device=dev)
                self.rounds = [(1, None)]
                n_ph = 2
            else:
                owner_phase = (rank - torch.arange(P, device=dev)) % P       # owner q arrives in round (rank - q) mod P
                self.rounds = [(t, (rank - t) % P) for t in range(1, P)]
                n_ph = P
            ext_phase = torch.cat([torch.zeros(n_local, dtype=torch.int64, device=dev), owner_phase[halo_owner]])
            edge_phase = ext_phase[cols]
        self.plans = []
        deg_f = deg.to(torch.float32)
        for ph in range(n_ph):
            m = edge_phase == ph
            cnt = torch.bincount(row_of[m], minlength=n_local)
            rows = torch.nonzero(cnt).flatten()
            if rows.numel() == 0:
                self.plans.append(None)
                continue
            ipp = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
            ipp[1:] = torch.cumsum(cnt, 0)
            order = rows[torch.sort(cnt[rows], descending=True, stable=True).indices]
            plan = build_stream_plan(ipp, topo.indices[m], vals[m], chunk_edges, order, subset=True,
                                     row_deg=(deg_f if n_ph > 1 else None))
            self.plans.append(_SubGraph(plan, step_fn))
            del m, cnt, ipp
        del row_of, cols, vals, edge_phase
        self.comm_stream = torch.cuda.Stream(device=dev) if self.on_gpu else None
        self.exchange = getattr(self.transport, "hx", None)

    def n_ext(self):
        return self.topo.n_local + self.topo.n_halo

    def alloc(self, F):
        """H, Z, S buffers of [>= n_local + n_halo, F] usable with this transport."""
        return self.transport.alloc(F, 3)

    def _transfer(self, src, rnd):
        ph, owner = rnd
        t = self.topo
        if owner is None:                                  # all owners at once
            if isinstance(self.transport, PeerPush):
                self.transport.push_all(src)
            elif isinstance(self.transport, PeerPull):
                for q in range(t.world):
                    if q != t.rank:
                        self.transport.fetch(src, q)
            else:
                self.transport.hx.exchange(src[: t.n_local + t.n_halo])
        elif isinstance(self.transport, PeerPull):
            self.transport.fetch(src, owner)
        else:
            self.transport.exchange_round(src, dest=(t.rank + ph) % t.world, source=owner)

    def propagate(self, H_ext, Z_ext, S_ext, K, alpha):
        """H_ext[:n_local] holds the input; result in Z_ext[:n_local].  Value-free Y-space iteration as on
        one GPU (epilogues PPNP_EPI_*)."""
        from . import _lib
        t = self.topo
        cur = torch.cuda.current_stream() if self.on_gpu else None
        if t.world > 1 and self.on_gpu:
            rank_barrier(None, H_ext.device, getattr(self, "group", None))     # entry barrier: peers' buffers are ready
        src = H_ext
        for k in range(1, K + 1):
            dst = Z_ext if (K - k) % 2 == 0 else S_ext
            if K == 1:
                epi, use_vals = _lib.EPI_PLAIN, True
            elif k == 1:
                epi, use_vals = _lib.EPI_Z2Y, True
            elif k == K:
                epi, use_vals = _lib.EPI_Y2Z, False
            else:
                epi, use_vals = _lib.EPI_Y, False
            self.transport.step_barrier(src)               # every rank has finished writing `src` (and reading `dst`)
            if not self.rounds:                            # single phase: full exchange, then everything
                if t.world > 1:
                    self._transfer(src, (1, None))
                self.plans[0].step(src, H_ext, dst, alpha, epi, use_vals)
            elif self.on_gpu:
                self.comm_stream.wait_stream(cur)
                events = []
                with torch.cuda.stream(self.comm_stream):
                    for rnd in self.rounds:
                        self._transfer(src, rnd)
                        ev = torch.cuda.Event()
                        ev.record(self.comm_stream)
                        events.append(ev)
                self.plans[0].step(src, H_ext, dst, alpha, epi, use_vals)
                for rnd, ev in zip(self.rounds, events):
                    cur.wait_event(ev)
                    if self.plans[rnd[0]] is not None:
                        self.plans[rnd[0]].step(src, dst, dst, alpha, epi | _lib.EPI_ACC, use_vals)
            else:
                self.plans[0].step(src, H_ext, dst, alpha, epi, use_vals)
                for rnd in self.rounds:
                    self._transfer(src, rnd)
                    if self.plans[rnd[0]] is not None:
                        self.plans[rnd[0]].step(src, dst, dst, alpha, epi | _lib.EPI_ACC, use_vals)
            src = dst
        return Z_ext[: t.n_local]


class PipelinedPushPropagation:
    """K-step APPNP over one shard with a SENDER-side pipelined halo push.

    The shard's rows are cut into ``row_groups`` contiguous groups of equal non-zeros; a step runs one
    kernel per group.  As soon as group g of step k is final, its rows that peers need are gathered
    and written into the peers' halo slots of the step-(k+1) source buffer (symmetric memory over
    NVLink, csrc/rows.cu) on a side stream -- while the kernels of groups g+1.. are still running.
    Only the push of the LAST group (1/row_groups of the halo) plus one device-side barrier is exposed
    per step.  Rows are never split between kernels, so there is no accumulate pass and no extra
    traffic; the arithmetic (and its order) is that of the single-GPU kernel.
    On CPU tensors (gloo tests) the push is emulated with point-to-point messages of the same slices."""

    def __init__(self, topo: ShardTopology, deg_global_dinv, chunk_edges=256, row_groups=4, group=None, step_fn=None):
        from .plan import build_stream_plan
        self.topo, self.group = topo, group
        dev = topo.indices.device
        self.on_gpu = dev.type == "cuda"
        P, rank, n_local = topo.world, topo.rank, topo.n_local
        ip = topo.indptr
        deg = ip[1:] - ip[:-1]
        lo = topo.bounds[rank]
        dinv_ext = torch.cat([deg_global_dinv[lo: lo + n_local], deg_global_dinv[topo.halo_cols]])
        row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), deg)
        vals = dinv_ext[row_of] * dinv_ext[topo.indices.to(torch.int64)]
        del row_of
        G = max(1, min(row_groups, n_local))
        self.gb = balanced_row_blocks(deg, G)                     # local row boundaries of the groups
        self.plans = []
        for g in range(G):
            a, b = self.gb[g], self.gb[g + 1]
            if b <= a:
                self.plans.append(None)
                continue
            rows = torch.arange(a, b, device=dev)
            order = rows[torch.sort(deg[a:b], descending=True, stable=True).indices]
            self.plans.append(_SubGraph(build_stream_plan(ip, topo.indices, vals, chunk_edges, order, subset=True), step_fn))
        del vals
        self.G = G
        # send lists (sorted local ids per destination) cut at the group boundaries
        self.hx = HaloExchange(topo, group)
        gbt = torch.tensor(self.gb, dtype=torch.int64, device=dev)
        self.soffs, o = [], 0
        scnt = torch.zeros((P, G), dtype=torch.int64)
        for q in range(P):
            self.soffs.append(o)
            ns = self.hx.send_counts[q]
            if ns:
                ids = self.hx.send_idx[o: o + ns]
                cuts = torch.searchsorted(ids, gbt)                # position of every group boundary in the list
                scnt[q] = (cuts[1:] - cuts[:-1]).cpu()
            o += ns
        self.scnt = scnt                                          # rows I send to q out of my group g
        # what every peer sends me per group, and where my rows land at every peer
        allc = torch.empty(P * P * G, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, scnt.to(dev).flatten().contiguous(), group=group)
        allc = allc.view(P, P, G).cpu()                           # allc[s, q, g]: s sends q in s's group g
        self.rcnt = allc[:, rank, :].clone()                      # [source, g]
        mine = torch.tensor(topo.recv_counts + [topo.n_local], dtype=torch.int64, device=dev)
        rc_all = torch.empty(P * (P + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(rc_all, mine, group=group)
        rc_all = rc_all.view(P, P + 1).cpu()
        # my rows for q land after q's local rows and after the rows of all owners < me
        self.dst_off = [int(rc_all[q, P]) + int(rc_all[q, :rank].sum()) for q in range(P)]
        self.roffs, o = [], 0
        for q in range(P):
            self.roffs.append(o)
            o += topo.recv_counts[q]
        for q in range(P):                                        # consistency of the two views of the halo
            assert int(self.rcnt[q].sum()) == topo.recv_counts[q]
        ext = torch.tensor([topo.n_local + topo.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        self.rows_alloc = int(ext)
        self.handles = {}
        self._sendbuf = {}
        self.comm_stream = torch.cuda.Stream(device=dev) if self.on_gpu else None
        self.transport_name = "pipelined-push" if self.on_gpu else "pipelined-p2p"
        self.phases = f"{G} row groups"
        self.rounds = []

    # -- buffers
    def alloc(self, F, count=3):
        dev = self.topo.indices.device
        if not self.on_gpu:
            return [torch.zeros((self.rows_alloc, F), dtype=torch.float32, device=dev) for _ in range(count)]
        import torch.distributed._symmetric_memory as symm_mem
        bufs = []
        for _ in range(count):
            t = symm_mem.empty((self.rows_alloc, F), dtype=torch.float32, device=dev)
            h = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self.handles[t.data_ptr()] = (h, F)
            bufs.append(t)
        return bufs

    def n_ext(self):
        return self.topo.n_local + self.topo.n_halo

    # -- transfers
    def _push_group(self, buf, g):
        """Ship the rows of my group g that peers need into their halo slots of ``buf``."""
        t = self.topo
        P, rank = t.world, t.rank
        if self.on_gpu:
            from .ops import gather_rows
            h, F = self.handles[buf.data_ptr()]
            for d in range(1, P):
                q = (rank + d) % P
                ns = int(self.scnt[q, g])
                if ns == 0:
                    continue
                s0 = self.soffs[q] + int(self.scnt[q, :g].sum())
                peer = h.get_buffer(q, (self.rows_alloc, F), torch.float32)
                o = self.dst_off[q] + int(self.scnt[q, :g].sum())
                gather_rows(buf[: t.n_local], self.hx.send_idx[s0: s0 + ns], peer[o: o + ns])
        else:
            ops, keep = [], []
            for d in range(1, P):
                q, s = (rank + d) % P, (rank - d) % P
                ns = int(self.scnt[q, g])
                if ns:
                    s0 = self.soffs[q] + int(self.scnt[q, :g].sum())
                    sb = torch.index_select(buf[: t.n_local], 0, self.hx.send_idx[s0: s0 + ns])
                    keep.append(sb)
                    ops.append(dist.P2POp(dist.isend, sb, q, self.group))
                nr = int(self.rcnt[s, g])
                if nr:
                    o = t.n_local + self.roffs[s] + int(self.rcnt[s, :g].sum())
                    ops.append(dist.P2POp(dist.irecv, buf[o: o + nr], s, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()

    def _barrier(self, buf):
        if self.on_gpu:
            rank_barrier(self.handles[buf.data_ptr()][0], buf.device, self.group)

    def transfers_only(self, buf):
        for g in range(self.G):
            self._push_group(buf, g)
        self._barrier(buf)

    # -- the iteration
    def propagate(self, H_ext, Z_ext, S_ext, K, alpha):
        from . import _lib
        t = self.topo
        multi = t.world > 1
        cur = torch.cuda.current_stream() if self.on_gpu else None
        ready = None
        if multi and self.on_gpu:
            rank_barrier(None, H_ext.device, getattr(self, "group", None))     # entry barrier: peers' buffers are ready
        if multi:                                              # halo of the input for step 1
            if self.on_gpu:
                self.comm_stream.wait_stream(cur)
                with torch.cuda.stream(self.comm_stream):
                    self.transfers_only(H_ext)
                    ready = torch.cuda.Event()
                    ready.record(self.comm_stream)
            else:
                self.transfers_only(H_ext)
        src = H_ext
        for k in range(1, K + 1):
            dst = Z_ext if (K - k) % 2 == 0 else S_ext
            if K == 1:
                epi, use_vals = _lib.EPI_PLAIN, True
            elif k == 1:
                epi, use_vals = _lib.EPI_Z2Y, True
            elif k == K:
                epi, use_vals = _lib.EPI_Y2Z, False
            else:
                epi, use_vals = _lib.EPI_Y, False
            if ready is not None:
                cur.wait_event(ready)
            push = multi and k < K
            for g in range(self.G):
                if self.plans[g] is not None:
                    self.plans[g].step(src, H_ext, dst, alpha, epi, use_vals)
                if push:
                    if self.on_gpu:
                        ev = torch.cuda.Event()
                        ev.record(cur)
                        with torch.cuda.stream(self.comm_stream):
                            self.comm_stream.wait_event(ev)
                            self._push_group(dst, g)
                    else:
                        self._push_group(dst, g)
            if push and self.on_gpu:
                with torch.cuda.stream(self.comm_stream):
                    self._barrier(dst)
                    ready = torch.cuda.Event()
                    ready.record(self.comm_stream)
            else:
                ready = None
            src = dst
        return Z_ext[: t.n_local]


class FusedPushPropagation:
    """K-step APPNP over one shard with the halo push FUSED into the propagation kernel.

    One kernel per step, the single-GPU one: whenever its epilogue finishes a row that other ranks
    reference, it also stores the row into those ranks' halo slots of the step's output buffer
    through the NVLink peer mappings (symmetric memory; csrc/appnp_spmm.cu push_row).  The transfer
    therefore overlaps the gathers row by row, there is no pack, no collective and no second kernel;
    a device-side barrier between two steps is all that is left of the exchange.  Only the halo of
    the INPUT of a propagation (rows of H, produced by the caller) is shipped by a separate gather
    kernel (csrc/rows.cu) once per call.
    On CPU tensors (gloo tests) the same push lists are applied with point-to-point messages."""

    MAX_PEERS = 8

    def __init__(self, topo: ShardTopology, deg_global_dinv, chunk_edges=256, group=None, step_fn=None, carve=None,
                 idx16=False, rows_below=None, rows_order="dest", window=None):
        """window: "first" | "mid" | "last": the rows of the edge stream keep their columns sorted by rank (hottest first)
        and the whole-segment chunks of the hub rows are processed in column-window order (plan.window_order_chunks):
        plan-only, no new partial sums; raises the L2 hit rate of the gathers.
        carve: keyword arguments of plan.build_carved_plan (block_cols, n_blocks, min_piece): stream the
        shard with its hot column blocks first (blocks sized for the L2: the cold gathers of the hub rows
        stay inside an L2-resident window) instead of in plain degree order.
        idx16: 16-byte staging of a lane-transposed copy of the stream (feature widths 16 and 64), as on one GPU.
        rows_below: rows of the shard with fewer stored entries go through the one-lane-group-per-row kernel
        (csrc/appnp_rows.cu, same fused push in its epilogue), the stream keeps the rows of higher degree.
        rows_order = "dest": those rows are processed grouped by the peer that wants them and, inside a peer, in the
        order of their halo slots there, so that the rows a warp finishes together land in CONSECUTIVE slots of one
        peer: one 64 * GPW-byte contiguous NVLink write per warp store instead of GPW scattered 64-byte ones
        (measured: 64-byte peer writes move at ~490 GB/s, profiles/r01_scaling.md).  "degree": descending degree
        (equal trip counts inside a warp), as on one GPU."""
        import ctypes as C
        from .plan import build_carved_plan, build_stream_plan, rank_sorted_csr, window_order_chunks
        if carve and window:
            raise ValueError("carve and window are alternative stream orders")
        self.topo, self.group, self._step_fn = topo, group, step_fn
        dev = topo.indices.device
        self.on_gpu = dev.type == "cuda"
        P, rank, n_local = topo.world, topo.rank, topo.n_local
        if P > self.MAX_PEERS:
            raise ValueError(f"the fused push addresses at most {self.MAX_PEERS} ranks")
        ip = topo.indptr
        deg = ip[1:] - ip[:-1]
        lo = topo.bounds[rank]
        dinv_ext = torch.cat([deg_global_dinv[lo: lo + n_local], deg_global_dinv[topo.halo_cols]])
        row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), deg)
        vals = dinv_ext[row_of] * dinv_ext[topo.indices.to(torch.int64)]
        del row_of
        self.rows_part = None
        self._rows_below, self._rows_order = rows_below, rows_order
        if carve:
            plan = build_carved_plan(ip, topo.indices, vals, chunk_edges, n_cols=n_local + topo.n_halo, wide_cta=False,
                                     **carve)
        else:
            order = torch.sort(deg, descending=True, stable=True).indices
            if rows_below is not None and step_fn is None and self.on_gpu:
                n_hub = int((deg >= int(rows_below)).sum().item())
                self._low_rows = order[n_hub:]                     # ordered for good below, once the push lists exist
                self._csr = (ip.to(torch.int32).contiguous(), topo.indices.to(torch.int32).contiguous(), vals.contiguous())
                order = order[:n_hub]
                subset = True
            else:
                n_hub, subset = int(order.numel()), False
            if window and n_hub:
                sidx, svals, crank = rank_sorted_csr(ip, topo.indices, vals, n_cols=n_local + topo.n_halo)
                plan = window_order_chunks(build_stream_plan(ip, sidx, svals, chunk_edges, order, subset=subset), crank, key=window)
                del sidx, svals, crank
            else:
                plan = build_stream_plan(ip, topo.indices, vals, chunk_edges, order, subset=subset) if n_hub else None
        self.sub = _SubGraph(plan, step_fn)
        self.plans = [self.sub]
        self.idx16, self._plans16 = bool(idx16) and step_fn is None, {}
        del vals
        # push lists: for every local row, the (peer, slot) pairs that want it
        self.hx = HaloExchange(topo, group)
        mine = torch.tensor(topo.recv_counts + [topo.n_local], dtype=torch.int64, device=dev)
        rc_all = torch.empty(P * (P + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(rc_all, mine, group=group)
        rc_all = rc_all.view(P, P + 1).cpu()
        self.dst_off = [int(rc_all[q, P]) + int(rc_all[q, :rank].sum()) for q in range(P)]
        ext = torch.tensor([topo.n_local + topo.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        self.rows_alloc = int(ext)
        if self.rows_alloc >= (1 << 28):
            raise ValueError("shard too large for the 28-bit slot field of the push code")
        rows, codes, self.soffs, o = [], [], [], 0
        for q in range(P):
            self.soffs.append(o)
            ns = self.hx.send_counts[q]
            if ns:
                rows.append(self.hx.send_idx[o: o + ns])
                codes.append((q << 28) + self.dst_off[q] + torch.arange(ns, device=dev, dtype=torch.int64))
            o += ns
        if rows:
            rows, codes = torch.cat(rows), torch.cat(codes)
            perm = torch.sort(rows, stable=True).indices
            rows, codes = rows[perm], codes[perm]
            cnt = torch.bincount(rows, minlength=n_local)
        else:
            rows = torch.zeros(0, dtype=torch.int64, device=dev)
            codes, cnt = rows.clone(), torch.zeros(n_local, dtype=torch.int64, device=dev)
        pp = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
        pp[1:] = torch.cumsum(cnt, 0)
        self.push_ptr = pp.to(torch.int32).contiguous()
        self.push_code = torch.cat([codes, torch.zeros(1, dtype=torch.int64, device=dev)]).to(torch.int32).contiguous()
        self.push_rows = rows                                     # (CPU emulation)
        # per-row summary: -1 none, >= 0 the only destination, <= -2 several (walk the list)
        first = torch.full((n_local,), -1, dtype=torch.int64, device=dev)
        if rows.numel():
            one = cnt == 1
            start = pp[:-1]
            first[one] = codes[start[one]]
            first[cnt > 1] = -2
        self.push_first = first.to(torch.int32).contiguous()
        if getattr(self, "_low_rows", None) is not None and self._low_rows.numel():
            low = self._low_rows
            if rows_order == "dest":
                # key = (peer of the first destination, slot there); rows nobody wants go last, by descending degree
                if codes.numel():
                    c0 = codes[pp[:-1][low].clamp(max=int(codes.numel()) - 1)]
                    key = torch.where(cnt[low] > 0, c0, torch.full_like(c0, 1 << 40))
                    low = low[torch.sort(key, stable=True).indices]
            rs = _lib_mod().RowsPlanStruct()
            rs.n, rs.n_rows = n_local, int(low.numel())
            self._low_rows_i32 = low.to(torch.int32).contiguous()
            rs.indptr, rs.indices, rs.vals = self._csr[0].data_ptr(), self._csr[1].data_ptr(), self._csr[2].data_ptr()
            rs.rows = self._low_rows_i32.data_ptr()
            self.rows_part = rs
        self.handles, self._bases = {}, {}
        self._C = C
        self.transport_name = (("fused-push" if self.on_gpu else "fused-p2p") + ("" if self.rows_part is None else f"+rows<{rows_below}/{rows_order}")
                               + (f"+window/{window}" if window else ""))
        self.phases, self.rounds = "one kernel per step, push in the epilogue", []

    def alloc(self, F, count=3):
        dev = self.topo.indices.device
        if not self.on_gpu:
            return [torch.zeros((self.rows_alloc, F), dtype=torch.float32, device=dev) for _ in range(count)]
        import torch.distributed._symmetric_memory as symm_mem
        bufs = []
        grp = self.group if self.group is not None else dist.group.WORLD
        for _ in range(count):
            t = symm_mem.empty((self.rows_alloc, F), dtype=torch.float32, device=dev)
            h = symm_mem.rendezvous(t, grp)
            self.handles[t.data_ptr()] = (h, F)
            arr = (self._C.c_void_p * self.MAX_PEERS)()
            for q in range(self.topo.world):
                arr[q] = t.data_ptr() if q == self.topo.rank else h.get_buffer(q, (self.rows_alloc, F), torch.float32).data_ptr()
            self._bases[t.data_ptr()] = arr
            bufs.append(t)
        return bufs

    def n_ext(self):
        return self.topo.n_local + self.topo.n_halo

    def _barrier(self, buf):
        if not self.on_gpu:
            return
        rank_barrier(self.handles[buf.data_ptr()][0], buf.device, self.group)

    def _push_input(self, buf):
        """Halo of the caller's input: one gather kernel per peer into its slots, then the barrier."""
        t = self.topo
        P, rank = t.world, t.rank
        if self.on_gpu:
            from .ops import gather_rows
            h, F = self.handles[buf.data_ptr()]
            for d in range(1, P):
                q = (rank + d) % P
                ns = self.hx.send_counts[q]
                if ns:
                    peer = h.get_buffer(q, (self.rows_alloc, F), torch.float32)
                    gather_rows(buf[: t.n_local], self.hx.send_idx[self.soffs[q]: self.soffs[q] + ns],
                                peer[self.dst_off[q]: self.dst_off[q] + ns])
            self._barrier(buf)
        else:
            self._apply_push_lists_cpu(buf)

    def _apply_push_lists_cpu(self, buf):
        """gloo stand-in for the in-kernel stores: ship (slot, row) pairs peer by peer."""
        t = self.topo
        P, rank = t.world, t.rank
        codes = self.push_code[:-1].to(torch.int64)
        peer_of, slot_of = codes >> 28, codes & 0x0FFFFFFF
        ops, keep, recvs = [], [], []
        for d in range(1, P):
            q, s = (rank + d) % P, (rank - d) % P
            m = peer_of == q
            if int(m.sum()):
                sl, data = slot_of[m].contiguous(), buf[self.push_rows[m]].contiguous()
                keep += [sl, data]
                ops += [dist.P2POp(dist.isend, sl, q, self.group), dist.P2POp(dist.isend, data, q, self.group)]
            nr = t.recv_counts[s]
            if nr:
                rs = torch.empty(nr, dtype=torch.int64)
                rd = torch.empty((nr, buf.shape[1]), dtype=buf.dtype)
                recvs.append((rs, rd))
                ops += [dist.P2POp(dist.irecv, rs, s, self.group), dist.P2POp(dist.irecv, rd, s, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for rs, rd in recvs:
            buf[rs] = rd

    def transfers_only(self, buf):
        if self.topo.world > 1:
            self._push_input(buf)

    def _step(self, src, T, dst, alpha, epi, use_vals, push):
        if self._step_fn is not None:
            self._step_fn(self.sub.plan, src, T, dst, alpha, epi, use_vals)
            if push:
                self._apply_push_lists_cpu(dst)
            return
        from . import _lib
        lib = _lib.load()
        F = src.shape[1]
        plan = self.sub.plan
        if self.rows_part is not None:
            rs = self.rows_part
            rc = lib.ppnp_spmm_step_rows(rs.indptr, rs.indices, rs.vals, rs.rows, rs.n_rows, rs.n, _lib.ptr(src), _lib.ptr(T), _lib.ptr(dst),
                                         F, F, float(alpha), int(epi), int(bool(use_vals)),
                                         _lib.ptr(self.push_ptr) if push else None, _lib.ptr(self.push_code) if push else None,
                                         _lib.ptr(self.push_first) if push else None, self._bases[dst.data_ptr()] if push else None,
                                         self.topo.world if push else 0, _lib.current_stream())
            _lib.check(rc, "ppnp_spmm_step_rows")
            if plan is None:
                return
        if self.idx16 and F in (16, 64):
            from .plan import lane_group_for, lane_transpose
            G = lane_group_for(F)
            if G not in self._plans16:
                self._plans16[G] = lane_transpose(plan, G)
            plan = self._plans16[G]
        partial = None
        if plan.n_slots:
            partial = self.sub._partial.get(F)
            if partial is None:
                partial = torch.empty(plan.n_slots * F, dtype=torch.float32, device=src.device)
                self.sub._partial[F] = partial
        if push:
            rc = lib.ppnp_spmm_step_push(plan.struct(), _lib.ptr(src), _lib.ptr(T), _lib.ptr(dst), _lib.ptr(partial), F, F,
                                         float(alpha), int(epi), int(bool(use_vals)), _lib.ptr(self.push_ptr),
                                         _lib.ptr(self.push_code), _lib.ptr(self.push_first), self._bases[dst.data_ptr()],
                                         self.topo.world,
                                         _lib.current_stream())
        else:
            rc = lib.ppnp_spmm_step(plan.struct(), _lib.ptr(src), _lib.ptr(T), _lib.ptr(dst), _lib.ptr(partial), F, F,
                                    float(alpha), int(epi), int(bool(use_vals)), _lib.current_stream())
        _lib.check(rc, "ppnp_spmm_step_push")

    def propagate(self, H_ext, Z_ext, S_ext, K, alpha):
        from . import _lib
        t = self.topo
        multi = t.world > 1
        if multi:
            # entry barrier: a peer that is ahead must not push this call's input halo into my buffers while I am
            # still initialising them (the caller zeroes / fills H_ext right before the call) or, for K = 1, while
            # the previous call's only step still gathers the halo of ITS input.  One 4-byte all-reduce per call.
            self._barrier(H_ext)
            self._push_input(H_ext)
        src = H_ext
        for k in range(1, K + 1):
            dst = Z_ext if (K - k) % 2 == 0 else S_ext
            if K == 1:
                epi, use_vals = _lib.EPI_PLAIN, True
            elif k == 1:
                epi, use_vals = _lib.EPI_Z2Y, True
            elif k == K:
                epi, use_vals = _lib.EPI_Y2Z, False
            else:
                epi, use_vals = _lib.EPI_Y, False
            push = multi and k < K
            self._step(src, H_ext, dst, alpha, epi, use_vals, push)
            if push:
                self._barrier(dst)
            src = dst
        return Z_ext[: t.n_local]


def _lib_mod():
    from . import _lib
    return _lib


def _push_arrays(n_rows, rows, codes, dev):
    """Per-row push lists in the layout the kernel epilogue reads (csrc/appnp_spmm.cu PushArgs):
    ptr[n_rows + 1] ranges into code[], first[r] = -1 none / >= 0 the only destination / -2 several."""
    if rows.numel():
        perm = torch.sort(rows, stable=True).indices
        rows, codes = rows[perm], codes[perm]
        cnt = torch.bincount(rows, minlength=n_rows)
    else:
        cnt = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    pp = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    pp[1:] = torch.cumsum(cnt, 0)
    first = torch.full((n_rows,), -1, dtype=torch.int64, device=dev)
    if rows.numel():
        one = cnt == 1
        first[one] = codes[pp[:-1][one]]
        first[cnt > 1] = -2
    code = torch.cat([codes.to(torch.int64), torch.zeros(1, dtype=torch.int64, device=dev)])
    return {"ptr": pp.to(torch.int32).contiguous(), "code": code.to(torch.int32).contiguous(),
            "first": first.to(torch.int32).contiguous(), "rows": rows, "codes": codes.to(torch.int64)}


class HybridPushPropagation:
    """Fused-push propagation in which HUB rows are computed where their columns live.

    A row of degree >= ``hub_degree`` keeps only its LOCAL columns at its owner.  Every other rank sums
    the columns it owns into one partial row per such hub (a *virtual row* of its own stream -- by the
    symmetry of the pattern it finds those edges in its own rows) and the kernel epilogue stores that
    partial into a slot of the owner's buffer over NVLink, exactly like a halo row.  After a rank barrier a
    second, small launch of the same kernel (accumulate epilogue, PPNP_EPI_ACC | PPNP_EPI_INPLACE) adds the
    received partials to the hub rows, finishes them and pushes them to the ranks that reference them; a
    second barrier ends the step.  The owner's halo then holds only what its non-hub rows reference:
    tools/halo_model.py puts the rows a rank receives per step at ~62 % of the 1-D form at 8 ranks.

    Buffer rows: [local | virtual rows (outputs only) | halo | partial slots].  Everything is built from the
    kernel features the 1-D form already uses (partial-row streams with a per-row degree, the accumulate
    epilogue, push lists); the virtual rows get the "degree" that makes their epilogue the identity
    ((1-alpha) for the 1/d epilogues, (1-alpha)^2 for the 1/sqrt(d) ones) and a zero teleport row.
    Host logic is covered under gloo (tests/test_dist_cpu.py); not yet measured on GPUs."""

    MAX_PEERS = 8

    def __init__(self, topo: ShardTopology, deg_global_dinv, hub_degree=64, chunk_edges=256, group=None, step_fn=None,
                 alpha=0.1):
        import ctypes as C
        from .plan import build_stream_plan
        self.topo, self.group, self._step_fn, self.alpha = topo, group, step_fn, float(alpha)
        dev = topo.indices.device
        self.on_gpu = dev.type == "cuda"
        P, rank, n_local = topo.world, topo.rank, topo.n_local
        if P > self.MAX_PEERS:
            raise ValueError(f"the fused push addresses at most {self.MAX_PEERS} ranks")
        lo = topo.bounds[rank]
        bnd = torch.tensor(topo.bounds, dtype=torch.int64, device=dev)
        ip = topo.indptr
        cnt = ip[1:] - ip[:-1]
        dinv_ext = torch.cat([deg_global_dinv[lo: lo + n_local], deg_global_dinv[topo.halo_cols]]).to(torch.float64)
        deg_ext = torch.round(1.0 / (dinv_ext * dinv_ext)).to(torch.int64)
        hub_ext = deg_ext >= int(hub_degree)
        row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), cnt)
        col = topo.indices.to(torch.int64)
        remote = col >= n_local
        drop = hub_ext[row_of] & remote                      # a hub's remote columns are summed where they live
        virt = remote & hub_ext[col]                         # (remote hub h, local row c): c is a column of h's virtual row
        # ---- new column space
        vh_old = torch.unique(col[virt], sorted=True)        # old halo slots of the remote hubs I contribute to
        nv = int(vh_old.numel())
        keep = ~drop
        halo_old = torch.unique(col[keep & remote], sorted=True)
        nh = int(halo_old.numel())
        self.n_virtual, self.n_halo = nv, nh
        halo_global = topo.halo_cols[halo_old - n_local]
        owner_h = torch.searchsorted(bnd, halo_global, right=True) - 1
        recv_counts = torch.bincount(owner_h, minlength=P).tolist() if nh else [0] * P
        # ---- main CSR: local rows (hub rows truncated) then virtual rows
        kcol = col[keep]
        kcol = torch.where(kcol < n_local, kcol, n_local + nv + torch.searchsorted(halo_old, kcol))
        kcnt = torch.bincount(row_of[keep], minlength=n_local)
        kval = (dinv_ext[row_of[keep]] * dinv_ext[col[keep]]).to(torch.float32)
        v_of = torch.searchsorted(vh_old, col[virt])         # virtual row of every contributing edge
        vperm = torch.sort(v_of * n_local + row_of[virt], stable=True).indices
        vcol = row_of[virt][vperm]
        vval = (dinv_ext[col[virt]] * dinv_ext[row_of[virt]]).to(torch.float32)[vperm]
        vcnt = torch.bincount(v_of, minlength=nv)
        n_rows = n_local + nv
        self.n_rows = n_rows
        mip = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        mip[1:] = torch.cumsum(torch.cat([kcnt, vcnt]), 0)
        mcol = torch.cat([kcol, vcol]).to(torch.int32)
        mval = torch.cat([kval, vval])
        oma = 1.0 - self.alpha
        rd_y = torch.cat([deg_ext[:n_local].to(torch.float32), torch.full((nv,), oma, dtype=torch.float32, device=dev)])
        rd_z = torch.cat([deg_ext[:n_local].to(torch.float32), torch.full((nv,), oma * oma, dtype=torch.float32, device=dev)])
        order = torch.sort(mip[1:] - mip[:-1], descending=True, stable=True).indices
        plan_y = build_stream_plan(mip, mcol, mval, chunk_edges, order, row_deg=rd_y)
        import dataclasses
        plan_z = dataclasses.replace(plan_y, row_deg=rd_z.contiguous(),
                                     fix_deg=rd_z[plan_y.fix_row.to(torch.int64)].contiguous(), _struct=None)
        self.main_y, self.main_z = _SubGraph(plan_y, step_fn), _SubGraph(plan_z, step_fn)
        # ---- who contributes to whom: hub ids I add to, grouped by owner -> the owners learn their contributors
        vh_global = topo.halo_cols[vh_old - n_local]
        v_owner = torch.searchsorted(bnd, vh_global, right=True) - 1
        out_counts = torch.bincount(v_owner, minlength=P) if nv else torch.zeros(P, dtype=torch.int64, device=dev)
        in_counts = torch.empty(P, dtype=torch.int64, device=dev)
        dist.all_to_all_single(in_counts, out_counts, group=group)
        ic, oc = [int(x) for x in in_counts.tolist()], [int(x) for x in out_counts.tolist()]
        contrib = torch.empty(sum(ic), dtype=torch.int64, device=dev)      # global ids of MY hubs, grouped by contributor
        dist.all_to_all_single(contrib, vh_global.contiguous(), output_split_sizes=ic, input_split_sizes=oc, group=group)
        n_ps = int(contrib.numel())
        self.n_pslots = n_ps
        ps_base = n_local + nv + nh                                       # first partial slot of MY buffer
        # each contributor learns where its block of slots starts in my buffer
        offs = torch.tensor([ps_base + sum(ic[:q]) for q in range(P)], dtype=torch.int64, device=dev)
        their = torch.empty(P, dtype=torch.int64, device=dev)
        dist.all_to_all_single(their, offs, group=group)                 # their[q]: start of my block in q's buffer
        # ---- combine CSR: hub row <- its partial slots
        crow = contrib - lo
        if n_ps:
            assert int(crow.min()) >= 0 and int(crow.max()) < n_local and bool(hub_ext[crow].all())
        cslot = ps_base + torch.arange(n_ps, device=dev, dtype=torch.int64)
        cperm = torch.sort(crow, stable=True).indices
        ccnt = torch.bincount(crow, minlength=n_rows) if n_ps else torch.zeros(n_rows, dtype=torch.int64, device=dev)
        cip = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        cip[1:] = torch.cumsum(ccnt, 0)
        combined = torch.nonzero(ccnt > 0).flatten()
        self.combined = combined
        self.comb_y = self.comb_z = None
        if n_ps:
            cplan_y = build_stream_plan(cip, cslot[cperm].to(torch.int32), None, chunk_edges, combined, subset=True, row_deg=rd_y)
            cplan_z = dataclasses.replace(cplan_y, row_deg=rd_z.contiguous(),
                                          fix_deg=rd_z[cplan_y.fix_row.to(torch.int64)].contiguous(), _struct=None)
            self.comb_y, self.comb_z = _SubGraph(cplan_y, step_fn), _SubGraph(cplan_z, step_fn)
        # ---- halo lists over the reduced halo
        topo2 = ShardTopology(rank=rank, world=P, bounds=list(topo.bounds), n_local=n_local, indptr=mip[: n_local + 1],
                              indices=mcol[: int(mip[n_local])], halo_cols=halo_global, recv_counts=[int(c) for c in recv_counts],
                              interior=torch.zeros(n_local, dtype=torch.bool, device=dev))
        self.hx = HaloExchange(topo2, group)
        mine = torch.tensor(recv_counts + [n_local + nv, n_local + nv + nh + n_ps], dtype=torch.int64, device=dev)
        allv = torch.empty(P * (P + 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allv, mine, group=group)
        allv = allv.view(P, P + 2).cpu()
        self.dst_off = [int(allv[q, P]) + int(allv[q, :rank].sum()) for q in range(P)]     # my rows in q's halo region
        self.rows_alloc = int(allv[:, P + 1].max())
        if self.rows_alloc >= (1 << 28):
            raise ValueError("shard too large for the 28-bit slot field of the push code")
        rows, codes, self.soffs, o = [], [], [], 0
        for q in range(P):
            self.soffs.append(o)
            ns = self.hx.send_counts[q]
            if ns:
                rows.append(self.hx.send_idx[o: o + ns])
                codes.append((q << 28) + self.dst_off[q] + torch.arange(ns, device=dev, dtype=torch.int64))
            o += ns
        rows = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64, device=dev)
        codes = torch.cat(codes) if codes else torch.zeros(0, dtype=torch.int64, device=dev)
        is_comb = torch.zeros(n_rows, dtype=torch.bool, device=dev)
        is_comb[combined] = True
        late = is_comb[rows] if rows.numel() else torch.zeros(0, dtype=torch.bool, device=dev)
        # virtual rows: one destination each, the slot the owner reserved for (me, hub)
        vrows = n_local + torch.arange(nv, device=dev, dtype=torch.int64)
        if nv:
            first_of_owner = torch.cumsum(out_counts, 0) - out_counts
            vcodes = (v_owner << 28) + their[v_owner] + (torch.arange(nv, device=dev, dtype=torch.int64) - first_of_owner[v_owner])
        else:
            vcodes = torch.zeros(0, dtype=torch.int64, device=dev)
        self.push_input = _push_arrays(n_rows, rows, codes, dev)                              # halo of the caller's input
        self.push_main = _push_arrays(n_rows, torch.cat([rows[~late], vrows]), torch.cat([codes[~late], vcodes]), dev)
        self.push_comb = _push_arrays(n_rows, rows[late], codes[late], dev)
        self.handles, self._bases = {}, {}
        self._C = C
        self.transport_name = "hybrid-push" if self.on_gpu else "hybrid-p2p"
        self.phases, self.rounds = "hub rows summed where their columns live; two launches and two barriers per step", []
        self.plans = [self.main_y] + ([self.comb_y] if self.comb_y is not None else [])
        self.stats = {"n_local": n_local, "n_virtual": nv, "n_halo": nh, "n_halo_1d": topo.n_halo, "n_pslots": n_ps,
                      "hub_rows_local": int(hub_ext[:n_local].sum()), "combined_rows": int(combined.numel())}

    # buffers, barrier and the input halo are those of the 1-D fused form
    alloc = FusedPushPropagation.alloc
    _barrier = FusedPushPropagation._barrier

    def n_ext(self):
        return self.rows_alloc

    def _push_cpu(self, buf, lst):
        """gloo stand-in for the in-kernel peer stores of one launch: (slot, row) pairs, peer by peer."""
        P, rank = self.topo.world, self.topo.rank
        peer_of, slot_of = lst["codes"] >> 28, lst["codes"] & 0x0FFFFFFF
        out_n = torch.bincount(peer_of, minlength=P) if peer_of.numel() else torch.zeros(P, dtype=torch.int64)
        in_n = torch.empty(P, dtype=torch.int64)
        dist.all_to_all_single(in_n, out_n, group=self.group)
        perm = torch.sort(peer_of, stable=True).indices if peer_of.numel() else peer_of
        F = buf.shape[1]
        send_slots, send_rows = slot_of[perm].contiguous(), buf[lst["rows"][perm]].contiguous()
        isz, osz = [int(x) for x in in_n.tolist()], [int(x) for x in out_n.tolist()]
        rs = torch.empty(sum(isz), dtype=torch.int64)
        rd = torch.empty((sum(isz), F), dtype=buf.dtype)
        dist.all_to_all_single(rs, send_slots, output_split_sizes=isz, input_split_sizes=osz, group=self.group)
        dist.all_to_all_single(rd, send_rows, output_split_sizes=isz, input_split_sizes=osz, group=self.group)
        buf[rs] = rd

    def _push_input(self, buf):
        t = self.topo
        if self.on_gpu:
            from .ops import gather_rows
            h, F = self.handles[buf.data_ptr()]
            for d in range(1, t.world):
                q = (t.rank + d) % t.world
                ns = self.hx.send_counts[q]
                if ns:
                    peer = h.get_buffer(q, (self.rows_alloc, F), torch.float32)
                    gather_rows(buf[: t.n_local], self.hx.send_idx[self.soffs[q]: self.soffs[q] + ns],
                                peer[self.dst_off[q]: self.dst_off[q] + ns])
            self._barrier(buf)
        else:
            self._push_cpu(buf, self.push_input)

    def transfers_only(self, buf):
        if self.topo.world > 1:
            self._push_input(buf)

    def _launch(self, sub, src, T, dst, alpha, epi, use_vals, lst):
        if self._step_fn is not None:
            self._step_fn(sub.plan, src, T, dst, alpha, epi, use_vals)
            if lst is not None:
                self._push_cpu(dst, lst)
            return
        from . import _lib
        lib = _lib.load()
        F = src.shape[1]
        plan = sub.plan
        partial = None
        if plan.n_slots:
            partial = sub._partial.get(F)
            if partial is None:
                partial = torch.empty(plan.n_slots * F, dtype=torch.float32, device=src.device)
                sub._partial[F] = partial
        if lst is not None:
            rc = lib.ppnp_spmm_step_push(plan.struct(), _lib.ptr(src), _lib.ptr(T), _lib.ptr(dst), _lib.ptr(partial), F, F,
                                         float(alpha), int(epi), int(bool(use_vals)), _lib.ptr(lst["ptr"]), _lib.ptr(lst["code"]),
                                         _lib.ptr(lst["first"]), self._bases[dst.data_ptr()], self.topo.world, _lib.current_stream())
        else:
            rc = lib.ppnp_spmm_step(plan.struct(), _lib.ptr(src), _lib.ptr(T), _lib.ptr(dst), _lib.ptr(partial), F, F,
                                    float(alpha), int(epi), int(bool(use_vals)), _lib.current_stream())
        _lib.check(rc, "ppnp_spmm_step_push")

    def propagate(self, H_ext, Z_ext, S_ext, K, alpha):
        """H_ext rows past n_local must be zero on entry apart from what this call ships (the teleport rows of
        the virtual rows are read as zeros)."""
        from . import _lib
        if K < 2:
            raise NotImplementedError("the hybrid form runs the value-free iteration (K >= 2)")
        if abs(float(alpha) - self.alpha) > 1e-12:
            raise ValueError("alpha is baked into the virtual rows' epilogue: build the object with the alpha you propagate with")
        t = self.topo
        multi = t.world > 1
        if multi:
            self._barrier(H_ext)          # entry barrier, see FusedPushPropagation.propagate
            self._push_input(H_ext)
        src = H_ext
        for k in range(1, K + 1):
            dst = Z_ext if (K - k) % 2 == 0 else S_ext
            if k == 1:
                epi, use_vals, main, comb = _lib.EPI_Z2Y, True, self.main_z, self.comb_z
            elif k == K:
                epi, use_vals, main, comb = _lib.EPI_Y2Z, False, self.main_z, self.comb_z
            else:
                epi, use_vals, main, comb = _lib.EPI_Y, False, self.main_y, self.comb_y
            self._launch(main, src, H_ext, dst, alpha, epi, use_vals, self.push_main if multi else None)
            if multi:
                self._barrier(dst)                       # every partial row has landed
                if comb is not None:
                    last = k == K
                    self._launch(comb, dst, dst, dst, alpha, epi | _lib.EPI_ACC | _lib.EPI_INPLACE, False,
                                 None if last else self.push_comb)
                if k < K:
                    self._barrier(dst)                   # every finished hub row has landed
            src = dst
        return Z_ext[: t.n_local]


_PUSH_CLASSES = (PipelinedPushPropagation, FusedPushPropagation, HybridPushPropagation)


class _SubGraph:
    def __init__(self, plan, step_fn=None):
        self.plan, self._step_fn = plan, step_fn
        self._partial = {}

    def step(self, Zin, T, Zout, alpha, epi, use_vals):
        if self._step_fn is not None:
            return self._step_fn(self.plan, Zin, T, Zout, alpha, epi, use_vals)
        from . import _lib
        lib = _lib.load()
        F = Zin.shape[1]
        partial = None
        if self.plan.n_slots:
            partial = self._partial.get(F)
            if partial is None:
                partial = torch.empty(self.plan.n_slots * F, dtype=torch.float32, device=Zin.device)
                self._partial[F] = partial
        rc = lib.ppnp_spmm_step(self.plan.struct(), _lib.ptr(Zin), _lib.ptr(T), _lib.ptr(Zout), _lib.ptr(partial),
                                F, F, float(alpha), int(epi), int(bool(use_vals)), _lib.current_stream())
        _lib.check(rc, "ppnp_spmm_step")


# ------------------------------------------------------------------------------ shard generation (bench)
def stripe_relabel(ids, n, world, stripes):
    """Block-cyclic relabelling of the vertex ids: the id space is cut into world * stripes equal
    blocks that are dealt round-robin, so that every contiguous 1/world-th of the NEW id space holds
    `stripes` blocks spread over the whole old range.  R-MAT puts its hubs at the low ids: without
    this, the nnz-balanced contiguous cut gives rank 0 only hubs and the last rank only leaves, and
    the last rank ships 11x the rows rank 0 does (measured, profiles/r01_scaling.md).  A bijection
    when n is a multiple of world * stripes (else the ids are returned unchanged)."""
    if world == 1 or stripes <= 1 or n % (world * stripes) != 0:
        return ids
    B = n // (world * stripes)
    blk = torch.div(ids, B, rounding_mode="floor")
    return (blk % world) * (n // world) + torch.div(blk, world, rounding_mode="floor") * B + ids % B


def auto_stripes(n, world, target_block=4096):
    """Stripes per rank such that a stripe is ~target_block ids and world * stripes divides n (0 if none):
    fine enough that the few hundred thousand hub ids of an R-MAT graph are dealt evenly to all ranks."""
    if world == 1 or n % world != 0:
        return 1
    per = n // world
    for b in range(target_block, 255, -1):
        if per % b == 0:
            return per // b
    return 1


def degree_sort_relabel(w, bounds):
    """Second relabelling, inside every rank's block: ids by descending weight ``w`` (ties by id).  Rows of the feature
    matrices are STORED in id order, so this puts a block's hot rows next to each other in memory: at F = 16 a row is
    64 bytes, half a 128-byte line -- packed by degree the lines of the hot set are all useful, in the generator's
    order every hot row drags a cold line-mate through the L2 (measured on one GPU, 16 M nodes, F = 16: 5.53 -> 4.74
    ms per step; random ids: 7.1; no effect at F = 64 -- profiles/r02_relabel_probe.txt).  Block membership, hence
    the non-zero balance of the cut, does not change.  Returns new_of_old (int64 [n])."""
    n = int(w.numel())
    new_of_old = torch.empty(n, dtype=torch.int64, device=w.device)
    for p in range(len(bounds) - 1):
        lo, hi = bounds[p], bounds[p + 1]
        if hi > lo:
            order = torch.sort(w[lo:hi], descending=True, stable=True).indices
            new_of_old[lo + order] = torch.arange(lo, hi, device=w.device, dtype=torch.int64)
    return new_of_old


def auto_row_cost(world):
    """Cost of a ROW of a shard in units of one stored entry, for the cut of the row blocks: a row costs its teleport
    read, its result write and its self gather whatever its degree (~4 entries' worth at F = 16) plus the pushes of
    its result to the peers that reference it (~6 entries' worth each, 1.4 peers per row at 8 ranks).  Fitted to
    the 8-GPU run of profiles/r02_bench_n8_degree_sorted.jsonl: with the cut by non-zeros alone the last rank owns
    16.3 M rows and ships 21.7 M (the first: 9.5 M and 15.0 M) and every step waits for it -- 244 M entries + 12.4 x
    16.3 M rows = 446 M units there against a mean of 398 M."""
    if world <= 1:
        return 0.0
    return 4.0 + 6.0 * 1.6 * (1.0 - 1.0 / world)


def rmat_shard(n, raw_draws, scale, seed, dev, rank, world, batch=1 << 26, group=None, stripes=0, degree_sort=True,
               return_relabel=False, row_cost=None):
    """Rows of A + I of the R-MAT graph owned by ``rank`` (global column ids) and the partition.
    Pass 1 estimates the per-row weight from the raw draws to place the boundaries by non-zeros;
    pass 2 keeps the (de-duplicated) edges whose row falls inside this rank's block.  Vertex ids
    are the striped relabelling of the generator's ids (``stripe_relabel``) followed, with
    ``degree_sort``, by ``degree_sort_relabel`` inside every block (both are the partitioner's choice of
    row order; ``return_relabel=True`` also returns new_of_old [n] (int64, device) from the
    generator's ids to the ids used here, for callers that hold data in the generator's order).
    ``row_cost``: per-row weight added to the row's non-zeros when the blocks are cut (None: ``auto_row_cost``)."""
    from . import _lib
    lib = _lib.load()
    if stripes == 0:
        stripes = auto_stripes(n, world)
    striped = world > 1 and stripes > 1 and n % (world * stripes) == 0

    def keys_of(e0, e1, second=None):
        k = torch.empty(2 * (e1 - e0), dtype=torch.int64, device=dev)
        rc = lib.ppnp_rmat_keys(int(seed), int(scale), int(n), int(e0), int(e1), _lib.ptr(k), _lib.current_stream())
        _lib.check(rc, "ppnp_rmat_keys")
        k = k[k >= 0]
        if striped or second is not None:
            r, c = k >> 32, k & 0xFFFFFFFF
            if striped:
                r, c = stripe_relabel(r, n, world, stripes), stripe_relabel(c, n, world, stripes)
            if second is not None:
                r, c = second[r], second[c]
            k = (r << 32) | c
        return k

    # pass 1: this rank histograms its slice of the draws, all-reduce -> approximate degrees
    w = torch.ones(n, dtype=torch.int32, device=dev)       # the self loop
    per = (raw_draws + world - 1) // world
    a, b = rank * per, min(raw_draws, (rank + 1) * per)
    w_part = torch.zeros(n, dtype=torch.int32, device=dev)
    for e0 in range(a, b, batch):
        k = keys_of(e0, min(b, e0 + batch))
        w_part += torch.bincount(k >> 32, minlength=n).to(torch.int32)
        del k
    if world > 1:
        dist.all_reduce(w_part, group=group)
    w += w_part
    del w_part
    rc_ = auto_row_cost(world) if row_cost is None else float(row_cost)
    bounds = balanced_row_blocks(w if rc_ == 0 else w.to(torch.float64) + rc_, world)
    second = degree_sort_relabel(w, bounds) if degree_sort else None
    del w
    lo, hi = bounds[rank], bounds[rank + 1]
    # pass 2: all draws, keep my rows
    kept = [(torch.arange(lo, hi, device=dev, dtype=torch.int64) << 32) | torch.arange(lo, hi, device=dev, dtype=torch.int64)]
    pending = 0
    for e0 in range(0, raw_draws, batch):
        k = keys_of(e0, min(raw_draws, e0 + batch), second)
        k = k[(k >= (lo << 32)) & (k < (hi << 32))]
        kept.append(k)
        pending += k.numel()
        if pending > (1 << 28):                              # compact now and then to bound memory
            kept = [torch.unique(torch.cat(kept))]
            pending = 0
    keys = torch.unique(torch.cat(kept), sorted=True)
    del kept
    rows = (keys >> 32) - lo
    cols = (keys & 0xFFFFFFFF)
    del keys
    counts = torch.bincount(rows, minlength=hi - lo)
    indptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    if return_relabel:
        ids = torch.arange(n, device=dev, dtype=torch.int64)
        new_of_old = stripe_relabel(ids, n, world, stripes) if striped else ids
        if second is not None:
            new_of_old = second[new_of_old]
        return indptr, cols, bounds, new_of_old
    return indptr, cols, bounds


def global_dinv(indptr_local, bounds, rank, world, dev, group=None):
    """D^-1/2 for ALL rows (fp32 [n]) from every rank's exact local degrees (all_gather)."""
    n = bounds[-1]
    deg_local = (indptr_local[1:] - indptr_local[:-1]).to(torch.float32)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    if world == 1:
        out.copy_(deg_local)
    else:
        # blocks have different lengths: gather padded to the longest block, then slice
        width = max(bounds[r + 1] - bounds[r] for r in range(world))
        mine = torch.ones(width, dtype=torch.float32, device=dev)
        mine[: deg_local.numel()] = deg_local
        allb = torch.empty(world * width, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allb, mine, group=group)
        for r in range(world):
            out[bounds[r]: bounds[r + 1]] = allb[r * width: r * width + (bounds[r + 1] - bounds[r])]
    return 1.0 / torch.sqrt(out)


def bench_partitioned(wl, n, raw, scale, F, K, alpha, steps, warmup, dev, rank, world, phases="peer", transport="auto", stripes=0, row_groups=4,
                      carve=None, hub_degree=64, idx16=False, check_small=None, rows_below=None, rows_order="dest", window=None,
                      degree_sort=True, row_cost=None):
    """bench.py's multi-GPU leg: strong scaling of one pass (K forward + K backward steps) on the
    row-partitioned graph.  Times on the device with CUDA events, max over ranks."""
    import time
    t0 = time.perf_counter()
    if stripes == 0:
        stripes = auto_stripes(n, world)
    indptr, cols, bounds = rmat_shard(n, raw, scale, 0, dev, rank, world, stripes=stripes, degree_sort=degree_sort, row_cost=row_cost)
    dinv = global_dinv(indptr, bounds, rank, world, dev)
    topo = build_shard_topology(indptr, cols, bounds, rank)
    del cols
    if (carve or window) and transport not in ("auto", "fused"):
        raise ValueError("carved / window-ordered shard streams exist for the fused transport (--transport fused)")

    def make_prop(topo_, dinv_, carve_):
        # the fused class also runs a single shard (no pushes, no barriers): T_1 of the scaling study then goes through
        # exactly the kernels of the N > 1 runs (edge stream for the rows of high degree + rows kernel for the rest)
        if transport in ("auto", "fused") and (world > 1 or rows_below or carve_ or window):
            try:
                pr = FusedPushPropagation(topo_, dinv_, carve=carve_, idx16=idx16, rows_below=rows_below, rows_order=rows_order,
                                          window=window)
                pr.alloc(4, 1)                                  # peer mappings must be obtainable on this box
                return pr
            except Exception as e:  # noqa: BLE001
                if transport == "fused":
                    raise
                import warnings
                warnings.warn(f"peer-memory transport unavailable ({type(e).__name__}: {e}); falling back to NCCL point-to-point")
                return PartitionedPropagation(topo_, dinv_, phases="one", transport="p2p")
        if transport == "hybrid" and world > 1:
            return HybridPushPropagation(topo_, dinv_, hub_degree=hub_degree, alpha=alpha)
        if transport == "pipe" and world > 1:
            return PipelinedPushPropagation(topo_, dinv_, row_groups=row_groups)
        return PartitionedPropagation(topo_, dinv_, phases=phases, transport=("p2p" if transport in ("pipe", "fused", "hybrid") else transport))

    def alloc_of(pr, count):
        return pr.alloc(F, count) if isinstance(pr, _PUSH_CLASSES) else pr.transport.alloc(F, count)

    prop = make_prop(topo, dinv, carve)
    del dinv
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    H, G, Z, S, Z2 = alloc_of(prop, 5)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    for b in (H, G, Z, S, Z2):
        b.zero_()
    H[: topo.n_local].normal_(generator=g)
    G[: topo.n_local].normal_(generator=g)

    def one_pass():
        prop.propagate(H, Z, S, K, alpha)
        prop.propagate(G, Z, S, K, alpha)

    for _ in range(warmup):
        one_pass()
    torch.cuda.synchronize()
    dist.barrier()
    import bench as _bench
    sampler = _bench.ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    # ---- parity at full size: adjointness <P(H), G> == <H, P(G)> over all ranks (A_hat is symmetric; a halo row that
    # failed to arrive, a wrong push list or a wrong degree breaks it), fp64 dot products, all-reduced
    nl = topo.n_local
    zf = prop.propagate(H, Z, S, K, alpha)
    lhs = (zf.double() * G[:nl].double()).sum()
    zb = prop.propagate(G, Z2, S, K, alpha)
    rhs = (H[:nl].double() * zb.double()).sum()
    dots = torch.stack([lhs, rhs])
    dist.all_reduce(dots)
    adj_rel = float((dots[0] - dots[1]).abs() / dots[0].abs().clamp_min(1e-30))

    # ---- e2e: host shards in, host results out; copy-in, compute and copy-out on three streams like the one-GPU arm
    # (the next input rides under the running propagation, the previous result leaves under the next one)
    Hh = torch.empty((nl, F), dtype=torch.float32, pin_memory=True).copy_(H[:nl])
    Gh = torch.empty((nl, F), dtype=torch.float32, pin_memory=True).copy_(G[:nl])
    Zh = torch.empty((nl, F), dtype=torch.float32, pin_memory=True)
    dHh = torch.empty((nl, F), dtype=torch.float32, pin_memory=True)
    s_in, s_out, cur = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    last = {"fwd": None, "bwd": None, "z": None, "dh": None}

    def one_pass_e2e():
        with torch.cuda.stream(s_in):
            if last["fwd"] is not None:
                s_in.wait_event(last["fwd"])              # H is free once the previous forward has finished with it
            H[:nl].copy_(Hh, non_blocking=True)
            eH = torch.cuda.Event(); eH.record(s_in)
            if last["bwd"] is not None:
                s_in.wait_event(last["bwd"])
            G[:nl].copy_(Gh, non_blocking=True)
            eG = torch.cuda.Event(); eG.record(s_in)
        cur.wait_event(eH)
        if last["z"] is not None:
            cur.wait_event(last["z"])                     # Z is free once its previous read-back has finished
        prop.propagate(H, Z, S, K, alpha)
        f = torch.cuda.Event(); f.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(f)
            Zh.copy_(Z[:nl], non_blocking=True)
            zo = torch.cuda.Event(); zo.record(s_out)
        cur.wait_event(eG)
        if last["dh"] is not None:
            cur.wait_event(last["dh"])
        prop.propagate(G, Z2, S, K, alpha)
        b = torch.cuda.Event(); b.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(b)
            dHh.copy_(Z2[:nl], non_blocking=True)
            do = torch.cuda.Event(); do.record(s_out)
        last.update(fwd=f, bwd=b, z=zo, dh=do)

    def drain():
        cur.wait_stream(s_in)
        cur.wait_stream(s_out)

    one_pass_e2e()
    drain()
    torch.cuda.synchronize()
    dist.barrier()
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    e2e_steps = max(3, min(int(steps), 8))
    e0.record()
    for _ in range(e2e_steps):
        one_pass_e2e()
    drain()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = torch.tensor([e0.elapsed_time(e1) / e2e_steps], device=dev)
    dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_ok = bool(torch.equal(Zh, Z[:nl].cpu()) and torch.equal(dHh, Z2[:nl].cpu()))
    del Hh, Gh, Zh, dHh

    # diagnostics: the transfers alone and the compute alone (not part of the headline number)
    def timed(fn, reps=4):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    from . import _lib as _l

    def transfers_only():
        if isinstance(prop, _PUSH_CLASSES):
            if world > 1:
                prop.transfers_only(Z)
            return
        prop.transport.step_barrier(Z)
        for rnd in (prop.rounds or ([(1, None)] if world > 1 else [])):
            prop._transfer(Z, rnd)
    ms_exchange = timed(transfers_only)

    def compute_only():
        if isinstance(prop, FusedPushPropagation):
            prop._step(Z, H, S, alpha, _l.EPI_Y, False, False)
            return
        first = True
        for p in prop.plans:
            if p is not None:
                acc_pass = (not first) and not isinstance(prop, _PUSH_CLASSES)
                p.step(Z, S if acc_pass else H, S, alpha, _l.EPI_Y | (_l.EPI_ACC if acc_pass else 0), False)
            first = False
    ms_compute = timed(compute_only)

    # per-rank (not max-reduced) transfer time and egress volume
    torch.cuda.synchronize(); dist.barrier()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(4):
        transfers_only()
    b_.record(); torch.cuda.synchronize()
    my_xfer_us = int(a_.elapsed_time(b_) / 4 * 1e3)
    hx = prop.hx if isinstance(prop, _PUSH_CLASSES) else getattr(prop.transport, "hx", None)
    sent_rows = sum(hx.send_counts) if hx is not None else 0
    stats = torch.tensor([int(indptr[-1]), topo.n_halo, topo.n_local, int(topo.interior.sum()), sent_rows, my_xfer_us], dtype=torch.int64, device=dev)
    allstats = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    nnz = int(sum(int(s[0]) for s in allstats))
    launches = 0
    for p in prop.plans:
        if p is not None and p.plan is not None:
            launches += 2 if p.plan.n_fix > 0 else 1
    if getattr(prop, "rows_part", None) is not None:
        launches += 1
    launches += (world - 1) if prop.transport_name in ("pull", "push") else 0
    if isinstance(prop, _PUSH_CLASSES):
        launches += (world - 1) * getattr(prop, "G", 0)
    work = 2 * K * nnz * F
    parity = {"adjointness_rel_full_size": adj_rel, "e2e_results_reached_host": e2e_ok}
    if check_small is not None:     # bench.py's checker: the same code path on a graph its CPU oracle finishes in a second
        parity["small_graph_vs_oracle"] = check_small(make_prop, alloc_of)
    parity["ok"] = bool(adj_rel < 1e-5 and e2e_ok and (check_small is None or parity["small_graph_vs_oracle"]["rel_fro_max_over_ranks"] < 1e-5))
    return {
        "parity": parity,
        "ms_per_step": float(ms), "nnz": nnz, "clocks": clocks,
        "partition": {"rule": f"block-cyclic relabelling ({stripes} stripes per rank), then contiguous row blocks cut at the non-zero prefix sum"
                              + f" of (non-zeros + {auto_row_cost(world) if row_cost is None else row_cost:g} per row)"
                              + (", rows of a block stored in degree order" if degree_sort else ""), "phases": prop.phases,
                      "transport": prop.transport_name,
                      "rows": [int(s[2]) for s in allstats], "nnz": [int(s[0]) for s in allstats],
                      "halo_rows": [int(s[1]) for s in allstats], "interior_rows": [int(s[3]) for s in allstats],
                      "sent_rows": [int(s[4]) for s in allstats], "transfer_us_per_rank": [int(s[5]) for s in allstats]},
        "e2e": {"value": work / (float(ms_e2e) * 1e-3), "unit": "edge*feature/s", "ms_per_step": float(ms_e2e),
                "h2d_bytes_per_step": 2 * n * F * 4, "d2h_bytes_per_step": 2 * n * F * 4},
        "gpu_launches": launches * 2 * K * steps, "graph_build_s": round(t_build, 1),
        "halo_bytes_received_per_step_rank0": topo.n_halo * F * 4,
        "transfers_ms_alone": ms_exchange, "spmm_step_ms_alone": ms_compute,
    }
