"""Row-partitioned propagation on 2 real GPUs over NCCL against the single-GPU kernel and the C
oracle.  Skipped unless two CUDA devices are visible (run with gpurun --gpus 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import oracle, relerr

pytestmark = pytest.mark.gpu


def worker(rank, world, port, mode, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ppnp_b200 import dist as pd
        n, raw, scale, F, K, alpha = 204_800, 3_000_000, 18, 16, 10, 0.1   # multiple of world * 16 stripes
        indptr, cols, bounds, relabel = pd.rmat_shard(n, raw, scale, 0, dev, rank, world, batch=1 << 20, return_relabel=True)
        dinv = pd.global_dinv(indptr, bounds, rank, world, dev)
        topo = pd.build_shard_topology(indptr, cols, bounds, rank)
        phases, transport = mode.split("/")
        if transport == "fused":
            prop = pd.FusedPushPropagation(topo, dinv)
        elif transport == "fused16":           # 16-byte index staging with the halo push 
            prop = pd.FusedPushPropagation(topo, dinv, idx16=True)
        elif transport == "fusedcarve":        # L2-sized hot column blocks of the shard first 
            prop = pd.FusedPushPropagation(topo, dinv, carve=dict(block_cols=16384, n_blocks=8, min_piece=8))
        elif transport == "fusedwindow":       # rows kernel below `phases` entries, window-ordered hub stream
            prop = pd.FusedPushPropagation(topo, dinv, rows_below=int(phases) or None, window="mid")
        elif transport == "fusedrows":         # low-degree rows through the rows kernel, ordered by destination slot / by degree
            prop = pd.FusedPushPropagation(topo, dinv, rows_below=int(phases.rstrip("d")), rows_order="degree" if phases.endswith("d") else "dest")
        elif transport == "hybrid":            # hub rows summed where their columns live 
            prop = pd.HybridPushPropagation(topo, dinv, hub_degree=int(phases), alpha=alpha)
        elif transport == "pipe":
            prop = pd.PipelinedPushPropagation(topo, dinv, row_groups=int(phases))
        else:
            prop = pd.PartitionedPropagation(topo, dinv, phases=phases, transport=transport)
        lo, hi = bounds[rank], bounds[rank + 1]
        # rows are the striped relabelling of the generator's ids: new id -> old id
        new_of_old = relabel.cpu().numpy()
        old_of_new = np.argsort(new_of_old)
        mine = old_of_new[lo:hi]
        Hg = np.random.RandomState(0).randn(n, F).astype(np.float32)
        H, Z, S = prop.alloc(F, 3) if transport in ("pipe", "fused", "fused16", "fusedcarve", "fusedrows", "fusedwindow", "hybrid") else prop.transport.alloc(F, 3)
        H.zero_()
        H[: topo.n_local] = torch.from_numpy(Hg[mine]).to(dev)
        out = prop.propagate(H, Z, S, K, alpha).cpu().numpy()
        # oracle on the host: the same recipe through the C generator
        ip, idx = oracle.rmat_graph(n, raw, scale, seed=0)
        oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
        assert int(indptr[-1]) == int((oip[mine + 1] - oip[mine]).sum())  # the shard holds exactly its rows of A + I
        ref = oracle.c_appnp_f64(oip, oidx, oval, Hg.astype(np.float64), K, alpha)[mine]
        err = relerr(out, ref)
        assert err < 1e-5, err
        with open(os.path.join(outdir, f"ok_{rank}"), "w") as f:
            f.write(str(err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["x/fused", "4/pipe", "1/pipe", "two/push", "one/push", "peer/pull", "peer/p2p", "one/p2p",
                                  "x/fused16", "x/fusedcarve", "64/hybrid", "512/hybrid", "32/fusedrows", "32d/fusedrows", "1000000/fusedrows",
                                  "0/fusedwindow", "32/fusedwindow"])
def test_partitioned_matches_oracle_on_two_gpus(tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + len(mode) + (hash(mode) % 40) + (os.getpid() % 50)
    mp.spawn(worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok_{r}") for r in range(2))
