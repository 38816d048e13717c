"""Two-GPU tests of the sharded index (NCCL): sharding a built index and the distributed build. They need two
devices and are skipped on a one-GPU box (the round-end run); run them with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import quake_b200 as qb
        from quake_b200.sharded import ShardedQuakeIndex
        out = {}
        for metric in ("l2", "ip"):
            torch.manual_seed(1234)
            n, d, nlist = 40000, 96, 64
            x = torch.randn(n, d)
            ids = torch.arange(n, dtype=torch.int64) + 3
            bp = qb.IndexBuildParams()
            bp.nlist, bp.metric = nlist, metric
            full = qb.QuakeIndex()
            full.build(x, ids, bp)  # deterministic: the same index on every rank
            torch.manual_seed(4321)
            q = torch.randn(100, d)
            sp = qb.SearchParams()
            sp.k, sp.nprobe = 10, 12
            want = full.search(q, sp)
            sh = ShardedQuakeIndex()
            sh.shard_from(full)
            got = sh.search(q, sp)
            out[f"shard_{metric}"] = bool(torch.equal(got.ids, want.ids) and torch.equal(got.distances, want.distances))
            # distributed build from per-rank slices: exhaustive search must equal brute force
            lo, hi = rank * n // world, (rank + 1) * n // world
            sb = ShardedQuakeIndex()
            sb.build(x[lo:hi], ids[lo:hi], bp)
            out[f"ntotal_{metric}"] = sb.ntotal() == n
            sp2 = qb.SearchParams()
            sp2.k, sp2.nprobe = 10, nlist
            r2 = sb.search(q, sp2)
            xs = x / x.norm(dim=1, keepdim=True) if metric == "ip" else x
            if metric == "l2":
                gt = torch.cdist(q, xs).topk(10, largest=False).indices + 3
            else:
                gt = (q @ xs.T).topk(10).indices + 3
            out[f"build_{metric}"] = float((r2.ids == gt).float().mean()) > 0.999
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_index_two_gpus():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert all(ret[r].values()), ret[r]
