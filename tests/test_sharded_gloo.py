"""World-size-2 `gloo` tests (CPU) of the host-side logic of the sharded index (quake_b200/sharded.py):
partition ownership, probe masking, and the all-gather + merge plumbing. The merge kernel itself
(qk_merge_topk) is covered on the GPU in test_gpu_index.py; here a host merge is injected so that the
collective path runs without a device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quake_b200 import sharded


def _host_merge(part_dist, part_ids, k, metric):
    """(distance, id)-ordered k-way merge on the host; -1 ids are padding."""
    S, Q, _ = part_ids.shape
    desc = metric == 0
    out_i = torch.full((Q, k), -1, dtype=torch.int64)
    out_d = torch.full((Q, k), float("-inf") if desc else float("inf"), dtype=torch.float32)
    for q in range(Q):
        pairs = [(float(part_dist[s, q, j]), int(part_ids[s, q, j])) for s in range(S) for j in range(part_ids.shape[2])
                 if part_ids[s, q, j] >= 0]
        pairs.sort(key=lambda t: ((-t[0]) if desc else t[0], t[1]))
        for j, (dv, iv) in enumerate(pairs[:k]):
            out_i[q, j], out_d[q, j] = iv, dv
    return out_i, out_d


def _pool(metric):
    """A global candidate pool: for every query, 40 (distance, id, partition) triples."""
    g = torch.Generator().manual_seed(11 + metric)
    Q, n = 9, 40
    d = torch.randn(Q, n, generator=g).abs()
    d[:, ::7] = 0.5  # ties across partitions
    ids = torch.stack([torch.randperm(1000, generator=g)[:n] for _ in range(Q)]).to(torch.int64)
    part = torch.randint(0, 6, (Q, n), generator=g)
    return d, ids, part


def _local_topk(d, ids, part, rank, world, k, metric):
    """What a rank would produce: the top-k of the candidates living in the partitions it owns."""
    Q = d.shape[0]
    desc = metric == 0
    oi = torch.full((Q, k), -1, dtype=torch.int64)
    od = torch.full((Q, k), float("-inf") if desc else float("inf"), dtype=torch.float32)
    for q in range(Q):
        mine = [(float(d[q, j]), int(ids[q, j])) for j in range(d.shape[1]) if int(part[q, j]) % world == rank]
        mine.sort(key=lambda t: ((-t[0]) if desc else t[0], t[1]))
        for j, (dv, iv) in enumerate(mine[:k]):
            oi[q, j], od[q, j] = iv, dv
    return oi, od


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for metric in (1, 0):
            k = 7
            d, ids, part = _pool(metric)
            li, ld = _local_topk(d, ids, part, rank, world, k, metric)
            mi, md = sharded.gather_and_merge(li, ld, k, metric, merge=_host_merge)
            wi, wd = _local_topk(d, ids, torch.zeros_like(part), 0, 1, k, metric)  # single-process answer
            ok = ok and torch.equal(mi, wi) and torch.equal(md, wd)
        # every rank must hold the same merged result
        chk = [torch.zeros_like(mi) for _ in range(world)]
        dist.all_gather(chk, mi)
        ok = ok and all(torch.equal(c, mi) for c in chk)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(120)
def test_gather_and_merge_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_ownership_and_probe_masking():
    pids = torch.tensor([[0, 5, 2, -1], [7, 3, 3, 4]], dtype=torch.int64)
    slots = torch.tensor([[10, 11, 12, -1], [13, 14, 14, 15]], dtype=torch.int32)
    assert sharded.owner_of(torch.arange(6), 4).tolist() == [0, 1, 2, 3, 0, 1]
    m0 = sharded.mask_foreign_probes(pids, slots, 0, 2)
    m1 = sharded.mask_foreign_probes(pids, slots, 1, 2)
    assert m0.tolist() == [[10, -1, 12, -1], [-1, -1, -1, 15]]
    assert m1.tolist() == [[-1, 11, -1, -1], [13, 14, 14, -1]]
    # every valid probe is scanned by exactly one rank
    assert torch.equal((m0 >= 0).int() + (m1 >= 0).int(), (slots >= 0).int())


def test_single_process_is_identity():
    ids = torch.arange(12, dtype=torch.int64).reshape(3, 4)
    dd = torch.rand(3, 4)
    a, b = sharded.gather_and_merge(ids, dd, 4, 1)
    assert a is ids and b is dd
