"""Host logic of bench.py and of the k-means filter-precision policy (no GPU).

* bench.agreed_repeats: under N ranks every rank must derive the SAME number of timed regions (each region is bracketed
  by collectives) -- world-size-2 gloo run with deliberately different per-rank probe times.
* clustering.AssignFilter: starts a run with the 2xTF32 filter, falls back to 3xTF32 for good once more than 0.5 % of
  the points of the reviewed calls needed the exact re-scan."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        # 4.6 ms vs 4.8 ms per probe: ceil(100 / 4.6) = 22, ceil(100 / 4.8) = 21 -- on their own the ranks disagree
        probe = 4.6 if rank == 0 else 4.8
        alone = bench.agreed_repeats(probe, 1, torch.device("cpu"))
        together = bench.agreed_repeats(probe, world, torch.device("cpu"))
        got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, torch.tensor([alone, together], dtype=torch.int64))
        ret[rank] = [g.tolist() for g in got]
    finally:
        dist.destroy_process_group()


def test_repeat_count_is_agreed_across_ranks():
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        rows = ret[0]
    assert ret is not None and rows == [[22, 21], [21, 21]]


def test_repeat_count_bounds():
    import bench
    assert bench.agreed_repeats(1000.0, 1, None) == 3      # long probes: at least three regions
    assert bench.agreed_repeats(0.01, 1, None) == 60       # short probes: capped
    assert bench.agreed_repeats(5.0, 1, None) == 20


def test_assign_filter_policy():
    from quake_b200 import clustering
    cpu = torch.device("cpu")
    f = clustering.AssignFilter(cpu)
    assert f.terms == 2 and not f.fixed
    f.review()                      # nothing to review yet
    assert f.terms == 2
    f.points, f.stats[0] = 100_000, 400      # 0.4 % re-scanned: stays
    f.review()
    assert f.terms == 2 and f.points == 0 and int(f.stats[0]) == 0
    f.points, f.stats[0] = 100_000, 501      # > 0.5 %: three terms for the rest of the run
    f.review()
    assert f.terms == 3
    f.points, f.stats[0] = 100_000, 0
    f.review()
    assert f.terms == 3
    g = clustering.AssignFilter(cpu)
    g.points, g.stats[0] = 100, 8            # tiny calls: an absolute floor of 8 points
    g.review()
    assert g.terms == 2
    old = os.environ.get("QK_ASSIGN_FILTER")
    os.environ["QK_ASSIGN_FILTER"] = "3"
    try:
        h = clustering.AssignFilter(cpu)
        assert h.terms == 3 and h.fixed
    finally:
        if old is None:
            del os.environ["QK_ASSIGN_FILTER"]
        else:
            os.environ["QK_ASSIGN_FILTER"] = old
