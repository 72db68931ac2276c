"""world_size-2 gloo test (CPU) of the only exchange step on the path: the all-gather of per-rank
breakpoint lists, with LPT chromosome ownership."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from canvas_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [500, 100, 300, 200, 50]
    owner = multi.assign_chromosomes_lpt(lengths, world)
    truth = [np.array([0, 10 * (c + 1), 20 * (c + 1)], np.int32) for c in range(len(lengths))]
    mine = [truth[c] if owner[c] == rank else np.zeros(0, np.int32) for c in range(len(lengths))]
    got = multi.all_gather_breakpoints(mine, len(lengths), capacity=256, device="cpu")
    ok = all(g.tolist() == t.tolist() for g, t in zip(got, truth))
    ret[rank] = ok
    dist.destroy_process_group()


def test_all_gather_breakpoints_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.get_context("spawn").Manager()  # no fork() of a process that already runs threads
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


class _FakeEngine:
    """Stands in for the GPU engine: returns the known segmentation of the chromosomes it is asked for."""

    def __init__(self, truth):
        self.truth = truth

    def partition_cbs(self, chrom_off, coverage, chrom_selected=None, **kw):
        segs = [{"len": np.asarray(t, np.int32) if chrom_selected[c] else np.zeros(0, np.int32)} for c, t in enumerate(self.truth)]
        return {"segments": segs}


def _cbs_worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from canvas_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    truth = [[100, 250, 150], [100], [40, 260], [], [5, 45]]
    off = np.concatenate([[0], np.cumsum([sum(t) for t in truth])])
    r = multi.partition_cbs_sharded(_FakeEngine(truth), off, np.zeros(int(off[-1])))
    ret[rank] = [s["len"].tolist() for s in r["segments"]] == truth and sorted(set(r["owner"].tolist())) == [0, 1]
    dist.destroy_process_group()


def test_partition_cbs_sharded_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.get_context("spawn").Manager()  # no fork() of a process that already runs threads
    ret = mgr.dict()
    mp.spawn(_cbs_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


class _FakeHmmEngine:
    def __init__(self, truth):
        self.truth = truth

    def partition_hmm(self, chrom_off, coverage, chrom_selected=None, **kw):
        bps = [np.asarray(t, np.int32) if chrom_selected[c] else np.zeros(0, np.int32) for c, t in enumerate(self.truth)]
        return {"breakpoints": bps, "states": np.zeros(int(chrom_off[-1]), np.uint8)}


def _hmm_worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from canvas_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    truth = [[0, 40, 90], [0], [0, 7], [], [0, 3, 5, 9]]
    off = np.array([0, 500, 600, 900, 905, 1000])
    r = multi.partition_hmm_sharded(_FakeHmmEngine(truth), off, np.zeros(1000))
    ret[rank] = [b.tolist() for b in r["breakpoints"]] == truth and "states" not in r
    dist.destroy_process_group()


def test_partition_hmm_sharded_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.get_context("spawn").Manager()  # no fork() of a process that already runs threads
    ret = mgr.dict()
    mp.spawn(_hmm_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _overflow_worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from canvas_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 alone holds more breakpoints than the capacity both ranks were given: the size is agreed collectively, so
    # nobody raises on its own and nobody is left waiting in the all-gather
    truth = [np.arange(3, dtype=np.int32), np.arange(500, dtype=np.int32) * 2]
    mine = [truth[c] if c == rank else np.zeros(0, np.int32) for c in range(2)]
    got = multi.all_gather_breakpoints(mine, 2, capacity=64, device="cpu")
    ret[rank] = all(g.tolist() == t.tolist() for g, t in zip(got, truth))
    dist.destroy_process_group()


def test_all_gather_breakpoints_beyond_the_default_capacity_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_overflow_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_library_lpt_equals_the_python_one():
    # cg_shard_assign (host code of the library, what the *_sharded entry points use) against multi.assign_chromosomes_lpt
    sys.path.insert(0, ROOT)
    from canvas_b200 import multi, native
    rng = np.random.default_rng(2)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 24, 72, 300):
            w = rng.integers(0, 250000, n)
            w[rng.random(n) < 0.2] = 1000  # ties
            assert native.shard_assign(w, world).tolist() == multi.assign_chromosomes_lpt(w, world).tolist()
