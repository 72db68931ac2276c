"""Multi-GPU parity on real hardware: the *_sharded entry points of the C-ABI (LPT chromosome assignment + NCCL all-gather
inside libcanvasgpu) must return, on every rank, exactly what the single-GPU call returns.

  one process per GPU (torchrun; torch.distributed only carries the 128-byte NCCL id):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [scale]
  one process, N GPUs, one host thread per context (cg_comm_init_all — how a single C# host would drive a box):
    python tools/multi_gpu_check.py --single-process 2 [scale]
"""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def same_bp(a, b):
    return all(x.tolist() == y.tolist() for x, y in zip(a, b))


def checks(eng, rank, world, scale):
    """Every comparison of one rank; returns (dict of booleans, info)."""
    from canvas_b200 import synth, textcodec
    s = synth.make_sample(config=2, sample=3, scale=scale, n_events=60)
    ew = max(2000, int(100000 * scale))
    ok = {}
    fused = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, evenness_window=ew)
    fs = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, evenness_window=ew,
                                     sharded=True)
    ok["fused"] = (same_bp(fused["breakpoints"], fs["breakpoints"]) and fused["cv"] == fs["cv"]
                   and np.array_equal(fused["kept_index"], fs["kept_index"]))
    x_ms = eng.last_exchange_ms
    off = fused["chrom_off"]
    cov = textcodec.f2_roundtrip(fused["count"])
    full = eng.partition_wavelet(off, cov, evenness_window=ew)
    shard = eng.partition_wavelet(off, cov, evenness_window=ew, sharded=True)
    ok["wavelet"] = same_bp(full["breakpoints"], shard["breakpoints"]) and np.array_equal(full["factor_of_three"], shard["factor_of_three"])
    ok["owners_cover_ranks"] = sorted(set(shard["owner"].tolist())) == list(range(min(world, len(off) - 1)))
    # more breakpoints than the fixed first round holds: the exact second round
    rng = np.random.default_rng(3)
    n = 300_000
    lev = rng.choice([100., 2000., 4000., 8000., 16000.], n // 15 + 1)
    stair = np.round(np.repeat(lev, 15)[:n] + rng.normal(0, 0.5, n), 2)
    soff = np.array([0, 120_000, 200_000, n])
    a, b = eng.partition_wavelet(soff, stair, evenness_window=20000), eng.partition_wavelet(soff, stair, evenness_window=20000, sharded=True)
    # (run with CANVAS_COMM_PACK_INTS=512 this takes the exact second round; the allgather_lists check below always does)
    ok["wavelet_many_breakpoints"] = same_bp(a["breakpoints"], b["breakpoints"]) and sum(len(x) for x in a["breakpoints"]) > 600
    # CBS on a few chromosomes' worth of bins (permutation tests are heavier)
    ncb = min(len(off) - 1, 6)
    off_c = off[:ncb + 1]
    cov_c = cov[:off_c[-1]]
    fc, sc = eng.partition_cbs(off_c, cov_c), eng.partition_cbs(off_c, cov_c, sharded=True)
    ok["cbs"] = all(np.array_equal(x["len"], y["len"]) and np.array_equal(x["mean"], y["mean"]) for x, y in zip(fc["segments"], sc["segments"]))
    fh, sh = eng.partition_hmm(off, cov, per_sample=True), eng.partition_hmm(off, cov, per_sample=True, sharded=True)
    ok["hmm"] = same_bp(fh["breakpoints"], sh["breakpoints"]) and np.array_equal(fh["states"], sh["states"])
    # config 5's gather: lists of different lengths, one of them beyond the first-round capacity
    mine = np.arange(5 + 40_000 * (rank == world - 1) + 7 * rank, dtype=np.int32) + 1000 * rank
    got = eng.allgather_lists(mine)
    ok["allgather_lists"] = all(np.array_equal(g, np.arange(5 + 40_000 * (r == world - 1) + 7 * r, dtype=np.int32) + 1000 * r)
                                for r, g in enumerate(got))
    buf = (np.arange(1_000_003, dtype=np.float32) * (1 if rank == world - 1 else 0))
    eng.broadcast(buf, world - 1)
    ok["broadcast"] = bool(np.array_equal(buf, np.arange(1_000_003, dtype=np.float32)))
    # the pedigree chain (config 4): Clean once per sample on rank s mod R, GPU-to-GPU broadcast of the cleaned lists, merge,
    # (sample, chromosome) units of the HMM over the ranks, one all-gather -- must equal the one-GPU chain on every rank
    trio = [synth.make_sample(config=4, sample=k, scale=scale, n_events=40) for k in range(3)]
    t0 = trio[0]
    cols = [t.count for t in trio]
    p1 = eng.pedigree_hmm(t0.chrom, t0.is_autosome, t0.is_chr_y, t0.start, t0.stop, cols, t0.gc)
    # a rank only needs the counts of the samples it cleans
    mine_cols = [c if k % world == rank else None for k, c in enumerate(cols)]
    # ... and only the rank that writes the merged files (the last one here) downloads the merged table
    tables = rank == world - 1
    pn = eng.pedigree_hmm(t0.chrom, t0.is_autosome, t0.is_chr_y, t0.start, t0.stop, mine_cols, t0.gc, sharded=True, want_tables=tables)
    ok["pedigree"] = (p1["n_common"] == pn["n_common"]
                      and ((pn["common_index"] is None and pn["count"] is None) if not tables else
                           (np.array_equal(p1["common_index"], pn["common_index"]) and np.array_equal(p1["count"].view(np.uint32), pn["count"].view(np.uint32))))
                      and p1["n_kept"].tolist() == pn["n_kept"].tolist() and p1["local_sd"].tolist() == pn["local_sd"].tolist()
                      and all(same_bp(a, b) for a, b in zip(p1["breakpoints"], pn["breakpoints"]))
                      and sum(len(b) for per in p1["breakpoints"] for b in per) > 0)
    info = {"bins": int(len(cov)), "breakpoints": sum(len(b) for b in full["breakpoints"]), "owners": shard["owner"].tolist(),
            "pedigree_units_per_rank": np.bincount(pn["owner"].ravel(), minlength=world).tolist(),
            "fused_exchange_ms": x_ms}
    return ok, info


def main_torchrun(scale):
    import torch
    import torch.distributed as dist
    from canvas_b200 import native
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = native.Engine(local)
    eng.comm_init_torch()
    ok, info = checks(eng, rank, world, scale)
    flag = torch.tensor([int(v) for v in ok.values()], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-gpu check (one process per GPU): world", world, "nccl", native.load().cg_comm_nccl_version(),
              {k: bool(f) for k, f in zip(ok, flag.tolist())}, info)
    dist.destroy_process_group()
    eng.close()
    return 0 if all(flag.tolist()) else 1


def main_single_process(n, scale):
    import ctypes as C
    from canvas_b200 import native
    engs = [native.Engine(i) for i in range(n)]
    arr = (C.c_void_p * n)(*[e.h for e in engs])
    rc = native.load().cg_comm_init_all(n, arr)
    assert rc == 0, engs[0].lib.cg_last_error(engs[0].h)
    res = [None] * n

    def work(r):
        try:
            res[r] = checks(engs[r], r, n, scale)
        except Exception as e:  # noqa
            res[r] = ({"exception": False}, {"error": repr(e)})

    th = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    good = all(all(r[0].values()) for r in res)
    print("multi-gpu check (one process,", n, "contexts, one thread each):", [r[0] for r in res] if not good else res[0][0], res[0][1])
    [e.close() for e in engs]
    return 0 if good else 1


if __name__ == "__main__":
    a = sys.argv[1:]
    if a and a[0] == "--single-process":
        sys.exit(main_single_process(int(a[1]), float(a[2]) if len(a) > 2 else 0.25))
    sys.exit(main_torchrun(float(a[0]) if a else 0.25))
