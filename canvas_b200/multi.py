"""torch.distributed form of the partition stage's multi-GPU exchange — for the CPU (gloo) tests of the host logic only.

The PRODUCT path does this inside the library (csrc/comm.cu: cg_comm_init + cg_*_sharded, NCCL bound with dlopen; the
Python binding is Engine.comm_init / sharded=True).  The logic is the same and is what tests/test_multi_cpu.py pins with
world-size-2 gloo runs: chromosomes are independent once the genome-wide scalars exist, so every rank
computes the scalars on the full coverage (replicated, microseconds), segments only the chromosomes assigned to it
(longest-processing-time-first by bin count) and ONE all-gather reassembles the genome-wide breakpoint list on every
rank; the buffer size is agreed collectively first, so a rank with long lists cannot leave the others waiting.
"""
import numpy as np


def assign_chromosomes_lpt(lengths, world):
    """Greedy LPT: longest chromosome first onto the least loaded rank.  Returns owner[c]."""
    lengths = np.asarray(lengths, np.int64)
    owner = np.zeros(len(lengths), np.int32)
    load = np.zeros(world, np.int64)
    for c in np.argsort(-lengths, kind="stable"):
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += lengths[c]
    return owner


def pack_breakpoints(breakpoints, capacity):
    """[count, (chrom, bp) ...] padded to `capacity` int32 entries."""
    flat = np.zeros(capacity, np.int32)
    k = 1
    total = 0
    for c, b in enumerate(breakpoints):
        m = len(b)
        if k + 2 * m > capacity:
            raise ValueError("segment-list buffer too small")
        flat[k:k + 2 * m:2] = c
        flat[k + 1:k + 2 * m:2] = b
        k += 2 * m
        total += m
    flat[0] = total
    return flat


def unpack_breakpoints(buffers, n_chrom):
    """Inverse of pack_breakpoints over the gathered [world, capacity] array: per chromosome, the
    breakpoints reported by whichever rank owned it."""
    out = [np.zeros(0, np.int32) for _ in range(n_chrom)]
    for row in np.asarray(buffers):
        m = int(row[0])
        pairs = row[1:1 + 2 * m].reshape(m, 2)
        for c in np.unique(pairs[:, 0]):
            out[int(c)] = pairs[pairs[:, 0] == c, 1].astype(np.int32)
    return out


def all_gather_breakpoints(breakpoints, n_chrom, capacity=16384, device=None):
    """Every rank contributes the breakpoints of its chromosomes, every rank gets all.  The buffer size is agreed
    collectively first (all_reduce MAX of what each rank needs): a rank whose lists outgrow the default capacity makes
    EVERY rank use the larger buffer, instead of raising alone and leaving the others waiting in the all-gather.
    (The product path does this inside the library: cg_*_sharded / cg_comm_allgather_lists; this torch.distributed form
    serves the CPU tests of the host logic.)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    need = torch.tensor([1 + 2 * sum(len(b) for b in breakpoints)], dtype=torch.int64, device=dev)
    dist.all_reduce(need, op=dist.ReduceOp.MAX)
    capacity = max(int(capacity), int(need.item()))
    local = torch.from_numpy(pack_breakpoints(breakpoints, capacity)).to(dev)
    gathered = torch.empty(world * capacity, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, local)
    return unpack_breakpoints(gathered.cpu().numpy().reshape(world, capacity), n_chrom)


def partition_wavelet_sharded(engine, chrom_off, coverage, **kw):
    """cg_partition_wavelet_shard on this rank's chromosomes + the all-gather.  Every rank returns the
    full per-chromosome breakpoint list (identical to the single-GPU result)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    lengths = np.diff(np.asarray(chrom_off, np.int64))
    owner = assign_chromosomes_lpt(lengths, world)
    mask = (owner == rank).astype(np.uint8)
    r = engine.partition_wavelet(chrom_off, coverage, chrom_selected=mask, **kw)
    r["breakpoints"] = all_gather_breakpoints(r["breakpoints"], len(lengths))
    r["owner"] = owner
    return r


def partition_cbs_sharded(engine, chrom_off, coverage, **kw):
    """cg_partition_cbs_shard on this rank's chromosomes + the same single all-gather: segment lengths travel
    as (chromosome, cumulative end) pairs.  Every chromosome keeps its own random stream, so the merged result
    equals the single-GPU call on every rank."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    lengths = np.diff(np.asarray(chrom_off, np.int64))
    owner = assign_chromosomes_lpt(lengths, world)
    r = engine.partition_cbs(chrom_off, coverage, chrom_selected=(owner == rank).astype(np.uint8), **kw)
    ends = [np.cumsum(s["len"]).astype(np.int32) for s in r["segments"]]
    merged = all_gather_breakpoints(ends, len(lengths))
    r["segments"] = [{"len": np.diff(np.concatenate([[0], e])).astype(np.int32)} for e in merged]
    r["owner"] = owner
    return r


def clean_partition_wavelet_sharded(engine, sample_arrays, lengths_hint, **kw):
    """Fused Clean + wavelet partition of ONE sample over all ranks: every rank cleans the whole sample (genome-wide
    order statistics), segments the chromosomes LPT assigns to it (by input bin count: `lengths_hint`), and one
    all-gather reassembles the breakpoint lists.  sample_arrays = (chrom, is_autosome, is_chr_y, start, stop, count, gc)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    owner = assign_chromosomes_lpt(lengths_hint, world)
    r = engine.clean_partition_wavelet(*sample_arrays, chrom_selected=(owner == rank).astype(np.uint8), **kw)
    r["breakpoints"] = all_gather_breakpoints(r["breakpoints"], len(lengths_hint))
    r["owner"] = owner
    return r


def partition_hmm_sharded(engine, chrom_off, coverage, **kw):
    """cg_partition_hmm_shard on this rank's chromosomes + the single all-gather.  The emission statistics are
    whole-genome (per-sample mode) or per chromosome (joint mode), so the union over ranks equals the single-GPU call."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    lengths = np.diff(np.asarray(chrom_off, np.int64))
    owner = assign_chromosomes_lpt(lengths, world)
    r = engine.partition_hmm(chrom_off, coverage, chrom_selected=(owner == rank).astype(np.uint8), **kw)
    r["breakpoints"] = all_gather_breakpoints(r["breakpoints"], len(lengths))
    r["owner"] = owner
    r.pop("states", None)  # the Viterbi path stays on the rank that computed it
    return r
