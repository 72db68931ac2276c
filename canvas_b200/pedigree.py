"""SmallPedigree chain above the C-ABI: CanvasClean per sample -> bins common to every sample -> PerSampleHMM per sample
(reference Canvas/CanvasRunner.cs:883-937: CanvasClean per sample :883-893, NormalizeCanvasClean / Utilities.
MergeMultiSampleCleanedBedFile :895-903 + CanvasCommon/Utilities.cs:834-920, CanvasPartition -m PerSampleHMM :927).

The product call for this chain is Engine.pedigree_hmm = cg_pedigree_hmm (csrc/pedigree.cu): ONE device-resident call that keeps
the cleaned lists in HBM between the stages.  trio_segments below drives the same stages one C-ABI call at a time through
host memory — the form a host that wants the intermediate files would use, and the cross-check of the one-call chain in the
tests.

One GPU: every step on the one engine.  N ranks (BASELINE config 4: "chromosomes sharded across 8 x B200"): sample s is
cleaned ONCE, on rank s mod N, and its cleaned bins are broadcast over NCCL (cg_comm_broadcast) instead of repeating Clean
on every rank; the common-bin merge is cheap and runs on every rank; the 3 x 24 (sample, chromosome) units of the HMM are
assigned longest-processing-time-first over the ranks (cg_shard_assign), each rank segments its units
(cg_partition_hmm_shard), and one all-gather of the packed breakpoint lists (cg_comm_allgather_lists) reassembles the
whole result on every rank.
"""
import numpy as np

from . import native, synth


def bin_positions(sample):
    """File coordinates of a pedigree's shared bin layout: rebuilt from the bin index (1 kb bins) — the synthetic oversized
    bins push the generator's own coordinates past int32 at full scale."""
    off = synth.chrom_offsets(sample.chrom, len(sample.names))
    return ((np.arange(len(sample)) - off[sample.chrom]) * 1000).astype(np.int32)


def trio_segments(eng, samples, sharded=False, timings=None, layout_off=None):
    """samples: synth samples of one pedigree (shared bin layout).  Returns {"breakpoints": [sample][chrom] arrays,
    "chrom_off": offsets of the common bins, "n_common": int, "owner": [sample][chrom] rank of every unit}."""
    import time
    S = len(samples)
    nc = len(samples[0].names)
    rank, world = (eng.comm_rank, eng.comm_size) if sharded else (0, 1)
    kernel_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    # ---- CanvasClean, each sample once
    mine = {}
    for s in range(S):
        if s % world == rank:
            sm = samples[s]
            mine[s] = eng.clean(sm.chrom, sm.is_autosome, sm.is_chr_y, sm.start, sm.stop, sm.count, sm.gc)
            kernel_ms += eng.last_kernel_ms
            launches += eng.last_launches
    t1 = time.perf_counter()
    # ---- the cleaned bins of every sample on every rank
    if world > 1:
        n_out = np.zeros(S, np.int32)
        for s, r in mine.items():
            n_out[s] = len(r["kept_index"])
        got = eng.allgather_lists(n_out)
        n_out = np.max(np.stack(got), axis=0)
    kept_lists, count_lists = [], []
    for s in range(S):
        if world > 1:
            root = s % world
            kept = mine[s]["kept_index"].copy() if root == rank else np.empty(int(n_out[s]), np.int32)
            cnt = mine[s]["count"].copy() if root == rank else np.empty(int(n_out[s]), np.float32)
            kept = eng.broadcast(kept, root)
            cnt = eng.broadcast(cnt, root)
        else:
            kept, cnt = mine[s]["kept_index"], mine[s]["count"]
        kept_lists.append(kept)
        count_lists.append(cnt)
    t2 = time.perf_counter()
    # ---- bins common to every sample (every rank).  The samples share one bin layout, so cg_clean's kept_index lists are
    # the keys (cg_merge_kept_indices); the general, coordinate-keyed form is cg_merge_common_bins
    m = eng.merge_kept_indices(len(samples[0]), kept_lists, count_lists)
    kernel_ms += m["kernel_ms"]
    launches += eng.last_launches
    if layout_off is None:
        layout_off = synth.chrom_offsets(samples[0].chrom, nc)
    off = np.searchsorted(m["common_index"], layout_off).astype(np.int64)  # common bins are in layout order
    n_common = int(len(m["common_index"]))
    t3 = time.perf_counter()
    # ---- PerSampleHMM, (sample, chromosome) units over the ranks
    lens = np.diff(off)
    owner = native.shard_assign(np.tile(lens, S), world).reshape(S, nc)
    local = []
    for s in range(S):
        mask = (owner[s] == rank).astype(np.uint8)
        if not mask.any():
            local.append([np.zeros(0, np.int32)] * nc)
            continue
        # the merged file prints float.ToString(): text_mode 2 reproduces that round trip on the device
        r = eng.partition_hmm_counts(off, m["count"][s], text_mode=2, per_sample=True, chrom_selected=None if world == 1 else mask)
        kernel_ms += r["kernel_ms"]
        launches += r["launches"]
        local.append(r["breakpoints"])
    t4 = time.perf_counter()
    # ---- one all-gather of the packed lists: [sample, chromosome, count, breakpoints ...] per unit
    if world > 1:
        flat = []
        for s in range(S):
            for c in range(nc):
                if owner[s][c] == rank and len(local[s][c]):
                    flat.append(np.concatenate([[s, c, len(local[s][c])], local[s][c]]).astype(np.int32))
        got = eng.allgather_lists(np.concatenate(flat) if flat else np.zeros(0, np.int32))
        bps = [[np.zeros(0, np.int32) for _ in range(nc)] for _ in range(S)]
        for lst in got:
            i = 0
            while i < len(lst):
                s, c, k = int(lst[i]), int(lst[i + 1]), int(lst[i + 2])
                bps[s][c] = lst[i + 3:i + 3 + k].copy()
                i += 3 + k
    else:
        bps = local
    t5 = time.perf_counter()
    if timings is not None:
        for k, v in (("clean", t1 - t0), ("broadcast", t2 - t1), ("merge", t3 - t2), ("hmm", t4 - t3), ("gather", t5 - t4)):
            timings[k] = timings.get(k, 0.0) + v * 1e3
        timings["kernel_ms"] = timings.get("kernel_ms", 0.0) + kernel_ms
        timings["launches"] = timings.get("launches", 0) + launches
    return {"breakpoints": bps, "chrom_off": off, "n_common": n_common, "common_index": m["common_index"], "owner": owner}
