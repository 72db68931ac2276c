"""Host side of CanvasBin's counting step (reference CanvasBin/FragmentBinner.cs, HitArray.cs).

BAM decoding and the read-name bookkeeping stay on the host; the geometric part — which bin a fragment
overlaps most — and the per-bin counting run on the GPU through cg_bin_fragments / cg_bin_hits.

Alignments are expected in coordinate order (a sorted, indexed BAM is what BinTask.DoIt jumps into,
FragmentBinner.cs:204-243); that is what makes the reference's monotone `binIndexStart` scan equal to
the kernel's binary search, and what lets the pairing below run without knowing the geometric outcome.
"""
import numpy as np

# filter bits (the same encoding the tests use)
MAPPED, MATE_MAPPED, PRIMARY, PAIRED, PROPER, DUPLICATE, FAILED_QC = 1, 2, 4, 8, 16, 32, 64
MAPQ_NOT_AVAILABLE = 255  # FragmentBinnerConstants.MappingQualityNotAvailable


def flags_from_sam(sam_flag):
    """SAM FLAG -> the filter bits BinOneAlignment looks at (FragmentBinner.cs:259-262, :322-326)."""
    f = np.asarray(sam_flag, np.int64)
    out = np.zeros(f.shape, np.uint8)
    out |= np.where(f & 0x4, 0, MAPPED).astype(np.uint8)
    out |= np.where(f & 0x8, 0, MATE_MAPPED).astype(np.uint8)
    out |= np.where(f & 0x100, 0, PRIMARY).astype(np.uint8)
    out |= np.where(f & 0x1, PAIRED, 0).astype(np.uint8)
    out |= np.where(f & 0x2, PROPER, 0).astype(np.uint8)
    out |= np.where(f & 0x400, DUPLICATE, 0).astype(np.uint8)
    out |= np.where(f & 0x200, FAILED_QC, 0).astype(np.uint8)
    return out


def hit_array(length, positions):
    """HitArray.Set (HitArray.cs:61-64): one saturating byte counter per position."""
    h = np.bincount(np.asarray(positions, np.int64), minlength=length)[:length]
    return np.minimum(h, 255).astype(np.uint8)


def pair_fragments(flags, pos, mate_pos, ref_id, mate_ref_id, frag_len, mapq, names, quality_threshold):
    """The name-dictionary half of BinOneAlignment (FragmentBinner.cs:256-295).

    Returns (frag_start, frag_stop, undo_index): the fragments the left read of every usable pair
    defines, and the indices (into those fragments) whose mate later failed the filters."""
    flags = np.asarray(flags, np.uint8)
    need = MAPPED | MATE_MAPPED | PRIMARY | PAIRED | PROPER
    ok = (flags & need) == need
    mq = np.asarray(mapq, np.int64)
    bad = ((flags & (DUPLICATE | FAILED_QC)) != 0) | (mq == MAPQ_NOT_AVAILABLE) | (mq < quality_threshold)
    fs, fe, undo = [], [], []
    name_to_frag = {}
    same_pos = set()
    pos = np.asarray(pos, np.int64); mate_pos = np.asarray(mate_pos, np.int64)
    for i in np.nonzero(ok)[0].tolist():
        nm = names[i]
        if nm in name_to_frag:
            if bad[i]:
                undo.append(name_to_frag[nm])
            del name_to_frag[nm]
            continue
        if bad[i] or ref_id[i] != mate_ref_id[i] or pos[i] > mate_pos[i]:
            continue
        if pos[i] == mate_pos[i]:
            if nm in same_pos:
                same_pos.remove(nm)
                continue
            same_pos.add(nm)
        if frag_len[i] == 0:
            continue
        name_to_frag[nm] = len(fs)
        fs.append(int(pos[i]))
        fe.append(int(pos[i]) + int(frag_len[i]))
    return np.array(fs, np.int32), np.array(fe, np.int32), np.array(undo, np.int32)


def bin_paired_alignments(engine, flags, pos, mate_pos, ref_id, mate_ref_id, frag_len, mapq, names, quality_threshold,
                          bin_start, bin_stop):
    """BinTask.DoIt's counting for one chromosome: per-bin fragment counts and usableFragmentCount."""
    fs, fe, undo = pair_fragments(flags, pos, mate_pos, ref_id, mate_ref_id, frag_len, mapq, names, quality_threshold)
    r = engine.bin_fragments(fs, fe, bin_start, bin_stop, undo)
    best = r["best_bin"]
    usable = int((best >= 0).sum()) - int((best[undo] >= 0).sum()) if len(best) else 0
    return {"count": r["count"], "usable": usable}


def bin_size_from_rates(counts_per_bin, rates):
    """SampleHitArrays.GetBinSize (CanvasBin.cs:79-83): (int)(countsPerBin / Median(rates)); rates = observed / possible of
    every autosome (GetRates, :30-71, from Engine.bin_screen's counts).  Median = mean of the middles for an even count."""
    import numpy as np
    r = np.sort(np.asarray(rates, np.float64))
    if len(r) == 0:
        med = 0.0
    elif len(r) % 2:
        med = float(r[len(r) // 2])
    else:
        med = (float(r[len(r) // 2 - 1]) + float(r[len(r) // 2])) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        q = float(np.float64(counts_per_bin) / np.float64(med))
    if q != q or abs(q) >= 2.0 ** 31:
        return -2 ** 31  # (int) of NaN, an infinity or an out-of-range double: 0x80000000 on x64 .NET
    return int(q)


def mean_fragment_size(stats):
    """MeanFragmentSize (CanvasBin.cs:164-174): NonZeroMean over the chromosomes of their NonZeroMean fragment length;
    stats = [(sum, count)] per chromosome from Engine.bin_fragment_stats."""
    means = [s // c if c else 0 for s, c in stats]       # Convert.ToInt16(sum / counter): integer division of longs
    pos = [m for m in means if m > 0]
    return sum(pos) // len(pos) if pos else 0


def observed_vs_expected_gc(expected, observed):
    """The ratio table of ComputeObservedVsExpectedGC (CanvasBin.cs:374-387), single precision as in the reference."""
    import numpy as np
    exp = np.array(expected, np.int64)
    obs = np.array(observed, np.int64)
    sum_obs, sum_exp = int(obs.sum()), int(exp.sum())      # taken before the zero counts are replaced by 1
    exp[exp == 0] = 1
    obs[obs == 0] = 1
    with np.errstate(divide="ignore", invalid="ignore"):
        return (obs.astype(np.float32) / exp.astype(np.float32)) * (np.float32(sum_exp) / np.float32(sum_obs))
