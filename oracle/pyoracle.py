"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libcanvas_oracle.so (CPU restatement of the reference hot path).
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; nothing under
canvas_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcanvas_oracle.so")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, no dependencies)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h", "Makefile"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class CleanOpts(C.Structure):
    _fields_ = [("size_filter", C.c_int), ("outlier_filter", C.c_int), ("gc_norm", C.c_int),
                ("gc_mode", C.c_int), ("want_local_sd", C.c_int), ("min_bins_per_gc", C.c_int)]


class WaveletOpts(C.Structure):
    _fields_ = [("is_germline", C.c_int), ("mad_factor", C.c_double), ("thr_lower", C.c_double),
                ("thr_upper", C.c_double), ("min_size", C.c_int), ("evenness_window", C.c_int),
                ("n_threads", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ora_golden_section_quadratic.restype = C.c_double
        _lib.ora_golden_section_quadratic.argtypes = [C.c_double, C.c_double]
        _lib.ora_median_f32.restype = C.c_double
        _lib.ora_median_f64.restype = C.c_double
        _lib.ora_uh_tree.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def clean(chrom, is_auto, is_chry, start, stop, count, gc, size_filter=True, outlier_filter=True,
          gc_norm=True, gc_mode=0, want_local_sd=True, min_bins_per_gc=100):
    n = len(count)
    chrom = np.ascontiguousarray(chrom, np.uint8)
    is_auto = np.ascontiguousarray(is_auto, np.uint8)
    is_chry = np.ascontiguousarray(is_chry, np.uint8)
    start = np.ascontiguousarray(start, np.int32)
    stop = np.ascontiguousarray(stop, np.int32)
    count = np.ascontiguousarray(count, np.float32)
    gc = np.ascontiguousarray(gc, np.uint8)
    o = CleanOpts(int(size_filter), int(outlier_filter), int(gc_norm), int(gc_mode),
                  int(want_local_sd), int(min_bins_per_gc))
    n_out = C.c_int64(0)
    kept = np.empty(max(n, 1), np.int32)
    out = np.empty(max(n, 1), np.float32)
    lsd = C.c_double(0)
    skipped = C.c_int(0)
    rc = lib().ora_clean(C.byref(o), C.c_int64(n), _p(chrom, C.c_uint8), _p(is_auto, C.c_uint8),
                         _p(is_chry, C.c_uint8), C.c_int(len(is_auto)), _p(start, C.c_int32),
                         _p(stop, C.c_int32), _p(count, C.c_float), _p(gc, C.c_uint8),
                         C.byref(n_out), _p(kept, C.c_int32), _p(out, C.c_float), C.byref(lsd),
                         C.byref(skipped))
    assert rc == 0
    k = n_out.value
    return {"kept_index": kept[:k].copy(), "count": out[:k].copy(), "local_sd": lsd.value,
            "gc_norm_skipped": bool(skipped.value)}


def partition_wavelet(chrom_off, coverage, is_germline=True, mad_factor=5.0, thr_lower=0.05,
                      thr_upper=80.0, min_size=10, evenness_window=100000, n_threads=1):
    chrom_off = np.ascontiguousarray(chrom_off, np.int64)
    coverage = np.ascontiguousarray(coverage, np.float64)
    nc = len(chrom_off) - 1
    n = int(chrom_off[-1])
    o = WaveletOpts(int(is_germline), mad_factor, thr_lower, thr_upper, min_size, evenness_window,
                    n_threads)
    n_bp = np.zeros(nc, np.int32)
    bp = np.zeros(max(n, 1), np.int32)
    ev = C.c_double(0)
    ev_ok = C.c_int(0)
    cv = C.c_double(0)
    cv_has = C.c_int(0)
    f3 = np.zeros(9, np.float64)
    rc = lib().ora_partition_wavelet(C.byref(o), C.c_int(nc), _p(chrom_off, C.c_int64),
                                     _p(coverage, C.c_double), _p(n_bp, C.c_int32),
                                     _p(bp, C.c_int32), C.byref(ev), C.byref(ev_ok), C.byref(cv),
                                     C.byref(cv_has), _p(f3, C.c_double))
    assert rc == 0
    bps = [bp[chrom_off[c]:chrom_off[c] + n_bp[c]].copy() for c in range(nc)]
    return {"breakpoints": bps, "evenness": ev.value if ev_ok.value else None,
            "cv": cv.value if cv_has.value else None, "factor_of_three": f3}


def coverage_variability(window, chrom_off, coverage):
    chrom_off = np.ascontiguousarray(chrom_off, np.int64)
    coverage = np.ascontiguousarray(coverage, np.float64)
    cv = C.c_double(0)
    has = lib().ora_coverage_variability(C.c_int(window), C.c_int(len(chrom_off) - 1),
                                         _p(chrom_off, C.c_int64), _p(coverage, C.c_double),
                                         C.byref(cv))
    return cv.value if has else None


def factor_of_three(chrom_off, coverage):
    chrom_off = np.ascontiguousarray(chrom_off, np.int64)
    coverage = np.ascontiguousarray(coverage, np.float64)
    f3 = np.zeros(9, np.float64)
    lib().ora_factor_of_three(C.c_int(len(chrom_off) - 1), _p(chrom_off, C.c_int64),
                              _p(coverage, C.c_double), _p(f3, C.c_double))
    return f3


def evenness_score(window, chrom_off, coverage):
    chrom_off = np.ascontiguousarray(chrom_off, np.int64)
    coverage = np.ascontiguousarray(coverage, np.float64)
    s = C.c_double(0)
    ok = lib().ora_evenness_score(C.c_int(window), C.c_int(len(chrom_off) - 1),
                                  _p(chrom_off, C.c_int64), _p(coverage, C.c_double), C.byref(s))
    return s.value if ok else None


def haar_wavelets(ratio, thr_lower, thr_upper, is_germline, mad_factor, cv, f3):
    ratio = np.ascontiguousarray(ratio, np.float64)
    f3 = np.ascontiguousarray(f3, np.float64)
    bp = np.zeros(len(ratio), np.int32)
    k = lib().ora_haar_wavelets(C.c_int64(len(ratio)), _p(ratio, C.c_double), C.c_double(thr_lower),
                                C.c_double(thr_upper), C.c_int(int(is_germline)),
                                C.c_double(mad_factor), C.c_int(cv is not None),
                                C.c_double(cv if cv is not None else 0.0), _p(f3, C.c_double),
                                C.c_int(len(f3)), _p(bp, C.c_int32))
    return bp[:k].copy()


def uh_tree(x):
    x = np.ascontiguousarray(x, np.float64)
    n = len(x)
    level = np.zeros(n, np.int32)
    start = np.zeros(n, np.int32)
    brk = np.zeros(n, np.int32)
    end = np.zeros(n, np.int32)
    coef = np.zeros(n, np.float64)
    smooth = C.c_double(0)
    k = lib().ora_uh_tree(C.c_int64(n), _p(x, C.c_double), _p(level, C.c_int32), _p(start, C.c_int32),
                          _p(brk, C.c_int32), _p(end, C.c_int32), _p(coef, C.c_double),
                          C.byref(smooth))
    return {"level": level[:k], "start": start[:k], "brk": brk[:k], "end": end[:k],
            "coef": coef[:k], "smooth": smooth.value}


def loess_train(x, y, bandwidth, robustness_iters, x_step, xq=None):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    fitted = np.zeros(len(x), np.float64)
    xq = np.ascontiguousarray(xq if xq is not None else [], np.float64)
    yq = np.zeros(max(len(xq), 1), np.float64)
    rc = lib().ora_loess_train(C.c_int(len(x)), _p(x, C.c_double), _p(y, C.c_double),
                               C.c_double(bandwidth), C.c_int(robustness_iters), C.c_double(x_step),
                               _p(fitted, C.c_double), C.c_int(len(xq)), _p(xq, C.c_double),
                               _p(yq, C.c_double))
    assert rc == 0
    return fitted, yq[:len(xq)]


def golden_section_quadratic(a, b):
    return lib().ora_golden_section_quadratic(a, b)


def median_f32(x):
    x = np.ascontiguousarray(x, np.float32)
    return lib().ora_median_f32(C.c_int64(len(x)), _p(x, C.c_float))


def median_f64(x):
    x = np.ascontiguousarray(x, np.float64)
    return lib().ora_median_f64(C.c_int64(len(x)), _p(x, C.c_double))


def quartiles_f32(x):
    x = np.ascontiguousarray(x, np.float32)
    q = np.zeros(3, np.float32)
    lib().ora_quartiles_f32(C.c_int64(len(x)), _p(x, C.c_float), _p(q, C.c_float))
    return q


def weighted_quantiles(v, w, probs):
    v = np.ascontiguousarray(v, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    probs = np.ascontiguousarray(probs, np.float32)
    q = np.zeros(len(probs), np.float64)
    lib().ora_weighted_quantiles(C.c_int64(len(v)), _p(v, C.c_float), _p(w, C.c_float),
                                 C.c_int(len(probs)), _p(probs, C.c_float), _p(q, C.c_double))
    return q


def dotnet_sort_levels(counts):
    counts = np.ascontiguousarray(counts, np.int32)
    idx = np.zeros(len(counts), np.int32)
    lib().ora_dotnet_sort_levels(C.c_int(len(counts)), _p(counts, C.c_int32), _p(idx, C.c_int32))
    return idx


def f2_roundtrip(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x), np.float64)
    lib().ora_f2_roundtrip(C.c_int64(len(x)), _p(x, C.c_float), _p(out, C.c_double))
    return out


def bin_hits(hits, possible, bases, bin_size, mode=0, read_gc=None, obs_vs_exp=None):
    hits = np.ascontiguousarray(hits, np.uint8)
    poss = np.ascontiguousarray(possible, np.uint8)
    n = len(hits)
    cap = n // max(bin_size, 1) + 1
    start = np.zeros(cap, np.int32); stop = np.zeros(cap, np.int32); count = np.zeros(cap, np.int32)
    gc = np.zeros(cap, np.uint8)
    rgc = np.ascontiguousarray(read_gc if read_gc is not None else np.zeros(max(n, 1)), np.uint8)
    ratio = np.ascontiguousarray(obs_vs_exp if obs_vs_exp is not None else np.ones(101), np.float32)
    f = lib().ora_bin_hits
    f.restype = C.c_int64
    nb = f(C.c_int64(n), _p(hits, C.c_uint8), _p(poss, C.c_uint8), C.c_char_p(bytes(bases)), C.c_int(bin_size), C.c_int(mode),
           _p(rgc, C.c_uint8), _p(ratio, C.c_float), C.c_int64(cap), _p(start, C.c_int32), _p(stop, C.c_int32),
           _p(count, C.c_int32), _p(gc, C.c_uint8))
    return {"start": start[:nb], "stop": stop[:nb], "count": count[:nb], "gc": gc[:nb]}


def bin_alignments(flags, pos, mate_pos, ref_id, mate_ref_id, frag_len, mapq, name_id, quality_threshold, bin_start, bin_stop):
    a = lambda x, t: np.ascontiguousarray(x, t)
    flags = a(flags, np.uint8); pos = a(pos, np.int32); mate_pos = a(mate_pos, np.int32); ref_id = a(ref_id, np.int32)
    mate_ref_id = a(mate_ref_id, np.int32); frag_len = a(frag_len, np.int32); mapq = a(mapq, np.uint32); name_id = a(name_id, np.int64)
    bs = a(bin_start, np.int32); be = a(bin_stop, np.int32)
    count = np.zeros(max(len(bs), 1), np.int32)
    f = lib().ora_bin_alignments
    f.restype = C.c_int64
    usable = f(C.c_int64(len(flags)), _p(flags, C.c_uint8), _p(pos, C.c_int32), _p(mate_pos, C.c_int32), _p(ref_id, C.c_int32),
               _p(mate_ref_id, C.c_int32), _p(frag_len, C.c_int32), _p(mapq, C.c_uint32), _p(name_id, C.c_int64),
               C.c_uint32(quality_threshold), C.c_int64(len(bs)), _p(bs, C.c_int32), _p(be, C.c_int32), _p(count, C.c_int32))
    return {"count": count[:len(bs)], "usable": int(usable)}


class CbsOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("n_perm", C.c_uint32), ("hybrid", C.c_int), ("min_width", C.c_int), ("k_max", C.c_int),
                ("n_min", C.c_uint32), ("undo", C.c_int), ("seed", C.c_uint32), ("trim", C.c_double), ("undo_sd", C.c_double), ("undo_prune", C.c_double)]


_BDRY = {}


def cbs_boundary(n_perm=10000, alpha=0.01, eta=0.05):
    key = (n_perm, alpha, eta)
    if key not in _BDRY:
        m = int(np.floor(n_perm * alpha) + 1)
        out = np.zeros(m * (m + 1) // 2, np.uint32)
        f = lib().ora_cbs_boundary
        f.restype = C.c_int64
        k = f(C.c_uint32(n_perm), C.c_double(alpha), C.c_double(eta), _p(out, C.c_uint32), C.c_int64(len(out)))
        assert k == len(out)
        _BDRY[key] = out
    return _BDRY[key]


def cbs_tailp(b, delta, m):
    f = lib().ora_cbs_tailp
    f.restype = C.c_double
    return f(C.c_double(b), C.c_double(delta), C.c_int(m))


def mt19937(seed, n):
    out = np.zeros(n, np.uint32)
    lib().ora_mt19937(C.c_uint32(seed), C.c_int64(n), _p(out, C.c_uint32))
    return out


def cbs_tmaxo(x, al0=2):
    x = np.ascontiguousarray(x, np.float64)
    seg = np.zeros(2, np.int32)
    f = lib().ora_cbs_tmaxo
    f.restype = C.c_double
    v = f(_p(x, C.c_double), C.c_int(len(x)), C.c_int(al0), _p(seg, C.c_int32))
    return v, int(seg[0]), int(seg[1])


def cbs_htmaxp(px, k, tss, al0=2):
    px = np.ascontiguousarray(px, np.float64)
    f = lib().ora_cbs_htmaxp
    f.restype = C.c_double
    return f(_p(px, C.c_double), C.c_int(len(px)), C.c_int(k), C.c_double(tss), C.c_int(al0))


def cbs_prune(g, seg_len, cutoff=0.05):
    """ChangePointsPrune (ChangePoint.cs:205-271) on one chromosome: new segment lengths."""
    g = np.ascontiguousarray(g, np.float64)
    ln = np.ascontiguousarray(seg_len, np.int32)
    out = np.zeros(len(ln), np.int32)
    k = lib().ora_cbs_prune(_p(g, C.c_double), C.c_int(len(g)), _p(ln, C.c_int32), C.c_int(len(ln)), C.c_double(cutoff),
                            _p(out, C.c_int32))
    assert k >= 1
    return out[:k].copy()


def partition_cbs(chrom_off, coverage, alpha=0.01, n_perm=10000, hybrid=True, min_width=2, k_max=25, n_min=200, seed=0,
                  sbdry=None, n_threads=1, undo=0, trim=0.025, undo_sd=3.0, undo_prune=0.05):
    off = np.ascontiguousarray(chrom_off, np.int64)
    cov = np.ascontiguousarray(coverage, np.float64)
    nc = len(off) - 1
    n = max(len(cov), 1)
    if sbdry is None:
        sbdry = cbs_boundary(n_perm, alpha, 0.05)
    sbdry = np.ascontiguousarray(sbdry, np.uint32)
    o = CbsOpts(alpha, n_perm, int(hybrid), min_width, k_max, n_min, undo, seed, trim, undo_sd, undo_prune)
    n_seg = np.zeros(max(nc, 1), np.int32)
    seg_len = np.zeros(n, np.int32); seg_mean = np.zeros(n, np.float64)
    first = np.zeros(n, np.int32); last = np.zeros(n, np.int32)
    stats = np.zeros(4, np.int64)
    rc = lib().ora_partition_cbs(C.byref(o), _p(sbdry, C.c_uint32), C.c_int64(len(sbdry)), C.c_int(nc), _p(off, C.c_int64),
                                 _p(cov, C.c_double), _p(n_seg, C.c_int32), _p(seg_len, C.c_int32), _p(seg_mean, C.c_double),
                                 _p(first, C.c_int32), _p(last, C.c_int32), _p(stats, C.c_int64), C.c_int(n_threads))
    assert rc == 0
    segs = []
    for c in range(nc):
        a = int(off[c]); k = int(n_seg[c])
        segs.append({"len": seg_len[a:a + k].copy(), "mean": seg_mean[a:a + k].copy(), "first": first[a:a + k].copy(),
                     "last": last[a:a + k].copy()})
    return {"segments": segs, "tests": int(stats[0]), "perms": int(stats[1]), "perm_steps": int(stats[2]),
            "edge_steps": int(stats[3])}


def cbs_inflation_factor(trim=0.025):
    f = lib().ora_cbs_inflation_factor
    f.restype = C.c_double
    return f(C.c_double(trim))


def cbs_trimmed_variance(x, trim=0.025):
    x = np.ascontiguousarray(x, np.float64)
    f = lib().ora_cbs_trimmed_variance
    f.restype = C.c_double
    return f(_p(x, C.c_double), C.c_int64(len(x)), C.c_double(trim))


class HmmOpts(C.Structure):
    _fields_ = [("n_states", C.c_int), ("per_sample", C.c_int), ("min_size", C.c_int), ("n_threads", C.c_int)]


def partition_hmm(chrom_off, coverage, per_sample=True, min_size=10, n_threads=1):
    """HiddenMarkovModelsRunner.Run.  coverage: [N] (one sample) or [n_samples, N]."""
    chrom_off = np.ascontiguousarray(chrom_off, np.int64)
    cov = np.ascontiguousarray(np.atleast_2d(np.asarray(coverage, np.float64)))
    ns, n = cov.shape
    nc = len(chrom_off) - 1
    assert n == int(chrom_off[-1])
    o = HmmOpts(5, int(per_sample), min_size, n_threads)
    n_bp = np.zeros(max(nc, 1), np.int32)
    bp = np.zeros(max(n, 1), np.int32)
    states = np.zeros(max(n, 1), np.uint8)
    rc = lib().ora_partition_hmm(C.byref(o), C.c_int(ns), C.c_int(nc), _p(chrom_off, C.c_int64), _p(cov, C.c_double),
                                 _p(n_bp, C.c_int32), _p(bp, C.c_int32), _p(states, C.c_uint8))
    assert rc == 0
    return {"breakpoints": [bp[chrom_off[c]:chrom_off[c] + n_bp[c]].copy() for c in range(nc)], "states": states[:n]}


def gamma_ln(z):
    f = lib().ora_gamma_ln
    f.restype = C.c_double
    f.argtypes = [C.c_double]
    return f(float(z))


def negative_binomial(mean, variance, max_value):
    out = np.zeros(max(max_value, 1), np.float64)
    k = lib().ora_negative_binomial(C.c_double(mean), C.c_double(variance), C.c_int(max_value), _p(out, C.c_double))
    return out[:k]


def merge_multi_sample_cleaned(samples):
    """Utilities.MergeMultiSampleCleanedBedFile (CanvasCommon/Utilities.cs:834-920), dictionary for dictionary
    (pure Python: small cases only).  samples: list of lists of (chr, start, stop, count) rows in file order.
    Returns rows (chr, start, stop, [count per sample]) in the reference's output order: chromosomes by first
    appearance over all files, positions by first appearance, kept when every file listed them."""
    chromosomes = []
    for rows in samples:
        for r in rows:
            if r[0] not in chromosomes:
                chromosomes.append(r[0])
    start = {c: {} for c in chromosomes}
    stop = {c: {} for c in chromosomes}
    counts = {c: {} for c in chromosomes}
    for rows in samples:
        for c, a, b, v in rows:
            start[c][a] = a
            stop[c][a] = b
            counts[c].setdefault(a, []).append(np.float32(v))
    out = []
    for c in chromosomes:
        for a in list(start[c].keys()):
            if len(counts[c][a]) < len(samples):
                continue
            if a < 0:
                raise ValueError("Start must be non-negative")
            if a >= stop[c][a]:
                raise ValueError("Start must be less than Stop")
            out.append((c, a, stop[c][a], list(counts[c][a])))
    return out


def median_filter(values, half_window):
    """Utilities.MedianFilter (Utilities.cs:767-791)."""
    v = np.ascontiguousarray(values, np.float32)
    out = np.zeros(max(len(v), 1), np.float32)
    f = lib().ora_median_filter
    f.restype = C.c_int64
    m = f(C.c_int64(len(v)), _p(v, C.c_float), C.c_uint32(half_window), _p(out, C.c_float))
    return out[:m].copy()


def repeated_median_filter(values, max_half_window):
    """RepeatedMedianSmoother (CanvasSmooth.cs:66-77)."""
    v = np.ascontiguousarray(values, np.float32)
    out = np.zeros(max(len(v), 1), np.float32)
    f = lib().ora_repeated_median_filter
    f.restype = C.c_int64
    m = f(C.c_int64(len(v)), _p(v, C.c_float), C.c_uint32(max_half_window), _p(out, C.c_float))
    return out[:m].copy()


def normalize_reference(counts, on_target=None):
    """WeightedAverageReferenceGenerator.Run (:43-70): counts [n_samples, n] doubles -> medians, weights, reference."""
    c = np.ascontiguousarray(np.atleast_2d(np.asarray(counts, np.float64)))
    s, n = c.shape
    on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
    med = np.zeros(s); w = np.zeros(s); ref = np.zeros(max(n, 1))
    lib().ora_normalize_reference(C.c_int(s), C.c_int64(n), _p(c, C.c_double), _p(on, C.c_uint8) if on is not None else None,
                                  _p(med, C.c_double), _p(w, C.c_double), _p(ref, C.c_double))
    return {"median": med, "weight": w, "reference": ref[:n]}


def normalize_ratio(sample, reference, on_target=None, mode="lsnorm", min_ref=1.0, max_ref=np.inf, ploidy=None):
    """LSNormRatioCalculator / RawRatioCalculator + RatiosToCounts: kept bin indices, ratios and counts (float)."""
    a = np.ascontiguousarray(sample, np.float32)
    b = np.ascontiguousarray(reference, np.float32)
    n = min(len(a), len(b))
    on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
    pl = None if ploidy is None else np.ascontiguousarray(ploidy, np.int32)
    idx = np.zeros(max(n, 1), np.int32); ratio = np.zeros(max(n, 1), np.float32); count = np.zeros(max(n, 1), np.float32)
    lsf = C.c_double(0)
    f = lib().ora_normalize_ratio
    f.restype = C.c_int64
    k = f(C.c_int64(n), _p(a, C.c_float), _p(b, C.c_float), _p(on, C.c_uint8) if on is not None else None,
          C.c_int(mode == "lsnorm"), C.c_double(min_ref), C.c_double(max_ref), _p(pl, C.c_int32) if pl is not None else None,
          _p(idx, C.c_int32), _p(ratio, C.c_float), _p(count, C.c_float), C.byref(lsf))
    return {"kept_index": idx[:k].copy(), "ratio": ratio[:k].copy(), "count": count[:k].copy(), "library_size_factor": lsf.value}


def normalize_best_lr2(sample, controls, on_target=None):
    """BestLR2ReferenceGenerator.Run (:32-90): index of the best control, mean squared log ratios, ignored bins."""
    t = np.ascontiguousarray(sample, np.float64)
    c = np.ascontiguousarray(np.atleast_2d(np.asarray(controls, np.float64)))
    s, n = c.shape
    on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
    mean = np.zeros(s); ign = np.zeros(s, np.int64)
    best = lib().ora_normalize_best_lr2(C.c_int(s), C.c_int64(n), _p(t, C.c_double), _p(c, C.c_double),
                                        _p(on, C.c_uint8) if on is not None else None, _p(mean, C.c_double), _p(ign, C.c_int64))
    return {"best": int(best), "mean_sq_log_ratio": mean, "ignored": ign}


def normalize_pca_reference(sample, mu, axes, on_target=None, min_ref=1.0, max_ref=np.inf):
    """PCAReferenceGenerator.Run (:37-78): reference counts (float) and the median ratio; None when the axes are not orthogonal."""
    a = np.ascontiguousarray(sample, np.float32)
    m = np.ascontiguousarray(mu, np.float32)
    ax = np.ascontiguousarray(np.atleast_2d(np.asarray(axes, np.float64)))
    k, n = ax.shape
    on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
    ref = np.zeros(max(n, 1), np.float32)
    med = C.c_double(0)
    rc = lib().ora_normalize_pca_reference(C.c_int64(n), C.c_int(k), _p(a, C.c_float), _p(m, C.c_float), _p(ax, C.c_double),
                                           _p(on, C.c_uint8) if on is not None else None, C.c_double(min_ref), C.c_double(max_ref),
                                           _p(ref, C.c_float), C.byref(med))
    if rc != 0:
        return None
    return {"reference": ref[:n], "median_ratio": med.value}


def bin_screen(hits, possible, filter_start=(), filter_stop=()):
    """ExcludeTagsOverlappingFilterFile + ScreenObservedTags + the counts of GetRates on one chromosome."""
    h = np.array(hits, np.uint8)
    p = np.array(possible, np.uint8)
    fs = np.ascontiguousarray(filter_start, np.int32)
    fe = np.ascontiguousarray(filter_stop, np.int32)
    obs, pos = C.c_int64(0), C.c_int64(0)
    lib().ora_bin_screen(C.c_int64(len(h)), _p(h, C.c_uint8), _p(p, C.c_uint8), C.c_int64(len(fs)), _p(fs, C.c_int32), _p(fe, C.c_int32),
                         C.byref(obs), C.byref(pos))
    return {"hits": h, "possible": p.astype(bool), "observed": obs.value, "n_possible": pos.value}


def bin_read_gc(bases, frag_len, mean_frag, hits):
    """Read GC content per position (CanvasBin.cs:450-497) and the chromosome's expected / observed counts per GC bin (:341-358)."""
    b = bytes(bases)
    f = np.ascontiguousarray(frag_len, np.int16)
    h = np.ascontiguousarray(hits, np.uint8)
    n = len(b)
    gc = np.zeros(max(n, 1), np.uint8)
    exp = np.zeros(101, np.int64); obs = np.zeros(101, np.int64)
    lib().ora_bin_read_gc(C.c_int64(n), b, _p(f, C.c_int16), C.c_int(mean_frag), _p(h, C.c_uint8), _p(gc, C.c_uint8),
                          _p(exp, C.c_int64), _p(obs, C.c_int64))
    return {"read_gc": gc[:n], "expected": exp, "observed": obs}


def merge_common_bins_np(samples):
    """The same merge on column arrays (numpy; what bench.py's CPU arm times at 3 x 3 M bins, where the dictionary walk above
    would take minutes): samples = [(chrom_id, start, stop, count) arrays], every sample ordered by (chromosome id, start) and
    the ids shared — then the reference's output order (chromosomes, then positions, by first appearance) is sample 0's order.
    Returns kept_index into sample 0, the stop of the LAST sample (the reference overwrites it per file, Utilities.cs:885) and
    count[n_samples][n_common].  Checked against merge_multi_sample_cleaned in tests/test_host_layer.py."""
    keys = [(np.asarray(c, np.int64) << 32) | np.asarray(a, np.int64) for c, a, _, _ in samples]
    common = keys[0]
    for k in keys[1:]:
        common = common[np.isin(common, k, assume_unique=True)]
    kept = np.flatnonzero(np.isin(keys[0], common, assume_unique=True))
    counts = np.empty((len(samples), len(kept)), np.float32)
    stop = None
    for s, (k, smp) in enumerate(zip(keys, samples)):
        order = np.argsort(k, kind="stable")
        at = order[np.searchsorted(k[order], keys[0][kept])]
        counts[s] = np.asarray(smp[3], np.float32)[at]
        stop = np.asarray(smp[2])[at]
    return {"kept_index": kept, "stop": stop, "count": counts}
