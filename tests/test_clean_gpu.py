"""cg_clean (CUDA, through the C-ABI) against the oracle's restatement of CanvasClean.
Bit-exact: kept bins, normalised counts (float32) and the local-SD metric (float64)."""
import numpy as np
import pytest

from canvas_b200 import synth
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def _run_both(engine, s, **kw):
    a = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, **kw)
    b = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, **kw)
    return a, b


def _assert_same(a, b):
    assert len(a["kept_index"]) == len(b["kept_index"]), (len(a["kept_index"]), len(b["kept_index"]))
    assert np.array_equal(a["kept_index"], b["kept_index"])
    bad = np.nonzero(a["count"].view(np.uint32) != b["count"].view(np.uint32))[0]
    assert len(bad) == 0, (len(bad), a["count"][bad[:5]], b["count"][bad[:5]])
    assert a["gc_norm_skipped"] == b["gc_norm_skipped"]
    if np.isnan(b["local_sd"]):
        assert np.isnan(a["local_sd"])
    else:
        assert a["local_sd"] == b["local_sd"], (a["local_sd"], b["local_sd"])


def test_chr20_config1(engine):
    # BASELINE config 1: chr20 at 1 kb (~63k bins): local-SD on (>= 50000), variance step off
    s = synth.make_sample(config=1, chromosomes=["chr20"])
    a, b = _run_both(engine, s)
    _assert_same(a, b)
    assert b["local_sd"] > 0


def test_scaled_genome_all_flag_combinations(engine):
    s = synth.make_sample(config=2, sample=3, scale=0.08)  # ~250k bins, 24 chromosomes
    for size_filter in (True, False):
        for outlier_filter in (True, False):
            for gc_norm in (True, False):
                for want in (True, False):
                    a, b = _run_both(engine, s, size_filter=size_filter, outlier_filter=outlier_filter,
                                     gc_norm=gc_norm, want_local_sd=want)
                    _assert_same(a, b)


def test_full_genome_config2(engine):
    # BASELINE config 2: ~3.1M bins; exercises the > 500000-bin variance branch
    s = synth.make_sample(config=2)
    a, b = _run_both(engine, s)
    _assert_same(a, b)
    assert len(a["kept_index"]) > 2_900_000


def test_variance_normalisation_fires(engine):
    # inflate the spread of a few GC buckets so NormalizeVarianceByGC rescales and the second
    # NormalizeByGC runs (CanvasClean.cs:512-517)
    s = synth.make_sample(config=2, sample=1, scale=0.25)
    rng = np.random.default_rng(7)
    hot = (s.gc >= 44) & (s.gc <= 47)
    noise = rng.normal(0, 60, len(s)).astype(np.float32)
    s.count = np.where(hot, np.maximum(0, s.count + noise), s.count).astype(np.float32)
    a, b = _run_both(engine, s)
    _assert_same(a, b)
    # the oracle without the variance branch gives something else: the branch really ran
    c = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, want_local_sd=False)
    common = np.intersect1d(c["kept_index"], b["kept_index"], return_indices=True)
    assert not np.array_equal(c["count"][common[1]], b["count"][common[2]])


def test_ffpe_like_sample_drops_local_sd_windows(engine):
    # noisy stretches push the local-SD average above 5 so RemoveBinsWithExtremeLocalSD removes bins
    s = synth.make_sample(config=2, sample=2, scale=0.1)
    rng = np.random.default_rng(11)
    cnt = s.count.copy()
    for c in range(len(s.names)):
        idx = np.nonzero(s.chrom == c)[0]
        half = idx[: len(idx) // 2]
        cnt[half] = np.maximum(0, cnt[half] + rng.normal(0, 45, len(half))).astype(np.float32)
    s.count = np.rint(cnt).astype(np.float32)
    a, b = _run_both(engine, s, outlier_filter=False)
    _assert_same(a, b)
    assert b["local_sd"] > 5.0
    plain = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc,
                      outlier_filter=False, want_local_sd=False)
    assert len(b["kept_index"]) < len(plain["kept_index"])


def test_edge_cases(engine):
    names = ["chr1", "chrX"]
    for n in (0, 1, 2, 3, 49, 50, 101, 5000):
        rng = np.random.default_rng(n)
        chrom = np.sort(rng.integers(0, 2, n)).astype(np.uint8)
        start = (np.arange(n) * 1000).astype(np.int32)
        stop = start + 1000 + (rng.integers(0, 3, n) * 500).astype(np.int32)
        count = rng.poisson(100, n).astype(np.float32)
        gc = np.full(n, 40, np.uint8)  # one bucket so the >= 100-bin rule is predictable
        s = synth.Sample(names, chrom, start, stop, count, gc)
        for kw in ({}, {"gc_norm": False}, {"size_filter": False, "outlier_filter": False}):
            a, b = _run_both(engine, s, **kw)
            _assert_same(a, b)


def test_all_zero_counts_and_constant_sizes(engine):
    s = synth.make_sample(config=2, sample=4, scale=0.03)
    s.count[:] = 0
    s.stop = (s.start + 1000).astype(np.int32)
    a, b = _run_both(engine, s)
    _assert_same(a, b)


def test_unsorted_chromosomes_rejected(engine):
    from canvas_b200 import native
    s = synth.make_sample(config=2, sample=5, scale=0.01)
    s.chrom = s.chrom[::-1].copy()
    with pytest.raises(native.CanvasGpuError) as e:
        engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    assert e.value.code == native.CG_ERR_UNSORTED


@pytest.mark.parametrize("n", [40_000, 40_004, 1_000_000 + 12, 2044])
def test_normalize_apply_stream_matches_formula(engine, n):
    # 40_004: the GC array of sample 1 is not 16-byte aligned (plain-load path); 1_000_012: many tiles per CTA
    # plus a ragged end; 2044: shorter than one tile
    rng = np.random.default_rng(5)
    batch = 3
    count = rng.poisson(100, (batch, n)).astype(np.float32)
    gc = rng.integers(0, 101, (batch, n)).astype(np.uint8)
    med = rng.uniform(50, 150, (batch, 101))
    med[:, 7] = 0.0  # a disabled bucket leaves counts untouched
    gmed = rng.uniform(90, 110, batch)
    out, ms = engine.normalize_apply(count, gc, med, gmed, repeats=2)
    exp = count.copy()
    for b in range(batch):
        m = med[b][gc[b]]
        ok = m > 0
        exp[b][ok] = (gmed[b] * count[b][ok].astype(np.float64) / m[ok]).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))
    assert ms > 0


def test_normalize_apply_near_float_rounding_boundaries(engine):
    """The stream kernel multiplies by a rounded reciprocal and falls back to the exact divide near float rounding
    boundaries: quotients placed within +-50 double ulps of float midpoints (and exactly on them) must still round
    as gMed * count / median does in double (CanvasClean.cs:194)."""
    rng = np.random.default_rng(11)
    batch, n = 64, 4096
    f = rng.uniform(0.5, 5000.0, batch).astype(np.float32)
    mid = (f.astype(np.float64) + np.nextafter(f, np.float32(np.inf)).astype(np.float64)) / 2.0  # exact in double
    gmed = rng.uniform(50.0, 150.0, batch)
    base = gmed / mid
    med = np.empty((batch, 101))
    for g in range(101):
        m = base.copy()
        for _ in range(abs(g - 50)):
            m = np.nextafter(m, np.inf if g > 50 else -np.inf)
        med[:, g] = m
    # one sample where the quotient is exactly a midpoint: 3 * (1 + 2^-24) / 3
    gmed[0] = 3.0 * (1.0 + 2.0 ** -24)
    med[0, :] = 3.0
    count = np.ones((batch, n), np.float32)
    gc = np.tile(np.arange(n) % 101, (batch, 1)).astype(np.uint8)
    out, _ = engine.normalize_apply(count, gc, med, gmed, repeats=1)
    exp = (gmed[:, None] * count.astype(np.float64) / np.take_along_axis(med, gc.astype(np.int64), axis=1)).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))
    assert out[0, 0] == np.float32(1.0)  # the tie rounds to even
    assert len(np.unique(exp[1])) >= 2  # the perturbations do straddle a boundary


# ---------------------------------------------------------------------------------------------
# -m LOESS: floating point; the contract is 1e-5 relative on the normalised counts (bucket-wise instead of
# point-wise summation order, device log / exp), kept bins identical
# ---------------------------------------------------------------------------------------------
LOESS_RTOL = 1e-5


def _assert_close_loess(a, b):
    assert np.array_equal(a["kept_index"], b["kept_index"])
    assert a["gc_norm_skipped"] == b["gc_norm_skipped"]
    x, y = a["count"].astype(np.float64), b["count"].astype(np.float64)
    rel = np.abs(x - y) / np.maximum(np.abs(y), 1e-30)
    rel[(x == y)] = 0
    assert np.nanmax(rel) <= LOESS_RTOL, (np.nanmax(rel), int(np.nanargmax(rel)))
    assert np.array_equal(np.isnan(x), np.isnan(y))


@pytest.mark.parametrize("scale,sample", [(0.02, 5), (0.004, 6)])
def test_loess_mode_matches_oracle(engine, scale, sample):
    s = synth.make_sample(config=2, sample=sample, scale=scale)
    a, b = _run_both(engine, s, gc_mode=1, want_local_sd=False)
    _assert_close_loess(a, b)
    # the normalisation did something: counts moved, and low / high GC bins moved in opposite directions to the middle
    assert np.abs(a["count"] - s.count[a["kept_index"]]).max() > 1


def test_loess_mode_with_metric_and_zero_counts(engine):
    s = synth.make_sample(config=1, chromosomes=["chr20"], sample=2)
    s.count[::977] = 0  # log(0) = -Infinity: dropped from the fit, written back as exp(-inf) = 0
    a, b = _run_both(engine, s, gc_mode=1, outlier_filter=False)
    _assert_close_loess(a, b)
    assert (a["count"] == 0).sum() > 10 and a["local_sd"] == b["local_sd"]


def test_weighted_quantiles_for_small_gc_buckets(engine):
    # -w 10 on a small sample: buckets with 10..99 autosomal bins stay and take the weighted median of
    # their neighbours (GetWeightedCounts / WeightedMedian); bit-exact like the rest of the median mode
    s = synth.make_sample(config=2, sample=8, scale=0.002)
    auto = np.array(s.is_autosome, bool)[s.chrom]
    h = np.bincount(s.gc[auto], minlength=101)
    assert ((h >= 10) & (h < 100)).sum() >= 5
    for w in (10, 70, 100):
        a, b = _run_both(engine, s, min_bins_per_gc=w, outlier_filter=False)
        _assert_same(a, b)


def test_weighted_quantiles_with_a_huge_neighbouring_bucket(engine):
    # a sparse GC bucket next to one with far more bins than the shared-memory sorter of the weighted quantiles holds
    # (16384): the reference collects all of the neighbour's values (CanvasClean.cs:107-132), so does the device (global
    # scratch, gc_weighted_big_kernel) — bit-exact, for the median (-w 10 keeps sparse buckets) and for the quartiles of
    # the variance step (> 500000 bins with the metric on)
    rng = np.random.default_rng(31)
    for n, w in ((120_000, 10), (560_000, 100)):
        names = ["chr1", "chr2", "chrX"]
        chrom = np.sort(rng.integers(0, 3, n)).astype(np.uint8)
        start = (np.arange(n) * 1000).astype(np.int32)
        stop = (start + 1000).astype(np.int32)
        gc = np.full(n, 41, np.uint8)
        gc[rng.choice(n, 40, replace=False)] = 40   # sparse, its neighbour 41 holds ~everything
        gc[rng.choice(n, 55, replace=False)] = 43   # sparse, two buckets away from the big one
        gc[rng.choice(n, n // 5, replace=False)] = 47
        count = rng.poisson(100 + (gc.astype(np.int64) - 41) * 3).astype(np.float32)
        s = synth.Sample(names, chrom, start, stop, count, gc)
        a, b = _run_both(engine, s, min_bins_per_gc=w, outlier_filter=False)
        _assert_same(a, b)
        assert len(a["kept_index"]) > n // 2


def test_loess_with_variance_step_uses_weighted_quartiles(engine):
    # > 500000 bins with the metric on: NormalizeVarianceByGC runs on unfiltered GC buckets, the sparse
    # ones through WeightedQuantiles (CanvasClean.cs:57-66)
    s = synth.make_sample(config=2, sample=9, scale=0.18)
    auto = np.array(s.is_autosome, bool)[s.chrom]
    h = np.bincount(s.gc[auto], minlength=101)
    assert ((h[10:90] > 0) & (h[10:90] < 100)).sum() >= 2
    a, b = _run_both(engine, s, gc_mode=1)
    _assert_close_loess(a, b)
    assert a["local_sd"] == b["local_sd"]
