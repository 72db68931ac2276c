"""cg_partition_wavelet (CUDA, through the C-ABI) against the oracle's restatement of
CanvasPartition's wavelet branch.  Breakpoints are integer bin indices: compared exactly.  Scalars:
factor-of-three and CV are exact order statistics (compared exactly); evenness sums 10^4-10^5 doubles
in a different order than the reference's sequential LINQ Sum (1e-9 relative, north_star allows 1e-5)."""
import json
import os

import numpy as np
import pytest

from canvas_b200 import synth
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def _cleaned_coverage(s):
    """Partition's input as the reference sees it: CanvasClean output rounded by the .cleaned file."""
    r = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    chrom = s.chrom[r["kept_index"]]
    off = synth.chrom_offsets(chrom, len(s.names))
    cov = ora.f2_roundtrip(r["count"])
    return off, cov


def _compare(a, b, label=""):
    for c, (x, y) in enumerate(zip(a["breakpoints"], b["breakpoints"])):
        assert x.tolist() == y.tolist(), (label, "chromosome", c, x.tolist()[:40], y.tolist()[:40])
    assert (a["cv"] is None) == (b["cv"] is None)
    if b["cv"] is not None:
        assert a["cv"] == b["cv"] or (np.isnan(a["cv"]) and np.isnan(b["cv"])), (a["cv"], b["cv"])
    assert np.array_equal(a["factor_of_three"], b["factor_of_three"], equal_nan=True), (a["factor_of_three"], b["factor_of_three"])
    assert (a["evenness"] is None) == (b["evenness"] is None)
    if b["evenness"] is not None:
        assert abs(a["evenness"] - b["evenness"]) <= 1e-9 * abs(b["evenness"]), (a["evenness"], b["evenness"])


def test_reference_golden_vector(engine, golden_dir):
    # CanvasTest/CanvasPartition/WaveletTests.cs:9-90 through the C-ABI: the CV window of the test
    # (11) is the evenness window; 550 bins >= 10 * 11 so CV has a value.
    g = json.load(open(os.path.join(golden_dir, "wavelet_minimal.json")))
    cov = np.array(g["coverage"], np.float64)
    off = np.array([0, len(cov)], np.int64)
    r = engine.partition_wavelet(off, cov, is_germline=False, mad_factor=5.0, thr_lower=5.0, thr_upper=80.0,
                                 min_size=10, evenness_window=11)
    assert r["breakpoints"][0].tolist() == g["breakpoints"]
    o = ora.partition_wavelet(off, cov, is_germline=False, mad_factor=5.0, thr_lower=5.0, thr_upper=80.0,
                              min_size=10, evenness_window=11)
    _compare(r, o, "golden")


@pytest.mark.parametrize("germline", [True, False])
def test_chr20_config1(engine, germline):
    s = synth.make_sample(config=1, chromosomes=["chr20"])
    off, cov = _cleaned_coverage(s)
    a = engine.partition_wavelet(off, cov, is_germline=germline)
    b = ora.partition_wavelet(off, cov, is_germline=germline, n_threads=8)
    _compare(a, b, "chr20")
    assert b["cv"] is None  # < 10 windows of 100000: threshold falls back to the MAD


@pytest.mark.parametrize("germline", [True, False])
def test_scaled_genome(engine, germline):
    s = synth.make_sample(config=2, sample=6, scale=0.12, n_events=300)
    off, cov = _cleaned_coverage(s)
    # small windows so that the CV / evenness branches that need >= 10 windows are exercised
    a = engine.partition_wavelet(off, cov, is_germline=germline, evenness_window=12000)
    b = ora.partition_wavelet(off, cov, is_germline=germline, evenness_window=12000, n_threads=8)
    _compare(a, b, "scaled")
    assert b["cv"] is not None and b["evenness"] is not None
    assert sum(len(x) for x in b["breakpoints"]) > 40


def test_full_genome_config2(engine):
    s = synth.make_sample(config=2)
    off, cov = _cleaned_coverage(s)
    a = engine.partition_wavelet(off, cov, is_germline=True)
    b = ora.partition_wavelet(off, cov, is_germline=True, n_threads=8)
    _compare(a, b, "full")


def test_tumour_like_config3(engine):
    s = synth.make_sample(config=3, sample=0, scale=0.3, n_events=150, tumour=True)
    off, cov = _cleaned_coverage(s)
    a = engine.partition_wavelet(off, cov, is_germline=False, evenness_window=30000)
    b = ora.partition_wavelet(off, cov, is_germline=False, evenness_window=30000, n_threads=8)
    _compare(a, b, "tumour")


def test_zero_runs_and_short_chromosomes(engine):
    # runs of exact zeros (unmappable stretches, chrY of a female sample) make the reference peel one
    # bin per level; chromosomes at or below MinSize are skipped
    rng = np.random.default_rng(21)
    lens = [5000, 10, 11, 3000, 1, 0, 2500]
    parts = []
    for i, n in enumerate(lens):
        x = np.round(rng.normal(100, 10, n), 2)
        if i == 0:
            x[1200:2900] = 0.0
            x[2900:3400] = np.round(rng.normal(150, 10, 500), 2)
        if i == 3:
            x[:] = 0.0
        parts.append(x)
    cov = np.concatenate(parts)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for germline in (False, True):
        a = engine.partition_wavelet(off, cov, is_germline=germline, evenness_window=1000)
        b = ora.partition_wavelet(off, cov, is_germline=germline, evenness_window=1000, n_threads=4)
        _compare(a, b, "zeros")
        assert len(b["breakpoints"][1]) == 0 and len(b["breakpoints"][2]) >= 1


def test_shard_mask_matches_full_run(engine):
    s = synth.make_sample(config=2, sample=7, scale=0.05, n_events=200)
    off, cov = _cleaned_coverage(s)
    full = engine.partition_wavelet(off, cov, is_germline=True, evenness_window=5000)
    nc = len(off) - 1
    for rank in range(2):
        mask = np.array([1 if c % 2 == rank else 0 for c in range(nc)], np.uint8)
        part = engine.partition_wavelet(off, cov, is_germline=True, evenness_window=5000, chrom_selected=mask)
        assert part["cv"] == full["cv"] and part["evenness"] == full["evenness"]
        for c in range(nc):
            if mask[c]:
                assert part["breakpoints"][c].tolist() == full["breakpoints"][c].tolist()
            else:
                assert len(part["breakpoints"][c]) == 0


def test_empty_input(engine):
    off = np.array([0, 0], np.int64)
    r = engine.partition_wavelet(off, np.zeros(0), is_germline=True)
    assert len(r["breakpoints"][0]) == 0 and r["cv"] is None and r["evenness"] is None


def test_fused_clean_partition_matches_two_calls(engine):
    # cg_clean_partition_wavelet == cg_clean, .cleaned text round trip on the host, cg_partition_wavelet
    from canvas_b200 import textcodec
    s = synth.make_sample(config=2, sample=8, scale=0.2, n_events=120)
    fused = engine.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc,
                                           is_germline=True, evenness_window=20000)
    o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    assert np.array_equal(fused["kept_index"], o["kept_index"])
    assert np.array_equal(fused["count"].view(np.uint32), o["count"].view(np.uint32))
    assert fused["local_sd"] == o["local_sd"]
    off = synth.chrom_offsets(s.chrom[o["kept_index"]], len(s.names))
    assert np.array_equal(fused["chrom_off"], off)
    cov = ora.f2_roundtrip(o["count"])
    assert np.array_equal(cov, textcodec.f2_roundtrip(o["count"]))
    p = ora.partition_wavelet(off, cov, is_germline=True, evenness_window=20000, n_threads=8)
    _compare(fused, p, "fused")
    stages = engine.last_stage_ms()
    assert all(v > 0 for v in stages.values()), stages
    st = engine.last_partition_stats()
    assert st["bins"] == len(cov) and st["visits"] > 10 * len(cov)


def test_prefetched_inputs_give_the_same_result(engine):
    # cg_prefetch_bins stages a sample's columns ahead of the call; the call that is given the same arrays reads the staged
    # copy.  Results must be those of the plain call, whatever the interleaving: two samples in flight, a prefetch that is
    # never consumed, a call without prefetch in between, Clean alone.
    a = synth.make_sample(config=2, sample=5, scale=0.03, n_events=60)
    b = synth.make_sample(config=2, sample=6, scale=0.02, n_events=60)

    def cols(s):
        return dict(chrom=np.ascontiguousarray(s.chrom, np.uint8), start=np.ascontiguousarray(s.start, np.int32),
                    stop=np.ascontiguousarray(s.stop, np.int32), count=np.ascontiguousarray(s.count, np.float32),
                    gc=np.ascontiguousarray(s.gc, np.uint8))

    def call(s, c):
        return engine.clean_partition_wavelet(c["chrom"], s.is_autosome, s.is_chr_y, c["start"], c["stop"], c["count"], c["gc"],
                                              evenness_window=4000)

    def same(x, y):
        assert np.array_equal(x["kept_index"], y["kept_index"]) and np.array_equal(x["count"].view(np.uint32), y["count"].view(np.uint32))
        assert all(p.tolist() == q.tolist() for p, q in zip(x["breakpoints"], y["breakpoints"])) and x["cv"] == y["cv"]

    ca, cb = cols(a), cols(b)
    ra, rb = call(a, ca), call(b, cb)
    pre = lambda c: engine.prefetch_bins(c["chrom"], c["start"], c["stop"], c["count"], c["gc"])  # noqa: E731
    pre(ca)
    for _ in range(3):      # steady state of a cohort: next sample staged before the current call
        pre(cb)
        same(call(a, ca), ra)
        pre(ca)
        same(call(b, cb), rb)
    same(call(b, cb), rb)   # nothing staged for b now: plain copies
    same(call(a, ca), ra)   # consumes the copy staged in the last round
    pre(ca)
    pre(ca)                 # both slots hold a; b is called without prefetch in between
    same(call(b, cb), rb)
    same(call(a, ca), ra)
    same(call(a, ca), ra)
    same(call(a, ca), ra)
    pre(cb)
    c1 = engine.clean(cb["chrom"], b.is_autosome, b.is_chr_y, cb["start"], cb["stop"], cb["count"], cb["gc"])
    assert np.array_equal(c1["kept_index"], rb["kept_index"]) and np.array_equal(c1["count"].view(np.uint32), rb["count"].view(np.uint32))
    # a staged copy that two calls in a row did not ask for is dropped: a host buffer refilled later is read afresh
    pre(ca)
    same(call(b, cb), rb)
    same(call(b, cb), rb)
    ca["count"][:] = np.roll(ca["count"], 12345)
    fresh = {k: v.copy() for k, v in ca.items()}
    same(call(a, ca), call(a, fresh))
    with pytest.raises(ValueError):
        engine.prefetch_bins(a.chrom.astype(np.int64), ca["start"], ca["stop"], ca["count"], ca["gc"])


def test_graph_replay_across_samples_of_one_layout(engine):
    # The pipelines' CUDA graph is keyed on the INPUT chromosome lengths: every sample binned on one layout replays the graph
    # captured for the first two, with its own cleaned lengths read from device memory.  Three samples of one layout, visited
    # repeatedly in mixed order (with and without prefetch), must each keep reproducing the oracle's result.
    samples = [synth.make_sample(config=2, sample=10 + k, scale=0.04, n_events=50 + 10 * k) for k in range(3)]
    assert all(len(s) == len(samples[0]) for s in samples)
    want = []
    for s in samples:
        o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
        off = synth.chrom_offsets(s.chrom[o["kept_index"]], len(s.names))
        p = ora.partition_wavelet(off, ora.f2_roundtrip(o["count"]), is_germline=True, evenness_window=5000, n_threads=4)
        want.append((o, p))
    assert len({len(o["kept_index"]) for o, _ in want}) == 3  # the cleaned lengths differ: the replayed graph sees new offsets
    cols = [dict(chrom=np.ascontiguousarray(s.chrom, np.uint8), start=np.ascontiguousarray(s.start, np.int32),
                 stop=np.ascontiguousarray(s.stop, np.int32), count=np.ascontiguousarray(s.count, np.float32),
                 gc=np.ascontiguousarray(s.gc, np.uint8)) for s in samples]
    order = [0, 1, 2, 2, 0, 1, 0, 2, 1, 1, 0]
    for step, k in enumerate(order):
        s, c = samples[k], cols[k]
        if step % 3 == 1:
            engine.prefetch_bins(c["chrom"], c["start"], c["stop"], c["count"], c["gc"])
        r = engine.clean_partition_wavelet(c["chrom"], s.is_autosome, s.is_chr_y, c["start"], c["stop"], c["count"], c["gc"],
                                           is_germline=True, evenness_window=5000)
        o, p = want[k]
        assert np.array_equal(r["kept_index"], o["kept_index"]), (step, k)
        assert np.array_equal(r["count"].view(np.uint32), o["count"].view(np.uint32)), (step, k)
        _compare(r, p, f"step {step} sample {k}")
