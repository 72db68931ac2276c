"""CanvasBin counting kernels through the C-ABI against the oracle: integers, bit-exact."""
import numpy as np
import pytest

from canvas_b200 import binning, native
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return native.Engine(0)


def _chromosome(rng, n, n_lead=1000, p_possible=0.8):
    bases = rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02])
    bases[:n_lead] = ord("n")
    # long unmappable stretches as well as speckle
    possible = rng.random(n) < p_possible
    for s in rng.integers(0, n, 20):
        possible[s:s + int(rng.integers(1, max(2, n // 100)))] = False
    hits = np.minimum(rng.poisson(3.0, n), 255).astype(np.uint8)
    hits[rng.integers(0, n, 50)] = 255
    return bases.tobytes(), possible, hits


@pytest.mark.parametrize("n,bin_size,seed", [(1, 1, 0), (70_000, 100, 1), (1_000_003, 777, 2), (5_000_000, 1000, 3), (300_000, 400_000, 4)])
def test_bin_hits_matches_oracle(eng, n, bin_size, seed):
    rng = np.random.default_rng(seed)
    bases, possible, hits = _chromosome(rng, n, n_lead=min(1000, n // 2))
    want = pyoracle.bin_hits(hits, possible, bases, bin_size)
    got = eng.bin_hits(hits, possible, bases, bin_size)
    for k in ("start", "stop", "count", "gc"):
        assert np.array_equal(got[k], want[k]), k
    if n > 1000 and bin_size < n // 10:
        assert len(got["start"]) > 0


def test_bin_hits_weighted_matches_oracle(eng):
    rng = np.random.default_rng(11)
    n = 2_000_000
    bases, possible, hits = _chromosome(rng, n)
    read_gc = rng.integers(0, 101, n).astype(np.uint8)
    ratio = (0.5 + rng.random(101)).astype(np.float32)
    want = pyoracle.bin_hits(hits, possible, bases, 500, mode=1, read_gc=read_gc, obs_vs_exp=ratio)
    got = eng.bin_hits(hits, possible, bases, 500, mode=1, read_gc=read_gc, obs_vs_exp_gc=ratio)
    for k in ("start", "stop", "count", "gc"):
        assert np.array_equal(got[k], want[k]), k


def test_bin_hits_all_n_and_empty(eng):
    got = eng.bin_hits(np.zeros(0, np.uint8), np.zeros(0, bool), b"", 10)
    assert len(got["start"]) == 0
    got = eng.bin_hits(np.ones(100, np.uint8), np.ones(100, bool), b"n" * 100, 10)
    assert len(got["start"]) == 0


def _alignments(rng, n_pairs, chr_len):
    left = np.sort(rng.integers(0, chr_len - 1000, n_pairs))
    tlen = rng.integers(0, 600, n_pairs)  # 0 = unavailable
    same = rng.random(n_pairs) < 0.02
    right = np.where(same, left, left + np.maximum(tlen - 100, 1))
    names = [f"r{i}" for i in range(n_pairs)]
    rec = []
    for i in range(n_pairs):
        f1 = 0x1 | (0x2 if rng.random() > 0.05 else 0) | (0x400 if rng.random() < 0.05 else 0) | (0x200 if rng.random() < 0.02 else 0)
        f2 = 0x1 | 0x2 | (0x400 if rng.random() < 0.05 else 0) | (0x100 if rng.random() < 0.01 else 0)
        q1, q2 = int(rng.choice([0, 2, 10, 60, 255], p=[.03, .03, .2, .7, .04])), int(rng.choice([0, 2, 10, 60, 255], p=[.03, .03, .2, .7, .04]))
        mref = 0 if rng.random() > 0.01 else 1
        rec.append((int(left[i]), 0, f1, int(right[i]), mref, int(tlen[i]), q1, i))
        rec.append((int(right[i]), 1, f2, int(left[i]), mref, -int(tlen[i]), q2, i))
    rec.sort(key=lambda r: (r[0], r[1]))
    a = np.array([(r[0], r[2], r[3], r[4], r[5], r[6], r[7]) for r in rec], np.int64).reshape(-1, 7)
    return dict(pos=a[:, 0], flags=binning.flags_from_sam(a[:, 1]), mate_pos=a[:, 2], ref_id=np.zeros(len(a), np.int64),
                mate_ref_id=a[:, 3], frag_len=a[:, 4], mapq=a[:, 5], name_id=a[:, 6], names=[names[j] for j in a[:, 6]])


@pytest.mark.parametrize("n_pairs,seed", [(0, 0), (2000, 1), (200_000, 2)])
def test_bin_fragments_matches_oracle(eng, n_pairs, seed):
    rng = np.random.default_rng(seed)
    chr_len = 2_000_000
    edges = np.sort(rng.choice(np.arange(1, chr_len), 4000, replace=False))
    bs, be = edges[:-1].copy(), edges[1:].copy()
    keep = rng.random(len(bs)) > 0.1  # gaps between bins
    bs, be = bs[keep], be[keep]
    a = _alignments(rng, n_pairs, chr_len)
    want = pyoracle.bin_alignments(a["flags"], a["pos"], a["mate_pos"], a["ref_id"], a["mate_ref_id"], a["frag_len"], a["mapq"],
                                   a["name_id"], 3, bs, be)
    got = binning.bin_paired_alignments(eng, a["flags"], a["pos"], a["mate_pos"], a["ref_id"], a["mate_ref_id"], a["frag_len"],
                                        a["mapq"], a["names"], 3, bs, be)
    assert np.array_equal(got["count"], want["count"])
    assert got["usable"] == want["usable"]
    if n_pairs >= 2000:
        assert want["usable"] > n_pairs // 2


def test_bin_fragments_reference_cases(eng):
    # CanvasTest/TestCanvasBin.cs:14-78 through the host pairing + the kernel
    for pos1, pos2 in [(100, 120), (100, 100)]:
        for mq1, mq2, expect in [(10, 10, 1), (10, 2, 0), (2, 10, 0), (2, 2, 0)]:
            flags = binning.flags_from_sam([0x3, 0x3])
            got = binning.bin_paired_alignments(eng, flags, [pos1, pos2], [pos2, pos1], [0, 0], [0, 0], [100, -100], [mq1, mq2],
                                                ["ReadName", "ReadName"], 3, [100], [200])
            assert got["count"].tolist() == [expect]


@pytest.mark.parametrize("n,n_filter", [(1000, 0), (64, 3), (100003, 40), (5000000, 2000), (7, 1)])
def test_bin_screen_matches_oracle(eng, n, n_filter):
    rng = np.random.default_rng(n + n_filter)
    hits = np.where(rng.random(n) < 0.3, rng.integers(1, 256, n), 0).astype(np.uint8)
    possible = rng.random(n) < 0.8
    fs = rng.integers(0, n, n_filter).astype(np.int32)
    fe = np.minimum(n, fs + rng.integers(0, max(2, min(n, 700)), n_filter)).astype(np.int32)
    if n_filter > 2:
        fe[1] = fs[1]                      # empty interval
        fs[2], fe[2] = 0, min(n, 200)      # overlaps whatever else starts there
    want = pyoracle.bin_screen(hits, possible, fs, fe)
    got = eng.bin_screen(hits, possible, fs, fe)
    assert np.array_equal(want["possible"], got["possible"])
    assert np.array_equal(want["hits"], got["hits"])
    assert (want["observed"], want["n_possible"]) == (got["observed"], got["n_possible"])


def test_bin_screen_rejects_intervals_past_the_end(eng):
    with pytest.raises(native.CanvasGpuError):
        eng.bin_screen(np.zeros(100, np.uint8), np.ones(100, bool), [90], [101])
    r = eng.bin_screen(np.zeros(0, np.uint8), np.zeros(0, bool))
    assert (r["observed"], r["n_possible"]) == (0, 0)


@pytest.mark.parametrize("n,mean", [(5000, 150), (100003, 300), (4097, 1), (2000000, 420), (700, 400)])
def test_read_gc_matches_oracle(eng, n, mean):
    rng = np.random.default_rng(n + mean)
    bases = rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]).tobytes()
    frag = np.where(rng.random(n) < 0.3, rng.integers(-20, 5 * mean + 2, n), 0).astype(np.int16)
    hits = np.where(rng.random(n) < 0.3, rng.integers(1, 256, n), 0).astype(np.uint8)
    if n > 200000:  # the oracle counts every fragment base by base: keep it to seconds
        frag = np.minimum(frag, 60).astype(np.int16)
        mean = 20
    want = pyoracle.bin_read_gc(bases, frag, mean, hits)
    got = eng.bin_read_gc(bases, frag, mean, hits)
    assert np.array_equal(want["read_gc"], got["read_gc"])
    assert np.array_equal(want["expected"], got["expected"]) and np.array_equal(want["observed"], got["observed"])
    # the histograms accumulate over chromosomes
    again = eng.bin_read_gc(bases, frag, mean, hits, got["expected"], got["observed"])
    assert np.array_equal(again["expected"], 2 * want["expected"]) and np.array_equal(again["observed"], 2 * want["observed"])
    s, c = eng.bin_fragment_stats(frag)
    assert (s, c) == (int(frag[frag > 0].astype(np.int64).sum()), int((frag > 0).sum()))


def test_read_gc_rejects_bad_mean(eng):
    with pytest.raises(native.CanvasGpuError):
        eng.bin_read_gc(b"ACGT", np.zeros(4, np.int16), 0, np.zeros(4, np.uint8))
    assert eng.bin_fragment_stats(np.zeros(0, np.int16)) == (0, 0)
