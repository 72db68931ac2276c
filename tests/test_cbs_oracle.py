"""CBS oracle: known answers for the published third-party pieces it restates, and behaviour on
planted change points.  The reference has no CBS test: parity is unpinned beyond these.
The prune undo step is checked against a written-out search and against the library's host code."""
import numpy as np
import pytest

from oracle import pyoracle as po


def test_mt19937_known_answers():
    # MT19937 reference outputs for init_genrand(5489): first five draws and the 10000th
    out = po.mt19937(5489, 10000)
    assert out[:5].tolist() == [3499211612, 581869302, 3890346734, 3586334585, 545404204]
    assert int(out[-1]) == 4123659995


def test_boundary_shape_and_first_row():
    sb = po.cbs_boundary(10000, 0.01, 0.05)
    assert len(sb) == 101 * 102 // 2
    assert sb[0] == 9500  # nPerm - (uint)(nPerm * eta), GetBoundary.cs:31
    at = 0
    for j in range(1, 102):
        row = sb[at:at + j].astype(np.int64)
        assert np.all(np.diff(row) > 0) and row[-1] <= 10000 and row[0] >= 1
        at += j


def test_tailp_decreases_with_statistic():
    p = [po.cbs_tailp(b, 26 / 5000, 5000) for b in (3.5, 4.0, 4.5, 5.0, 6.0)]
    assert all(a > b for a, b in zip(p, p[1:])) and p[-1] < 1e-4


def test_tmaxo_finds_planted_arc():
    rng = np.random.default_rng(5)
    x = rng.normal(0, 1, 3000)
    x[1200:1500] += 3.0
    stat, i, j = po.cbs_tmaxo(x)
    assert (i, j) == (1200, 1500) and stat > 100


def test_segments_of_piecewise_constant_signal():
    rng = np.random.default_rng(7)
    x = np.round(100 + rng.normal(0, 5, 6000), 2)
    x[1000:1400] += 50
    x[4000:4006] -= 80  # a 6-bin event: its edges go through the permutation t-test (m1 < 10)
    r = po.partition_cbs([0, 6000], x)
    assert r["segments"][0]["len"].tolist() == [1000, 400, 2600, 6, 1994]
    assert r["edge_steps"] > 0 and r["tests"] >= 5
    m = r["segments"][0]["mean"]
    assert abs(m[1] - 150) < 1 and abs(m[3] - 20) < 6
    # first / last bin of every segment (no non-finite bins: the identity mapping of CBSRunner.cs:127-138)
    assert r["segments"][0]["first"].tolist() == [0, 1000, 1400, 4000, 4006]
    assert r["segments"][0]["last"].tolist() == [999, 1399, 3999, 4005, 5999]


def test_short_and_constant_chromosomes():
    r = po.partition_cbs([0, 3, 3, 13, 43], np.concatenate([[1.0, 2.0, 3.0], np.full(10, 7.0), np.arange(30.0)]))
    assert [s["len"].tolist() for s in r["segments"][:3]] == [[3], [], [10]]
    assert sum(r["segments"][3]["len"]) == 30


def test_inflation_factor_and_trimmed_variance():
    # DNAcopy's inflfact(0.025) = 1.3178; the trimmed variance of white noise recovers sigma^2
    assert abs(po.cbs_inflation_factor(0.025) - 1.3178) < 1e-4
    x = np.random.default_rng(0).normal(0, 3, 200000)
    assert abs(po.cbs_trimmed_variance(x) - 9.0) < 0.15


def test_sdundo_merges_a_weak_event():
    rng = np.random.default_rng(0)
    x = np.round(100 + rng.normal(0, 5, 6000), 2)
    x[1000:1400] += 50
    x[3000:3300] += 6
    assert len(po.partition_cbs([0, 6000], x)["segments"][0]["len"]) > 3
    assert po.partition_cbs([0, 6000], x, undo=2)["segments"][0]["len"].tolist() == [1000, 400, 4600]


def _prune_python(g, seg_len, cutoff):
    """ChangePointsPrune (ChangePoint.cs:205-271) written out with itertools: the same lexicographic order of
    subsets as Prune.Combination, `<=` keeps the last best, falling through every j prunes everything."""
    import itertools
    import math
    S, K = len(seg_len), len(seg_len) - 1
    ends = np.cumsum(seg_len)
    starts = ends - seg_len
    ssq = 0.0
    for v in g:
        ssq += math.pow(v, 2)
    sx = []
    for a, b in zip(starts, ends):
        s = 0.0
        for v in g[a:b]:
            s += v
        sx.append(s)

    def errssq(loc):
        cuts = [0] + list(loc) + [S]
        total = 0.0
        for a, b in zip(cuts, cuts[1:]):
            s, cnt = 0.0, 0
            for i in range(a, b):
                s += sx[i]
                cnt += int(seg_len[i])
            total += math.pow(s, 2) / cnt
        return total

    wssqk = ssq - errssq(range(1, K + 1))
    prev, kept = list(range(1, K + 1)), []
    for j in range(K - 1, 0, -1):
        best, wj = None, None
        for loc in itertools.combinations(range(1, K + 1), j):
            w = ssq - errssq(loc)
            if wj is None or w <= wj:
                wj, best = w, loc
        ratio = wj / wssqk if wssqk != 0 else (math.inf if wj > 0 else math.nan)
        if ratio > 1 + cutoff:
            kept = prev
            break
        prev = list(best)
    cuts = [0] + [int(ends[l - 1]) for l in kept] + [len(g)]
    return np.diff(cuts).astype(np.int32)


def _prune_case(seed, n_seg, noise=1.0, integer=False):
    rng = np.random.default_rng(seed)
    seg_len = rng.integers(2, 40, n_seg).astype(np.int32)
    levels = rng.choice([100.0, 100.0, 100.5, 103.0, 120.0, 60.0], n_seg)
    g = np.concatenate([np.full(l, v) for l, v in zip(seg_len, levels)]) + rng.normal(0, noise, int(seg_len.sum()))
    g = np.round(g) if integer else np.round(g, 2)
    return g, seg_len


def test_prune_matches_the_written_out_search():
    # integer data makes ties between subsets frequent: the `<=` (last best wins) rule is exercised
    for seed, n_seg, noise, integer in [(0, 2, 1, False), (1, 3, 1, False), (2, 6, 2, False), (3, 9, 4, False),
                                        (4, 8, 0, True), (5, 10, 1, True), (6, 11, 8, False), (7, 5, 0, True)]:
        g, seg_len = _prune_case(seed, n_seg, noise, integer)
        for cutoff in (0.05, 0.5, 1e-9):
            want = _prune_python(g, seg_len, cutoff)
            got = po.cbs_prune(g, seg_len, cutoff)
            assert np.array_equal(want, got), (seed, cutoff)
            assert got.sum() == len(g)


def test_prune_quirks_of_the_reference():
    # two segments: the j loop never runs and the single change point is always dropped (ChangePoint.cs:227, :264)
    g = np.concatenate([np.full(50, 10.0), np.full(50, 90.0)])
    assert po.cbs_prune(g, [50, 50]).tolist() == [100]
    # three clearly different levels survive
    g = np.concatenate([np.full(50, 10.0), np.full(50, 90.0), np.full(50, 40.0)]) + np.random.default_rng(1).normal(0, 1, 150)
    assert po.cbs_prune(g, [50, 50, 50]).tolist() == [50, 50, 50]
    # through the runner: a weak event found by CBS is merged away, the strong one stays
    rng = np.random.default_rng(0)
    x = np.round(100 + rng.normal(0, 5, 6000), 2)
    x[1000:1400] += 50
    x[3000:3300] += 6
    plain = po.partition_cbs([0, 6000], x)["segments"][0]["len"]
    pruned = po.partition_cbs([0, 6000], x, undo=1, undo_prune=0.1)["segments"][0]["len"]
    assert len(plain) > 3 and pruned.tolist() == [1000, 400, 4600]


def test_library_prune_equals_oracle():
    # host code of the C-ABI library (no device): depth-first walk with tabulated group terms = the oracle's search
    from canvas_b200 import native
    total = 0
    for seed in range(12):
        g, seg_len = _prune_case(100 + seed, 2 + seed % 13 + (6 if seed > 8 else 0), noise=[0, 1, 6][seed % 3], integer=seed % 2 == 0)
        for cutoff in (0.05, 0.3):
            got, scored = native.cbs_prune(g, seg_len, cutoff)
            assert np.array_equal(got, po.cbs_prune(g, seg_len, cutoff)), (seed, cutoff)
            total += scored
    assert total > 1000
    with pytest.raises(native.CanvasGpuError):
        native.cbs_prune(np.zeros(10), [10])          # fewer than two segments
    with pytest.raises(native.CanvasGpuError):
        native.cbs_prune(np.zeros(10), [5, 4])        # lengths do not cover the data
    # exponential search: 79 evenly good change points exceed the budget (2^28 inside cg_partition_cbs) -> refused, not approximated
    g = np.tile(np.r_[np.zeros(5), np.ones(5)], 40) + np.random.default_rng(3).normal(0, 0.1, 400)
    with pytest.raises(native.CanvasGpuError) as e:
        native.cbs_prune(g, np.full(80, 5, np.int32), 1e9, max_subsets=1 << 24)
    assert e.value.code == native.CG_ERR_UNSUPPORTED
