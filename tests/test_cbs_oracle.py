"""CBS oracle: known answers for the published third-party pieces it restates, and behaviour on
planted change points.  The reference has no CBS test: parity is unpinned beyond these."""
import numpy as np

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
