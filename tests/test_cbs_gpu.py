"""cg_partition_cbs against the oracle: identical segments, bit-identical means, and the same number of
tests / permutations / random draws (i.e. the same walk through every sequential stopping rule)."""
import numpy as np
import pytest

from canvas_b200 import native
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return native.Engine(0)


def _chrom(rng, n, events=True):
    x = np.full(n, 100.0)
    if events:
        for _ in range(max(1, n // 3000)):
            s = int(rng.integers(0, max(1, n - 50)))
            ln = int(rng.integers(3, 800)) if rng.random() < 0.6 else int(rng.integers(2, 12))
            x[s:s + ln] *= rng.choice([0.5, 1.5, 2.0, 0.0, 1.1, 0.9, 1.25])
    return np.round(x + rng.normal(0, 8, n), 2)


def _compare(eng, off, cov, **kw):
    want = po.partition_cbs(off, cov, **kw)
    got = eng.partition_cbs(off, cov, **kw)
    for c, (w, g) in enumerate(zip(want["segments"], got["segments"])):
        assert np.array_equal(w["len"], g["len"]), c
        assert np.array_equal(w["mean"], g["mean"]), c
    for k in ("tests", "perms", "perm_steps", "edge_steps"):
        assert want[k] == got[k], k
    return want


def test_boundary_table_matches_oracle(eng):
    assert np.array_equal(eng.cbs_boundary(10000, 0.01, 0.05), po.cbs_boundary(10000, 0.01, 0.05))
    assert np.array_equal(eng.cbs_boundary(1000, 0.05, 0.05), po.cbs_boundary(1000, 0.05, 0.05))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cbs_matches_oracle(eng, seed):
    rng = np.random.default_rng(seed)
    lens = [8000, 6000, 5000, 2500, 300, 150, 20, 3, 0, 5]
    cov = np.concatenate([_chrom(rng, n, n > 200) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    want = _compare(eng, off, cov)
    assert want["edge_steps"] > 0 and want["perms"] > 1000


def test_cbs_many_chromosomes_and_other_seed(eng):
    rng = np.random.default_rng(11)
    lens = [int(v) for v in rng.integers(50, 4000, 40)]  # more chromosomes than clusters: the work queue
    cov = np.concatenate([_chrom(rng, n) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    _compare(eng, off, cov, seed=12345)


def test_cbs_other_parameters(eng):
    rng = np.random.default_rng(4)
    lens = [5000, 900]
    cov = np.concatenate([_chrom(rng, n) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    _compare(eng, off, cov, alpha=0.05, n_perm=1000, k_max=20, min_width=3, n_min=150)


def test_cbs_sdundo_matches_oracle(eng):
    # -s SDUndo: neighbours whose medians differ by less than 3 trimmed SDs are merged again
    rng = np.random.default_rng(9)
    lens = [6000, 3000]
    cov = np.concatenate([_chrom(rng, n) for n in lens])
    cov[3000:3300] += 6  # a weak event that CBS finds and the undo step removes
    off = np.concatenate([[0], np.cumsum(lens)])
    plain = po.partition_cbs(off, cov)
    want = _compare(eng, off, cov, undo=2)
    assert sum(len(s["len"]) for s in want["segments"]) < sum(len(s["len"]) for s in plain["segments"])


def test_cbs_prune_matches_oracle(eng):
    # -s Prune: the subset search over the found change points (ChangePoint.cs:205-271), several chromosomes
    rng = np.random.default_rng(12)
    parts = [_chrom(rng, n) for n in (6000, 2500, 9000, 300)]
    weak = parts[0]
    weak[3000:3300] += 6
    off = np.concatenate([[0], np.cumsum([len(p) for p in parts])])
    cov = np.concatenate(parts)
    plain = eng.partition_cbs(off, cov)
    want = _compare(eng, off, cov, undo=1)
    assert sum(len(s["len"]) for s in want["segments"]) < sum(len(s["len"]) for s in plain["segments"])
    _compare(eng, off, cov, undo=1, undo_prune=0.5)


def test_cbs_degenerate_inputs(eng):
    cov = np.concatenate([[1.0, 2.0, 3.0], np.full(10, 7.0), np.arange(30.0)])
    want = _compare(eng, [0, 3, 3, 13, 43], cov)
    assert [s["len"].tolist() for s in want["segments"][:3]] == [[3], [], [10]]
    got = eng.partition_cbs([0], np.zeros(0))
    assert got["segments"] == []


def test_cbs_rejects_what_it_does_not_implement(eng):
    cov = np.arange(100.0)
    with pytest.raises(native.CanvasGpuError):
        eng.partition_cbs([0, 100], cov, undo=3)
    with pytest.raises(native.CanvasGpuError):
        eng.partition_cbs([0, 100], cov, hybrid=False)
    bad = cov.copy()
    bad[5] = np.nan
    with pytest.raises(native.CanvasGpuError):
        eng.partition_cbs([0, 100], bad)


def test_cbs_shards_union_equals_whole(eng):
    # every chromosome keeps its own random stream: two disjoint shard calls give the whole-genome answer
    rng = np.random.default_rng(21)
    lens = [3000, 50, 2000, 700, 1500]
    cov = np.concatenate([_chrom(rng, n) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    whole = eng.partition_cbs(off, cov)
    a = eng.partition_cbs(off, cov, chrom_selected=[1, 0, 0, 1, 0])
    b = eng.partition_cbs(off, cov, chrom_selected=[0, 1, 1, 0, 1])
    for c in range(len(lens)):
        pick = a if c in (0, 3) else b
        other = b if c in (0, 3) else a
        assert np.array_equal(pick["segments"][c]["len"], whole["segments"][c]["len"])
        assert np.array_equal(pick["segments"][c]["mean"], whole["segments"][c]["mean"])
        assert len(other["segments"][c]["len"]) == 0
    assert a["perms"] + b["perms"] == whole["perms"]
