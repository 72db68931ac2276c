"""cg_merge_common_bins against the dictionary restatement of MergeMultiSampleCleanedBedFile."""
import gzip

import numpy as np
import pytest

from canvas_b200 import fileio, native, textcodec
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def _samples(rng, n_samples, n_chrom, n_bins, drop=0.05):
    names = [f"chr{c + 1}" for c in range(n_chrom)]
    base = []
    for c in range(n_chrom):
        starts = np.sort(rng.choice(np.arange(0, n_bins * 3) * 1000, n_bins, replace=False))
        for a in starts.tolist():
            base.append((c, a, a + int(rng.choice([1000, 1000, 1500]))))
    out = []
    for s in range(n_samples):
        keep = rng.random(len(base)) >= drop
        rows = [(c, a, b + (7 if s == n_samples - 1 and i % 50 == 0 else 0)) for i, (c, a, b) in enumerate(base) if keep[i]]
        cnt = np.round(rng.gamma(20, 5, len(rows)), 2).astype(np.float32)
        out.append((names, rows, cnt))
    return out


def _run(engine, data):
    samples = [(np.array([r[0] for r in rows], np.uint8), np.array([r[1] for r in rows], np.int32),
                np.array([r[2] for r in rows], np.int32), cnt) for _, rows, cnt in data]
    got = engine.merge_common_bins(samples)
    want = ora.merge_multi_sample_cleaned([[(names[c], a, b, v) for (c, a, b), v in zip(rows, cnt.tolist())]
                                           for names, rows, cnt in data])
    names, rows0, _ = data[0]
    assert len(got["kept_index"]) == len(want)
    for k, (i, w) in enumerate(zip(got["kept_index"].tolist(), want)):
        assert (names[rows0[i][0]], rows0[i][1]) == (w[0], w[1])
        assert got["stop"][k] == w[2]
        assert [np.float32(x) for x in got["count"][:, k]] == w[3]
    return got


@pytest.mark.parametrize("n_samples,n_chrom,n_bins", [(1, 2, 50), (2, 3, 400), (3, 5, 3000), (8, 2, 700)])
def test_matches_dictionary_restatement(engine, n_samples, n_chrom, n_bins):
    _run(engine, _samples(np.random.default_rng(n_samples * 100 + n_bins), n_samples, n_chrom, n_bins))


def test_disjoint_and_empty(engine):
    a = (np.zeros(3, np.uint8), np.array([0, 10, 20], np.int32), np.array([10, 20, 30], np.int32), np.ones(3, np.float32))
    b = (np.zeros(2, np.uint8), np.array([5, 15], np.int32), np.array([15, 25], np.int32), np.ones(2, np.float32))
    assert len(engine.merge_common_bins([a, b])["kept_index"]) == 0
    e = tuple(np.zeros(0, t) for t in (np.uint8, np.int32, np.int32, np.float32))
    assert len(engine.merge_common_bins([e, a])["kept_index"]) == 0
    assert len(engine.merge_common_bins([a, e])["kept_index"]) == 0


def test_reference_exceptions_and_order(engine):
    bad = (np.zeros(2, np.uint8), np.array([0, 10], np.int32), np.array([10, 10], np.int32), np.ones(2, np.float32))
    with pytest.raises(native.CanvasGpuError) as e:
        engine.merge_common_bins([bad, bad])
    assert "Start must be less than Stop" in str(e.value)
    neg = (np.zeros(2, np.uint8), np.array([-5, 10], np.int32), np.array([10, 20], np.int32), np.ones(2, np.float32))
    with pytest.raises(native.CanvasGpuError) as e:
        engine.merge_common_bins([neg])
    assert "Start must be non-negative" in str(e.value)
    uns = (np.zeros(2, np.uint8), np.array([10, 0], np.int32), np.array([20, 10], np.int32), np.ones(2, np.float32))
    with pytest.raises(native.CanvasGpuError) as e:
        engine.merge_common_bins([uns, uns])
    assert e.value.code == native.CG_ERR_UNSORTED


def test_normalize_canvas_clean_files(engine, tmp_path):
    rng = np.random.default_rng(9)
    data = _samples(rng, 3, 3, 500)
    paths = []
    for k, (names, rows, cnt) in enumerate(data):
        p = tmp_path / f"s{k}.cleaned"
        with gzip.open(p, "wt") as f:
            for (c, a, b), t in zip(rows, textcodec.f2_text(cnt)):
                f.write(f"{names[c]}\t{a}\t{b}\t{t}\t{40 + k}\n")
        paths.append(str(p))
    n = fileio.normalize_canvas_clean(engine, paths)
    want = ora.merge_multi_sample_cleaned([[(names[c], a, b, float(np.float32(float(t)))) for (c, a, b), t in
                                            zip(rows, textcodec.f2_text(cnt))] for names, rows, cnt in data])
    assert n == len(want)
    for k, p in enumerate(paths):
        lines = gzip.open(p, "rt").read().splitlines()
        assert len(lines) == n
        for line, w in zip(lines, want):
            c, a, b, v = line.split("\t")
            assert (c, int(a), int(b)) == (w[0], w[1], w[2])
            assert np.float32(float(v)) == w[3][k]
