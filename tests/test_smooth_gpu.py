"""cg_smooth against the streaming restatement of Utilities.MedianFilter / RepeatedMedianSmoother (bit-exact floats)."""
import gzip

import numpy as np
import pytest

from canvas_b200 import modules
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def test_reference_vector(engine):
    # CanvasTest/TestUtilities.cs:195-206
    got = engine.smooth(np.array([0, 8]), np.array([2, 1, 3, 5, 4, 6, 7, 8], np.float32), 1)
    assert got[0].tolist() == [1.5, 2, 3, 4, 5, 6, 7, 7.5]


@pytest.mark.parametrize("half", [0, 1, 2, 3, 5, 8, 32])
def test_matches_oracle_on_ragged_chromosomes(engine, half):
    rng = np.random.default_rng(half)
    lens = [0, 1, 2, 3, 5, 7, 16, 17, 33, 64, 65, 66, 1000, 40000]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cnt = np.round(rng.gamma(20, 5, int(off[-1])), 2).astype(np.float32)
    cnt[rng.integers(0, len(cnt), 200)] = 0.0  # ties
    got = engine.smooth(off, cnt, half)
    for c, n in enumerate(lens):
        want = ora.repeated_median_filter(cnt[off[c]:off[c + 1]], half)
        assert len(got[c]) == len(want), (c, n, half)
        assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), (c, n, half)


def test_whole_genome_and_module(engine, tmp_path):
    rng = np.random.default_rng(2)
    names = ["chr1", "chr2", "chrX"]
    lens = [30000, 12, 5000]
    inp = tmp_path / "in.cleaned"
    rows = []
    with gzip.open(inp, "wt") as f:
        for c, n in zip(names, lens):
            for i in range(n):
                v = round(float(rng.gamma(20, 5)), 2)
                rows.append((c, i * 1000, i * 1000 + 1000, v))
                f.write(f"{c}\t{i * 1000}\t{i * 1000 + 1000}\t{v:.2f}\t{40 + i % 5}\n")
    out = tmp_path / "out.smoothed"
    assert modules.main(["CanvasSmooth", "-i", str(inp), "-o", str(out), "-w", "7"]) == 0
    lines = gzip.open(out, "rt").read().splitlines()
    k = 0
    for c, n in zip(names, lens):
        x = np.array([r[3] for r in rows if r[0] == c], np.float32)
        want = ora.repeated_median_filter(x, 7)
        # chr2 has 12 bins: half windows 6 and 7 need 13 / 15 values, so it loses bins
        for i, w in enumerate(want.tolist()):
            p = lines[k].split("\t")
            assert (p[0], int(p[1])) == (c, i * 1000)
            assert abs(float(p[3]) - w) <= 0.005 + 1e-6 * abs(w)
            k += 1
    assert k == len(lines)
    assert modules.main(["CanvasSmooth", "-i", str(tmp_path / "missing"), "-o", str(out)]) == 1
