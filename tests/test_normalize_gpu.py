"""cg_normalize_reference / cg_normalize_ratio against the oracle: bit-identical medians, weights, reference counts,
kept bins, ratios and counts; the CanvasNormalize module end to end on files."""
import gzip
import os

import numpy as np
import pytest

from canvas_b200 import modules, native, textcodec
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return native.Engine(0)


def _controls(rng, s, n, integer=True):
    depth = rng.uniform(40, 160, (s, 1))
    shape = rng.gamma(8, 1 / 8, (1, n))
    c = rng.poisson(depth * shape).astype(np.float64)
    return c if integer else np.round(c + rng.normal(0, 0.3, c.shape), 2)


@pytest.mark.parametrize("s,n,integer,masked", [(2, 1000, True, False), (5, 70001, True, True), (3, 4096, False, True),
                                                (8, 300000, True, False), (1, 17, True, False), (4, 2, False, False)])
def test_reference_matches_oracle(eng, s, n, integer, masked):
    rng = np.random.default_rng(s * 1000 + n)
    c = _controls(rng, s, n, integer)
    on = (rng.random(n) < 0.3).astype(np.uint8) if masked else None
    want = po.normalize_reference(c, on)
    got = eng.normalize_reference(c, on)
    for k in ("median", "weight", "reference"):
        assert np.array_equal(want[k], got[k], equal_nan=True), k


def test_reference_degenerate_cases(eng):
    c = np.array([[0.0, 0.0, 5.0], [2.0, 4.0, 8.0]])
    for on in (None, [0, 0, 0], [0, 1, 0]):
        want, got = po.normalize_reference(c, on), eng.normalize_reference(c, on)
        for k in ("median", "weight", "reference"):
            assert np.array_equal(want[k], got[k], equal_nan=True), (on, k)
    z = eng.normalize_reference(np.zeros((3, 0)))
    assert len(z["reference"]) == 0 and z["median"].tolist() == [0, 0, 0]
    with pytest.raises(ValueError):
        eng.normalize_reference(c, on_target=[1, 1])


@pytest.mark.parametrize("n,mode,masked,ploidy", [(1000, "lsnorm", False, False), (250001, "lsnorm", True, True),
                                                  (65536, "raw", False, True), (3, "lsnorm", False, False),
                                                  (3100000, "lsnorm", False, False)])
def test_ratio_matches_oracle(eng, n, mode, masked, ploidy):
    rng = np.random.default_rng(n)
    sample = rng.poisson(rng.gamma(8, 12, n)).astype(np.float32)
    ref = np.round(rng.gamma(8, 10, n), 3).astype(np.float32)
    ref[rng.random(n) < 0.02] = np.float32(0.4)
    ref[rng.random(n) < 0.001] = 0
    on = (rng.random(n) < 0.5).astype(np.uint8) if masked else None
    pl = rng.choice([0, 1, 2, 3], n).astype(np.int32) if ploidy else None
    kw = dict(mode=mode, min_ref=20.0, max_ref=150.0) if mode == "raw" else dict(mode=mode)
    want = po.normalize_ratio(sample, ref, on, ploidy=pl, **kw)
    got = eng.normalize_ratio(sample, ref, on, ploidy=pl, **kw)
    assert want["library_size_factor"] == got["library_size_factor"]
    assert 0 < len(want["kept_index"]) < n or n == 3
    for k in ("kept_index", "ratio", "count"):
        assert np.array_equal(want[k], got[k], equal_nan=True), k


def test_ratio_special_values(eng):
    sample = np.array([1, np.nan, 3, 0, 5, np.inf, 7], np.float32)
    ref = np.array([np.nan, 2, np.inf, 0.99, 1, 1, -3], np.float32)
    for mode in ("lsnorm", "raw"):
        want, got = po.normalize_ratio(sample, ref, mode=mode), eng.normalize_ratio(sample, ref, mode=mode)
        for k in ("kept_index", "ratio", "count"):
            assert np.array_equal(want[k], got[k], equal_nan=True), (mode, k)
    assert len(eng.normalize_ratio(np.zeros(0), np.zeros(0))["ratio"]) == 0
    # the shorter list ends the enumeration (eSampleBins.MoveNext() && eReferenceBins.MoveNext())
    assert len(eng.normalize_ratio(np.ones(10), np.ones(4), mode="raw")["ratio"]) == 4
    # ... but LSNorm takes each median over its WHOLE file first (LSNormRatioCalculator.cs:28-36): lists of different
    # length are refused rather than normalised with medians of truncated lists
    with pytest.raises(native.CanvasGpuError) as e:
        eng.normalize_ratio(np.ones(10), np.ones(4), mode="lsnorm")
    assert e.value.code == native.CG_ERR_UNSUPPORTED


@pytest.mark.parametrize("s,n,masked", [(2, 1000, False), (6, 200001, True), (3, 3100000, False), (4, 5, False)])
def test_best_lr2_matches_oracle(eng, s, n, masked):
    rng = np.random.default_rng(s * 7 + n)
    c = _controls(rng, s, n)
    sample = np.round(c[rng.integers(0, s)] * rng.uniform(0.5, 2.0) + rng.normal(0, 3, n))
    sample[rng.random(n) < 0.01] = 0
    c[:, rng.random(n) < 0.005] = 0
    on = (rng.random(n) < 0.4).astype(np.uint8) if masked else None
    want, got = po.normalize_best_lr2(sample, c, on), eng.normalize_best_lr2(sample, c, on)
    assert want["best"] == got["best"] and np.array_equal(want["ignored"], got["ignored"])
    # chunk-wise sums on the device, one left-to-right sum of up to 3 M terms in the reference (and libm against CUDA log): 1e-9 relative is far inside 1e-5
    assert np.allclose(want["mean_sq_log_ratio"], got["mean_sq_log_ratio"], rtol=1e-9, atol=0)


def test_best_lr2_degenerate(eng):
    sample = np.array([10.0, 20.0, 30.0, 40.0])
    for controls in ([[0.0, 0.0, 0.0, 9.0], [20.0, 20.0, 20.0, 20.0]], [[np.nan] * 4, [1.0, 2.0, 3.0, 4.0]], [[5.0, 10.0, 15.0, 20.0]]):
        want, got = po.normalize_best_lr2(sample, controls), eng.normalize_best_lr2(sample, controls)
        assert want["best"] == got["best"] and np.array_equal(want["ignored"], got["ignored"])
        assert np.allclose(want["mean_sq_log_ratio"], got["mean_sq_log_ratio"], rtol=1e-9, atol=0, equal_nan=True)
    with pytest.raises(ValueError):
        eng.normalize_best_lr2(sample, np.ones((2, 3)))


def _pca_case(rng, n, k):
    mu = rng.gamma(8, 12, n).astype(np.float32)
    q, _ = np.linalg.qr(rng.normal(size=(n, k)))          # orthonormal columns, then arbitrary lengths
    axes = (q.T * rng.uniform(0.5, 30, (k, 1))).copy()
    sample = np.round(mu * rng.uniform(0.7, 1.3) + (axes * rng.normal(0, 40, (k, 1))).sum(0) + rng.normal(0, 4, n)).astype(np.float32)
    sample[rng.random(n) < 0.01] = 0
    return sample, mu, axes


@pytest.mark.parametrize("n,k,masked", [(1000, 1, False), (70001, 3, True), (500000, 5, False), (9, 2, False)])
def test_pca_reference_matches_oracle(eng, n, k, masked):
    rng = np.random.default_rng(n + k)
    sample, mu, axes = _pca_case(rng, n, k)
    on = (rng.random(n) < 0.5).astype(np.uint8) if masked else None
    want = po.normalize_pca_reference(sample, mu, axes, on, 5.0, 400.0)
    got = eng.normalize_pca_reference(sample, mu, axes, on, 5.0, 400.0)
    # dot products are chunk-wise device sums against the reference's left-to-right sums: tolerance 1e-6 relative on the
    # float outputs (one float ulp is 6e-8), inside the 1e-5 the path allows
    assert abs(want["median_ratio"] - got["median_ratio"]) <= 1e-6 * abs(want["median_ratio"])
    assert np.allclose(want["reference"], got["reference"], rtol=1e-6, atol=0)
    assert np.mean(want["reference"] == got["reference"]) > 0.99


def test_pca_reference_refuses_skewed_axes(eng):
    rng = np.random.default_rng(3)
    sample, mu, axes = _pca_case(rng, 5000, 2)
    axes[1] += 0.01 * axes[0]
    assert po.normalize_pca_reference(sample, mu, axes) is None
    with pytest.raises(native.CanvasGpuError):
        eng.normalize_pca_reference(sample, mu, axes)
    with pytest.raises(ValueError):
        eng.normalize_pca_reference(sample[:10], mu, axes)


def _write_binned(path, chrom, start, count, fmt="{:.0f}"):
    lines = [f"{c}\t{s}\t{s + 1000}\t{fmt.format(v)}\t{40 + i % 20}" for i, (c, s, v) in enumerate(zip(chrom, start, count))]
    with open(path, "wb") as f:
        f.write(gzip.compress(("\n".join(lines) + "\n").encode()))


def test_canvas_normalize_module(tmp_path):
    rng = np.random.default_rng(4)
    n = 5000
    chrom = ["chr1"] * 3000 + ["chrX"] * 2000
    start = list(range(0, 3000000, 1000)) + list(range(0, 2000000, 1000))
    controls = _controls(rng, 3, n)
    controls[:, 100:110] = 0  # reference below 1: dropped
    tumor = rng.poisson(90, n).astype(np.float64)
    tumor[3000:3500] *= 2
    paths = []
    for k in range(3):
        p = str(tmp_path / f"normal{k}.binned")
        _write_binned(p, chrom, start, controls[k])
        paths.append(p)
    t = str(tmp_path / "tumor.binned")
    _write_binned(t, chrom, start, tumor)
    out, w = str(tmp_path / "out.binned"), str(tmp_path / "ref.binned")
    argv = ["-t", t, "-o", out, "-w", w]
    for p in paths:
        argv += ["-n", p]
    assert modules.canvas_normalize_main(argv) == 0
    ref = po.normalize_reference(controls)["reference"]
    ref_lines = gzip.open(w, "rt").read().splitlines()
    assert len(ref_lines) == n
    ref_text = [ln.split("\t")[3] for ln in ref_lines]
    assert all(abs(float(a) - b) <= 1e-14 * max(1.0, abs(b)) for a, b in zip(ref_text, ref.tolist()))
    ref_f = np.array([float(a) for a in ref_text]).astype(np.float32)
    want = po.normalize_ratio(tumor.astype(np.float32), ref_f)
    got = [ln.split("\t") for ln in gzip.open(out, "rt").read().splitlines()]
    assert len(got) == len(want["kept_index"]) == n - 10
    assert [int(g[1]) for g in got] == [start[i] for i in want["kept_index"]]
    assert [g[3] for g in got] == textcodec.f2_text(want["count"])
    cnd = open(out + ".cnd").read().splitlines()
    assert cnd[0] == "Fragment Count,Reference Count,Chromosome,Start,End,Unsmoothed Log Ratio" and len(cnd) == len(got) + 1
    # a single control is copied; missing files and missing options are reported as the reference does
    out2, w2 = str(tmp_path / "out2.binned"), str(tmp_path / "ref2.binned")
    assert modules.canvas_normalize_main(["-t", t, "-n", paths[0], "-o", out2, "-w", w2]) == 0
    assert open(w2, "rb").read() == open(paths[0], "rb").read()
    # BestLR2: the reference is a copy of the control closest to the tumour
    out3, w3 = str(tmp_path / "out3.binned"), str(tmp_path / "ref3.binned")
    assert modules.canvas_normalize_main(argv[:2] + ["-o", out3, "-w", w3, "-m", "BestLR2"] + argv[6:]) == 0
    best = po.normalize_best_lr2(tumor, controls)["best"]
    assert open(w3, "rb").read() == open(paths[best], "rb").read()
    # PCA: model file = chr, start, stop, mean, axes...; raw ratios against the projected reference
    rng2 = np.random.default_rng(9)
    mu = controls.mean(0).astype(np.float32)
    q, _ = np.linalg.qr(rng2.normal(size=(n, 2)))
    axes = (q.T * 25.0).copy()
    model = str(tmp_path / "model.txt.gz")
    with open(model, "wb") as f:
        f.write(gzip.compress(("\n".join(f"{c}\t{s_}\t{s_ + 1000}\t{m!r}\t{a!r}\t{b!r}" for c, s_, m, a, b in
                                          zip(chrom, start, mu.tolist(), axes[0].tolist(), axes[1].tolist())) + "\n").encode()))
    out4, w4 = str(tmp_path / "out4.binned"), str(tmp_path / "ref4.binned")
    assert modules.canvas_normalize_main(["-t", t, "-n", model, "-o", out4, "-w", w4, "-m", "PCA", "-r", "1", "-r", "100000"]) == 0
    want_ref = po.normalize_pca_reference(tumor.astype(np.float32), mu, axes, None, 1.0, 100000.0)
    got_ref = [ln.split("\t")[3] for ln in gzip.open(w4, "rt").read().splitlines()]
    assert len(got_ref) == n
    assert np.mean(np.array(got_ref) == np.array(textcodec.f2_text(want_ref["reference"]))) > 0.99
    assert len(gzip.open(out4, "rt").read().splitlines()) > 0.9 * n
    assert modules.canvas_normalize_main(["-t", t, "-n", model, "-n", model, "-o", out4, "-w", w4, "-m", "PCA"]) == 1
    assert modules.canvas_normalize_main(["-t", t, "-n", str(tmp_path / "nope"), "-o", out2, "-w", w2]) == 1
    assert modules.canvas_normalize_main(["-n", paths[0], "-o", out2]) == 1
