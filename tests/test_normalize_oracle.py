"""CanvasNormalize oracle (weighted-average reference, ratio step): hand-computed cases.  The reference has no test
for CanvasNormalize: parity is unpinned beyond the source restatement."""
import numpy as np

from oracle import pyoracle as po


def test_weighted_average_reference_by_hand():
    # medians 2 and 4 -> weights (1/2, 1/4) / (3/4) = (2/3, 1/3)
    counts = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 8.0]])
    r = po.normalize_reference(counts)
    assert r["median"].tolist() == [2.0, 4.0]
    w = np.array([0.5, 0.25]) / 0.75
    assert r["weight"].tolist() == w.tolist()
    assert r["reference"].tolist() == [(0.0 + w[0] * a) + w[1] * b for a, b in zip(counts[0], counts[1])]
    # on-target bins only feed the medians; an even count takes the mean of the middles
    r = po.normalize_reference(counts, on_target=[1, 0, 1])
    assert r["median"].tolist() == [2.0, 5.0]
    # a control whose median is not positive gets weight 0
    r = po.normalize_reference(np.array([[0.0, 0.0, 5.0], [2.0, 4.0, 8.0]]))
    assert r["weight"].tolist() == [0.0, 1.0] and r["reference"].tolist() == [2.0, 4.0, 8.0]


def test_ratio_steps_by_hand():
    sample = np.array([10, 20, 30, 40, 50], np.float32)
    ref = np.array([20, 0.5, 30, 10, 25], np.float32)
    r = po.normalize_ratio(sample, ref, mode="lsnorm")
    # medians 30 and 20 -> factor 2/3; the bin with reference 0.5 < 1 is dropped
    assert r["library_size_factor"] == 20.0 / 30.0
    assert r["kept_index"].tolist() == [0, 2, 3, 4]
    q = (sample / ref)[[0, 2, 3, 4]]
    want = (q.astype(np.float64) * (20.0 / 30.0)).astype(np.float32)
    assert np.array_equal(r["ratio"], want)
    assert np.array_equal(r["count"], (want.astype(np.float64) * 40.0).astype(np.float32))
    # raw mode: range filter on the reference count, no library size factor; ploidy 1 halves the count
    r = po.normalize_ratio(sample, ref, mode="raw", min_ref=10, max_ref=25, ploidy=[2, 2, 2, 1, 2])
    assert r["kept_index"].tolist() == [0, 3, 4] and r["library_size_factor"] == 1.0
    assert r["ratio"].tolist() == [0.5, 4.0, 2.0] and r["count"].tolist() == [20.0, 80.0, 80.0]
    # a sample median of zero leaves the factor at 1
    assert po.normalize_ratio(np.zeros(4, np.float32), np.full(4, 3, np.float32))["library_size_factor"] == 1.0
    assert len(po.normalize_ratio(np.zeros(0, np.float32), np.zeros(0, np.float32))["ratio"]) == 0


def test_best_lr2_by_hand():
    import math
    sample = np.array([10.0, 20.0, 30.0, 40.0])
    controls = np.array([[5.0, 10.0, 15.0, 20.0],      # the sample's shape: every normalised ratio is 1, log ratio 0
                         [20.0, 20.0, 20.0, 20.0],
                         [10.0, 0.0, 30.0, 41.0]])
    r = po.normalize_best_lr2(sample, controls)
    assert r["best"] == 0 and r["mean_sq_log_ratio"][0] == 0.0 and r["ignored"].tolist() == [0, 0, 1]
    t = sample / 25.0
    want1 = sum(math.log(a / 1.0) ** 2 for a in t) / 4
    assert abs(r["mean_sq_log_ratio"][1] - want1) < 1e-15
    # only on-target bins count, for the medians too
    r = po.normalize_best_lr2(sample, controls, on_target=[0, 1, 1, 1])
    assert r["ignored"].tolist() == [0, 0, 1]
    # a control whose median is not positive normalises to zeros: every bin is ignored, its mean is the empty sum 0 and it wins
    r = po.normalize_best_lr2(sample, np.array([[0.0, 0.0, 0.0, 9.0], [20.0, 20.0, 20.0, 20.0]]))
    assert r["best"] == 0 and r["ignored"][0] == 4 and r["mean_sq_log_ratio"][0] == 0.0


def test_pca_reference_by_hand():
    # one axis along (1, 1, 0, 0) / sqrt 2: the projection replaces the first two centred values by their mean
    mu = np.array([100, 100, 100, 100], np.float32)
    sample = np.array([120, 140, 90, 0.5], np.float32)
    r = po.normalize_pca_reference(sample, mu, [[3.0, 3.0, 0.0, 0.0]])
    ref = np.array([130.0, 130.0, 100.0, 100.0])          # mu + projection; the last two are untouched
    ratios = sample / ref.astype(np.float32)
    med = float(np.median(ratios.astype(np.float64)))
    assert abs(r["median_ratio"] - med) < 1e-12
    assert np.allclose(r["reference"], (ref * med).astype(np.float32), rtol=1e-6)
    # axes that are not orthogonal are refused (PCAModel.LoadModel)
    assert po.normalize_pca_reference(sample, mu, [[1.0, 1.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]]) is None


def _helmert_basis(p):
    # CanvasTest/TestUtilities.cs:125-142 (GetHelmertBasis)
    axes = []
    for i in range(1, p + 1):
        v = np.zeros(p)
        size = np.sqrt(i + i * i if i < p else i)
        v[:i] = 1 / size
        if i < p:
            v[i] = -i / size
        axes.append(v)
    return np.array(axes)


def test_pca_projection_pinned_by_the_reference_projection_tests():
    # TestProject2 (TestUtilities.cs:144-169): projecting onto the complete Helmert basis returns the vector.  Through the
    # PCA reference generator: mu = 0, so the reference vector is max(1, max(1, sample)) and every raw ratio that survives
    # the [1, inf) filter is sample / max(1, sample)
    u = np.arange(10, dtype=np.float32)
    r = po.normalize_pca_reference(u, np.zeros(10, np.float32), _helmert_basis(10))
    want = np.maximum(1.0, u)
    ratios = u / want
    assert abs(r["median_ratio"] - np.median(ratios.astype(np.float64))) < 1e-12
    assert np.allclose(r["reference"], want * r["median_ratio"], rtol=1e-6)
    # axes of any length are scaled to unit length first (TestNormalizeBy2Norm, :65-78): the same answer
    r2 = po.normalize_pca_reference(u, np.zeros(10, np.float32), _helmert_basis(10) * np.arange(1, 11)[:, None])
    assert np.allclose(r2["reference"], r["reference"], rtol=1e-6)
    # TestProject (:106-123): one axis (1, .., 1) / sqrt(10) and the vector e0 project to 1/10 everywhere.
    # mu = 1 and sample = (2, 1, .., 1) centre to e0: reference vector 1.1, written as "1.10"
    sample = np.ones(10, np.float32)
    sample[0] = 2
    r = po.normalize_pca_reference(sample, np.ones(10, np.float32), [np.ones(10)])
    ratios = (sample / np.float32(1.1)).astype(np.float64)
    assert abs(r["median_ratio"] - np.median(ratios)) < 1e-12
    assert np.allclose(r["reference"], np.float32(1.1 * r["median_ratio"]), rtol=1e-6)
    # TestAreOrthogonal (:96-104)
    two = np.array([3.0, 4.0], np.float32)
    assert po.normalize_pca_reference(two, two, [[1.0, 1.0], [1.0, -1.0]]) is not None
    assert po.normalize_pca_reference(two, two, [[1.0, 1.0], [1.0, 1.0]]) is None
