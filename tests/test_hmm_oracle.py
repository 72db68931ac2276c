"""The HMM oracle (oracle/hmm.cpp) against closed forms and planted segments.  The reference holds no test for
CanvasPartition's HMM (parity unpinned); these checks pin the restated third-party pieces (MathNet GammaLn /
FactorialLn, the negative binomial density) against independent implementations and the Viterbi pass against a
plain numpy restatement of HMM.cs:62-130 on small inputs."""
import math

import numpy as np

from oracle import pyoracle as ora


def test_gamma_ln_matches_libm():
    for z in [0.05, 0.3, 0.5, 0.75, 1.0, 1.5, 2.0, 3.7, 10.0, 55.5, 100.0, 171.0, 500.0, 1234.5, 1e5]:
        want = math.lgamma(z)
        assert abs(ora.gamma_ln(z) - want) <= 1e-12 * max(1.0, abs(want)), z


def test_negative_binomial_is_a_density_with_the_requested_mean():
    for mean, var in [(100.0, 400.0), (50.0, 55.0), (5.0, 169.0), (200.0, 150.0)]:
        d = ora.negative_binomial(mean, var, 4000)
        assert abs(d.sum() - 1.0) < 1e-9
        assert abs((d * np.arange(len(d))).sum() - mean) < 1e-6 * mean
        # NegativeBinomialWrapper's variance floor and its floor of 2 on the clumping parameter (DistributionUtilities.cs:59-60)
        k = max(2.0, mean * mean / (max(var, 1.2 * mean) - mean))
        v_eff = mean + mean * mean / k
        assert abs((d * (np.arange(len(d)) - mean) ** 2).sum() - v_eff) < 1e-6 * v_eff


def _numpy_viterbi(x, haploid, variance, per_sample):
    """HiddenMarkovModel.BestPathViterbi for one sample, straight from the formulas."""
    thr = haploid * 5
    x = np.where(x > thr, thr, x)
    xi = np.rint(x).astype(int)
    L = xi.max() + 10
    P = np.stack([ora.negative_binomial(max(cn, 0.1) * haploid, variance, L) for cn in range(5)])
    if not per_sample:
        lo, hi = np.maximum(P[0], P[1]), np.maximum(P[3], P[4])
        P = np.stack([lo, lo, P[2], hi, hi])
    with np.errstate(divide="ignore"):
        le = np.log(P[:, xi])  # [5, n]
    T = np.full((5, 5), (1.0 - 0.99) / 4)
    np.fill_diagonal(T, 0.99)
    lt = np.log(T)
    n = len(x)
    s = np.array([math.log(float(np.float32(1.0) / np.float32(5))) + (le[j, 0] + lt[0, j]) - lt[0, j] for j in range(5)])
    back = np.zeros((n, 5), int)
    for t in range(1, n):
        ns = np.empty(5)
        for j in range(5):
            mx, st = -np.finfo(float).max, 0
            for i in range(5):
                v = s[i] + (le[j, t] + lt[i, j])
                if v > mx:
                    mx, st = v, i
            ns[j], back[t, j] = mx, st
        s = ns
    best, mx = -1, -np.finfo(float).max
    for i in range(5):
        if s[i] > mx:
            best, mx = i, s[i]
    path = np.zeros(n, int)
    for t in range(n - 1, 0, -1):
        path[t] = best
        best = back[t, best]
    path[0] = best
    return path


def _planted(rng, n, haploid=50.0):
    cn = np.full(n, 2)
    for _ in range(max(1, n // 400)):
        a = int(rng.integers(0, n - 5))
        cn[a:a + int(rng.integers(3, 120))] = rng.choice([0, 1, 3, 4])
    return np.round(rng.poisson(haploid * np.maximum(cn, 0.03)).astype(np.float64) + rng.uniform(0, 0.99, n), 2), cn


def test_per_sample_viterbi_matches_numpy_restatement():
    rng = np.random.default_rng(11)
    x, _ = _planted(rng, 3000)
    q = np.sort(x.astype(np.float32))
    q1, q2, q3 = ora.quartiles_f32(x.astype(np.float32))
    iqr = np.float32(q3) - np.float32(q1)
    want = _numpy_viterbi(x, float(q2) / 2.0, float(np.float32(iqr * iqr)), True)
    got = ora.partition_hmm(np.array([0, len(x)]), x, per_sample=True)
    assert np.array_equal(got["states"], want)
    assert got["breakpoints"][0].tolist() == [0] + [t for t in range(1, len(x)) if want[t] != want[t - 1]]
    del q


def test_joint_single_sample_viterbi_matches_numpy_restatement():
    rng = np.random.default_rng(12)
    x, _ = _planted(rng, 2500)
    med = max(1.0, float(np.median(x)))
    mu = x.sum() / len(x)
    var = float(((x - mu) ** 2).sum() / (len(x) - 1))
    want = _numpy_viterbi(x, med / 2.0, var, False)
    got = ora.partition_hmm(np.array([0, len(x)]), x, per_sample=False)
    # states 0|1 and 3|4 share their emission: the path uses whichever index wins the ties, as the reference does
    assert np.array_equal(got["states"], want)


def test_planted_events_are_found_and_short_chromosomes_skipped():
    rng = np.random.default_rng(13)
    n = 60000
    cn = np.full(n, 2)
    cn[10000:12000] = 3
    cn[30000:30400] = 1
    cn[45000:45060] = 0
    x = rng.poisson(50 * np.maximum(cn, 0.03)).astype(np.float64)
    off = np.array([0, 8, n])  # first chromosome has 8 bins (<= MinSize): no segmentation
    r = ora.partition_hmm(off, x, per_sample=True, n_threads=2)
    assert r["breakpoints"][0].tolist() == []
    bps = (r["breakpoints"][1] + 8).tolist()
    for edge in (10000, 12000, 30000, 30400, 45000, 45060):
        assert any(abs(b - edge) <= 3 for b in bps), (edge, bps)
    assert len(bps) <= 12


def test_joint_three_samples_runs_and_shares_breakpoints():
    rng = np.random.default_rng(14)
    n = 20000
    cn = np.full(n, 2)
    cn[5000:5600] = 1
    x = np.stack([rng.poisson(50 * cn).astype(np.float64), rng.poisson(100, n).astype(np.float64),
                  rng.poisson(40 * cn).astype(np.float64)])
    r = ora.partition_hmm(np.array([0, n]), x, per_sample=False)
    bps = r["breakpoints"][0].tolist()
    assert any(abs(b - 5000) <= 3 for b in bps) and any(abs(b - 5600) <= 3 for b in bps)


def test_negative_binomial_reference_vector():
    # DistributionUtilitiesTests.TestNegativeBinomialWrapperMaxDensityEqualsMean (CanvasTest/DistributionUtilitiesTests.cs:39-49):
    # mean = variance = 50 over 200 values -> the density peaks at index 49 (the variance floor 1.2 * mean applies)
    d = ora.negative_binomial(50.0, 50.0, 200)
    assert len(d) == 200 and int(np.argmax(d)) == 49
    assert abs(float(np.sum(d)) - 1.0) < 1e-9


def test_genotype_arrangements_reference_vectors():
    # DistributionUtilitiesTests.TestGetGenotypeCombinations(+SingleSample) (:13-37): every assignment of {altCN, 2} to the
    # samples except "all diploid"; for altCN == 2 the one all-diploid arrangement.  The oracle and the kernels enumerate them
    # as bit masks (bit s set: sample s is diploid) — this is that enumeration, spelled out
    def arrangements(n_samples, alt):
        full = (1 << n_samples) - 1
        masks = [m for m in range(1 << n_samples) if (m == full) == (alt == 2)]
        return sorted([2 if (m >> s) & 1 else alt for s in range(n_samples)] for m in masks)
    assert arrangements(2, 1) == [[1, 1], [1, 2], [2, 1]]
    assert arrangements(1, 1) == [[1]]
    assert arrangements(3, 2) == [[2, 2, 2]]
