"""cg_partition_hmm against the oracle: identical Viterbi paths and breakpoints (integer results: exact), for the
PerSampleHMM default, the joint HMM mode with one and three samples, the strictly sequential cross-check kernel,
chromosome sharding, and the reference's edge cases."""
import numpy as np
import pytest

from canvas_b200 import native, synth, textcodec
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def _coverage(rng, lens, haploid=50.0, events_per=3000):
    out = []
    for n in lens:
        cn = np.full(n, 2)
        for _ in range(max(1, n // events_per)):
            a = int(rng.integers(0, max(1, n - 5)))
            cn[a:a + int(rng.integers(3, 400))] = rng.choice([0, 1, 3, 4])
        lam = haploid * np.maximum(cn, 0.03) * rng.uniform(0.9, 1.1)
        out.append(np.round(rng.poisson(lam).astype(np.float64) + rng.uniform(0, 0.99, n), 2))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return off, np.concatenate(out)


def _same(got, want):
    assert np.array_equal(got["states"], want["states"])
    for c, (a, b) in enumerate(zip(got["breakpoints"], want["breakpoints"])):
        assert a.tolist() == b.tolist(), c


@pytest.mark.parametrize("lens", [[300], [257, 258, 12, 1025], [40000, 7, 25000, 11, 10], [250000, 60000]])
def test_per_sample_matches_oracle(engine, lens):
    off, cov = _coverage(np.random.default_rng(len(lens) * 7 + lens[0]), lens)
    want = ora.partition_hmm(off, cov, per_sample=True, n_threads=4)
    _same(engine.partition_hmm(off, cov, per_sample=True), want)
    _same(engine.partition_hmm(off, cov, per_sample=True, exact_sequential=True), want)


def test_joint_one_sample_matches_oracle(engine):
    off, cov = _coverage(np.random.default_rng(3), [30000, 9000, 1200])
    want = ora.partition_hmm(off, cov, per_sample=False, n_threads=4)
    _same(engine.partition_hmm(off, cov, per_sample=False), want)


def test_joint_three_samples_matches_oracle(engine):
    rng = np.random.default_rng(4)
    lens = [20000, 5000]
    off, a = _coverage(rng, lens, haploid=50.0)
    _, b = _coverage(rng, lens, haploid=45.0)
    _, c = _coverage(rng, lens, haploid=60.0)
    cov = np.stack([a, b, c])
    want = ora.partition_hmm(off, cov, per_sample=False, n_threads=4)
    got = engine.partition_hmm(off, cov, per_sample=False)
    # the device takes the logarithm of the joint emission itself (libdevice log, <= 1 ulp from libm): paths are
    # compared exactly and have matched on every seed tried; a flip would need two scores within ~1e-13
    _same(got, want)


def test_config2_whole_genome_after_clean(engine):
    s = synth.make_sample(config=4, sample=1, scale=0.3, n_events=60)
    c = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
    cov = textcodec.f2_roundtrip(c["count"])
    want = ora.partition_hmm(off, cov, per_sample=True, n_threads=8)
    got = engine.partition_hmm(off, cov, per_sample=True)
    _same(got, want)
    assert sum(len(b) for b in got["breakpoints"]) > len(s.names)  # the planted events were found


def test_sharded_union_equals_whole(engine):
    off, cov = _coverage(np.random.default_rng(5), [9000, 300, 20000, 4000, 700])
    whole = engine.partition_hmm(off, cov, per_sample=True)
    mask = np.array([1, 0, 0, 1, 1], np.uint8)
    a = engine.partition_hmm(off, cov, per_sample=True, chrom_selected=mask)
    b = engine.partition_hmm(off, cov, per_sample=True, chrom_selected=1 - mask)
    for c in range(5):
        pick = a if mask[c] else b
        other = b if mask[c] else a
        assert pick["breakpoints"][c].tolist() == whole["breakpoints"][c].tolist()
        assert other["breakpoints"][c].tolist() == []


def test_zero_probability_states_follow_the_reference(engine):
    # a coverage scale at which the CN 3/4 densities underflow to exactly 0 at empty bins: log 0 = -inf and the
    # scores sit at Double.MinValue for a step (HMM.cs:91-99); the blocked pass must still agree
    rng = np.random.default_rng(6)
    n = 5000
    cov = rng.poisson(3000, n).astype(np.float64)
    cov[1000:1040] = 0.0
    cov[2047:2052] = 0.0
    off = np.array([0, n])
    want = ora.partition_hmm(off, cov, per_sample=True)
    _same(engine.partition_hmm(off, cov, per_sample=True), want)
    _same(engine.partition_hmm(off, cov, per_sample=True, exact_sequential=True), want)


def test_edge_cases(engine):
    r = engine.partition_hmm(np.array([0], np.int64), np.zeros(0))
    assert r["breakpoints"] == []
    off = np.array([0, 10, 10, 21])
    cov = np.concatenate([np.full(10, 100.0), np.full(11, 100.0)])
    r = engine.partition_hmm(off, cov)  # 10 bins: not longer than MinSize; empty chromosome; 11 bins: segmented
    assert [b.tolist() for b in r["breakpoints"]] == [[], [], [0]]
    with pytest.raises(native.CanvasGpuError) as e:
        engine.partition_hmm(np.array([0, 20]), np.concatenate([np.full(19, 50.0), [-1.0]]))
    assert e.value.code == native.CG_ERR_ARG
    with pytest.raises(native.CanvasGpuError):
        engine.partition_hmm(np.array([0, 20]), np.concatenate([np.full(19, 50.0), [np.nan]]))


def test_counts_entry_point_reproduces_the_text_round_trips(engine):
    # cg_partition_hmm_counts: float counts in, the .cleaned text round trip on the device — same paths as the double entry
    # point fed with the host's emulation of float.ToString("F2") / float.ToString()
    rng = np.random.default_rng(12)
    lens = [30000, 257, 9000]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cn = np.repeat(rng.choice([1, 2, 3], 40), sum(lens) // 40 + 1)[:sum(lens)]
    cnt = (rng.poisson(50.0 * cn) * rng.uniform(0.97, 1.03, sum(lens))).astype(np.float32)
    cnt[::977] = 0.0
    cnt[5::1201] *= 1e-3  # small values: other decimal exponents
    for mode, conv in ((1, textcodec.f2_roundtrip), (2, textcodec.float_default_roundtrip), (0, lambda v: v.astype(np.float64))):
        want = engine.partition_hmm(off, conv(cnt), per_sample=True)
        got = engine.partition_hmm_counts(off, cnt, text_mode=mode, per_sample=True)
        _same(got, want)
    mask = np.array([1, 0, 1], np.uint8)
    want = engine.partition_hmm(off, textcodec.float_default_roundtrip(cnt), per_sample=True, chrom_selected=mask)
    got = engine.partition_hmm_counts(off, cnt, text_mode=2, per_sample=True, chrom_selected=mask)
    assert all(a.tolist() == b.tolist() for a, b in zip(got["breakpoints"], want["breakpoints"]))
