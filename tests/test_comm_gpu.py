"""The multi-GPU entry points of the C-ABI on one GPU: with a loopback communicator (n_ranks = 1, no NCCL call) every
*_sharded call must equal its single-GPU form, and the variable-length all-gather must survive lists beyond its
fixed first-round capacity.  The N > 1 path runs in tools/multi_gpu_check.py (gpurun --gpus 2 / 8)."""
import numpy as np
import pytest

from canvas_b200 import native, synth, textcodec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = native.Engine(0)
    e.comm_init(1, 0)
    assert e.comm_size == 1 and e.comm_rank == 0
    yield e
    e.close()


def _same_bp(a, b):
    for c, (x, y) in enumerate(zip(a, b)):
        assert x.tolist() == y.tolist(), c


def test_sharded_calls_need_a_communicator():
    e = native.Engine(0)
    with pytest.raises(native.CanvasGpuError) as err:
        e.partition_wavelet(np.array([0, 100]), np.full(100, 50.0), sharded=True)
    assert err.value.code == native.CG_ERR_ARG
    e.close()


def test_wavelet_sharded_equals_plain(eng):
    s = synth.make_sample(config=2, sample=4, scale=0.05, n_events=80)
    plain = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, evenness_window=5000)
    shard = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, evenness_window=5000,
                                        sharded=True)
    _same_bp(plain["breakpoints"], shard["breakpoints"])
    assert np.array_equal(plain["kept_index"], shard["kept_index"]) and plain["cv"] == shard["cv"]
    assert set(shard["owner"].tolist()) == {0}
    assert sum(len(b) for b in plain["breakpoints"]) > 30
    off = plain["chrom_off"]
    cov = textcodec.f2_roundtrip(plain["count"])
    p2 = eng.partition_wavelet(off, cov, evenness_window=5000)
    s2 = eng.partition_wavelet(off, cov, evenness_window=5000, sharded=True)
    _same_bp(p2["breakpoints"], s2["breakpoints"])
    _same_bp(p2["breakpoints"], plain["breakpoints"])
    assert eng.last_exchange_ms >= 0


def test_wavelet_sharded_more_breakpoints_than_the_first_round_holds(monkeypatch):
    # first-round capacity lowered to 512 ints (CANVAS_COMM_PACK_INTS): ~1000 breakpoints travel in the exact second round
    monkeypatch.setenv("CANVAS_COMM_PACK_INTS", "512")
    e = native.Engine(0)
    e.comm_init(1, 0)
    rng = np.random.default_rng(3)
    n = 300_000
    lev = rng.choice([100., 2000., 4000., 8000., 16000.], n // 15 + 1)
    cov = np.round(np.repeat(lev, 15)[:n] + rng.normal(0, 0.5, n), 2)
    off = np.array([0, 200_000, n])
    plain = e.partition_wavelet(off, cov, evenness_window=20000)
    shard = e.partition_wavelet(off, cov, evenness_window=20000, sharded=True)
    assert sum(len(b) for b in plain["breakpoints"]) > 600
    _same_bp(plain["breakpoints"], shard["breakpoints"])
    big = np.arange(5000, dtype=np.int32)
    assert np.array_equal(e.allgather_lists(big)[0], big)
    e.close()


def test_cbs_and_hmm_sharded_equal_plain(eng):
    rng = np.random.default_rng(5)
    lens = [4000, 2500, 300, 0, 5]
    cov = np.concatenate([np.round(np.where(np.arange(n) // 500 % 2 == 0, 100.0, 140.0) + rng.normal(0, 8, n), 2) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    a, b = eng.partition_cbs(off, cov), eng.partition_cbs(off, cov, sharded=True)
    for x, y in zip(a["segments"], b["segments"]):
        assert np.array_equal(x["len"], y["len"]) and np.array_equal(x["mean"], y["mean"])
    cov = np.abs(cov)
    a, b = eng.partition_hmm(off, cov), eng.partition_hmm(off, cov, sharded=True)
    _same_bp(a["breakpoints"], b["breakpoints"])
    assert np.array_equal(a["states"], b["states"])


def test_allgather_lists_and_empty_input(eng):
    for n in (0, 5, 16383, 16384, 70_000):
        loc = np.arange(n, dtype=np.int32) * 3 - 7
        got = eng.allgather_lists(loc)
        assert len(got) == 1 and np.array_equal(got[0], loc)
    r = eng.clean_partition_wavelet(np.zeros(0, np.uint8), np.ones(2, np.uint8), np.zeros(2, np.uint8), np.zeros(0, np.int32),
                                    np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.uint8), sharded=True)
    assert len(r["kept_index"]) == 0 and all(len(b) == 0 for b in r["breakpoints"])


def test_pedigree_chain_one_call_equals_the_staged_chain(eng):
    # cg_pedigree_hmm (Clean x S -> common bins -> PerSampleHMM, device resident) against the same stages called one by one
    # through host memory, and against the oracle; with and without the (loopback) communicator
    from oracle import pyoracle as ora
    trio = [synth.make_sample(config=4, sample=k, scale=0.04, n_events=40) for k in range(3)]
    t0 = trio[0]
    one = eng.pedigree_hmm(t0.chrom, t0.is_autosome, t0.is_chr_y, t0.start, t0.stop, [t.count for t in trio], t0.gc)
    cleaned = [eng.clean(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, t.count, t.gc) for t in trio]
    for k, c in enumerate(cleaned):
        o = ora.clean(trio[k].chrom, trio[k].is_autosome, trio[k].is_chr_y, trio[k].start, trio[k].stop, trio[k].count, trio[k].gc)
        assert np.array_equal(c["kept_index"], o["kept_index"])
        assert one["n_kept"][k] == len(c["kept_index"]) and one["local_sd"][k] == c["local_sd"] == o["local_sd"]
    m = eng.merge_kept_indices(len(t0), [c["kept_index"].copy() for c in cleaned], [c["count"].copy() for c in cleaned])
    assert one["n_common"] == len(m["common_index"]) > 1000
    assert np.array_equal(one["common_index"], m["common_index"])
    assert np.array_equal(one["count"].view(np.uint32), m["count"].view(np.uint32))
    off = np.searchsorted(m["common_index"], synth.chrom_offsets(t0.chrom, len(t0.names))).astype(np.int64)
    assert np.array_equal(one["chrom_off"], off)
    total = 0
    for k in range(3):
        want = ora.partition_hmm(off, textcodec.float_default_roundtrip(m["count"][k]), per_sample=True, n_threads=4)
        _same_bp(one["breakpoints"][k], want["breakpoints"])
        total += sum(len(b) for b in want["breakpoints"])
    assert total > 20
    loop = eng.pedigree_hmm(t0.chrom, t0.is_autosome, t0.is_chr_y, t0.start, t0.stop, [t.count for t in trio], t0.gc, sharded=True)
    assert loop["n_common"] == one["n_common"] and set(loop["owner"].ravel().tolist()) == {0}
    for k in range(3):
        _same_bp(loop["breakpoints"][k], one["breakpoints"][k])


def test_pedigree_chain_edge_cases(eng):
    t = synth.make_sample(config=4, sample=0, scale=0.02, n_events=10)
    # one sample: the "common" bins are its own cleaned bins
    one = eng.pedigree_hmm(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, [t.count], t.gc)
    c = eng.clean(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, t.count, t.gc)
    assert np.array_equal(one["common_index"], c["kept_index"]) and np.array_equal(one["count"][0].view(np.uint32), c["count"].view(np.uint32))
    # a rank that does not write the merged files skips the table download: same breakpoints, no tables
    lean = eng.pedigree_hmm(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, [t.count], t.gc, want_tables=False)
    assert lean["common_index"] is None and lean["count"] is None and lean["n_common"] == one["n_common"]
    _same_bp(lean["breakpoints"][0], one["breakpoints"][0])
    # empty layout
    z = np.zeros(0, np.int32)
    e = eng.pedigree_hmm(np.zeros(0, np.uint8), t.is_autosome, t.is_chr_y, z, z, [np.zeros(0, np.float32)] * 2, np.zeros(0, np.uint8))
    assert e["n_common"] == 0 and all(len(b) == 0 for per in e["breakpoints"] for b in per) and e["gc_norm_skipped"].all()
    # a sample this GPU has to clean must come with its counts
    with pytest.raises(native.CanvasGpuError) as err:
        eng.pedigree_hmm(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, [t.count, None], t.gc)
    assert err.value.code == native.CG_ERR_ARG
    # unsorted chromosome ids are refused as in cg_clean
    bad = t.chrom.copy()
    bad[10] = bad.max()
    with pytest.raises(native.CanvasGpuError) as err:
        eng.pedigree_hmm(bad, t.is_autosome, t.is_chr_y, t.start, t.stop, [t.count], t.gc)
    assert err.value.code == native.CG_ERR_UNSORTED
