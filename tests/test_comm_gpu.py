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
