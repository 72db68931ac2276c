"""Small inputs through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).
Results are still compared with the oracle where that is cheap, so a sanitizer run also proves the instrumented
kernels computed the right thing.
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_driver.py [fused|k8|cbs|hmm|bin|loess|pedigree|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import native, synth, textcodec
from oracle import pyoracle as ora

what = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
eng = native.Engine(0)
rng = np.random.default_rng(7)


def fused():
    s = synth.make_sample(config=2, sample=9, scale=scale, n_events=60)
    r = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc,
                                    is_germline=True, evenness_window=2000)
    o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    assert np.array_equal(r["kept_index"], o["kept_index"])
    assert np.array_equal(r["count"].view(np.uint32), o["count"].view(np.uint32))
    off = synth.chrom_offsets(s.chrom[o["kept_index"]], len(s.names))
    p = ora.partition_wavelet(off, ora.f2_roundtrip(o["count"]), is_germline=True, evenness_window=2000, n_threads=4)
    for a, b in zip(r["breakpoints"], p["breakpoints"]):
        assert a.tolist() == b.tolist()
    print("fused ok: bins", len(s), "breakpoints", sum(len(b) for b in r["breakpoints"]))


def k8():
    n, batch = 300_008, 3  # cg_normalize_apply takes n in multiples of 4
    count = rng.poisson(100, (batch, n)).astype(np.float32)
    gc = rng.integers(0, 101, (batch, n)).astype(np.uint8)
    med = rng.uniform(50, 150, (batch, 101))
    gmed = rng.uniform(90, 110, batch)
    out, _ = eng.normalize_apply(count, gc, med, gmed, repeats=1)
    exp = (gmed[:, None] * count.astype(np.float64) / np.take_along_axis(med, gc.astype(np.int64), 1)).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))
    print("k8 ok")


def cbs():
    lens = [2500, 900, 150, 20]
    cov = np.concatenate([np.round(np.where(np.arange(n) // 400 % 2 == 0, 100.0, 150.0) + rng.normal(0, 8, n), 2) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    want = ora.partition_cbs(off, cov)
    got = eng.partition_cbs(off, cov)
    for w, g in zip(want["segments"], got["segments"]):
        assert np.array_equal(w["len"], g["len"]) and np.array_equal(w["mean"], g["mean"])
    print("cbs ok: perms", got["perms"])


def hmm():
    lens = [6000, 257, 1025, 12]
    cov = np.concatenate([np.round(rng.poisson(np.where(np.arange(n) // 700 % 2 == 0, 100.0, 150.0)) + rng.uniform(0, 0.99, n), 2) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    want = ora.partition_hmm(off, cov, per_sample=True, n_threads=2)
    got = eng.partition_hmm(off, cov, per_sample=True)
    assert np.array_equal(got["states"], want["states"])
    print("hmm ok")


def binning():
    n = 300_000
    hits = np.where(rng.random(n) < 0.3, rng.integers(1, 40, n), 0).astype(np.uint8)
    possible = rng.random(n) < 0.85
    fs = np.sort(rng.integers(0, n - 2000, 50)).astype(np.int32)
    fe = (fs + rng.integers(1, 2000, 50)).astype(np.int32)
    eng.bin_screen(hits, possible, fs, fe)
    bases = rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=n).tobytes()
    r = eng.bin_hits(hits, possible, bases, 100, mode=0)
    frag = np.where(rng.random(n) < 0.3, rng.integers(100, 900, n), 0).astype(np.int16)
    eng.bin_read_gc(bases, frag, 350, hits)
    print("bin ok: bins", len(r["start"]))


def loess():
    s = synth.make_sample(config=2, sample=3, scale=scale)
    r = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, gc_mode=1)
    o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, gc_mode=1)
    assert np.array_equal(r["kept_index"], o["kept_index"])
    print("loess ok")


def pedigree():
    # the device-resident trio chain (loopback communicator) and a prefetched fused call
    eng.comm_init(1, 0)
    trio = [synth.make_sample(config=4, sample=k, scale=scale, n_events=20) for k in range(3)]
    t0 = trio[0]
    one = eng.pedigree_hmm(t0.chrom, t0.is_autosome, t0.is_chr_y, t0.start, t0.stop, [t.count for t in trio], t0.gc, sharded=True)
    c = [ora.clean(t.chrom, t.is_autosome, t.is_chr_y, t.start, t.stop, t.count, t.gc) for t in trio]
    assert [len(x["kept_index"]) for x in c] == one["n_kept"].tolist()
    common = np.intersect1d(np.intersect1d(c[0]["kept_index"], c[1]["kept_index"]), c[2]["kept_index"])
    assert np.array_equal(common, one["common_index"])
    cols = [np.ascontiguousarray(a, d) for a, d in ((t0.chrom, np.uint8), (t0.start, np.int32), (t0.stop, np.int32), (t0.count, np.float32), (t0.gc, np.uint8))]
    plain = eng.clean_partition_wavelet(cols[0], t0.is_autosome, t0.is_chr_y, cols[1], cols[2], cols[3], cols[4], evenness_window=2000)
    for _ in range(2):
        eng.prefetch_bins(*cols)
        pre = eng.clean_partition_wavelet(cols[0], t0.is_autosome, t0.is_chr_y, cols[1], cols[2], cols[3], cols[4], evenness_window=2000)
        assert all(a.tolist() == b.tolist() for a, b in zip(plain["breakpoints"], pre["breakpoints"]))
    print("pedigree + prefetch ok: common bins", one["n_common"])


runs = {"pedigree": pedigree, "fused": fused, "k8": k8, "cbs": cbs, "hmm": hmm, "bin": binning, "loess": loess}
rc = 0
for name, fn in runs.items():
    if what in ("all", name):
        try:
            fn()
        except Exception:
            import traceback
            traceback.print_exc()
            rc = 1
eng.close()
sys.exit(rc)
