"""CanvasBin counting kernels at chr1 scale (BASELINE north_star's "bin-count histogramming"): device time of
cg_bin_hits (TruncatedDynamicRange and GCContentWeighted), cg_bin_screen, cg_bin_read_gc and cg_bin_fragments against the
HBM roofline by SURVEY 8(d)'s algorithmic bytes (2.125 B / genome position for the hit-array binning).  Kernel time = CUDA
events on the library stream around the kernels of one call (cg_last_kernel_ms), inputs already on the device; best of 3
calls after one warm-up.  One JSON line per entry point.  With `check`, the integer results of the big run are compared
with the oracle (bit-exact).
Usage: python tools/bin_bench.py [positions (default 249e6)] [check] > profiles/<tag>_bin_bench.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import native


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6543.1


PEAK = peak()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 and sys.argv[1] != "check" else 249_250_621  # hg19 chr1
check = "check" in sys.argv
eng = native.Engine(0)
rng = np.random.default_rng(1)


def emit(name, nbytes, ms, note, **extra):
    gbps = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "positions": n, "algorithmic_bytes": nbytes, "ms": ms, "achieved_GBps": gbps, "peak_GBps": PEAK,
                      "frac": gbps / PEAK, "Mpositions_per_s": n / ms / 1e3, "note": note, **extra}), flush=True)


def best(fn, reps=3):
    fn()
    return min(fn() for _ in range(reps))


t0 = time.time()
# 30x WGS of 100-bp reads: ~0.15 alignment starts per position per strand; most positions 0, a few piled up
hits = rng.poisson(0.3, n).astype(np.uint8)
hits[rng.integers(0, n, n // 5000)] = 200
possible = rng.random(n) < 0.85
possible[: n // 250] = False                      # leading N block
bases_arr = rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), size=n, p=[.2, .2, .2, .2, .045, .045, .045, .045, .02])
bases_arr[: n // 250] = ord("n")
bases = bases_arr.tobytes()
bin_size = 850                                   # ~1 kb bins at 85 % uniqueness
sys.stderr.write("inputs in %.1f s\n" % (time.time() - t0))


def hits_ms(mode, **kw):
    def f():
        r = eng.bin_hits(hits, possible, bases, bin_size, mode=mode, **kw)
        f.last = r
        return eng.last_kernel_ms
    return f


f0 = hits_ms(0)
ms0 = best(f0)
extra = {}
if check:
    from oracle import pyoracle as ora
    w = ora.bin_hits(hits, possible, bases, bin_size, mode=0)
    extra["parity_bit_exact"] = bool(all(np.array_equal(f0.last[k], w[k]) for k in ("start", "stop", "count", "gc")))
emit("cg_bin_hits mode 0 (TruncatedDynamicRange)", 2.125 * n + 14.0 * len(f0.last["start"]), ms0,
     "hits + bases 1 B/position, possible-position bitmap 1/8 B/position, 14 B per bin written", bins=int(len(f0.last["start"])), **extra)

read_gc = np.clip(np.rint(rng.normal(41, 8, n)), 0, 100).astype(np.uint8)
ratio = rng.uniform(0.5, 1.5, 101).astype(np.float32)
f1 = hits_ms(1, read_gc=read_gc, obs_vs_exp_gc=ratio)
ms1 = best(f1)
extra = {}
if check:
    w = ora.bin_hits(hits, possible, bases, bin_size, mode=1, read_gc=read_gc, obs_vs_exp=ratio)
    extra["parity_bit_exact"] = bool(all(np.array_equal(f1.last[k], w[k]) for k in ("start", "stop", "count", "gc")))
emit("cg_bin_hits mode 1 (GCContentWeighted)", 3.125 * n + 14.0 * len(f1.last["start"]), ms1,
     "as mode 0 plus the read-GC byte per position; sequential float adds per bin as the reference", **extra)
bins = f0.last

fs = np.sort(rng.integers(0, n - 2000, 20000)).astype(np.int32)
fe = (fs + rng.integers(1, 2000, 20000)).astype(np.int32)
ms = best(lambda: eng.bin_screen(hits, possible, fs, fe)["kernel_ms"])
emit("cg_bin_screen", 2.25 * n, ms, "hits read + written where cleared, bitmap read + written; 20000 filter intervals")

frag = np.where(rng.random(n) < 0.15, rng.integers(100, 900, n), 0).astype(np.int16)
ms = best(lambda: eng.bin_read_gc(bases, frag, 350, hits)["kernel_ms"])
emit("cg_bin_read_gc", 5.0 * n, ms, "bases + fragment lengths + hits read, read GC written (the prefix array is internal)")
del frag, read_gc

# fragment binning: 3e7 fragments (30x of 2x150 reads over chr1), ~400 bp, sorted by start as a coordinate-sorted BAM gives them
nf = int(3e7 * n / 249_250_621)
fstart = np.sort(rng.integers(0, n - 1000, nf)).astype(np.int32)
fstop = (fstart + rng.integers(150, 700, nf)).astype(np.int32)


def frag_call():
    r = eng.bin_fragments(fstart, fstop, bins["start"], bins["stop"])
    frag_call.last = r
    return eng.last_kernel_ms


ms = best(frag_call)
extra = {}
if check:
    # FindBestBin restated in numpy (FragmentBinner.cs:353-371): first bin whose stop lies right of the fragment start,
    # then the largest overlap among the following bins, the first one on ties (fragments span at most a few bins here)
    bs, be = bins["start"].astype(np.int64), bins["stop"].astype(np.int64)
    lo = np.searchsorted(be, fstart, side="right")
    best_bin = np.full(nf, -1, np.int64)
    best_ov = np.zeros(nf, np.int64)
    alive = np.ones(nf, bool)
    for k in range(8):
        b = lo + k
        ok = alive & (b < len(bs))
        bb = np.minimum(b, len(bs) - 1)
        ov = np.minimum(be[bb], fstop) - np.maximum(bs[bb], fstart)
        ok &= ov > 0
        alive = ok
        better = ok & (ov > best_ov)
        best_bin[better] = b[better]
        best_ov[better] = ov[better]
    assert not alive.any()
    cnt = np.bincount(best_bin[best_bin >= 0], minlength=len(bs))
    extra["parity_bit_exact"] = bool(np.array_equal(frag_call.last["count"], cnt) and np.array_equal(frag_call.last["best_bin"], best_bin))
emit("cg_bin_fragments", 12.0 * nf + 12.0 * len(bins["start"]), ms,
     "fragment start/stop read + best bin written (12 B/fragment), bin start/stop read + count written (12 B/bin)",
     fragments=nf, Mfragments_per_s=nf / ms / 1e3, **extra)
eng.close()
