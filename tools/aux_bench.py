"""Device times of the kernels beside the headline path (CanvasBin passes, CanvasNormalize, CanvasSmooth) on WGS-sized inputs,
against the HBM roofline by their algorithmic bytes.  Kernel time = CUDA events on the library stream around the kernels of
one call (cg_last_kernel_ms), inputs already on the device; best of 3 calls after one warm-up.  One JSON line per kernel group.
Usage: python tools/aux_bench.py > profiles/<tag>_aux_kernels.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import native, synth


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbps", "hbm_GBps", "hbm_copy_gbps", "hbm_gb_s"):
            if k in d:
                return float(d[k])
        for v in d.values():
            if isinstance(v, dict):
                for k, x in v.items():
                    if "hbm" in k.lower() and isinstance(x, (int, float)):
                        return float(x)
    except Exception:
        pass
    return 6543.1


def run(name, nbytes, fn, note):
    fn()
    ms = min(fn() for _ in range(3))
    gbps = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": nbytes, "ms": ms, "achieved_GBps": gbps, "peak_GBps": PEAK,
                      "frac": gbps / PEAK, "note": note}), flush=True)


PEAK = peak()
eng = native.Engine(0)
rng = np.random.default_rng(0)

# CanvasBin passes on a 64 M-position chromosome piece
n = 64 * 1024 * 1024
hits = np.where(rng.random(n) < 0.3, rng.integers(1, 40, n), 0).astype(np.uint8)
possible = rng.random(n) < 0.85
fs = np.sort(rng.integers(0, n - 2000, 5000)).astype(np.int32)
fe = (fs + rng.integers(1, 2000, 5000)).astype(np.int32)
run("cg_bin_screen (bin_filter_clear + bin_screen)", 2.25 * n, lambda: eng.bin_screen(hits, possible, fs, fe)["kernel_ms"],
    "64 Mi positions, 5000 filter intervals; hits read + written, bitmap read + written")
bases = rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]).tobytes()
frag = np.where(rng.random(n) < 0.3, rng.integers(100, 900, n), 0).astype(np.int16)
run("cg_bin_read_gc (gc_tile_count + scan + gc_prefix + read_gc)", 5.0 * n, lambda: eng.bin_read_gc(bases, frag, 350, hits)["kernel_ms"],
    "64 Mi positions, mean fragment 350; bases + fragment lengths + hits read, read GC written (the 4-byte prefix array is internal)")
del hits, possible, bases, frag

# CanvasNormalize on config-2-sized bin lists
s = synth.make_sample(config=2)
nb = len(s.count)
S = 8
controls = np.stack([rng.poisson(np.maximum(s.count, 1.0) * rng.uniform(0.6, 1.6)).astype(np.float64) for _ in range(S)])
run("cg_normalize_reference (8 controls)", (8.0 * S + 8.0) * nb, lambda: eng.normalize_reference(controls)["kernel_ms"],
    "8 x 3.1 M doubles: 8 u64 radix-select passes over all controls for the medians, then one weighted stream")
sample = s.count.astype(np.float64)
run("cg_normalize_best_lr2 (8 controls)", 8.0 * (S + 1) * nb, lambda: eng.normalize_best_lr2(sample, controls)["kernel_ms"],
    "medians of 9 vectors (8 passes) + one pass of squared log ratios per control")
ref = eng.normalize_reference(controls)["reference"].astype(np.float32)
run("cg_normalize_ratio (lsnorm)", 8.0 * nb + 12.0 * nb, lambda: eng.normalize_ratio(s.count, ref)["kernel_ms"],
    "3.1 M bins: 4 u32 select passes for the two medians, keep filter + compaction with the ratio arithmetic in the emit")
K = 4
q, _ = np.linalg.qr(rng.normal(size=(nb, K)))
axes = (q.T * 20.0).copy()
mu = controls.mean(0).astype(np.float32)
run("cg_normalize_pca_reference (4 axes)", (8.0 * K + 12.0) * nb, lambda: eng.normalize_pca_reference(s.count, mu, axes)["kernel_ms"],
    "3.1 M bins: 4 + 6 + 4 chunk-wise dot products, projection stream, one u32 median select, scaled reference")
