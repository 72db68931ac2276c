"""GPU-side check of cg_partition_cbs against the oracle on small seeded inputs (run under gpurun)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from canvas_b200 import native  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def chrom(rng, n, events=True):
    x = np.full(n, 100.0)
    if events:
        for _ in range(max(1, n // 3000)):
            s = int(rng.integers(0, max(1, n - 50)))
            ln = int(rng.integers(3, 800)) if rng.random() < 0.6 else int(rng.integers(2, 12))
            x[s:s + ln] *= rng.choice([0.5, 1.5, 2.0, 0.0, 1.1, 0.9, 1.25])
    return np.round(x + rng.normal(0, 8, n), 2)


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    lens = [int(v * scale) for v in (8000, 6000, 5000, 2500)] + [300, 150, 20, 3, 0, 5]
    cov = np.concatenate([chrom(rng, n, n > 200) for n in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    eng = native.Engine(0)
    b_gpu = eng.cbs_boundary()
    b_ora = po.cbs_boundary()
    print("boundary equal:", np.array_equal(b_gpu, b_ora))
    t = time.time()
    want = po.partition_cbs(off, cov)
    t_ora = time.time() - t
    t = time.time()
    got = eng.partition_cbs(off, cov)
    t_gpu = time.time() - t
    print("oracle %.2fs gpu call %.3fs kernel %.2f ms" % (t_ora, t_gpu, got["kernel_ms"]))
    print("oracle", {k: v for k, v in want.items() if k != "segments"})
    print("gpu   ", {k: v for k, v in got.items() if k != "segments"})
    ok = True
    for c, (w, g) in enumerate(zip(want["segments"], got["segments"])):
        same = np.array_equal(w["len"], g["len"]) and np.array_equal(w["mean"], g["mean"])
        ok &= same
        if not same:
            print("chrom", c, "n", lens[c], "\n  want", w["len"][:20], "\n  got ", g["len"][:20])
    print("PARITY", "OK" if ok else "FAIL")


if __name__ == "__main__":
    main()
