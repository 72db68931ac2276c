"""Small driver for ncu: runs cg_clean + cg_partition_wavelet `iters` times on a synthetic sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from canvas_b200 import synth, native

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
what = sys.argv[3] if len(sys.argv) > 3 else "both"
eng = native.Engine(0)
s = synth.make_sample(config=2, scale=scale)
for it in range(iters):
    if what == "fused":
        r = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, is_germline=True)
        print("fused kernels ms", eng.last_kernel_ms, "launches", eng.last_launches, "bp", sum(len(b) for b in r["breakpoints"]),
              "stages", eng.last_stage_ms())
        continue
    if what in ("both", "clean"):
        r = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
        print("clean kernels ms", eng.last_kernel_ms, "launches", eng.last_launches, "kept", len(r["kept_index"]))
    else:
        r = {"kept_index": np.arange(len(s)), "count": s.count}
    if what in ("both", "partition"):
        chrom = s.chrom[r["kept_index"]]
        off = synth.chrom_offsets(chrom, len(s.names))
        cov = np.round(r["count"].astype(np.float64), 2)
        p = eng.partition_wavelet(off, cov, is_germline=True)
        print("partition kernels ms", eng.last_kernel_ms, "launches", eng.last_launches, "bp", sum(len(b) for b in p["breakpoints"]))
