"""CBS on a synthetic WGS sample: GPU timing, optional comparison with the oracle (threads = host cores)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canvas_b200 import native, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
check = len(sys.argv) > 2 and sys.argv[2] == "check"
eng = native.Engine(0)
s = synth.make_sample(config=2, scale=scale)
r = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
chrom = s.chrom[r["kept_index"]]
off = synth.chrom_offsets(chrom, len(s.names))
cov = np.round(r["count"].astype(np.float64), 2)
eng.cbs_boundary()
for it in range(2):
    t = time.time()
    got = eng.partition_cbs(off, cov)
    dt = time.time() - t
    print("gpu: bins %d call %.3fs kernel %.1f ms tests %d perms %d steps %.3g edge %.3g segs %d" % (
        len(cov), dt, got["kernel_ms"], got["tests"], got["perms"], got["perm_steps"], got["edge_steps"],
        sum(len(x["len"]) for x in got["segments"])), got["phase_ms"], flush=True)
if check:
    from oracle import pyoracle as po
    nt = os.cpu_count() or 1
    t = time.time()
    want = po.partition_cbs(off, cov, n_threads=nt)
    dt = time.time() - t
    print("oracle: %.2fs on %d threads; %.3g Mbins/s" % (dt, nt, len(cov) / dt / 1e6))
    same = all(np.array_equal(w["len"], g["len"]) and np.array_equal(w["mean"], g["mean"])
               for w, g in zip(want["segments"], got["segments"]))
    print("PARITY", "OK" if same and all(want[k] == got[k] for k in ("tests", "perms", "perm_steps", "edge_steps")) else "FAIL",
          {k: want[k] for k in ("tests", "perms", "perm_steps", "edge_steps")})
