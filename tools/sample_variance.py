"""Device time of the fused Clean + Partition call on each of the eight config-5 samples (bench.py gives sample r to rank r,
and the reported step is the maximum over ranks): which sample is slow, and in which stage.
Usage: python tools/sample_variance.py [n_samples]   (CANVAS_DEBUG=1 adds the per-chromosome pipeline timeline)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canvas_b200 import native, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else range(n)
eng = native.Engine(0)
for k in only:
    s = synth.make_sample(config=2, sample=k)
    best = None
    for it in range(4):
        r = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, is_germline=True)
        ms = eng.last_kernel_ms
        if it > 0 and (best is None or ms < best[0]):
            best = (ms, eng.last_stage_ms(), eng.last_partition_stats())
    ms, st, ps = best
    print(json.dumps({"sample": k, "device_ms": round(ms, 3), "stages_ms": {a: round(b, 3) for a, b in st.items()},
                      "breakpoints": int(sum(len(b) for b in r["breakpoints"])), "l_eff": round(ps["visits"] / ps["bins"], 1),
                      "max_depth": ps["max_depth"], "big_phase_ms": round(ps["big_phase_ms"], 3),
                      "decompose_span_ms": round(ps["decompose_span_ms"], 3), "candidates": ps["candidates"]}), flush=True)
eng.close()
