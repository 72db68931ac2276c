"""PerSampleHMM on a whole config-2 genome: device time of cg_partition_hmm beside the oracle on the host cores.
Usage (GPU box): python tools/hmm_bench.py [scale]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canvas_b200 import native, synth, textcodec  # noqa: E402
from oracle import pyoracle as ora  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    eng = native.Engine(0)
    s = synth.make_sample(config=4, sample=0, scale=scale)
    c = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
    cov = textcodec.f2_roundtrip(c["count"])
    for _ in range(3):
        r = eng.partition_hmm(off, cov, per_sample=True)
    t0 = time.perf_counter()
    r = eng.partition_hmm(off, cov, per_sample=True)
    wall = time.perf_counter() - t0
    stages = eng.last_stage_ms()
    rs = eng.partition_hmm(off, cov, per_sample=True, exact_sequential=True)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    w = ora.partition_hmm(off, cov, per_sample=True, n_threads=threads)
    cpu = time.perf_counter() - t0
    same = np.array_equal(r["states"], w["states"]) and np.array_equal(rs["states"], w["states"])
    print(json.dumps({"bins": int(len(cov)), "gpu_kernel_ms": r["kernel_ms"], "gpu_call_ms": wall * 1e3, "stages_ms": stages,
                      "gpu_sequential_kernel_ms": rs["kernel_ms"], "launches": r["launches"],
                      "oracle_ms": cpu * 1e3, "oracle_threads": threads, "identical_paths": bool(same),
                      "breakpoints": int(sum(len(b) for b in r["breakpoints"])),
                      "Mbins_per_s_kernel": len(cov) / r["kernel_ms"] / 1e3}))
    eng.close()


if __name__ == "__main__":
    main()
