"""Diagnostic run of cg_partition_wavelet against the oracle with verbose output (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from canvas_b200 import synth, native
from oracle import pyoracle as ora

eng = native.Engine(0)
print(eng.describe())

def run(label, off, cov, **kw):
    t = time.time(); a = eng.partition_wavelet(off, cov, **kw); ta = time.time() - t
    okw = dict(kw); okw["n_threads"] = 8
    t = time.time(); b = ora.partition_wavelet(off, cov, **okw); tb = time.time() - t
    print(f"== {label}: N={len(cov)} gpu {ta*1e3:.1f} ms (kernels {eng.last_kernel_ms:.3f} ms, {eng.last_launches} launches) oracle {tb*1e3:.1f} ms")
    print("   cv", a["cv"], b["cv"], "evenness", a["evenness"], b["evenness"])
    print("   f3 equal", np.array_equal(a["factor_of_three"], b["factor_of_three"]), a["factor_of_three"][:4], b["factor_of_three"][:4])
    bad = 0
    for c, (x, y) in enumerate(zip(a["breakpoints"], b["breakpoints"])):
        if x.tolist() != y.tolist():
            bad += 1
            print("   MISMATCH chrom", c, "gpu", x.tolist()[:30], "ora", y.tolist()[:30])
    print("   chromosomes mismatching:", bad, "total bp", sum(len(x) for x in b["breakpoints"]))

import json
g = json.load(open("tests/golden/wavelet_minimal.json"))
cov = np.array(g["coverage"]); off = np.array([0, len(cov)], np.int64)
run("golden", off, cov, is_germline=False, mad_factor=5.0, thr_lower=5.0, thr_upper=80.0, min_size=10, evenness_window=11)
print("   expected", g["breakpoints"])

def cleaned(s):
    r = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    chrom = s.chrom[r["kept_index"]]
    return synth.chrom_offsets(chrom, len(s.names)), np.round(r["count"].astype(np.float64), 2)

s = synth.make_sample(config=1, chromosomes=["chr20"]); off, cov = cleaned(s)
run("chr20 germline", off, cov, is_germline=True)
run("chr20 somatic", off, cov, is_germline=False)
s = synth.make_sample(config=2, sample=6, scale=0.12, n_events=300); off, cov = cleaned(s)
run("scaled germline", off, cov, is_germline=True, evenness_window=12000)
if len(sys.argv) > 1 and sys.argv[1] == "full":
    s = synth.make_sample(config=2); off, cov = cleaned(s)
    run("full germline", off, cov, is_germline=True)
    run("full germline (2nd call)", off, cov, is_germline=True)
