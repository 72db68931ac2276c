"""Golden digest of CanvasClean -m LOESS on the WHOLE config-2 sample (3.1 M bins): the oracle needs ~2 minutes for it (the
golden-section bandwidth search of LoessGCNormalizer evaluates O(n) fits), too long for the GPU test tier, so its result is
committed as a fixture: the kept bins as a SHA-256, every 997th normalised count, the local-SD metric.
Usage: python tools/make_golden_loess_full.py   (writes tests/golden/loess_config2_full.json)"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import synth
from oracle import pyoracle as ora

SAMPLE, STEP = 1, 997
s = synth.make_sample(config=2, sample=SAMPLE)
o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, gc_mode=1)
out = {"config": 2, "sample": SAMPLE, "bins": int(len(s)), "kept": int(len(o["kept_index"])),
       "kept_sha256": hashlib.sha256(np.ascontiguousarray(o["kept_index"], np.int32).tobytes()).hexdigest(),
       "local_sd": float(o["local_sd"]), "step": STEP, "counts": [float(v) for v in o["count"][::STEP]],
       "how": "oracle/clean.cpp + oracle/loess.cpp (pinned by TestLoessInterpolator), gc_mode=1, all filters on"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loess_config2_full.json")
json.dump(out, open(path, "w"))
print(path, out["bins"], out["kept"], out["local_sd"])
