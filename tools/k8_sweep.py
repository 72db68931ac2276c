"""K8 (normalise apply) on batches larger than L2: variants x batch size x CTAs per SM.
Usage (GPU box): python tools/k8_sweep.py [quick]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canvas_b200 import native, synth  # noqa: E402


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    eng = native.Engine(0)
    s = synth.make_sample(config=2, sample=0, scale=1.0)
    n = (len(s) // 2048) * 2048 if os.environ.get("K8_ALIGN_TILE") else (len(s) // 16) * 16
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.default_rng(1)
    ref = None
    for batch in ([8] if quick else [8, 16, 32]):
        cnt = np.tile(s.count[:n], (batch, 1))
        gcb = np.tile(s.gc[:n], (batch, 1))
        med = rng.uniform(80, 120, (batch, 101))
        gm = np.full(batch, 100.0)
        for variant in ["reciprocal", "divide", "reciprocal", "divide"]:
            for cps in [4]:
                if variant == "divide":
                    os.environ["CANVAS_K8_EXACT_DIV"] = "1"
                else:
                    os.environ.pop("CANVAS_K8_EXACT_DIV", None)
                out, ms = eng.normalize_apply(cnt, gcb, med, gm, repeats=int(os.environ.get('K8_REPEATS', '20')))
                if batch == 8:
                    if ref is None:
                        ref = out.copy()
                    else:
                        assert np.array_equal(ref.view(np.uint32), out.view(np.uint32)), "variants disagree"
                gbs = 9.0 * batch * n / (ms * 1e-3) / 1e9
                print(json.dumps({"variant": variant, "batch": batch, "ctas_per_sm": cps, "bins": batch * n,
                                  "ms": round(ms, 5), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
