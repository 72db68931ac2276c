#!/usr/bin/env python
"""Extract the known-answer vectors the reference's own unit tests hold for the hot path into
tests/golden/*.json.  Run in the build container (needs /root/reference); the JSON fixtures are
committed so nothing at test time reads the reference tree.

Sources (reference @ v1.40.0, Src/Canvas/CanvasTest/):
  CanvasPartition/WaveletTests.cs:9-90      coverage[550] -> 12 breakpoints
  TestLoessInterpolator.cs:11-81            x, y, fittedR, weightedFittedR (R loess span=.3 degree=1)
  TestUtilities.cs:33-41                    golden-section intervals
  TestUtilities.cs:195-206                  median-filter vector (pins SortedList<float>.Median)
"""
import json
import os
import re
import sys

REF = "/root/reference/Src/Canvas/CanvasTest"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def array_after(src, name):
    m = re.search(name + r"\s*=\s*new\s+\w+\[\]\s*\{(.*?)\}", src, flags=re.S)
    body = strip_comments(m.group(1))
    return [float(t.rstrip("f")) for t in body.replace("\n", " ").split(",") if t.strip()]


def main():
    os.makedirs(OUT, exist_ok=True)
    src = open(os.path.join(REF, "CanvasPartition", "WaveletTests.cs"), encoding="utf-8-sig").read()
    cov = array_after(src, "coverage")
    expected = [int(v) for v in re.findall(r"Assert\.Equal\((\d+), breakpoints\[\d+\]\)", src)]
    count = int(re.search(r"Assert\.Equal\((\d+), breakpoints\.Count\)", src).group(1))
    assert len(expected) == count == 12 and len(cov) == 550, (len(expected), count, len(cov))
    json.dump({"source": "CanvasTest/CanvasPartition/WaveletTests.cs:9-90",
               "call": {"cv_window": 11, "thr_lower": 5, "thr_upper": 80, "is_germline": False,
                        "mad_factor": 5},
               "coverage": cov, "breakpoints": expected},
              open(os.path.join(OUT, "wavelet_minimal.json"), "w"))

    src = open(os.path.join(REF, "TestLoessInterpolator.cs"), encoding="utf-8-sig").read()
    d = {"source": "CanvasTest/TestLoessInterpolator.cs:11-81", "bandwidth": 0.3, "x_step": 0.01,
         "bound": 0.31}
    for k in ("x", "y", "fittedR", "weightedFittedR"):
        d[k] = array_after(src, r"\b" + k)
    assert len(d["x"]) == len(d["y"]) == len(d["fittedR"]) == len(d["weightedFittedR"])
    json.dump(d, open(os.path.join(OUT, "loess_train.json"), "w"))

    json.dump({"source": "CanvasTest/TestUtilities.cs:33-41,195-206",
               "golden_section": {"intervals": [[-5, 5], [0, 5], [-5, 0]], "abs_bound": 1e-3},
               "median_filter": {"values": [2, 1, 3, 5, 4, 6, 7, 8], "half_window": 1,
                                 "expected": [1.5, 2, 3, 4, 5, 6, 7, 7.5]}},
              open(os.path.join(OUT, "utilities.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
