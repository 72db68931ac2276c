"""CPU study for the next selection kernel (point 1): how small is the bracket a per-segment sample puts
around a requested rank, and how often does it miss?  For every segment the partition and Clean select on (chromosomes,
10 000- and 100 000-bin windows, GC buckets, the global list) the script draws a strided sample of m elements, brackets the
median rank by +-z standard deviations of the sample rank, and reports the share of elements that fall inside the bracket
(what a gather pass would have to hold in shared memory) and whether the true median is inside.
Usage: python tools/select_bracket_study.py > profiles/<tag>_select_bracket_study.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import synth


def study(name, segments, m, z):
    inside_max, inside_sum, total, misses, nseg = 0, 0, 0, 0, 0
    for x in segments:
        n = len(x)
        if n < 4:
            continue
        nseg += 1
        s = np.sort(x[np.linspace(0, n - 1, min(m, n)).astype(np.int64)])  # evenly spaced sample over the whole segment
        k = len(s)
        sd = 0.5 * np.sqrt(k)                      # sd of the sample rank of the population median
        lo_r = int(max(0, np.floor(k / 2 - z * sd - 1)))
        hi_r = int(min(k - 1, np.ceil(k / 2 + z * sd + 1)))
        lo, hi = s[lo_r], s[hi_r]
        inside = int(np.count_nonzero((x >= lo) & (x <= hi)))
        med = np.partition(x, n // 2)[n // 2]
        if not (lo <= med <= hi):
            misses += 1
        inside_max = max(inside_max, inside)
        inside_sum += inside
        total += n
    print(f"{name:34s} m={m:5d} z={z:.0f}: segments {nseg:4d}, elements {total:9d}, inside the brackets {100.0 * inside_sum / max(total, 1):5.2f} % "
          f"(largest bracket {inside_max:6d} elements = {inside_max * 8 / 1024:6.1f} KiB of 8-byte keys), misses {misses}")


def main():
    s = synth.make_sample(config=2)
    # coverage as CanvasPartition sees it: counts scaled per GC bucket to the global median (a plain numpy stand-in for
    # NormalizeByGC, good enough for the distribution of values), stored as float and printed with two decimals
    gmed = np.median(s.count)
    med = np.array([np.median(s.count[s.gc == g]) if np.any(s.gc == g) else 1.0 for g in range(101)])
    med[med <= 0] = 1.0
    cov = np.round((gmed * s.count.astype(np.float64) / med[s.gc]).astype(np.float32).astype(np.float64), 2)
    chroms = [cov[s.chrom == c] for c in range(len(s.names))]
    w10 = [c[i:i + 10000] for c in chroms for i in range(0, max(0, len(c) - 10000), 10000)]
    w100 = [c[i:i + 100000] for c in chroms for i in range(0, max(0, len(c) - 100000), 100000)]
    gcb = [s.count[(s.gc == g)] for g in range(101)]
    print("# synthetic config 2 (3.1 M bins); sample = m evenly spaced elements of the segment, bracket = sample ranks k/2 +- z * sqrt(k) / 2")
    for m, z in ((1024, 4), (2048, 4), (4096, 4), (4096, 5)):
        study("chromosomes (coverage)", chroms, m, z)
        study("100 000-bin windows", w100, m, z)
        study("10 000-bin windows", w10, m, z)
        study("GC buckets (raw counts)", gcb, m, z)
        study("global list (raw counts)", [s.count], m, z)
        print()
    # how discrete the keys are: distinct values per segment kind
    print(f"distinct coverage values genome-wide: {len(np.unique(cov))} (two-decimal text); distinct raw counts: {len(np.unique(s.count))}; "
          f"largest share of one value: coverage {np.unique(cov, return_counts=True)[1].max() / len(cov) * 100:.2f} %, "
          f"raw counts {np.unique(s.count, return_counts=True)[1].max() / len(s.count) * 100:.2f} %")


if __name__ == "__main__":
    main()
