"""CPU prototype for the next selection kernel (point 1): order statistics of Partition's coverage on
16-bit integer hundredths instead of 64-bit double keys.

The coverage CanvasPartition reads is `hundredths / 100.0` by construction (two-decimal text, IO.cs:21).  Median: select the
integer(s), rebuild the double(s), average as SortedList<double>.Median() does.  MAD = median of |x - median| in doubles:
the integer distance |2h - 2m| orders the doubles except INSIDE a distance class, where the bins above and below the median
can give two different doubles (rounding of h/100 on either side); the class of the requested rank is resolved from the
populations of its two sides.  The script checks both against the plain double computation, bit for bit, on every
chromosome and every 10 000- / 100 000-bin window of the synthetic genome, and counts how often the two-sided classes occur.
Usage: python tools/hundredths_select_study.py > profiles/<tag>_hundredths_select_study.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from canvas_b200 import synth


def median_double(x):
    s = np.sort(x)
    n = len(s)
    return s[n // 2] if n & 1 else (s[n // 2 - 1] + s[n // 2]) / 2


def median_from_hundredths(h):
    """h: int64 hundredths.  Counting select (what two 8-bit digit passes resolve), doubles rebuilt from the integers."""
    n = len(h)
    cnt = np.bincount(h)
    cum = np.cumsum(cnt)
    hi = int(np.searchsorted(cum, n // 2 + 1))            # value at rank n/2
    if n & 1:
        return hi / 100.0, 2 * hi
    lo = int(np.searchsorted(cum, n // 2))                # value at rank n/2 - 1
    return (lo / 100.0 + hi / 100.0) / 2, lo + hi          # second value: twice the median in hundredths


def mad_from_hundredths(h, med, med2):
    """k-th smallest |h/100 - med| via the integer distance |2h - med2| and a two-sided resolution of the selected class."""
    n = len(h)
    dist = np.abs(2 * h - med2)
    cnt = np.bincount(dist)
    cum = np.cumsum(cnt)
    two_sided = 0

    def value_at(rank):
        nonlocal two_sided
        d = int(np.searchsorted(cum, rank + 1))
        before = int(cum[d - 1]) if d > 0 else 0
        r = rank - before                                   # rank inside the class
        up, dn = (med2 + d), (med2 - d)                     # twice the hundredths of the two sides
        vals = []
        if up % 2 == 0 and up // 2 < len(hist) and hist[up // 2]:
            vals.append((abs((up // 2) / 100.0 - med), int(hist[up // 2])))
        if d > 0 and dn % 2 == 0 and 0 <= dn // 2 < len(hist) and hist[dn // 2]:
            vals.append((abs((dn // 2) / 100.0 - med), int(hist[dn // 2])))
        vals.sort()
        if len(vals) == 2 and vals[0][0] != vals[1][0]:
            two_sided += 1
        return vals[0][0] if r < vals[0][1] else vals[1][0]

    hist = np.bincount(h)
    if n & 1:
        out = value_at(n // 2)
    else:
        out = (value_at(n // 2 - 1) + value_at(n // 2)) / 2
    return out, two_sided


def main():
    s = synth.make_sample(config=2)
    gmed = np.median(s.count)
    med_gc = np.array([np.median(s.count[s.gc == g]) if np.any(s.gc == g) else 1.0 for g in range(101)])
    med_gc[med_gc <= 0] = 1.0
    cnt = (gmed * s.count.astype(np.float64) / med_gc[s.gc]).astype(np.float32)
    h_all = np.rint(cnt.astype(np.float64) * 100).astype(np.int64)          # stand-in for the F2 text
    cov_all = h_all / 100.0
    segs = []
    for c in range(len(s.names)):
        idx = np.flatnonzero(s.chrom == c)
        segs.append(("chromosome", idx))
        for w in (10000, 100000):
            for i in range(0, max(0, len(idx) - w), w):
                segs.append((f"{w}-bin window", idx[i:i + w]))
    bad_med = bad_mad = two = 0
    for _, idx in segs:
        x, h = cov_all[idx], h_all[idx]
        m_ref = median_double(x)
        m, m2 = median_from_hundredths(h)
        bad_med += m != m_ref
        mad_ref = median_double(np.abs(x - m_ref))
        mad, t = mad_from_hundredths(h, m_ref, m2)
        bad_mad += mad != mad_ref
        two += t
    print(f"segments {len(segs)} (24 chromosomes + windows), keys: {h_all.max()} max hundredths = {int(h_all.max()).bit_length()} bits")
    print(f"median from integer hundredths: {bad_med} mismatches against the double computation")
    print(f"MAD from integer distances + two-sided class resolution: {bad_mad} mismatches; classes whose two sides gave different doubles: {two}")


if __name__ == "__main__":
    main()
