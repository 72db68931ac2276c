"""Host-side text semantics of the intermediate files (reference CanvasCommon/IO.cs:15-52).

`f2_roundtrip` reproduces what a count goes through between CanvasClean and CanvasPartition:
`string.Format("{3:F2}", float)` on write (IO.cs:21) and `Convert.ToDouble(string)` on read
(CanvasSegment.cs:1147).  .NET Core 2.0 formats a float from its first 7 significant decimal digits
and then rounds the digit string half-up to two decimals; the parse is correctly rounded.
"""
import numpy as np


def f2_hundredths(count):
    """Integer hundredths that `{0:F2}` prints for each float32 count (sign kept separately)."""
    v = np.asarray(count, np.float32)
    x = np.abs(v.astype(np.float64))
    h = np.zeros(x.shape, np.int64)
    ok = np.isfinite(x) & (x > 0)
    xs = x[ok]
    e = np.floor(np.log10(xs)).astype(np.int64)
    # fix the rare off-by-one of log10 at powers of ten (products below are exact)
    e = np.where(xs < 10.0 ** e.astype(np.float64), e - 1, e)
    e = np.where(xs >= 10.0 ** (e + 1).astype(np.float64), e + 1, e)
    e = np.clip(e, -30, 18)
    scaled = np.where(e <= 6, xs * 10.0 ** np.clip(6 - e, 0, 36).astype(np.float64),
                      xs / 10.0 ** np.clip(e - 6, 0, 18).astype(np.float64))
    d7 = np.rint(scaled)
    over = d7 >= 1e7
    d7 = np.where(over, d7 / 10.0, d7)
    e = np.where(over, e + 1, e)
    q = d7.astype(np.int64)
    k = 4 - e
    hh = np.zeros(len(xs), np.int64)
    big = k <= 0
    hh[big] = q[big] * (10 ** np.clip(-k[big], 0, 12))
    small = (k >= 1) & (k <= 7)
    p = 10 ** np.clip(k[small], 0, 7)
    hs = q[small] // p
    hs += ((q[small] % p) * 2 >= p).astype(np.int64)
    hh[small] = hs
    h[ok] = hh
    return h


def f2_roundtrip(count):
    """float32 counts -> the doubles CanvasPartition parses from the .cleaned file."""
    v = np.asarray(count, np.float32)
    h = f2_hundredths(v)
    out = h.astype(np.float64) / 100.0
    out = np.where(v < 0, -out, out)
    bad = ~np.isfinite(v)
    if bad.any():
        out = np.where(bad, v.astype(np.float64), out)
    return out


def f2_text(count):
    """The strings `{0:F2}` prints (used by the .cleaned writer)."""
    v = np.asarray(count, np.float32)
    h = f2_hundredths(v)
    neg = (v < 0) & (h != 0)
    return [("-" if n else "") + f"{a // 100}.{a % 100:02d}" for a, n in zip(h.tolist(), neg.tolist())]


def float_default_text(count):
    """float.ToString() of .NET Core 2.0 — general format with 7 significant digits, scientific from 1E+07 up and
    below 1E-04 — as NormalizeCanvasClean interpolates the merged counts (CanvasRunner.cs:895-897)."""
    out = []
    for v in np.asarray(count, np.float32).tolist():
        if v != v:
            out.append("NaN")
        elif v in (float("inf"), float("-inf")):
            out.append("Infinity" if v > 0 else "-Infinity")
        else:
            t = "%.7g" % v
            if "e" in t:
                m, e = t.split("e")
                t = f"{m}E{'-' if e[0] == '-' else '+'}{abs(int(e)):02d}"
            out.append(t)
    return out


def float_default_roundtrip(count):
    """float32 counts -> the doubles CanvasPartition parses from the merged (4-column) .cleaned file."""
    v = np.asarray(count, np.float32)
    x = np.abs(v.astype(np.float64))
    out = np.zeros(x.shape, np.float64)
    ok = np.isfinite(x) & (x > 0)
    xs = x[ok]
    e = np.floor(np.log10(xs)).astype(np.int64)
    e = np.where(xs < 10.0 ** e.astype(np.float64), e - 1, e)
    e = np.where(xs >= 10.0 ** (e + 1).astype(np.float64), e + 1, e)
    k = 6 - e  # decimal shift that leaves seven significant digits before the point
    scaled = np.where(k >= 0, xs * 10.0 ** np.clip(k, 0, 22).astype(np.float64), xs / 10.0 ** np.clip(-k, 0, 22).astype(np.float64))
    d7 = np.rint(scaled)
    out[ok] = np.where(k >= 0, d7 / 10.0 ** np.clip(k, 0, 22).astype(np.float64), d7 * 10.0 ** np.clip(-k, 0, 22).astype(np.float64))
    out = np.where(v < 0, -out, out)
    bad = ~np.isfinite(v)
    return np.where(bad, v.astype(np.float64), out)
