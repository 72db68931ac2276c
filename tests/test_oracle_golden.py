"""The oracle against every known-answer vector the reference's own tests hold for the hot path
(SURVEY.md §4 / §8c).  CPU only."""
import json
import os

import numpy as np

from oracle import pyoracle as ora


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_wavelet_minimal_12_breakpoints(golden_dir):
    # CanvasTest/CanvasPartition/WaveletTests.cs:9-90
    g = _load(golden_dir, "wavelet_minimal.json")
    cov = np.array(g["coverage"], np.float64)
    off = np.array([0, len(cov)], np.int64)
    cv = ora.coverage_variability(g["call"]["cv_window"], off, cov)
    f3 = ora.factor_of_three(off, cov)
    bp = ora.haar_wavelets(cov, g["call"]["thr_lower"], g["call"]["thr_upper"],
                           g["call"]["is_germline"], g["call"]["mad_factor"], cv, f3)
    assert bp.tolist() == g["breakpoints"]


def test_loess_train_matches_r(golden_dir):
    # CanvasTest/TestLoessInterpolator.cs:11-81
    g = _load(golden_dir, "loess_train.json")
    x, y = np.array(g["x"]), np.array(g["y"])
    fitted, pred = ora.loess_train(x, y, g["bandwidth"], 0, g["x_step"], xq=x)
    d_fit = np.abs(fitted - np.array(g["fittedR"])).sum()
    d_pred = np.abs(pred - np.array(g["fittedR"])).sum()
    assert d_fit < g["bound"] and d_pred < g["bound"]
    # regression targets measured on a scratch restatement during the survey
    assert abs(d_fit - 0.30472) < 1e-4
    fitted2, _ = ora.loess_train(x, y, g["bandwidth"], 2, g["x_step"])
    d2 = np.abs(fitted2 - np.array(g["weightedFittedR"])).sum()
    assert d2 < g["bound"] and abs(d2 - 0.26636) < 1e-4
    # Predict(x[i]) one at a time == Predict(x[]) (the reference asserts this)
    singles = np.array([ora.loess_train(x, y, g["bandwidth"], 0, g["x_step"], xq=[xi])[1][0] for xi in x[:20]])
    assert np.array_equal(singles, pred[:20])


def test_golden_section_search(golden_dir):
    # CanvasTest/TestUtilities.cs:33-41
    g = _load(golden_dir, "utilities.json")["golden_section"]
    for a, b in g["intervals"]:
        assert abs(ora.golden_section_quadratic(a, b)) < g["abs_bound"]


def test_sortedlist_median_rule(golden_dir):
    # CanvasTest/TestUtilities.cs:195-206 pins SortedList<float>.Median(): mean of the middles
    g = _load(golden_dir, "utilities.json")["median_filter"]
    v = np.array(g["values"], np.float32)
    hw = g["half_window"]
    got = [ora.median_f32(v[max(0, i - hw):i + hw + 1]) for i in range(len(v))]
    assert got == g["expected"]


def test_quartiles_against_definition():
    # Utilities.cs:361-419 restated independently in numpy for every n mod 4
    rng = np.random.default_rng(1)
    for n in (4, 5, 6, 7, 8, 9, 100, 101, 102, 103):
        x = rng.normal(100, 10, n).astype(np.float32)
        s = np.sort(x)
        mid = n // 2
        if n % 2 == 0:
            q2 = (s[mid - 1] + s[mid]) / np.float32(2)
            mm = mid // 2
            if mid % 2 == 0:
                q1 = (s[mm - 1] + s[mm]) / np.float32(2)
                q3 = (s[mid + mm - 1] + s[mid + mm]) / np.float32(2)
            else:
                q1, q3 = s[mm], s[mm + mid]
        else:
            q2 = s[mid]
            if (n - 1) % 4 == 0:
                k = (n - 1) // 4
                q1 = s[k - 1] * np.float32(.25) + s[k] * np.float32(.75)
                q3 = s[3 * k] * np.float32(.75) + s[3 * k + 1] * np.float32(.25)
            else:
                k = (n - 3) // 4
                q1 = s[k] * np.float32(.75) + s[k + 1] * np.float32(.25)
                q3 = s[3 * k + 1] * np.float32(.25) + s[3 * k + 2] * np.float32(.75)
        got = ora.quartiles_f32(x)
        assert got.tolist() == [np.float32(q1), np.float32(q2), np.float32(q3)], n


def test_weighted_quantiles_last_value_rule():
    # Utilities.cs:493-515: value of the LAST element (stable ascending) with cum/total <= p
    v = [1, 2, 3, 4]
    w = [1, 1, 1, 1]
    assert ora.weighted_quantiles(v, w, [0.25, 0.5, 0.75]).tolist() == [1.0, 2.0, 3.0]
    # nothing at or below p -> stays 0
    assert ora.weighted_quantiles([5, 6], [3, 1], [0.5]).tolist() == [0.0]
    # ties keep insertion order: the heavy copy of 2 comes first and overshoots p
    assert ora.weighted_quantiles([1, 2, 2, 3], [1, 2, 0.25, 1], [0.5]).tolist() == [1.0]
    assert ora.weighted_quantiles([1, 2, 2, 3], [1, 0.25, 2, 1], [0.5]).tolist() == [2.0]


def test_dotnet_introsort_is_a_descending_sort():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 16, 17, 40, 130, 1000):
        counts = rng.integers(0, 6, n).astype(np.int32)
        idx = ora.dotnet_sort_levels(counts)
        assert sorted(idx.tolist()) == list(range(n))
        c = counts[idx]
        assert np.all(c[:-1] >= c[1:])
    # small arrays go through insertion sort, which is stable
    counts = np.array([3, 1, 3, 2, 1, 3], np.int32)
    assert ora.dotnet_sort_levels(counts).tolist() == [0, 2, 5, 3, 1, 4]


def test_median_filter_streaming_restatement(golden_dir):
    # CanvasTest/TestUtilities.cs:195-206 through the statement-for-statement restatement of the streaming filter,
    # plus its behaviour on sequences shorter than the window (fewer outputs than inputs)
    g = _load(golden_dir, "utilities.json")["median_filter"]
    assert ora.median_filter(g["values"], g["half_window"]).tolist() == g["expected"]
    assert [len(ora.median_filter(np.arange(n, dtype=np.float32), 3)) for n in range(9)] == [0, 0, 0, 0, 1, 3, 5, 7, 8]
    x = np.arange(20, dtype=np.float32)
    assert ora.repeated_median_filter(x, 0).tolist() == x.tolist()
