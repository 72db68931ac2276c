// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product
// path (canvas_b200/, libcanvasgpu.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
// baseline legs may use it.
//
// CPU restatement of the statistics helpers of Illumina/canvas (reference @ v1.40.0):
//   Src/Canvas/CanvasCommon/Utilities.cs:241-257   StandardDeviation(double[], start, end)
//   Src/Canvas/CanvasCommon/Utilities.cs:340-344   Median(IEnumerable<double>)
//   Src/Canvas/CanvasCommon/Utilities.cs:361-419   Quartiles(List<float>)
//   Src/Canvas/CanvasCommon/Utilities.cs:428-463   Median(x,start,end) / Mad(x,start,end)
//   Src/Canvas/CanvasCommon/Utilities.cs:470-474   Median(IEnumerable<float>)
//   Src/Canvas/CanvasCommon/Utilities.cs:493-520   WeightedQuantiles / WeightedMedian
//   Src/Canvas/CanvasCommon/Utilities.cs:1014-1044 GoldenSectionSearch
// Third-party arithmetic restated here because its source is not in the tree:
//   Illumina.Common 6.2.0.419 SortedList<T>.Median(): mean of the two middle elements for even n,
//   computed in T — pinned by CanvasTest/TestUtilities.cs:195-206 (TestMedianFilter).
//   .NET ordering of floating point: NaN compares below every number (Double.CompareTo).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <tuple>
#include <utility>
#include <vector>

namespace ora {

// .NET Double.CompareTo / Single.CompareTo ordering: NaN first, then numeric order.
template <typename T>
inline bool dotnet_less(T a, T b) {
    if (std::isnan(a)) return !std::isnan(b);
    if (std::isnan(b)) return false;
    return a < b;
}

template <typename T>
inline void dotnet_sort(std::vector<T>& v) {
    std::sort(v.begin(), v.end(), dotnet_less<T>);
}

// SortedList<T>.Median() [Illumina.Common, restated]: sorts a copy; odd n -> middle, even n -> mean
// of the middles in T.  Empty input is not reachable on the hot path; we return 0.
template <typename T>
inline T sorted_median(std::vector<T> v) {
    if (v.empty()) return T(0);
    dotnet_sort(v);
    size_t n = v.size();
    if (n & 1) return v[n / 2];
    return (v[n / 2 - 1] + v[n / 2]) / T(2);
}

// Utilities.cs:470-474 — float list, returned widened to double.
inline double median_f(const std::vector<float>& x) { return (double)sorted_median<float>(x); }
// Utilities.cs:340-344
inline double median_d(const std::vector<double>& x) { return sorted_median<double>(x); }

// Utilities.cs:428-442 — Median(x, start, end) over x[start, end)
inline double median_range(const double* x, long start, long end) {
    return sorted_median<double>(std::vector<double>(x + start, x + end));
}

// Utilities.cs:451-463 — Mad(x, start, end)
inline double mad_range(const double* x, long start, long end) {
    double med = median_range(x, start, end);
    std::vector<double> diffs((size_t)(end - start));
    for (long i = start; i < end; i++) diffs[(size_t)(i - start)] = std::fabs(x[i] - med);
    return sorted_median<double>(std::move(diffs));
}

// Utilities.cs:241-257 — sample SD over x[start,end) with sequential mean (Utilities.Mean).
inline double stddev_range(const double* x, long start, long end) {
    double s = 0;
    for (long i = start; i < end; i++) s += x[i];
    double mu = s / (double)(end - start);
    double sum = 0;
    for (long i = start; i < end; i++) {
        double d = x[i] - mu;
        sum += d * d;
    }
    return std::sqrt(sum / (double)(end - start - 1));
}

// Utilities.cs:361-419 — Quartiles in single precision.
inline std::tuple<float, float, float> quartiles_f(const std::vector<float>& x) {
    std::vector<float> s(x);
    dotnet_sort(s);
    int n = (int)s.size();
    int mid = n / 2;
    float q1 = 0, q2 = 0, q3 = 0;
    if (n == 0) return {q1, q2, q3};
    if (n % 2 == 0) {
        q2 = (s[mid - 1] + s[mid]) / 2;
        int mm = mid / 2;
        if (mid % 2 == 0) {
            q1 = (s[mm - 1] + s[mm]) / 2;
            q3 = (s[mid + mm - 1] + s[mid + mm]) / 2;
        } else {
            q1 = s[mm];
            q3 = s[mm + mid];
        }
    } else {
        q2 = s[mid];
        if ((n - 1) % 4 == 0) {
            int k = (n - 1) / 4;
            // n == 1 would index s[-1] in the reference (throws); unreachable on the hot path.
            if (k >= 1) {
                q1 = (s[k - 1] * 0.25f) + (s[k] * 0.75f);
                q3 = (s[3 * k] * 0.75f) + (s[3 * k + 1] * 0.25f);
            }
        } else if ((n - 3) % 4 == 0) {
            int k = (n - 3) / 4;
            q1 = (s[k] * 0.75f) + (s[k + 1] * 0.25f);
            q3 = (s[3 * k + 1] * 0.25f) + (s[3 * k + 2] * 0.75f);
        }
    }
    return {q1, q2, q3};
}

// Utilities.cs:493-515 — weighted quantiles: value of the last element (stable ascending order by
// value) whose cumulative weight / total weight is <= p.  LINQ Sum over float accumulates in double
// and returns float; the result is then widened again.
inline std::vector<double> weighted_quantiles(const std::vector<std::pair<float, float>>& x,
                                              const std::vector<float>& probs) {
    double acc = 0;
    for (auto& t : x) acc += (double)t.second;
    double total = (double)(float)acc;
    std::vector<size_t> order(x.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](size_t a, size_t b) { return dotnet_less<float>(x[a].first, x[b].first); });
    std::vector<double> q(probs.size(), 0.0);
    double cw = 0;
    for (size_t oi : order) {
        cw += (double)x[oi].second;
        double cp = cw / total;
        for (size_t i = 0; i < probs.size(); i++)
            if (cp <= (double)probs[i]) q[i] = (double)x[oi].first;
    }
    return q;
}

inline double weighted_median(const std::vector<std::pair<float, float>>& x) {
    return weighted_quantiles(x, {0.5f})[0];
}

// Utilities.cs:1014-1044
inline double golden_section_search(const std::function<double(double)>& f, double a, double b,
                                    double tol = 1e-5) {
    const double g = 0.618034;
    double c = b - g * (b - a);
    double d = a + g * (b - a);
    double fc = f(c), fd = f(d);
    while (std::fabs(d - c) > tol) {
        if (fc < fd) {
            b = d; d = c; fd = fc;
            c = b - g * (b - a);
            fc = f(c);
        } else {
            a = c; c = d; fc = fd;
            d = a + g * (b - a);
            fd = f(d);
        }
    }
    return (b + a) / 2;
}

}  // namespace ora
