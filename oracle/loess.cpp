// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_stats.hpp header).
//
// CPU restatement of the LOESS GC normaliser of CanvasClean (reference @ v1.40.0):
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:61-76    Train (stable argsort of x)
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:85-175   train (fit + robustness iterations)
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:177-197  computeIntervals
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:199-251  computeCoefficients
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:271-301  updateBandwidthInterval
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:311-315  tricube
//   Src/Canvas/CanvasClean/LoessInterpolator.cs:422-444  LoessModel.Predict(IEnumerable)
//   Src/Canvas/CanvasClean/LoessGCNormalizer.cs:35-59    initialize (log transform, drop +-Inf, chrY list)
//   Src/Canvas/CanvasClean/LoessGCNormalizer.cs:61-82    Normalize
//   Src/Canvas/CanvasClean/LoessGCNormalizer.cs:84-131   findBestBandwith / objective
// Pinned by CanvasTest/TestLoessInterpolator.cs:11-81 (R loess fitted values, sum|delta| < 0.31).
#include <numeric>
#include <stdexcept>

#include "oracle.h"
#include "ref_stats.hpp"

namespace ora {

struct LoessInterval {
    double xmin, xmax;  // [xmin, xmax)
    int left, right;
};

struct LoessModel {
    std::vector<double> xs, ys;        // sorted by x
    std::vector<double> fitted;        // sorted by x (may be empty)
    std::vector<double> robust;        // sorted by x (may be empty)
    std::vector<int> ascending_order;  // argsort
    std::vector<LoessInterval> intervals;
};

static inline double tricube(double x) {
    double t = 1 - x * x * x;
    return t * t * t;
}

// LoessInterpolator.cs:271-301 — returns true when the interval moved.
static bool update_bandwidth_interval(double x, const std::vector<double>& xv, int& left, int& right) {
    bool updated = false;
    int n = (int)xv.size();
    while (right < n - 1 && x > xv[right]) { left++; right++; updated = true; }
    while (right < n - 1 && xv[right + 1] - x < x - xv[left]) { left++; right++; updated = true; }
    return updated;
}

// LoessInterpolator.cs:199-251 (+ predict :253-262)
static double fit_at(double x, const std::vector<double>& xv, const std::vector<double>& yv,
                     const std::vector<double>* rw, int left, int right) {
    int edge = (x - xv[left] > xv[right] - x) ? left : right;
    double sw = 0, sx = 0, sxx = 0, sy = 0, sxy = 0;
    double denom = std::fabs(1.0 / (xv[edge] - x));
    for (int k = left; k <= right; ++k) {
        double xk = xv[k], yk = yv[k];
        double dist = std::fabs(x - xk);
        double r = rw ? (*rw)[k] : 1.0;
        double w = tricube(dist * denom) * r;
        double xkw = xk * w;
        sw += w;
        sx += xkw;
        sxx += xk * xkw;
        sy += yk * w;
        sxy += yk * xkw;
    }
    double mx = sx / sw, my = sy / sw, mxy = sxy / sw, mxx = sxx / sw;
    double beta = (mxx == mx * mx) ? 0 : (mxy - mx * my) / (mxx - mx * mx);
    double alpha = my - beta * mx;
    double y = 0;
    y += 1.0 * alpha;  // Math.Pow(x, 0) * coefficients[0]
    y += x * beta;     // Math.Pow(x, 1) * coefficients[1]
    return y;
}

// LoessInterpolator.cs:61-175
static LoessModel loess_train(const std::vector<double>& xin, const std::vector<double>& yin,
                              double bandwidth, int robustness_iters, double x_step,
                              bool compute_fitted) {
    LoessModel m;
    int n = (int)xin.size();
    m.ascending_order.resize(n);
    std::iota(m.ascending_order.begin(), m.ascending_order.end(), 0);
    std::stable_sort(m.ascending_order.begin(), m.ascending_order.end(),
                     [&](int a, int b) { return dotnet_less<double>(xin[a], xin[b]); });
    m.xs.resize(n);
    m.ys.resize(n);
    for (int i = 0; i < n; i++) {
        m.xs[i] = xin[m.ascending_order[i]];
        m.ys[i] = yin[m.ascending_order[i]];
    }
    if (n <= 1) {
        m.fitted = {n ? m.ys[0] : 0.0};
        return m;
    }
    int bw = (int)std::ceil(bandwidth * n);
    if (bw < 2) throw std::runtime_error("loess: bandwidth too small");
    if (robustness_iters > 0) compute_fitted = true;
    if (compute_fitted) m.fitted.assign(n, 0.0);
    std::vector<double> residuals;
    if (robustness_iters > 0) {
        residuals.assign(n, 0.0);
        m.robust.assign(n, 1.0);
    }
    const std::vector<double>* rw = m.robust.empty() ? nullptr : &m.robust;
    for (int iter = 0; iter <= robustness_iters; ++iter) {
        int left = 0, right = bw - 1;
        for (int i = 0; i < n; ++i) {
            double x = m.xs[i];
            if (i > 0) update_bandwidth_interval(x, m.xs, left, right);
            if (compute_fitted) m.fitted[i] = fit_at(x, m.xs, m.ys, rw, left, right);
            if (robustness_iters > 0) residuals[i] = std::fabs(m.ys[i] - m.fitted[i]);
        }
        if (iter == robustness_iters) break;
        double med = median_d(residuals);
        if (med == 0) break;
        for (int i = 0; i < n; ++i) {
            double arg = residuals[i] / (6 * med);
            m.robust[i] = (arg >= 1) ? 0 : std::pow(1 - arg * arg, 2);
        }
    }
    // computeIntervals :177-197 — NB the closed interval carries the *previous* window.
    {
        int left = 0, right = bw - 1;
        double xmin = -std::numeric_limits<double>::infinity();
        for (double x = m.xs[0]; x <= m.xs[n - 1]; x += x_step) {
            int nl = left, nr = right;
            if (update_bandwidth_interval(x, m.xs, nl, nr)) {
                m.intervals.push_back({xmin, x, left, right});
                xmin = x;
                left = nl;
                right = nr;
            }
        }
        m.intervals.push_back({xmin, std::numeric_limits<double>::infinity(), left, right});
    }
    return m;
}

// LoessInterpolator.cs:422-444
static std::vector<double> loess_predict(const LoessModel& m, const std::vector<double>& xq) {
    std::vector<int> order(xq.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return dotnet_less<double>(xq[a], xq[b]); });
    std::vector<double> y(xq.size());
    const std::vector<double>* rw = m.robust.empty() ? nullptr : &m.robust;
    size_t ii = 0;
    for (size_t i = 0; i < xq.size(); i++) {
        double x = xq[order[i]];
        while (ii + 1 < m.intervals.size() && m.intervals[ii].xmax <= x) ii++;
        y[order[i]] = fit_at(x, m.xs, m.ys, rw, m.intervals[ii].left, m.intervals[ii].right);
    }
    return y;
}

static std::vector<double> fitted_by_gc(const std::vector<double>& gcs, const std::vector<double>& ys,
                                        double bandwidth, int min_gc, int max_gc) {
    LoessModel model = loess_train(gcs, ys, bandwidth, 0, 1.0, false);
    // Enumerable.Range(minGC, maxGC): maxGC is a COUNT (LoessGCNormalizer.cs:74,111).
    std::vector<double> xq;
    for (int i = 0; i < max_gc; i++) xq.push_back((double)(min_gc + i));
    return loess_predict(model, xq);
}

static double sd_all(const std::vector<double>& x) {
    // Utilities.StandardDeviation(double[]) — same sequential mean / (n-1) form.
    return stddev_range(x.data(), 0, (long)x.size());
}

// LoessGCNormalizer.cs:97-131
static double loess_objective(double bandwidth, const std::vector<double>& gcs,
                              const std::vector<double>& counts) {
    double median_y = median_d(counts);
    int min_gc = (int)*std::min_element(gcs.begin(), gcs.end());
    int max_gc = (int)*std::max_element(gcs.begin(), gcs.end());
    size_t n = counts.size();
    std::vector<double> normalized(n), fitted(n);
    {
        auto f = fitted_by_gc(gcs, counts, bandwidth, min_gc, max_gc);
        for (size_t i = 0; i < n; i++) {
            // reference indexes f[gc - minGC] unguarded (IndexOutOfRange when minGC == 0 and
            // gc == maxGC); clamped here, flagged in DESIGN.md.
            int idx = std::min((int)f.size() - 1, (int)gcs[i] - min_gc);
            normalized[i] = counts[i] - f[idx] + median_y;
        }
    }
    {
        auto f = fitted_by_gc(gcs, normalized, bandwidth, min_gc, max_gc);
        for (size_t i = 0; i < n; i++) {
            int idx = std::min((int)f.size() - 1, (int)gcs[i] - min_gc);
            fitted[i] = f[idx];
        }
    }
    return sd_all(fitted);
}

// LoessGCNormalizer.cs:35-82 with countTransformer = log, invCountTransformer = (float)exp
// (CanvasClean.cs:147-151, robustnessIter 0).
void loess_gc_normalize(std::vector<float>& count, const std::vector<int>& gc,
                        const std::vector<uint8_t>& is_chr_y) {
    std::vector<double> gcs, ys, gcs_noy, ys_noy;
    for (size_t i = 0; i < count.size(); i++) {
        double c = std::log((double)count[i]);
        if (std::isinf(c)) continue;
        gcs.push_back((double)gc[i]);
        ys.push_back(c);
        if (!is_chr_y[i]) {
            gcs_noy.push_back((double)gc[i]);
            ys_noy.push_back(c);
        }
    }
    if (gcs.empty()) return;
    double lo = std::max(2.0 / (double)gcs_noy.size(), 0.3);
    double hi = std::min(1.0, 0.75);
    if (hi < lo) hi = lo;
    double best = golden_section_search(
        [&](double b) { return loess_objective(b, gcs_noy, ys_noy); }, lo, hi);
    double median_y = median_d(ys);
    int min_gc = (int)*std::min_element(gcs.begin(), gcs.end());
    int max_gc = (int)*std::max_element(gcs.begin(), gcs.end());
    auto f = fitted_by_gc(gcs, ys, best, min_gc, max_gc);
    for (size_t i = 0; i < count.size(); i++) {
        int idx = std::min((int)f.size() - 1, std::max(0, gc[i] - min_gc));
        double smoothed = std::log((double)count[i]) - f[idx] + median_y;
        count[i] = (float)std::exp(smoothed);
    }
}

}  // namespace ora

using namespace ora;

// Direct entry to LoessInterpolator.Train/Predict for the known-answer test.
extern "C" int ora_loess_train(int n, const double* x, const double* y, double bandwidth,
                               int robustness_iters, double x_step, double* fitted_orig_order,
                               int n_query, const double* xq, double* yq) {
    try {
        std::vector<double> xv(x, x + n), yv(y, y + n);
        LoessModel m = loess_train(xv, yv, bandwidth, robustness_iters, x_step, true);
        // LoessModel.Fitted: OriginalOrder[ascendingOrder[i]] = i
        for (int i = 0; i < n; i++) fitted_orig_order[m.ascending_order[i]] = m.fitted[i];
        if (n_query > 0) {
            auto p = loess_predict(m, std::vector<double>(xq, xq + n_query));
            for (int i = 0; i < n_query; i++) yq[i] = p[i];
        }
    } catch (const std::exception&) {
        return -1;
    }
    return 0;
}

extern "C" double ora_golden_section_quadratic(double a, double b) {
    return golden_section_search([](double x) { return x * x; }, a, b);
}
