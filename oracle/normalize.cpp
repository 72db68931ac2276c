// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path.
//
// CPU restatement of CanvasNormalize's weighted-average reference and ratio steps (reference @ v1.40.0):
//   CanvasNormalize/BinCounts.cs:24-62                         OnTargetMedianBinCount: Utilities.Median of the (on-target) counts as doubles
//   CanvasNormalize/WeightedAverageReferenceGenerator.cs:43-70 weights 1 / median (0 when the median is not positive), divided by their
//                                                              sum; weighted bin count = sum of weight * count in sample order
//   CanvasNormalize/LSNormRatioCalculator.cs:28-47             library size factor = reference median / sample median when both are
//                                                              positive, else 1; bins with a reference count below 1 are skipped;
//                                                              ratio = (float division) * factor, stored as float
//   CanvasNormalize/RawRatioCalculator.cs:36-45                bins with a reference count outside [min, max] are skipped; ratio = float division
//   CanvasNormalize/CanvasNormalizeUtilities.cs:22-31          RatiosToCounts: count = (float)(ratio * (40 * ploidy / 2.0))
//   CanvasNormalize/BestLR2ReferenceGenerator.cs:32-125        the control with the smallest mean squared log ratio of the median-normalised
//                                                              on-target counts (first strict minimum; -1 when none is below +infinity)
//   CanvasNormalize/PCAReferenceGenerator.cs:37-78, :113-146    unit axes (Utilities.NormalizeBy2Norm :650-667), pairwise orthogonality
//                                                              (AreOrthogonal :685-692, tolerance 1e-4), projection of the centred
//                                                              sample (Project :700-745), F2 round trip of the temporary reference,
//                                                              raw ratios, their median, scaled reference
// Pinned in part: the projection helpers behind the PCA mode (TwoNorm, NormalizeBy2Norm, DotProduct, AreOrthogonal, Project)
// by CanvasTest/TestUtilities.cs:53-169 through tests/test_normalize_oracle.py.  Parity unpinned for the rest: the reference
// has no test of the CanvasNormalize classes themselves; restated from source only.
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "oracle.h"
#include "ref_stats.hpp"

extern "C" void ora_normalize_reference(int n_samples, int64_t n, const double* counts, const uint8_t* on_target, double* median,
                                        double* weight, double* reference) {
    for (int s = 0; s < n_samples; s++) {
        std::vector<double> on;
        for (int64_t i = 0; i < n; i++)
            if (!on_target || on_target[i]) on.push_back(counts[(size_t)s * n + i]);
        median[s] = ora::median_d(on);
        weight[s] = median[s] > 0 ? 1.0 / median[s] : 0;
    }
    double sum = 0;
    for (int s = 0; s < n_samples; s++) sum += weight[s];
    for (int s = 0; s < n_samples; s++) weight[s] /= sum;
    for (int64_t i = 0; i < n; i++) {
        double w = 0;
        for (int s = 0; s < n_samples; s++) w += weight[s] * counts[(size_t)s * n + i];
        reference[i] = w;
    }
}

extern "C" int64_t ora_normalize_ratio(int64_t n, const float* sample, const float* reference, const uint8_t* on_target, int lsnorm,
                                       double min_ref, double max_ref, const int32_t* ploidy, int32_t* kept_index, float* ratio,
                                       float* count, double* library_size_factor) {
    double factor = 1;
    if (lsnorm) {
        std::vector<double> a, b;
        for (int64_t i = 0; i < n; i++)
            if (!on_target || on_target[i]) { a.push_back((double)sample[i]); b.push_back((double)reference[i]); }
        const double sample_median = ora::median_d(a), reference_median = ora::median_d(b);
        factor = (sample_median > 0 && reference_median > 0) ? reference_median / sample_median : 1;
    }
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        if (lsnorm) {
            if (reference[i] < 1) continue;
        } else {
            if ((double)reference[i] < min_ref) continue;
            if ((double)reference[i] > max_ref) continue;
        }
        const float q = sample[i] / reference[i];
        const double r = lsnorm ? (double)q * factor : (double)q;
        const float rf = (float)r;
        const double f = 40.0 * (ploidy ? ploidy[i] : 2) / 2.0;
        kept_index[k] = (int32_t)i;
        ratio[k] = rf;
        count[k] = (float)((double)rf * f);
        k++;
    }
    if (library_size_factor) *library_size_factor = factor;
    return k;
}

extern "C" int ora_normalize_best_lr2(int n_controls, int64_t n, const double* sample, const double* controls, const uint8_t* on_target,
                                      double* mean_sq_log_ratio, int64_t* ignored) {
    auto normalised = [&](const double* x) {
        std::vector<double> on;
        for (int64_t i = 0; i < n; i++)
            if (!on_target || on_target[i]) on.push_back(x[i]);
        const double median = ora::median_d(on);
        const double weight = median > 0 ? 1.0 / median : 0;
        for (double& v : on) v = v * weight;
        return on;
    };
    const std::vector<double> tumor = normalised(sample);
    int best = -1;
    double mn = std::numeric_limits<double>::infinity();
    for (int c = 0; c < n_controls; c++) {
        const std::vector<double> normal = normalised(controls + (size_t)c * n);
        double sum = 0;
        int64_t ign = 0, used = 0;
        for (size_t i = 0; i < tumor.size(); i++) {
            if (normal[i] <= 0) { ign++; continue; }
            const double lr = std::log(tumor[i] / normal[i]);
            const double sq = lr * lr;
            if (std::isinf(sq) || std::isnan(sq)) { ign++; continue; }
            sum += sq;
            used++;
        }
        const double mean = used > 0 ? sum / used : sum;
        mean_sq_log_ratio[c] = mean;
        ignored[c] = ign;
        if (mean < mn) { mn = mean; best = c; }
    }
    return best;
}

// returns 0, or -1 when the axes are not orthogonal
extern "C" int ora_normalize_pca_reference(int64_t n, int n_axes, const float* sample, const float* mu, const double* axes,
                                           const uint8_t* on_target, double min_ref, double max_ref, float* reference,
                                           double* median_ratio) {
    std::vector<std::vector<double>> ax(n_axes, std::vector<double>(n));
    for (int k = 0; k < n_axes; k++) {
        double norm2 = 0;
        for (int64_t i = 0; i < n; i++) norm2 += axes[(size_t)k * n + i] * axes[(size_t)k * n + i];
        const double size = std::sqrt(norm2);
        for (int64_t i = 0; i < n; i++) ax[k][i] = size == 0 ? axes[(size_t)k * n + i] : axes[(size_t)k * n + i] / size;
    }
    for (int a = 0; a < n_axes; a++)
        for (int b = a + 1; b < n_axes; b++) {
            double dot = 0;
            for (int64_t i = 0; i < n; i++) dot += ax[a][i] * ax[b][i];
            if (std::fabs(dot) > 1e-4) return -1;
        }
    std::vector<double> centred(n), projected(n, 0.0), ref(n);
    for (int64_t i = 0; i < n; i++) centred[i] = (double)std::max(1.0f, sample[i]) - (double)mu[i];
    for (int k = 0; k < n_axes; k++) {
        double size = 0;
        for (int64_t i = 0; i < n; i++) size += centred[i] * ax[k][i];
        for (int64_t i = 0; i < n; i++) {
            const double t = size * ax[k][i];
            projected[i] = k == 0 ? t : projected[i] + t;
        }
    }
    std::vector<double> ratios;
    for (int64_t i = 0; i < n; i++) {
        ref[i] = std::max(1.0, (double)mu[i] + projected[i]);
        const float as_float = (float)ref[i];
        double back = 0;
        ora_f2_roundtrip(1, &as_float, &back);
        const float rf = (float)back;  // float.Parse of the two-decimal text
        if ((double)rf < min_ref) continue;
        if ((double)rf > max_ref) continue;
        const float ratio = sample[i] / rf;
        if (!on_target || on_target[i]) ratios.push_back((double)ratio);
    }
    const double med = ora::median_d(ratios);
    *median_ratio = med;
    for (int64_t i = 0; i < n; i++) reference[i] = (float)(ref[i] * med);
    return 0;
}
