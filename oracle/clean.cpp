// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_stats.hpp header).
//
// CPU restatement of CanvasClean (reference Src/Canvas/CanvasClean/CanvasClean.cs @ v1.40.0):
//   :34-97    NormalizeVarianceByGC        :107-132  GetWeightedCounts
//   :163-196  NormalizeByGC (MedianByGC)   :207-237  RemoveBinsWithExtremeGC
//   :243-258  GetLocalStandardDeviationAverage        :268-298 GetLocalStandardDeviation
//   :308-322  RemoveBinsWithExtremeLocalSD :328-355  RemoveBigBins
//   :363-381  SignificantlyDifferent       :387-413  RemoveOutliers
//   :474-530  Main (numeric block between ReadFromTextFile and WriteToTextFile)
//   Src/Canvas/CanvasClean/EnrichmentUtilities.cs:65-84 GetCountsByGC
// Structure follows the reference on purpose (array-of-records, list rebuilt by every filter,
// sort-based order statistics): it is also the CPU baseline of bench.py.
//
// Un-vendored dependency: Isas.SequencingFiles GenomeMetadata.SequenceMetadata.IsAutosome(string) —
// the caller passes a per-chromosome flag table instead (parity unpinned for the name rule itself).
#include <cstdio>
#include <cstring>

#include "oracle.h"
#include "ref_stats.hpp"

namespace ora {

struct Bin {
    int32_t chrom;
    int32_t start, stop;
    int32_t gc;
    float count;
    double dev;    // SampleGenomicBin.CountDeviation (GenomicBin.cs:49, initialised to -1 :84)
    int32_t orig;  // index in the caller's arrays
};

static const int kGcBins = 101;           // EnrichmentUtilities.numberOfGCbins
static const int kDefaultMinBinsPerGC = 100;  // CanvasClean.cs:14

// CanvasClean.cs:328-355
static std::vector<Bin> remove_big_bins(const std::vector<Bin>& bins) {
    std::vector<int> sizes;
    sizes.reserve(bins.size());
    for (auto& b : bins) sizes.push_back(b.stop - b.start);
    std::sort(sizes.begin(), sizes.end());
    int index = (int)(0.98 * (double)bins.size());
    if (index >= (int)sizes.size()) return bins;  // "Too few bins to do outlier removal"
    int thresh = sizes[index];
    std::vector<Bin> out;
    out.reserve(bins.size());
    for (auto& b : bins)
        if (b.stop - b.start <= thresh) out.push_back(b);
    return out;
}

// CanvasClean.cs:363-381
static bool significantly_different(float a, float b) {
    double mu = ((double)a + (double)b) / 2;
    if (a + b == 0) return false;
    double da = (double)a - mu, db = (double)b - mu;
    double chi2 = (da * da + db * db) / mu;
    return chi2 > 6.635;
}

// CanvasClean.cs:387-413
static std::vector<Bin> remove_outliers(const std::vector<Bin>& bins) {
    std::vector<Bin> out;
    out.reserve(bins.size());
    long n = (long)bins.size();
    for (long i = 0; i < n; i++) {
        bool has_prev = i > 0, has_next = i < n - 1;
        int c = bins[i].chrom;
        bool prev_same = has_prev && bins[i - 1].chrom == c;
        bool next_same = has_next && bins[i + 1].chrom == c;
        if ((has_prev && !prev_same) && (has_next && !next_same)) continue;
        if ((prev_same && !significantly_different(bins[i].count, bins[i - 1].count)) ||
            (next_same && !significantly_different(bins[i].count, bins[i + 1].count)) ||
            (!has_prev && !has_next))
            out.push_back(bins[i]);
    }
    return out;
}

// CanvasClean.cs:243-258
static double local_sd_average(const std::vector<double>& sds, const std::vector<int>& chrom) {
    std::vector<double> mads;
    long start = 0;
    for (long i = 0; i < (long)sds.size(); i++) {
        if (chrom[i] != chrom[start]) {
            mads.push_back(mad_range(sds.data(), start, i));
            start = i;
        }
    }
    // With an empty list the reference would throw inside Mad(); unreachable (needs >= 50000 bins).
    if (sds.empty()) return std::numeric_limits<double>::quiet_NaN();
    mads.push_back(mad_range(sds.data(), start, (long)sds.size()));
    double s = 0;
    for (double m : mads) s += m;  // Enumerable.Average: sequential double sum / count
    return s / (double)mads.size();
}

// CanvasClean.cs:268-298
static double local_standard_deviation(std::vector<Bin>& bins) {
    long n = (long)bins.size();
    std::vector<double> diffs((size_t)std::max<long>(n - 1, 0));
    for (long i = 0; i + 1 < n; i++) diffs[i] = (double)(bins[i + 1].count - bins[i].count);
    std::vector<double> sds;
    std::vector<int> chrom;
    const long w = 20;
    for (long we = w, ws = 0; we < (long)diffs.size(); ws += w, we += w) {
        double sd = stddev_range(diffs.data(), ws, we);
        sds.push_back(sd);
        chrom.push_back(bins[ws].chrom);
        for (long i = ws; i < we; i++) bins[i].dev = sd;
    }
    return local_sd_average(sds, chrom);
}

// EnrichmentUtilities.cs:65-84
static void counts_by_gc(const std::vector<Bin>& bins, const uint8_t* is_auto,
                         std::vector<std::vector<float>>& by_gc, std::vector<float>& all) {
    by_gc.assign(kGcBins, {});
    all.clear();
    all.reserve(bins.size());
    for (auto& b : bins) {
        if (!is_auto[b.chrom]) continue;
        by_gc[b.gc].push_back(b.count);
        all.push_back(b.count);
    }
}

// CanvasClean.cs:107-132
static std::vector<std::pair<float, float>> weighted_counts(
    const std::vector<std::vector<float>>& by_gc, int gc) {
    std::vector<std::pair<float, float>> wc;
    int radius = 0;
    float weight = 1;
    while ((int)wc.size() < kDefaultMinBinsPerGC) {
        int hi = gc + radius, lo = gc - radius;
        if (hi >= (int)by_gc.size() && lo < 0) break;
        if (hi < (int)by_gc.size())
            for (float c : by_gc[hi]) wc.emplace_back(c, weight);
        if (lo != hi && lo >= 0)
            for (float c : by_gc[lo]) wc.emplace_back(c, weight);
        radius++;
        weight /= 2;
    }
    return wc;
}

// CanvasClean.cs:207-237
static std::vector<Bin> remove_extreme_gc(const std::vector<Bin>& bins, int threshold,
                                          int min_bins_weighted, const uint8_t* is_auto) {
    int counts[kGcBins] = {0};
    double total = 0;
    for (auto& b : bins) {
        if (!is_auto[b.chrom]) continue;
        counts[b.gc]++;
        total++;
    }
    int avg = std::max(min_bins_weighted, (int)(total / kGcBins));
    threshold = std::min(threshold, avg);
    std::vector<Bin> out;
    out.reserve(bins.size());
    for (auto& b : bins)
        if (counts[b.gc] >= threshold) out.push_back(b);
    return out;
}

// CanvasClean.cs:163-196
static void normalize_by_gc(std::vector<Bin>& bins, const uint8_t* is_auto) {
    std::vector<std::vector<float>> by_gc;
    std::vector<float> all;
    counts_by_gc(bins, is_auto, by_gc, all);
    double global_median = median_f(all);
    double med[kGcBins];
    for (int g = 0; g < kGcBins; g++) {
        if ((int)by_gc[g].size() >= kDefaultMinBinsPerGC)
            med[g] = median_f(by_gc[g]);
        else
            med[g] = weighted_median(weighted_counts(by_gc, g));
    }
    for (auto& b : bins) {
        double m = med[b.gc];
        if (m > 0) b.count = (float)(global_median * (double)b.count / m);
    }
}

// CanvasClean.cs:34-97
static bool normalize_variance_by_gc(std::vector<Bin>& bins, const uint8_t* is_auto) {
    std::vector<std::vector<float>> by_gc;
    std::vector<float> all;
    counts_by_gc(bins, is_auto, by_gc, all);
    auto gq = quartiles_f(all);
    float local_iqr[kGcBins], local_q2[kGcBins];
    for (int g = 0; g < kGcBins; g++) {
        if (by_gc[g].empty()) {
            local_iqr[g] = -1.f;
            local_q2[g] = -1.f;
        } else if ((int)by_gc[g].size() >= kDefaultMinBinsPerGC) {
            auto q = quartiles_f(by_gc[g]);
            local_q2[g] = std::get<1>(q);
            local_iqr[g] = std::get<2>(q) - std::get<0>(q);
        } else {
            auto q = weighted_quantiles(weighted_counts(by_gc, g), {0.25f, 0.5f, 0.75f});
            local_q2[g] = (float)q[1];
            local_iqr[g] = (float)(q[2] - q[0]);
        }
    }
    float global_iqr = std::get<2>(gq) - std::get<0>(gq);
    int significant = 0;
    for (int g = 10; g < 90; g++)
        if (global_iqr * 2.f < local_iqr[g]) significant++;
    if (significant <= 0) return false;
    for (auto& b : bins) {
        float scaled = local_iqr[b.gc] * 0.8f;
        if (global_iqr >= scaled) continue;
        float ratio = scaled / global_iqr;
        float m = local_q2[b.gc];
        b.count = m + (b.count - m) / ratio;
    }
    return true;
}

// LOESS mode lives in loess.cpp (LoessGCNormalizer.cs).
void loess_gc_normalize(std::vector<float>& count, const std::vector<int>& gc,
                        const std::vector<uint8_t>& is_chr_y);

// CanvasClean.cs:308-322
static std::vector<Bin> remove_extreme_local_sd(const std::vector<Bin>& bins, double avg,
                                                double threshold) {
    std::vector<Bin> out;
    out.reserve(bins.size());
    for (auto& b : bins) {
        if (b.dev > threshold * 2.0 && avg > 5.0) continue;
        out.push_back(b);
    }
    return out;
}

}  // namespace ora

using namespace ora;

// float.ToString("F2") then Convert.ToDouble (IO.cs:21 -> CanvasSegment.cs:1147) as .NET Core 2.0 does
// it: 7 significant digits first (FLOAT_PRECISION; number.cpp DoubleToNumber), then the digit string
// is rounded half-up to 2 decimals (RoundNumber).  Implemented on decimal strings, independently of
// the arithmetic version in the product.  The runtime's formatter is not in the tree: parity unpinned.
extern "C" void ora_f2_roundtrip(int64_t n, const float* in, double* out) {
    for (int64_t i = 0; i < n; i++) {
        float v = in[i];
        if (std::isnan(v) || std::isinf(v)) { out[i] = (double)v; continue; }
        char buf[64];
        snprintf(buf, sizeof buf, "%.6e", (double)std::fabs(v));  // d.dddddde[+-]xx
        int digits[7];
        digits[0] = buf[0] - '0';
        for (int k = 0; k < 6; k++) digits[k + 1] = buf[2 + k] - '0';
        int exp10 = atoi(buf + 9);
        // keep pos = exp10 + 3 digits (integer digits + 2 decimals)
        int pos = exp10 + 3;
        long long kept = 0;
        if (pos >= 7) {
            for (int k = 0; k < 7; k++) kept = kept * 10 + digits[k];
            for (int k = 7; k < pos && k < 18; k++) kept *= 10;
        } else if (pos >= 0) {
            for (int k = 0; k < pos; k++) kept = kept * 10 + digits[k];
            if (digits[pos] >= 5) kept += 1;
        } else {
            kept = 0;
        }
        char txt[64];
        snprintf(txt, sizeof txt, "%s%lld.%02lld", (v < 0 && kept != 0) ? "-" : "", kept / 100, kept % 100);
        out[i] = strtod(txt, nullptr);
    }
}

// CanvasClean.cs:474-530 — everything between ReadFromTextFile and WriteToTextFile.
extern "C" int ora_clean(const ora_clean_opts* o, int64_t n, const uint8_t* chrom,
                         const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
                         const int32_t* start, const int32_t* stop, const float* count,
                         const uint8_t* gc, int64_t* n_out, int32_t* kept_index, float* count_out,
                         double* local_sd_out, int* gc_norm_skipped) {
    (void)n_chrom;
    std::vector<Bin> bins((size_t)n);
    for (int64_t i = 0; i < n; i++)
        bins[i] = Bin{chrom[i], start[i], stop[i], gc[i], count[i], -1.0, (int32_t)i};
    if (o->size_filter) bins = remove_big_bins(bins);
    if (o->outlier_filter) bins = remove_outliers(bins);
    bool metric = o->want_local_sd != 0;
    if (metric && bins.size() < 50000) metric = false;  // :483-486
    double local_sd = -1.0;
    if (metric) local_sd = local_standard_deviation(bins);
    *gc_norm_skipped = 0;
    if (o->gc_norm) {
        std::vector<Bin> stripped =
            o->gc_mode == 0 ? remove_extreme_gc(bins, kDefaultMinBinsPerGC, o->min_bins_per_gc,
                                                chrom_is_autosome)
                            : bins;
        if (stripped.empty()) {
            *gc_norm_skipped = 1;  // :502-505
        } else {
            bins = std::move(stripped);
            auto normalize = [&]() {
                if (o->gc_mode == 0) {
                    normalize_by_gc(bins, chrom_is_autosome);
                } else {
                    std::vector<float> c(bins.size());
                    std::vector<int> g(bins.size());
                    std::vector<uint8_t> y(bins.size());
                    for (size_t i = 0; i < bins.size(); i++) {
                        c[i] = bins[i].count;
                        g[i] = bins[i].gc;
                        y[i] = chrom_is_chrY[bins[i].chrom];
                    }
                    loess_gc_normalize(c, g, y);
                    for (size_t i = 0; i < bins.size(); i++) bins[i].count = c[i];
                }
            };
            normalize();
            if (metric && bins.size() > 500000) {  // :512
                if (normalize_variance_by_gc(bins, chrom_is_autosome)) normalize();
            }
        }
    }
    if (metric) bins = remove_extreme_local_sd(bins, local_sd, 20);
    *n_out = (int64_t)bins.size();
    for (size_t i = 0; i < bins.size(); i++) {
        kept_index[i] = bins[i].orig;
        count_out[i] = bins[i].count;
    }
    *local_sd_out = local_sd;
    return 0;
}
