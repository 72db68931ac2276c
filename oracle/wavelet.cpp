// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_stats.hpp header).
//
// CPU restatement of the wavelet branch of CanvasPartition (reference @ v1.40.0):
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:19-48    GetInnerProdIter (recurrence form)
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:54-67    GetInnerProdMax (first index of max |ip|)
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:72-115   HardThresh
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:118-185  GetUnbalHaarVector / GetReconstructedVector / GetSegments
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:194-232  GetBreakpointsAfterHealingBadSplits
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:237-258  RefineSegments
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:264-379  FindBestUnbalancedHaarDecomposition
//   Src/Canvas/CanvasPartition/WaveletSegmentation.cs:385-426  HaarWavelets
//   Src/Canvas/CanvasPartition/Segmentation.cs:260-297         GetEvennessScore / reportScoresByWindow
//   Src/Canvas/CanvasPartition/Segmentation.cs:309-347         GetCoverageVariability / reportVariabilityByWindow
//   Src/Canvas/CanvasPartition/Segmentation.cs:364-429         FactorOfThreeCoverageVariabilities / GetTripletMediansAndCMADs
//   Src/Canvas/CanvasPartition/WaveletsRunner.cs:52-139        Run / LaunchWavelets (per-chromosome fan-out)
// Pinned by CanvasTest/CanvasPartition/WaveletTests.cs:9-90.
//
// Runtime behaviour restated because it is not in the tree:
//   .NET Core 2.0 Array.Sort<T>(T[], Comparison<T>) = ArraySortHelper<T> introspective sort
//   (insertion sort <= 16, median-of-three quicksort, heapsort at depth 2*floor(log2 n)); it is
//   unstable, and HardThresh's germline level weights depend on how it orders equal counts.
//   Parity for that tie order is unpinned (no reference test covers germline thresholds).
#include <atomic>
#include <cstdio>
#include <thread>

#include "oracle.h"
#include "ref_stats.hpp"

namespace ora {

// ---------------------------------------------------------------------------------------------
// .NET introsort with a Comparison<T> (restated from the published coreclr algorithm).
// ---------------------------------------------------------------------------------------------
namespace dotnet {
template <typename Cmp>
struct Introsort {
    int* k;
    Cmp cmp;
    void swap_if_greater(int a, int b) {
        if (a != b && cmp(k[a], k[b]) > 0) std::swap(k[a], k[b]);
    }
    void insertion(int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            int j = i;
            int t = k[i + 1];
            while (j >= lo && cmp(t, k[j]) < 0) {
                k[j + 1] = k[j];
                j--;
            }
            k[j + 1] = t;
        }
    }
    void down_heap(int i, int n, int lo) {
        int d = k[lo + i - 1];
        while (i <= n / 2) {
            int child = 2 * i;
            if (child < n && cmp(k[lo + child - 1], k[lo + child]) < 0) child++;
            if (!(cmp(d, k[lo + child - 1]) < 0)) break;
            k[lo + i - 1] = k[lo + child - 1];
            i = child;
        }
        k[lo + i - 1] = d;
    }
    void heapsort(int lo, int hi) {
        int n = hi - lo + 1;
        for (int i = n / 2; i >= 1; i--) down_heap(i, n, lo);
        for (int i = n; i > 1; i--) {
            std::swap(k[lo], k[lo + i - 1]);
            down_heap(1, i - 1, lo);
        }
    }
    int partition(int lo, int hi) {
        int mid = lo + (hi - lo) / 2;
        swap_if_greater(lo, mid);
        swap_if_greater(lo, hi);
        swap_if_greater(mid, hi);
        int pivot = k[mid];
        std::swap(k[mid], k[hi - 1]);
        int left = lo, right = hi - 1;
        while (left < right) {
            while (cmp(k[++left], pivot) < 0) {}
            while (cmp(pivot, k[--right]) < 0) {}
            if (left >= right) break;
            std::swap(k[left], k[right]);
        }
        std::swap(k[left], k[hi - 1]);
        return left;
    }
    void introsort(int lo, int hi, int depth) {
        while (hi > lo) {
            int size = hi - lo + 1;
            if (size <= 16) {
                if (size == 1) return;
                if (size == 2) { swap_if_greater(lo, hi); return; }
                if (size == 3) {
                    swap_if_greater(lo, hi - 1);
                    swap_if_greater(lo, hi);
                    swap_if_greater(hi - 1, hi);
                    return;
                }
                insertion(lo, hi);
                return;
            }
            if (depth == 0) { heapsort(lo, hi); return; }
            depth--;
            int p = partition(lo, hi);
            introsort(p + 1, hi, depth);
            hi = p - 1;
        }
    }
};
static int floor_log2(int n) {
    int r = 0;
    while (n >= 1) { r++; n /= 2; }
    return r;
}
template <typename Cmp>
void array_sort(int* keys, int n, Cmp cmp) {
    if (n < 2) return;
    Introsort<Cmp> s{keys, cmp};
    s.introsort(0, n - 1, 2 * floor_log2(n));
}
}  // namespace dotnet

// ---------------------------------------------------------------------------------------------
// Unbalanced Haar decomposition
// ---------------------------------------------------------------------------------------------
struct Node {
    double coef;
    int64_t start, brk, end;  // 1-based inclusive, as stored by the reference
};
typedef std::vector<std::vector<Node>> Tree;

// WaveletSegmentation.cs:19-48 — inner products with all n-1 Unbalanced Haar vectors.
static void inner_prod_iter(const double* x, long n, std::vector<double>& ip, double& mean) {
    ip.resize((size_t)(n - 1));
    std::vector<double> plus((size_t)(n - 1)), minus((size_t)(n - 1));
    plus[0] = std::sqrt(1 - 1.0 / n) * x[0];
    double sum_x = 0;
    for (long i = 1; i < n; i++) sum_x += x[i];
    mean = (x[0] + sum_x) / n;
    minus[0] = (1.0 / std::sqrt((double)(n * (n - 1)))) * sum_x;
    if (n > 2) {
        for (long m = 1; m < n - 1; m++) {
            double factor =
                std::sqrt((double)(n - m - 1) * (double)m / (double)(m + 1) / (double)(n - m));
            plus[m] = plus[m - 1] * factor + x[m] * std::sqrt(1.0 / (m + 1) - 1.0 / n);
            minus[m] = minus[m - 1] / factor -
                       x[m] / std::sqrt(((double)n * n / (double)(m + 1)) - (double)n);
        }
    }
    for (long i = 0; i < n - 1; i++) ip[i] = plus[i] - minus[i];
}

// WaveletSegmentation.cs:54-67 — 1-based index of the first maximum of |ip|.
// Enumerable.Max over doubles skips NaN unless everything is NaN; with an all-NaN vector the
// reference walks off the end and throws — we return n (caller would index out of range).
static long inner_prod_max(const std::vector<double>& ip) {
    double mx = -1;
    bool any = false;
    for (double v : ip) {
        double a = std::fabs(v);
        if (std::isnan(a)) continue;
        if (!any || a > mx) { mx = a; any = true; }
    }
    size_t idx = 0;
    for (; idx < ip.size(); idx++)
        if (std::fabs(ip[idx]) == mx) break;
    return (long)idx + 1;
}

// WaveletSegmentation.cs:264-379
static double best_unbalanced_haar(const double* x, long n, Tree& tree) {
    tree.clear();
    const double meanscale = 200.0;
    std::vector<double> ip;
    std::vector<double> sub;
    double mean;
    inner_prod_iter(x, n, ip, mean);
    long ind = inner_prod_max(ip);
    tree.push_back({Node{ip[ind - 1] / std::max(0.5, mean / meanscale), 1, ind, n}});
    size_t j = 0;
    auto bp_sum = [&](size_t lvl) {
        double s = 0;
        for (auto& nd : tree[lvl]) s += (double)(nd.end - nd.start) - 1.0;
        return s;
    };
    while (bp_sum(j) != 0) {
        std::vector<Node> next;
        for (const Node& p : tree[j]) {
            if (p.brk - p.start >= 1) {
                long skip = p.start - 1, take = p.brk - skip;
                sub.assign(x + skip, x + skip + take);  // Array.Copy into subX
                inner_prod_iter(sub.data(), take, ip, mean);
                ind = inner_prod_max(ip);
                next.push_back(Node{ip[ind - 1] / std::max(0.5, mean / meanscale), p.start,
                                    ind + p.start - 1, p.brk});
            }
            if (p.end - p.brk >= 2) {
                long skip = p.brk, take = p.end - skip;
                sub.assign(x + skip, x + skip + take);
                inner_prod_iter(sub.data(), take, ip, mean);
                ind = inner_prod_max(ip);
                next.push_back(Node{ip[ind - 1] / std::max(0.5, mean / meanscale), p.brk + 1,
                                    ind + p.brk, p.end});
            }
        }
        tree.push_back(std::move(next));
        j++;
    }
    double smooth = 0;
    for (long i = 0; i < n; i++) smooth += x[i];
    return smooth / std::sqrt((double)n);
}

// WaveletSegmentation.cs:72-115
static void hard_thresh(Tree& tree, double sigma, bool is_germline) {
    int tsize = (int)tree.size();
    std::vector<double> thresholds;
    std::vector<int> indices(tsize);
    if (is_germline) {
        std::vector<int> counts(tsize);
        for (int i = 0; i < tsize; i++) {
            counts[i] = (int)tree[i].size();
            indices[i] = i;
        }
        dotnet::array_sort(indices.data(), tsize, [&](int a, int b) {
            return counts[b] < counts[a] ? -1 : (counts[b] > counts[a] ? 1 : 0);  // counts[b].CompareTo(counts[a])
        });
        for (int x = 1; x <= tsize; x++) thresholds.push_back(((double)x * (1.0 - 0.8)) / tsize + 0.8);
    } else {
        for (int i = 0; i < tsize; i++) {
            thresholds.push_back(1.0);
            indices[i] = i;
        }
    }
    double n = (double)tree[0][0].end;
    for (int lvl = 0; lvl < tsize; lvl++)
        for (Node& nd : tree[lvl])
            if (std::fabs(nd.coef) <= 2 * sigma * (thresholds[indices[lvl]]) * std::sqrt(2 * std::log(n)))
                nd.coef = 0;
}

// WaveletSegmentation.cs:118-168
static std::vector<double> reconstruct(const Tree& tree, double smooth) {
    long n = (long)tree[0][0].end;
    std::vector<double> rec((size_t)n);
    for (long i = 0; i < n; i++) rec[i] = 1.0 / std::sqrt((double)n) * smooth;
    for (auto& lvl : tree)
        for (const Node& nd : lvl) {
            double nn = (double)(nd.end - nd.start + 1);
            double m = (double)(nd.brk - nd.start + 1);
            double v1 = std::sqrt(1 / m - 1 / nn);
            double v2 = -1.0 / std::sqrt(nn * nn / m - nn);
            for (long i = nd.start - 1; i < nd.end; i++) {
                long k = i - nd.start + 1;
                rec[i] = rec[i] + ((double)k < m ? v1 : v2) * nd.coef;
            }
        }
    return rec;
}

// WaveletSegmentation.cs:174-185
static std::vector<int> get_segments(const Tree& tree, double smooth) {
    std::vector<double> rec = reconstruct(tree, smooth);
    std::vector<int> bp{0};
    for (size_t i = 1; i < rec.size(); i++)
        if (rec[i] - rec[i - 1] != 0) bp.push_back((int)i);
    return bp;
}

// WaveletSegmentation.cs:194-232
static std::vector<int> heal_bad_splits(const std::vector<int>& prelim, const double* ratio, long N,
                                        const std::vector<double>& f3) {
    std::vector<int> bp;
    int L = (int)prelim.size();
    bp.push_back(prelim[0]);
    for (int i = 1; i < L; ++i) {
        int left_start = bp.back();
        int right_start = prelim[i];
        int right_end = (i < L - 1) ? prelim[i + 1] : (int)N;
        int left_len = right_start - left_start, right_len = right_end - right_start;
        double lm = median_range(ratio, left_start, left_start + left_len);
        double rm = median_range(ratio, right_start, right_start + right_len);
        double wm = (left_len * lm + right_len * rm) / (right_end - left_start);
        int smaller = std::min(left_len, right_len);
        int scale = std::min((int)f3.size() - 1, (int)std::ceil(std::log((double)smaller) / std::log(3.0)));
        double cutoff = f3[scale];
        if (std::fabs(lm - rm) > cutoff * 4 * std::max(wm, 50.0)) bp.push_back(prelim[i]);
    }
    return bp;
}

// WaveletSegmentation.cs:237-258
static void refine_segments(std::vector<int>& bp, const double* cov, long N) {
    const int half = 5;
    double total_median = median_range(cov, 0, N);
    for (size_t i = 1; i + 1 < bp.size(); i++) {
        int li = std::min(half, (bp[i] - bp[i - 1]) / 2);
        int ri = std::min(half, (bp[i + 1] - bp[i]) / 2);
        double best = std::fabs(median_range(cov, bp[i - 1], bp[i]) - total_median);
        int best_bp = bp[i];
        for (int j = bp[i] - li; j < bp[i] + ri; j++) {
            double d = std::fabs(median_range(cov, bp[i - 1], j) - total_median);
            if (d > best) { best = d; best_bp = j; }
        }
        bp[i] = best_bp;
    }
}

// WaveletSegmentation.cs:385-426
static std::vector<int> haar_wavelets(const double* ratio, long n, double thr_lower, double thr_upper,
                                      bool is_germline, double mad_factor, bool has_cv, double cv,
                                      const std::vector<double>& f3) {
    Tree tree;
    double smooth = best_unbalanced_haar(ratio, n, tree);
    double median = median_range(ratio, 0, n);
    double variability = has_cv ? median * cv : mad_range(ratio, 0, n);
    double threshold = mad_factor * variability;
    if (threshold < thr_lower) threshold = thr_lower;
    if (threshold > thr_upper) threshold = thr_upper;
    hard_thresh(tree, threshold, is_germline);
    std::vector<int> prelim = get_segments(tree, smooth);
    std::vector<int> bp = heal_bad_splits(prelim, ratio, n, f3);
    if (is_germline) refine_segments(bp, ratio, n);
    return bp;
}

// ---------------------------------------------------------------------------------------------
// Genome-wide scalars
// ---------------------------------------------------------------------------------------------
// Segmentation.cs:333-347
static std::vector<float> variability_by_window(int w, int n_chrom, const int64_t* off, const double* cov) {
    std::vector<float> out;
    for (int c = 0; c < n_chrom; c++) {
        const double* x = cov + off[c];
        long len = (long)(off[c + 1] - off[c]);
        for (long idx = 0; idx < len - w; idx += w) {
            double mad = mad_range(x, idx, idx + w);
            double med = median_range(x, idx, idx + w);
            out.push_back((float)(mad / med));
        }
    }
    return out;
}

// Segmentation.cs:309-327
static bool coverage_variability(int window, int n_chrom, const int64_t* off, const double* cov, double& cv) {
    long total = (long)(off[n_chrom] - off[0]);
    if (total < 10L * window) return false;
    const int window_iqr = 10000;
    if (window > window_iqr) {
        auto rv = variability_by_window(window_iqr, n_chrom, off, cov);
        auto q = quartiles_f(rv);
        if ((std::get<2>(q) - std::get<0>(q)) / std::get<1>(q) > 0.015) {
            cv = (double)std::get<0>(q);
            return true;
        }
    }
    auto rv = variability_by_window(window, n_chrom, off, cov);
    cv = median_f(rv);
    return true;
}

// Segmentation.cs:404-429
static std::vector<double> triplet_medians(const std::vector<double>& data, std::vector<double>& cmads) {
    size_t n = data.size() / 3;
    std::vector<double> med(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = i * 3 + 1;
        double a = data[j - 1], b = data[j], c = data[j + 1];
        if (a > b) std::swap(a, b);
        if (a > c) std::swap(a, c);
        if (b > c) std::swap(b, c);
        med[i] = b;
        cmads.push_back((c - a) / 2.0 / b);
    }
    return med;
}

// Segmentation.cs:364-402
static std::vector<double> factor_of_three(int n_chrom, const int64_t* off, const double* cov, int max_exp = 8) {
    std::vector<double> f3{0.0};
    std::vector<std::vector<double>> results(n_chrom);
    for (int c = 0; c < n_chrom; c++) results[c].assign(cov + off[c], cov + off[c + 1]);
    int exponent = 1;
    while (exponent <= max_exp) {
        std::vector<double> cmads;
        for (int c = 0; c < n_chrom; c++) results[c] = triplet_medians(results[c], cmads);
        if (cmads.size() < 50) {
            double last = f3.back();
            int pad = max_exp - (int)f3.size() + 1;
            for (int i = 0; i < pad; i++) f3.push_back(last);
            break;
        }
        f3.push_back(median_d(cmads));
        ++exponent;
    }
    return f3;
}

// Segmentation.cs:276-297 — evaluated the way the reference does it (one pass per integer depth).
static std::vector<double> evenness_by_window(int w, int n_chrom, const int64_t* off, const double* cov) {
    std::vector<double> out;
    for (int c = 0; c < n_chrom; c++) {
        const double* x = cov + off[c];
        long len = (long)(off[c + 1] - off[c]);
        for (long idx = 0; idx < len - w; idx += w) {
            long cnt = std::min<long>(w - 1, len - idx);  // Take(windowSize - 1)
            const double* t = x + idx;
            double sum = 0;
            for (long i = 0; i < cnt; i++) sum += t[i];
            double average = sum / (double)cnt;
            double ev = 0;
            for (int depth = 0; (double)depth <= average; depth++) {
                int ge = 0;
                for (long i = 0; i < cnt; i++) ge += (t[i] >= (double)depth);
                ev += (double)ge / sum;
            }
            if (!std::isinf(ev) && !std::isnan(ev)) out.push_back(ev);
        }
    }
    return out;
}

// Segmentation.cs:260-269.  Quartiles()/Median() of an empty list throw in the reference and the
// caller (WaveletsRunner.cs:56-66) swallows the exception: reported as ok = false.
static bool evenness_score(int window, int n_chrom, const int64_t* off, const double* cov, double& score) {
    auto iqr_scores = evenness_by_window(10000, n_chrom, off, cov);
    if (iqr_scores.empty()) return false;
    std::vector<float> f(iqr_scores.begin(), iqr_scores.end());
    if (f.size() == 1) return false;  // Quartiles indexes sorted[-1] for a single element
    auto q = quartiles_f(f);
    auto scores = evenness_by_window(window, n_chrom, off, cov);
    if (scores.empty()) return false;  // SortedList.Median() on an empty list [EXT]: treated as a throw
    double median = median_d(scores);
    score = (std::get<2>(q) - std::get<0>(q) > 0.015) ? std::get<2>(q) * 100.0 : median * 100.0;
    return true;
}

}  // namespace ora

using namespace ora;

extern "C" int ora_coverage_variability(int window, int n_chrom, const int64_t* chrom_off,
                                        const double* cov, double* cv) {
    double v = 0;
    bool ok = coverage_variability(window, n_chrom, chrom_off, cov, v);
    *cv = v;
    return ok ? 1 : 0;
}

extern "C" void ora_factor_of_three(int n_chrom, const int64_t* chrom_off, const double* cov, double* f3) {
    auto v = factor_of_three(n_chrom, chrom_off, cov);
    for (size_t i = 0; i < 9; i++) f3[i] = i < v.size() ? v[i] : v.back();
}

extern "C" int ora_evenness_score(int window, int n_chrom, const int64_t* chrom_off, const double* cov,
                                  double* score) {
    double s = 0;
    bool ok = evenness_score(window, n_chrom, chrom_off, cov, s);
    *score = s;
    return ok ? 1 : 0;
}

extern "C" int ora_haar_wavelets(int64_t n, const double* ratio, double thr_lower, double thr_upper,
                                 int is_germline, double mad_factor, int has_cv, double cv,
                                 const double* f3, int n_f3, int32_t* bp) {
    std::vector<double> f(f3, f3 + n_f3);
    auto v = haar_wavelets(ratio, (long)n, thr_lower, thr_upper, is_germline != 0, mad_factor,
                           has_cv != 0, cv, f);
    for (size_t i = 0; i < v.size(); i++) bp[i] = v[i];
    return (int)v.size();
}

extern "C" int64_t ora_uh_tree(int64_t n, const double* x, int32_t* level, int32_t* start, int32_t* brk,
                               int32_t* end, double* coef, double* smooth) {
    Tree tree;
    *smooth = best_unbalanced_haar(x, (long)n, tree);
    int64_t k = 0;
    for (size_t l = 0; l < tree.size(); l++)
        for (auto& nd : tree[l]) {
            level[k] = (int32_t)l;
            start[k] = (int32_t)nd.start;
            brk[k] = (int32_t)nd.brk;
            end[k] = (int32_t)nd.end;
            coef[k] = nd.coef;
            k++;
        }
    return k;
}

extern "C" void ora_dotnet_sort_levels(int n, const int32_t* counts, int32_t* indices) {
    for (int i = 0; i < n; i++) indices[i] = i;
    dotnet::array_sort(indices, n, [&](int a, int b) {
        return counts[b] < counts[a] ? -1 : (counts[b] > counts[a] ? 1 : 0);
    });
}

// WaveletsRunner.cs:52-139 — genome-wide scalars, then one task per chromosome.
extern "C" int ora_partition_wavelet(const ora_wavelet_opts* o, int n_chrom, const int64_t* chrom_off,
                                     const double* coverage, int32_t* n_bp, int32_t* bp,
                                     double* evenness, int* evenness_ok, double* cv, int* cv_has_value,
                                     double* factor_of_three_out) {
    double cvv = 0;
    bool has_cv = coverage_variability(o->evenness_window, n_chrom, chrom_off, coverage, cvv);
    std::vector<double> f3 = factor_of_three(n_chrom, chrom_off, coverage);
    double ev = 0;
    bool ev_ok = evenness_score(o->evenness_window, n_chrom, chrom_off, coverage, ev);
    *cv = cvv;
    *cv_has_value = has_cv;
    *evenness = ev;
    *evenness_ok = ev_ok;
    for (size_t i = 0; i < 9; i++) factor_of_three_out[i] = i < f3.size() ? f3[i] : f3.back();

    std::atomic<int> next{0};
    auto worker = [&]() {
        for (;;) {
            int c = next.fetch_add(1);
            if (c >= n_chrom) break;
            long len = (long)(chrom_off[c + 1] - chrom_off[c]);
            n_bp[c] = 0;
            if (std::max<long>(len, 1) > o->min_size) {
                auto v = haar_wavelets(coverage + chrom_off[c], len, o->thr_lower, o->thr_upper,
                                       o->is_germline != 0, o->mad_factor, has_cv, cvv, f3);
                n_bp[c] = (int32_t)v.size();
                for (size_t i = 0; i < v.size(); i++) bp[chrom_off[c] + (int64_t)i] = v[i];
            }
        }
    };
    int nt = std::max(1, o->n_threads);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}

extern "C" double ora_median_f32(int64_t n, const float* x) {
    return median_f(std::vector<float>(x, x + n));
}
extern "C" double ora_median_f64(int64_t n, const double* x) {
    return median_d(std::vector<double>(x, x + n));
}
extern "C" void ora_quartiles_f32(int64_t n, const float* x, float* q) {
    auto t = quartiles_f(std::vector<float>(x, x + n));
    q[0] = std::get<0>(t);
    q[1] = std::get<1>(t);
    q[2] = std::get<2>(t);
}
extern "C" void ora_weighted_quantiles(int64_t n, const float* v, const float* w, int n_probs,
                                       const float* probs, double* q) {
    std::vector<std::pair<float, float>> x((size_t)n);
    for (int64_t i = 0; i < n; i++) x[i] = {v[i], w[i]};
    auto r = weighted_quantiles(x, std::vector<float>(probs, probs + n_probs));
    for (int i = 0; i < n_probs; i++) q[i] = r[i];
}
