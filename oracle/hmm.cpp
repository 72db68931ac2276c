// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path.
//
// CPU restatement of CanvasPartition's HMM segmentation (reference @ v1.40.0, `-m HMM` / `-m PerSampleHMM`,
// the SmallPedigree default: Canvas/CanvasRunner.cs:927):
//   CanvasPartition/HiddenMarkovModelsRunner.cs:23-109   Run: global quartiles (per-sample mode), per chromosome
//                                                         emission set-up, Viterbi, breakpoints where the state changes
//   CanvasPartition/HiddenMarkovModelsRunner.cs:111-163  InitializeNegativeBinomialEmission, RemoveOutliers
//   CanvasPartition/Distributions.cs:29-36,62-76         MultivariateNegativeBinomial
//   CanvasPartition/Distributions.cs:187-204             GetGenotypeCombinations
//   CanvasCommon/DistributionUtilities.cs:51-69          NegativeBinomialWrapper (the one MultivariateNegativeBinomial calls)
//   CanvasPartition/Distributions.cs:257-323             NegativeBinomialMixture.EstimateViterbiLikelihood
//   CanvasPartition/HMM.cs:25-52,62-130                  transition matrix, BestPathViterbi
//   CanvasCommon/Utilities.cs:290-302,340-344,361-419    Variance, Median, Quartiles
//
// Parity unpinned: the reference holds no test for the HMM.  Third-party arithmetic restated from the published
// algorithm: MathNet.Numerics 3.17.0 SpecialFunctions.GammaLn (Lanczos, g = 10.900511, 11 terms) and FactorialLn
// (log of a cached factorial below 171, GammaLn(x + 1) above); Combinatorics Permutations(WithoutRepetition) =
// the distinct arrangements of a multiset.
//
// Two observations that simplify the restatement without changing a result:
//  * the transition charge of EstimateViterbiLikelihood (:297-320) always equals transition[i][j]: `bestState` holds
//    only the values j and 2, with at least one j, and the matrix has 0.99 on the diagonal and one other value off it;
//  * the genotype arrangements of state j are all assignments of {j, 2} to the samples with at least one j (j = 2:
//    only the all-diploid one); the maximum over them does not depend on the enumeration order.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "oracle.h"
#include "ref_stats.hpp"

namespace {

// MathNet.Numerics SpecialFunctions.GammaLn
const double kGammaDk[11] = {2.48574089138753565546e-5, 1.05142378581721974210,   -3.45687097222016235469,
                             4.51227709466894823700,    -2.98285225323576655721,  1.05639711577126713077,
                             -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
                             4.63399473359905636708e-6, -2.71994908488607703910e-9};
const double kGammaR = 10.900511;
const double kLnPi = 1.1447298858494001741434273513530587116472948129153;
const double kLogTwoSqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;

double gamma_ln(double z) {
    if (z < 0.5) {
        double s = kGammaDk[0];
        for (int i = 1; i <= 10; i++) s += kGammaDk[i] / ((double)i - z);
        return kLnPi - std::log(std::sin(M_PI * z)) - std::log(s) - kLogTwoSqrtEOverPi -
               ((0.5 - z) * std::log((0.5 - z + kGammaR) / M_E));
    }
    double s = kGammaDk[0];
    for (int i = 1; i <= 10; i++) s += kGammaDk[i] / (z + (double)i - 1.0);
    return std::log(s) + kLogTwoSqrtEOverPi + ((z - 0.5) * std::log((z - 0.5 + kGammaR) / M_E));
}

double factorial_ln(int x) {
    if (x <= 1) return 0.0;
    if (x < 171) {
        double f = 1.0;
        for (int i = 2; i <= x; i++) f *= (double)i;  // the cached table is built by the same running product
        return std::log(f);
    }
    return gamma_ln((double)x + 1.0);
}

// Convert.ToInt32(double): round half to even
int to_int32(double v) { return (int)std::nearbyint(v); }

// CanvasCommon/DistributionUtilities.cs:51-69 — the wrapper MultivariateNegativeBinomial calls (Distributions.cs:33), with its
// floor of 2 on the clumping parameter (adjustClumpingParameter is false on this path); CanvasPartition's own copy without
// the floor (Distributions.cs:206-217) has no caller.  Pinned by DistributionUtilitiesTests.cs:39-49.
std::vector<double> negative_binomial_wrapper(double mean, double variance, int max_value) {
    std::vector<double> density((size_t)std::max(max_value, 0), 0.0);
    double r = std::pow(std::max(mean, 0.1), 2) / (std::max(variance, mean * 1.2) - mean);
    r = std::max(2.0, r);
    for (int x = 0; x < max_value; x++) {
        double t = std::exp(std::log(std::pow(1 + mean / r, -r)) + std::log(std::pow(mean / (mean + r), x)) + gamma_ln(r + x) -
                            factorial_ln(x) - gamma_ln(r));
        density[(size_t)x] = (std::isnan(t) || std::isinf(t)) ? 0.0 : t;
    }
    return density;
}

struct Emission {
    int n_states, n_samples, len;
    std::vector<double> p;  // [state][sample][x]
    double at(int g, int s, int x) const { return p[((size_t)g * n_samples + s) * len + x]; }
};

// Distributions.cs:257-296 without the transition term: max over the genotype arrangements of the product over
// samples, multiplied in sample order starting from 1.0
double emission_max(const Emission& e, const int* x, int j, bool use_all_states) {
    const int S = e.n_samples;
    auto factor = [&](int g, int s) {
        if (use_all_states) return e.at(g, s, x[s]);
        if (g == 0 || g == 1) return std::max(e.at(0, s, x[s]), e.at(1, s, x[s]));
        if (g == 3 || g == 4) return std::max(e.at(3, s, x[s]), e.at(4, s, x[s]));
        return e.at(g, s, x[s]);
    };
    double best = -DBL_MAX;  // Double.MinValue
    const unsigned n_assign = 1u << S;
    for (unsigned mask = 0; mask < n_assign; mask++) {  // bit s set: sample s carries genotype 2
        if (j != 2 && mask == n_assign - 1) continue;   // at least one sample keeps state j
        if (j == 2 && mask != n_assign - 1) continue;
        double l = 1.0;
        for (int s = 0; s < S; s++) l *= factor((mask >> s) & 1u ? 2 : j, s);
        if (std::isnan(l) || std::isinf(l)) l = 0;
        if (best < l) best = l;
    }
    return best;
}

struct HmmChrom {
    std::vector<int32_t> bp;
    std::vector<uint8_t> states;
};

void hmm_chromosome(const ora_hmm_opts* o, int S, int64_t n, const double* const* cov, const double* g_median,
                    const double* g_pvar, HmmChrom& out) {
    const int NS = o->n_states;
    out.bp.clear();
    out.states.assign((size_t)n, 0);
    if (n <= o->min_size) return;
    // InitializeNegativeBinomialEmission (:111-153)
    std::vector<double> haploid(S), variance(S);
    for (int s = 0; s < S; s++) {
        std::vector<double> v(cov[s], cov[s] + n);
        double median = std::max(1.0, ora::median_d(v));
        if (!o->per_sample) {
            haploid[s] = median / 2.0;
            double sum = 0;
            for (double t : v) sum += t;
            double mu = sum / (double)n, ss = 0;
            for (double t : v) { double d = t - mu; ss += d * d; }
            variance[s] = ss / (double)(n - 1);
        } else {
            haploid[s] = g_median[s] / 2.0;
            variance[s] = g_pvar[s];
        }
    }
    const double max_thr = *std::max_element(haploid.begin(), haploid.end()) * NS;
    std::vector<std::vector<double>> data(S, std::vector<double>((size_t)n));
    int max_values = INT32_MIN;
    for (int64_t t = 0; t < n; t++) {
        double mx = -HUGE_VAL;
        for (int s = 0; s < S; s++) {
            double v = cov[s][t] > max_thr ? max_thr : cov[s][t];
            data[s][(size_t)t] = v;
            mx = s == 0 ? v : std::max(mx, v);
        }
        max_values = std::max(max_values, to_int32(mx));
    }
    Emission em;
    em.n_states = NS; em.n_samples = S; em.len = max_values + 10;
    em.p.assign((size_t)NS * S * em.len, 0.0);
    for (int cn = 0; cn < NS; cn++)
        for (int s = 0; s < S; s++) {
            std::vector<double> d = negative_binomial_wrapper(std::max((double)cn, 0.1) * haploid[s], variance[s], em.len);
            std::copy(d.begin(), d.end(), em.p.begin() + ((size_t)cn * S + s) * em.len);
        }
    // HiddenMarkovModel (:25-52)
    const double self_t = 0.99;
    std::vector<double> log_t((size_t)NS * NS);
    for (int i = 0; i < NS; i++)
        for (int j = 0; j < NS; j++) log_t[(size_t)i * NS + j] = std::log(i == j ? self_t : (1.0 - self_t) / (NS - 1));
    const double log_start = std::log((double)(1.0f / NS));
    // BestPathViterbi (:62-130)
    std::vector<double> prev(NS), cur(NS), le(NS);
    std::vector<uint8_t> back((size_t)n * NS);
    std::vector<int> x(S);
    for (int64_t t = 0; t < n; t++) {
        for (int s = 0; s < S; s++) x[s] = to_int32(data[s][(size_t)t]);
        for (int j = 0; j < NS; j++) le[j] = std::log(emission_max(em, x.data(), j, o->per_sample != 0));
        if (t == 0) {
            for (int j = 0; j < NS; j++) {
                cur[j] = log_start + (le[j] + log_t[j]) - log_t[j];  // transition row 0, then "subtract it off" (:80)
                back[j] = 255;
            }
        } else {
            for (int j = 0; j < NS; j++) {
                int state = 0;
                double mx = -DBL_MAX;
                for (int i = 0; i < NS; i++) {
                    double v = prev[i] + (le[j] + log_t[(size_t)i * NS + j]);
                    if (v > mx) { state = i; mx = v; }
                }
                cur[j] = mx;
                back[(size_t)t * NS + j] = (uint8_t)state;
            }
        }
        prev.swap(cur);
    }
    int best = -1;
    double mx = -DBL_MAX;
    for (int i = 0; i < NS; i++)
        if (prev[i] > mx) { best = i; mx = prev[i]; }
    // with every final score at Double.MinValue the reference indexes bestStateSequence[..][-1] and throws
    if (best < 0) best = 0;
    for (int64_t t = n - 1; t > 0; t--) {
        out.states[(size_t)t] = (uint8_t)best;
        best = back[(size_t)t * NS + best];
    }
    out.states[0] = (uint8_t)best;
    out.bp.push_back(0);
    for (int64_t t = 1; t < n; t++)
        if (out.states[(size_t)t] != out.states[(size_t)t - 1]) out.bp.push_back((int32_t)t);
}

}  // namespace

extern "C" double ora_gamma_ln(double z) { return gamma_ln(z); }

extern "C" int ora_negative_binomial(double mean, double variance, int max_value, double* out) {
    std::vector<double> d = negative_binomial_wrapper(mean, variance, max_value);
    std::copy(d.begin(), d.end(), out);
    return (int)d.size();
}

extern "C" int ora_partition_hmm(const ora_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off,
                                 const double* coverage, int32_t* n_bp, int32_t* bp, uint8_t* states) {
    if (!o || n_samples < 1 || n_samples > 8 || n_chrom < 0 || o->n_states != 5) return -1;
    const int64_t N = chrom_off[n_chrom];
    // whole-genome median and IQR-based pseudo-variance per sample, in single precision (:38-50)
    std::vector<double> g_median(n_samples), g_pvar(n_samples);
    for (int s = 0; s < n_samples; s++) {
        std::vector<float> v((size_t)N);
        for (int64_t i = 0; i < N; i++) v[(size_t)i] = (float)coverage[(size_t)s * N + i];
        auto [q1, q2, q3] = ora::quartiles_f(v);
        g_median[s] = (double)q2;
        float iqr = q3 - q1;
        g_pvar[s] = (double)(iqr * iqr);
    }
    std::vector<HmmChrom> res((size_t)n_chrom);
    auto work = [&](int c) {
        std::vector<const double*> cov(n_samples);
        for (int s = 0; s < n_samples; s++) cov[s] = coverage + (size_t)s * N + chrom_off[c];
        hmm_chromosome(o, n_samples, chrom_off[c + 1] - chrom_off[c], cov.data(), g_median.data(), g_pvar.data(), res[(size_t)c]);
    };
    const int nt = std::max(1, std::min(o->n_threads, n_chrom));
    std::vector<std::thread> th;
    for (int w = 0; w < nt; w++)
        th.emplace_back([&, w] {
            for (int c = w; c < n_chrom; c += nt) work(c);
        });
    for (auto& t : th) t.join();
    for (int c = 0; c < n_chrom; c++) {
        n_bp[c] = (int32_t)res[(size_t)c].bp.size();
        std::copy(res[(size_t)c].bp.begin(), res[(size_t)c].bp.end(), bp + chrom_off[c]);
        if (states) std::copy(res[(size_t)c].states.begin(), res[(size_t)c].states.end(), states + chrom_off[c]);
    }
    return 0;
}
