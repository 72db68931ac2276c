/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement of the Illumina/canvas hot path (reference @ v1.40.0,
 * CanvasClean -> CanvasPartition).  It exists to check libcanvasgpu.so and to be timed as the CPU
 * baseline.  Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs)
 * may load it.  The product path must never call into this library.
 *
 * Parity status: pinned by the reference's own known-answer tests
 *   CanvasTest/CanvasPartition/WaveletTests.cs:9-90      (12 breakpoints)
 *   CanvasTest/TestLoessInterpolator.cs:11-81            (R loess fitted values)
 *   CanvasTest/TestUtilities.cs:33-41,195-206            (golden section, SortedList median rule)
 *   CanvasTest/CanvasPartition/SegmentationResultsProcessorTests.cs:10-95
 * Everything in CanvasClean except LOESS has no reference test: restated from source only.
 */
#ifndef CANVAS_ORACLE_H
#define CANVAS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int size_filter;     /* -s */
    int outlier_filter;  /* -r */
    int gc_norm;         /* -g */
    int gc_mode;         /* -m : 0 MedianByGC, 1 LOESS */
    int want_local_sd;   /* --local-sd-metric-file given */
    int min_bins_per_gc; /* -w, default 100 */
} ora_clean_opts;

int ora_clean(const ora_clean_opts* o, int64_t n, const uint8_t* chrom,
              const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
              const int32_t* start, const int32_t* stop, const float* count, const uint8_t* gc,
              int64_t* n_out, int32_t* kept_index, float* count_out, double* local_sd,
              int* gc_norm_skipped);

/* Circular binary segmentation (oracle/cbs.cpp). */
typedef struct {
    double alpha;      /* CanvasPartitionParameters.CBSalpha, 0.01 */
    uint32_t n_perm;   /* 10000 */
    int hybrid;        /* pMethod == "hybrid" */
    int min_width;     /* 2 */
    int k_max;         /* 25 */
    uint32_t n_min;    /* 200 */
    int undo;          /* 0 none, 1 prune, 2 sdundo */
    uint32_t seed;     /* seed of the per-chromosome seed generator; 0 in the reference */
    double trim;       /* 0.025 */
    double undo_sd;    /* 3 */
    double undo_prune; /* 0.05 */
} ora_cbs_opts;
int64_t ora_cbs_boundary(uint32_t n_perm, double alpha, double eta, uint32_t* out, int64_t cap);
double ora_cbs_tailp(double b, double delta, int m);
double ora_cbs_inflation_factor(double trim);
double ora_cbs_trimmed_variance(const double* x, int64_t n, double trim);
void ora_mt19937(uint32_t seed, int64_t n, uint32_t* out);
double ora_cbs_tmaxo(const double* x, int n, int al0, int* seg);
double ora_cbs_htmaxp(const double* px, int n, int k, double tss, int al0);
/* ChangePointsPrune on one chromosome's segment lengths (n_seg >= 2); returns the new segment count. */
int ora_cbs_prune(const double* g, int n, const int32_t* len, int n_seg, double cutoff, int32_t* len_out);
int ora_partition_cbs(const ora_cbs_opts* o, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom, const int64_t* chrom_off,
                      const double* coverage, int32_t* n_seg, int32_t* seg_len, double* seg_mean, int32_t* seg_first,
                      int32_t* seg_last, int64_t* stats, int n_threads);

/* HMM segmentation (oracle/hmm.cpp): `-m HMM` (per_sample = 0, all samples jointly) and `-m PerSampleHMM`
 * (per_sample = 1, called with one sample).  coverage is [n_samples][N]; breakpoints per chromosome at
 * bp[chrom_off[c] ...] as for the wavelets; states (optional, [N]) the Viterbi path. */
typedef struct {
    int n_states;   /* 5 */
    int per_sample; /* isPerSample: whole-genome median / pseudo-variance, all states distinct */
    int min_size;   /* 10 */
    int n_threads;
} ora_hmm_opts;
int ora_partition_hmm(const ora_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off, const double* coverage,
                      int32_t* n_bp, int32_t* bp, uint8_t* states);
double ora_gamma_ln(double z);
int ora_negative_binomial(double mean, double variance, int max_value, double* out);

/* CanvasNormalize (oracle/normalize.cpp): weighted-average reference of the controls; ratio + RatiosToCounts (returns kept bins). */
void ora_normalize_reference(int n_samples, int64_t n, const double* counts, const uint8_t* on_target, double* median, double* weight,
                             double* reference);
int64_t ora_normalize_ratio(int64_t n, const float* sample, const float* reference, const uint8_t* on_target, int lsnorm, double min_ref,
                            double max_ref, const int32_t* ploidy, int32_t* kept_index, float* ratio, float* count,
                            double* library_size_factor);
/* BestLR2ReferenceGenerator: returns the index of the best control (-1: none). */
int ora_normalize_best_lr2(int n_controls, int64_t n, const double* sample, const double* controls, const uint8_t* on_target,
                           double* mean_sq_log_ratio, int64_t* ignored);
/* PCAReferenceGenerator.Run; -1 when the axes are not orthogonal. */
int ora_normalize_pca_reference(int64_t n, int n_axes, const float* sample, const float* mu, const double* axes, const uint8_t* on_target,
                                double min_ref, double max_ref, float* reference, double* median_ratio);
/* CanvasBin pre-binning passes on one chromosome (possible = one byte per position); arrays are modified in place. */
void ora_bin_screen(int64_t len, uint8_t* hits, uint8_t* possible, int64_t n_filter, const int32_t* filter_start,
                    const int32_t* filter_stop, int64_t* n_observed, int64_t* n_possible);
/* GCContentWeighted tables of one chromosome: read GC per position; expected / observed [101] are accumulated (+=). */
void ora_bin_read_gc(int64_t len, const char* bases, const int16_t* frag_len, int mean_frag, const uint8_t* hits, uint8_t* read_gc,
                     int64_t* expected, int64_t* observed);
/* CanvasSmooth (oracle/smooth.cpp): Utilities.MedianFilter and the repeated filter; return the output length (<= n). */
int64_t ora_median_filter(int64_t n, const float* in, uint32_t half_window, float* out);
int64_t ora_repeated_median_filter(int64_t n, const float* in, uint32_t max_half_window, float* out);

/* CanvasBin counting loops (oracle/bin.cpp). possible: one byte per position. Returns the number of bins. */
int64_t ora_bin_hits(int64_t len, const uint8_t* hits, const uint8_t* possible, const char* bases, int bin_size, int mode,
                     const uint8_t* read_gc, const float* obs_vs_exp, int64_t max_bins, int32_t* start, int32_t* stop,
                     int32_t* count, uint8_t* gc);
/* flags: bit0 mapped, 1 mate mapped, 2 primary, 3 paired, 4 proper pair, 5 duplicate, 6 failed QC. Returns usableFragmentCount. */
int64_t ora_bin_alignments(int64_t n, const uint8_t* flags, const int32_t* pos, const int32_t* mate_pos, const int32_t* ref_id,
                           const int32_t* mate_ref_id, const int32_t* frag_len, const uint32_t* mapq, const int64_t* name_id,
                           uint32_t quality_threshold, int64_t n_bins, const int32_t* bin_start, const int32_t* bin_stop,
                           int32_t* count);

/* float.ToString("F2") then Convert.ToDouble: the .cleaned file round trip (IO.cs:21). */
void ora_f2_roundtrip(int64_t n, const float* in, double* out);

typedef struct {
    int is_germline;
    double mad_factor;  /* CanvasPartitionParameters.MadFactor, 5.0 */
    double thr_lower;   /* WaveletsRunnerParams.ThresholdLower (= ThresholdLowerMaf 0.05) */
    double thr_upper;   /* 80 */
    int min_size;       /* 10 */
    int evenness_window; /* EvennessScoreWindow 100000 */
    int n_threads;      /* CPU threads for the per-chromosome fan-out (Parallel.ForEach) */
} ora_wavelet_opts;

int ora_partition_wavelet(const ora_wavelet_opts* o, int n_chrom, const int64_t* chrom_off,
                          const double* coverage, int32_t* n_bp, int32_t* bp, double* evenness,
                          int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three);

/* Piecewise entry points used by the known-answer tests. */
int ora_coverage_variability(int window, int n_chrom, const int64_t* chrom_off, const double* cov,
                             double* cv);                                  /* returns has_value */
void ora_factor_of_three(int n_chrom, const int64_t* chrom_off, const double* cov, double* f3 /*[9]*/);
int ora_evenness_score(int window, int n_chrom, const int64_t* chrom_off, const double* cov,
                       double* score);                                     /* returns ok */
int ora_haar_wavelets(int64_t n, const double* ratio, double thr_lower, double thr_upper,
                      int is_germline, double mad_factor, int has_cv, double cv, const double* f3,
                      int n_f3, int32_t* bp /*[n]*/);                       /* returns n_bp */
/* Unbalanced-Haar tree of one chromosome, flattened level by level: per node level, start, brk,
 * end (1-based as in the reference) and coefficient.  Returns the node count. */
int64_t ora_uh_tree(int64_t n, const double* x, int32_t* level, int32_t* start, int32_t* brk,
                    int32_t* end, double* coef, double* smooth);

int ora_loess_train(int n, const double* x, const double* y, double bandwidth, int robustness_iters,
                    double x_step, double* fitted_orig_order, int n_query, const double* xq,
                    double* yq);
double ora_golden_section_quadratic(double a, double b);
double ora_median_f32(int64_t n, const float* x);
double ora_median_f64(int64_t n, const double* x);
void ora_quartiles_f32(int64_t n, const float* x, float* q /*[3]*/);
void ora_weighted_quantiles(int64_t n, const float* v, const float* w, int n_probs,
                            const float* probs, double* q);
/* .NET Core 2.0 Array.Sort(T[], Comparison<T>) introsort, restated; sorts indices 0..n-1 by
 * descending counts the way WaveletSegmentation.HardThresh does (WaveletSegmentation.cs:87). */
void ora_dotnet_sort_levels(int n, const int32_t* counts, int32_t* indices);

#ifdef __cplusplus
}
#endif
#endif
