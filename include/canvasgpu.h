/* libcanvasgpu — C-ABI of the B200-native coverage-to-segments engine.
 *
 * Drop-in boundary for the numeric blocks of the Illumina/canvas module executables
 * (reference @ v1.40.0).  The reference has no FFI: its only interface is "process + gzip-TSV file"
 * (Src/Canvas/CanvasClean/CanvasClean.cs:415-533, Src/Canvas/CanvasPartition/CanvasPartition.cs:24-187).
 * Each entry point below replaces the in-memory computation between a module's file reader and its
 * file writer; the C# hosts keep parsing flags and files and call these through [DllImport]
 * (binding shown in INTEGRATION.md).
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (CG_ERR_*); cg_last_error() gives the text;
 *  - all pointers are HOST memory owned by the caller; the library copies host<->device inside the
 *    call and never keeps a pointer after returning.  Buffers obtained from cg_host_alloc() are
 *    page-locked, which makes those copies run at PCIe speed;
 *  - outputs are caller-allocated at worst-case capacity (documented per argument);
 *  - one cg_ctx per GPU; a ctx is not thread-safe; different ctxs may be used concurrently;
 *  - calls are synchronous.  There is no CPU fallback: without a CUDA device cg_create fails.
 */
#ifndef CANVASGPU_H
#define CANVASGPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_ctx cg_ctx;

enum {
    CG_OK = 0,
    CG_ERR_CUDA = -1,        /* CUDA runtime failure */
    CG_ERR_ARG = -2,         /* invalid argument */
    CG_ERR_UNSORTED = -3,    /* chromosome ids not grouped in non-decreasing runs */
    CG_ERR_UNSUPPORTED = -4, /* option not implemented by this build */
    CG_ERR_CAPACITY = -5     /* internal work queue overflow */
};

int cg_create(int device, cg_ctx** out);
void cg_destroy(cg_ctx* ctx);
const char* cg_last_error(cg_ctx* ctx);
/* library + device description, e.g. "canvasgpu 0.1 sm_100 NVIDIA B200 148 SMs" */
const char* cg_describe(cg_ctx* ctx);

/* Page-locked host buffers (optional; plain malloc'ed memory is accepted everywhere). */
void* cg_host_alloc(size_t bytes);
void cg_host_free(void* p);

/* Device time (ms, CUDA events on the ctx stream) of the kernels of the last call, excluding the
 * host<->device copies; and the number of kernel launches it made. */
double cg_last_kernel_ms(cg_ctx* ctx);
int cg_last_launches(cg_ctx* ctx);
/* Device time (ms) between the events that bracket one stage of the last call: 0 Clean pipeline, 1 partition prefix sums +
 * order statistics + thresholds, 2 the per-chromosome pipelines (Unbalanced-Haar decomposition stages and the finish of every
 * chromosome, from the seed kernel to the last finish), 3 packing of the results.  -1 when the stage did not run.  Stages 1
 * and 2 OVERLAP: the pipelines start as soon as the prefix sums exist and run beside the order statistics, so the stage times
 * add up to more than cg_last_kernel_ms.  Fused call only: 4 = from the end of Clean to the end of the work enqueued before
 * the mid-call wait (coverage round trip, clears), 5 = from there to the start of the prefix sums (main stream idle while the
 * host plans the partition — the range-quantile index is built on a side stream meanwhile — plus the plan upload). */
double cg_last_stage_ms(cg_ctx* ctx, int stage);
/* Work counters of the last partition call: out[0] = bin visits of the decomposition (sum over tree
 * nodes of their length: L_eff * N), out[1] = tree nodes, out[2] = candidate nodes recorded for the threshold (every chain /
 * mid-stage node + the small / tiny nodes above the candidate threshold), out[3] = bins, out[4..6] = bin visits of the
 * big (chain + mid) / warp / per-thread tiers, out[7..9] = nodes of those tiers, out[10] / out[11] = ms (device %globaltimer)
 * from the first decomposition kernel to the last chain node / to the last tiny-stage thread of any chromosome, out[12],
 * out[13] = unused (0), out[14] = deepest tree, out[15] = 1 when the order statistics ran on integer hundredths (0: a coverage
 * value was not a two-decimal number below 2^22 / 100 and the f64 keys were used).
 * After cg_pedigree_hmm: the phase figures listed there.  Returns the number of values written (<= min(n, 16)). */
int cg_last_partition_stats(cg_ctx* ctx, double* out, int n);

/* ---------------------------------------------------------------------------------------------
 * CanvasClean — replaces CanvasClean.Main between CanvasIO.ReadFromTextFile (CanvasClean.cs:474)
 * and CanvasIO.WriteToTextFile (:530): RemoveBigBins :328-355, RemoveOutliers :387-413,
 * GetLocalStandardDeviation :268-298, RemoveBinsWithExtremeGC :207-237, NormalizeByGC :163-196,
 * NormalizeVarianceByGC :34-97, RemoveBinsWithExtremeLocalSD :308-322.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int size_filter;     /* -s|filtsize */
    int outlier_filter;  /* -r|outliers */
    int gc_norm;         /* -g|gcnorm */
    int gc_mode;         /* -m|mode: 0 MedianByGC (default), 1 LOESS */
    int want_local_sd;   /* --local-sd-metric-file given */
    int min_bins_per_gc; /* -w|weightedmedian, default 100 */
} cg_clean_opts;

/* chrom[i]: dense id of the chromosome run bin i belongs to, in file order (non-decreasing);
 * chrom_is_autosome / chrom_is_chrY: [n_chrom] flags decided by the host from the names
 * (GenomeMetadata.SequenceMetadata.IsAutosome; LoessGCNormalizer.cs:52-53).
 * Outputs: *n_out surviving bins; kept_index[0..n_out) their positions in the input (ascending);
 * count_out[0..n_out) their normalised counts; *local_sd the "#localSD" metric (-1 when not
 * computed, CanvasClean.cs:489); *gc_norm_skipped set when every bin was GC-filtered and
 * normalisation was skipped (:502-505).  kept_index and count_out need capacity n. */
int cg_clean(cg_ctx* ctx, const cg_clean_opts* opts, int64_t n, const uint8_t* chrom,
             const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
             const int32_t* start, const int32_t* stop, const float* count, const uint8_t* gc,
             int64_t* n_out, int32_t* kept_index, float* count_out, double* local_sd,
             int* gc_norm_skipped);

/* ---------------------------------------------------------------------------------------------
 * CanvasPartition, wavelets — replaces WaveletsRunner.Run (WaveletsRunner.cs:52-72):
 * GetCoverageVariability (Segmentation.cs:309-347), FactorOfThreeCoverageVariabilities (:364-429),
 * GetEvennessScore (:260-297) and, per chromosome, WaveletSegmentation.HaarWavelets
 * (WaveletSegmentation.cs:385-426).  Breakpoint -> segment conversion (Segmentation.cs:83-125) and
 * SegmentationResultsProcessor stay on the host.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int is_germline;     /* -g */
    double mad_factor;   /* CanvasPartitionParameters.MadFactor (5.0) */
    double thr_lower;    /* ThresholdLowerMaf (0.05), WaveletsRunner.cs:35 */
    double thr_upper;    /* 80 */
    int min_size;        /* WaveletsRunnerParams.MinSize (10) */
    int evenness_window; /* EvennessScoreWindow (100000) */
} cg_wavelet_opts;

/* coverage: all chromosomes concatenated; chrom_off[n_chrom + 1] their offsets.
 * Outputs: n_bp[c] breakpoints of chromosome c written at bp[chrom_off[c] ...] (bin indices within
 * the chromosome, ascending, first is 0; n_bp[c] = 0 when the chromosome has <= min_size bins);
 * *evenness (x100 score) valid when *evenness_ok; *cv valid when *cv_has_value;
 * factor_of_three[9].  bp needs capacity chrom_off[n_chrom]. */
int cg_partition_wavelet(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom,
                         const int64_t* chrom_off, const double* coverage, int32_t* n_bp,
                         int32_t* bp, double* evenness, int* evenness_ok, double* cv,
                         int* cv_has_value, double* factor_of_three);

/* The same, for the chromosomes of one shard only (multi-GPU: each rank passes the full coverage
 * array so that the genome-wide scalars are identical everywhere, and a 0/1 mask of the chromosomes
 * it segments; n_bp of unselected chromosomes is 0). */
int cg_partition_wavelet_shard(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom,
                               const int64_t* chrom_off, const double* coverage,
                               const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                               double* evenness, int* evenness_ok, double* cv, int* cv_has_value,
                               double* factor_of_three);

/* Clean followed by wavelet partition without leaving the device between the two: the cleaned
 * counts are rounded exactly as the .cleaned file round trip does (IO.cs:21 "{3:F2}" then
 * Convert.ToDouble, CanvasSegment.cs:1147) before segmentation.  Outputs of both stages as above;
 * chrom_off_out[n_chrom + 1] receives the per-chromosome offsets of the surviving bins. */
int cg_clean_partition_wavelet(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts,
                               int64_t n, const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                               const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                               const int32_t* stop, const float* count, const uint8_t* gc,
                               int64_t* n_out, int32_t* kept_index, float* count_out,
                               double* local_sd, int* gc_norm_skipped, int64_t* chrom_off_out,
                               int32_t* n_bp, int32_t* bp, double* evenness, int* evenness_ok,
                               double* cv, int* cv_has_value, double* factor_of_three);

/* Multi-GPU form of the fused call: Clean and the genome-wide scalars run on every rank (they need genome-wide order
 * statistics), only the chromosomes with chrom_selected[c] != 0 are segmented (n_bp = 0 for the others). */
int cg_clean_partition_wavelet_shard(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts,
                                     int64_t n, const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                                     const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                                     const int32_t* stop, const float* count, const uint8_t* gc,
                                     const uint8_t* chrom_selected, int64_t* n_out, int32_t* kept_index,
                                     float* count_out, double* local_sd, int* gc_norm_skipped, int64_t* chrom_off_out,
                                     int32_t* n_bp, int32_t* bp, double* evenness, int* evenness_ok, double* cv,
                                     int* cv_has_value, double* factor_of_three);

/* ---------------------------------------------------------------------------------------------
 * CanvasPartition -m CBS: circular binary segmentation, CBSRunner.Run (CBSRunner.cs:40-151) with
 * ChangePoint.ChangePoints (ChangePoint.cs:44-153) per chromosome on its own MersenneTwister stream
 * (seeds drawn from MersenneTwister(seed) in chromosome order, :107-112).  Coverage must be finite.
 * sbdry = the sequential boundary of GetBoundary.ComputeBoundary (GetBoundary.cs:19-160); NULL makes
 * the library compute it (cg_cbs_boundary) from n_perm, alpha, eta.  Outputs per chromosome c at
 * chrom_off[c]: n_seg[c] segment lengths (in bins) and means (lengthSeg / segmentMeans, :140-152).
 * stats (optional, [4]) = tests run, permutations consumed, permuted bins, edge-test draws.
 * Supported: hybrid p-value method, every undo method (prune up to 2^28 scored subsets per chromosome),
 * k_max <= 32, n_min <= 256 (the CanvasPartition defaults are hybrid / none / 25 / 200); anything else
 * returns CG_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    double alpha;      /* CanvasPartitionParameters.CBSalpha, 0.01 */
    uint32_t n_perm;   /* 10000 */
    int hybrid;        /* 1 */
    int min_width;     /* 2 */
    int k_max;         /* 25 */
    uint32_t n_min;    /* 200 */
    double eta;        /* 0.05 */
    double trim;       /* 0.025 (only used by the sdundo method) */
    int undo;          /* 0 none, 1 prune, 2 sdundo */
    double undo_prune; /* 0.05 */
    double undo_sd;    /* 3 */
    uint32_t seed;     /* 0 = the reference's seed generator */
} cg_cbs_opts;
int cg_partition_cbs(cg_ctx* ctx, const cg_cbs_opts* opts, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                     const int64_t* chrom_off, const double* coverage, int32_t* n_seg, int32_t* seg_len, double* seg_mean,
                     int64_t* stats);
/* Multi-GPU: segment only the chromosomes with chrom_selected[c] != 0 (the others get n_seg = 0).  Every chromosome
 * keeps the random stream it has in the whole-genome call, so the union over ranks equals cg_partition_cbs. */
int cg_partition_cbs_shard(cg_ctx* ctx, const cg_cbs_opts* opts, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                           const int64_t* chrom_off, const double* coverage, const uint8_t* chrom_selected, int32_t* n_seg,
                           int32_t* seg_len, double* seg_mean, int64_t* stats);
/* -s Prune on one chromosome (ChangePoint.ChangePointsPrune, ChangePoint.cs:205-271; Prune.cs:18-76), the host-side
 * step cg_partition_cbs applies with undo = 1: seg_len[n_seg] (n_seg >= 2, summing to n) -> seg_len_out, returns the
 * new segment count (>= 1) or an error code; CG_ERR_UNSUPPORTED when more than max_subsets (0 = the 2^28 that
 * cg_partition_cbs allows) subsets would be scored.  Host code only (no device needed). */
int cg_cbs_prune(const double* g, int64_t n, const int32_t* seg_len, int n_seg, double cutoff, int64_t max_subsets,
                 int32_t* seg_len_out, int64_t* subsets_scored);
/* Sequential boundary table; returns its length maxOnes (maxOnes + 1) / 2 (out may be NULL), < 0 on bad arguments. */
int64_t cg_cbs_boundary(uint32_t n_perm, double alpha, double eta, uint32_t* out, int64_t cap);

/* ---------------------------------------------------------------------------------------------
 * CanvasPartition -m HMM / -m PerSampleHMM (the SmallPedigree default, CanvasRunner.cs:927):
 * HiddenMarkovModelsRunner.Run (HiddenMarkovModelsRunner.cs:23-109) — negative-binomial emissions for copy
 * numbers 0..4 (InitializeNegativeBinomialEmission :111-153, Distributions.cs:206-217), outliers clipped at
 * 5 x the largest haploid mean (:155-163), HiddenMarkovModel.BestPathViterbi (HMM.cs:62-130) per chromosome,
 * breakpoints where the state changes.  coverage is [n_samples][N] (sample-major, all chromosomes concatenated,
 * chrom_off[n_chrom + 1]); finite and >= 0.  per_sample = 1 (PerSampleHMM): one sample per call, whole-genome
 * median and IQR-based pseudo-variance, five distinct states; per_sample = 0 (HMM): up to 4 samples jointly,
 * per-chromosome median and variance, states 0|1 and 3|4 share their emission.  Outputs as for the wavelets:
 * n_bp[c] breakpoints at bp[chrom_off[c] ...] (first is 0; none when the chromosome has <= min_size bins);
 * states (optional, [N]) the Viterbi path.  SegmentationInput.DeriveSegments and SplitOverlappingSegments
 * stay on the host.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int n_states;         /* 5 */
    int per_sample;       /* isPerSample */
    int min_size;         /* 10 */
    int exact_sequential; /* 1: one thread per chromosome replays the reference loop (cross-check; slow) */
} cg_hmm_opts;
int cg_partition_hmm(cg_ctx* ctx, const cg_hmm_opts* opts, int n_samples, int n_chrom, const int64_t* chrom_off,
                     const double* coverage, int32_t* n_bp, int32_t* bp, uint8_t* states);
/* Multi-GPU: only the chromosomes with chrom_selected[c] != 0 are segmented (the statistics of the emission
 * model are whole-genome in per-sample mode, so every rank passes the full coverage). */
int cg_partition_hmm_shard(cg_ctx* ctx, const cg_hmm_opts* opts, int n_samples, int n_chrom, const int64_t* chrom_off,
                           const double* coverage, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                           uint8_t* states);

/* The same from the float counts of the .cleaned table, with the text round trip that stands between CanvasClean and
 * CanvasPartition applied on the device: text_mode 1 = "{3:F2}" + Convert.ToDouble (IO.cs:21, CanvasSegment.cs:1147),
 * 2 = float.ToString() of the pedigree workflow's merged four-column file (CanvasRunner.cs:895-897), 0 = plain widening.
 * count is [n_samples][N]; chrom_selected may be NULL (every chromosome).  Half the bytes cross PCIe and the host does
 * not format / parse three million numbers per sample. */
int cg_partition_hmm_counts(cg_ctx* ctx, const cg_hmm_opts* opts, int n_samples, int n_chrom, const int64_t* chrom_off,
                            const float* count, int text_mode, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                            uint8_t* states);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: chromosomes are independent once the genome-wide scalars exist, so every rank
 * segments the chromosomes a longest-processing-time-first assignment gives it, and ONE NCCL all-gather over NVLink
 * reassembles the whole-genome result on every rank.  This replaces the lock-protected dictionary merges of the
 * reference's Parallel.ForEach over chromosomes (WaveletsRunner.cs:128-131, CBSRunner.cs:140-143,
 * HiddenMarkovModelsRunner.cs:90-105).  One communicator per context:
 *   one process per GPU  : rank 0 calls cg_comm_unique_id, the host hands the 128 bytes to every rank (any channel),
 *                          every rank calls cg_comm_init(ctx, n_ranks, rank, id)  (ncclCommInitRank);
 *   one process, n GPUs  : cg_comm_init_all(n, ctxs) (ncclCommInitAll); the sharded calls of the n contexts must then be
 *                          issued from n host threads, because each blocks in the all-gather until every rank arrives.
 * n_ranks = 1 gives a loopback communicator (no NCCL call; the sharded entry points then equal the plain ones).
 * libnccl.so.2 is bound at run time on the first cg_comm_* call (the one already loaded in the process, if any).
 * The *_sharded entry points take the arguments of their single-GPU forms on EVERY rank (the genome-wide scalars and
 * Clean need the whole sample) and return the whole-genome result on every rank; owner (optional, [n_chrom]) receives
 * the chromosome -> rank assignment (cg_shard_assign on the chromosome lengths).  The exchange carries, per rank, a
 * packed int32 list [length, payload]: one fixed-capacity all-gather (64 KiB per rank); lists that do not fit travel in a
 * second, exactly sized round that all ranks enter together (the decision is taken from the gathered lengths, so no rank
 * can be left waiting in a collective).
 * ------------------------------------------------------------------------------------------- */
#define CG_COMM_ID_BYTES 128
int cg_comm_unique_id(uint8_t* id /* [CG_COMM_ID_BYTES] */);
int cg_comm_init(cg_ctx* ctx, int n_ranks, int rank, const uint8_t* id);
int cg_comm_init_all(int n, cg_ctx* const* ctxs);
int cg_comm_destroy(cg_ctx* ctx);
int cg_comm_rank(cg_ctx* ctx);               /* -1 without a communicator */
int cg_comm_size(cg_ctx* ctx);               /* 0 without a communicator */
int cg_comm_nccl_version(void);              /* e.g. 22809; -1 when libnccl cannot be bound */
double cg_comm_last_exchange_ms(cg_ctx* ctx); /* device time of the collectives of the last exchange */
/* LPT: heaviest unit first onto the least loaded rank (ties: earlier unit, lower rank).  Host code. */
int cg_shard_assign(int n_units, const int64_t* weight, int n_ranks, int32_t* owner);
/* All-gather of one int32 list per rank (BASELINE config 5: per-sample segment lists of independent samples):
 * counts[n_ranks] receives the list lengths, all[0 .. *n_total) the lists back to back in rank order. */
int cg_comm_allgather_lists(cg_ctx* ctx, int64_t n_local, const int32_t* local, int64_t* counts, int32_t* all, int64_t cap,
                            int64_t* n_total);
/* Broadcast of a host buffer from rank `root` (staged through the device, NCCL broadcast over NVLink): pedigree mode,
 * where the rank that cleaned a sample hands the cleaned bins to the ranks that segment its chromosomes. */
int cg_comm_broadcast(cg_ctx* ctx, void* buf, int64_t bytes, int root);

int cg_partition_wavelet_sharded(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom, const int64_t* chrom_off,
                                 const double* coverage, int32_t* n_bp, int32_t* bp, double* evenness, int* evenness_ok,
                                 double* cv, int* cv_has_value, double* factor_of_three, int32_t* owner);
/* LPT weights = the chromosome run lengths of the INPUT (known before Clean, identical on every rank). */
int cg_clean_partition_wavelet_sharded(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts, int64_t n,
                                       const uint8_t* chrom, const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY,
                                       int n_chrom, const int32_t* start, const int32_t* stop, const float* count,
                                       const uint8_t* gc, int64_t* n_out, int32_t* kept_index, float* count_out,
                                       double* local_sd, int* gc_norm_skipped, int64_t* chrom_off_out, int32_t* n_bp,
                                       int32_t* bp, double* evenness, int* evenness_ok, double* cv, int* cv_has_value,
                                       double* factor_of_three, int32_t* owner);
/* stats (optional, [4]) are those of this rank's chromosomes. */
int cg_partition_cbs_sharded(cg_ctx* ctx, const cg_cbs_opts* opts, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                             const int64_t* chrom_off, const double* coverage, int32_t* n_seg, int32_t* seg_len,
                             double* seg_mean, int64_t* stats, int32_t* owner);
/* states (optional, [N]): the Viterbi path of every chromosome, gathered like the breakpoints. */
int cg_partition_hmm_sharded(cg_ctx* ctx, const cg_hmm_opts* opts, int n_samples, int n_chrom, const int64_t* chrom_off,
                             const double* coverage, int32_t* n_bp, int32_t* bp, uint8_t* states, int32_t* owner);

/* ---------------------------------------------------------------------------------------------
 * Pedigree step between CanvasClean and CanvasPartition: keep the bins that survived CanvasClean in EVERY
 * sample — Utilities.MergeMultiSampleCleanedBedFile (CanvasCommon/Utilities.cs:834-920), written back per
 * sample by CanvasRunner.NormalizeCanvasClean (Canvas/CanvasRunner.cs:883-903).  Sample s has n[s] bins in the
 * columns chrom[s] (dense chromosome ids shared by all samples), start[s], stop[s], count[s], ordered by
 * (chromosome id, start) without duplicates (CG_ERR_UNSORTED otherwise).  Outputs, capacity n[0]: *n_out common
 * bins in the first sample's order; kept_index = their positions in sample 0; stop_out = the stop of the last
 * sample (the reference overwrites it per file, :885); count_out[s * n[0] + k] = sample s's count of common bin k.
 * A kept bin with start < 0 or start >= stop is the reference's IlluminaException (CG_ERR_ARG, same message).
 * ------------------------------------------------------------------------------------------- */
int cg_merge_common_bins(cg_ctx* ctx, int n_samples, const int64_t* n, const uint8_t* const* chrom,
                         const int32_t* const* start, const int32_t* const* stop, const float* const* count,
                         int64_t* n_out, int32_t* kept_index, int32_t* stop_out, float* count_out);

/* The same merge for samples cleaned from ONE bin layout (a pedigree is binned once, so a bin is identified by its index
 * in that layout): kept[s][0 .. n_kept[s]) = cg_clean's kept_index of sample s (strictly increasing indices below n_bins),
 * count[s] its normalised counts.  Outputs, capacity n_kept[0]: *n_out common bins, common_index[k] their indices in the
 * layout (ascending = the first sample's order), count_out[s * n_kept[0] + k] sample s's count of common bin k.  The
 * coordinates of the common bins are layout columns the host already holds; nothing is gathered per sample. */
int cg_merge_kept_indices(cg_ctx* ctx, int64_t n_bins, int n_samples, const int64_t* n_kept, const int32_t* const* kept,
                          const float* const* count, int64_t* n_out, int32_t* common_index, float* count_out);

/* Software pipelining across calls (a cohort run processes sample after sample): start copying the NEXT sample's columns into
 * a staging slot of the device while the call for the current sample computes.  Returns at once.  The next cg_clean /
 * cg_clean_partition_wavelet* call that is given exactly these arrays (same pointers, same n) reads the staged columns and
 * copies nothing; the arrays must stay unchanged (and, for the copy to overlap, page-locked: cg_host_alloc) until then.  Two
 * slots: one prefetch per call keeps one slot filling while the other is read; a staged copy that two calls in a row did not
 * ask for is dropped.  Replaces nothing in the reference (its
 * modules read their input files before they compute); it hides the 14 B/bin upload that the file read was. */
int cg_prefetch_bins(cg_ctx* ctx, int64_t n, const uint8_t* chrom, const int32_t* start, const int32_t* stop, const float* count,
                     const uint8_t* gc);

/* ---------------------------------------------------------------------------------------------
 * The SmallPedigree chain in one call, device resident: CanvasClean per sample (Canvas/CanvasRunner.cs:883-893) ->
 * bins common to every sample (CanvasRunner.cs:895-903, CanvasCommon/Utilities.cs:834-920) -> CanvasPartition
 * -m PerSampleHMM per sample (CanvasRunner.cs:927, CanvasPartition/HiddenMarkovModelsRunner.cs:23-109).  The samples of a
 * pedigree share one bin layout (chrom/start/stop/gc, n bins; CanvasRunner.cs:846-870); count[s] are sample s's counts.
 * The cleaned lists never leave the GPU between the stages: the layout is uploaded once, each sample adds 4 B/bin.
 * sharded != 0 (communicator of R ranks, every rank makes the same call): sample s is cleaned on rank s mod R (count[s]
 * may be NULL elsewhere), cleaned lists move GPU to GPU over NCCL, the n_samples x n_chrom (sample, chromosome) units of the
 * HMM are spread so that a rank touches as few samples as possible (R >= n_samples: the ranks r with r mod n_samples == s share
 * sample s's chromosomes longest-first; R < n_samples: whole samples round robin) and one all-gather completes the result on
 * EVERY rank.  owner (may be NULL): [n_samples][n_chrom] rank of every unit.
 * Outputs: n_kept / local_sd / gc_norm_skipped [n_samples] as cg_clean reports them; *n_common, common_index[k] (capacity
 * n) = layout index of common bin k, count_out[s * n + k] = sample s's cleaned count of it (what the merged .cleaned file
 * prints with float.ToString()) — both may be NULL together on a rank that does not write those files: the table is
 * 4 (n_samples + 1) bytes per bin, and N ranks downloading it through one host's PCIe root compete with each other; chrom_off_out[n_chrom + 1] offsets of the chromosomes among the common bins;
 * n_bp[s * n_chrom + c] breakpoints of (s, c) at bp[s * n + chrom_off_out[c] ...], as cg_partition_hmm numbers them.
 * cg_last_partition_stats: [0..4] wall ms of clean / exchange of cleaned lists / merge / HMM / gather on this rank,
 * [5] kernel ms, [6] launches, [7] device ms of the NCCL calls, [8] wall ms of the final download of the merged table.
 * ------------------------------------------------------------------------------------------- */
int cg_pedigree_hmm(cg_ctx* ctx, const cg_clean_opts* copts, const cg_hmm_opts* hopts, int n_samples, int64_t n,
                    const uint8_t* chrom, const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
                    const int32_t* start, const int32_t* stop, const float* const* count, const uint8_t* gc, int sharded,
                    int64_t* n_kept, double* local_sd, int* gc_norm_skipped, int64_t* n_common, int32_t* common_index,
                    float* count_out, int64_t* chrom_off_out, int32_t* n_bp, int32_t* bp, int32_t* owner);

/* ---------------------------------------------------------------------------------------------
 * CanvasSmooth — RepeatedMedianSmoother.Smooth (CanvasSmooth/CanvasSmooth.cs:44-77) over every chromosome:
 * Utilities.MedianFilter (CanvasCommon/Utilities.cs:767-791) with half windows 1 .. max_half_window, each pass on
 * the previous output.  count is all chromosomes concatenated (chrom_off[n_chrom + 1]).  n_out[c] smoothed counts of
 * chromosome c are written at count_out[chrom_off[c] ...]; a chromosome shorter than 2h + 1 bins loses bins exactly
 * as the reference's streaming window does (its first n_out[c] bins keep their coordinates, Enumerable.Zip :61).
 * ------------------------------------------------------------------------------------------- */
int cg_smooth(cg_ctx* ctx, int max_half_window, int n_chrom, const int64_t* chrom_off, const float* count,
              int64_t* n_out, float* count_out);

/* ---------------------------------------------------------------------------------------------
 * CanvasBin counting (BAM decoding, read pairing and FASTA handling stay on the host).
 *
 * cg_bin_hits — BinCountsForChromosome (CanvasBin.cs:568-661) for one chromosome without predefined
 * bins: hits[p] saturating per-position hit counts (HitArray.cs:61-64), possible_bits bit p of word
 * p/64 = unique-kmer start (:183-200), bases the FASTA characters.  A bin closes at its bin_size-th
 * possible position; count = sum of min(10, hits) over possible positions (mode 0, TruncatedDynamicRange
 * :618-625) or round(sum of min(10, hits / obs_vs_exp_gc[read_gc[p]])) accumulated in single precision
 * in position order (mode 1, GCContentWeighted :626-636); gc = (int)(100f * GC / length) (:638).  The
 * trailing incomplete bin is dropped as in the reference.  Outputs need capacity max_bins.
 *
 * cg_bin_screen — the passes over one chromosome's positions that precede the binning: ExcludeTagsOverlappingFilterFile
 * (CanvasBin.cs:668-692: possible[i] = false inside every filter interval [start, stop); an interval reaching past the
 * chromosome is the reference's ArgumentOutOfRangeException -> CG_ERR_ARG), ScreenObservedTags (:699-716: hits[i] = 0
 * where position i is not possible) and the two counts behind the chromosome's rate in GetRates (:56-58): positions with
 * hits[i] > 0 and possible positions.  hits and possible_bits are updated in place.  The host takes the median of
 * observed / possible over the autosomes and bin size = (int)(countsPerBin / median) (:79-83).
 *
 * cg_bin_fragment_stats / cg_bin_read_gc — the tables of the GCContentWeighted mode.  fragment_stats: sum and number of the
 * positive fragment lengths of one chromosome (Utilities.NonZeroMean, CanvasCommon/Utilities.cs:136-151; MeanFragmentSize,
 * CanvasBin.cs:164-174, is the host's integer division of the two, then the same mean over the chromosomes' means).
 * read_gc: GC content (0..100) of the read that starts at every position — fragment length frag_len[p], or mean_frag where
 * it is 0, at most 3 * mean_frag; positions from len - 3 * mean_frag - 1 on stay 0 (:450-497) — written to read_gc[len], and
 * this chromosome's share of expectedReadCountsByGC / observedReadCountsByGC ADDED to expected[101] / observed[101]
 * (ComputeObservedVsExpectedGC :341-358; the 101 ratios are host arithmetic, :374-387).  mean_frag in 1..10922.
 *
 * cg_bin_fragments — FragmentBinner.BinOneAlignment / FindBestBin (FragmentBinner.cs:296-311, :353-371):
 * fragment i = [frag_start, frag_stop) goes to the bin (sorted, non-overlapping) with the largest
 * overlap, the first one on ties; best_bin[i] = -1 when none.  undo_index lists fragments whose mate
 * later failed the duplicate / QC / MAPQ filters (:279-284): their bin is decremented again.
 * ------------------------------------------------------------------------------------------- */
int cg_bin_screen(cg_ctx* ctx, int64_t chr_len, uint8_t* hits, uint64_t* possible_bits, int64_t n_filter,
                  const int32_t* filter_start, const int32_t* filter_stop, int64_t* n_observed, int64_t* n_possible);
int cg_bin_fragment_stats(cg_ctx* ctx, int64_t len, const int16_t* frag_len, int64_t* sum, int64_t* count);
int cg_bin_read_gc(cg_ctx* ctx, int64_t len, const char* bases, const int16_t* frag_len, int mean_frag, const uint8_t* hits,
                   uint8_t* read_gc, int64_t* expected, int64_t* observed);
int cg_bin_hits(cg_ctx* ctx, int64_t chr_len, const uint8_t* hits, const uint64_t* possible_bits, const char* bases,
                int bin_size, int mode, const uint8_t* read_gc, const float* obs_vs_exp_gc, int64_t max_bins,
                int64_t* n_bins, int32_t* start, int32_t* stop, int32_t* count, uint8_t* gc);
int cg_bin_fragments(cg_ctx* ctx, int64_t n_frag, const int32_t* frag_start, const int32_t* frag_stop, int64_t n_undo,
                     const int32_t* undo_index, int64_t n_bins, const int32_t* bin_start, const int32_t* bin_stop,
                     int32_t* best_bin, int32_t* count);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone normalise-apply stream (the kernel BASELINE.json's roofline target names):
 * count_out[i] = (float)(global_median * (double)count[i] / median_by_gc[gc[i]]) when the bucket
 * median is > 0, else count[i] (CanvasClean.cs:190-195).  `batch` independent samples of n bins laid
 * out back to back; medians are [batch][101], global_median [batch].  Used by bench.py to measure
 * the kernel on arrays larger than L2; *kernel_ms returns its device time.
 * ------------------------------------------------------------------------------------------- */
int cg_normalize_apply(cg_ctx* ctx, int batch, int64_t n, const float* count, const uint8_t* gc,
                       const double* median_by_gc, const double* global_median, float* count_out,
                       int repeats, double* kernel_ms);

/* ---------------------------------------------------------------------------------------------
 * Text codecs of the .binned / .cleaned files (host code, multi-threaded; no ctx, no device work) — the per-line
 * loops of CanvasIO.WriteToTextFile / ReadFromTextFile (CanvasCommon/IO.cs:15-52) on the uncompressed text; gzip stays
 * with the caller.
 *
 * cg_format_bins: "chr \t start \t stop \t count \t gc \n" per bin, count as .NET Core 2.0 prints {0:F2} (IO.cs:21);
 * four_columns != 0: "chr \t start \t stop \t count \n" with float.ToString() (CanvasRunner.cs:895-897; gc may be
 * NULL).  names[chrom[i]] is the chromosome of bin i.  Returns the text length; when out is NULL or cap is too small
 * nothing is written (size query).  < 0: bad argument.
 *
 * cg_parse_bins: columns of every non-empty line (4 or 5 columns; gc = 0 when absent); chrom[i] = id of the
 * chromosome RUN of line i (a name that reappears later starts a new run); the run names are written back to back,
 * NUL-terminated, into names.  Returns the number of rows (if > max_rows the buffers were too small and nothing is
 * complete); -1 bad argument, -2 malformed line, -3 more than 256 runs or names_cap too small.
 * ------------------------------------------------------------------------------------------- */
int64_t cg_format_bins(int64_t n, int n_names, const char* const* names, const uint8_t* chrom, const int32_t* start,
                       const int32_t* stop, const float* count, const uint8_t* gc, int four_columns, char* out,
                       int64_t cap, int n_threads);
int64_t cg_parse_bins(const char* text, int64_t len, int64_t max_rows, uint8_t* chrom, int32_t* start, int32_t* stop,
                      float* count, uint8_t* gc, int* n_names, char* names, int64_t names_cap, int n_threads);

/* ---------------------------------------------------------------------------------------------
 * CanvasNormalize (tumour / control-panel coverage ratio; enrichment workflows).
 * cg_normalize_reference = WeightedAverageReferenceGenerator.Run (CanvasNormalize/WeightedAverageReferenceGenerator.cs:43-70):
 *   counts[n_samples][n] (doubles: double.Parse of column 4, or the widened float when a manifest is given,
 *   BinCounts.cs:86-100 / :102-160), on_target[n] (bins overlapping the manifest regions; NULL = every bin);
 *   out: median[s] = OnTargetMedianBinCount (BinCounts.cs:41-62), weight[s] = (1 / median or 0) / sum, reference[n] = the
 *   weighted bin counts.  The single-control case is a file copy in the reference (:37-41) and stays with the host.
 * cg_normalize_ratio = LSNormRatioCalculator.Run (mode 1, LSNormRatioCalculator.cs:22-48; min_ref / max_ref ignored, bins
 *   with a reference count < 1 are skipped) or RawRatioCalculator.Run (mode 0, RawRatioCalculator.cs:24-47; bins with a
 *   reference count outside [min_ref, max_ref] are skipped), followed by CanvasNormalizeUtilities.RatiosToCounts
 *   (CanvasNormalizeUtilities.cs:22-31) with the reference ploidy of every bin (ploidy[n]; NULL = 2).  Outputs for the
 *   *n_out kept bins, in order: kept_index (into the input), ratio (the .cnd column) and count (the output file's column 4).
 * ------------------------------------------------------------------------------------------- */
/* BestLR2ReferenceGenerator.Run (CanvasNormalize/BestLR2ReferenceGenerator.cs:32-125): sample[n] and controls[n_controls][n]
 * as doubles; every vector is divided by the median of its on-target counts, and *best_index is the control with the
 * smallest mean squared log ratio over the on-target bins (first strict minimum; -1 when none is below +infinity, where
 * the reference throws).  mean_sq_log_ratio / ignored: [n_controls].  The sums are added chunk-wise on the device:
 * they agree with the reference's left-to-right sum to better than 1e-9 relative, not bit for bit. */
int cg_normalize_best_lr2(cg_ctx* ctx, int n_controls, int64_t n, const double* sample, const double* controls,
                          const uint8_t* on_target, int* best_index, double* mean_sq_log_ratio, int64_t* ignored);
/* PCAReferenceGenerator.Run (CanvasNormalize/PCAReferenceGenerator.cs:37-78): sample[n] (the file's counts), the model's
 * mean mu[n] and axes[n_axes][n] as read from the model file (they are scaled to unit length and checked for pairwise
 * orthogonality as PCAModel.LoadModel does, :113-146: CG_ERR_ARG "Axes are not orthogonal to each other."); reference[n] =
 * (float)(max(1, mu + projection of the centred sample) * median ratio), the median taken over the raw sample / reference
 * ratios of the bins whose (two-decimal) reference count lies in [min_ref, max_ref].  Dot products are chunk-wise device
 * sums: results agree with the reference's left-to-right sums to ~1e-12 relative (an occasional last-bit difference in the
 * float outputs), not bit for bit.  n_axes <= 10. */
int cg_normalize_pca_reference(cg_ctx* ctx, int64_t n, int n_axes, const float* sample, const float* mu, const double* axes,
                               const uint8_t* on_target, double min_ref, double max_ref, float* reference, double* median_ratio);
int cg_normalize_reference(cg_ctx* ctx, int n_samples, int64_t n, const double* counts, const uint8_t* on_target,
                           double* median, double* weight, double* reference);
int cg_normalize_ratio(cg_ctx* ctx, int64_t n, const float* sample, const float* reference, const uint8_t* on_target,
                       int mode, double min_ref, double max_ref, const int32_t* ploidy, int64_t* n_out,
                       int32_t* kept_index, float* ratio, float* count, double* library_size_factor);

#ifdef __cplusplus
}
#endif
#endif
