// CanvasClean on the device: filters, local-SD metric, per-GC exact medians / quartiles and the
// streaming normalise kernels.  Reference: Src/Canvas/CanvasClean/CanvasClean.cs (line numbers in
// the comments of each kernel).  Nothing here syncs with the host: every data-dependent decision
// (list lengths after a filter, "metric enabled", "variance normalisation fired", ...) lives in a
// device-side CleanCtl block that later kernels read.
#pragma once
#include <memory>
#include "common.cuh"
#include "select.cuh"

constexpr int GC_BINS = 101;          // EnrichmentUtilities.numberOfGCbins
constexpr int GC_SEGS = GC_BINS + 1;  // + one "all autosomal bins" segment
constexpr int MIN_BINS_PER_GC = 100;  // CanvasClean.cs:14 defaultMinNumberOfBinsPerGC
constexpr int LOCAL_SD_WINDOW = 20;   // CanvasClean.cs:283

struct CleanCtl {
    int n0;             // input bins
    int n1;             // bins after RemoveBigBins
    int n2;             // bins after RemoveOutliers
    int n3;             // bins alive after RemoveBinsWithExtremeGC
    int n_out;          // bins written
    int size_thresh;    // 98th-percentile bin size
    int size_filter_on;
    int metric_on;      // local-SD metric enabled (:483-486)
    int n_windows;      // 20-bin windows of consecutive-count differences
    int gc_thresh;      // RemoveBinsWithExtremeGC threshold (:226-227)
    int gc_skipped;     // every bin GC-filtered -> normalisation skipped (:502-505)
    int do_norm;        // GC normalisation runs
    int do_variance;    // NormalizeVarianceByGC is evaluated (:512)
    int variance_fired; // ... and rescaled something (:81-96) -> second NormalizeByGC
    int unsorted;       // chromosome ids not grouped
    int need_weighted;  // some used GC bucket has < 100 autosomal bins (weighted-quantile path)
    int pad;
    double local_sd;    // "#localSD" metric
    double global_median;
    float global_q[3];
    float global_iqr;
    unsigned hist_auto[GC_BINS];  // autosomal alive bins per GC
    unsigned hist_all[GC_BINS];   // alive bins per GC, all chromosomes
    unsigned n_auto;              // autosomal bins before the GC filter (totalCount, :214-224)
    unsigned n_auto3;             // autosomal bins alive after the GC filter
    double med[GC_BINS];          // per-GC median (NormalizeByGC) ; <= 0 disables the bucket
    float q2[GC_BINS];            // per-GC quartile 2 (NormalizeVarianceByGC localQuartiles.Item2)
    float iqr[GC_BINS];           // per-GC IQR (localIQR), -1 for an empty bucket
    int wq_big[GC_BINS];          // buckets whose weighted quantiles need more neighbours than the shared-memory sorter holds
};

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024); returns the exclusive prefix
// and the block total through `total`.
__device__ inline int block_excl_scan(int v, int& total) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int x = lane < nw ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, off);
            if (lane >= off) xi += t;
        }
        s_warp[lane] = xi - x;
        if (lane == 31) s_total = xi;
    }
    __syncthreads();
    int r = s_warp[w] + incl - v;
    total = s_total;
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// Stream compaction in three small kernels (count per tile, scan of tile counts, scatter).
// Pred: __device__ bool operator()(int i) const;   Emit: __device__ void operator()(int src, int dst) const
// ---------------------------------------------------------------------------------------------
constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = 8;
constexpr int CMP_TILE = CMP_THREADS * CMP_ITEMS;

template <class Pred>
__global__ void compact_count_kernel(Pred p, const int* __restrict__ n_ptr, int* __restrict__ tile_counts) {
    const int n = *n_ptr;
    const int base = blockIdx.x * CMP_TILE;
    int c = 0;
    if (base < n) {
#pragma unroll
        for (int t = 0; t < CMP_ITEMS; t++) {
            int i = base + t * CMP_THREADS + threadIdx.x;
            if (i < n && p(i)) c++;
        }
    }
    int total;
    block_excl_scan(c, total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

static __global__ void compact_scan_kernel(int* __restrict__ tile_counts, int ntiles, int* __restrict__ total_out) {
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < ntiles ? tile_counts[i] : 0;
        int total;
        int ex = block_excl_scan(v, total);
        if (i < ntiles) tile_counts[i] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

template <class Pred, class Emit>
__global__ void compact_scatter_kernel(Pred p, Emit e, const int* __restrict__ n_ptr,
                                       const int* __restrict__ tile_offsets) {
    const int n = *n_ptr;
    const int base = blockIdx.x * CMP_TILE;
    if (base >= n) return;
    // thread owns CMP_ITEMS consecutive elements so that the output order is the input order
    const int first = base + threadIdx.x * CMP_ITEMS;
    bool f[CMP_ITEMS];
    int c = 0;
#pragma unroll
    for (int t = 0; t < CMP_ITEMS; t++) {
        int i = first + t;
        f[t] = i < n && p(i);
        c += f[t];
    }
    int total;
    int ex = block_excl_scan(c, total);
    int dst = tile_offsets[blockIdx.x] + ex;
#pragma unroll
    for (int t = 0; t < CMP_ITEMS; t++)
        if (f[t]) e(first + t, dst++);
}

template <class Pred, class Emit>
inline void compact_run(cg_ctx* ctx, const Pred& p, const Emit& e, const int* n_ptr, int n_upper,
                        int* tile_counts, int* total_out) {
    int ntiles = std::max(1, div_up(n_upper, CMP_TILE));
    CG_LAUNCH(ctx, compact_count_kernel<Pred>, ntiles, CMP_THREADS, 0, p, n_ptr, tile_counts);
    CG_LAUNCH(ctx, compact_scan_kernel, 1, 1024, 0, tile_counts, ntiles, total_out);
    CG_LAUNCH(ctx, (compact_scatter_kernel<Pred, Emit>), ntiles, CMP_THREADS, 0, p, e, n_ptr, tile_counts);
}

// float.ToString("F2") followed by Convert.ToDouble — what the .cleaned file does to a count on its way
// to CanvasPartition (IO.cs:21, CanvasSegment.cs:1147).  .NET Core 2.0 formats a float from its 7
// significant decimal digits (FLOAT_PRECISION) and then rounds that digit string half-up to two
// decimals; parsing "ddd.dd" gives the double nearest to hundredths / 100.
__host__ __device__ inline double cg_pow10(int k) {  // literals instead of a per-thread table in local memory
    switch (k) {
        case 0: return 1e0; case 1: return 1e1; case 2: return 1e2; case 3: return 1e3; case 4: return 1e4;
        case 5: return 1e5; case 6: return 1e6; case 7: return 1e7; case 8: return 1e8; case 9: return 1e9;
        case 10: return 1e10; case 11: return 1e11; case 12: return 1e12; case 13: return 1e13; case 14: return 1e14;
        case 15: return 1e15; case 16: return 1e16; case 17: return 1e17; case 18: return 1e18; default: return 1e19;
    }
}

// Hundredths of |v| as float.ToString("F2") prints them (seven significant digits first, then half-up to two decimals);
// false for NaN / +-Infinity.  The .cleaned value CanvasPartition parses back is hundredths / 100 (IO.cs:21,
// CanvasSegment.cs:1147): a correctly rounded decimal parse and one IEEE division give the same double.
__host__ __device__ inline bool dotnet_f2_hundredths(float v, double& hundredths) {
    hundredths = 0.0;
    if (v != v || v - v != 0.0f) return false;
    double x = v < 0 ? -(double)v : (double)v;
    if (x == 0.0) return true;
    int e = 0;  // 10^e <= x < 10^(e+1)
    if (x >= 1.0) { while (e < 18 && x >= cg_pow10(e + 1)) e++; }
    else { e = -1; while (e > -19 && x * cg_pow10(-e) < 1.0) e--; }
    // seven significant digits, round-half-even on the exact value (the product is exact: a 24-bit
    // significand times 10^k, k <= 12, fits 53 bits)
    double scaled = (6 - e >= 0) ? (6 - e < 20 ? x * cg_pow10(6 - e) : 0.0) : x / cg_pow10(e - 6);
    double d7 = rint(scaled);
    if (d7 >= 1e7) { d7 /= 10.0; e += 1; }
    if (e >= 4) {
        hundredths = d7 * cg_pow10(e - 4 < 19 ? e - 4 : 19);
    } else {
        const int k = 4 - e;  // digits to drop
        if (k > 7) hundredths = 0.0;
        else {
            const long long q = (long long)d7;
            long long h, rem, p;
            // constant divisors: the compiler turns these into multiplications
            switch (k) {
                case 1: p = 10LL; h = q / 10LL; rem = q % 10LL; break;
                case 2: p = 100LL; h = q / 100LL; rem = q % 100LL; break;
                case 3: p = 1000LL; h = q / 1000LL; rem = q % 1000LL; break;
                case 4: p = 10000LL; h = q / 10000LL; rem = q % 10000LL; break;
                case 5: p = 100000LL; h = q / 100000LL; rem = q % 100000LL; break;
                case 6: p = 1000000LL; h = q / 1000000LL; rem = q % 1000000LL; break;
                default: p = 10000000LL; h = q / 10000000LL; rem = q % 10000000LL; break;
            }
            if (rem * 2 >= p) h += 1;  // first dropped digit >= 5
            hundredths = (double)h;
        }
    }
    return true;
}

__host__ __device__ inline double dotnet_f2_roundtrip(float v) {
    double hundredths;
    if (!dotnet_f2_hundredths(v, hundredths)) return (double)v;  // NaN / +-Infinity survive as such
    const double r = hundredths / 100.0;
    return v < 0 ? -r : r;
}

// float.ToString() (".NET Core 2.0: G7") followed by Convert.ToDouble — what the merged four-column .cleaned file of the
// pedigree workflow does to a count (CanvasRunner.cs:895-897): seven significant digits d7 at decimal shift k, parsed back
// as the double nearest to d7 / 10^k (one IEEE division of two exactly representable numbers).
__host__ __device__ inline double dotnet_g7_roundtrip(float v) {
    if (v != v || v - v != 0.0f) return (double)v;
    const double x = v < 0 ? -(double)v : (double)v;
    if (x == 0.0) return 0.0;
    int e = 0;  // 10^e <= x < 10^(e+1)
    if (x >= 1.0) { while (e < 18 && x >= cg_pow10(e + 1)) e++; }
    else { e = -1; while (e > -19 && x * cg_pow10(-e) < 1.0) e--; }
    const int k = 6 - e;
    const double p = cg_pow10(k >= 0 ? (k < 19 ? k : 19) : (-k < 19 ? -k : 19));
    const double d7 = rint(k >= 0 ? x * p : x / p);
    const double r = k >= 0 ? d7 / p : d7 * p;
    return v < 0 ? -r : r;
}

struct LoessDev;

// Device-side buffers of one cg_clean call (slices of the ctx arena).
struct CleanDev {
    int64_t n;
    int n_chrom;
    int64_t max_chrom_bins = -1;  // longest chromosome run of the input when the host knows it (-1: unknown)
    // input
    uint8_t *chrom, *gc;
    int32_t *start, *stop;
    float* count;
    // after RemoveBigBins
    uint8_t *chrom1, *gc1;
    float* count1;
    int32_t* orig1;
    // after RemoveOutliers (the list every later stage works on) + GC-filter mask
    uint8_t *chrom2, *gc2, *alive;
    float* count2;
    int32_t* orig2;
    // local-SD windows
    double* wsd;
    uint8_t* wchrom;
    unsigned* wcnt;
    double *wmed, *wmad;
    // outputs
    int32_t* kept;
    float* count_out;
    // scratch / control
    int* tiles;
    unsigned long long* wq_key;  // scratch of gc_weighted_big_kernel: wq_cap keys (a power of two >= n)
    long long wq_cap;
    CleanCtl* ctl;
    uint8_t* is_auto;
    uint8_t* is_chry;
    SelState<uint32_t> sel_size, sel_gc;
    SelState<uint64_t> sel_win;
    std::shared_ptr<LoessDev> lo;  // buffers of the LOESS mode (clean_loess.cuh), only allocated for -m LOESS
};

size_t clean_workspace_bytes(int64_t n, int n_chrom, bool loess = false);
int clean_alloc(cg_ctx* ctx, int64_t n, int n_chrom, CleanDev& d, bool loess = false);
int clean_enqueue(cg_ctx* ctx, const cg_clean_opts* o, CleanDev& d);
