// Range-quantile index over the coverage of each chromosome.
//
// Healing (WaveletSegmentation.cs:194-232) and refinement (:237-258) ask for exact medians of bin
// ranges that are decided sequentially, one after the other, per chromosome.  Answering each of them
// by radix passes over the range costs O(range) per query through a single SM.  Instead, one parallel
// pre-pass builds, per chromosome:
//   * 1023 splitters (sorted sample of the chromosome's keys) -> 1024 value buckets;
//   * for every tile of 1024 consecutive bins: its bucket histogram H[tile][1024] (u16), the bucket-sorted
//     copy of its keys, and the start of every bucket inside that copy;
//   * C[tile][b]: bins of bucket b in all earlier tiles of the chromosome.
// A range median then needs C[tile_r] - C[tile_l] (one 1024-wide subtraction), the <= 2046 bins of the two
// partial tiles, and the ~range/1024 keys of the one bucket that holds the wanted rank, which are read
// from the bucket-sorted tile copies and ranked in shared memory: exact, and ~20 KB instead of
// several passes over megabytes.  Buckets only route; the answer is always a true element.
#pragma once
#include "wavelet.cuh"

constexpr int RQ_TILE = 1024;
constexpr int RQ_BUCKETS = 1024;
constexpr int RQ_SAMPLE = 4096;

struct RqIndex {
    const unsigned long long* spl;     // [n_chrom][RQ_BUCKETS] splitter keys (last entry unused)
    const unsigned short* hist;        // [ntiles][RQ_BUCKETS]
    const unsigned short* tstart;      // [ntiles][RQ_BUCKETS] bucket starts inside the sorted tile
    const unsigned* cum;               // [ntiles + n_chrom][RQ_BUCKETS]: row tfirst[c] + c + t = tiles < t of chromosome c
    const unsigned long long* sorted;  // [N] bucket-sorted keys, tile by tile (same offsets as the coverage)
    const int* tfirst;                 // [n_chrom + 1] first tile of each chromosome
};

// bucket of a key = number of splitters strictly below it (0..RQ_BUCKETS-1)
__device__ inline int rq_bucket(const unsigned long long* __restrict__ s_spl, unsigned long long key) {
    int lo = 0, hi = RQ_BUCKETS - 1;  // the last splitter slot is a sentinel (~0)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_spl[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- 1. splitters: sorted sample of every selected chromosome (one CTA per chromosome)
__global__ void __launch_bounds__(1024)
rq_splitter_kernel(const double* __restrict__ cov, const long long* __restrict__ off, const unsigned char* __restrict__ selected,
                   unsigned long long* __restrict__ spl) {
    __shared__ unsigned long long s_key[RQ_SAMPLE];
    const int c = blockIdx.x;
    const long long o = off[c];
    const long long n = off[c + 1] - o;
    if (!selected[c] || n < 1) return;
    const int S = (int)(n < RQ_SAMPLE ? n : RQ_SAMPLE);
    for (int i = threadIdx.x; i < RQ_SAMPLE; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < S) k = f64_key(cov[o + ((long long)i * n) / S]);
        s_key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= RQ_SAMPLE; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < RQ_SAMPLE; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = s_key[i], b = s_key[l];
                    if ((a > b) == up) { s_key[i] = b; s_key[l] = a; }
                }
            }
            __syncthreads();
        }
    // picks: sample quantiles (j + 1) / RQ_BUCKETS
    __shared__ unsigned long long s_pick[RQ_BUCKETS];
    for (int j = threadIdx.x; j < RQ_BUCKETS; j += blockDim.x) {
        unsigned long long v = ~0ull;  // the last slot is a sentinel
        if (j < RQ_BUCKETS - 1) {
            int idx = (int)(((long long)(j + 1) * S) / RQ_BUCKETS);
            if (idx >= S) idx = S - 1;
            v = s_key[idx];
        }
        s_pick[j] = v;
    }
    __syncthreads();
    // Coverage is heavily discretised (integer counts times a handful of GC ratios): a frequent value
    // shows up as a run of equal picks.  The first pick of such a run becomes "key - 1", which makes
    // the next bucket hold exactly that one value — a pure bucket is answered without reading any bin.
    for (int j = threadIdx.x; j < RQ_BUCKETS; j += blockDim.x) {
        unsigned long long v = s_pick[j];
        if (j + 1 < RQ_BUCKETS - 1 && v != 0ull && v != ~0ull && s_pick[j + 1] == v && (j == 0 || s_pick[j - 1] != v)) v -= 1ull;
        spl[(size_t)c * RQ_BUCKETS + j] = v;
    }
}

// ---- 2. tiles: bucket histogram + bucket-sorted copy (one CTA of 256 threads per tile)
__global__ void __launch_bounds__(RQ_TILE)
rq_tile_kernel(const double* __restrict__ cov, const long long* __restrict__ off, const int* __restrict__ tfirst, int n_chrom,
               const unsigned char* __restrict__ selected, const unsigned long long* __restrict__ spl,
               unsigned short* __restrict__ hist, unsigned short* __restrict__ tstart, unsigned long long* __restrict__ sorted) {
    __shared__ unsigned long long s_spl[RQ_BUCKETS];
    __shared__ int s_cnt[RQ_BUCKETS], s_start[RQ_BUCKETS], s_fill[RQ_BUCKETS];
    __shared__ int s_warp[32];
    const int tile = blockIdx.x;
    if (tile >= tfirst[n_chrom]) return;  // the grid may be sized for an upper bound of the tile count
    // chromosome of this tile
    int lo = 0, hi = n_chrom - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tfirst[mid] <= tile) lo = mid; else hi = mid - 1; }
    const int c = lo;
    if (!selected[c]) return;
    const long long o = off[c], n = off[c + 1] - o;
    const long long t0 = (long long)(tile - tfirst[c]) * RQ_TILE;  // first bin of the tile inside the chromosome
    const int len = (int)((n - t0) < RQ_TILE ? (n - t0) : RQ_TILE);
    const int t = threadIdx.x;
    s_spl[t] = spl[(size_t)c * RQ_BUCKETS + t];
    s_cnt[t] = 0;
    s_fill[t] = 0;
    __syncthreads();
    unsigned long long key = 0;
    int b = 0;
    if (t < len) {
        key = f64_key(cov[o + t0 + t]);
        b = rq_bucket(s_spl, key);
        atomicAdd(&s_cnt[b], 1);
    }
    __syncthreads();
    // exclusive scan of the bucket counts
    const int lane = t & 31, w = t >> 5;
    const int v = s_cnt[t];
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        const int x = s_warp[lane];
        int xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += u; }
        s_warp[lane] = xi - x;
    }
    __syncthreads();
    s_start[t] = s_warp[w] + incl - v;
    __syncthreads();
    hist[(size_t)tile * RQ_BUCKETS + t] = (unsigned short)v;
    tstart[(size_t)tile * RQ_BUCKETS + t] = (unsigned short)s_start[t];
    if (t < len) {
        const int pos = s_start[b] + atomicAdd(&s_fill[b], 1);
        sorted[o + t0 + pos] = key;
    }
}

// ---- 3. cumulative counts over the tiles of each chromosome (one CTA per chromosome, one thread per bucket)
__global__ void __launch_bounds__(RQ_BUCKETS)
rq_cumulate_kernel(const int* __restrict__ tfirst, const unsigned char* __restrict__ selected,
                   const unsigned short* __restrict__ hist, unsigned* __restrict__ cum) {
    const int c = blockIdx.x;
    if (!selected[c]) return;
    const int b = threadIdx.x;
    const int t0 = tfirst[c], t1 = tfirst[c + 1];
    unsigned run = 0;
    size_t row = (size_t)(t0 + c) * RQ_BUCKETS + b;
    for (int t = t0; t < t1; t++) {
        cum[row] = run;
        run += hist[(size_t)t * RQ_BUCKETS + b];
        row += RQ_BUCKETS;
    }
    cum[row] = run;  // row t1 + c: the whole chromosome
}
