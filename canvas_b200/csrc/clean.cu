// cg_clean: CanvasClean's numeric block on the device (reference CanvasClean.cs:474-530).
#include "clean.cuh"
#include "clean_loess.cuh"
#include "clean_stream.cuh"

// ---------------------------------------------------------------------------------------------
// Views for the select engine
// ---------------------------------------------------------------------------------------------
struct SizeView {  // RemoveBigBins: bin sizes, one segment
    const int32_t* start;
    const int32_t* stop;
    long long n;
    __device__ long long size() const { return n; }
    __device__ bool get(long long i, uint32_t& key, int& a, int& b) const {
        key = i32_key(stop[i] - start[i]);
        a = 0;
        b = -1;
        return true;
    }
};

struct GcCountView {  // GetCountsByGC: autosomal alive bins, segment = GC bucket plus the global list
    const float* count;
    const uint8_t* gc;
    const uint8_t* chrom;
    const uint8_t* alive;
    const uint8_t* is_auto;
    const CleanCtl* ctl;
    const int* enabled;
    __device__ long long size() const { return *enabled ? ctl->n2 : 0; }
    __device__ bool get(long long i, uint32_t& key, int& a, int& b) const {
        // every column is loaded before anything is tested: the loads of one element (and of the other elements of
        // the caller's batch) go out together instead of one L2 round trip after the other
        const uint8_t al = alive[i], ch = chrom[i], g = gc[i];
        const float c = count[i];
        const uint8_t au = is_auto[ch];
        key = f32_key(c);
        a = g;
        b = GC_BINS;
        return al != 0 && au != 0;
    }
};

struct WindowView {  // local-SD windows, segment = chromosome of the window's first bin
    const double* wsd;
    const uint8_t* wchrom;
    const CleanCtl* ctl;
    const double* center;  // nullptr: raw value; else |x - center[seg]|
    __device__ long long size() const { return ctl->metric_on ? ctl->n_windows : 0; }
    __device__ bool get(long long i, uint64_t& key, int& a, int& b) const {
        a = wchrom[i];
        b = -1;
        double x = wsd[i];
        if (center) x = fabs(x - center[a]);
        key = f64_key(x);
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------
__global__ void clean_init_kernel(CleanCtl* ctl, int n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CleanCtl z;
        memset(&z, 0, sizeof(z));
        z.n0 = z.n1 = z.n2 = z.n3 = n;
        z.local_sd = -1.0;
        *ctl = z;
    }
}

// input validation (chromosome runs, GC range)
__global__ void clean_validate_kernel(const uint8_t* __restrict__ chrom, const uint8_t* __restrict__ gc,
                                      int n, CleanCtl* ctl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((i > 0 && chrom[i] < chrom[i - 1]) || gc[i] > 100) ctl->unsorted = 1;
}

// size request: rank (int)(0.98 n)  (CanvasClean.cs:338)
__global__ void size_request_kernel(SelState<uint32_t> st, long long k, int on) {
    st.nreq[0] = on ? 1 : 0;
    st.req_k[0] = (unsigned long long)k;
}

__global__ void size_finish_kernel(SelState<uint32_t> st, CleanCtl* ctl, int on) {
    ctl->size_filter_on = on;
    ctl->size_thresh = on ? i32_unkey(st.req_key[0]) : 0x7fffffff;
}

struct SizePred {  // CanvasClean.cs:349-353
    const int32_t* start;
    const int32_t* stop;
    const CleanCtl* ctl;
    __device__ bool operator()(int i) const { return (stop[i] - start[i]) <= ctl->size_thresh; }
};

struct Emit1 {
    const uint8_t* chrom;
    const uint8_t* gc;
    const float* count;
    uint8_t* chrom1;
    uint8_t* gc1;
    float* count1;
    int32_t* orig1;
    __device__ void operator()(int src, int dst) const {
        chrom1[dst] = chrom[src];
        gc1[dst] = gc[src];
        count1[dst] = count[src];
        orig1[dst] = src;
    }
};

// CanvasClean.cs:363-381
__device__ inline bool significantly_different(float a, float b) {
    double mu = __ddiv_rn(__dadd_rn((double)a, (double)b), 2.0);
    if (__fadd_rn(a, b) == 0.0f) return false;
    double da = __dsub_rn((double)a, mu), db = __dsub_rn((double)b, mu);
    double chi2 = __ddiv_rn(__dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db)), mu);
    return chi2 > 6.635;
}

struct OutlierPred {  // CanvasClean.cs:387-413, on the size-filtered list
    const uint8_t* chrom1;
    const float* count1;
    const CleanCtl* ctl;
    int on;
    __device__ bool operator()(int r) const {
        if (!on) return true;
        const int n1 = ctl->n1;
        const bool has_prev = r > 0, has_next = r < n1 - 1;
        const uint8_t c = chrom1[r];
        const bool prev_same = has_prev && chrom1[r - 1] == c;
        const bool next_same = has_next && chrom1[r + 1] == c;
        if ((has_prev && !prev_same) && (has_next && !next_same)) return false;
        const float x = count1[r];
        return (prev_same && !significantly_different(x, count1[r - 1])) ||
               (next_same && !significantly_different(x, count1[r + 1])) || (!has_prev && !has_next);
    }
};

struct Emit2 {
    const uint8_t* chrom1;
    const uint8_t* gc1;
    const float* count1;
    const int32_t* orig1;
    uint8_t* chrom2;
    uint8_t* gc2;
    float* count2;
    int32_t* orig2;
    __device__ void operator()(int src, int dst) const {
        chrom2[dst] = chrom1[src];
        gc2[dst] = gc1[src];
        count2[dst] = count1[src];
        orig2[dst] = orig1[src];
    }
};

// decisions that depend on the length of the filtered list (CanvasClean.cs:483-486)
__global__ void clean_decide_metric_kernel(CleanCtl* ctl, int want_local_sd) {
    const int n2 = ctl->n2;
    ctl->metric_on = (want_local_sd && n2 >= 50000) ? 1 : 0;
    // windows of 20 differences while windowEnd < len(diffs) = n2 - 1  (:284)
    ctl->n_windows = (ctl->metric_on && n2 >= 2) ? (n2 - 2) / LOCAL_SD_WINDOW : 0;
}

// CanvasClean.cs:268-293 — one thread per window; sequential sums as Utilities.Mean/StandardDeviation.
__global__ void local_sd_windows_kernel(const float* __restrict__ count2, const uint8_t* __restrict__ chrom2,
                                        const CleanCtl* __restrict__ ctl, double* __restrict__ wsd,
                                        uint8_t* __restrict__ wchrom, unsigned* __restrict__ wcnt) {
    const int nw = ctl->n_windows;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = w < nw;
    int c = -1;
    if (ok) {
        const int ws = w * LOCAL_SD_WINDOW;
        double d[LOCAL_SD_WINDOW];
        float prev = count2[ws];
        double sum = 0;
#pragma unroll
        for (int t = 0; t < LOCAL_SD_WINDOW; t++) {
            float nxt = count2[ws + t + 1];
            d[t] = (double)__fsub_rn(nxt, prev);
            prev = nxt;
            sum = __dadd_rn(sum, d[t]);
        }
        const double mu = __ddiv_rn(sum, (double)LOCAL_SD_WINDOW);
        double ss = 0;
#pragma unroll
        for (int t = 0; t < LOCAL_SD_WINDOW; t++) {
            double diff = __dsub_rn(d[t], mu);
            ss = __dadd_rn(ss, __dmul_rn(diff, diff));
        }
        wsd[w] = sqrt(__ddiv_rn(ss, (double)(LOCAL_SD_WINDOW - 1)));
        c = chrom2[ws];
        wchrom[w] = (uint8_t)c;
    }
    // windows per chromosome (adjacent windows share the chromosome: aggregate per warp)
    unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
        unsigned m = __match_any_sync(act, c);
        if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&wcnt[c], (unsigned)__popc(m));
    }
}

// Median and MAD of one chromosome's window SDs (CanvasClean.cs:243-258) in one CTA: the list (<= LSD_SORT_CAP values)
// is staged in shared memory and the order statistics of SortedList<double>.Median() are radix-selected there, first
// on the values, then on the deviations |x - median|.  Replaces two 8-pass grid-wide selections over ~N/20 values.
constexpr int LSD_SORT_CAP = 16384;

// k-th smallest (0-based) of k[0..n) in shared memory: MSD radix select, 8-bit digits, block-wide
__device__ unsigned long long block_radix_select_u64(const unsigned long long* k, int n, int rank) {
    __shared__ unsigned s_h[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_rank;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = rank; }
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) s_h[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = k[i];
            if (shift == 56 || ((key ^ prefix) >> (shift + 8)) == 0ull) atomicAdd(&s_h[(unsigned)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            // every lane reads the rank before the one owning lane rewrites it below (racecheck: the shuffles in between
            // are the warp-level barrier that orders the two)
            const unsigned r = (unsigned)s_rank;
            unsigned c[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { c[t] = s_h[lane * 8 + t]; sum += c[t]; }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const unsigned excl = incl - sum;
            __syncwarp();
            if (r >= excl && r < excl + sum) {  // exactly one lane
                unsigned run = excl;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (r < run + c[t]) { s_prefix = prefix | ((unsigned long long)(lane * 8 + t) << shift); s_rank = (int)(r - run); break; }
                    run += c[t];
                }
            }
        }
        __syncthreads();
    }
    const unsigned long long out = s_prefix;
    __syncthreads();
    return out;
}

// SortedList<double>.Median() of the keys in shared memory
__device__ double block_median_u64(const unsigned long long* k, int n) {
    const unsigned long long kb = block_radix_select_u64(k, n, n / 2);
    const unsigned long long ka = (n & 1) ? kb : block_radix_select_u64(k, n, n / 2 - 1);
    const double a = f64_unkey(ka), b = f64_unkey(kb);
    return ka == kb ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
}

__global__ void __launch_bounds__(1024) local_sd_mad_kernel(const double* __restrict__ wsd, const unsigned* __restrict__ wcnt,
                                                            const CleanCtl* __restrict__ ctl, double* __restrict__ wmed,
                                                            double* __restrict__ wmad) {
    extern __shared__ unsigned long long lsd_k[];
    if (!ctl->metric_on) return;
    const int c = blockIdx.x;
    const int n = (int)wcnt[c];
    if (n == 0 || n > LSD_SORT_CAP) {  // the host only takes this path when no chromosome can exceed the capacity
        if (threadIdx.x == 0) { wmed[c] = 0.0; wmad[c] = 0.0; }
        return;
    }
    long long first = 0;
    for (int q = 0; q < c; q++) first += wcnt[q];
    for (int i = threadIdx.x; i < n; i += blockDim.x) lsd_k[i] = f64_key(wsd[first + i]);
    __syncthreads();
    const double med = block_median_u64(lsd_k, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) lsd_k[i] = f64_key(fabs(f64_unkey(lsd_k[i]) - med));
    __syncthreads();
    const double mad = block_median_u64(lsd_k, n);
    if (threadIdx.x == 0) { wmed[c] = med; wmad[c] = mad; }
}

// median-style requests {lower middle, upper middle} for every segment with cnt > 0
__global__ void median_request_u64_kernel(SelState<uint64_t> st, const unsigned* __restrict__ cnt, const int* on) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    unsigned n = *on ? cnt[s] : 0u;
    if (n == 0) { st.nreq[s] = 0; return; }
    st.nreq[s] = 2;
    st.req_k[s * SEL_G + 0] = (n & 1u) ? n / 2 : n / 2 - 1;
    st.req_k[s * SEL_G + 1] = n / 2;
}

// SortedList<double>.Median(): mean of the middles
__global__ void median_finish_u64_kernel(SelState<uint64_t> st, double* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    if (st.nreq[s] == 0) { out[s] = 0.0; return; }
    double a = f64_unkey(st.req_key[s * SEL_G + 0]), b = f64_unkey(st.req_key[s * SEL_G + 1]);
    out[s] = st.req_key[s * SEL_G + 0] == st.req_key[s * SEL_G + 1] ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
}

// CanvasClean.cs:243-258 — average of the per-chromosome MADs, in chromosome order
__global__ void local_sd_average_kernel(CleanCtl* ctl, const unsigned* __restrict__ wcnt,
                                        const double* __restrict__ mad, int n_chrom) {
    if (!ctl->metric_on) return;
    double s = 0;
    int k = 0;
    for (int c = 0; c < n_chrom; c++)
        if (wcnt[c] > 0) { s = __dadd_rn(s, mad[c]); k++; }
    ctl->local_sd = k > 0 ? __ddiv_rn(s, (double)k) : __longlong_as_double(0x7ff8000000000000ll);
}

// CanvasClean.cs:213-224 — GC histograms of the outlier-filtered list
__global__ void gc_hist_kernel(const uint8_t* __restrict__ gc2, const uint8_t* __restrict__ chrom2,
                               const uint8_t* __restrict__ is_auto, CleanCtl* ctl) {
    __shared__ unsigned h_auto[GC_BINS], h_all[GC_BINS];
    for (int t = threadIdx.x; t < GC_BINS; t += blockDim.x) h_auto[t] = h_all[t] = 0u;
    __syncthreads();
    const int n2 = ctl->n2;
    const int n_round = ((n2 + 31) / 32) * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool ok = i < n2;
        int g = ok ? gc2[i] : 0;
        bool au = ok && is_auto[chrom2[i]];
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned m = __match_any_sync(act, g);
            unsigned ma = __ballot_sync(m, au) & m;
            if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) {
                atomicAdd(&h_all[g], (unsigned)__popc(m));
                if (ma) atomicAdd(&h_auto[g], (unsigned)__popc(ma));
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < GC_BINS; t += blockDim.x) {
        if (h_all[t]) atomicAdd(&ctl->hist_all[t], h_all[t]);
        if (h_auto[t]) atomicAdd(&ctl->hist_auto[t], h_auto[t]);
    }
}

// CanvasClean.cs:226-227 and the bookkeeping that follows from it
__global__ void gc_threshold_kernel(CleanCtl* ctl, int gc_norm, int median_mode, int min_bins_weighted) {
    unsigned total = 0;
    for (int g = 0; g < GC_BINS; g++) total += ctl->hist_auto[g];
    ctl->n_auto = total;
    int thr = 0;
    if (gc_norm && median_mode) {
        int avg = max(min_bins_weighted, (int)((double)total / (double)GC_BINS));
        thr = min(MIN_BINS_PER_GC, avg);
    } else {
        thr = -0x7fffffff;  // no GC filter: every bin stays
    }
    long long alive = 0, alive_auto = 0;
    for (int g = 0; g < GC_BINS; g++) {
        if ((long long)ctl->hist_auto[g] >= (long long)thr) {
            alive += ctl->hist_all[g];
            alive_auto += ctl->hist_auto[g];
        }
    }
    ctl->do_norm = gc_norm;
    if (gc_norm && alive == 0) {  // :502-505 — keep the unfiltered list and skip normalisation
        ctl->gc_skipped = 1;
        ctl->do_norm = 0;
        thr = -0x7fffffff;
        alive = ctl->n2;
        alive_auto = total;
    }
    ctl->gc_thresh = thr;
    ctl->n3 = (int)alive;
    ctl->n_auto3 = (unsigned)alive_auto;
    ctl->do_variance = (ctl->do_norm && ctl->metric_on && alive > 500000) ? 1 : 0;  // :512
}

__global__ void gc_alive_kernel(const uint8_t* __restrict__ gc2, const CleanCtl* __restrict__ ctl,
                                uint8_t* __restrict__ alive) {
    const int n2 = ctl->n2;
    const int thr = ctl->gc_thresh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
        alive[i] = (long long)ctl->hist_auto[gc2[i]] >= (long long)thr ? 1 : 0;
}

__device__ inline unsigned alive_auto_count(const CleanCtl* ctl, int g) {
    return (long long)ctl->hist_auto[g] >= (long long)ctl->gc_thresh ? ctl->hist_auto[g] : 0u;
}

// NormalizeByGC requests (CanvasClean.cs:172-187): exact median per bucket with >= 100 bins + global
__global__ void gc_median_request_kernel(SelState<uint32_t> st, CleanCtl* ctl, const int* enabled) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= GC_SEGS) return;
    st.nreq[s] = 0;
    if (!*enabled) return;
    unsigned n = s < GC_BINS ? alive_auto_count(ctl, s) : ctl->n_auto3;
    if (s < GC_BINS && n < (unsigned)MIN_BINS_PER_GC) {
        return;  // weighted median of the neighbouring buckets: gc_weighted_kernel
    }
    if (n == 0) return;
    st.nreq[s] = 2;
    st.req_k[s * SEL_G + 0] = (n & 1u) ? n / 2 : n / 2 - 1;
    st.req_k[s * SEL_G + 1] = n / 2;
}

// SortedList<float>.Median(): mean of the middles in single precision, widened to double
__global__ void gc_median_finish_kernel(SelState<uint32_t> st, CleanCtl* ctl, const int* enabled) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= GC_SEGS || !*enabled) return;
    double m = 0.0;
    if (st.nreq[s] == 2) {
        float a = f32_unkey(st.req_key[s * SEL_G + 0]), b = f32_unkey(st.req_key[s * SEL_G + 1]);
        unsigned n = s < GC_BINS ? alive_auto_count(ctl, s) : ctl->n_auto3;
        m = (n & 1u) ? (double)a : (double)__fdiv_rn(__fadd_rn(a, b), 2.0f);
    }
    if (s < GC_BINS) ctl->med[s] = m;
    else ctl->global_median = m;
}

// Quartile ranks of Utilities.Quartiles (Utilities.cs:361-419) as six requests
// {Q1a, Q1b, Q2a, Q2b, Q3a, Q3b}
__device__ inline void quartile_ranks(unsigned n, unsigned long long* k) {
    const unsigned mid = n / 2;
    if ((n & 1u) == 0) {
        k[2] = mid - 1; k[3] = mid;
        const unsigned mm = mid / 2;
        if ((mid & 1u) == 0) { k[0] = mm - 1; k[1] = mm; k[4] = mid + mm - 1; k[5] = mid + mm; }
        else { k[0] = k[1] = mm; k[4] = k[5] = mm + mid; }
    } else {
        k[2] = k[3] = mid;
        if ((n - 1) % 4 == 0) { const unsigned q = (n - 1) / 4; k[0] = q - 1; k[1] = q; k[4] = 3 * q; k[5] = 3 * q + 1; }
        else { const unsigned q = (n - 3) / 4; k[0] = q; k[1] = q + 1; k[4] = 3 * q + 1; k[5] = 3 * q + 2; }
    }
}

__device__ inline void quartile_values(unsigned n, const float* v, float* q) {
    const unsigned mid = n / 2;
    if ((n & 1u) == 0) {
        q[1] = __fdiv_rn(__fadd_rn(v[2], v[3]), 2.0f);
        if ((mid & 1u) == 0) {
            q[0] = __fdiv_rn(__fadd_rn(v[0], v[1]), 2.0f);
            q[2] = __fdiv_rn(__fadd_rn(v[4], v[5]), 2.0f);
        } else { q[0] = v[0]; q[2] = v[4]; }
    } else {
        q[1] = v[2];
        if ((n - 1) % 4 == 0) {
            q[0] = __fadd_rn(__fmul_rn(v[0], 0.25f), __fmul_rn(v[1], 0.75f));
            q[2] = __fadd_rn(__fmul_rn(v[4], 0.75f), __fmul_rn(v[5], 0.25f));
        } else {
            q[0] = __fadd_rn(__fmul_rn(v[0], 0.75f), __fmul_rn(v[1], 0.25f));
            q[2] = __fadd_rn(__fmul_rn(v[4], 0.25f), __fmul_rn(v[5], 0.75f));
        }
    }
}

__global__ void gc_quartile_request_kernel(SelState<uint32_t> st, CleanCtl* ctl, const int* enabled) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= GC_SEGS) return;
    st.nreq[s] = 0;
    if (!*enabled) return;
    unsigned n = s < GC_BINS ? alive_auto_count(ctl, s) : ctl->n_auto3;
    if (s < GC_BINS && n > 0 && n < (unsigned)MIN_BINS_PER_GC) return;  // weighted quantiles: gc_weighted_kernel
    if (n < 2) return;
    st.nreq[s] = 6;
    quartile_ranks(n, st.req_k + s * SEL_G);
}

__global__ void gc_quartile_finish_kernel(SelState<uint32_t> st, CleanCtl* ctl, const int* enabled) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= GC_SEGS || !*enabled) return;
    unsigned n = s < GC_BINS ? alive_auto_count(ctl, s) : ctl->n_auto3;
    float q[3] = {0.f, 0.f, 0.f};
    if (st.nreq[s] == 6) {
        float v[6];
        for (int r = 0; r < 6; r++) v[r] = f32_unkey(st.req_key[s * SEL_G + r]);
        quartile_values(n, v, q);
    }
    if (s < GC_BINS) {
        if (n == 0) { ctl->iqr[s] = -1.f; ctl->q2[s] = -1.f; }   // CanvasClean.cs:52-56
        else { ctl->iqr[s] = __fsub_rn(q[2], q[0]); ctl->q2[s] = q[1]; }
    } else {
        ctl->global_q[0] = q[0]; ctl->global_q[1] = q[1]; ctl->global_q[2] = q[2];
        ctl->global_iqr = __fsub_rn(q[2], q[0]);
    }
}

// GC buckets with fewer than 100 autosomal bins borrow their neighbours (GetWeightedCounts, CanvasClean.cs:107-132:
// buckets gc +- r with weight 2^-r until 100 values are collected) and take weighted quantiles
// (Utilities.WeightedQuantiles, Utilities.cs:493-515: the last value, in stable value order, whose cumulative
// weight / total weight is <= p).  One CTA per such bucket gathers the (value, radius) pairs, sorts them in
// shared memory and walks them in order.  Weights are powers of two, so the running sums are exact.
constexpr int WQ_CAP = 16384;

__global__ void __launch_bounds__(256) gc_weighted_kernel(const float* __restrict__ count2, const uint8_t* __restrict__ gc2,
                                                          const uint8_t* __restrict__ chrom2, const uint8_t* __restrict__ alive,
                                                          const uint8_t* __restrict__ is_auto, CleanCtl* ctl, int mode,
                                                          const int* enabled) {
    extern __shared__ unsigned long long wq_key[];
    __shared__ signed char s_rad[GC_BINS];
    __shared__ int s_total, s_cnt, s_go;
    if (!*enabled) return;
    const int g = blockIdx.x, t = threadIdx.x;
    const unsigned ng = alive_auto_count(ctl, g);
    bool needed;
    if (mode == 0) needed = ng < (unsigned)MIN_BINS_PER_GC && (long long)ctl->hist_auto[g] >= (long long)ctl->gc_thresh && ctl->hist_all[g] > 0;
    else needed = ng > 0 && ng < (unsigned)MIN_BINS_PER_GC;
    if (!needed) return;
    if (t < GC_BINS) s_rad[t] = -1;
    __syncthreads();
    if (t == 0) {
        long long total = 0;
        int r = 0;
        while (total < MIN_BINS_PER_GC) {
            const int hi = g + r, lo = g - r;
            if (hi >= GC_BINS && lo < 0) break;
            if (hi < GC_BINS) { s_rad[hi] = (signed char)r; total += alive_auto_count(ctl, hi); }
            if (lo != hi && lo >= 0) { s_rad[lo] = (signed char)r; total += alive_auto_count(ctl, lo); }
            r++;
        }
        s_total = (int)min(total, (long long)0x7fffffff);
        s_cnt = 0;
        s_go = total <= WQ_CAP ? 1 : 0;
        if (!s_go) ctl->wq_big[g] = 1;  // more neighbours than the shared-memory sorter holds: gc_weighted_big_kernel takes the bucket
    }
    __syncthreads();
    if (!s_go) return;
    const int T = s_total;
    int pad = 1;
    while (pad < T) pad <<= 1;
    for (int k = t; k < pad; k += blockDim.x) wq_key[k] = ~0ull;
    __syncthreads();
    const int n2 = ctl->n2;
    for (int i = t; i < n2; i += blockDim.x) {
        const int r = s_rad[gc2[i]];
        if (r >= 0 && alive[i] && is_auto[chrom2[i]]) {
            const int slot = atomicAdd(&s_cnt, 1);
            if (slot < pad) wq_key[slot] = ((unsigned long long)f32_key(count2[i]) << 8) | (unsigned long long)r;
        }
    }
    __syncthreads();
    for (int k = 2; k <= pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < pad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = wq_key[i], b = wq_key[l];
                    if ((a > b) == up) { wq_key[i] = b; wq_key[l] = a; }
                }
            }
            __syncthreads();
        }
    if (t == 0) {
        double acc = 0.0;
        for (int i = 0; i < T; i++) acc += ldexp(1.0, -(int)(wq_key[i] & 0xffull));
        const double total_w = (double)(float)acc;  // Enumerable.Sum over float weights returns a float
        const double probs[3] = {0.25, 0.5, 0.75};
        double q[3] = {0.0, 0.0, 0.0};
        double cw = 0.0;
        for (int i = 0; i < T; i++) {
            cw += ldexp(1.0, -(int)(wq_key[i] & 0xffull));
            const double cp = cw / total_w;
            const double v = (double)f32_unkey((uint32_t)(wq_key[i] >> 8));
            for (int p = 0; p < 3; p++)
                if (cp <= probs[p]) q[p] = v;
        }
        if (mode == 0) ctl->med[g] = q[1];
        else { ctl->q2[g] = (float)q[1]; ctl->iqr[g] = (float)(q[2] - q[0]); }
    }
}

// The same for the buckets gc_weighted_kernel left (one neighbouring bucket alone can hold tens of thousands of bins, and
// the reference collects all of them, CanvasClean.cs:107-132): one CTA works through them one after the other with the keys
// in global scratch — block-wide bitonic sort, then the cumulative weights as a block scan.  Sums of the weights 2^-r are
// exact in double whatever their order, so the scan gives the reference's running sums; the reference's loop keeps the
// LAST value whose cumulative share is <= p, and since the shares ascend that is the largest index that satisfies it.
__global__ void __launch_bounds__(1024) gc_weighted_big_kernel(const float* __restrict__ count2, const uint8_t* __restrict__ gc2,
                                                               const uint8_t* __restrict__ chrom2, const uint8_t* __restrict__ alive,
                                                               const uint8_t* __restrict__ is_auto, CleanCtl* ctl, int mode, const int* enabled,
                                                               unsigned long long* __restrict__ key, long long key_cap) {
    __shared__ signed char s_rad[GC_BINS];
    __shared__ int s_total, s_cnt;
    __shared__ double s_w[32];
    __shared__ int s_i[3][32];
    if (!*enabled) return;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (int g = 0; g < GC_BINS; g++) {
        if (!ctl->wq_big[g]) continue;  // uniform
        __syncthreads();
        if (t < GC_BINS) s_rad[t] = -1;
        __syncthreads();
        if (t == 0) {
            long long total = 0;
            int r = 0;
            while (total < MIN_BINS_PER_GC) {
                const int hi = g + r, lo = g - r;
                if (hi >= GC_BINS && lo < 0) break;
                if (hi < GC_BINS) { s_rad[hi] = (signed char)r; total += alive_auto_count(ctl, hi); }
                if (lo != hi && lo >= 0) { s_rad[lo] = (signed char)r; total += alive_auto_count(ctl, lo); }
                r++;
            }
            s_total = (int)min(total, (long long)0x7fffffff);
            s_cnt = 0;
        }
        __syncthreads();
        const int T = s_total;
        long long pad = 1;
        while (pad < T) pad <<= 1;
        if (pad > key_cap) {  // cannot happen: the scratch holds every bin
            if (t == 0) { ctl->need_weighted = 1; ctl->wq_big[g] = 0; }
            continue;
        }
        for (long long k = t; k < pad; k += blockDim.x) key[k] = ~0ull;
        __syncthreads();
        const int n2 = ctl->n2;
        for (int i = t; i < n2; i += blockDim.x) {
            const int r = s_rad[gc2[i]];
            if (r >= 0 && alive[i] && is_auto[chrom2[i]]) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < pad) key[slot] = ((unsigned long long)f32_key(count2[i]) << 8) | (unsigned long long)r;
            }
        }
        __syncthreads();
        for (long long k = 2; k <= pad; k <<= 1)
            for (long long j = k >> 1; j > 0; j >>= 1) {
                for (long long i = t; i < pad; i += blockDim.x) {
                    const long long l = i ^ j;
                    if (l > i) {
                        const bool up = (i & k) == 0;
                        const unsigned long long a = key[i], b = key[l];
                        if ((a > b) == up) { key[i] = b; key[l] = a; }
                    }
                }
                __syncthreads();
            }
        // cumulative weights: a contiguous chunk per thread, block scan of the chunk sums
        const int chunk = (T + (int)blockDim.x - 1) / (int)blockDim.x;
        const int c0 = min(T, t * chunk), c1 = min(T, c0 + chunk);
        double mine = 0.0;
        for (int i = c0; i < c1; i++) mine += ldexp(1.0, -(int)(key[i] & 0xffull));
        double incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_w[wid] = incl;
        __syncthreads();
        double before = incl - mine, all = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { if (k < wid) before += s_w[k]; all += s_w[k]; }
        const double total_w = (double)(float)all;  // Enumerable.Sum over float weights returns a float
        const double probs[3] = {0.25, 0.5, 0.75};
        int last[3] = {-1, -1, -1};
        double cw = before;
        for (int i = c0; i < c1; i++) {
            cw += ldexp(1.0, -(int)(key[i] & 0xffull));
            const double cp = cw / total_w;
#pragma unroll
            for (int p = 0; p < 3; p++)
                if (cp <= probs[p]) last[p] = i;
        }
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int m = (int)__reduce_max_sync(0xffffffffu, (unsigned)(last[p] + 1));
            if (lane == 0) s_i[p][wid] = m;
        }
        __syncthreads();
        if (t == 0) {
            double q[3];
            for (int p = 0; p < 3; p++) {
                int m = 0;
                for (int k = 0; k < (int)(blockDim.x >> 5); k++) m = max(m, s_i[p][k]);
                q[p] = m > 0 ? (double)f32_unkey((uint32_t)(key[m - 1] >> 8)) : 0.0;
            }
            if (mode == 0) ctl->med[g] = q[1];
            else { ctl->q2[g] = (float)q[1]; ctl->iqr[g] = (float)(q[2] - q[0]); }
            ctl->wq_big[g] = 0;
        }
        __syncthreads();
    }
}

// CanvasClean.cs:71-82
__global__ void variance_decide_kernel(CleanCtl* ctl) {
    if (!ctl->do_variance) { ctl->variance_fired = 0; return; }
    int significant = 0;
    for (int g = 10; g < 90; g++)
        if (__fmul_rn(ctl->global_iqr, 2.f) < ctl->iqr[g]) significant++;
    ctl->variance_fired = significant > 0 ? 1 : 0;
}

// CanvasClean.cs:85-94 (single precision throughout)
__global__ void variance_apply_kernel(float* __restrict__ count2, const uint8_t* __restrict__ gc2,
                                      const uint8_t* __restrict__ alive, const CleanCtl* __restrict__ ctl) {
    if (!ctl->variance_fired) return;
    __shared__ float s_iqr[GC_BINS], s_q2[GC_BINS];
    for (int t = threadIdx.x; t < GC_BINS; t += blockDim.x) { s_iqr[t] = ctl->iqr[t]; s_q2[t] = ctl->q2[t]; }
    __syncthreads();
    const float giqr = ctl->global_iqr;
    const int n2 = ctl->n2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        if (!alive[i]) continue;
        const int g = gc2[i];
        const float scaled = __fmul_rn(s_iqr[g], 0.8f);
        if (giqr >= scaled) continue;
        const float ratio = __fdiv_rn(scaled, giqr);
        const float m = s_q2[g];
        count2[i] = __fadd_rn(m, __fdiv_rn(__fsub_rn(count2[i], m), ratio));
    }
}

struct FinalPred {  // GC filter survivors minus RemoveBinsWithExtremeLocalSD (CanvasClean.cs:308-322)
    const uint8_t* alive;
    const double* wsd;
    const CleanCtl* ctl;
    __device__ bool operator()(int i) const {
        if (!alive[i]) return false;
        if (!ctl->metric_on) return true;
        const int w = i / LOCAL_SD_WINDOW;
        const double dev = w < ctl->n_windows ? wsd[w] : -1.0;
        return !(dev > 20.0 * 2.0 && ctl->local_sd > 5.0);
    }
};

struct EmitOut {
    const float* count2;
    const int32_t* orig2;
    int32_t* kept;
    float* count_out;
    __device__ void operator()(int src, int dst) const {
        kept[dst] = orig2[src];
        count_out[dst] = count2[src];
    }
};

// ---------------------------------------------------------------------------------------------
// Device-side pipeline (shared by cg_clean and cg_clean_partition_wavelet)
// ---------------------------------------------------------------------------------------------
static long long wq_scratch_keys(int64_t n) {  // a power of two that holds every bin
    long long p = 1;
    while (p < n) p <<= 1;
    return p;
}

size_t clean_workspace_bytes(int64_t n, int n_chrom, bool loess) {
    size_t s = 0;
    s += arena_need(n, 1) * 2 + arena_need(n, 4) * 3;          // inputs
    s += arena_need(n, 1) * 2 + arena_need(n, 4) * 2;          // L1
    s += arena_need(n, 1) * 3 + arena_need(n, 4) * 2;          // L2 + alive
    s += arena_need(n / LOCAL_SD_WINDOW + 2, 8) + arena_need(n / LOCAL_SD_WINDOW + 2, 1);
    s += arena_need(n, 4) * 2;                                  // outputs
    s += arena_need(n / CMP_TILE + 2, 4);
    s += arena_need(wq_scratch_keys(n), 8);                     // weighted quantiles beyond the shared-memory sorter
    s += arena_need(1, sizeof(CleanCtl)) + arena_need(256, 1) * 2 + arena_need(256, 4) + arena_need(256, 8) * 2;
    s += sel_state_bytes<uint32_t>(1) + sel_state_bytes<uint32_t>(GC_SEGS) + sel_state_bytes<uint64_t>(std::max(n_chrom, 1));
    if (loess) s += loess_workspace_bytes(n);
    return s + (1 << 16);
}

int clean_alloc(cg_ctx* ctx, int64_t n, int n_chrom, CleanDev& d, bool loess) {
    d.n = n;
    d.n_chrom = n_chrom;
    d.chrom = arena_take<uint8_t>(ctx, n);
    d.gc = arena_take<uint8_t>(ctx, n);
    d.start = arena_take<int32_t>(ctx, n);
    d.stop = arena_take<int32_t>(ctx, n);
    d.count = arena_take<float>(ctx, n);
    d.chrom1 = arena_take<uint8_t>(ctx, n);
    d.gc1 = arena_take<uint8_t>(ctx, n);
    d.count1 = arena_take<float>(ctx, n);
    d.orig1 = arena_take<int32_t>(ctx, n);
    d.chrom2 = arena_take<uint8_t>(ctx, n);
    d.gc2 = arena_take<uint8_t>(ctx, n);
    d.alive = arena_take<uint8_t>(ctx, n);
    d.count2 = arena_take<float>(ctx, n);
    d.orig2 = arena_take<int32_t>(ctx, n);
    d.wsd = arena_take<double>(ctx, n / LOCAL_SD_WINDOW + 2);
    d.wchrom = arena_take<uint8_t>(ctx, n / LOCAL_SD_WINDOW + 2);
    d.kept = arena_take<int32_t>(ctx, n);
    d.count_out = arena_take<float>(ctx, n);
    d.tiles = arena_take<int>(ctx, n / CMP_TILE + 2);
    d.wq_cap = wq_scratch_keys(n);
    d.wq_key = arena_take<unsigned long long>(ctx, (size_t)d.wq_cap);
    d.ctl = arena_take<CleanCtl>(ctx, 1);
    d.is_auto = arena_take<uint8_t>(ctx, 256);
    d.is_chry = arena_take<uint8_t>(ctx, 256);
    d.wcnt = arena_take<unsigned>(ctx, 256);
    d.wmed = arena_take<double>(ctx, 256);
    d.wmad = arena_take<double>(ctx, 256);
    bool ok = d.chrom && d.gc && d.start && d.stop && d.count && d.chrom1 && d.gc1 && d.count1 && d.orig1 &&
              d.chrom2 && d.gc2 && d.alive && d.count2 && d.orig2 && d.wsd && d.wchrom && d.kept &&
              d.count_out && d.tiles && d.wq_key && d.ctl && d.is_auto && d.is_chry && d.wcnt && d.wmed && d.wmad;
    ok = ok && sel_state_alloc<uint32_t>(ctx, 1, d.sel_size) && sel_state_alloc<uint32_t>(ctx, GC_SEGS, d.sel_gc) &&
         sel_state_alloc<uint64_t>(ctx, std::max(n_chrom, 1), d.sel_win);
    if (ok && loess) {
        d.lo = std::make_shared<LoessDev>();
        ok = loess_alloc(ctx, n, *d.lo);
        d.lo->is_chry = d.is_chry;
    }
    return ok ? CG_OK : cg_fail(ctx, CG_ERR_CUDA, "clean: device arena exhausted");
}

// Enqueue the whole CanvasClean pipeline on ctx->stream.  Inputs must already be in d.chrom/...;
// results end in d.kept / d.count_out / d.ctl.
static int clean_enqueue_body(cg_ctx* ctx, const cg_clean_opts* o, CleanDev& d);

// The Clean pipeline is ~90 short launches whose arguments depend only on the problem shape (n, options, arena
// placement): every data-dependent decision is taken on the device.  The sequence is therefore captured once per
// shape into a CUDA graph and replayed — one launch instead of ninety, no host-side gaps between the kernels.
int clean_enqueue(cg_ctx* ctx, const cg_clean_opts* o, CleanDev& d) {
    cudaStream_t s = ctx->stream;
    cudaEventRecord(ctx->stage_ev[0], s);
    ctx->stage_used[0] = true;
    static const bool no_graph = getenv("CANVAS_NO_GRAPH") != nullptr;
    int rc = CG_OK;
    if (no_graph || ctx->tl) {
        rc = clean_enqueue_body(ctx, o, d);
    } else {
        CgGraphEntry want{};
        const long long key[12] = {(long long)(uintptr_t)ctx->arena, (long long)(uintptr_t)d.ctl, d.n, d.n_chrom, d.max_chrom_bins,
                                   o->size_filter, o->outlier_filter, o->gc_norm, o->gc_mode, o->want_local_sd, o->min_bins_per_gc,
                                   (long long)(uintptr_t)d.count};
        for (int i = 0; i < 12; i++) want.key[i] = key[i];
        const CgGraphEntry* hit = nullptr;
        for (const auto& g : ctx->clean_graphs)
            if (!memcmp(g.key, want.key, sizeof(want.key))) { hit = &g; break; }
        if (!hit) {
            const int launches_before = ctx->launches;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                rc = clean_enqueue_body(ctx, o, d);  // capture unavailable: plain launches
                cudaEventRecord(ctx->stage_ev[1], s);
                return rc;
            }
            rc = clean_enqueue_body(ctx, o, d);
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            want.launches = ctx->launches - launches_before;
            ctx->launches = launches_before;
            if (rc != CG_OK || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return rc != CG_OK ? rc : cg_fail(ctx, CG_ERR_CUDA, std::string("clean: graph capture failed: ") + cudaGetErrorString(ce));
            }
            const cudaError_t ie = cudaGraphInstantiate(&want.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, std::string("clean: graph instantiation failed: ") + cudaGetErrorString(ie));
            if (ctx->clean_graphs.size() >= 16) cg_graphs_clear(ctx);
            ctx->clean_graphs.push_back(want);
            hit = &ctx->clean_graphs.back();
        }
        const cudaError_t le = cudaGraphLaunch(hit->exec, s);
        if (le != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, std::string("clean: graph launch failed: ") + cudaGetErrorString(le));
        ctx->launches += hit->launches;
    }
    cudaEventRecord(ctx->stage_ev[1], s);
    return rc;
}

static int clean_enqueue_body(cg_ctx* ctx, const cg_clean_opts* o, CleanDev& d) {
    const int n = (int)d.n;
    const int nb = std::max(1, div_up(n, 256));
    const int grid_stream = std::max(1, std::min(nb, ctx->num_sms * 8));
    const int grid_k8 = std::max(1, std::min(div_up(n, K8_TILE), ctx->num_sms * K8_CTAS_PER_SM));
    cudaFuncSetAttribute(normalize_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k8_smem_bytes());
    CleanCtl* ctl = d.ctl;
    const bool loess = o->gc_norm && o->gc_mode != 0;
    if (loess && !d.lo) return cg_fail(ctx, CG_ERR_ARG, "clean: LOESS buffers were not allocated");

    cudaFuncSetAttribute(gc_weighted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WQ_CAP * 8);
    CG_LAUNCH(ctx, clean_init_kernel, 1, 32, 0, ctl, n);
    cudaMemsetAsync(d.wcnt, 0, 256 * sizeof(unsigned), ctx->stream);
    cudaMemsetAsync(d.sel_size.hist, 0, (size_t)1 * SEL_G * SEL_BINS * sizeof(unsigned), ctx->stream);
    cudaMemsetAsync(d.sel_gc.hist, 0, (size_t)GC_SEGS * SEL_G * SEL_BINS * sizeof(unsigned), ctx->stream);
    cudaMemsetAsync(d.sel_win.hist, 0, (size_t)d.sel_win.nseg * SEL_G * SEL_BINS * sizeof(unsigned), ctx->stream);
    CG_LAUNCH(ctx, clean_validate_kernel, nb, 256, 0, d.chrom, d.gc, n, ctl);

    // --- RemoveBigBins (:328-355)
    long long k98 = (long long)(0.98 * (double)n);
    int size_on = (o->size_filter && k98 < n) ? 1 : 0;
    CG_LAUNCH(ctx, size_request_kernel, 1, 1, 0, d.sel_size, k98, size_on);
    if (size_on) {
        SizeView sv{d.start, d.stop, n};
        sel_run_scatter<uint32_t, SizeView>(ctx, sv, d.sel_size, n);
    }
    CG_LAUNCH(ctx, size_finish_kernel, 1, 1, 0, d.sel_size, ctl, size_on);
    {
        SizePred p{d.start, d.stop, ctl};
        Emit1 e{d.chrom, d.gc, d.count, d.chrom1, d.gc1, d.count1, d.orig1};
        compact_run(ctx, p, e, &ctl->n0, n, d.tiles, &ctl->n1);
    }
    CG_TL(ctx, "size filter");
    // --- RemoveOutliers (:387-413)
    {
        OutlierPred p{d.chrom1, d.count1, ctl, o->outlier_filter};
        Emit2 e{d.chrom1, d.gc1, d.count1, d.orig1, d.chrom2, d.gc2, d.count2, d.orig2};
        compact_run(ctx, p, e, &ctl->n1, n, d.tiles, &ctl->n2);
    }
    CG_TL(ctx, "outlier filter");
    // --- local SD metric (:483-494)
    CG_LAUNCH(ctx, clean_decide_metric_kernel, 1, 1, 0, ctl, o->want_local_sd);
    if (o->want_local_sd && n >= 50000) {
        const int nwin_upper = n / LOCAL_SD_WINDOW + 1;
        CG_LAUNCH(ctx, local_sd_windows_kernel, div_up(nwin_upper, 128), 128, 0, d.count2, d.chrom2, ctl, d.wsd,
                  d.wchrom, d.wcnt);
        if (d.max_chrom_bins >= 0 && d.max_chrom_bins / LOCAL_SD_WINDOW + 2 <= LSD_SORT_CAP && d.n_chrom > 0) {
            // every chromosome's window list fits one CTA's shared memory: sort there
            cudaFuncSetAttribute(local_sd_mad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSD_SORT_CAP * 8);
            CG_LAUNCH(ctx, local_sd_mad_kernel, d.n_chrom, 1024, LSD_SORT_CAP * 8, d.wsd, d.wcnt, ctl, d.wmed, d.wmad);
        } else {
            WindowView wv{d.wsd, d.wchrom, ctl, nullptr};
            CG_LAUNCH(ctx, median_request_u64_kernel, div_up(d.sel_win.nseg, 128), 128, 0, d.sel_win, d.wcnt, &ctl->metric_on);
            sel_run_scatter<uint64_t, WindowView>(ctx, wv, d.sel_win, nwin_upper);
            CG_LAUNCH(ctx, median_finish_u64_kernel, div_up(d.sel_win.nseg, 128), 128, 0, d.sel_win, d.wmed);
            WindowView wv2{d.wsd, d.wchrom, ctl, d.wmed};
            CG_LAUNCH(ctx, median_request_u64_kernel, div_up(d.sel_win.nseg, 128), 128, 0, d.sel_win, d.wcnt, &ctl->metric_on);
            sel_run_scatter<uint64_t, WindowView>(ctx, wv2, d.sel_win, nwin_upper);
            CG_LAUNCH(ctx, median_finish_u64_kernel, div_up(d.sel_win.nseg, 128), 128, 0, d.sel_win, d.wmad);
        }
        CG_LAUNCH(ctx, local_sd_average_kernel, 1, 1, 0, ctl, d.wcnt, d.wmad, d.n_chrom);
    }
    CG_TL(ctx, "local sd");
    // --- RemoveBinsWithExtremeGC (:207-237) as an alive mask over the outlier-filtered list
    CG_LAUNCH(ctx, gc_hist_kernel, grid_stream, 256, 0, d.gc2, d.chrom2, d.is_auto, ctl);
    CG_LAUNCH(ctx, gc_threshold_kernel, 1, 1, 0, ctl, o->gc_norm, o->gc_mode == 0, o->min_bins_per_gc);
    CG_LAUNCH(ctx, gc_alive_kernel, grid_stream, 256, 0, d.gc2, ctl, d.alive);
    CG_TL(ctx, "gc filter");
    if (o->gc_norm) {
        if (loess) {
            // --- LoessGCNormalizer.Normalize (LoessGCNormalizer.cs:61-82)
            loess_enqueue(ctx, *d.lo, d.count2, d.gc2, d.chrom2, d.alive, ctl, &ctl->do_norm, n, grid_stream);
        } else {
            GcCountView gv{d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, &ctl->do_norm};
            // --- NormalizeByGC (:163-196)
            CG_LAUNCH(ctx, gc_median_request_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->do_norm);
            sel_run_scatter<uint32_t, GcCountView>(ctx, gv, d.sel_gc, n);
            CG_LAUNCH(ctx, gc_median_finish_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->do_norm);
            CG_LAUNCH(ctx, gc_weighted_kernel, GC_BINS, 256, WQ_CAP * 8, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 0, &ctl->do_norm);
            CG_LAUNCH(ctx, gc_weighted_big_kernel, 1, 1024, 0, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 0, &ctl->do_norm, d.wq_key, d.wq_cap);
            CG_LAUNCH(ctx, normalize_apply_bulk_kernel, dim3(grid_k8, 1), K8_THREADS, k8_smem_bytes(), d.count2, d.gc2, d.alive, d.count2,
                      &ctl->n2, 0LL, ctl->med, &ctl->global_median, &ctl->do_norm, 0LL);
        }
    CG_TL(ctx, "normalize 1");
        // --- NormalizeVarianceByGC (:34-97), evaluated only when the metric is on and > 500000 bins
        if (o->want_local_sd && n > 500000) {
            GcCountView gq{d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, &ctl->do_variance};
            CG_LAUNCH(ctx, gc_quartile_request_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->do_variance);
            sel_run_scatter<uint32_t, GcCountView>(ctx, gq, d.sel_gc, n);
            CG_LAUNCH(ctx, gc_quartile_finish_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->do_variance);
            CG_LAUNCH(ctx, gc_weighted_kernel, GC_BINS, 256, WQ_CAP * 8, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 1, &ctl->do_variance);
            CG_LAUNCH(ctx, gc_weighted_big_kernel, 1, 1024, 0, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 1, &ctl->do_variance, d.wq_key, d.wq_cap);
            CG_LAUNCH(ctx, variance_decide_kernel, 1, 1, 0, ctl);
            CG_LAUNCH(ctx, variance_apply_kernel, grid_stream, 256, 0, d.count2, d.gc2, d.alive, ctl);
            // second normalisation when the rescale fired (:516-517)
            if (loess) {
                loess_enqueue(ctx, *d.lo, d.count2, d.gc2, d.chrom2, d.alive, ctl, &ctl->variance_fired, n, grid_stream);
            } else {
                GcCountView gm{d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, &ctl->variance_fired};
                CG_LAUNCH(ctx, gc_median_request_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->variance_fired);
                sel_run_scatter<uint32_t, GcCountView>(ctx, gm, d.sel_gc, n);
                CG_LAUNCH(ctx, gc_median_finish_kernel, 1, 128, 0, d.sel_gc, ctl, &ctl->variance_fired);
                CG_LAUNCH(ctx, gc_weighted_kernel, GC_BINS, 256, WQ_CAP * 8, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 0, &ctl->variance_fired);
                CG_LAUNCH(ctx, gc_weighted_big_kernel, 1, 1024, 0, d.count2, d.gc2, d.chrom2, d.alive, d.is_auto, ctl, 0, &ctl->variance_fired, d.wq_key, d.wq_cap);
                CG_LAUNCH(ctx, normalize_apply_bulk_kernel, dim3(grid_k8, 1), K8_THREADS, k8_smem_bytes(), d.count2, d.gc2, d.alive,
                          d.count2, &ctl->n2, 0LL, ctl->med, &ctl->global_median, &ctl->variance_fired, 0LL);
            }
        }
    }
    CG_TL(ctx, "variance + normalize 2");
    // --- RemoveBinsWithExtremeLocalSD (:308-322) + final compaction
    {
        FinalPred p{d.alive, d.wsd, ctl};
        EmitOut e{d.count2, d.orig2, d.kept, d.count_out};
        compact_run(ctx, p, e, &ctl->n2, n, d.tiles, &ctl->n_out);
    }
    CG_TL(ctx, "final compaction");
    return CG_OK;
}

extern "C" int cg_clean(cg_ctx* ctx, const cg_clean_opts* opts, int64_t n, const uint8_t* chrom,
                        const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
                        const int32_t* start, const int32_t* stop, const float* count, const uint8_t* gc,
                        int64_t* n_out, int32_t* kept_index, float* count_out, double* local_sd,
                        int* gc_norm_skipped) {
    if (!ctx) return CG_ERR_ARG;
    if (n_chrom > 256) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean: more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (see DESIGN.md, Limits)");
    if (!opts || n < 0 || n > 0x7fff0000LL || n_chrom < 0 || !n_out || !local_sd || !gc_norm_skipped)
        return cg_fail(ctx, CG_ERR_ARG, "cg_clean: bad argument");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    *n_out = 0;
    *local_sd = -1.0;
    *gc_norm_skipped = 0;
    if (n == 0) {  // an empty list is "every bin GC-filtered" for the reference (:502-505)
        *gc_norm_skipped = opts->gc_norm ? 1 : 0;
        return CG_OK;
    }
    if (!chrom || !start || !stop || !count || !gc || !kept_index || !count_out || (n_chrom > 0 && !chrom_is_autosome))
        return cg_fail(ctx, CG_ERR_ARG, "cg_clean: null array");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool loess = opts->gc_norm && opts->gc_mode != 0;
    int rc = arena_reserve(ctx, clean_workspace_bytes(n, n_chrom, loess));
    if (rc) return rc;
    CleanDev d;
    rc = clean_alloc(ctx, n, n_chrom, d, loess);
    if (rc) return rc;
    {
        // longest chromosome run (ids are non-decreasing; the device rejects anything else): sizes the local-SD sort
        int64_t prev = 0, longest = 0;
        for (int c = 0; c < n_chrom; c++) {
            int64_t lo = prev, hi = n;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (chrom[mid] <= (uint8_t)c) lo = mid + 1; else hi = mid; }
            longest = std::max(longest, lo - prev);
            prev = lo;
        }
        d.max_chrom_bins = prev == n ? longest : -1;
    }
    cudaStream_t s = ctx->stream;
    if (CgStageSlot* sl = cg_stage_take(ctx, n, chrom, start, stop, count, gc)) {
        // staged by cg_prefetch_bins while the previous call ran: read the columns where they are
        CG_CUDA(ctx, cudaStreamWaitEvent(s, sl->ready, 0));
        d.chrom = sl->chrom; d.gc = sl->gc; d.start = sl->start; d.stop = sl->stop; d.count = sl->count;
    } else {
        CG_CUDA(ctx, cudaMemcpyAsync(d.chrom, chrom, n, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.gc, gc, n, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.start, start, n * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.stop, stop, n * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.count, count, n * 4, cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaMemsetAsync(d.is_auto, 0, 256, s));
    if (n_chrom > 0) CG_CUDA(ctx, cudaMemcpyAsync(d.is_auto, chrom_is_autosome, n_chrom, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d.is_chry, 0, 256, s));
    if (n_chrom > 0 && chrom_is_chrY) CG_CUDA(ctx, cudaMemcpyAsync(d.is_chry, chrom_is_chrY, n_chrom, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    rc = clean_enqueue(ctx, opts, d);
    if (rc) return rc;
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    CleanCtl* h = (CleanCtl*)ctx->pinned;
    CG_CUDA(ctx, cudaMemcpyAsync(h, d.ctl, sizeof(CleanCtl), cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    if (h->unsorted) return cg_fail(ctx, CG_ERR_UNSORTED, "cg_clean: chromosome ids must form non-decreasing runs and GC must be 0..100");
    if (h->need_weighted)
        return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean: a GC bucket with < 100 autosomal bins needs more than 16384 neighbouring values for its weighted quantiles");
    const int64_t m = h->n_out;
    if (m > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(kept_index, d.kept, m * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(count_out, d.count_out, m * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    *n_out = m;
    *local_sd = h->local_sd;
    *gc_norm_skipped = h->gc_skipped;
    return CG_OK;
}

// Stand-alone K8 for the roofline measurement (see canvasgpu.h).
extern "C" int cg_normalize_apply(cg_ctx* ctx, int batch, int64_t n, const float* count, const uint8_t* gc,
                                  const double* median_by_gc, const double* global_median, float* count_out,
                                  int repeats, double* kernel_ms) {
    if (!ctx) return CG_ERR_ARG;
    if (batch <= 0 || n <= 0 || (n & 3) || !count || !gc || !median_by_gc || !global_median || !count_out)
        return cg_fail(ctx, CG_ERR_ARG, "cg_normalize_apply: bad argument (n must be a multiple of 4)");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    const size_t total = (size_t)batch * (size_t)n;
    int rc = arena_reserve(ctx, arena_need(total, 4) * 2 + arena_need(total, 1) + arena_need((size_t)batch * GC_BINS, 8) +
                                    arena_need(batch, 8) + 4096);
    if (rc) return rc;
    float* d_in = arena_take<float>(ctx, total);
    float* d_out = arena_take<float>(ctx, total);
    uint8_t* d_gc = arena_take<uint8_t>(ctx, total);
    double* d_med = arena_take<double>(ctx, (size_t)batch * GC_BINS);
    double* d_gmed = arena_take<double>(ctx, batch);
    if (!d_in || !d_out || !d_gc || !d_med || !d_gmed) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_in, count, total * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_gc, gc, total, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_med, median_by_gc, (size_t)batch * GC_BINS * 8, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_gmed, global_median, (size_t)batch * 8, cudaMemcpyHostToDevice, s));
    if (repeats < 1) repeats = 1;
    const int exact_divide = getenv("CANVAS_K8_EXACT_DIV") ? 1 : 0;  // experiments: FP64 divide on every element
    // one untimed launch, then `repeats` timed ones; the grid fills every SM exactly once (no second wave)
    CG_CUDA(ctx, cudaFuncSetAttribute(normalize_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k8_smem_bytes()));
    const int per_sample_blocks = std::max(1, std::min(div_up(n, K8_TILE), ctx->num_sms * K8_CTAS_PER_SM / batch));
    dim3 grid(per_sample_blocks, batch);
    for (int r = 0; r <= repeats; r++) {
        if (r == 1) CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
        CG_LAUNCH(ctx, normalize_apply_bulk_kernel, grid, K8_THREADS, k8_smem_bytes(), d_in, d_gc, (const uint8_t*)nullptr, d_out,
                  (const int*)nullptr, (long long)n, d_med, d_gmed, (const int*)nullptr, (long long)n, exact_divide);
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    CG_CUDA(ctx, cudaMemcpyAsync(count_out, d_out, total * 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms / repeats;
    if (kernel_ms) *kernel_ms = ms / repeats;
    return CG_OK;
}
