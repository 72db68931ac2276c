// CanvasNormalize on the device: the weighted-average reference of the control samples and the sample / reference
// ratio step with its conversion back to counts.
//
//   cg_normalize_reference   WeightedAverageReferenceGenerator.Run (WeightedAverageReferenceGenerator.cs:33-70) with
//                            BinCounts.OnTargetMedianBinCount (BinCounts.cs:24-62)
//   cg_normalize_ratio       LSNormRatioCalculator.Run (LSNormRatioCalculator.cs:22-48), RawRatioCalculator.Run
//                            (RawRatioCalculator.cs:24-47), CanvasNormalizeUtilities.RatiosToCounts
//                            (CanvasNormalizeUtilities.cs:22-31)
//
// Both are one exact order-statistics step (the medians, select.cuh) followed by one stream over the bins:
// 8 S + 8 B/bin for the reference of S controls, 8 B/bin in and 12 B per kept bin out for the ratio step.
// Arithmetic follows the C# expressions operation by operation (float division, then double factors, float store).
#include <algorithm>
#include <vector>

#include "clean.cuh"
#include "select.cuh"

namespace {

// counts[s * n + bin], segment = control sample; bins off target are not part of any median
struct ControlCountView {
    const double* counts;
    const uint8_t* on_target;  // nullptr: every bin
    long long n;
    int n_samples;
    __device__ long long size() const { return n * n_samples; }
    __device__ bool get(long long i, uint64_t& key, int& a, int& b) const {
        const long long bin = i % n;
        if (on_target && !on_target[bin]) return false;
        a = (int)(i / n);
        b = -1;
        key = f64_key(counts[i]);
        return true;
    }
};

// segment 0 = the sample's counts, segment 1 = the reference's; float order = order of the widened doubles
struct PairCountView {
    const float* sample;
    const float* reference;
    const uint8_t* on_target;
    long long n;
    __device__ long long size() const { return 2 * n; }
    __device__ bool get(long long i, uint32_t& key, int& a, int& b) const {
        const int which = i >= n;
        const long long bin = which ? i - n : i;
        if (on_target && !on_target[bin]) return false;
        a = which;
        b = -1;
        key = f32_key(which ? reference[bin] : sample[bin]);
        return true;
    }
};

__global__ void count_on_target_kernel(const uint8_t* __restrict__ on_target, long long n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += on_target[i] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// SortedList<T>.Median(): ranks of the two middles (the same rank twice for an odd count)
template <typename K>
__global__ void middle_request_kernel(SelState<K> st, const unsigned long long* __restrict__ cnt_ptr, unsigned long long cnt_all) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    const unsigned long long m = cnt_ptr ? *cnt_ptr : cnt_all;
    if (m == 0) { st.nreq[s] = 0; return; }
    st.nreq[s] = 2;
    st.req_k[s * SEL_G + 0] = (m & 1ull) ? m / 2 : m / 2 - 1;
    st.req_k[s * SEL_G + 1] = m / 2;
}

// medians as doubles (mean of the middles in double); an empty list gives 0 (default(T) of the empty SortedList)
__global__ void control_weights_kernel(SelState<uint64_t> st, double* __restrict__ median, double* __restrict__ weight) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double sum = 0.0;
    for (int s = 0; s < st.nseg; s++) {
        double m = 0.0;
        if (st.nreq[s]) {
            const double a = f64_unkey(st.req_key[s * SEL_G + 0]), b = f64_unkey(st.req_key[s * SEL_G + 1]);
            m = st.req_key[s * SEL_G + 0] == st.req_key[s * SEL_G + 1] ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
        }
        median[s] = m;
        const double w = m > 0 ? __ddiv_rn(1.0, m) : 0.0;  // WeightedAverageReferenceGenerator.cs:52
        weight[s] = w;
        sum = __dadd_rn(sum, w);                            // weights.Sum()
    }
    for (int s = 0; s < st.nseg; s++) weight[s] = __ddiv_rn(weight[s], sum);  // :56
}

// weightedBinCount = sum_i weights[i] * counts[i][bin], added in sample order (:70)
__global__ void __launch_bounds__(256) weighted_average_kernel(const double* __restrict__ counts, const double* __restrict__ weight,
                                                               int n_samples, long long n, double* __restrict__ out) {
    extern __shared__ double s_w[];
    for (int t = threadIdx.x; t < n_samples; t += blockDim.x) s_w[t] = weight[t];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < n_samples; s++) acc = __dadd_rn(acc, __dmul_rn(s_w[s], __ldcs(counts + (size_t)s * n + i)));
        __stcs(out + i, acc);
    }
}

struct RatioCtl {
    double sample_median, reference_median, library_size_factor;
    int n, n_kept;
};

__global__ void ratio_factor_kernel(SelState<uint32_t> st, RatioCtl* ctl, int lsnorm, int n) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double med[2] = {0.0, 0.0};
    for (int s = 0; s < 2; s++)
        if (lsnorm && st.nreq[s]) {
            const double a = (double)f32_unkey(st.req_key[s * SEL_G + 0]), b = (double)f32_unkey(st.req_key[s * SEL_G + 1]);
            med[s] = st.req_key[s * SEL_G + 0] == st.req_key[s * SEL_G + 1] ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
        }
    ctl->sample_median = med[0];
    ctl->reference_median = med[1];
    // LSNormRatioCalculator.cs:31
    ctl->library_size_factor = (lsnorm && med[0] > 0 && med[1] > 0) ? __ddiv_rn(med[1], med[0]) : 1.0;
    ctl->n = n;
    ctl->n_kept = 0;
}

// a bin is skipped when its reference count is outside [min_ref, max_ref] (NaN compares false: kept, as in C#)
struct RatioKeep {
    const float* reference;
    double min_ref, max_ref;
    __device__ bool operator()(int i) const {
        const double r = (double)reference[i];
        return !(r < min_ref) && !(r > max_ref);
    }
};

struct RatioEmit {
    const float* sample;
    const float* reference;
    const int32_t* ploidy;  // nullptr: 2
    const RatioCtl* ctl;
    int32_t* kept_index;
    float* ratio;
    float* count;
    __device__ void operator()(int src, int dst) const {
        // `sampleBin.Count / referenceBin.Count` is a float division; the product with the double factor is a double
        const float q = __fdiv_rn(sample[src], reference[src]);
        const float r = (float)__dmul_rn((double)q, ctl->library_size_factor);
        // RatiosToCounts: factor = 40 * ploidy / 2.0; count = (float)(ratio * factor)
        const int p = ploidy ? ploidy[src] : 2;
        const double factor = __ddiv_rn(__dmul_rn(40.0, (double)p), 2.0);
        kept_index[dst] = src;
        ratio[dst] = r;
        count[dst] = (float)__dmul_rn((double)r, factor);
    }
};

// BestLR2: squared log ratios of the median-normalised sample (row n_controls of `counts`) against every control over
// the on-target bins.  Block (chunk, control) adds its chunk in a fixed order; the chunk sums are added in chunk order by
// lr2_finish_kernel: deterministic, but not the reference's single left-to-right sum (agreement better than 1e-9 relative).
constexpr int LR2_CHUNK = 8192;
struct Lr2Partial {
    double sum;
    long long used, ignored;
};

__global__ void __launch_bounds__(256) lr2_partial_kernel(const double* __restrict__ counts, const uint8_t* __restrict__ on_target,
                                                          const double* __restrict__ median, int n_controls, long long n,
                                                          Lr2Partial* __restrict__ part) {
    const int ctrl = blockIdx.y;
    const long long lo = (long long)blockIdx.x * LR2_CHUNK, hi = min(n, lo + LR2_CHUNK);
    const double mt = median[n_controls], mc = median[ctrl];
    const double wt = mt > 0 ? __ddiv_rn(1.0, mt) : 0.0, wc = mc > 0 ? __ddiv_rn(1.0, mc) : 0.0;
    const double* __restrict__ t = counts + (size_t)n_controls * n;
    const double* __restrict__ c = counts + (size_t)ctrl * n;
    // thread k takes every 256th element of the chunk (coalesced rows); its partial and the sum over threads are added in
    // a fixed order
    double sum = 0.0;
    long long used = 0, ign = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        if (on_target && !on_target[i]) continue;
        const double normal = __dmul_rn(c[i], wc);
        if (normal <= 0) { ign++; continue; }  // NaN is not <= 0: it goes on and is dropped by the test below
        const double lr = log(__ddiv_rn(__dmul_rn(t[i], wt), normal));
        const double sq = __dmul_rn(lr, lr);
        if (isinf(sq) || isnan(sq)) { ign++; continue; }
        sum = __dadd_rn(sum, sq);
        used++;
    }
    __shared__ double s_sum[256];
    __shared__ long long s_used[256], s_ign[256];
    s_sum[threadIdx.x] = sum; s_used[threadIdx.x] = used; s_ign[threadIdx.x] = ign;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0;
        long long tu = 0, ti = 0;
        for (int k = 0; k < 256; k++) { ts = __dadd_rn(ts, s_sum[k]); tu += s_used[k]; ti += s_ign[k]; }
        part[(size_t)ctrl * gridDim.x + blockIdx.x] = Lr2Partial{ts, tu, ti};
    }
}

// mean squared log ratio per control and the first strict minimum (BestLR2ReferenceGenerator.cs:62-78); -1 when no
// control beats +infinity
__global__ void lr2_finish_kernel(const Lr2Partial* __restrict__ part, int n_controls, int n_chunks, double* __restrict__ mean,
                                  long long* __restrict__ ignored, int* __restrict__ best) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int b = -1;
    double mn = __longlong_as_double(0x7ff0000000000000ll);
    for (int c = 0; c < n_controls; c++) {
        double s = 0.0;
        long long u = 0, ig = 0;
        for (int k = 0; k < n_chunks; k++) {
            const Lr2Partial p = part[(size_t)c * n_chunks + k];
            s = __dadd_rn(s, p.sum); u += p.used; ig += p.ignored;
        }
        const double m = u > 0 ? __ddiv_rn(s, (double)u) : s;
        mean[c] = m;
        ignored[c] = ig;
        if (m < mn) { mn = m; b = c; }
    }
    *best = b;
}

// ---------------------------------------------------------------------------------------------
// PCA reference (PCAReferenceGenerator.cs:37-78).  Dot products are chunk-wise device sums added in chunk order
// (deterministic; the reference adds left to right: agreement ~1e-13 relative, far inside the 1e-5 the path allows).
// ---------------------------------------------------------------------------------------------
constexpr int DOT_CHUNK = 8192;

// part[pair * n_chunks + chunk] = sum over the chunk of a[i] * b[i]; pair p reads rows ia[p], ib[p] of `rows` (row -1 =
// the centred sample: (double)max(1, sample) - (double)mu)
struct DotPairs {
    int ia[64], ib[64];
};

__device__ inline double pca_centred(const float* sample, const float* mu, long long i) {
    return __dsub_rn((double)fmaxf(1.0f, sample[i]), (double)mu[i]);
}

__global__ void __launch_bounds__(256) dot_partial_kernel(const double* __restrict__ rows, long long n, DotPairs pairs,
                                                          const float* __restrict__ sample, const float* __restrict__ mu,
                                                          double* __restrict__ part) {
    const int pr = blockIdx.y;
    const int ia = pairs.ia[pr], ib = pairs.ib[pr];
    const long long lo = (long long)blockIdx.x * DOT_CHUNK, hi = min(n, lo + DOT_CHUNK);
    double sum = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {  // coalesced rows; fixed order of additions
        const double x = ia < 0 ? pca_centred(sample, mu, i) : rows[(size_t)ia * n + i];
        const double y = ib < 0 ? pca_centred(sample, mu, i) : rows[(size_t)ib * n + i];
        sum = __dadd_rn(sum, __dmul_rn(x, y));
    }
    __shared__ double s_sum[256];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 256; k++) t = __dadd_rn(t, s_sum[k]);
        part[(size_t)pr * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void dot_finish_kernel(const double* __restrict__ part, int n_pairs, int n_chunks, double* __restrict__ out) {
    const int pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= n_pairs) return;
    double t = 0.0;
    for (int k = 0; k < n_chunks; k++) t = __dadd_rn(t, part[(size_t)pr * n_chunks + k]);
    out[pr] = t;
}

// NormalizeBy2Norm (Utilities.cs:650-667): v / sqrt(sum of squares), unchanged when the norm is zero
__global__ void __launch_bounds__(256) axes_normalize_kernel(double* __restrict__ rows, long long n, int n_axes, const double* __restrict__ sumsq) {
    const int k = blockIdx.y;
    const double size = sqrt(sumsq[k]);
    if (size == 0) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        rows[(size_t)k * n + i] = __ddiv_rn(rows[(size_t)k * n + i], size);
}

// AreOrthogonal (:685-692) over all pairs; gram[] holds the pair dot products after the K axis projections
__global__ void pca_check_kernel(const double* __restrict__ gram, int n_pairs, int* __restrict__ bad) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int b = 0;
    for (int p = 0; p < n_pairs; p++)
        if (fabs(gram[p]) > 1e-4) b = 1;
    *bad = b;
}

// reference vector, the temporary reference file's F2 round trip, the raw ratio and its keep flag (:49-63,
// RawRatioCalculator.cs:36-45)
__global__ void __launch_bounds__(256) pca_reference_kernel(const double* __restrict__ rows, long long n, int n_axes,
                                                            const double* __restrict__ size, const float* __restrict__ sample,
                                                            const float* __restrict__ mu, double min_ref, double max_ref,
                                                            double* __restrict__ ref_d, float* __restrict__ ratio,
                                                            uint8_t* __restrict__ kept) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double proj = __dmul_rn(size[0], rows[i]);
        for (int k = 1; k < n_axes; k++) proj = __dadd_rn(proj, __dmul_rn(size[k], rows[(size_t)k * n + i]));
        const double r = fmax(1.0, __dadd_rn((double)mu[i], proj));
        ref_d[i] = r;
        const float rf = (float)dotnet_f2_roundtrip((float)r);  // written with {0:F2}, read back with float.Parse
        const double rd = (double)rf;
        kept[i] = !(rd < min_ref) && !(rd > max_ref);
        ratio[i] = __fdiv_rn(sample[i], rf);
    }
}

struct RatioListView {  // the kept (and on-target) ratios: one list
    const float* ratio;
    const uint8_t* kept;
    const uint8_t* on_target;
    long long n;
    __device__ long long size() const { return n; }
    __device__ bool get(long long i, uint32_t& key, int& a, int& b) const {
        if (!kept[i] || (on_target && !on_target[i])) return false;
        a = 0;
        b = -1;
        key = f32_key(ratio[i]);
        return true;
    }
};

__global__ void count_kept_kernel(const uint8_t* __restrict__ kept, const uint8_t* __restrict__ on_target, long long n,
                                  unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c += kept[i] && (!on_target || on_target[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void __launch_bounds__(256) pca_scale_kernel(const double* __restrict__ ref_d, long long n, SelState<uint32_t> st,
                                                        double* __restrict__ median_out, float* __restrict__ out) {
    double med = 0.0;  // Median of an empty list
    if (st.nreq[0]) {
        const double a = (double)f32_unkey(st.req_key[0]), b = (double)f32_unkey(st.req_key[1]);
        med = st.req_key[0] == st.req_key[1] ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *median_out = med;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)__dmul_rn(ref_d[i], med);
}

void reset_call(cg_ctx* ctx) {
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
}

}  // namespace

extern "C" int cg_normalize_reference(cg_ctx* ctx, int n_samples, int64_t n, const double* counts, const uint8_t* on_target,
                                      double* median, double* weight, double* reference) {
    if (!ctx) return CG_ERR_ARG;
    if (n_samples < 1 || n < 0 || !median || !weight || (n > 0 && (!counts || !reference)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_normalize_reference: bad argument");
    if (n_samples > 256)  // the selection pass keeps one prefix row per sample in shared memory (48 KB without opt-in)
        return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_reference: more than 256 control samples");
    if (n > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_reference: too many bins");
    reset_call(ctx);
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t total = (size_t)n_samples * (size_t)n;
    int rc = arena_reserve(ctx, arena_need(total, 8) + arena_need(n, 8) + arena_need(n, 1) + arena_need(n_samples, 8) * 2 +
                                    sel_state_bytes<uint64_t>(n_samples) + (1u << 16));
    if (rc) return rc;
    double* d_counts = arena_take<double>(ctx, std::max<size_t>(total, 1));
    double* d_ref = arena_take<double>(ctx, std::max<int64_t>(n, 1));
    uint8_t* d_on = on_target ? arena_take<uint8_t>(ctx, std::max<int64_t>(n, 1)) : nullptr;
    double* d_med = arena_take<double>(ctx, n_samples);
    double* d_w = arena_take<double>(ctx, n_samples);
    unsigned long long* d_cnt = arena_take<unsigned long long>(ctx, 1);
    SelState<uint64_t> st;
    if (!d_counts || !d_ref || (on_target && !d_on) || !d_med || !d_w || !d_cnt || !sel_state_alloc<uint64_t>(ctx, n_samples, st))
        return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    if (total) CG_CUDA(ctx, cudaMemcpyAsync(d_counts, counts, total * 8, cudaMemcpyHostToDevice, s));
    if (on_target && n) CG_CUDA(ctx, cudaMemcpyAsync(d_on, on_target, (size_t)n, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_CUDA(ctx, cudaMemsetAsync(st.hist, 0, (size_t)n_samples * SEL_G * SEL_BINS * sizeof(unsigned), s));
    CG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 8, s));
    if (on_target && n) CG_LAUNCH(ctx, count_on_target_kernel, std::min<long long>(ctx->num_sms * 4, div_up(n, 256)), 256, 0, d_on, (long long)n, d_cnt);
    CG_LAUNCH(ctx, middle_request_kernel<uint64_t>, div_up(n_samples, 128), 128, 0, st, on_target ? d_cnt : nullptr, (unsigned long long)n);
    ControlCountView v{d_counts, d_on, (long long)n, n_samples};
    sel_run_scatter<uint64_t, ControlCountView>(ctx, v, st, (long long)total);
    CG_LAUNCH(ctx, control_weights_kernel, 1, 32, 0, st, d_med, d_w);
    if (n) CG_LAUNCH(ctx, weighted_average_kernel, std::min<long long>(ctx->num_sms * 8, div_up(n, 256)), 256, (size_t)n_samples * 8, d_counts, d_w, n_samples, (long long)n, d_ref);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    CG_CUDA(ctx, cudaMemcpyAsync(median, d_med, (size_t)n_samples * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(weight, d_w, (size_t)n_samples * 8, cudaMemcpyDeviceToHost, s));
    if (n) CG_CUDA(ctx, cudaMemcpyAsync(reference, d_ref, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}

extern "C" int cg_normalize_ratio(cg_ctx* ctx, int64_t n, const float* sample, const float* reference, const uint8_t* on_target,
                                  int mode, double min_ref, double max_ref, const int32_t* ploidy, int64_t* n_out,
                                  int32_t* kept_index, float* ratio, float* count, double* library_size_factor) {
    if (!ctx) return CG_ERR_ARG;
    if (n < 0 || !n_out || (mode != 0 && mode != 1) || (n > 0 && (!sample || !reference || !kept_index || !ratio || !count)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_normalize_ratio: bad argument");
    if (n > 0x3fff0000LL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_ratio: too many bins");
    reset_call(ctx);
    *n_out = 0;
    const int lsnorm = mode == 1;
    if (lsnorm) { min_ref = 1.0; max_ref = __builtin_inf(); }  // LSNormRatioCalculator.cs:44 drops reference counts below 1 only
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ntiles = std::max(1, div_up((int)n, CMP_TILE));
    int rc = arena_reserve(ctx, arena_need(n, 4) * 6 + arena_need(n, 1) + arena_need(ntiles, 4) + sel_state_bytes<uint32_t>(2) + (1u << 16));
    if (rc) return rc;
    const size_t n1 = (size_t)std::max<int64_t>(n, 1);
    float* d_s = arena_take<float>(ctx, n1);
    float* d_r = arena_take<float>(ctx, n1);
    int32_t* d_p = ploidy ? arena_take<int32_t>(ctx, n1) : nullptr;
    uint8_t* d_on = on_target ? arena_take<uint8_t>(ctx, n1) : nullptr;
    // kept index, ratio and count of the kept bins lie next to each other: one download
    int32_t* d_idx = arena_take<int32_t>(ctx, n1);
    float* d_ratio = arena_take<float>(ctx, n1);
    float* d_count = arena_take<float>(ctx, n1);
    int* d_tiles = arena_take<int>(ctx, ntiles);
    RatioCtl* d_ctl = arena_take<RatioCtl>(ctx, 1);
    unsigned long long* d_cnt = arena_take<unsigned long long>(ctx, 1);
    SelState<uint32_t> st;
    if (!d_s || !d_r || (ploidy && !d_p) || (on_target && !d_on) || !d_idx || !d_ratio || !d_count || !d_tiles || !d_ctl || !d_cnt ||
        !sel_state_alloc<uint32_t>(ctx, 2, st))
        return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    if (n) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_s, sample, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d_r, reference, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        if (ploidy) CG_CUDA(ctx, cudaMemcpyAsync(d_p, ploidy, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        if (on_target) CG_CUDA(ctx, cudaMemcpyAsync(d_on, on_target, (size_t)n, cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_CUDA(ctx, cudaMemsetAsync(st.hist, 0, (size_t)2 * SEL_G * SEL_BINS * sizeof(unsigned), s));
    CG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 8, s));
    if (lsnorm) {
        if (on_target && n) CG_LAUNCH(ctx, count_on_target_kernel, std::min<long long>(ctx->num_sms * 4, div_up(n, 256)), 256, 0, d_on, (long long)n, d_cnt);
        CG_LAUNCH(ctx, middle_request_kernel<uint32_t>, 1, 128, 0, st, on_target ? d_cnt : nullptr, (unsigned long long)n);
        PairCountView v{d_s, d_r, d_on, (long long)n};
        sel_run_scatter<uint32_t, PairCountView>(ctx, v, st, 2 * (long long)n);
    }
    CG_LAUNCH(ctx, ratio_factor_kernel, 1, 32, 0, st, d_ctl, lsnorm, (int)n);
    RatioKeep keep{d_r, min_ref, max_ref};
    RatioEmit emit{d_s, d_r, d_p, d_ctl, d_idx, d_ratio, d_count};
    compact_run(ctx, keep, emit, &d_ctl->n, (int)n, d_tiles, &d_ctl->n_kept);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    RatioCtl h;
    CG_CUDA(ctx, cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CHECK_LAUNCHES(ctx);
    if (h.n_kept > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(kept_index, d_idx, (size_t)h.n_kept * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(ratio, d_ratio, (size_t)h.n_kept * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(count, d_count, (size_t)h.n_kept * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    CG_CUDA(ctx, cudaGetLastError());
    *n_out = h.n_kept;
    if (library_size_factor) *library_size_factor = h.library_size_factor;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}

extern "C" int cg_normalize_best_lr2(cg_ctx* ctx, int n_controls, int64_t n, const double* sample, const double* controls,
                                     const uint8_t* on_target, int* best_index, double* mean_sq_log_ratio, int64_t* ignored) {
    if (!ctx) return CG_ERR_ARG;
    if (n_controls < 1 || n < 0 || !best_index || !mean_sq_log_ratio || !ignored || (n > 0 && (!sample || !controls)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_normalize_best_lr2: bad argument");
    if (n_controls > 255) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_best_lr2: more than 255 control samples");
    if (n > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_best_lr2: too many bins");
    reset_call(ctx);
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rows = n_controls + 1;  // the sample is the last row
    const size_t total = (size_t)rows * (size_t)n;
    const int n_chunks = std::max<int>(1, (int)div_up((long long)n, (long long)LR2_CHUNK));
    int rc = arena_reserve(ctx, arena_need(total, 8) + arena_need(n, 1) + arena_need(rows, 8) * 3 +
                                    arena_need((size_t)n_controls * n_chunks, sizeof(Lr2Partial)) + sel_state_bytes<uint64_t>(rows) + (1u << 16));
    if (rc) return rc;
    double* d_counts = arena_take<double>(ctx, std::max<size_t>(total, 1));
    uint8_t* d_on = on_target ? arena_take<uint8_t>(ctx, std::max<int64_t>(n, 1)) : nullptr;
    double* d_med = arena_take<double>(ctx, rows);
    double* d_w = arena_take<double>(ctx, rows);
    double* d_mean = arena_take<double>(ctx, rows);
    long long* d_ign = arena_take<long long>(ctx, rows);
    int* d_best = arena_take<int>(ctx, 1);
    unsigned long long* d_cnt = arena_take<unsigned long long>(ctx, 1);
    Lr2Partial* d_part = arena_take<Lr2Partial>(ctx, (size_t)n_controls * n_chunks);
    SelState<uint64_t> st;
    if (!d_counts || (on_target && !d_on) || !d_med || !d_w || !d_mean || !d_ign || !d_best || !d_cnt || !d_part ||
        !sel_state_alloc<uint64_t>(ctx, rows, st))
        return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    if (n) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_counts, controls, (size_t)n_controls * n * 8, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d_counts + (size_t)n_controls * n, sample, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        if (on_target) CG_CUDA(ctx, cudaMemcpyAsync(d_on, on_target, (size_t)n, cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_CUDA(ctx, cudaMemsetAsync(st.hist, 0, (size_t)rows * SEL_G * SEL_BINS * sizeof(unsigned), s));
    CG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 8, s));
    if (on_target && n) CG_LAUNCH(ctx, count_on_target_kernel, std::min<long long>(ctx->num_sms * 4, div_up(n, 256)), 256, 0, d_on, (long long)n, d_cnt);
    CG_LAUNCH(ctx, middle_request_kernel<uint64_t>, div_up(rows, 128), 128, 0, st, on_target ? d_cnt : nullptr, (unsigned long long)n);
    ControlCountView v{d_counts, d_on, (long long)n, rows};
    sel_run_scatter<uint64_t, ControlCountView>(ctx, v, st, (long long)total);
    CG_LAUNCH(ctx, control_weights_kernel, 1, 32, 0, st, d_med, d_w);  // medians; the weights are formed per pair below
    CG_LAUNCH(ctx, lr2_partial_kernel, dim3(n_chunks, n_controls), 256, 0, d_counts, d_on, d_med, n_controls, (long long)n, d_part);
    CG_LAUNCH(ctx, lr2_finish_kernel, 1, 32, 0, d_part, n_controls, n_chunks, d_mean, d_ign, d_best);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    int h_best = -1;
    std::vector<long long> h_ign(n_controls);
    CG_CUDA(ctx, cudaMemcpyAsync(mean_sq_log_ratio, d_mean, (size_t)n_controls * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(h_ign.data(), d_ign, (size_t)n_controls * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(&h_best, d_best, 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    for (int c = 0; c < n_controls; c++) ignored[c] = h_ign[c];
    *best_index = h_best;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}

extern "C" int cg_normalize_pca_reference(cg_ctx* ctx, int64_t n, int n_axes, const float* sample, const float* mu, const double* axes,
                                          const uint8_t* on_target, double min_ref, double max_ref, float* reference,
                                          double* median_ratio) {
    if (!ctx) return CG_ERR_ARG;
    if (n < 0 || n_axes < 1 || !median_ratio || (n > 0 && (!sample || !mu || !axes || !reference)))
        return cg_fail(ctx, CG_ERR_ARG, n_axes < 1 ? "No axes to project onto." : "cg_normalize_pca_reference: bad argument");
    if (n_axes > 10) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_pca_reference: more than 10 axes");
    if (n > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_normalize_pca_reference: too many bins");
    reset_call(ctx);
    *median_ratio = 0.0;
    if (n == 0) return CG_OK;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int K = n_axes, n_gram = K * (K - 1) / 2;
    const int n_chunks = (int)div_up((long long)n, (long long)DOT_CHUNK);
    const int max_pairs = std::max(K, std::max(n_gram, 1));
    int rc = arena_reserve(ctx, arena_need((size_t)K * n, 8) + arena_need(n, 8) + arena_need(n, 4) * 4 + arena_need(n, 1) * 2 +
                                    arena_need((size_t)max_pairs * n_chunks, 8) + arena_need(64, 8) * 3 + sel_state_bytes<uint32_t>(1) + (1u << 16));
    if (rc) return rc;
    double* d_rows = arena_take<double>(ctx, (size_t)K * n);
    double* d_ref = arena_take<double>(ctx, n);
    float* d_sample = arena_take<float>(ctx, n);
    float* d_mu = arena_take<float>(ctx, n);
    float* d_ratio = arena_take<float>(ctx, n);
    float* d_out = arena_take<float>(ctx, n);
    uint8_t* d_kept = arena_take<uint8_t>(ctx, n);
    uint8_t* d_on = on_target ? arena_take<uint8_t>(ctx, n) : nullptr;
    double* d_part = arena_take<double>(ctx, (size_t)max_pairs * n_chunks);
    double* d_sumsq = arena_take<double>(ctx, 64);
    double* d_gram = arena_take<double>(ctx, 64);
    double* d_size = arena_take<double>(ctx, 64);
    int* d_bad = arena_take<int>(ctx, 1);
    double* d_med = arena_take<double>(ctx, 1);
    unsigned long long* d_cnt = arena_take<unsigned long long>(ctx, 1);
    SelState<uint32_t> st;
    if (!d_rows || !d_ref || !d_sample || !d_mu || !d_ratio || !d_out || !d_kept || (on_target && !d_on) || !d_part || !d_sumsq || !d_gram ||
        !d_size || !d_bad || !d_med || !d_cnt || !sel_state_alloc<uint32_t>(ctx, 1, st))
        return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_rows, axes, (size_t)K * n * 8, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_sample, sample, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_mu, mu, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (on_target) CG_CUDA(ctx, cudaMemcpyAsync(d_on, on_target, (size_t)n, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_CUDA(ctx, cudaMemsetAsync(st.hist, 0, (size_t)SEL_G * SEL_BINS * sizeof(unsigned), s));
    CG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 8, s));
    const int grid_stream = (int)std::min<long long>(ctx->num_sms * 8, div_up((long long)n, 256LL));
    DotPairs pairs;
    // norms of the axes as read from the model file, then the unit axes
    for (int k = 0; k < K; k++) { pairs.ia[k] = k; pairs.ib[k] = k; }
    CG_LAUNCH(ctx, dot_partial_kernel, dim3(n_chunks, K), 256, 0, d_rows, (long long)n, pairs, d_sample, d_mu, d_part);
    CG_LAUNCH(ctx, dot_finish_kernel, 1, 64, 0, d_part, K, n_chunks, d_sumsq);
    CG_LAUNCH(ctx, axes_normalize_kernel, dim3(grid_stream, K), 256, 0, d_rows, (long long)n, K, d_sumsq);
    // pairwise orthogonality of the unit axes
    if (n_gram > 0) {
        int q = 0;
        for (int i = 0; i < K; i++)
            for (int j = i + 1; j < K; j++) { pairs.ia[q] = i; pairs.ib[q] = j; q++; }
        CG_LAUNCH(ctx, dot_partial_kernel, dim3(n_chunks, n_gram), 256, 0, d_rows, (long long)n, pairs, d_sample, d_mu, d_part);
        CG_LAUNCH(ctx, dot_finish_kernel, 1, 64, 0, d_part, n_gram, n_chunks, d_gram);
    }
    CG_LAUNCH(ctx, pca_check_kernel, 1, 32, 0, d_gram, n_gram, d_bad);
    // sizes of the projections of the centred sample
    for (int k = 0; k < K; k++) { pairs.ia[k] = -1; pairs.ib[k] = k; }
    CG_LAUNCH(ctx, dot_partial_kernel, dim3(n_chunks, K), 256, 0, d_rows, (long long)n, pairs, d_sample, d_mu, d_part);
    CG_LAUNCH(ctx, dot_finish_kernel, 1, 64, 0, d_part, K, n_chunks, d_size);
    CG_LAUNCH(ctx, pca_reference_kernel, grid_stream, 256, 0, d_rows, (long long)n, K, d_size, d_sample, d_mu, min_ref, max_ref, d_ref,
              d_ratio, d_kept);
    // median of the kept (on-target) ratios, then the scaled reference
    CG_LAUNCH(ctx, count_kept_kernel, grid_stream, 256, 0, d_kept, d_on, (long long)n, d_cnt);
    CG_LAUNCH(ctx, middle_request_kernel<uint32_t>, 1, 32, 0, st, d_cnt, 0ull);
    RatioListView v{d_ratio, d_kept, d_on, (long long)n};
    sel_run_scatter<uint32_t, RatioListView>(ctx, v, st, (long long)n);
    CG_LAUNCH(ctx, pca_scale_kernel, grid_stream, 256, 0, d_ref, (long long)n, st, d_med, d_out);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    int h_bad = 0;
    CG_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(median_ratio, d_med, 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(reference, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    if (h_bad) return cg_fail(ctx, CG_ERR_ARG, "Axes are not orthogonal to each other.");
    return CG_OK;
}
