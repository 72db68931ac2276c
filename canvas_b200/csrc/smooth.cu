// cg_smooth: CanvasSmooth's repeated median filter on the device (reference CanvasSmooth/CanvasSmooth.cs:44-77,
// Utilities.MedianFilter CanvasCommon/Utilities.cs:767-791).
//
// The reference streams every chromosome through a sorted window: output i is the median of x[max(0, i - h) ..
// min(n - 1, i + h)] — the window only grows at the start and only shrinks at the end — and a chromosome shorter than
// 2h + 1 bins yields FEWER outputs than inputs (n - h medians of the growing window, then n - h - 1 of the shrinking
// one; none at all when n <= h), which truncates its bin list (Enumerable.Zip, CanvasSmooth.cs:61).  The passes
// h = 1 .. maxHalfWindowSize depend on each other, each pass is an independent stencil: one thread per output sorts its
// <= 2h + 1 keys (.NET float order, NaN first) and takes SortedList<float>.Median() — the mean of the two middle values
// in float for an even count.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int SMOOTH_MAX_HALF = 32;  // window of at most 65 values per thread

struct SmoothChrom {
    long long in_off, out_off;
    int n_in, n_out;
};

__global__ void __launch_bounds__(128) median_filter_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            const SmoothChrom* __restrict__ chroms, int h) {
    const SmoothChrom c = chroms[blockIdx.y];
    const int n = c.n_in;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n_out; i += gridDim.x * blockDim.x) {
        int lo, hi;
        if (n >= 2 * h + 1) { lo = max(0, i - h); hi = min(n - 1, i + h); }
        else if (i <= n - h - 1) { lo = 0; hi = h + i; }       // the window is still growing (nothing has left it yet)
        else { lo = i - (n - h - 1); hi = n - 1; }               // input exhausted: the oldest values leave one by one
        unsigned k[2 * SMOOTH_MAX_HALF + 1];
        const int m = hi - lo + 1;
        const float* x = in + c.in_off + lo;
        for (int t = 0; t < m; t++) {  // insertion sort on the order-preserving keys
            const unsigned key = f32_key(x[t]);
            int q = t;
            while (q > 0 && k[q - 1] > key) { k[q] = k[q - 1]; q--; }
            k[q] = key;
        }
        const float a = f32_unkey(k[(m - 1) / 2]);
        out[c.out_off + i] = (m & 1) ? a : __fdiv_rn(__fadd_rn(a, f32_unkey(k[m / 2])), 2.0f);
    }
}

int out_len(int n, int h) {
    if (n >= 2 * h + 1) return n;
    return std::max(0, 2 * n - 2 * h - 1);
}

}  // namespace

extern "C" int cg_smooth(cg_ctx* ctx, int max_half_window, int n_chrom, const int64_t* chrom_off, const float* count,
                         int64_t* n_out, float* count_out) {
    if (!ctx) return CG_ERR_ARG;
    if (max_half_window < 0 || n_chrom < 0 || !chrom_off || !n_out) return cg_fail(ctx, CG_ERR_ARG, "cg_smooth: bad argument");
    if (max_half_window > SMOOTH_MAX_HALF) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_smooth: half windows up to 32 are supported");
    if (n_chrom > 65535) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_smooth: too many chromosomes");
    for (int c = 0; c < n_chrom; c++)
        if (chrom_off[c + 1] < chrom_off[c] || chrom_off[c + 1] - chrom_off[c] > 0x7fff0000LL)
            return cg_fail(ctx, CG_ERR_ARG, "cg_smooth: bad chromosome offsets");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    const long long N = n_chrom > 0 ? chrom_off[n_chrom] - chrom_off[0] : 0;
    for (int c = 0; c < n_chrom; c++) n_out[c] = chrom_off[c + 1] - chrom_off[c];
    if (N == 0) return CG_OK;
    if (!count || !count_out) return cg_fail(ctx, CG_ERR_ARG, "cg_smooth: null array");
    const long long base = chrom_off[0];
    if (max_half_window == 0) {  // the loop of RepeatedMedianFilter does not run: counts pass through
        std::copy(count + base, count + base + N, count_out + base);
        return CG_OK;
    }
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int H = max_half_window;
    int rc = arena_reserve(ctx, arena_need(N, 4) * 2 + arena_need((size_t)n_chrom * H, sizeof(SmoothChrom)) + (1u << 16));
    if (rc) return rc;
    float* buf[2] = {arena_take<float>(ctx, N), arena_take<float>(ctx, N)};
    SmoothChrom* d_tab = arena_take<SmoothChrom>(ctx, (size_t)n_chrom * H);
    if (!buf[0] || !buf[1] || !d_tab) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    // per pass: every chromosome keeps its slot (offset) and only its length shrinks
    std::vector<SmoothChrom> tab((size_t)n_chrom * H);
    std::vector<int> len(n_chrom);
    int max_out = 0;
    for (int c = 0; c < n_chrom; c++) len[c] = (int)(chrom_off[c + 1] - chrom_off[c]);
    for (int h = 1; h <= H; h++)
        for (int c = 0; c < n_chrom; c++) {
            SmoothChrom& t = tab[(size_t)(h - 1) * n_chrom + c];
            t.in_off = t.out_off = chrom_off[c] - base;
            t.n_in = len[c];
            t.n_out = out_len(len[c], h);
            len[c] = t.n_out;
            max_out = std::max(max_out, t.n_out);
        }
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(buf[0], count + base, (size_t)N * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(SmoothChrom), cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    const int gx = std::max(1, std::min(div_up(std::max(max_out, 1), 128), std::max(1, ctx->num_sms * 16 / std::max(1, n_chrom))));
    for (int h = 1; h <= H; h++)
        CG_LAUNCH(ctx, median_filter_kernel, dim3(gx, n_chrom), 128, 0, buf[(h - 1) & 1], buf[h & 1], d_tab + (size_t)(h - 1) * n_chrom, h);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    for (int c = 0; c < n_chrom; c++) {
        n_out[c] = len[c];
        if (len[c] > 0)
            CG_CUDA(ctx, cudaMemcpyAsync(count_out + chrom_off[c], buf[H & 1] + (chrom_off[c] - base), (size_t)len[c] * 4, cudaMemcpyDeviceToHost, s));
    }
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}
