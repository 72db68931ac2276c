// cg_merge_common_bins: the pedigree step between CanvasClean and CanvasPartition — keep only the bins that
// survived CanvasClean in EVERY sample (reference Utilities.MergeMultiSampleCleanedBedFile, CanvasCommon/Utilities.cs:
// 834-920, written back per sample by CanvasRunner.NormalizeCanvasClean, Canvas/CanvasRunner.cs:883-903).
//
// The reference keys dictionaries by (chromosome, start): a bin is kept when all samples list that key ("if outlier
// is removed in one sample, remove it in all samples", :901-903); kept bins come out in the first sample's order,
// each with its own count per sample and the stop of the LAST sample that listed it (:885).  Here every sample is a
// sorted key column (chromosome id << 32 | start, strictly increasing — .cleaned files are written in genome order),
// the bins of sample 0 are looked up in the other samples by binary search over the L2-resident key columns, and
// one order-preserving stream compaction gathers the survivors.
#include <algorithm>

#include "clean.cuh"

namespace {

constexpr int MERGE_MAX_SAMPLES = 8;

struct MergeCols {
    const unsigned long long* key[MERGE_MAX_SAMPLES];
    const int32_t* stop[MERGE_MAX_SAMPLES];
    const float* count[MERGE_MAX_SAMPLES];
    int n[MERGE_MAX_SAMPLES];
    int S;
};

struct MergeCtl {
    int n0;         // bins of sample 0
    int n_out;      // common bins
    int unsorted;   // a key column is not strictly increasing
    int bad_start;  // "Start must be non-negative" (:905-908)
    int bad_stop;   // "Start must be less than Stop" (:909-912)
};

__global__ void merge_key_kernel(const uint8_t* __restrict__ chrom, const int32_t* __restrict__ start, int n,
                                 unsigned long long* __restrict__ key, MergeCtl* ctl) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // order by chromosome id, then by start as a signed number (negative starts sort first and are rejected later)
        const unsigned long long k = ((unsigned long long)chrom[i] << 32) | (unsigned long long)i32_key(start[i]);
        key[i] = k;
        if (i > 0) {
            const unsigned long long p = ((unsigned long long)chrom[i - 1] << 32) | (unsigned long long)i32_key(start[i - 1]);
            if (p >= k) ctl->unsorted = 1;
        }
    }
}

// match[s][i] = index in sample s of the key of bin i of sample 0, or -1
__global__ void merge_match_kernel(MergeCols c, int* __restrict__ match) {
    const int n0 = c.n[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) {
        const unsigned long long k = c.key[0][i];
        for (int s = 1; s < c.S; s++) {
            const unsigned long long* ks = c.key[s];
            int lo = 0, hi = c.n[s];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ks[mid] < k) lo = mid + 1; else hi = mid;
            }
            match[(size_t)s * n0 + i] = (lo < c.n[s] && ks[lo] == k) ? lo : -1;
        }
    }
}

struct CommonPred {
    const int* match;
    int n0, S;
    __device__ bool operator()(int i) const {
        for (int s = 1; s < S; s++)
            if (match[(size_t)s * n0 + i] < 0) return false;
        return true;
    }
};

struct CommonEmit {
    MergeCols c;
    const int* match;
    const int32_t* start0;
    int32_t* kept;
    int32_t* stop_out;
    float* count_out;  // [S][n0]
    MergeCtl* ctl;
    __device__ void operator()(int src, int dst) const {
        const int n0 = c.n[0];
        kept[dst] = src;
        count_out[dst] = c.count[0][src];
        int stop = c.stop[0][src];
        for (int s = 1; s < c.S; s++) {
            const int j = match[(size_t)s * n0 + src];
            count_out[(size_t)s * n0 + dst] = c.count[s][j];
            stop = c.stop[s][j];  // the last sample's stop wins (stop[chr][pos] is overwritten per file)
        }
        stop_out[dst] = stop;
        const int st = start0[src];
        if (st < 0) ctl->bad_start = 1;
        else if (st >= stop) ctl->bad_stop = 1;
    }
};

}  // namespace

extern "C" int cg_merge_common_bins(cg_ctx* ctx, int n_samples, const int64_t* n, const uint8_t* const* chrom,
                                    const int32_t* const* start, const int32_t* const* stop, const float* const* count,
                                    int64_t* n_out, int32_t* kept_index, int32_t* stop_out, float* count_out) {
    if (!ctx) return CG_ERR_ARG;
    if (n_samples < 1 || !n || !chrom || !start || !stop || !count || !n_out)
        return cg_fail(ctx, CG_ERR_ARG, "cg_merge_common_bins: bad argument");
    if (n_samples > MERGE_MAX_SAMPLES) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_merge_common_bins: at most 8 samples");
    const int S = n_samples;
    size_t total = 0;
    for (int s = 0; s < S; s++) {
        if (n[s] < 0 || n[s] > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_common_bins: bad sample length");
        if (n[s] > 0 && (!chrom[s] || !start[s] || !stop[s] || !count[s])) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_common_bins: null column");
        total += (size_t)n[s];
    }
    *n_out = 0;
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    const int n0 = (int)n[0];
    if (n0 == 0) return CG_OK;
    if (!kept_index || !stop_out || !count_out) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_common_bins: null output");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t need = arena_need(1, sizeof(MergeCtl)) + arena_need(n0 / CMP_TILE + 2, 4) + arena_need((size_t)S * n0, 4) * 2 +
                  arena_need(n0, 4) * 2 + (1u << 16);
    for (int s = 0; s < S; s++) need += arena_need(n[s], 1) + arena_need(n[s], 4) * 3 + arena_need(n[s], 8);
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    MergeCtl* ctl = arena_take<MergeCtl>(ctx, 1);
    int* tiles = arena_take<int>(ctx, n0 / CMP_TILE + 2);
    int* match = arena_take<int>(ctx, (size_t)S * n0);
    float* d_count_out = arena_take<float>(ctx, (size_t)S * n0);
    int32_t* d_kept = arena_take<int32_t>(ctx, n0);
    int32_t* d_stop_out = arena_take<int32_t>(ctx, n0);
    if (!ctl || !tiles || !match || !d_count_out || !d_kept || !d_stop_out) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t st = ctx->stream;
    MergeCols cols;
    memset(&cols, 0, sizeof(cols));
    cols.S = S;
    const int32_t* d_start0 = nullptr;
    MergeCtl h0 = {n0, 0, 0, 0, 0};
    CG_CUDA(ctx, cudaMemcpyAsync(ctl, &h0, sizeof(MergeCtl), cudaMemcpyHostToDevice, st));
    std::vector<uint8_t*> d_chrom(S);
    std::vector<int32_t*> d_start(S);
    for (int s = 0; s < S; s++) {
        const size_t m = (size_t)n[s];
        d_chrom[s] = arena_take<uint8_t>(ctx, m);
        d_start[s] = arena_take<int32_t>(ctx, m);
        int32_t* d_stop = arena_take<int32_t>(ctx, m);
        float* d_count = arena_take<float>(ctx, m);
        unsigned long long* d_key = arena_take<unsigned long long>(ctx, m);
        if (!d_chrom[s] || !d_start[s] || !d_stop || !d_count || !d_key) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
        if (m > 0) {
            CG_CUDA(ctx, cudaMemcpyAsync(d_chrom[s], chrom[s], m, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d_start[s], start[s], m * 4, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d_stop, stop[s], m * 4, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d_count, count[s], m * 4, cudaMemcpyHostToDevice, st));
        }
        cols.key[s] = d_key; cols.stop[s] = d_stop; cols.count[s] = d_count; cols.n[s] = (int)m;
        if (s == 0) d_start0 = d_start[s];
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    for (int s = 0; s < S; s++)
        if (n[s] > 0)
            CG_LAUNCH(ctx, merge_key_kernel, std::max(1, std::min(div_up(n[s], 256), ctx->num_sms * 8)), 256, 0, d_chrom[s], d_start[s],
                      (int)n[s], const_cast<unsigned long long*>(cols.key[s]), ctl);
    if (S > 1) CG_LAUNCH(ctx, merge_match_kernel, std::max(1, std::min(div_up(n0, 256), ctx->num_sms * 8)), 256, 0, cols, match);
    {
        CommonPred p{match, n0, S};
        CommonEmit e{cols, match, d_start0, d_kept, d_stop_out, d_count_out, ctl};
        compact_run(ctx, p, e, &ctl->n0, n0, tiles, &ctl->n_out);
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    MergeCtl* h = (MergeCtl*)ctx->pinned;
    CG_CUDA(ctx, cudaMemcpyAsync(h, ctl, sizeof(MergeCtl), cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaStreamSynchronize(st));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    if (h->unsorted)
        return cg_fail(ctx, CG_ERR_UNSORTED, "cg_merge_common_bins: every sample must be ordered by (chromosome id, start) without duplicates");
    if (h->bad_start) return cg_fail(ctx, CG_ERR_ARG, "Start must be non-negative");
    if (h->bad_stop) return cg_fail(ctx, CG_ERR_ARG, "Start must be less than Stop");
    const int m = h->n_out;
    *n_out = m;
    if (m > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(kept_index, d_kept, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        CG_CUDA(ctx, cudaMemcpyAsync(stop_out, d_stop_out, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        for (int s = 0; s < S; s++)
            CG_CUDA(ctx, cudaMemcpyAsync(count_out + (size_t)s * n0, d_count_out + (size_t)s * n0, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        CG_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------
// cg_merge_kept_indices: the same merge for samples cleaned from ONE bin layout.  A pedigree is binned once
// (CanvasRunner.cs:846-870 hands the same bin definitions to every sample's CanvasBin), so a bin is identified by its
// index in that layout and cg_clean's kept_index lists are already the keys: a position table per sample
// (layout index -> row in that sample's list) replaces the key columns and the binary searches, and the host no longer
// gathers chromosome / start / stop columns per sample.
// ---------------------------------------------------------------------------------------------
namespace {

struct KeptCols {
    const int32_t* kept[MERGE_MAX_SAMPLES];
    const float* count[MERGE_MAX_SAMPLES];
    int* pos[MERGE_MAX_SAMPLES];  // [n_bins] row of the bin in sample s, -1 when that sample dropped it
    int n[MERGE_MAX_SAMPLES];
    int S;
};

__global__ void kept_scatter_kernel(const int32_t* __restrict__ kept, int n, int n_bins, int* __restrict__ pos, MergeCtl* ctl) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = kept[i];
        if (b < 0 || b >= n_bins || (i > 0 && kept[i - 1] >= b)) { ctl->unsorted = 1; continue; }
        pos[b] = i;
    }
}

struct KeptPred {
    KeptCols c;
    __device__ bool operator()(int i) const {
        const int b = c.kept[0][i];
        for (int s = 1; s < c.S; s++)
            if (c.pos[s][b] < 0) return false;
        return true;
    }
};

struct KeptEmit {
    KeptCols c;
    int32_t* common;
    float* count_out;  // [S][n0]
    __device__ void operator()(int src, int dst) const {
        const int n0 = c.n[0];
        const int b = c.kept[0][src];
        common[dst] = b;
        count_out[dst] = c.count[0][src];
        for (int s = 1; s < c.S; s++) count_out[(size_t)s * n0 + dst] = c.count[s][c.pos[s][b]];
    }
};

}  // namespace

// on_device: the lists and the outputs are device memory of this GPU (the pedigree chain, pedigree.cu) — nothing is staged,
// count_out rows are `out_stride` floats apart.  Otherwise host memory in and out, rows n_kept[0] apart (the C-ABI form).
int merge_kept_lists(cg_ctx* ctx, int64_t n_bins, int n_samples, const int64_t* n_kept, const int32_t* const* kept,
                     const float* const* count, int64_t* n_out, int32_t* common_index, float* count_out, bool on_device,
                     size_t out_stride) {
    if (!ctx) return CG_ERR_ARG;
    if (n_samples < 1 || n_bins < 0 || n_bins > 0x7fff0000LL || !n_kept || !kept || !count || !n_out)
        return cg_fail(ctx, CG_ERR_ARG, "cg_merge_kept_indices: bad argument");
    if (n_samples > MERGE_MAX_SAMPLES) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_merge_kept_indices: at most 8 samples");
    const int S = n_samples;
    for (int s = 0; s < S; s++) {
        if (n_kept[s] < 0 || n_kept[s] > n_bins) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_kept_indices: bad list length");
        if (n_kept[s] > 0 && (!kept[s] || !count[s])) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_kept_indices: null column");
    }
    *n_out = 0;
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    const int n0 = (int)n_kept[0];
    if (n0 == 0) return CG_OK;
    if (!common_index || !count_out) return cg_fail(ctx, CG_ERR_ARG, "cg_merge_kept_indices: null output");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t need = arena_need(1, sizeof(MergeCtl)) + arena_need(n0 / CMP_TILE + 2, 4) + arena_need((size_t)S * n0, 4) + arena_need(n0, 4) + (1u << 16);
    for (int s = 0; s < S; s++) need += arena_need(n_kept[s], 4) * 2 + arena_need(n_bins + 1, 4);
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    MergeCtl* ctl = arena_take<MergeCtl>(ctx, 1);
    int* tiles = arena_take<int>(ctx, n0 / CMP_TILE + 2);
    float* d_count_out = arena_take<float>(ctx, (size_t)S * n0);
    int32_t* d_common = arena_take<int32_t>(ctx, n0);
    if (!ctl || !tiles || !d_count_out || !d_common) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    if (!on_device) out_stride = (size_t)n0;
    cudaStream_t st = ctx->stream;
    KeptCols cols;
    memset(&cols, 0, sizeof(cols));
    cols.S = S;
    MergeCtl h0 = {n0, 0, 0, 0, 0};
    CG_CUDA(ctx, cudaMemcpyAsync(ctl, &h0, sizeof(MergeCtl), cudaMemcpyHostToDevice, st));
    for (int s = 0; s < S; s++) {
        const size_t m = (size_t)n_kept[s];
        int32_t* d_kept = arena_take<int32_t>(ctx, m);
        float* d_count = arena_take<float>(ctx, m);
        int* d_pos = arena_take<int>(ctx, n_bins + 1);
        if (!d_kept || !d_count || !d_pos) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
        if (m > 0 && !on_device) {
            CG_CUDA(ctx, cudaMemcpyAsync(d_kept, kept[s], m * 4, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d_count, count[s], m * 4, cudaMemcpyHostToDevice, st));
        }
        cols.kept[s] = on_device ? kept[s] : d_kept; cols.count[s] = on_device ? count[s] : d_count; cols.pos[s] = d_pos; cols.n[s] = (int)m;
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    for (int s = 0; s < S; s++) {
        if (s > 0) CG_CUDA(ctx, cudaMemsetAsync(cols.pos[s], 0xff, (size_t)(n_bins + 1) * 4, st));
        if (n_kept[s] > 0)
            CG_LAUNCH(ctx, kept_scatter_kernel, std::max(1, std::min(div_up(n_kept[s], 256), ctx->num_sms * 8)), 256, 0, cols.kept[s],
                      (int)n_kept[s], (int)n_bins, cols.pos[s], ctl);
    }
    {
        KeptPred p{cols};
        KeptEmit e{cols, d_common, d_count_out};
        compact_run(ctx, p, e, &ctl->n0, n0, tiles, &ctl->n_out);
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    MergeCtl* h = (MergeCtl*)ctx->pinned;
    CG_CUDA(ctx, cudaMemcpyAsync(h, ctl, sizeof(MergeCtl), cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaStreamSynchronize(st));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    if (h->unsorted) return cg_fail(ctx, CG_ERR_UNSORTED, "cg_merge_kept_indices: every list must be strictly increasing indices into the layout");
    const int m = h->n_out;
    *n_out = m;
    if (m > 0) {
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        CG_CUDA(ctx, cudaMemcpyAsync(common_index, d_common, (size_t)m * 4, kind, st));
        for (int s = 0; s < S; s++)
            CG_CUDA(ctx, cudaMemcpyAsync(count_out + (size_t)s * out_stride, d_count_out + (size_t)s * n0, (size_t)m * 4, kind, st));
        CG_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return CG_OK;
}

extern "C" int cg_merge_kept_indices(cg_ctx* ctx, int64_t n_bins, int n_samples, const int64_t* n_kept, const int32_t* const* kept,
                                     const float* const* count, int64_t* n_out, int32_t* common_index, float* count_out) {
    return merge_kept_lists(ctx, n_bins, n_samples, n_kept, kept, count, n_out, common_index, count_out, false, 0);
}
