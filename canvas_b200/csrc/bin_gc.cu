// CanvasBin, GCContentWeighted mode: the tables BinCountsForChromosome's weighted count needs (reference
// Src/Canvas/CanvasBin/CanvasBin.cs):
//   cg_bin_fragment_stats  Utilities.NonZeroMean(Int16[]) (CanvasCommon/Utilities.cs:136-151): sum and number of the positive
//                          fragment lengths of one chromosome (MeanFragmentSize, CanvasBin.cs:164-174, divides on the host)
//   cg_bin_read_gc         GC content of the "read" starting at every position (:450-497) and this chromosome's share of the
//                          expected / observed read counts per GC bin (ComputeObservedVsExpectedGC, :341-358)
// The reference counts the G/C bases of every fragment again from its first base (O(length x fragment size)); here one
// prefix sum of the G/C indicator makes every position a difference of two prefix values: 1 B (base) + 2 B (fragment
// length) + 1 B (hits) read, 1 B written per position, plus the 4-byte prefix array written once and read twice.
#include "clean.cuh"

namespace {

constexpr int GCT_THREADS = 256;
constexpr int GCT_ITEMS = 16;
constexpr int GCT_TILE = GCT_THREADS * GCT_ITEMS;
constexpr int GC_READ_BINS = 101;  // numberOfGCbins, CanvasBin.cs:114

__device__ inline int is_gc(char b) { return b == 'C' || b == 'c' || b == 'G' || b == 'g'; }

__global__ void __launch_bounds__(GCT_THREADS) gc_tile_count_kernel(const char* __restrict__ bases, long long len, int* __restrict__ tile_cnt) {
    const long long base = (long long)blockIdx.x * GCT_TILE + (long long)threadIdx.x * GCT_ITEMS;
    int c = 0;
#pragma unroll
    for (int t = 0; t < GCT_ITEMS; t++)
        if (base + t < len) c += is_gc(bases[base + t]);
    int total;
    block_excl_scan(c, total);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// prefix[i] = number of G/C bases before position i, i = 0 .. len
__global__ void __launch_bounds__(GCT_THREADS) gc_prefix_kernel(const char* __restrict__ bases, long long len, const int* __restrict__ tile_off,
                                                               unsigned* __restrict__ prefix) {
    const long long base = (long long)blockIdx.x * GCT_TILE + (long long)threadIdx.x * GCT_ITEMS;
    int f[GCT_ITEMS];
    int c = 0;
#pragma unroll
    for (int t = 0; t < GCT_ITEMS; t++) {
        f[t] = base + t < len ? is_gc(bases[base + t]) : 0;
        c += f[t];
    }
    int total;
    int run = tile_off[blockIdx.x] + block_excl_scan(c, total);
#pragma unroll
    for (int t = 0; t < GCT_ITEMS; t++) {
        if (base + t <= len) prefix[base + t] = (unsigned)run;
        run += f[t];
    }
}

// gcContent[pos] and the histogram rows of ComputeObservedVsExpectedGC
__global__ void __launch_bounds__(256) read_gc_kernel(const unsigned* __restrict__ prefix, const short* __restrict__ frag_len,
                                                      const unsigned char* __restrict__ hits, long long len, int mean_frag, int cutoff,
                                                      unsigned char* __restrict__ read_gc, unsigned long long* __restrict__ expected,
                                                      unsigned long long* __restrict__ observed) {
    __shared__ unsigned s_exp[GC_READ_BINS], s_obs[GC_READ_BINS];
    for (int t = threadIdx.x; t < GC_READ_BINS; t += blockDim.x) s_exp[t] = s_obs[t] = 0u;
    __syncthreads();
    const long long limit = len - (long long)mean_frag * cutoff - 1;  // pos < Bases.Length - mean * cutoff - 1 (:469)
    const int cap = mean_frag * cutoff;
    const long long per_block = 8192;  // positions per block: per-block counters stay far below 2^32
    const long long lo = (long long)blockIdx.x * per_block, hi = min(len, lo + per_block);
    for (long long pos = lo + threadIdx.x; pos < hi; pos += blockDim.x) {
        int g = 0;
        if (pos < limit) {
            const int f = frag_len[pos];
            const int cur = f == 0 ? mean_frag : min(f, cap);           // Convert.ToInt16(Math.Min(length, mean * cutoff))
            if (cur > 0) {
                const unsigned cnt = prefix[pos + cur] - prefix[pos];   // G/C bases of [pos, pos + cur)
                g = (int)min((long long)(100u * cnt) / (long long)cur, (long long)GC_READ_BINS);
            }
            // a negative length: the counting loop does not run and 0 / negative is 0
        }
        read_gc[pos] = (unsigned char)g;
        atomicAdd(&s_exp[g], 1u);
        const unsigned h = hits[pos];
        if (h) atomicAdd(&s_obs[g], h);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < GC_READ_BINS; t += blockDim.x) {
        if (s_exp[t]) atomicAdd(&expected[t], (unsigned long long)s_exp[t]);
        if (s_obs[t]) atomicAdd(&observed[t], (unsigned long long)s_obs[t]);
    }
}

// The same in one streaming pass when the longest read (mean * cutoff bases) fits a shared-memory halo: a block takes 8192
// positions, turns their bases plus the halo into a G/C bitmap with a running popcount per word (shared memory), and every
// thread computes the read GC of 32 consecutive positions from two bitmap lookups.  Neighbouring positions share all but two
// bases of their reads, so the GC bin changes slowly along a thread's run and the histogram updates are run-length
// aggregated (one shared-memory atomic per run instead of one per position).  Traffic = the algorithmic 5 B / position.
constexpr int RG_THREADS = 256;
constexpr int RG_PER = 32;
constexpr int RG_TILE = RG_THREADS * RG_PER;
constexpr int RG_HALO_MAX = 8192;
constexpr int RG_WORDS = (RG_TILE + RG_HALO_MAX) / 32 + 2;

__global__ void __launch_bounds__(RG_THREADS) read_gc_tile_kernel(const char* __restrict__ bases, const short* __restrict__ frag_len,
                                                                 const unsigned char* __restrict__ hits, long long len, int mean_frag, int cutoff,
                                                                 unsigned char* __restrict__ read_gc, unsigned long long* __restrict__ expected,
                                                                 unsigned long long* __restrict__ observed) {
    __shared__ unsigned s_bits[RG_WORDS];
    __shared__ unsigned s_pre[RG_WORDS];
    __shared__ unsigned s_warp[RG_THREADS / 32];
    __shared__ unsigned s_exp[GC_READ_BINS], s_obs[GC_READ_BINS];
    const int cap = mean_frag * cutoff;
    const long long limit = len - (long long)cap - 1;  // pos < Bases.Length - mean * cutoff - 1 (:469)
    const long long tile_start = (long long)blockIdx.x * RG_TILE;
    const int span = (int)(min(len, tile_start + RG_TILE + cap) - tile_start);  // bases this tile's reads can touch
    const int nwords = (span + 31) / 32 + 1;
    for (int t = threadIdx.x; t < GC_READ_BINS; t += blockDim.x) s_exp[t] = s_obs[t] = 0u;
    // ---- G/C bitmap of [tile_start, tile_start + span): 16 bases per 128-bit load -> 16 bits
    unsigned short* s_half = reinterpret_cast<unsigned short*>(s_bits);
    for (int i = threadIdx.x * 16; i < nwords * 32; i += RG_THREADS * 16) {
        unsigned bits16 = 0u;
        if (i + 16 <= span) {
            const uint4 v = *reinterpret_cast<const uint4*>(bases + tile_start + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned x = w[q] | 0x20202020u;
                const unsigned eq = (__vcmpeq4(x, 0x63636363u) | __vcmpeq4(x, 0x67676767u)) & 0x01010101u;
                bits16 |= ((eq * 0x10204080u) >> 28) << (4 * q);  // bit 0 of byte j -> bit j
            }
        } else {
            for (int j = 0; j < 16 && i + j < span; j++) bits16 |= (unsigned)is_gc(bases[tile_start + i + j]) << j;
        }
        s_half[i >> 4] = (unsigned short)bits16;
    }
    __syncthreads();
    // ---- s_pre[w] = G/C bases before word w
    {
        constexpr int WPT = (RG_WORDS + RG_THREADS - 1) / RG_THREADS;
        unsigned c[WPT], run = 0;
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            const int w = threadIdx.x * WPT + j;
            c[j] = w < nwords ? (unsigned)__popc(s_bits[w]) : 0u;
            run += c[j];
        }
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        unsigned base = incl - run;
        for (int k = 0; k < wid; k++) base += s_warp[k];
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            const int w = threadIdx.x * WPT + j;
            if (w < RG_WORDS) s_pre[w] = base;
            base += c[j];
        }
    }
    __syncthreads();
    // ---- 32 consecutive positions per thread
    const int a0 = threadIdx.x * RG_PER;
    const long long p0 = tile_start + a0;
    if (p0 < len) {
        const bool whole = p0 + RG_PER <= len;
        unsigned fw[16], hw[8], out[8];  // 32 lengths (two per word), 32 hit counts, 32 results
        if (whole) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 v = *reinterpret_cast<const uint4*>(frag_len + p0 + 8 * q);
                fw[4 * q] = v.x; fw[4 * q + 1] = v.y; fw[4 * q + 2] = v.z; fw[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint4 v = *reinterpret_cast<const uint4*>(hits + p0 + 16 * q);
                hw[4 * q] = v.x; hw[4 * q + 1] = v.y; hw[4 * q + 2] = v.z; hw[4 * q + 3] = v.w;
            }
        }
        int run_g = -1;
        unsigned run_n = 0, run_h = 0;
        auto count_before = [&](int a) { return s_pre[a >> 5] + (unsigned)__popc(s_bits[a >> 5] & ((1u << (a & 31)) - 1u)); };
#pragma unroll
        for (int i = 0; i < RG_PER; i++) {
            const long long pos = p0 + i;
            if (!whole && pos >= len) break;
            const int f = whole ? (int)(short)((fw[i >> 1] >> (16 * (i & 1))) & 0xffffu) : (int)frag_len[pos];
            const unsigned h = whole ? (hw[i >> 2] >> (8 * (i & 3))) & 0xffu : (unsigned)hits[pos];
            int g = 0;
            if (pos < limit) {
                const int cur = f == 0 ? mean_frag : min(f, cap);  // Convert.ToInt16(Math.Min(length, mean * cutoff))
                if (cur > 0) {
                    const unsigned cnt = count_before(a0 + i + cur) - count_before(a0 + i);  // G/C bases of [pos, pos + cur)
                    g = (int)min((100u * cnt) / (unsigned)cur, (unsigned)GC_READ_BINS);
                }
            }
            if (whole) {
                if ((i & 3) == 0) out[i >> 2] = 0u;
                out[i >> 2] |= (unsigned)g << (8 * (i & 3));
            } else {
                read_gc[pos] = (unsigned char)g;
            }
            if (g == run_g) { run_n++; run_h += h; }
            else {
                if (run_n) { atomicAdd(&s_exp[run_g], run_n); if (run_h) atomicAdd(&s_obs[run_g], run_h); }
                run_g = g; run_n = 1u; run_h = h;
            }
        }
        if (run_n) { atomicAdd(&s_exp[run_g], run_n); if (run_h) atomicAdd(&s_obs[run_g], run_h); }
        if (whole) {
            *reinterpret_cast<uint4*>(read_gc + p0) = make_uint4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<uint4*>(read_gc + p0 + 16) = make_uint4(out[4], out[5], out[6], out[7]);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < GC_READ_BINS; t += blockDim.x) {
        if (s_exp[t]) atomicAdd(&expected[t], (unsigned long long)s_exp[t]);
        if (s_obs[t]) atomicAdd(&observed[t], (unsigned long long)s_obs[t]);
    }
}

__global__ void __launch_bounds__(256) frag_stats_kernel(const short* __restrict__ frag_len, long long len, unsigned long long* __restrict__ out) {
    unsigned long long sum = 0, cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
        const int v = frag_len[i];
        if (v > 0) { sum += (unsigned long long)v; cnt++; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&out[0], sum); atomicAdd(&out[1], cnt); }
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

extern "C" int cg_bin_fragment_stats(cg_ctx* ctx, int64_t len, const int16_t* frag_len, int64_t* sum, int64_t* count) {
    if (!ctx) return CG_ERR_ARG;
    if (len < 0 || len > 0x7fff0000LL || !sum || !count || (len > 0 && !frag_len)) return cg_fail(ctx, CG_ERR_ARG, "cg_bin_fragment_stats: bad argument");
    reset_call(ctx);
    *sum = *count = 0;
    if (len == 0) return CG_OK;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = arena_reserve(ctx, arena_need(len, 2) + (1 << 16));
    if (rc) return rc;
    short* d_f = arena_take<short>(ctx, len);
    unsigned long long* d_out = arena_take<unsigned long long>(ctx, 2);
    if (!d_f || !d_out) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_f, frag_len, (size_t)len * 2, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_out, 0, 16, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_LAUNCH(ctx, frag_stats_kernel, (int)std::min<long long>(ctx->num_sms * 8, div_up((long long)len, 256LL)), 256, 0, d_f, (long long)len, d_out);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    unsigned long long h[2] = {0, 0};
    CG_CUDA(ctx, cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CHECK_LAUNCHES(ctx);
    *sum = (int64_t)h[0];
    *count = (int64_t)h[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}

extern "C" int cg_bin_read_gc(cg_ctx* ctx, int64_t len, const char* bases, const int16_t* frag_len, int mean_frag, const uint8_t* hits,
                              uint8_t* read_gc, int64_t* expected, int64_t* observed) {
    if (!ctx) return CG_ERR_ARG;
    if (len < 0 || len > 0x7fff0000LL || !expected || !observed || (len > 0 && (!bases || !frag_len || !hits || !read_gc)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_bin_read_gc: bad argument");
    // the reference refuses a non-positive mean (CanvasBin.cs:432-437); above 10922 its Convert.ToInt16(mean * 3) can overflow
    if (mean_frag <= 0 || mean_frag > 10922) return cg_fail(ctx, CG_ERR_ARG, "cg_bin_read_gc: mean fragment size must be in 1..10922");
    reset_call(ctx);
    if (len == 0) return CG_OK;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int cutoff = 3;  // meanFragmentCutoff, :427
    const int ntiles = (int)div_up((long long)len + 1, (long long)GCT_TILE);
    int rc = arena_reserve(ctx, arena_need(len, 1) * 3 + arena_need(len, 2) + arena_need(len + 1, 4) + arena_need(ntiles + 1, 4) +
                                    arena_need(GC_READ_BINS, 8) * 2 + (1 << 16));
    if (rc) return rc;
    char* d_bases = arena_take<char>(ctx, len);
    unsigned char* d_hits = arena_take<unsigned char>(ctx, len);
    unsigned char* d_gc = arena_take<unsigned char>(ctx, len);
    short* d_f = arena_take<short>(ctx, len);
    unsigned* d_prefix = arena_take<unsigned>(ctx, len + 1);
    int* d_tiles = arena_take<int>(ctx, ntiles + 1);
    int* d_total = arena_take<int>(ctx, 1);
    unsigned long long* d_exp = arena_take<unsigned long long>(ctx, GC_READ_BINS);
    unsigned long long* d_obs = arena_take<unsigned long long>(ctx, GC_READ_BINS);
    if (!d_bases || !d_hits || !d_gc || !d_f || !d_prefix || !d_tiles || !d_total || !d_exp || !d_obs) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_bases, bases, (size_t)len, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_hits, hits, (size_t)len, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_f, frag_len, (size_t)len * 2, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_exp, 0, GC_READ_BINS * 8, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_obs, 0, GC_READ_BINS * 8, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    if (mean_frag * cutoff <= RG_HALO_MAX) {
        // the longest read fits the shared-memory halo: one streaming pass, no prefix array
        CG_LAUNCH(ctx, read_gc_tile_kernel, (int)div_up((long long)len, (long long)RG_TILE), RG_THREADS, 0, d_bases, d_f, d_hits, (long long)len,
                  mean_frag, cutoff, d_gc, d_exp, d_obs);
    } else {
        CG_LAUNCH(ctx, gc_tile_count_kernel, ntiles, GCT_THREADS, 0, d_bases, (long long)len, d_tiles);
        CG_LAUNCH(ctx, compact_scan_kernel, 1, 1024, 0, d_tiles, ntiles, d_total);
        CG_LAUNCH(ctx, gc_prefix_kernel, ntiles, GCT_THREADS, 0, d_bases, (long long)len, d_tiles, d_prefix);
        CG_LAUNCH(ctx, read_gc_kernel, (int)div_up((long long)len, 8192LL), 256, 0, d_prefix, d_f, d_hits, (long long)len, mean_frag, cutoff, d_gc,
                  d_exp, d_obs);
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    unsigned long long h_exp[GC_READ_BINS], h_obs[GC_READ_BINS];
    CG_CUDA(ctx, cudaMemcpyAsync(read_gc, d_gc, (size_t)len, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(h_exp, d_exp, GC_READ_BINS * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(h_obs, d_obs, GC_READ_BINS * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    for (int g = 0; g < GC_READ_BINS; g++) { expected[g] += (int64_t)h_exp[g]; observed[g] += (int64_t)h_obs[g]; }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}
