// CanvasBin counting kernels (reference Src/Canvas/CanvasBin/):
//   cg_bin_screen     ExcludeTagsOverlappingFilterFile + ScreenObservedTags + the counts of GetRates, CanvasBin.cs:668-716, :56-58
//   cg_bin_hits       BinCountsForChromosome, CanvasBin.cs:568-661 (bins of `bin_size` possible positions;
//                     TruncatedDynamicRange :618-625 and GCContentWeighted :626-636 counts, GC% :638)
//   cg_bin_fragments  FragmentBinner.BinOneAlignment / FindBestBin, FragmentBinner.cs:296-311, :353-371
// BAM decoding, read-name pairing and the unique-kmer FASTA stay on the host; integers are bit-exact.
#include "common.cuh"

namespace {

constexpr int BIN_TILE_WORDS = 1024;  // 64-bit words of the possible-alignment bitmap per scan tile

struct BinCtl {
    unsigned long long first_pos;   // first base that is not 'n' (CanvasBin.cs:582-583)
    unsigned long long total_possible;
    int n_bins;
    int pad;
};

__global__ void bin_init_kernel(BinCtl* ctl) {
    ctl->first_pos = ~0ull;
    ctl->total_possible = 0ull;
    ctl->n_bins = 0;
}

// "Skip past leading Ns": the first position whose base is not a lower-case 'n' (:582-583).  Threads take 16 bases per
// step; a thread stops as soon as its position lies beyond the best answer so far.
__global__ void __launch_bounds__(256) bin_first_base_kernel(const char* __restrict__ bases, long long len, BinCtl* ctl) {
    const long long stride = (long long)gridDim.x * blockDim.x * 16;
    unsigned long long best = ~0ull;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < len; i += stride) {
        if ((unsigned long long)i > *(volatile unsigned long long*)&ctl->first_pos) break;
        if (i + 16 <= len) {
            const uint4 v = *reinterpret_cast<const uint4*>(bases + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            int at = -1;
#pragma unroll
            for (int q = 3; q >= 0; q--) {
                const unsigned ne = ~__vcmpeq4(w[q], 0x6e6e6e6eu);  // 0xff where the base is not 'n'
                if (ne) at = 4 * q + ((__ffs((int)ne) - 1) >> 3);
            }
            if (at >= 0) { best = (unsigned long long)(i + at); break; }
        } else {
            for (long long j = i; j < len; j++)
                if (bases[j] != 'n') { best = (unsigned long long)j; break; }
            break;
        }
    }
    if (best != ~0ull) atomicMin(&ctl->first_pos, best);
}

__device__ inline unsigned long long masked_word(const unsigned long long* __restrict__ bits, long long w, long long nwords,
                                                  unsigned long long first_pos, long long len) {
    if (w >= nwords) return 0ull;
    unsigned long long v = bits[w];
    const long long lo = w * 64;
    if ((long long)first_pos > lo) {
        const long long k = (long long)first_pos - lo;
        v = k >= 64 ? 0ull : (v & (~0ull << k));
    }
    if (lo + 64 > len) {
        const long long k = len - lo;
        v = k <= 0 ? 0ull : (v & (~0ull >> (64 - k)));
    }
    return v;
}

// possible positions per tile of the bitmap (only positions >= first_pos count)
__global__ void bin_tile_count_kernel(const unsigned long long* __restrict__ bits, long long nwords, long long len,
                                      const BinCtl* __restrict__ ctl, unsigned* __restrict__ tile_cnt) {
    __shared__ unsigned s_w[32];
    const unsigned long long first = min(ctl->first_pos, (unsigned long long)len);  // all 'n': nothing is possible
    const long long w0 = (long long)blockIdx.x * BIN_TILE_WORDS;
    unsigned c = 0;
    for (int i = threadIdx.x; i < BIN_TILE_WORDS; i += blockDim.x) c += __popcll(masked_word(bits, w0 + i, nwords, first, len));
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += s_w[k];
        tile_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts (single block) -> tile_off (u64); number of complete bins
__global__ void bin_tile_scan_kernel(const unsigned* __restrict__ tile_cnt, int ntiles, unsigned long long* __restrict__ tile_off,
                                     int bin_size, BinCtl* ctl) {
    __shared__ unsigned long long s_carry;
    __shared__ unsigned long long s_w[32];
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < ntiles ? tile_cnt[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_w[w] = incl;
        __syncthreads();
        unsigned long long wb = 0;
        for (int k = 0; k < w; k++) wb += s_w[k];
        if (i < ntiles) tile_off[i] = s_carry + wb + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += wb + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_off[ntiles] = s_carry;
        ctl->total_possible = s_carry;
        ctl->n_bins = bin_size > 0 ? (int)(s_carry / (unsigned long long)bin_size) : 0;
    }
}

// end position of bin k = position of the ((k+1) * bin_size)-th possible position (:611-612).  One block per tile of the
// bitmap: the tile's masked words and their running popcount are staged in shared memory, then every bin that ends inside
// the tile finds its word by binary search there and its bit with __fns.
__global__ void __launch_bounds__(256) bin_end_tile_kernel(const unsigned long long* __restrict__ bits, long long nwords, long long len,
                                                          const unsigned long long* __restrict__ tile_off, int bin_size,
                                                          const BinCtl* __restrict__ ctl, int max_bins, int* __restrict__ end_pos) {
    __shared__ unsigned long long s_w[BIN_TILE_WORDS];
    __shared__ unsigned s_pre[BIN_TILE_WORDS];  // possible positions of the tile before word i
    __shared__ unsigned s_warp[8];
    const int t = blockIdx.x;
    const unsigned long long lo_rank = tile_off[t], hi_rank = tile_off[t + 1];
    const unsigned long long B = (unsigned long long)bin_size;
    const long long nb = min(ctl->n_bins, max_bins);
    const long long k_first = (long long)(lo_rank / B);               // first bin whose closing rank (k+1) B lies beyond lo_rank
    const long long k_last = min((long long)(hi_rank / B) - 1, nb - 1);  // last bin whose closing rank is <= hi_rank
    if (k_last < k_first) return;
    const unsigned long long first = min(ctl->first_pos, (unsigned long long)len);
    const long long w0 = (long long)t * BIN_TILE_WORDS;
    // each thread owns four consecutive words: popcounts, then a block-wide exclusive scan
    unsigned c4[4], run = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = threadIdx.x * 4 + j;
        const unsigned long long w = masked_word(bits, w0 + i, nwords, first, len);
        s_w[i] = w;
        c4[j] = (unsigned)__popcll(w);
        run += c4[j];
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
    for (int j = 0; j < 4; j++) { s_pre[threadIdx.x * 4 + j] = base; base += c4[j]; }
    __syncthreads();
    for (long long k = k_first + threadIdx.x; k <= k_last; k += blockDim.x) {
        const unsigned need = (unsigned)((unsigned long long)(k + 1) * B - lo_rank);  // 1-based rank inside the tile
        int lo = 0, hi = BIN_TILE_WORDS - 1;  // last word with s_pre < need
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[mid] < need) lo = mid; else hi = mid - 1; }
        const unsigned long long w = s_w[lo];
        unsigned r = need - s_pre[lo];  // r-th set bit of w
        const unsigned wl = (unsigned)w, wh = (unsigned)(w >> 32);
        const unsigned cl = (unsigned)__popc(wl);
        const int bit = r <= cl ? (int)__fns(wl, 0, (int)r) : 32 + (int)__fns(wh, 0, (int)(r - cl));
        end_pos[k] = (int)((w0 + lo) * 64 + bit);
    }
}

// Per-bin sums in one streaming pass over the positions (TruncatedDynamicRange count :618-625 and the G/C bases of :592-593):
// a block takes 4096 consecutive positions, a thread 16 of them (one 128-bit load of hits and of bases, 16 bits of the
// bitmap); bin ends that fall inside the tile are marked in a shared bitmap, a block scan of their counts tells every
// thread the bin of its first position, four positions are summed at a time with byte-wise SIMD, and the per-bin partial
// sums go through shared-memory accumulators to integer atomics in global memory (integers: any order gives the same sums).
constexpr int BIN_ACC_THREADS = 256;
constexpr int BIN_ACC_PER_THREAD = 32;  // positions per thread: two 128-bit loads of hits and of bases, 32 bits of the bitmap
constexpr int BIN_ACC_TILE = BIN_ACC_THREADS * BIN_ACC_PER_THREAD;
constexpr int BIN_ACC_LOCAL = 1024;  // bins of one tile accumulated in shared memory (more go straight to global atomics)

__device__ inline void bin_acc_add(unsigned* s_obs, unsigned* s_gc, unsigned* g_obs, unsigned* g_gc, long long k_lo, long long bin, long long nb,
                                   unsigned obs, unsigned gcv) {
    if (bin >= nb || (obs | gcv) == 0u) return;
    const long long loc = bin - k_lo;
    if (loc < BIN_ACC_LOCAL) {
        if (obs) atomicAdd(&s_obs[loc], obs);
        if (gcv) atomicAdd(&s_gc[loc], gcv);
    } else {
        if (obs) atomicAdd(&g_obs[bin], obs);
        if (gcv) atomicAdd(&g_gc[bin], gcv);
    }
}

// four positions at once: h = four hit counts, b = four bases, m4 = their four possible bits
__device__ inline void bin_quad(unsigned h, unsigned b, unsigned m4, unsigned& obs, unsigned& gcv) {
    const unsigned mask = ((m4 * 0x00204081u) & 0x01010101u) * 0xffu;  // bit j -> byte j = 0xff
    obs += __vsadu4(__vminu4(h, 0x0a0a0a0au) & mask, 0u);
    const unsigned x = b | 0x20202020u;
    gcv += (unsigned)__popc((__vcmpeq4(x, 0x63636363u) | __vcmpeq4(x, 0x67676767u)) & 0x01010101u);
}

// the same restricted to the positions of the quad selected by the 4-bit mask `part`
__device__ inline void bin_quad_part(unsigned h, unsigned b, unsigned m4, unsigned part, unsigned& obs, unsigned& gcv) {
    const unsigned pmask = ((part * 0x00204081u) & 0x01010101u) * 0xffu;
    const unsigned mask = ((m4 * 0x00204081u) & 0x01010101u) * 0xffu & pmask;
    obs += __vsadu4(__vminu4(h, 0x0a0a0a0au) & mask, 0u);
    const unsigned x = b | 0x20202020u;
    gcv += (unsigned)__popc((__vcmpeq4(x, 0x63636363u) | __vcmpeq4(x, 0x67676767u)) & 0x01010101u & pmask);
}

// first bin of every tile of the streaming pass: tile_first[t] = number of bins that end before position t * BIN_ACC_TILE.
// Bin k is that bin for the tiles whose start lies in (end[k-1], end[k]]; thread nb fills the tiles past the last bin.
__global__ void bin_tile_first_kernel(const BinCtl* __restrict__ ctl, const int* __restrict__ end_pos, int max_bins, long long n_tiles,
                                      int* __restrict__ tile_first) {
    const long long nb = min(ctl->n_bins, max_bins);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nb) return;
    const long long prev = k == 0 ? -1 : (long long)end_pos[k - 1];
    const long long t_lo = prev < 0 ? 0 : prev / BIN_ACC_TILE + 1;
    const long long t_hi = k == nb ? n_tiles : (long long)end_pos[k] / BIN_ACC_TILE;  // inclusive
    for (long long t = t_lo; t <= t_hi && t <= n_tiles; t++) tile_first[t] = (int)k;
}

__global__ void __launch_bounds__(BIN_ACC_THREADS) bin_accum_kernel(const unsigned char* __restrict__ hits, const unsigned long long* __restrict__ bits,
                                                                   const char* __restrict__ bases, long long nwords, long long len,
                                                                   const BinCtl* __restrict__ ctl, const int* __restrict__ end_pos, int max_bins,
                                                                   const int* __restrict__ tile_first, unsigned* __restrict__ g_obs,
                                                                   unsigned* __restrict__ g_gc) {
    __shared__ unsigned s_flag[BIN_ACC_THREADS];  // one word of end flags per thread (32 positions)
    __shared__ unsigned s_obs[BIN_ACC_LOCAL], s_gc[BIN_ACC_LOCAL];
    __shared__ unsigned s_warp[BIN_ACC_THREADS / 32];
    const long long nb = min(ctl->n_bins, max_bins);
    if (nb <= 0) return;
    const long long tile_start = (long long)blockIdx.x * BIN_ACC_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long p0 = tile_start + (long long)threadIdx.x * BIN_ACC_PER_THREAD;
    // bins k_lo .. k_hi - 1 end inside the tile, bin k_hi continues past it
    const long long k_lo = tile_first[blockIdx.x], k_hi = tile_first[blockIdx.x + 1];
    if (k_lo >= nb) return;  // the whole tile lies past the last complete bin
    // the loads of this thread's 32 positions do not depend on the bin lookup: issue them first
    const bool whole = p0 + BIN_ACC_PER_THREAD <= len;
    uint4 h4[2], b4[2];
    h4[0] = h4[1] = b4[0] = b4[1] = make_uint4(0u, 0u, 0u, 0u);
    if (whole) {
        h4[0] = *reinterpret_cast<const uint4*>(hits + p0);
        h4[1] = *reinterpret_cast<const uint4*>(hits + p0 + 16);
        b4[0] = *reinterpret_cast<const uint4*>(bases + p0);
        b4[1] = *reinterpret_cast<const uint4*>(bases + p0 + 16);
    }
    unsigned m32 = 0u;
    if (p0 < len) m32 = (unsigned)(masked_word(bits, p0 >> 6, nwords, min(ctl->first_pos, (unsigned long long)len), len) >> (p0 & 63));
    s_flag[threadIdx.x] = 0u;
    const long long nloc = min((long long)BIN_ACC_LOCAL, min(k_hi + 1, nb) - k_lo);
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) { s_obs[i] = 0u; s_gc[i] = 0u; }
    __syncthreads();
    for (long long k = k_lo + threadIdx.x; k < k_hi; k += blockDim.x) {
        const int r = (int)((long long)end_pos[k] - tile_start);
        atomicOr(&s_flag[r >> 5], 1u << (r & 31));
    }
    __syncthreads();
    // ends inside this thread's positions; position p belongs to bin k_lo + (ends of the tile before p)
    const unsigned f32 = s_flag[threadIdx.x];
    const unsigned mine = (unsigned)__popc(f32);
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned before = incl - mine;
    for (int k = 0; k < wid; k++) before += s_warp[k];
    long long bin = k_lo + before;
    unsigned obs = 0, gcv = 0;
    if (p0 < len && bin < nb) {
        if (whole) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const unsigned hq[4] = {h4[half].x, h4[half].y, h4[half].z, h4[half].w};
                const unsigned bq[4] = {b4[half].x, b4[half].y, b4[half].z, b4[half].w};
                const unsigned m16 = (m32 >> (16 * half)) & 0xffffu, f16 = (f32 >> (16 * half)) & 0xffffu;
                if (f16 == 0u) {
#pragma unroll
                    for (int q = 0; q < 4; q++) bin_quad(hq[q], bq[q], (m16 >> (4 * q)) & 0xfu, obs, gcv);
                } else {
                    // a bin closes inside these 16 positions (one thread in thirty, but most warps have such a thread, so this
                    // branch has to stay short): quads without an end as above, a quad with one end as two masked halves
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const unsigned fq = (f16 >> (4 * q)) & 0xfu, mq = (m16 >> (4 * q)) & 0xfu;
                        if (fq == 0u) {
                            bin_quad(hq[q], bq[q], mq, obs, gcv);
                        } else if ((fq & (fq - 1u)) == 0u) {
                            const unsigned upto = (fq << 1) - 1u;  // positions of the quad up to and including the end
                            bin_quad_part(hq[q], bq[q], mq, upto, obs, gcv);
                            bin_acc_add(s_obs, s_gc, g_obs, g_gc, k_lo, bin, nb, obs, gcv);
                            obs = 0; gcv = 0; bin++;
                            bin_quad_part(hq[q], bq[q], mq, ~upto & 0xfu, obs, gcv);
                        } else {
                            for (int j = 0; j < 4; j++) {  // bins shorter than four positions
                                const unsigned h = (hq[q] >> (8 * j)) & 0xffu, b = ((bq[q] >> (8 * j)) & 0xffu) | 0x20u;
                                if ((mq >> j) & 1u) obs += min(10u, h);
                                gcv += (b == 0x63u || b == 0x67u) ? 1u : 0u;
                                if ((fq >> j) & 1u) {
                                    bin_acc_add(s_obs, s_gc, g_obs, g_gc, k_lo, bin, nb, obs, gcv);
                                    obs = 0; gcv = 0; bin++;
                                }
                            }
                        }
                    }
                }
            }
        } else {
            for (int i = 0; i < BIN_ACC_PER_THREAD && p0 + i < len; i++) {  // ragged end of the chromosome
                const unsigned h = hits[p0 + i], b = (unsigned)(unsigned char)bases[p0 + i] | 0x20u;
                if ((m32 >> i) & 1u) obs += min(10u, h);
                gcv += (b == 0x63u || b == 0x67u) ? 1u : 0u;
                if ((f32 >> i) & 1u) {
                    bin_acc_add(s_obs, s_gc, g_obs, g_gc, k_lo, bin, nb, obs, gcv);
                    obs = 0; gcv = 0; bin++;
                }
            }
        }
    }
    // what is left belongs to `bin`; lanes of a warp mostly share it: one shared-memory atomic per distinct bin of the warp
    {
        const bool live = bin < nb && (obs | gcv) != 0u;
        unsigned todo = __ballot_sync(0xffffffffu, live);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const long long lb = __shfl_sync(0xffffffffu, bin, leader);
            const bool same = live && bin == lb;
            const unsigned grp = __ballot_sync(0xffffffffu, same);
            const unsigned so = __reduce_add_sync(0xffffffffu, same ? obs : 0u);
            const unsigned sg = __reduce_add_sync(0xffffffffu, same ? gcv : 0u);
            if (lane == leader) bin_acc_add(s_obs, s_gc, g_obs, g_gc, k_lo, lb, nb, so, sg);
            todo &= ~grp;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
        const unsigned o = s_obs[i], g = s_gc[i];
        if (o) atomicAdd(&g_obs[k_lo + i], o);
        if (g) atomicAdd(&g_gc[k_lo + i], g);
    }
}

// bin coordinates, count and GC percentage from the accumulated sums: gc = (int)(100f * GC / nucleotides) with every base
// counted as a nucleotide (:592-593, :638); stop = end + 1 "to conform to bed specification" (:652)
__global__ void bin_finalize_kernel(const BinCtl* __restrict__ ctl, const int* __restrict__ end_pos, int max_bins, const unsigned* __restrict__ g_obs,
                                    const unsigned* __restrict__ g_gc, int32_t* __restrict__ start, int32_t* __restrict__ stop,
                                    int32_t* __restrict__ count, unsigned char* __restrict__ gc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = min(ctl->n_bins, max_bins);
    if (k >= nb) return;
    const int s = k == 0 ? (int)ctl->first_pos : end_pos[k - 1] + 1;
    const int e = end_pos[k];
    start[k] = s;
    stop[k] = e + 1;
    count[k] = (int)g_obs[k];
    gc[k] = (unsigned char)(int)__fdiv_rn(__fmul_rn(100.0f, (float)g_gc[k]), (float)(e - s + 1));
}

// GCContentWeighted (:626-636): float accumulation in position order, Math.Round (half to even).  One warp per bin reads the
// positions coalesced (a lane per position, four rows in flight); terms that are exactly zero cannot change the running sum,
// the others are added one by one in position order by every lane alike.
__global__ void __launch_bounds__(256) bin_sum_weighted_kernel(const unsigned char* __restrict__ hits, const unsigned long long* __restrict__ bits,
                                                               const unsigned char* __restrict__ read_gc, const float* __restrict__ obs_vs_exp,
                                                               const BinCtl* __restrict__ ctl, const int* __restrict__ end_pos, int max_bins,
                                                               int32_t* __restrict__ count) {
    __shared__ float s_ratio[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_ratio[i] = i < 101 ? obs_vs_exp[i] : 1.0f;
    __syncthreads();
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nb = min(ctl->n_bins, max_bins);
    if (k >= nb) return;
    const int s = k == 0 ? (int)ctl->first_pos : end_pos[k - 1] + 1;
    const int e = end_pos[k];
    float acc = 0.0f;
    constexpr int U = 4;
    for (int base = s; base <= e; base += 32 * U) {
        float term[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int p = base + 32 * u + lane;
            term[u] = 0.0f;
            if (p <= e && ((bits[p >> 6] >> (p & 63)) & 1ull))
                term[u] = fminf(10.0f, __fdiv_rn((float)hits[p], s_ratio[min((int)read_gc[p], 127)]));
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            unsigned nz = __ballot_sync(0xffffffffu, !(term[u] == 0.0f));
            while (nz) {
                const int l = __ffs(nz) - 1;
                acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, term[u], l));
                nz &= nz - 1;
            }
        }
    }
    if (lane == 0) count[k] = (int)rint((double)acc);
}

// FragmentBinner.cs:296-311 + FindBestBin :353-371: first bin whose Stop is right of the fragment start,
// then the bin with the largest overlap (the first one on ties)
__global__ void bin_fragments_kernel(const int32_t* __restrict__ fstart, const int32_t* __restrict__ fstop, long long nfrag,
                                     const int32_t* __restrict__ bstart, const int32_t* __restrict__ bstop, int nbins,
                                     int32_t* __restrict__ best_bin, int32_t* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfrag) return;
    const int fs = fstart[i], fe = fstop[i];
    int lo = 0, hi = nbins;  // first bin with stop > fs
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (bstop[mid] <= fs) lo = mid + 1; else hi = mid; }
    int best = -1, best_ov = 0;
    for (int b = lo; b < nbins; b++) {
        const int ov = min(bstop[b], fe) - max(bstart[b], fs);
        if (ov <= 0) break;
        if (ov > best_ov) { best_ov = ov; best = b; }
    }
    best_bin[i] = best;
    if (best >= 0) atomicAdd(&count[best], 1);
}

// a fragment whose mate later fails the duplicate / QC / MAPQ filters is taken out again (:279-284)
__global__ void bin_fragments_undo_kernel(const int32_t* __restrict__ undo, long long nundo, const int32_t* __restrict__ best_bin,
                                          int32_t* __restrict__ count) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nundo) return;
    const int b = best_bin[undo[j]];
    if (b >= 0) atomicSub(&count[b], 1);
}

// ---------------------------------------------------------------------------------------------
// cg_bin_screen: the passes over the genome positions that precede the binning (CanvasBin.cs:777-780, :30-76).
// ---------------------------------------------------------------------------------------------
// ExcludeTagsOverlappingFilterFile (:668-692): possible[i] = false for i in [start, stop).  One warp per interval:
// whole words are stored as zero (idempotent when intervals overlap), the two edge words are cleared atomically.
__global__ void bin_filter_clear_kernel(unsigned long long* __restrict__ bits, const int32_t* __restrict__ fstart,
                                        const int32_t* __restrict__ fstop, long long n_filter) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_filter) return;
    const long long a = fstart[w], b = fstop[w];  // validated by the host: 0 <= a, b <= length
    if (a >= b) return;
    const long long wa = a >> 6, wb = (b - 1) >> 6;
    const unsigned long long head = ~0ull << (a & 63);                    // bits >= a inside word wa
    const unsigned long long tail = ~0ull >> (63 - ((b - 1) & 63));       // bits <= b-1 inside word wb
    if (wa == wb) {
        if (lane == 0) atomicAnd(&bits[wa], ~(head & tail));
        return;
    }
    if (lane == 0) atomicAnd(&bits[wa], ~head);
    if (lane == 1) atomicAnd(&bits[wb], ~tail);
    for (long long k = wa + 1 + lane; k < wb; k += 32) bits[k] = 0ull;
}

// ScreenObservedTags (:699-716): hits[i] = 0 where position i is not possible; and the two counts of GetRates (:56-58):
// positions with a hit left (HitArray.CountSetBits, HitArray.cs:24-32) and possible positions.  A thread takes one
// 64-position word: 8 eight-byte loads of hits, masked by the word's bits spread to bytes.
__global__ void __launch_bounds__(256) bin_screen_kernel(unsigned long long* __restrict__ hits8, const unsigned long long* __restrict__ bits,
                                                         long long nwords, long long len, unsigned long long* __restrict__ counts) {
    unsigned long long obs = 0, pos = 0;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (long long)gridDim.x * blockDim.x) {
        unsigned long long m = bits[w];
        const long long base = w << 6;
        if (base + 64 > len) m &= (len - base >= 64) ? ~0ull : ((1ull << (len - base)) - 1ull);  // bits past the end do not exist
        pos += (unsigned long long)__popcll(m);
        if (base + 64 <= len) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned b8 = (unsigned)(m >> (8 * k)) & 0xffu;
                // spread 8 bits to 8 byte masks: bit j -> byte j = 0xff
                unsigned long long spread = (unsigned long long)b8 * 0x0101010101010101ull & 0x8040201008040201ull;
                spread = ((spread + 0x7f7f7f7f7f7f7f7full) >> 7) & 0x0101010101010101ull;  // non-zero byte -> 1
                const unsigned long long mask = spread * 0xffull;
                const unsigned long long h = hits8[(base >> 3) + k];
                const unsigned long long kept = h & mask;
                if (kept != h) hits8[(base >> 3) + k] = kept;
                // count non-zero bytes
                unsigned long long nz = ((kept & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | kept;
                obs += (unsigned long long)__popcll(nz & 0x8080808080808080ull);
            }
        } else {
            unsigned char* hb = reinterpret_cast<unsigned char*>(hits8);
            for (long long i = base; i < len; i++) {
                if (!((m >> (i - base)) & 1ull)) { if (hb[i]) hb[i] = 0; }
                else if (hb[i]) obs++;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        obs += __shfl_down_sync(0xffffffffu, obs, o);
        pos += __shfl_down_sync(0xffffffffu, pos, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (obs) atomicAdd(&counts[0], obs);
        if (pos) atomicAdd(&counts[1], pos);
    }
}

}  // namespace

extern "C" int cg_bin_screen(cg_ctx* ctx, int64_t chr_len, uint8_t* hits, uint64_t* possible_bits, int64_t n_filter,
                             const int32_t* filter_start, const int32_t* filter_stop, int64_t* n_observed, int64_t* n_possible) {
    if (!ctx) return CG_ERR_ARG;
    if (chr_len < 0 || chr_len > 0x7fff0000LL || n_filter < 0 || !n_observed || !n_possible || (n_filter > 0 && (!filter_start || !filter_stop)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_bin_screen: bad argument");
    *n_observed = *n_possible = 0;
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    if (chr_len == 0) return CG_OK;
    if (!hits || !possible_bits) return cg_fail(ctx, CG_ERR_ARG, "cg_bin_screen: null array");
    // `tags[chr][i] = false` throws past the end of the BitArray (and for a negative index)
    for (int64_t k = 0; k < n_filter; k++)
        if (filter_start[k] < filter_stop[k] && (filter_start[k] < 0 || filter_stop[k] > chr_len))
            return cg_fail(ctx, CG_ERR_ARG, "Index was out of range. Must be non-negative and less than the size of the collection.");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long nwords = (chr_len + 63) / 64;
    int rc = arena_reserve(ctx, arena_need(nwords * 64, 1) + arena_need(nwords, 8) + arena_need(n_filter + 1, 4) * 2 + (1 << 16));
    if (rc) return rc;
    unsigned long long* d_hits8 = arena_take<unsigned long long>(ctx, nwords * 8);
    unsigned long long* d_bits = arena_take<unsigned long long>(ctx, nwords);
    int32_t* d_fs = arena_take<int32_t>(ctx, n_filter + 1);
    int32_t* d_fe = arena_take<int32_t>(ctx, n_filter + 1);
    unsigned long long* d_counts = arena_take<unsigned long long>(ctx, 2);
    if (!d_hits8 || !d_bits || !d_fs || !d_fe || !d_counts) return cg_fail(ctx, CG_ERR_CUDA, "cg_bin_screen: device arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_hits8, hits, chr_len, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_bits, possible_bits, nwords * 8, cudaMemcpyHostToDevice, s));
    if (n_filter) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_fs, filter_start, n_filter * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d_fe, filter_stop, n_filter * 4, cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 16, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    if (n_filter) CG_LAUNCH(ctx, bin_filter_clear_kernel, div_up(n_filter * 32, 256), 256, 0, d_bits, d_fs, d_fe, (long long)n_filter);
    CG_LAUNCH(ctx, bin_screen_kernel, (int)std::min<long long>(ctx->num_sms * 8, div_up(nwords, 256)), 256, 0, d_hits8, d_bits, nwords,
              (long long)chr_len, d_counts);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    unsigned long long h_counts[2] = {0, 0};
    CG_CUDA(ctx, cudaMemcpyAsync(hits, d_hits8, chr_len, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(possible_bits, d_bits, nwords * 8, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(h_counts, d_counts, 16, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    *n_observed = (int64_t)h_counts[0];
    *n_possible = (int64_t)h_counts[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}

extern "C" int cg_bin_hits(cg_ctx* ctx, int64_t chr_len, const uint8_t* hits, const uint64_t* possible_bits, const char* bases,
                           int bin_size, int mode, const uint8_t* read_gc, const float* obs_vs_exp_gc, int64_t max_bins,
                           int64_t* n_bins, int32_t* start, int32_t* stop, int32_t* count, uint8_t* gc) {
    if (!ctx) return CG_ERR_ARG;
    if (chr_len < 0 || chr_len > 0x7fff0000LL || bin_size <= 0 || max_bins < 0 || !n_bins || (mode != 0 && mode != 1) ||
        (mode == 1 && (!read_gc || !obs_vs_exp_gc)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_bin_hits: bad argument");
    *n_bins = 0;
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    if (chr_len == 0) return CG_OK;
    if (!hits || !possible_bits || !bases || (max_bins > 0 && (!start || !stop || !count || !gc)))
        return cg_fail(ctx, CG_ERR_ARG, "cg_bin_hits: null array");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long nwords = (chr_len + 63) / 64;
    const int ntiles = (int)((nwords + BIN_TILE_WORDS - 1) / BIN_TILE_WORDS);
    const long long cap = std::min<long long>(max_bins, chr_len / bin_size + 1);
    size_t need = arena_need(chr_len, 1) * 3 + arena_need(nwords, 8) + arena_need(ntiles + 1, 4) + arena_need(ntiles + 2, 8) +
                  arena_need(cap + 1, 4) * 6 + arena_need(cap + 1, 1) + arena_need(256, 4) + arena_need(1, sizeof(BinCtl)) +
                  arena_need(chr_len / BIN_ACC_TILE + 4, 4) + (1 << 16);
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    unsigned char* d_hits = arena_take<unsigned char>(ctx, chr_len);
    char* d_bases = arena_take<char>(ctx, chr_len);
    unsigned char* d_rgc = arena_take<unsigned char>(ctx, chr_len);
    unsigned long long* d_bits = arena_take<unsigned long long>(ctx, nwords);
    unsigned* d_tcnt = arena_take<unsigned>(ctx, ntiles + 1);
    unsigned long long* d_toff = arena_take<unsigned long long>(ctx, ntiles + 2);
    int* d_end = arena_take<int>(ctx, cap + 1);
    int32_t* d_start = arena_take<int32_t>(ctx, cap + 1);
    int32_t* d_stop = arena_take<int32_t>(ctx, cap + 1);
    int32_t* d_count = arena_take<int32_t>(ctx, cap + 1);
    unsigned char* d_gc = arena_take<unsigned char>(ctx, cap + 1);
    unsigned* d_obs = arena_take<unsigned>(ctx, cap + 1);
    unsigned* d_gcc = arena_take<unsigned>(ctx, cap + 1);
    const long long n_acc_tiles = div_up(chr_len, BIN_ACC_TILE);
    int* d_tfirst = arena_take<int>(ctx, n_acc_tiles + 2);
    float* d_ratio = arena_take<float>(ctx, 256);
    BinCtl* d_ctl = arena_take<BinCtl>(ctx, 1);
    if (!d_hits || !d_bases || !d_rgc || !d_bits || !d_tcnt || !d_toff || !d_end || !d_start || !d_stop || !d_count || !d_gc || !d_obs || !d_gcc || !d_tfirst || !d_ratio || !d_ctl)
        return cg_fail(ctx, CG_ERR_CUDA, "cg_bin_hits: device arena exhausted");
    cudaStream_t s = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_hits, hits, chr_len, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_bases, bases, chr_len, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_bits, possible_bits, nwords * 8, cudaMemcpyHostToDevice, s));
    if (mode == 1) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_rgc, read_gc, chr_len, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d_ratio, obs_vs_exp_gc, 101 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    CG_LAUNCH(ctx, bin_init_kernel, 1, 1, 0, d_ctl);
    CG_LAUNCH(ctx, bin_first_base_kernel, ctx->num_sms, 256, 0, d_bases, (long long)chr_len, d_ctl);
    CG_LAUNCH(ctx, bin_tile_count_kernel, ntiles, 256, 0, d_bits, nwords, (long long)chr_len, d_ctl, d_tcnt);
    CG_LAUNCH(ctx, bin_tile_scan_kernel, 1, 1024, 0, d_tcnt, ntiles, d_toff, bin_size, d_ctl);
    if (cap > 0) {
        CG_CUDA(ctx, cudaMemsetAsync(d_obs, 0, (size_t)(cap + 1) * 4, s));
        CG_CUDA(ctx, cudaMemsetAsync(d_gcc, 0, (size_t)(cap + 1) * 4, s));
        CG_LAUNCH(ctx, bin_end_tile_kernel, ntiles, 256, 0, d_bits, nwords, (long long)chr_len, d_toff, bin_size, d_ctl, (int)cap, d_end);
        CG_LAUNCH(ctx, bin_tile_first_kernel, div_up(cap + 1, 256), 256, 0, d_ctl, d_end, (int)cap, n_acc_tiles, d_tfirst);
        CG_LAUNCH(ctx, bin_accum_kernel, (int)n_acc_tiles, BIN_ACC_THREADS, 0, d_hits, d_bits, d_bases, nwords, (long long)chr_len,
                  d_ctl, d_end, (int)cap, d_tfirst, d_obs, d_gcc);
        CG_LAUNCH(ctx, bin_finalize_kernel, div_up(cap, 256), 256, 0, d_ctl, d_end, (int)cap, d_obs, d_gcc, d_start, d_stop, d_count, d_gc);
        if (mode == 1)
            CG_LAUNCH(ctx, bin_sum_weighted_kernel, div_up(cap * 32, 256), 256, 0, d_hits, d_bits, d_rgc, d_ratio, d_ctl, d_end, (int)cap, d_count);
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    BinCtl* h = (BinCtl*)ctx->pinned;
    CG_CUDA(ctx, cudaMemcpyAsync(h, d_ctl, sizeof(BinCtl), cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    const long long nb = h->n_bins;
    if (nb > max_bins) return cg_fail(ctx, CG_ERR_CAPACITY, "cg_bin_hits: max_bins too small");
    if (nb > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(start, d_start, nb * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(stop, d_stop, nb * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(count, d_count, nb * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(gc, d_gc, nb, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    *n_bins = nb;
    return CG_OK;
}

extern "C" int cg_bin_fragments(cg_ctx* ctx, int64_t n_frag, const int32_t* frag_start, const int32_t* frag_stop, int64_t n_undo,
                                const int32_t* undo_index, int64_t n_bins, const int32_t* bin_start, const int32_t* bin_stop,
                                int32_t* best_bin, int32_t* count) {
    if (!ctx) return CG_ERR_ARG;
    if (n_frag < 0 || n_undo < 0 || n_bins < 0 || n_bins > 0x7fff0000LL || (n_bins > 0 && (!bin_start || !bin_stop || !count)) ||
        (n_frag > 0 && (!frag_start || !frag_stop || !best_bin)) || (n_undo > 0 && !undo_index))
        return cg_fail(ctx, CG_ERR_ARG, "cg_bin_fragments: bad argument");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    if (n_bins == 0) {
        for (int64_t i = 0; i < n_frag; i++) best_bin[i] = -1;
        return CG_OK;
    }
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = arena_reserve(ctx, arena_need(n_frag + 1, 4) * 3 + arena_need(n_undo + 1, 4) + arena_need(n_bins + 1, 4) * 3 + (1 << 16));
    if (rc) return rc;
    int32_t* d_fs = arena_take<int32_t>(ctx, n_frag + 1);
    int32_t* d_fe = arena_take<int32_t>(ctx, n_frag + 1);
    int32_t* d_best = arena_take<int32_t>(ctx, n_frag + 1);
    int32_t* d_undo = arena_take<int32_t>(ctx, n_undo + 1);
    int32_t* d_bs = arena_take<int32_t>(ctx, n_bins + 1);
    int32_t* d_be = arena_take<int32_t>(ctx, n_bins + 1);
    int32_t* d_cnt = arena_take<int32_t>(ctx, n_bins + 1);
    if (!d_fs || !d_fe || !d_best || !d_undo || !d_bs || !d_be || !d_cnt) return cg_fail(ctx, CG_ERR_CUDA, "cg_bin_fragments: device arena exhausted");
    cudaStream_t s = ctx->stream;
    if (n_frag > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_fs, frag_start, n_frag * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d_fe, frag_stop, n_frag * 4, cudaMemcpyHostToDevice, s));
    }
    if (n_undo > 0) CG_CUDA(ctx, cudaMemcpyAsync(d_undo, undo_index, n_undo * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_bs, bin_start, n_bins * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_be, bin_stop, n_bins * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, n_bins * 4, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    if (n_frag > 0) CG_LAUNCH(ctx, bin_fragments_kernel, div_up(n_frag, 256), 256, 0, d_fs, d_fe, (long long)n_frag, d_bs, d_be, (int)n_bins, d_best, d_cnt);
    if (n_undo > 0) CG_LAUNCH(ctx, bin_fragments_undo_kernel, div_up(n_undo, 256), 256, 0, d_undo, (long long)n_undo, d_best, d_cnt);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    if (n_frag > 0) CG_CUDA(ctx, cudaMemcpyAsync(best_bin, d_best, n_frag * 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(count, d_cnt, n_bins * 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return CG_OK;
}
