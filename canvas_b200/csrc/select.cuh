// Exact multi-segment, multi-rank order statistics by MSD radix select (8-bit digits).
//
// Every median / quartile / percentile of the reference is an exact element of a sorted list
// (Utilities.cs:340-474), and the result feeds a division, so approximations are not acceptable.
// Instead of sorting, each requested rank is resolved digit by digit: one histogram pass over the
// data per digit, followed by a tiny "resolve" kernel that walks the histograms and narrows every
// request to one bucket.  Requests of a segment that still share a prefix share one histogram row
// ("group"), so asking for rank k and k+1 (even-length median) or all quartile ranks costs the same
// passes as one rank.
//
// Layout: up to SEL_G requests and SEL_G groups per segment; row (seg, j) of `hist` has 256 bins.
#pragma once
#include <algorithm>

#include "common.cuh"

constexpr int SEL_G = 6;
constexpr int SEL_BINS = 256;

template <typename K>
struct SelState {
    int nseg;
    int* nreq;                   // [nseg]        requests per segment (0..SEL_G)
    unsigned long long* req_k;   // [nseg*SEL_G]  0-based rank, relative to the request's group
    int* req_grp;                // [nseg*SEL_G]  group slot of the request
    K* req_key;                  // [nseg*SEL_G]  result: key of the requested order statistic
    int* ngrp;                   // [nseg]
    K* gprefix;                  // [nseg*SEL_G]  decided high bits of each group
    unsigned* hist;              // [nseg*SEL_G*256]
    int* seg_done;               // [nseg + 1] blocks that finished the current pass (slot nseg: whole grid)
    unsigned* req_aux;           // [nseg*SEL_G*4] pair mode (see sel_resolve_segment): {count of the even key, count of the odd key,
                                 //                rank inside the pair, 0} of the pair of keys that holds the requested rank
};

template <typename K>
inline size_t sel_state_bytes(int nseg) {
    size_t s = 0;
    s += arena_need(nseg, sizeof(int)) * 2;
    s += arena_need((size_t)nseg * SEL_G, sizeof(unsigned long long));
    s += arena_need((size_t)nseg * SEL_G, sizeof(int));
    s += arena_need((size_t)nseg * SEL_G, sizeof(K)) * 2;
    s += arena_need((size_t)nseg * SEL_G * SEL_BINS, sizeof(unsigned));
    s += arena_need(nseg + 1, sizeof(int));
    s += arena_need((size_t)nseg * SEL_G * 4, sizeof(unsigned));
    return s;
}

template <typename K>
inline bool sel_state_alloc(cg_ctx* ctx, int nseg, SelState<K>& st) {
    st.nseg = nseg;
    st.nreq = arena_take<int>(ctx, nseg);
    st.ngrp = arena_take<int>(ctx, nseg);
    st.req_k = arena_take<unsigned long long>(ctx, (size_t)nseg * SEL_G);
    st.req_grp = arena_take<int>(ctx, (size_t)nseg * SEL_G);
    st.req_key = arena_take<K>(ctx, (size_t)nseg * SEL_G);
    st.gprefix = arena_take<K>(ctx, (size_t)nseg * SEL_G);
    st.hist = arena_take<unsigned>(ctx, (size_t)nseg * SEL_G * SEL_BINS);
    st.seg_done = arena_take<int>(ctx, nseg + 1);
    st.req_aux = arena_take<unsigned>(ctx, (size_t)nseg * SEL_G * 4);
    return st.seg_done && st.req_aux && st.nreq && st.ngrp && st.req_k && st.req_grp && st.req_key && st.gprefix && st.hist;
}

template <typename K>
__device__ void sel_resolve_segment(SelState<K>& st, int s, int shift, int last, int pair = 0);

// After the caller filled nreq / req_k: one group per segment with an empty prefix.
template <typename K>
__global__ void sel_begin_kernel(SelState<K> st) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    int nr = st.nreq[s];
    st.ngrp[s] = nr > 0 ? 1 : 0;
    st.seg_done[s] = 0;
    if (s == 0) st.seg_done[st.nseg] = 0;
    for (int j = 0; j < SEL_G; j++) {
        st.gprefix[s * SEL_G + j] = (K)0;
        st.req_grp[s * SEL_G + j] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Histogram pass, scattered segments: every element may belong to two segments (its own bucket and
// a "global" one).  Row 0 of every segment is privatised in shared memory when it fits (PRIV) and
// updates to it are warp-aggregated (degenerate digits — e.g. all bin sizes equal — would otherwise
// serialise 32-way on one address); rows >= 1 only exist once ranks of a segment have diverged,
// when few elements still match, and go straight to global atomics.
//
// View: __device__ long long size() const;
//       __device__ bool get(long long i, K& key, int& segA, int& segB) const;  (seg < 0: none)
// ---------------------------------------------------------------------------------------------
template <typename K, class View, int PRIV>
__global__ void __launch_bounds__(1024) sel_hist_scatter_kernel(View v, SelState<K> st, int shift, int first, int last) {
    extern __shared__ unsigned char sel_smem[];
    const int nseg = st.nseg;
    K* s_prefix = (K*)sel_smem;
    int* s_ngrp = (int*)(s_prefix + (size_t)nseg * SEL_G);
    unsigned* s_hist = (unsigned*)(s_ngrp + nseg);
    // PRIV == 2: rows 1 and 2 (the other quartile / median-pair groups once ranks have diverged) as two 16-bit
    // counters per word; a block sees fewer than 65536 elements (checked by the host driver)
    unsigned* s_hist12 = s_hist + (size_t)nseg * SEL_BINS;
    for (int t = threadIdx.x; t < nseg * SEL_G; t += blockDim.x) s_prefix[t] = st.gprefix[t];
    for (int t = threadIdx.x; t < nseg; t += blockDim.x) s_ngrp[t] = st.ngrp[t];
    if (PRIV)
        for (int t = threadIdx.x; t < nseg * SEL_BINS * (PRIV == 2 ? 2 : 1); t += blockDim.x) s_hist[t] = 0u;
    __syncthreads();

    const long long n = v.size();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int hi_shift = shift + 8;
    // Row-0 updates are aggregated per thread as run lengths: the leading digits of a data set are nearly constant
    // (one flush per thread instead of one atomic per element), the trailing digits only concern the few elements
    // that still match a prefix.  One run per segment slot (own bucket / global list).
    int run_idx[2] = {-1, -1};
    unsigned run_len[2] = {0u, 0u};
    auto flush = [&](int a) {
        if (run_len[a]) {
            if (PRIV) atomicAdd(&s_hist[run_idx[a]], run_len[a]);
            else atomicAdd(&st.hist[(size_t)(run_idx[a] / SEL_BINS * SEL_G) * SEL_BINS + (run_idx[a] % SEL_BINS)], run_len[a]);
        }
    };
    auto update = [&](K key, const int* seg) {
        const int d = (int)((key >> shift) & (K)255);
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const int s = seg[a];
            if (s < 0) continue;
            const int ng = s_ngrp[s];
            if (ng <= 0) continue;
            if (first || (((key ^ s_prefix[s * SEL_G]) >> hi_shift) == 0)) {
                const int idx = s * SEL_BINS + d;
                if (idx == run_idx[a]) run_len[a]++;
                else { flush(a); run_idx[a] = idx; run_len[a] = 1u; }
            }
            for (int j = 1; j < ng; j++)
                if (((key ^ s_prefix[s * SEL_G + j]) >> hi_shift) == 0) {
                    if (PRIV == 2 && j <= 2) atomicAdd(&s_hist12[s * SEL_BINS + d], j == 1 ? 1u : 0x10000u);
                    else atomicAdd(&st.hist[(size_t)(s * SEL_G + j) * SEL_BINS + d], 1u);
                }
        }
    };
    // four elements per step: their (dependent) loads overlap
    constexpr int SEL_U = 4;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (SEL_U - 1) * stride < n; i += SEL_U * stride) {
        K key[SEL_U];
        int seg[SEL_U][2];
        bool ok[SEL_U];
#pragma unroll
        for (int u = 0; u < SEL_U; u++) {
            key[u] = 0; seg[u][0] = -1; seg[u][1] = -1;
            ok[u] = v.get(i + u * stride, key[u], seg[u][0], seg[u][1]);
        }
#pragma unroll
        for (int u = 0; u < SEL_U; u++)
            if (ok[u]) update(key[u], seg[u]);
    }
    for (; i < n; i += stride) {
        K key = 0;
        int seg[2] = {-1, -1};
        if (v.get(i, key, seg[0], seg[1])) update(key, seg);
    }
    flush(0);
    flush(1);
    if (PRIV) {
        __syncthreads();
        for (int t = threadIdx.x; t < nseg * SEL_BINS; t += blockDim.x) {
            unsigned c = s_hist[t];
            if (c) atomicAdd(&st.hist[(size_t)((t / SEL_BINS) * SEL_G) * SEL_BINS + (t % SEL_BINS)], c);
            if (PRIV == 2) {
                const unsigned c12 = s_hist12[t];
                if (c12 & 0xffffu) atomicAdd(&st.hist[(size_t)((t / SEL_BINS) * SEL_G + 1) * SEL_BINS + (t % SEL_BINS)], c12 & 0xffffu);
                if (c12 >> 16) atomicAdd(&st.hist[(size_t)((t / SEL_BINS) * SEL_G + 2) * SEL_BINS + (t % SEL_BINS)], c12 >> 16);
            }
        }
    }
    // the last block to finish the pass resolves every segment (one warp each); every thread fences its
    // own histogram updates before the block reports completion
    __shared__ int s_last_block;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last_block = atomicAdd(&st.seg_done[nseg], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last_block) {
        __threadfence();
        const int nwarps = blockDim.x >> 5;
        for (int s = threadIdx.x >> 5; s < nseg; s += nwarps) sel_resolve_segment<K>(st, s, shift, last);
        if (threadIdx.x == 0) st.seg_done[nseg] = 0;
#ifdef SEL_DEBUG
        __syncthreads();
        if (threadIdx.x == 0 && sizeof(K) == 8) printf("[sel] shift=%d grid=%d last=%d nseg=%d nreq0=%d ngrp0=%d k0=%llu key0=%llx\n", shift, (int)gridDim.x, last, nseg, st.nreq[0], st.ngrp[0], st.req_k[0], (unsigned long long)st.req_key[0]);
#endif
    }
}

template <typename K>
inline size_t sel_scatter_smem(int nseg, int priv) {
    size_t s = (size_t)nseg * SEL_G * sizeof(K) + (size_t)nseg * sizeof(int);
    s += (size_t)priv * nseg * SEL_BINS * sizeof(unsigned);
    return s;
}

// ---------------------------------------------------------------------------------------------
// Histogram pass, contiguous segments: a host-built work list cuts every segment into chunks;
// block b handles work[b] = {seg, lo, hi} with all group rows of that segment private in shared
// memory.  Segments may overlap in the underlying arrays (windows and whole chromosomes).
// View: __device__ bool get(long long i, int seg, K& key) const;   (false: skip element)
//       __device__ bool plain(int seg, const double*& base, double& centre, bool& has_centre) const;
//         (true: the segment's elements are base[i], or |base[i] - centre|, every one of them valid)
// ---------------------------------------------------------------------------------------------
constexpr int SEL_WSEG = 3;  // segments one work item can feed (e.g. chromosome + the two windows containing the chunk)

struct SelWork {
    int seg[SEL_WSEG];  // seg[0] >= 0; unused slots are -1.  All of them cover [lo, hi) completely.
    int pad;
    long long lo, hi;
};

template <typename K, class View>
__global__ void __launch_bounds__(256, 4) sel_hist_contig_kernel(View v, const SelWork* __restrict__ work, const int* __restrict__ seg_nwork,
                                       SelState<K> st, int shift, int first, int last, int pair) {
    __shared__ unsigned s_hist[SEL_WSEG][SEL_G * SEL_BINS];
    __shared__ K s_prefix[SEL_WSEG][SEL_G];
    __shared__ int s_last_block[SEL_WSEG];
    const SelWork w = work[blockIdx.x];
    int ng[SEL_WSEG];
    bool any = false;
#pragma unroll
    for (int a = 0; a < SEL_WSEG; a++) {
        ng[a] = w.seg[a] >= 0 ? st.ngrp[w.seg[a]] : 0;
        any = any || ng[a] > 0;
    }
    if (!any) return;
#pragma unroll
    for (int a = 0; a < SEL_WSEG; a++) {
        for (int t = threadIdx.x; t < ng[a] * SEL_BINS; t += blockDim.x) s_hist[a][t] = 0u;
        if (ng[a] > 0 && threadIdx.x < SEL_G) s_prefix[a][threadIdx.x] = st.gprefix[w.seg[a] * SEL_G + threadIdx.x];
    }
    __syncthreads();
    const int hi_shift = shift + 8;
    const long long span = w.hi - w.lo;
    // per-thread run-length aggregation (see the scattered kernel); an element matches at most one group of a segment
    int run_idx[SEL_WSEG];
    unsigned run_len[SEL_WSEG];
#pragma unroll
    for (int a = 0; a < SEL_WSEG; a++) { run_idx[a] = -1; run_len[a] = 0u; }
    auto update = [&](int a, K key) {  // a is a compile-time constant at every call site
        const int d = (int)((key >> shift) & (K)255);
        int idx = -1;
        if (first) idx = d;
        else {
            for (int j = 0; j < ng[a]; j++)
                if (((key ^ s_prefix[a][j]) >> hi_shift) == 0) { idx = j * SEL_BINS + d; break; }
        }
        if (idx < 0) return;
        if (idx == run_idx[a]) run_len[a]++;
        else {
            if (run_len[a]) atomicAdd(&s_hist[a][run_idx[a]], run_len[a]);
            run_idx[a] = idx;
            run_len[a] = 1u;
        }
    };
    constexpr int SEL_U = 4;  // loads of four elements in flight
    const long long bd = blockDim.x;
    long long o = threadIdx.x;
    if constexpr (sizeof(K) == 8) {
        // segments that are a plain double array (optionally |x - centre|): one load per element feeds every segment
        // of the work item, no per-element dispatch on the segment kind
        const double* fp = nullptr;
        double centre[SEL_WSEG] = {0.0, 0.0, 0.0};
        bool has_centre = false;
        bool plain = true;
#pragma unroll
        for (int a = 0; a < SEL_WSEG; a++)
            if (ng[a] > 0) {
                const double* base = nullptr;
                bool hc = false;
                if (!v.plain(w.seg[a], base, centre[a], hc)) plain = false;
                else if (fp && (fp != base || hc != has_centre)) plain = false;
                else { fp = base; has_centre = hc; }
            }
        if (plain && fp) {
            fp += w.lo;
            auto feed = [&](double x) {
                if (!has_centre) {
                    const K key = (K)f64_key(x);
#pragma unroll
                    for (int a = 0; a < SEL_WSEG; a++)
                        if (ng[a] > 0) update(a, key);
                } else {
#pragma unroll
                    for (int a = 0; a < SEL_WSEG; a++)
                        if (ng[a] > 0) update(a, (K)f64_key(fabs(x - centre[a])));
                }
            };
            for (; o + (SEL_U - 1) * bd < span; o += SEL_U * bd) {
                double x[SEL_U];
#pragma unroll
                for (int u = 0; u < SEL_U; u++) x[u] = fp[o + u * bd];
#pragma unroll
                for (int u = 0; u < SEL_U; u++) feed(x[u]);
            }
            for (; o < span; o += bd) feed(fp[o]);
            o = span;  // the generic loop below has nothing left
        }
    }
    if constexpr (sizeof(K) == 4) {
        // integer keys (hundredths of the coverage): one 4-byte load per element feeds every segment of the work item.
        // View: __device__ bool plain32(int seg, const uint32_t*& base, int& twice_centre, bool& has_centre) const;
        //   key = base[i], or the distance class (|2 base[i] - twice_centre| << 1) | (2 base[i] > twice_centre)
        const uint32_t* hp = nullptr;
        int c2[SEL_WSEG] = {0, 0, 0};
        bool has_centre = false;
        bool plain = true;
#pragma unroll
        for (int a = 0; a < SEL_WSEG; a++)
            if (ng[a] > 0) {
                const uint32_t* base = nullptr;
                bool hc = false;
                if (!v.plain32(w.seg[a], base, c2[a], hc)) plain = false;
                else if (hp && (hp != base || hc != has_centre)) plain = false;
                else { hp = base; has_centre = hc; }
            }
        if (plain && hp) {
            hp += w.lo;
            auto feed = [&](uint32_t h) {
                if (!has_centre) {
#pragma unroll
                    for (int a = 0; a < SEL_WSEG; a++)
                        if (ng[a] > 0) update(a, (K)h);
                } else {
#pragma unroll
                    for (int a = 0; a < SEL_WSEG; a++)
                        if (ng[a] > 0) {
                            const int t = 2 * (int)h - c2[a];
                            update(a, (K)(((unsigned)abs(t) << 1) | (t > 0 ? 1u : 0u)));
                        }
                }
            };
            constexpr int SEL_U32 = 8;
            for (; o + (SEL_U32 - 1) * bd < span; o += SEL_U32 * bd) {
                uint32_t x[SEL_U32];
#pragma unroll
                for (int u = 0; u < SEL_U32; u++) x[u] = hp[o + u * bd];
#pragma unroll
                for (int u = 0; u < SEL_U32; u++) feed(x[u]);
            }
            for (; o < span; o += bd) feed(hp[o]);
            o = span;
        }
    }
    for (; o < span; o += bd) {
#pragma unroll
        for (int a = 0; a < SEL_WSEG; a++)
            if (ng[a] > 0) {
                K key = 0;
                if (v.get(w.lo + o, w.seg[a], key)) update(a, key);
            }
    }
#pragma unroll
    for (int a = 0; a < SEL_WSEG; a++)
        if (run_len[a]) atomicAdd(&s_hist[a][run_idx[a]], run_len[a]);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < SEL_WSEG; a++)
        for (int t = threadIdx.x; t < ng[a] * SEL_BINS; t += blockDim.x) {
            unsigned c = s_hist[a][t];
            if (c) atomicAdd(&st.hist[(size_t)(w.seg[a] * SEL_G) * SEL_BINS + t], c);
        }
    // the last block of a segment resolves it (every thread fences its own updates first); warp a takes slot a
    __threadfence();
    __syncthreads();
    if (threadIdx.x < SEL_WSEG) {
        const int a = threadIdx.x;
        __threadfence();
        s_last_block[a] = ng[a] > 0 && atomicAdd(&st.seg_done[w.seg[a]], 1) == seg_nwork[w.seg[a]] - 1;
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5;
    if (wid < SEL_WSEG && s_last_block[wid]) {
        __threadfence();
        sel_resolve_segment<K>(st, w.seg[wid], shift, last, pair);
        if ((threadIdx.x & 31) == 0) st.seg_done[w.seg[wid]] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Resolve: one warp per segment walks the rows of its groups, moves each request into the bucket
// that holds its rank, regroups requests by their new prefix and clears the rows for the next pass.
// ---------------------------------------------------------------------------------------------
// Resolve one segment with one warp: walk the rows of its groups, move each request into the bucket
// that holds its rank, regroup requests by their new prefix and clear the rows for the next pass.
// pair != 0 (last pass only): keys 2j and 2j+1 form one class whose two members the caller orders itself (the two sides of a
// distance class in the MAD wave); besides the key, the populations of both members and the rank inside the pair are recorded.
template <typename K>
__device__ void sel_resolve_segment(SelState<K>& st, int s, int shift, int last, int pair) {
    const int lane = threadIdx.x & 31;
    const int nr = st.nreq[s];
    const int ng = st.ngrp[s];
    if (nr == 0 || ng == 0) return;
    unsigned long long rk[SEL_G];
    int rgrp[SEL_G], rnew[SEL_G];
    K newpref[SEL_G];
    int newng = 0;
#pragma unroll
    for (int r = 0; r < SEL_G; r++) {
        rk[r] = r < nr ? st.req_k[s * SEL_G + r] : 0ull;
        rgrp[r] = r < nr ? st.req_grp[s * SEL_G + r] : -1;
        rnew[r] = 0;
        newpref[r] = (K)0;
    }
    for (int j = 0; j < ng; j++) {
        unsigned* row = st.hist + (size_t)(s * SEL_G + j) * SEL_BINS;
        unsigned c[8];
        unsigned long long sum = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            c[t] = __ldcg(row + lane * 8 + t);
            sum += c[t];
            row[lane * 8 + t] = 0u;
        }
        unsigned long long incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const unsigned long long excl = incl - sum;
        const K pref = st.gprefix[s * SEL_G + j];
#pragma unroll
        for (int r = 0; r < SEL_G; r++) {
            if (rgrp[r] != j) continue;  // uniform across the warp
            const unsigned long long k = rk[r];
            const bool mine = sum > 0 && k >= excl && k < excl + sum;
            unsigned b = __ballot_sync(0xffffffffu, mine);
            if (b == 0u) {  // rank beyond the group (never for well-formed requests): clamp to the top bucket
                const unsigned nz = __ballot_sync(0xffffffffu, sum > 0);
                b = nz ? (1u << (31 - __clz(nz))) : 1u;
            }
            const int owner = __ffs(b) - 1;
            int d = lane * 8 + 7;
            unsigned long long cum = excl;
            if (lane == owner) {
                unsigned long long run = excl;
                d = lane * 8 + 7;
                cum = excl + sum - c[7];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (k < run + c[t]) { d = lane * 8 + t; cum = run; break; }
                    run += c[t];
                }
            }
            if (pair && last) {  // the pair (d & ~1, d | 1) lies inside the owner lane's eight bins
                unsigned c0 = 0, c1 = 0, rin = 0;
                if (lane == owner) {
                    const int t0 = (d & 7) & ~1;
#pragma unroll
                    for (int t = 0; t < 8; t += 2)
                        if (t == t0) { c0 = c[t]; c1 = c[t + 1]; }
                    const unsigned long long cum_pair = (d & 1) ? cum - c0 : cum;
                    rin = (unsigned)(k - cum_pair);
                    unsigned* aux = st.req_aux + ((size_t)s * SEL_G + r) * 4;
                    aux[0] = c0; aux[1] = c1; aux[2] = rin; aux[3] = 0u;
                }
            }
            d = __shfl_sync(0xffffffffu, d, owner);
            cum = __shfl_sync(0xffffffffu, cum, owner);
            rk[r] = k >= cum ? k - cum : 0ull;
            const K np = pref | ((K)d << shift);
            int idx = -1;
            for (int q = 0; q < newng; q++)
                if (newpref[q] == np) idx = q;
            if (idx < 0) { idx = newng; newpref[newng++] = np; }
            rnew[r] = idx;
        }
    }
    if (lane == 0) {
        for (int r = 0; r < nr; r++) {
            st.req_k[s * SEL_G + r] = rk[r];
            st.req_grp[s * SEL_G + r] = rnew[r];
            if (last) st.req_key[s * SEL_G + r] = newpref[rnew[r]];
        }
        for (int q = 0; q < newng; q++) st.gprefix[s * SEL_G + q] = newpref[q];
        st.ngrp[s] = newng;
    }
}

// Host drivers.  The caller has already written nreq/req_k (device side) before calling.
template <typename K, class View>
inline void sel_run_scatter(cg_ctx* ctx, const View& v, SelState<K>& st, long long n_upper) {
    const int bits = (int)sizeof(K) * 8;
    CG_LAUNCH(ctx, sel_begin_kernel<K>, div_up(st.nseg + 1, 128), 128, 0, st);
    // histogram rows in shared memory: rows 0..2 when they fit next to each other (a block must then see fewer
    // than 65536 elements: rows 1 and 2 are 16-bit counters), else row 0, else none
    const long long per_block = (n_upper + ctx->num_sms - 1) / std::max(1, ctx->num_sms);
    int priv = 0;
    if (sel_scatter_smem<K>(st.nseg, 2) <= 220 * 1024 && per_block < 60000 && n_upper >= (long long)ctx->num_sms * 1024) priv = 2;
    else if (sel_scatter_smem<K>(st.nseg, 1) <= 200 * 1024) priv = 1;
    const size_t smem = sel_scatter_smem<K>(st.nseg, priv);
    int grid = (int)std::min<long long>(std::max<long long>(1, (n_upper + 1023) / 1024), (long long)ctx->num_sms * (priv ? 1 : 4));
    if (priv == 2) cudaFuncSetAttribute(sel_hist_scatter_kernel<K, View, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (priv == 1) cudaFuncSetAttribute(sel_hist_scatter_kernel<K, View, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int shift = bits - 8; shift >= 0; shift -= 8) {
        int first = shift == bits - 8;
        if (priv == 2)
            CG_LAUNCH(ctx, (sel_hist_scatter_kernel<K, View, 2>), grid, 1024, smem, v, st, shift, first, shift == 0);
        else if (priv == 1)
            CG_LAUNCH(ctx, (sel_hist_scatter_kernel<K, View, 1>), grid, 1024, smem, v, st, shift, first, shift == 0);
        else
            CG_LAUNCH(ctx, (sel_hist_scatter_kernel<K, View, 0>), grid, 256, smem, v, st, shift, first, shift == 0);
    }
}

// key_bits: keys are known to be below 2^key_bits (a multiple of 8): the passes over the constant leading digits are skipped
template <typename K, class View>
inline void sel_run_contig(cg_ctx* ctx, const View& v, const SelWork* work_dev, const int* seg_nwork_dev, int nwork,
                           SelState<K>& st, int key_bits = (int)sizeof(K) * 8, int pair = 0) {
    CG_LAUNCH(ctx, sel_begin_kernel<K>, div_up(st.nseg + 1, 128), 128, 0, st);
    if (nwork <= 0) return;
    for (int shift = key_bits - 8; shift >= 0; shift -= 8) {
        int first = shift == key_bits - 8;
        CG_LAUNCH(ctx, (sel_hist_contig_kernel<K, View>), nwork, 256, 0, v, work_dev, seg_nwork_dev, st, shift, first, shift == 0, pair);
    }
}
