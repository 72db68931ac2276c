// Per-chromosome tail of HaarWavelets (WaveletSegmentation.cs:406-425), one CTA per chromosome:
//   HardThresh :72-115 on the candidate nodes, GetSegments :174-185 evaluated only on the pieces
//   delimited by surviving nodes, GetBreakpointsAfterHealingBadSplits :194-232, RefineSegments
//   :237-258.  Range medians are exact (block-wide radix select over the L2-resident coverage).
#pragma once
#include "wavelet.cuh"
#include "wavelet_rqindex.cuh"

constexpr int FIN_THREADS = 1024;
constexpr int FIN_SORT_SMEM = 4096;  // survivors sorted in shared memory up to this many
constexpr int BMS_R = 24;            // requests of one block multi-select
constexpr int BMS_X = 12;            // "extra" bins per request beyond the shared core range

struct FinParams {
    const double* x;           // coverage
    const double* pz;          // prefix sums (for the smooth term)
    const long long* off;
    const unsigned char* selected;
    const UhCand* cand;        // candidate lists, one slice per chromosome (cp[c].cand_base, cc[c].cand_count_)
    const UhChromCtl* cc;
    const UhChromPlan* cp;
    const WvCtl* ctl;
    const unsigned* lvlcnt;
    const int* depth;
    const double* sigma;       // [n_chrom]
    const double* chrom_median;  // [n_chrom] Median(ratio)
    const double* log3_scale_tab;  // ceil(log(3^k)/log 3) as the host libm evaluates it
    RqIndex rq;
    int is_germline, min_size, n_chrom, pad;
    // scratch, all indexed per chromosome at off[c]
    int* lvl_idx;              // [N]   introsort permutation of the levels
    int* sv;                   // [N]   candidate indices of the survivors
    unsigned long long* svkey; // [N]   sort keys (level << 32 | start)
    unsigned* bitmap;          // [N/32 + n_chrom] boundary bitmap, chromosome c at (off[c] >> 5) + c
    int* piece;                // [N]   piece start positions
    double* rec;               // [N]   reconstructed value per piece
    int* prelim;               // [N]   preliminary breakpoints
    int* lvl_first;            // [N]   first survivor of each distinct level
    // outputs
    unsigned long long* phase_ns;  // [n_chrom][8] debug timeline (nullable)
    int* n_bp;                 // [n_chrom]
    int* bp;                   // [N] at off[c]
};

// ------------------------------------------------------------------------------------------------
// .NET Core 2.0 Array.Sort<int>(indices, (a, b) => counts[b].CompareTo(counts[a])) — the introspective
// sort of ArraySortHelper<T> restated (unstable: the order of equal counts is part of the result).
// ------------------------------------------------------------------------------------------------
struct LevelSorter {
    int* k;            // level ids being permuted
    unsigned* kc;      // count of the level currently at each position (moves with k)
    __device__ static int cmpc(unsigned ca, unsigned cb) { return cb < ca ? -1 : (cb > ca ? 1 : 0); }  // counts[b].CompareTo(counts[a])
    __device__ void swap(int i, int j) {
        const int t = k[i]; k[i] = k[j]; k[j] = t;
        const unsigned c = kc[i]; kc[i] = kc[j]; kc[j] = c;
    }
    __device__ void swap_if_greater(int a, int b) {
        if (a != b && cmpc(kc[a], kc[b]) > 0) swap(a, b);
    }
    __device__ void insertion(int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            int j = i;
            const int t = k[i + 1];
            const unsigned tc = kc[i + 1];
            while (j >= lo && cmpc(tc, kc[j]) < 0) { k[j + 1] = k[j]; kc[j + 1] = kc[j]; j--; }
            k[j + 1] = t;
            kc[j + 1] = tc;
        }
    }
    __device__ void down_heap(int i, int n, int lo) {
        const int d = k[lo + i - 1];
        const unsigned dc = kc[lo + i - 1];
        while (i <= n / 2) {
            int child = 2 * i;
            if (child < n && cmpc(kc[lo + child - 1], kc[lo + child]) < 0) child++;
            if (!(cmpc(dc, kc[lo + child - 1]) < 0)) break;
            k[lo + i - 1] = k[lo + child - 1];
            kc[lo + i - 1] = kc[lo + child - 1];
            i = child;
        }
        k[lo + i - 1] = d;
        kc[lo + i - 1] = dc;
    }
    __device__ void heapsort(int lo, int hi) {
        const int n = hi - lo + 1;
        for (int i = n / 2; i >= 1; i--) down_heap(i, n, lo);
        for (int i = n; i > 1; i--) { swap(lo, lo + i - 1); down_heap(1, i - 1, lo); }
    }
    __device__ int partition(int lo, int hi) {
        const int mid = lo + (hi - lo) / 2;
        swap_if_greater(lo, mid);
        swap_if_greater(lo, hi);
        swap_if_greater(mid, hi);
        const unsigned pivot = kc[mid];
        swap(mid, hi - 1);
        int left = lo, right = hi - 1;
        while (left < right) {
            while (cmpc(kc[++left], pivot) < 0) {}
            while (cmpc(pivot, kc[--right]) < 0) {}
            if (left >= right) break;
            swap(left, right);
        }
        swap(left, hi - 1);
        return left;
    }
};

// The replay above is inherently sequential inside one sub-array, but after a partition the two parts
// never touch each other again, so they can be replayed by different threads and still give the
// identical permutation.  Round r runs every pending sub-array on its own thread (one partition, or
// the insertion / heap sort that finishes it) and queues the two parts for round r+1.  Needed because a
// run of exact zeros makes the tree (and with it the number of levels to sort) thousands of levels deep.
//
// Large sub-arrays are partitioned by the whole block instead (level_partition_block): the sequential
// Hoare loop swaps the t-th element from the left that is not below the pivot with the t-th element from
// the right that is not above it for as long as the former lies left of the latter, so ranking both kinds
// of "stops" with a block scan gives every swap pair at once.  In detail, with l_0 < l_1 < ... the stops
// of the left scan over (lo, hi-1] and r_0 > r_1 > ... those of the right scan over [lo, hi-1) — the pivot
// parked at hi-1 and the median-of-three minimum at lo are the sentinels — iteration t of the loop finds
// left = min(l_t, r_{t-1}) and right = max(r_t, l_{t-1}): untouched positions keep their original
// elements and an already swapped position holds a stop.  Hence pairs (l_t, r_t) are swapped while
// l_t < r_t, a prefix property, and the loop exits with left = min(l_T, r_{T-1}) for the first T with
// l_T >= r_T (r_{-1} = hi-1).
constexpr int FIN_SORT_PAR_MIN = 160;  // sub-arrays longer than this are partitioned by the block
constexpr int FIN_SORT_BIG_CAP = 64;   // pending large sub-arrays (disjoint, each > FIN_SORT_PAR_MIN of <= 4096)

__device__ int level_partition_block(LevelSorter ls, int lo, int hi, int* Lpos, int* Rpos) {
    __shared__ unsigned s_pivot;
    __shared__ int s_wl[32], s_wr[32];
    __shared__ int s_totl, s_totr, s_left;
    const int tid = threadIdx.x, P = blockDim.x;
    if (tid == 0) {
        const int mid = lo + (hi - lo) / 2;
        ls.swap_if_greater(lo, mid);
        ls.swap_if_greater(lo, hi);
        ls.swap_if_greater(mid, hi);
        s_pivot = ls.kc[mid];
        ls.swap(mid, hi - 1);
    }
    __syncthreads();
    const unsigned pivot = s_pivot;
    const int size = hi - lo + 1;
    const int E = (size + P - 1) / P;
    const int p0 = lo + tid * E, p1 = min(p0 + E, hi + 1);
    int cl = 0, cr = 0;
    for (int q = p0; q < p1; q++) {
        const unsigned v = ls.kc[q];
        if (q > lo && q <= hi - 1 && !(LevelSorter::cmpc(v, pivot) < 0)) cl++;
        if (q >= lo && q < hi - 1 && !(LevelSorter::cmpc(pivot, v) < 0)) cr++;
    }
    // block-wide exclusive scan of (cl, cr)
    int il = cl, ir = cr;
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, il, o), b2 = __shfl_up_sync(0xffffffffu, ir, o);
        if ((tid & 31) >= o) { il += a; ir += b2; }
    }
    if ((tid & 31) == 31) { s_wl[tid >> 5] = il; s_wr[tid >> 5] = ir; }
    __syncthreads();
    if (tid < 32) {
        const int nw = (P + 31) >> 5;
        int a = tid < nw ? s_wl[tid] : 0, b2 = tid < nw ? s_wr[tid] : 0;
        int ia = a, ib = b2;
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
            if (tid >= o) { ia += x; ib += y; }
        }
        s_wl[tid] = ia - a;
        s_wr[tid] = ib - b2;
        if (tid == 31) { s_totl = ia; s_totr = ib; }
    }
    __syncthreads();
    const int totl = s_totl, totr = s_totr;
    int rl = s_wl[tid >> 5] + il - cl;  // rank of this thread's first left stop
    int rr = s_wr[tid >> 5] + ir - cr;  // ascending rank of its first right stop
    for (int q = p0; q < p1; q++) {
        const unsigned v = ls.kc[q];
        if (q > lo && q <= hi - 1 && !(LevelSorter::cmpc(v, pivot) < 0)) Lpos[rl++] = q;
        if (q >= lo && q < hi - 1 && !(LevelSorter::cmpc(pivot, v) < 0)) { Rpos[totr - 1 - rr] = q; rr++; }
    }
    __syncthreads();
    // number of swapped pairs: l_t < r_t holds for a prefix of t
    int a = 0, b2 = min(totl, totr);
    while (a < b2) {
        const int m = (a + b2) >> 1;
        if (Lpos[m] < Rpos[m]) a = m + 1; else b2 = m;
    }
    const int nswap = a;
    for (int t = tid; t < nswap; t += P) ls.swap(Lpos[t], Rpos[t]);
    if (tid == 0) s_left = min(Lpos[nswap], nswap > 0 ? Rpos[nswap - 1] : hi - 1);
    __syncthreads();
    const int left = s_left;
    if (tid == 0) ls.swap(left, hi - 1);
    __syncthreads();
    return left;
}

// Lpos / Rpos: scratch for the block partition (n ints each) or nullptr (every sub-array on one thread).
__device__ void level_sort_parallel(LevelSorter ls, int n, int* task_a, int* task_b, int task_cap, int* s_counts2,
                                    int* Lpos = nullptr, int* Rpos = nullptr) {
    if (n < 2) return;
    __shared__ int s_big[2][3 * FIN_SORT_BIG_CAP];
    __shared__ int s_nbig[2];
    const bool coop = Lpos != nullptr && Rpos != nullptr;
    int depth0 = 0;
    for (int t = n; t >= 1; t /= 2) depth0++;
    depth0 *= 2;
    int* cur = task_a;
    int* nxt = task_b;
    int bc = 0;  // index of the current big list
    if (threadIdx.x == 0) {
        s_nbig[0] = s_nbig[1] = 0;
        s_counts2[0] = 0; s_counts2[1] = 0;
        if (coop && n > FIN_SORT_PAR_MIN) { s_big[0][0] = 0; s_big[0][1] = n - 1; s_big[0][2] = depth0; s_nbig[0] = 1; }
        else { cur[0] = 0; cur[1] = n - 1; cur[2] = depth0; s_counts2[0] = 1; }
    }
    __syncthreads();
    // queue a sub-array for the next round (any thread)
    auto push = [&](int lo, int hi, int depth) {
        if (coop && hi - lo + 1 > FIN_SORT_PAR_MIN) {
            const int i = atomicAdd(&s_nbig[bc ^ 1], 1);
            if (i < FIN_SORT_BIG_CAP) { s_big[bc ^ 1][3 * i] = lo; s_big[bc ^ 1][3 * i + 1] = hi; s_big[bc ^ 1][3 * i + 2] = depth; }
        } else {
            const int i = atomicAdd(&s_counts2[1], 1);
            if (i < task_cap) { nxt[3 * i] = lo; nxt[3 * i + 1] = hi; nxt[3 * i + 2] = depth; }
        }
    };
    for (;;) {
        const int ncur = s_counts2[0];
        const int nbig = min(s_nbig[bc], FIN_SORT_BIG_CAP);
        if (ncur <= 0 && nbig <= 0) break;
        // large sub-arrays: one after the other, all threads
        for (int b = 0; b < nbig; b++) {
            const int lo = s_big[bc][3 * b], hi = s_big[bc][3 * b + 1];
            int depth = s_big[bc][3 * b + 2];
            if (depth == 0) {
                if (threadIdx.x == 0) ls.heapsort(lo, hi);
                continue;
            }
            depth--;
            const int pv = level_partition_block(ls, lo, hi, Lpos, Rpos);
            if (threadIdx.x == 0) {
                if (hi > pv + 1) push(pv + 1, hi, depth);
                if (pv - 1 > lo) push(lo, pv - 1, depth);
            }
        }
        for (int t = threadIdx.x; t < ncur; t += blockDim.x) {
            const int lo = cur[3 * t], hi = cur[3 * t + 1];
            int depth = cur[3 * t + 2];
            if (hi <= lo) continue;
            const int size = hi - lo + 1;
            if (size <= 16) {
                if (size == 2) ls.swap_if_greater(lo, hi);
                else if (size == 3) { ls.swap_if_greater(lo, hi - 1); ls.swap_if_greater(lo, hi); ls.swap_if_greater(hi - 1, hi); }
                else ls.insertion(lo, hi);
                continue;
            }
            if (depth == 0) { ls.heapsort(lo, hi); continue; }
            depth--;
            const int pv = ls.partition(lo, hi);
            if (hi > pv + 1) push(pv + 1, hi, depth);
            if (pv - 1 > lo) push(lo, pv - 1, depth);
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_counts2[0] = min(s_counts2[1], task_cap); s_counts2[1] = 0; s_nbig[bc] = 0; }
        int* tmp = cur; cur = nxt; nxt = tmp;
        bc ^= 1;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Block multi-select: exact order statistics of up to BMS_R requests.  Request r asks for rank rk[r]
// of the bins  core(range id) ∪ [xlo, xhi)  (a few extra bins right after the core).  All requests
// that still share (range, prefix) share one histogram.  MSD radix passes over the core ranges narrow
// every request to one bucket; as soon as the buckets still in play hold <= BMS_GCAP bins in total
// (immediately for short ranges, after ~3 passes for a whole chromosome arm) they are gathered into
// shared memory, sorted there, and every request reads its answer from the sorted list.
// ------------------------------------------------------------------------------------------------
constexpr int BMS_GCAP = 4096;        // gathered candidates, all groups together
constexpr int BMS_CAND = 2 * BMS_GCAP + 2 * BMS_R;  // slices are padded to powers of two

struct BmsState {
    unsigned hist[BMS_R][256];
    unsigned long long cand[BMS_CAND];
    unsigned long long gprefix[BMS_R];
    int grange[BMS_R];
    int gcount[BMS_R], goff[BMS_R], gcap[BMS_R];
    unsigned gcore[BMS_R];   // bins of the core range inside the group's current bucket
    int gbucket[BMS_R];      // indexed path: value bucket of the group
    int gpure[BMS_R];        // indexed path: the bucket holds one distinct key (answered from the splitter)
    int r_tl[2], r_tr[2];    // indexed path: full tiles [tl, tr) of each range
    int ngroups;
    int nranges;
    int rlo[2], rhi[2];
    int nreq;
    int rgrp[BMS_R];
    unsigned long long rk[BMS_R];
    int xlo[BMS_R], xhi[BMS_R];
    unsigned long long rkey[BMS_R];
    int rdig[BMS_R];
    int newgrp[BMS_R];
    unsigned long long dbg_ok, dbg_fallback, dbg_na_sum, dbg_na_max;
    int mode;        // 0: keep histogramming, 1: gather + sort, 2: all digits decided
    int match_all;   // gather without any decided digit
    int shift;
};

__device__ inline bool bms_match(unsigned long long key, unsigned long long prefix, int low_bit, int match_all) {
    return match_all || ((key ^ prefix) >> low_bit) == 0ull;
}

// thread 0: decide whether the buckets in play are small enough to gather
__device__ inline void bms_plan_gather(BmsState& st) {
    unsigned total = 0;
    for (int g = 0; g < st.ngroups; g++) total += st.gcore[g];
    if (total > (unsigned)BMS_GCAP) return;
    int off = 0;
    for (int g = 0; g < st.ngroups; g++) {
        int cap = 1;
        while (cap < (int)st.gcore[g]) cap <<= 1;
        st.gcap[g] = cap;
        st.goff[g] = off;
        st.gcount[g] = 0;
        off += cap;
    }
    st.mode = 1;
}

// Sort every group's gathered slice and answer each request: the k-th smallest of (sorted core
// candidates A) ∪ (the request's extra bins E that fall into the group's bucket); ties: A before E.
// Extras are matched by radix prefix (s_spl == nullptr) or by value bucket (indexed path).
__device__ void bms_sort_and_answer(BmsState& st, const double* __restrict__ x, int low_bit, int match_all,
                                    const unsigned long long* s_spl) {
    const int B = blockDim.x;
    const int ng = st.ngroups;
    for (int g = 0; g < ng; g++) {
        const int n2 = st.gcap[g];
        unsigned long long* a = st.cand + st.goff[g];
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < n2; i += B) {
                    const int l = i ^ j;
                    if (l > i) {
                        const bool up = (i & k) == 0;
                        const unsigned long long p = a[i], q = a[l];
                        if ((p > q) == up) { a[i] = q; a[l] = p; }
                    }
                }
                __syncthreads();
            }
    }
    __syncthreads();
    if ((int)threadIdx.x < st.nreq) {
        const int r = threadIdx.x;
        const int g = st.rgrp[r];
        const unsigned long long* a = st.cand + st.goff[g];
        const int na = min(st.gcount[g], st.gcap[g]);
        unsigned long long e[BMS_X];
        int ne = 0;
        for (int i = st.xlo[r]; i < st.xhi[r] && ne < BMS_X; i++) {
            const unsigned long long key = f64_key(x[i]);
            const bool in = s_spl ? (rq_bucket(s_spl, key) == st.gbucket[g]) : bms_match(key, st.gprefix[g], low_bit, match_all);
            if (in) {
                int j = ne++;
                while (j > 0 && e[j - 1] > key) { e[j] = e[j - 1]; j--; }
                e[j] = key;
            }
        }
        const long long k = (long long)st.rk[r];
        unsigned long long ans = na > 0 ? a[na - 1] : (ne > 0 ? e[ne - 1] : 0ull);
        bool found = false;
        for (int j = 0; j < ne && !found; j++) {
            // rank of e[j] = j + #{a <= e[j]}
            int lo = 0, hi = na;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= e[j]) lo = mid + 1; else hi = mid; }
            if ((long long)j + lo == k) { ans = e[j]; found = true; }
        }
        if (!found) {
            // rank of a[i] = i + #{e < a[i]}: monotone in i, find i with rank == k
            int lo = 0, hi = na - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                int less = 0;
                for (int j = 0; j < ne; j++) less += e[j] < a[mid];
                if ((long long)mid + less <= k) lo = mid; else hi = mid - 1;
            }
            if (na > 0) ans = a[lo];
        }
        st.rkey[r] = ans;
    }
    __syncthreads();
}

// caller (thread 0) fills nranges, rlo/rhi, nreq, rk, xlo/xhi and rgrp = range id; then all threads call
__device__ void bms_run(BmsState& st, const double* __restrict__ x) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int B = blockDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        st.ngroups = st.nranges;
        for (int g = 0; g < st.nranges; g++) {
            st.gprefix[g] = 0ull;
            st.grange[g] = g;
            st.gcore[g] = (unsigned)max(0, st.rhi[g] - st.rlo[g]);
        }
        st.mode = 0;
        st.match_all = 1;
        st.shift = 56;
        bms_plan_gather(st);  // short ranges: no histogram pass at all
    }
    __syncthreads();
    while (st.mode == 0) {
        const int shift = st.shift;
        const int ng = st.ngroups;
        const bool first = shift == 56;
        for (int t = threadIdx.x; t < ng * 256; t += B) st.hist[t >> 8][t & 255] = 0u;
        __syncthreads();
        for (int r = 0; r < st.nranges; r++) {
            const int lo = st.rlo[r], hi = st.rhi[r];
            // four independent loads in flight per thread: the range is L2-resident and one CTA has
            // to pull it through a single SM
            // (the loop bound is rounded up so that whole warps stay converged for the match below)
            const int span = hi - lo;
            const int span_round = ((span + 4 * B - 1) / (4 * B)) * (4 * B);
            for (int o0 = (int)threadIdx.x; o0 < span_round; o0 += 4 * B) {
                const int i0 = lo + o0;
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (i0 + u * B < hi) ? __ldg(x + i0 + u * B) : 0.0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool ok = i0 + u * B < hi;
                    const unsigned long long key = f64_key(v[u]);
                    const int d = (int)((key >> shift) & 255ull);
                    for (int g = 0; g < ng; g++) {
                        const bool hit = ok && st.grange[g] == r && (first || ((key ^ st.gprefix[g]) >> (shift + 8)) == 0ull);
                        // lanes that hit the same bin add once: noise-free stretches share most digits
                        const unsigned act = __ballot_sync(0xffffffffu, hit);
                        if (hit) {
                            const unsigned m = __match_any_sync(act, d);
                            if ((int)(__ffs(m) - 1) == lane) atomicAdd(&st.hist[g][d], (unsigned)__popc(m));
                        }
                    }
                }
            }
        }
        __syncthreads();
        // resolve: one warp per request; lane l owns bins 8l..8l+7 of the request's group, merged with
        // the request's extra bins
        for (int r = warp; r < st.nreq; r += (B >> 5)) {
            const int g = st.rgrp[r];
            const unsigned long long pref = st.gprefix[g];
            unsigned c[8];
            unsigned long long sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) c[t] = st.hist[g][lane * 8 + t];
            for (int i = st.xlo[r]; i < st.xhi[r]; i++) {
                const unsigned long long key = f64_key(x[i]);
                if (first || ((key ^ pref) >> (shift + 8)) == 0ull) {
                    const int d = (int)((key >> shift) & 255ull);
                    if ((d >> 3) == lane) c[d & 7]++;
                }
            }
#pragma unroll
            for (int t = 0; t < 8; t++) sum += c[t];
            unsigned long long incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            const unsigned long long excl = incl - sum;
            const unsigned long long k = st.rk[r];
            const bool mine = sum > 0 && k >= excl && k < excl + sum;
            unsigned b = __ballot_sync(0xffffffffu, mine);
            if (b == 0u) {
                const unsigned nz = __ballot_sync(0xffffffffu, sum > 0);
                b = nz ? (1u << (31 - __clz(nz))) : 1u;
            }
            const int owner = __ffs(b) - 1;
            if (lane == owner) {
                unsigned long long run = excl;
                int d = lane * 8 + 7;
                unsigned long long cum = excl + sum - c[7];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (k < run + c[t]) { d = lane * 8 + t; cum = run; break; }
                    run += c[t];
                }
                st.rk[r] = k >= cum ? k - cum : 0ull;
                st.rkey[r] = pref | ((unsigned long long)d << shift);
                st.rdig[r] = d;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // regroup by (range, new prefix); the core count of a new group is one histogram bin
            int nn = 0;
            unsigned long long np[BMS_R];
            int nr[BMS_R];
            unsigned ncore[BMS_R];
            for (int r = 0; r < st.nreq; r++) {
                const int og = st.rgrp[r];
                const int range = st.grange[og];
                int idx = -1;
                for (int q = 0; q < nn; q++)
                    if (np[q] == st.rkey[r] && nr[q] == range) idx = q;
                if (idx < 0) { idx = nn; np[nn] = st.rkey[r]; nr[nn] = range; ncore[nn] = st.hist[og][st.rdig[r]]; nn++; }
                st.newgrp[r] = idx;
            }
            for (int q = 0; q < nn; q++) { st.gprefix[q] = np[q]; st.grange[q] = nr[q]; st.gcore[q] = ncore[q]; }
            for (int r = 0; r < st.nreq; r++) st.rgrp[r] = st.newgrp[r];
            st.ngroups = nn;
            st.match_all = 0;
            if (shift == 0) st.mode = 2;
            else {
                bms_plan_gather(st);
                if (st.mode == 0) st.shift = shift - 8;
            }
        }
        __syncthreads();
    }
    if (st.mode != 1) return;
    // ---- gather the buckets in play, sort each group's slice, answer every request from it
    const int low_bit = st.shift;
    const int match_all = st.match_all;
    const int ng = st.ngroups;
    for (int g = 0; g < ng; g++)
        for (int t = st.goff[g] + (int)threadIdx.x; t < st.goff[g] + st.gcap[g]; t += B) st.cand[t] = ~0ull;
    __syncthreads();
    for (int r = 0; r < st.nranges; r++) {
        const int lo = st.rlo[r], hi = st.rhi[r];
        for (int i0 = lo + (int)threadIdx.x; i0 < hi; i0 += 4 * B) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (i0 + u * B < hi) ? __ldg(x + i0 + u * B) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i0 + u * B >= hi) break;
                const unsigned long long key = f64_key(v[u]);
                for (int g = 0; g < ng; g++)
                    if (st.grange[g] == r && bms_match(key, st.gprefix[g], low_bit, match_all)) {
                        const int idx = atomicAdd(&st.gcount[g], 1);
                        if (idx < st.gcap[g]) st.cand[st.goff[g] + idx] = key;
                    }
            }
        }
    }
    __syncthreads();
    bms_sort_and_answer(st, x, low_bit, match_all, nullptr);
}

// ------------------------------------------------------------------------------------------------
// The same requests answered from the range-quantile index (wavelet_rqindex.cuh): bucket counts of a
// range = difference of two cumulative rows + the bins of its two partial tiles; the bucket holding a
// wanted rank is gathered from the bucket-sorted tile copies.  Returns false (nothing answered) when
// the buckets in play hold more than BMS_GCAP bins — heavy duplicates — and the caller falls back to
// the radix passes of bms_run.
// ------------------------------------------------------------------------------------------------
struct RqChrom {  // index slices of one chromosome
    const unsigned long long* spl;
    const unsigned short* hist;
    const unsigned short* tstart;
    const unsigned* cum;
    const unsigned long long* sorted;
};

__device__ bool bms_run_indexed(BmsState& st, const double* __restrict__ x, const RqChrom& rq,
                                const unsigned long long* s_spl) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int B = blockDim.x;
    unsigned* cnt = &st.hist[0][0];  // [nranges][RQ_BUCKETS] bucket counts (the radix histograms are idle here)
    __syncthreads();
    // ---- 1. bucket counts of every range
    for (int r = 0; r < st.nranges; r++) {
        const int lo = st.rlo[r], hi = st.rhi[r];
        const int tl = (lo + RQ_TILE - 1) / RQ_TILE, tr = hi / RQ_TILE;
        if (threadIdx.x == 0) { st.r_tl[r] = tl; st.r_tr[r] = tr; }
        for (int b = threadIdx.x; b < RQ_BUCKETS; b += B)
            cnt[r * RQ_BUCKETS + b] = tl < tr ? rq.cum[(size_t)tr * RQ_BUCKETS + b] - rq.cum[(size_t)tl * RQ_BUCKETS + b] : 0u;
    }
    __syncthreads();
    for (int r = 0; r < st.nranges; r++) {
        const int lo = st.rlo[r], hi = st.rhi[r];
        const int tl = st.r_tl[r], tr = st.r_tr[r];
        // partial pieces: [lo, min(hi, tl*T)) and, when a tile boundary lies inside, [max(lo, tr*T), hi)
        const int a_end = tl <= tr ? min(hi, tl * RQ_TILE) : hi;
        for (int i = lo + (int)threadIdx.x; i < a_end; i += B) atomicAdd(&cnt[r * RQ_BUCKETS + rq_bucket(s_spl, f64_key(x[i]))], 1u);
        if (tl <= tr) {
            const int b_start = max(max(lo, tr * RQ_TILE), a_end);
            for (int i = b_start + (int)threadIdx.x; i < hi; i += B) atomicAdd(&cnt[r * RQ_BUCKETS + rq_bucket(s_spl, f64_key(x[i]))], 1u);
        }
    }
    __syncthreads();
    // ---- 2. one warp per request: bucket that holds its rank (extras merged in); lane l owns buckets 32l..32l+31
    constexpr int PER = RQ_BUCKETS / 32;
    for (int r = warp; r < st.nreq; r += (B >> 5)) {
        const unsigned* row = cnt + st.rgrp[r] * RQ_BUCKETS + lane * PER;
        int xb[BMS_X];
        int nx = 0;
        for (int i = st.xlo[r]; i < st.xhi[r] && nx < BMS_X; i++) xb[nx++] = rq_bucket(s_spl, f64_key(x[i]));
        unsigned long long sum = 0;
        for (int t = 0; t < PER; t++) sum += row[t];
        for (int q = 0; q < nx; q++) sum += (xb[q] / PER) == lane;
        unsigned long long incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const unsigned long long excl = incl - sum;
        const unsigned long long k = st.rk[r];
        const bool mine = sum > 0 && k >= excl && k < excl + sum;
        unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (bal == 0u) {
            const unsigned nz = __ballot_sync(0xffffffffu, sum > 0);
            bal = nz ? (1u << (31 - __clz(nz))) : 1u;
        }
        const int owner = __ffs(bal) - 1;
        if (lane == owner) {
            unsigned long long run = excl, cum = excl;
            int d = lane * PER + PER - 1;
            for (int t = 0; t < PER; t++) {
                unsigned long long c = row[t];
                for (int q = 0; q < nx; q++) c += xb[q] == lane * PER + t;
                if (k < run + c) { d = lane * PER + t; cum = run; break; }
                run += c;
                cum = run;
            }
            st.rkey[r] = k >= cum ? k - cum : 0ull;  // rank inside the bucket (committed only if we go on)
            st.rdig[r] = d;
        }
    }
    __syncthreads();
    // ---- 3. groups = distinct (range, bucket); plan the gather
    if (threadIdx.x == 0) {
        int nn = 0;
        int nb[BMS_R], nr[BMS_R];
        for (int r = 0; r < st.nreq; r++) {
            const int range = st.rgrp[r];
            int idx = -1;
            for (int q = 0; q < nn; q++)
                if (nb[q] == st.rdig[r] && nr[q] == range) idx = q;
            if (idx < 0) { idx = nn; nb[nn] = st.rdig[r]; nr[nn] = range; nn++; }
            st.newgrp[r] = idx;
        }
        unsigned total = 0;
        int pure[BMS_R];
        for (int q = 0; q < nn; q++) {
            // bucket b = keys in (spl[b-1], spl[b]]: one distinct key when the two splitters are adjacent
            pure[q] = nb[q] > 0 && s_spl[nb[q] - 1] + 1ull == s_spl[nb[q]];
            if (!pure[q]) total += cnt[nr[q] * RQ_BUCKETS + nb[q]];
        }
        st.mode = 0;
        if (total <= (unsigned)(BMS_CAND - 2 * BMS_R)) {
            int off = 0;
            for (int q = 0; q < nn; q++) {
                st.gbucket[q] = nb[q]; st.grange[q] = nr[q]; st.gpure[q] = pure[q];
                st.gcore[q] = pure[q] ? 0u : cnt[nr[q] * RQ_BUCKETS + nb[q]];
                st.gcap[q] = (int)st.gcore[q]; st.goff[q] = off; st.gcount[q] = 0;
                off += (int)st.gcore[q];
            }
            st.ngroups = nn;
            for (int r = 0; r < st.nreq; r++) { st.rgrp[r] = st.newgrp[r]; st.rk[r] = st.rkey[r]; }
            st.mode = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (st.mode == 1) { st.dbg_ok++; for (int q = 0; q < st.ngroups; q++) { if (st.gpure[q]) st.dbg_na_max += 1ull << 20; st.dbg_na_sum += st.gcore[q]; if (st.gcore[q] > st.dbg_na_max) st.dbg_na_max = st.gcore[q]; } }
        else st.dbg_fallback++;
    }
    if (st.mode != 1) return false;  // requests untouched: the caller falls back to the radix passes
    const int ng = st.ngroups;
    // ---- 4. gather: bucket slices of the full tiles + matching bins of the partial tiles
    for (int g = 0; g < ng; g++) {
        if (st.gpure[g]) continue;
        const int r = st.grange[g], b = st.gbucket[g];
        const int lo = st.rlo[r], hi = st.rhi[r];
        const int tl = st.r_tl[r], tr = st.r_tr[r];
        for (int t = tl + (int)threadIdx.x; t < tr; t += B) {
            const int c = rq.hist[(size_t)t * RQ_BUCKETS + b];
            if (c) {
                const int src = t * RQ_TILE + rq.tstart[(size_t)t * RQ_BUCKETS + b];
                const int dst = atomicAdd(&st.gcount[g], c);
                for (int q = 0; q < c; q++)
                    if (dst + q < st.gcap[g]) st.cand[st.goff[g] + dst + q] = rq.sorted[src + q];
            }
        }
        const int a_end = tl <= tr ? min(hi, tl * RQ_TILE) : hi;
        for (int i = lo + (int)threadIdx.x; i < a_end; i += B) {
            const unsigned long long key = f64_key(x[i]);
            if (rq_bucket(s_spl, key) == b) { const int dst = atomicAdd(&st.gcount[g], 1); if (dst < st.gcap[g]) st.cand[st.goff[g] + dst] = key; }
        }
        if (tl <= tr) {
            const int b_start = max(max(lo, tr * RQ_TILE), a_end);
            for (int i = b_start + (int)threadIdx.x; i < hi; i += B) {
                const unsigned long long key = f64_key(x[i]);
                if (rq_bucket(s_spl, key) == b) { const int dst = atomicAdd(&st.gcount[g], 1); if (dst < st.gcap[g]) st.cand[st.goff[g] + dst] = key; }
            }
        }
    }
    __syncthreads();
    // ---- 5. one warp per request: k-th smallest key of A ∪ E by binary radix descent over the gathered
    //         candidates (only the bits in which they differ): linear in the candidate count, no sorting
    for (int r = warp; r < st.nreq; r += (B >> 5)) {
        const int g = st.rgrp[r];
        if (st.gpure[g]) {
            if (lane == 0) st.rkey[r] = s_spl[st.gbucket[g]];
            continue;
        }
        const unsigned long long* a = st.cand + st.goff[g];
        const int na = min(st.gcount[g], st.gcap[g]);
        unsigned long long e[BMS_X];
        int ne = 0;
        for (int i = st.xlo[r]; i < st.xhi[r] && ne < BMS_X; i++) {
            const unsigned long long key = f64_key(x[i]);
            if (rq_bucket(s_spl, key) == st.gbucket[g]) e[ne++] = key;
        }
        unsigned long long vor = 0ull, vand = ~0ull;
        for (int i = lane; i < na; i += 32) { const unsigned long long v = a[i]; vor |= v; vand &= v; }
        for (int q = 0; q < ne; q++) { vor |= e[q]; vand &= e[q]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vor |= __shfl_xor_sync(0xffffffffu, vor, o);
            vand &= __shfl_xor_sync(0xffffffffu, vand, o);
        }
        const unsigned long long diff = vor ^ vand;   // bits that are not common to all keys
        unsigned long long k = st.rk[r], mask = 0ull, value = 0ull;
        for (int bit = 63; bit >= 0; bit--) {
            if (!((diff >> bit) & 1ull)) continue;
            const unsigned long long bm = 1ull << bit;
            unsigned c0 = 0;
            for (int i = lane; i < na; i += 32) {
                const unsigned long long v = a[i];
                c0 += (((v ^ value) & mask) == 0ull) && !(v & bm);
            }
            c0 = __reduce_add_sync(0xffffffffu, c0);
            for (int q = 0; q < ne; q++) c0 += (((e[q] ^ value) & mask) == 0ull) && !(e[q] & bm);
            if (k >= c0) { k -= c0; value |= bm; }
            mask |= bm;
        }
        if (lane == 0) st.rkey[r] = (vand & ~diff) | value;
    }
    __syncthreads();
    return true;
}

// median of x[lo, hi) from a pair of requests {lower middle, upper middle}
__device__ inline void bms_add_median(BmsState& st, int range, int len_total, int xlo, int xhi) {
    const int r = st.nreq;
    st.rgrp[r] = range; st.rgrp[r + 1] = range;
    st.rk[r] = (len_total & 1) ? len_total / 2 : len_total / 2 - 1;
    st.rk[r + 1] = len_total / 2;
    st.xlo[r] = st.xlo[r + 1] = xlo;
    st.xhi[r] = st.xhi[r + 1] = xhi;
    st.nreq = r + 2;
}
__device__ inline double bms_median_value(const BmsState& st, int r) {
    const double a = f64_unkey(st.rkey[r]), b = f64_unkey(st.rkey[r + 1]);
    return st.rkey[r] == st.rkey[r + 1] ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
}

__device__ inline int fin_block_scan(int v, int& total) {
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
        const int nw = (blockDim.x + 31) >> 5;
        int xv = lane < nw ? s_warp[lane] : 0;
        int xi = xv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, off);
            if (lane >= off) xi += t;
        }
        s_warp[lane] = xi - xv;
        if (lane == 31) s_total = xi;
    }
    __syncthreads();
    const int r = s_warp[w] + incl - v;
    total = s_total;
    __syncthreads();
    return r;
}

// bitonic sort of (key, value) pairs, n padded to a power of two by the caller with key = ~0
__device__ void fin_bitonic(unsigned long long* key, int* val, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[l];
                    if ((a > b) == up) {
                        key[i] = b; key[l] = a;
                        const int t = val[i]; val[i] = val[l]; val[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

#define FIN_STAMP(k)                                                                         \
    do {                                                                                     \
        __syncthreads();                                                                     \
        if (threadIdx.x == 0 && p.phase_ns) {                                                \
            unsigned long long t__;                                                          \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t__));                          \
            p.phase_ns[(size_t)(c_first + blockIdx.x) * 8 + (k)] = t__;                                  \
        }                                                                                    \
    } while (0)

__global__ void __launch_bounds__(FIN_THREADS, 1)
uh_finish_kernel(FinParams p, int c_first) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    BmsState& s_bms = *reinterpret_cast<BmsState*>(fin_smem);
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(fin_smem + ((sizeof(BmsState) + 15) & ~(size_t)15));
    int* s_val = reinterpret_cast<int*>(s_key + FIN_SORT_SMEM);
    unsigned long long* s_spl = reinterpret_cast<unsigned long long*>(s_val + FIN_SORT_SMEM);
    const int c = c_first + blockIdx.x;
    const long long o = p.off[c];
    const int n = (int)(p.off[c + 1] - o);
    if (threadIdx.x == 0) p.n_bp[c] = 0;
    if (!p.selected[c] || n <= p.min_size || n < 2) return;
    const double* __restrict__ x = p.x + o;
    int* lvl_idx = p.lvl_idx + o;
    int* sv = p.sv + o;
    unsigned long long* svkey = p.svkey + o;
    unsigned* bitmap = p.bitmap + (o >> 5) + c;
    int* piece = p.piece + o;
    double* rec = p.rec + o;
    int* prelim = p.prelim + o;
    int* lvl_first = p.lvl_first + o;
    int* bp = p.bp + o;
    const unsigned* lvlcnt = p.lvlcnt + o;
    const int T = p.depth[c];  // number of levels (tree.Count)
    for (int j = threadIdx.x; j < RQ_BUCKETS; j += blockDim.x) s_spl[j] = p.rq.spl[(size_t)c * RQ_BUCKETS + j];
    RqChrom rqc;
    {
        const int tf = p.rq.tfirst[c];
        rqc.spl = p.rq.spl + (size_t)c * RQ_BUCKETS;
        rqc.hist = p.rq.hist + (size_t)tf * RQ_BUCKETS;
        rqc.tstart = p.rq.tstart + (size_t)tf * RQ_BUCKETS;
        rqc.cum = p.rq.cum + (size_t)(tf + c) * RQ_BUCKETS;
        rqc.sorted = p.rq.sorted + o;
    }
    const UhCand* __restrict__ cand = p.cand + p.cp[c].cand_base;
    const int ncand_all = min(p.cc[c].cand_count_.v, p.cp[c].cand_cap);

    if (threadIdx.x == 0) { s_bms.dbg_ok = s_bms.dbg_fallback = s_bms.dbg_na_sum = s_bms.dbg_na_max = 0; }
    FIN_STAMP(0);
    // ---- HardThresh level weights (:78-91): germline only
    if (p.is_germline) {
        __shared__ int s_sort_counts[2];
        if (T <= FIN_SORT_SMEM) {
            // sort in shared memory (s_key doubles as the count table, s_val as the permutation, the
            // candidate buffer of the multi-select as the task lists)
            unsigned* s_cnt = reinterpret_cast<unsigned*>(s_key);
            for (int l = threadIdx.x; l < T; l += blockDim.x) { s_cnt[l] = lvlcnt[l]; s_val[l] = l; }
            __syncthreads();
            int* tasks = reinterpret_cast<int*>(s_bms.cand);
            const int cap = (int)(sizeof(s_bms.cand) / sizeof(int)) / 6;
            // scratch of the block partition: the upper half of s_key (s_cnt takes the lower) and the multi-select's histograms
            int* lpos = reinterpret_cast<int*>(s_key) + FIN_SORT_SMEM;
            int* rpos = reinterpret_cast<int*>(&s_bms.hist[0][0]);
            static_assert(sizeof(s_bms.hist) >= FIN_SORT_SMEM * sizeof(int), "right-stop scratch does not fit");
            level_sort_parallel(LevelSorter{s_val, s_cnt}, T, tasks, tasks + 3 * cap, cap, s_sort_counts, lpos, rpos);
            __syncthreads();
            for (int l = threadIdx.x; l < T; l += blockDim.x) lvl_idx[l] = s_val[l];
        } else {
            // deeper than the shared-memory tables: same replay on global scratch (rec doubles as the
            // moving count table)
            unsigned* gcnt = reinterpret_cast<unsigned*>(rec);
            for (int l = threadIdx.x; l < T; l += blockDim.x) { lvl_idx[l] = l; gcnt[l] = lvlcnt[l]; }
            __syncthreads();
            const int cap = n / 3;
            // large sub-arrays are partitioned by the whole block here too (stop positions in the survivor scratch, which
            // is not in use yet): a run of a few thousand exact zeros (chrY of a female sample) makes the tree that deep,
            // and one thread partitioning thousands of levels out of L2 took 1.6 ms
            level_sort_parallel(LevelSorter{lvl_idx, gcnt}, T, piece, prelim, cap, s_sort_counts, sv, reinterpret_cast<int*>(svkey));
        }
    }
    __syncthreads();

    FIN_STAMP(1);
    // ---- survivors of the threshold among this chromosome's candidates (:102-114)
    const double sigma = p.sigma[c];
    const double root = sqrt(2.0 * log((double)n));
    int K = 0;
    for (int base = 0; base < ncand_all; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        if (i < ncand_all) {
            const UhCand k = cand[i];
            double w = 1.0;
            if (p.is_germline) w = __dadd_rn(__ddiv_rn(__dmul_rn((double)(lvl_idx[k.level] + 1), 1.0 - 0.8), (double)T), 0.8);
            const double thr = __dmul_rn(__dmul_rn(__dmul_rn(2.0, sigma), w), root);
            keep = !(fabs(k.coef) <= thr);
        }
        int total;
        const int ex = fin_block_scan(keep ? 1 : 0, total);
        if (keep) {
            sv[K + ex] = i;
            svkey[K + ex] = ((unsigned long long)cand[i].level << 32) | (unsigned)cand[i].s;
        }
        K += total;
    }
    __syncthreads();

    // ---- sort survivors by (level, start)
    if (K > 1) {
        int n2 = 1;
        while (n2 < K) n2 <<= 1;
        if (n2 <= FIN_SORT_SMEM) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                s_key[i] = i < K ? svkey[i] : ~0ull;
                s_val[i] = i < K ? sv[i] : -1;
            }
            __syncthreads();
            fin_bitonic(s_key, s_val, n2);
            for (int i = threadIdx.x; i < K; i += blockDim.x) { svkey[i] = s_key[i]; sv[i] = s_val[i]; }
        } else if (n2 <= n) {
            for (int i = K + threadIdx.x; i < n2; i += blockDim.x) { svkey[i] = ~0ull; sv[i] = -1; }
            __syncthreads();
            fin_bitonic(svkey, sv, n2);
        } else {
            // more survivors than fit the padded scratch: cannot happen (K <= n - 1, n2 < 2n) unless
            // n2 > n; fall back to an odd-even transposition sort in place
            for (int pass = 0; pass < K; pass++) {
                for (int i = (pass & 1) + 2 * (int)threadIdx.x; i + 1 < K; i += 2 * blockDim.x)
                    if (svkey[i] > svkey[i + 1]) {
                        const unsigned long long a = svkey[i]; svkey[i] = svkey[i + 1]; svkey[i + 1] = a;
                        const int t = sv[i]; sv[i] = sv[i + 1]; sv[i + 1] = t;
                    }
                __syncthreads();
            }
        }
        __syncthreads();
    }

    FIN_STAMP(2);
    // ---- pieces: positions where a surviving node starts, breaks or ends (+ position 0)
    const int nwords = (n + 31) >> 5;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) atomicOr(&bitmap[0], 1u);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const UhCand k = cand[sv[i]];
        const int q[3] = {k.s, k.b + 1, k.e + 1};
        for (int t = 0; t < 3; t++)
            if (q[t] > 0 && q[t] < n) atomicOr(&bitmap[q[t] >> 5], 1u << (q[t] & 31));
    }
    __syncthreads();
    int P = 0;
    for (int base = 0; base < nwords; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned wbits = i < nwords ? bitmap[i] : 0u;
        int total;
        int ex = fin_block_scan(__popc(wbits), total);
        unsigned b = wbits;
        while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1;
            piece[P + ex++] = (i << 5) + bit;
        }
        P += total;
    }
    __syncthreads();

    // ---- distinct levels among the sorted survivors
    int NL = 0;
    for (int base = 0; base < K; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const bool head = i < K && (i == 0 || (svkey[i] >> 32) != (svkey[i - 1] >> 32));
        int total;
        const int ex = fin_block_scan(head ? 1 : 0, total);
        if (head) lvl_first[NL + ex] = i;
        NL += total;
    }
    __syncthreads();

    // ---- reconstruction on the pieces (:138-168): additions in level order, exactly one node per level
    {
        const long long p0 = o + c;
        const double total_sum = p.pz[p0 + n] - p.pz[p0];
        const double smooth = total_sum / sqrt((double)n);              // :376-378
        const double base_val = __dmul_rn(__ddiv_rn(1.0, sqrt((double)n)), smooth);  // :146
        for (int pi = threadIdx.x; pi < P; pi += blockDim.x) {
            const int q = piece[pi];
            double r = base_val;
            for (int l = 0; l < NL; l++) {
                const int lo = lvl_first[l], hi = (l + 1 < NL) ? lvl_first[l + 1] : K;
                // last survivor of this level with start <= q
                int a = lo, b = hi;
                while (a < b) {
                    const int mid = (a + b) >> 1;
                    if ((int)(svkey[mid] & 0xffffffffu) <= q) a = mid + 1; else b = mid;
                }
                if (a == lo) continue;
                const UhCand k = cand[sv[a - 1]];
                if (q > k.e) continue;
                const double nn = (double)(k.e - k.s + 1), m = (double)(k.b - k.s + 1);
                const double v = q <= k.b ? sqrt(__dsub_rn(__ddiv_rn(1.0, m), __ddiv_rn(1.0, nn)))
                                          : __ddiv_rn(-1.0, sqrt(__dsub_rn(__ddiv_rn(__dmul_rn(nn, nn), m), nn)));
                r = __dadd_rn(r, __dmul_rn(v, k.coef));
            }
            rec[pi] = r;
        }
    }
    __syncthreads();

    // ---- preliminary breakpoints (:174-185)
    int L = 0;
    for (int base = 0; base < P; base += blockDim.x) {
        const int pi = base + threadIdx.x;
        const bool isbp = pi < P && (pi == 0 || __dsub_rn(rec[pi], rec[pi - 1]) != 0.0);
        int total;
        const int ex = fin_block_scan(isbp ? 1 : 0, total);
        if (isbp) prelim[L + ex] = piece[pi];
        L += total;
    }
    __syncthreads();

    FIN_STAMP(3);
    // ---- healing (:194-232): greedy left to right, exact medians of the two implied segments
    int nb = 0;  // kept breakpoints so far (uniform across the block)
    if (threadIdx.x == 0) bp[0] = prelim[0];
    nb = 1;
    {
        double left_median_cache = 0.0;
        int cache_lo = -1, cache_hi = -1;
        for (int i = 1; i < L; i++) {
            __syncthreads();
            const int left_start = bp[nb - 1];
            const int right_start = prelim[i];
            const int right_end = (i < L - 1) ? prelim[i + 1] : n;
            const int left_len = right_start - left_start, right_len = right_end - right_start;
            const bool have_left = cache_lo == left_start && cache_hi == right_start;
            if (threadIdx.x == 0) {
                s_bms.nreq = 0;
                s_bms.nranges = have_left ? 1 : 2;
                s_bms.rlo[0] = right_start; s_bms.rhi[0] = right_end;
                bms_add_median(s_bms, 0, right_len, 0, 0);
                if (!have_left) {
                    s_bms.rlo[1] = left_start; s_bms.rhi[1] = right_start;
                    bms_add_median(s_bms, 1, left_len, 0, 0);
                }
            }
            if (!bms_run_indexed(s_bms, x, rqc, s_spl)) bms_run(s_bms, x);
            const double rm = bms_median_value(s_bms, 0);
            const double lm = have_left ? left_median_cache : bms_median_value(s_bms, 2);
            const double wm = __ddiv_rn(__dadd_rn(__dmul_rn((double)left_len, lm), __dmul_rn((double)right_len, rm)),
                                        (double)(right_end - left_start));
            const int smaller = min(left_len, right_len);
            int scale = (int)ceil(__ddiv_rn(log((double)smaller), log(3.0)));
            // exact powers of three: log(3^k)/log(3) must not depend on the last bit of log()
            {
                int pw = 1, kk = 0;
                while (pw < smaller && kk < 19) { pw *= 3; kk++; }
                if (pw == smaller) scale = p.log3_scale_tab ? (int)p.log3_scale_tab[kk] : kk;
            }
            scale = min(WV_F3_LEVELS, scale);
            const double cutoff = p.ctl->f3[scale];
            const bool keep = fabs(__dsub_rn(lm, rm)) > __dmul_rn(__dmul_rn(cutoff, 4.0), fmax(wm, 50.0));
            __syncthreads();
            if (keep) {
                if (threadIdx.x == 0) bp[nb] = right_start;
                nb++;
                // the right segment becomes the next left segment
                left_median_cache = rm; cache_lo = right_start; cache_hi = right_end;
            } else {
                cache_lo = cache_hi = -1;
            }
        }
    }
    __syncthreads();

    FIN_STAMP(4);
    // ---- RefineSegments (:237-258), germline only
    if (p.is_germline && nb > 2) {
        const double total_median = p.chrom_median[c];
        for (int i = 1; i < nb - 1; i++) {
            __syncthreads();
            const int prev = bp[i - 1], cur = bp[i], next = bp[i + 1];
            const int li = min(5, (cur - prev) / 2), ri = min(5, (next - cur) / 2);
            const int core_hi = cur - li;
            if (threadIdx.x == 0) {
                s_bms.nreq = 0;
                s_bms.nranges = 1;
                s_bms.rlo[0] = prev; s_bms.rhi[0] = core_hi;
                bms_add_median(s_bms, 0, cur - prev, core_hi, cur);             // Median(coverage, prev, cur)
                for (int j = cur - li; j < cur + ri; j++) bms_add_median(s_bms, 0, j - prev, core_hi, j);
            }
            if (!bms_run_indexed(s_bms, x, rqc, s_spl)) bms_run(s_bms, x);
            if (threadIdx.x == 0) {
                double best = fabs(__dsub_rn(bms_median_value(s_bms, 0), total_median));
                int best_bp = cur;
                int r = 2;
                for (int j = cur - li; j < cur + ri; j++, r += 2) {
                    const double d = fabs(__dsub_rn(bms_median_value(s_bms, r), total_median));
                    if (d > best) { best = d; best_bp = j; }
                }
                bp[i] = best_bp;
            }
            __syncthreads();
        }
    }
    FIN_STAMP(5);
    if (threadIdx.x == 0 && p.phase_ns) {
        p.phase_ns[(size_t)(c_first + blockIdx.x) * 8 + 6] = (s_bms.dbg_fallback << 32) | s_bms.dbg_ok;
        p.phase_ns[(size_t)(c_first + blockIdx.x) * 8 + 7] = (s_bms.dbg_na_max << 32) | s_bms.dbg_na_sum;
    }
    if (threadIdx.x == 0) p.n_bp[c] = nb;
}

inline size_t fin_smem_bytes() {
    return ((sizeof(BmsState) + 15) & ~(size_t)15) + (size_t)FIN_SORT_SMEM * (sizeof(unsigned long long) + sizeof(int)) +
           (size_t)RQ_BUCKETS * sizeof(unsigned long long);
}
