// Unbalanced-Haar top-down decomposition (WaveletSegmentation.cs:264-379) as a pipeline of four
// kernels, one per node granularity, connected by task lists (a kernel boundary is the only
// synchronisation between stages; inside a stage nothing waits for another SM except stage A's ring).
//
// The reference walks the tree level by level and, for every node, evaluates the inner products
// with all n-1 Unbalanced-Haar vectors by a sequential recurrence (GetInnerProdIter :19-48), then
// takes the first arg-max of |ip| (:54-67).  The tree is 130-220 levels deep on WGS data because
// noise splits are lopsided: a long chain of big nodes with small subtrees peeling off.  Closed form:
// with P the prefix sums of the node, a = m+1, b = n-a,
//     ip[m] = (P_m - a*T/n) * sqrt(n / (a*b)),   arg-max |ip| = arg-max (P_m - a*T/n)^2 / (a*b)
// so a node costs one pass over its slice of the chromosome-wide prefix-sum array (L2-resident).
//
//   A  uh_chain_kernel  n > UH_MID_MAX      one thread-block CLUSTER per chain of big nodes: every CTA
//        scans a slice (16 loads in flight per thread; one SM alone gets ~80-120 GB/s out of L2), posts
//        its best to all CTAs through distributed shared memory, one barrier.cluster per node; the
//        prefix values bounding the children are carried from the arg-max winner, never re-read
//   M  uh_mid_kernel    UH_SMALL_MAX < n <= UH_MID_MAX   one CTA per subtree, depth first on a small
//        shared-memory stack, one __syncthreads per node
//   S  uh_small_kernel  UH_TINY_MAX < n <= UH_SMALL_MAX  one warp per subtree, depth first
//   T  uh_tiny_kernel   n <= UH_TINY_MAX    one thread per subtree, the reference's own recurrence
//        operation for operation, so exact ties in tiny nodes (symmetric patterns on 2-decimal data)
//        break exactly as in the reference
// A run of exact zeros is emitted at once as the reference's one-bin-per-level comb.  Nodes are not
// stored: only per-level node counts (HardThresh's germline weights need them) and the few nodes
// whose coefficient can survive the threshold ("candidates").
//
// Every kernel works on ONE chromosome (argument c): the host runs the four stages of a chromosome back to back on that
// chromosome's stream, followed by its finish kernel, and the chromosomes' pipelines run side by side — a stage of one
// chromosome no longer waits for the slowest chromosome of the stage before it.  Queue counters and list slices are per
// chromosome (UhChromCtl / UhChromPlan).
#pragma once
#include <cooperative_groups.h>

#include "wavelet.cuh"

struct UhParams {
    const double* x;         // coverage [N]
    const double* pz;        // prefix sums with leading zero per chromosome [N + n_chrom]
    const long long* off;    // [n_chrom + 1]
    const double* cand_thr;  // [n_chrom]
    unsigned* lvlcnt;        // [N]: node count of level l of chromosome c at off[c] + l
    UhBigTask* big;          // stage A rings [UH_QCAP], one slice per chromosome; c < 0 = empty slot
    UhTask* mid;             // stage M lists, one slice per chromosome
    UhTask* small;           // stage S lists
    UhTinyTask* tiny;        // stage T lists
    UhCand* cand;            // candidate lists
    UhChromCtl* cc;          // [n_chrom] queue counters
    const UhChromPlan* cp;   // [n_chrom] slices of the arrays above
    WvCtl* ctl;
    unsigned long long* tl_ns;  // debug timeline (nullable): [n_chrom][16], per stage k: [2k] = ~(earliest start), [2k+1] = latest end
    unsigned long long* task_dbg;  // debug (nullable): [0] = entries used, then per mid task {chrom << 32 | bins, nodes << 32 | ns}
};
constexpr int UH_TASK_DBG_CAP = 16384;

// debug timeline of the per-chromosome pipelines (CANVAS_DEBUG): called by one thread per block
__device__ inline void uh_stamp(unsigned long long* tl, int c, int stage, bool end) {
    if (!tl) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMax(&tl[(size_t)c * 16 + 2 * stage + (end ? 1 : 0)], end ? t : ~t);
}

__device__ inline void uh_emit_candidate(const UhParams& p, int c, int level, int s, int b, int e, double coef) {
    if (fabs(coef) <= p.cand_thr[c]) return;  // NaN falls through on purpose (never zeroed by HardThresh)
    const UhChromPlan& cp = p.cp[c];
    const int i = atomicAdd(&p.cc[c].cand_count_.v, 1);
    if (i >= cp.cand_cap) { p.ctl->overflow_.v = 1; return; }
    UhCand k;
    k.key = ((unsigned long long)c << 56) | ((unsigned long long)level << 32) | (unsigned)s;
    k.s = s; k.b = b; k.e = e; k.level = level; k.c = c; k.pad = 0; k.coef = coef;
    p.cand[cp.cand_base + i] = k;
}

// Chains and the mid stage run before the thresholds exist (they overlap the order statistics that produce them): every one of
// their nodes — a few hundred per chromosome — is recorded, and the finish stage applies the real threshold as to any candidate
// ... without stalling the node loop on the slot counter: the atomic that reserves a slot takes a round trip to L2 (~0.5 us,
// a sixth of a node), so the record of node i waits in registers and is written when node i+1 is emitted — by then the slot
// index has long arrived.  The last record is flushed when the kernel ends.
struct UhPendingCand {
    UhCand k;
    int slot;
    bool has;
};

__device__ inline void uh_cand_flush(const UhParams& p, const UhChromPlan& cp, UhPendingCand& pc) {
    if (!pc.has) return;
    if (pc.slot >= cp.cand_cap) p.ctl->overflow_.v = 1;
    else p.cand[cp.cand_base + pc.slot] = pc.k;
    pc.has = false;
}

__device__ inline void uh_cand_defer(const UhParams& p, UhChromCtl* cc, const UhChromPlan& cp, UhPendingCand& pc, int c, int level,
                                     int s, int b, int e, double coef) {
    uh_cand_flush(p, cp, pc);
    pc.k.key = ((unsigned long long)c << 56) | ((unsigned long long)level << 32) | (unsigned)s;
    pc.k.s = s; pc.k.b = b; pc.k.e = e; pc.k.level = level; pc.k.c = c; pc.k.pad = 0; pc.k.coef = coef;
    pc.slot = atomicAdd(&cc->cand_count_.v, 1);
    pc.has = true;
}

// Can the node's coefficient exceed the candidate threshold at all?  score = D^2 / (a b) of the chosen split, and
// coef^2 = score * n / scale^2 with scale = max(0.5, mean / 200): three multiplications instead of the square root and two
// divisions of the exact coefficient, which only the few nodes that pass (or a NaN) still compute.  The factor leaves the
// rounding of the exact formula far inside the margin.
__device__ inline bool uh_may_be_candidate(double score, double nn, double mu, double thr2) {
    const double sc = fmax(0.5, mu * 0.005);
    return !(score * nn < 0.999999 * thr2 * sc * sc);
}

// Stage-A ring: a producer reserves a position with one atomicAdd, fills the slot and publishes it by
// storing the chromosome id last; a consumer reserves a position the same way and waits for it.
__device__ inline void uh_push_big(const UhParams& p, int c, int s, int e, int level, double base, double endv) {
    const UhChromPlan& cp = p.cp[c];
    const unsigned long long pos = atomicAdd(&p.cc[c].q_tail_.v, 1ull);
    UhBigTask* t = p.big + cp.ring_base + (int)(pos & (unsigned long long)(cp.ring_cap - 1));
    while (*(volatile int*)&t->c >= 0) __nanosleep(64);  // slot still held by an unconsumed task (ring wrapped)
    t->s = s; t->e = e; t->level = level; t->base = base; t->endv = endv;
    __threadfence();
    *(volatile int*)&t->c = c;
}

// Lists for the later stages are only read after the producing kernel has finished: no publication
// protocol, and producers batch their appends in shared memory (one atomicAdd per flush).
template <typename T, int CAP>
struct UhOutBuf {
    T item[CAP];
    int count;
};

template <typename T, int CAP>
__device__ inline void uh_buf_flush(UhOutBuf<T, CAP>& b, T* list, int list_cap, int* tail, WvCtl* ctl) {
    const int n = b.count;
    if (n == 0) return;
    const int base = atomicAdd(tail, n);
    for (int i = 0; i < n; i++) {
        if (base + i < list_cap) list[base + i] = b.item[i];
        else ctl->overflow_.v = 1;
    }
    b.count = 0;
}

template <typename T, int CAP>
__device__ inline void uh_buf_push(UhOutBuf<T, CAP>& b, const T& t, T* list, int list_cap, int* tail, WvCtl* ctl) {
    if (b.count == CAP) uh_buf_flush(b, list, list_cap, tail, ctl);
    b.item[b.count++] = t;
}

// (score, m) arg-max inside a warp: largest score, smallest m among equals.  Scores are >= 0
// (or NaN, mapped to 0), so their bit patterns order like unsigned integers.
__device__ inline void warp_argmax(double& score, int& m) {
    if (!(score >= 0.0)) score = 0.0;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(score);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo_c = hi == mhi ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, lo_c);
    const bool top = hi == mhi && lo == mlo;
    const unsigned mm = __reduce_min_sync(0xffffffffu, top ? (unsigned)m : 0xffffffffu);
    score = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    m = (int)mm;
}

// Scan split positions [m_lo, m_hi) of a node with NT cooperating threads, U loads in flight each.
// Candidates are compared inside a thread by cross-multiplication (no division in the loop).
// GLOBAL = false: q points into shared memory (a subtree's prefix sums staged there).
template <int NT, int U, bool GLOBAL = true>
__device__ inline void uh_scan(const double* __restrict__ q, int m_lo, int m_hi, int tid, double base, double mu, double nn,
                               double& bnum, double& bden, int& bm, double& bv) {
    int m = m_lo + tid;
    for (; m + (U - 1) * NT < m_hi; m += U * NT) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = GLOBAL ? __ldg(q + m + u * NT) : q[m + u * NT];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double a = (double)(m + u * NT + 1);
            const double D = fma(-a, mu, v[u] - base);
            const double num = D * D, den = a * (nn - a);
            if (num * bden > bnum * den) { bnum = num; bden = den; bm = m + u * NT; bv = v[u]; }
        }
    }
    for (; m < m_hi; m += NT) {
        const double vv = GLOBAL ? __ldg(q + m) : q[m];
        const double a = (double)(m + 1);
        const double D = fma(-a, mu, vv - base);
        const double num = D * D, den = a * (nn - a);
        if (num * bden > bnum * den) { bnum = num; bden = den; bm = m; bv = vv; }
    }
}

// coefficient of the chosen split (closed form of :36-45, scaling of :282-283)
__device__ inline double uh_coef_from(double fv, double base, double T, int n, int fm) {
    const double nn = (double)n, a = (double)(fm + 1), b = (double)(n - fm - 1);
    const double mu = T / nn;
    const double D = fma(-a, mu, fv - base);
    return D * sqrt(nn / (a * b)) / fmax(0.5, mu / 200.0);
}

// which stage a node of n bins belongs to
enum { UH_TIER_NONE = 0, UH_TIER_TINY, UH_TIER_SMALL, UH_TIER_MID, UH_TIER_BIG };
__device__ inline int uh_tier(int n) {
    if (n < 2) return UH_TIER_NONE;
    if (n <= UH_TINY_MAX) return UH_TIER_TINY;
    if (n <= UH_SMALL_MAX) return UH_TIER_SMALL;
    if (n <= UH_MID_MAX) return UH_TIER_MID;
    return UH_TIER_BIG;
}

// ---------------------------------------------------------------------------------------------
// Stage A — chains of big nodes on thread-block clusters
// ---------------------------------------------------------------------------------------------
struct UhMail {
    double score, v;
    int m, pad;
};

__global__ void __cluster_dims__(UH_CLUSTER, 1, 1) __launch_bounds__(UH_THREADS, 2)
uh_chain_kernel(UhParams p, int c_self) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    __shared__ UhMail s_mail[2][UH_CLUSTER];  // written by every CTA of the cluster (DSMEM)
    __shared__ int s_task[4];                  // the leader's copy is the authoritative one
    __shared__ double s_taskd[2];
    __shared__ double s_ws[UH_THREADS / 32], s_wv[UH_THREADS / 32];
    __shared__ int s_wm2[UH_THREADS / 32];
    __shared__ UhOutBuf<UhTask, 32> s_mid_out, s_small_out;
    __shared__ UhOutBuf<UhTinyTask, 32> s_tiny_out;
    WvCtl* ctl = p.ctl;
    UhChromCtl* cc = p.cc + c_self;
    const UhChromPlan cp = p.cp[c_self];
    UhTask* const mid_list = p.mid + cp.mid_base;
    UhTask* const small_list = p.small + cp.small_base;
    UhTinyTask* const tiny_list = p.tiny + cp.tiny_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* __restrict__ pz = p.pz;
    unsigned long long v_big = 0, n_big = 0;
    UhPendingCand pend;
    pend.has = false;
    if (threadIdx.x == 0) {
        s_mid_out.count = 0; s_small_out.count = 0; s_tiny_out.count = 0;
        uh_stamp(p.tl_ns, c_self, 0, false);
        if (crank == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            atomicMin(&ctl->t_first, t);
        }
    }
    int parity = 0;
    for (;;) {
        // ---- the leader CTA takes a big task; everybody reads it from the leader's shared memory
        if (crank == 0 && threadIdx.x == 0) {
            int c = -1;
            const unsigned long long pos = atomicAdd(&cc->q_head_.v, 1ull);
            UhBigTask* t = p.big + cp.ring_base + (int)(pos & (unsigned long long)(cp.ring_cap - 1));
            volatile int* ready = &t->c;
            unsigned polls = 0;
            for (;;) {
                c = *ready;
                if (c >= 0) break;
                if ((++polls & 3u) == 0u && (*(volatile int*)&cc->big_done_.v || *(volatile int*)&ctl->overflow_.v)) {
                    __threadfence();
                    c = *ready;
                    break;
                }
                __nanosleep(40);
            }
            if (c >= 0) {
                __threadfence();
                s_task[1] = *(volatile int*)&t->s; s_task[2] = *(volatile int*)&t->e; s_task[3] = *(volatile int*)&t->level;
                s_taskd[0] = *(volatile double*)&t->base; s_taskd[1] = *(volatile double*)&t->endv;
                __threadfence();
                *ready = -1;  // free the slot
            }
            s_task[0] = c;
        }
        cluster.sync();
        const int* lt = cluster.map_shared_rank(s_task, 0);
        const double* ltd = cluster.map_shared_rank(s_taskd, 0);
        const int c = lt[0];
        int s = lt[1], e = lt[2], level = lt[3];
        double base = ltd[0], endv = ltd[1];
        cluster.sync();  // the leader may overwrite its task slot only after everybody has read it
        if (c < 0) break;
        const long long p0 = p.off[c] + c;
        const long long loff = p.off[c];
        // ---- walk the chain
        for (;;) {
            const int n = e - s + 1;
            const double nn = (double)n;
            const double T = endv - base;
            const double mu = T / nn;
            const double* __restrict__ q = pz + p0 + s + 1;
            int per = (n - 1 + UH_CLUSTER - 1) / UH_CLUSTER;
            per = (per + UH_THREADS - 1) / UH_THREADS * UH_THREADS;
            const int m_lo = crank * per;
            const int m_hi = min(n - 1, m_lo + per);
            double bnum = -1.0, bden = 1.0, bv = 0.0;
            int bm = 0x7fffffff;
            uh_scan<UH_THREADS, 16>(q, m_lo, m_hi, threadIdx.x, base, mu, nn, bnum, bden, bm, bv);
            double best = bnum >= 0.0 ? bnum / bden : -1.0;
            int best_m = bm;
            warp_argmax(best, best_m);
            {
                const unsigned own = __ballot_sync(0xffffffffu, bm == best_m && best_m != 0x7fffffff);
                const double wv = __shfl_sync(0xffffffffu, bv, own ? __ffs(own) - 1 : 0);
                if (lane == 0) { s_ws[warp] = best; s_wm2[warp] = best_m; s_wv[warp] = wv; }
            }
            __syncthreads();
            if (warp == 0) {
                // CTA best, then post it to every CTA of the cluster (lane d writes to CTA d)
                double cb = lane < UH_THREADS / 32 ? s_ws[lane] : 0.0;
                int cm = lane < UH_THREADS / 32 ? s_wm2[lane] : 0x7fffffff;
                const int mine_m = cm;
                warp_argmax(cb, cm);
                const unsigned own = __ballot_sync(0xffffffffu, lane < UH_THREADS / 32 && mine_m == cm && cm != 0x7fffffff);
                const int src = own ? __ffs(own) - 1 : 0;
                const double cv = s_wv[src < UH_THREADS / 32 ? src : 0];
                if (lane < UH_CLUSTER) {
                    UhMail* dst = cluster.map_shared_rank(&s_mail[parity][crank], lane);
                    dst->score = cb; dst->v = cv; dst->m = cm;
                }
            }
            cluster.sync();
            double fbest = s_mail[parity][0].score, fv = s_mail[parity][0].v;
            int fm = s_mail[parity][0].m;
#pragma unroll
            for (int w = 1; w < UH_CLUSTER; w++) {
                const double sc = s_mail[parity][w].score;
                const int mm = s_mail[parity][w].m;
                if (sc > fbest || (sc == fbest && mm < fm)) { fbest = sc; fm = mm; fv = s_mail[parity][w].v; }
            }
            parity ^= 1;
            if (threadIdx.x == 0 && crank == 0) v_big += (unsigned long long)n;
            if (fbest == 0.0 || fm == 0x7fffffff) {
                // run of exact zeros: comb of n-1 nodes with coefficient 0 — one node per level
                for (int k = crank * UH_THREADS + threadIdx.x; k < n - 1; k += UH_CLUSTER * UH_THREADS)
                    atomicAdd(&p.lvlcnt[loff + level + k], 1u);
                if (threadIdx.x == 0 && crank == 0) n_big += (unsigned long long)(n - 1);
                break;  // chain ends
            }
            const int ls = s, le = s + fm, rs = s + fm + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            const int lt_ = uh_tier(ln), rt_ = uh_tier(rn);
            const bool lbig = lt_ == UH_TIER_BIG, rbig = rt_ == UH_TIER_BIG;
            const bool cont_left = lbig && (!rbig || ln >= rn);  // larger big child, left on ties
            const bool cont_right = rbig && !cont_left;
            if (crank == 0 && threadIdx.x == 32) {
                atomicAdd(&p.lvlcnt[loff + level], 1u);
                uh_cand_defer(p, cc, cp, pend, c, level, s, s + fm, e, uh_coef_from(fv, base, T, n, fm));
                n_big++;
            }
            if (crank == 0 && threadIdx.x == 0) {
                // children that are not continued: other big child -> ring, the rest -> stage lists
                if (lbig && !cont_left) { atomicAdd(&cc->outstanding_.v, 1); uh_push_big(p, c, ls, le, level + 1, base, fv); }
                if (rbig && !cont_right) { atomicAdd(&cc->outstanding_.v, 1); uh_push_big(p, c, rs, re, level + 1, fv, endv); }
                const UhTask tl = {c, ls, le, level + 1, base, fv}, tr = {c, rs, re, level + 1, fv, endv};
                if (lt_ == UH_TIER_MID) uh_buf_push(s_mid_out, tl, mid_list, cp.mid_cap, &cc->mid_tail_.v, ctl);
                if (rt_ == UH_TIER_MID) uh_buf_push(s_mid_out, tr, mid_list, cp.mid_cap, &cc->mid_tail_.v, ctl);
                if (lt_ == UH_TIER_SMALL) uh_buf_push(s_small_out, tl, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
                if (rt_ == UH_TIER_SMALL) uh_buf_push(s_small_out, tr, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
                const UhTinyTask yl = {c, ls, le, level + 1}, yr = {c, rs, re, level + 1};
                if (lt_ == UH_TIER_TINY) uh_buf_push(s_tiny_out, yl, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
                if (rt_ == UH_TIER_TINY) uh_buf_push(s_tiny_out, yr, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
            }
            if (cont_left) { e = le; endv = fv; level++; }
            else if (cont_right) { s = rs; base = fv; level++; }
            else break;  // no big child: chain ends
        }
        // ---- chain finished: one big task less
        if (crank == 0 && threadIdx.x == 0) {
            __threadfence();
            const int now = atomicSub(&cc->outstanding_.v, 1) - 1;
            if (now == 0) {
                __threadfence();
                *(volatile int*)&cc->big_done_.v = 1;
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                atomicMax(&ctl->t_big_done, t);
            }
        }
    }
    if (crank == 0 && threadIdx.x == 0) {
        uh_buf_flush(s_mid_out, mid_list, cp.mid_cap, &cc->mid_tail_.v, ctl);
        uh_buf_flush(s_small_out, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
        uh_buf_flush(s_tiny_out, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
    }
    uh_cand_flush(p, cp, pend);
    if (v_big) atomicAdd(&ctl->visits_big, v_big);
    if (n_big) atomicAdd(&ctl->nodes_big, n_big);
    if (threadIdx.x == 0) uh_stamp(p.tl_ns, c_self, 0, true);
}

// ---------------------------------------------------------------------------------------------
// Stage M — subtrees of UH_SMALL_MAX < n <= UH_MID_MAX bins: one CTA each, depth first
// ---------------------------------------------------------------------------------------------
constexpr int UH_MID_STACK = 32;

// A chromosome has only a dozen or two of these subtrees, and the big ones are chains of 100-200 dependent nodes of
// 10-16 thousand bins each (noise splits peel a few hundred bins off a node), so the stage lasts as long as its slowest
// chain and a node costs what one SM can pull out of L2 (~100 GB/s: 1.5-5 us).  STAGED: the subtree's prefix sums (at most
// UH_MID_MAX + 1 doubles = 128 KB) are copied to shared memory once and every node of the subtree scans them there.
template <int NT, int U, bool STAGED>
__global__ void __launch_bounds__(NT, NT >= 1024 ? 1 : (NT >= 512 ? 2 : 4))
uh_mid_kernel(UhParams p, int c_self) {
    extern __shared__ __align__(16) double s_pz[];  // STAGED: prefix sums of the current subtree, s_pz[0] = the one before its first bin
    __shared__ UhTask s_stack[UH_MID_STACK];
    __shared__ double s_ws[2][NT / 32], s_wv[2][NT / 32];
    __shared__ int s_wm2[2][NT / 32];
    __shared__ UhOutBuf<UhTask, 32> s_small_out;
    __shared__ UhOutBuf<UhTinyTask, 32> s_tiny_out;
    __shared__ int s_idx;
    WvCtl* ctl = p.ctl;
    UhChromCtl* cc = p.cc + c_self;
    const UhChromPlan cp = p.cp[c_self];
    const UhTask* const mid_list = p.mid + cp.mid_base;
    UhTask* const small_list = p.small + cp.small_base;
    UhTinyTask* const tiny_list = p.tiny + cp.tiny_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* __restrict__ pz = p.pz;
    unsigned long long v_mid = 0, n_mid = 0;
    UhPendingCand pend;
    pend.has = false;
    const int total = min(*(volatile int*)&cc->mid_tail_.v, cp.mid_cap);
    if (threadIdx.x == 0) { s_small_out.count = 0; s_tiny_out.count = 0; uh_stamp(p.tl_ns, c_self, 1, false); }
    int parity = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            s_idx = atomicAdd(&cc->mid_head_.v, 1);
            if (s_idx < total) s_stack[0] = mid_list[s_idx];
        }
        __syncthreads();
        if (s_idx >= total) break;
        unsigned long long dbg_t0 = 0;
        unsigned dbg_nodes = 0;
        const int dbg_bins = s_stack[0].e - s_stack[0].s + 1;
        if (p.task_dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_t0));
        const int root_s = s_stack[0].s;
        if (STAGED) {
            const double* __restrict__ src = pz + (p.off[s_stack[0].c] + s_stack[0].c) + root_s;
            for (int i = threadIdx.x; i <= dbg_bins; i += NT) s_pz[i] = __ldg(src + i);
            __syncthreads();
        }
        // every thread mirrors the stack pointer; thread 0 alone writes the stack
        int sp = 1;
        while (sp > 0) {
            sp--;
            const UhTask t = s_stack[sp];
            const int c = t.c, s = t.s, e = t.e, level = t.level;
            const double base = t.base, endv = t.endv;
            const long long p0 = p.off[c] + c;
            const long long loff = p.off[c];
            const int n = e - s + 1;
            const double nn = (double)n;
            const double T = endv - base;
            const double mu = T / nn;
            double bnum = -1.0, bden = 1.0, bv = 0.0;
            int bm = 0x7fffffff;
            if (STAGED) uh_scan<NT, U, false>(s_pz + (s - root_s) + 1, 0, n - 1, threadIdx.x, base, mu, nn, bnum, bden, bm, bv);
            else uh_scan<NT, U>(pz + p0 + s + 1, 0, n - 1, threadIdx.x, base, mu, nn, bnum, bden, bm, bv);
            double best = bnum >= 0.0 ? bnum / bden : -1.0;
            int best_m = bm;
            warp_argmax(best, best_m);
            {
                const unsigned own = __ballot_sync(0xffffffffu, bm == best_m && best_m != 0x7fffffff);
                const double wv = __shfl_sync(0xffffffffu, bv, own ? __ffs(own) - 1 : 0);
                if (lane == 0) { s_ws[parity][warp] = best; s_wm2[parity][warp] = best_m; s_wv[parity][warp] = wv; }
            }
            __syncthreads();  // everybody has read s_stack[sp]; the per-warp bests are visible
            double fbest = s_ws[parity][0], fv = s_wv[parity][0];
            int fm = s_wm2[parity][0];
#pragma unroll
            for (int w = 1; w < NT / 32; w++) {
                const double sc = s_ws[parity][w];
                const int mm = s_wm2[parity][w];
                if (sc > fbest || (sc == fbest && mm < fm)) { fbest = sc; fm = mm; fv = s_wv[parity][w]; }
            }
            parity ^= 1;
            if (threadIdx.x == 0) { v_mid += (unsigned long long)n; dbg_nodes++; }
            if (fbest == 0.0 || fm == 0x7fffffff) {
                for (int k = threadIdx.x; k < n - 1; k += NT) atomicAdd(&p.lvlcnt[loff + level + k], 1u);
                if (threadIdx.x == 0) n_mid += (unsigned long long)(n - 1);
                continue;
            }
            const int ls = s, le = s + fm, rs = s + fm + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            const int lt_ = uh_tier(ln), rt_ = uh_tier(rn);
            const bool lmid = lt_ >= UH_TIER_MID, rmid = rt_ >= UH_TIER_MID;  // a child of a mid node is never BIG
            if (threadIdx.x == 32) {
                atomicAdd(&p.lvlcnt[loff + level], 1u);
                uh_cand_defer(p, cc, cp, pend, c, level, s, s + fm, e, uh_coef_from(fv, base, T, n, fm));
                n_mid++;
            }
            if (threadIdx.x == 0) {
                const UhTask tl = {c, ls, le, level + 1, base, fv}, tr = {c, rs, re, level + 1, fv, endv};
                // larger child first so that the smaller one is popped next (stack depth <= log2)
                int q = sp;
                if (ln >= rn) { if (lmid) s_stack[q++] = tl; if (rmid) s_stack[q++] = tr; }
                else { if (rmid) s_stack[q++] = tr; if (lmid) s_stack[q++] = tl; }
                if (lt_ == UH_TIER_SMALL) uh_buf_push(s_small_out, tl, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
                if (rt_ == UH_TIER_SMALL) uh_buf_push(s_small_out, tr, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
                const UhTinyTask yl = {c, ls, le, level + 1}, yr = {c, rs, re, level + 1};
                if (lt_ == UH_TIER_TINY) uh_buf_push(s_tiny_out, yl, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
                if (rt_ == UH_TIER_TINY) uh_buf_push(s_tiny_out, yr, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
            }
            sp += (lmid ? 1 : 0) + (rmid ? 1 : 0);
            __syncthreads();  // thread 0's stack writes before the next pop
        }
        if (p.task_dbg && threadIdx.x == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            const unsigned long long i = atomicAdd(&p.task_dbg[0], 1ull);
            if (i < UH_TASK_DBG_CAP) {
                p.task_dbg[1 + 2 * i] = ((unsigned long long)c_self << 32) | (unsigned)dbg_bins;
                p.task_dbg[2 + 2 * i] = ((unsigned long long)dbg_nodes << 32) | (unsigned)min(t1 - dbg_t0, 0xffffffffull);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uh_buf_flush(s_small_out, small_list, cp.small_cap, &cc->small_tail_.v, ctl);
        uh_buf_flush(s_tiny_out, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
    }
    uh_cand_flush(p, cp, pend);
    if (v_mid) atomicAdd(&ctl->visits_big, v_mid);
    if (n_mid) atomicAdd(&ctl->nodes_big, n_mid);
    if (threadIdx.x == 0) uh_stamp(p.tl_ns, c_self, 1, true);
}

// ---------------------------------------------------------------------------------------------
// Stage S — subtrees of UH_TINY_MAX < n <= UH_SMALL_MAX bins: one warp each, depth first
// ---------------------------------------------------------------------------------------------
constexpr int UH_WARP_STACK = 16;

struct UhWarpScratch {
    UhTask st[UH_WARP_STACK];
    UhOutBuf<UhTinyTask, 32> tiny_out;
};

__global__ void __launch_bounds__(UH_SMALL_THREADS, 4)
uh_small_kernel(UhParams p, int c_self) {
    extern __shared__ __align__(16) unsigned char uh_smem[];
    UhWarpScratch& ws = reinterpret_cast<UhWarpScratch*>(uh_smem)[threadIdx.x >> 5];
    WvCtl* ctl = p.ctl;
    UhChromCtl* cc = p.cc + c_self;
    const UhChromPlan cp = p.cp[c_self];
    const UhTask* const small_list = p.small + cp.small_base;
    UhTinyTask* const tiny_list = p.tiny + cp.tiny_base;
    const double thr2 = p.cand_thr[c_self] * p.cand_thr[c_self];
    const int lane = threadIdx.x & 31;
    const int c = c_self;
    const long long p0 = p.off[c] + c;
    unsigned* const lvl = p.lvlcnt + p.off[c];  // level counts go straight to L2 (one RED per node)
    const double* __restrict__ pz = p.pz;
    unsigned long long v_small = 0, n_small = 0;
    const int total = min(*(volatile int*)&cc->small_tail_.v, cp.small_cap);
    if (lane == 0) ws.tiny_out.count = 0;
    if (threadIdx.x == 0) uh_stamp(p.tl_ns, c_self, 2, false);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&cc->small_head_.v, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= total) break;
        // (staging the subtree's prefix sums in shared memory, as the mid stage can, was measured here too: 75 KB per CTA cost
        // more occupancy than the L2 round trips it saved — profiles/rd2p_*)
        if (lane == 0) ws.st[0] = small_list[idx];
        __syncwarp();
        int sp = 1;
        while (sp > 0) {
            sp--;
            const UhTask t = ws.st[sp];
            __syncwarp();
            const int s = t.s, e = t.e, level = t.level;
            const double base = t.base, endv = t.endv;
            const int n = e - s + 1;
            const double nn = (double)n;
            const double T = endv - base;
            const double mu = T / nn;
            double bnum = -1.0, bden = 1.0, bv = 0.0;
            int bm = 0x7fffffff;
            uh_scan<32, 4>(pz + p0 + s + 1, 0, n - 1, lane, base, mu, nn, bnum, bden, bm, bv);
            double best = bnum >= 0.0 ? bnum / bden : -1.0;
            int best_m = bm;
            warp_argmax(best, best_m);
            const unsigned own = __ballot_sync(0xffffffffu, bm == best_m && best_m != 0x7fffffff);
            const double fv = __shfl_sync(0xffffffffu, bv, own ? __ffs(own) - 1 : 0);
            if (lane == 0) v_small += (unsigned long long)n;
            if (best == 0.0 || best_m == 0x7fffffff) {
                for (int k = lane; k < n - 1; k += 32) atomicAdd(&lvl[level + k], 1u);
                if (lane == 0) n_small += (unsigned long long)(n - 1);
                __syncwarp();
                continue;
            }
            const int fm = best_m;
            const int ls = s, le = s + fm, rs = s + fm + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            const int lt_ = uh_tier(ln), rt_ = uh_tier(rn);
            const bool lsm = lt_ >= UH_TIER_SMALL, rsm = rt_ >= UH_TIER_SMALL;
            // the serial part is spread over two lanes: 0 pushes, 1 counts the node and emits the (rare) candidate
            if (lane == 1) {
                atomicAdd(&lvl[level], 1u);
                if (uh_may_be_candidate(best, nn, mu, thr2)) uh_emit_candidate(p, c, level, s, s + fm, e, uh_coef_from(fv, base, T, n, fm));
                n_small++;
            }
            if (lane == 0) {
                const UhTask tl = {c, ls, le, level + 1, base, fv}, tr = {c, rs, re, level + 1, fv, endv};
                int q = sp;
                if (ln >= rn) { if (lsm) ws.st[q++] = tl; if (rsm) ws.st[q++] = tr; }
                else { if (rsm) ws.st[q++] = tr; if (lsm) ws.st[q++] = tl; }
                const UhTinyTask yl = {c, ls, le, level + 1}, yr = {c, rs, re, level + 1};
                if (lt_ == UH_TIER_TINY) uh_buf_push(ws.tiny_out, yl, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
                if (rt_ == UH_TIER_TINY) uh_buf_push(ws.tiny_out, yr, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
            }
            sp += (lsm ? 1 : 0) + (rsm ? 1 : 0);
            __syncwarp();
        }
    }
    if (lane == 0) uh_buf_flush(ws.tiny_out, tiny_list, cp.tiny_cap, &cc->tiny_tail_.v, ctl);
    if (v_small) atomicAdd(&ctl->visits_small, v_small);
    if (n_small) atomicAdd(&ctl->nodes_small, n_small);
    if (lane == 0) uh_stamp(p.tl_ns, c_self, 2, true);
}

// ---------------------------------------------------------------------------------------------
// Stage T — subtrees of <= UH_TINY_MAX bins: one thread each, the reference recurrence verbatim
// (GetInnerProdIter :19-48, GetInnerProdMax :54-67; Enumerable.Max skips NaN)
// ---------------------------------------------------------------------------------------------
// The recurrence's constants depend on (n, m) only.  They are tabulated once with the same IEEE
// operations the reference performs (sqrt and division are correctly rounded on both sides), so using
// the table instead of recomputing them per node changes no bit.
struct UhTinyTab {
    double first_plus[UH_TINY_MAX + 1];    // sqrt(1 - 1/n)
    double first_minus[UH_TINY_MAX + 1];   // 1 / sqrt(n (n - 1))
    double factor[UH_TINY_MAX + 1][UH_TINY_MAX];
    double cplus[UH_TINY_MAX + 1][UH_TINY_MAX];   // sqrt(1/(m+1) - 1/n)
    double cminus[UH_TINY_MAX + 1][UH_TINY_MAX];  // sqrt(n n/(m+1) - n)
};

__global__ void uh_tiny_table_kernel(UhTinyTab* tab) {
    const int n = blockIdx.x + 2;  // 2..UH_TINY_MAX
    const int m = threadIdx.x;
    const double nn = (double)n;
    if (m == 0) {
        tab->first_plus[n] = sqrt(__dsub_rn(1.0, __ddiv_rn(1.0, nn)));
        tab->first_minus[n] = __ddiv_rn(1.0, sqrt((double)((long long)n * (long long)(n - 1))));
    }
    if (m >= 1 && m < n - 1) {
        tab->factor[n][m] = sqrt(__ddiv_rn(__ddiv_rn(__dmul_rn((double)(n - m - 1), (double)m), (double)(m + 1)), (double)(n - m)));
        tab->cplus[n][m] = sqrt(__dsub_rn(__ddiv_rn(1.0, (double)(m + 1)), __ddiv_rn(1.0, nn)));
        tab->cminus[n][m] = sqrt(__dsub_rn(__ddiv_rn(__dmul_rn(nn, nn), (double)(m + 1)), nn));
    }
}

__global__ void __launch_bounds__(128)
uh_tiny_kernel(UhParams p, const UhTinyTab* __restrict__ gtab, int c_self) {
    __shared__ UhTinyTab s_tab;
    {
        const double* src = reinterpret_cast<const double*>(gtab);
        double* dst = reinterpret_cast<double*>(&s_tab);
        for (int i = threadIdx.x; i < (int)(sizeof(UhTinyTab) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    WvCtl* ctl = p.ctl;
    if (threadIdx.x == 0) uh_stamp(p.tl_ns, c_self, 3, false);
    const UhChromPlan cp = p.cp[c_self];
    const UhTinyTask* const tiny_list = p.tiny + cp.tiny_base;
    const int total = min(*(volatile int*)&p.cc[c_self].tiny_tail_.v, cp.tiny_cap);
    unsigned visits = 0, nodes = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const UhTinyTask root = tiny_list[idx];
        const int c = c_self;
        const double* __restrict__ xc = p.x + p.off[c];
        unsigned* __restrict__ lv = p.lvlcnt + p.off[c];
        int st_s[6], st_e[6], st_l[6];
        int sp = 1;
        st_s[0] = root.s; st_e[0] = root.e; st_l[0] = root.level;
        while (sp > 0) {
            sp--;
            const int s = st_s[sp], e = st_e[sp], level = st_l[sp];
            const int n = e - s + 1;
            double xl[UH_TINY_MAX];
#pragma unroll
            for (int i = 0; i < UH_TINY_MAX; i++) xl[i] = i < n ? xc[s + i] : 0.0;
            const double nn = (double)n;
            double sum_x = 0.0;
            for (int i = 1; i < n; i++) sum_x = __dadd_rn(sum_x, xl[i]);
            const double mean = __ddiv_rn(__dadd_rn(xl[0], sum_x), nn);
            double plus = __dmul_rn(s_tab.first_plus[n], xl[0]);
            double minus = __dmul_rn(s_tab.first_minus[n], sum_x);
            double best_ip = __dsub_rn(plus, minus);
            double best_abs = fabs(best_ip);
            if (!(best_abs >= 0.0)) best_abs = -1.0;
            int best_m = 0;
            for (int m = 1; m < n - 1; m++) {
                const double factor = s_tab.factor[n][m];
                plus = __dadd_rn(__dmul_rn(plus, factor), __dmul_rn(xl[m], s_tab.cplus[n][m]));
                minus = __dsub_rn(__ddiv_rn(minus, factor), __ddiv_rn(xl[m], s_tab.cminus[n][m]));
                const double ip = __dsub_rn(plus, minus);
                const double a = fabs(ip);
                if (a > best_abs) { best_abs = a; best_m = m; best_ip = ip; }
            }
            const double coef = __ddiv_rn(best_ip, fmax(0.5, __ddiv_rn(mean, 200.0)));
            atomicAdd(&lv[level], 1u);
            uh_emit_candidate(p, c, level, s, s + best_m, e, coef);
            visits += (unsigned)n;
            nodes++;
            // children: left [s, s+m] and right [s+m+1, e], each needs >= 2 bins; larger first
            const int ls = s, le = s + best_m, rs = s + best_m + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            if (ln >= rn) {
                if (ln >= 2) { st_s[sp] = ls; st_e[sp] = le; st_l[sp] = level + 1; sp++; }
                if (rn >= 2) { st_s[sp] = rs; st_e[sp] = re; st_l[sp] = level + 1; sp++; }
            } else {
                if (rn >= 2) { st_s[sp] = rs; st_e[sp] = re; st_l[sp] = level + 1; sp++; }
                if (ln >= 2) { st_s[sp] = ls; st_e[sp] = le; st_l[sp] = level + 1; sp++; }
            }
        }
    }
    // statistics: one atomic per warp
    visits = __reduce_add_sync(0xffffffffu, visits);
    nodes = __reduce_add_sync(0xffffffffu, nodes);
    if ((threadIdx.x & 31) == 0) {
        if (visits) atomicAdd(&ctl->visits_tiny, (unsigned long long)visits);
        if (nodes) atomicAdd(&ctl->nodes_tiny, (unsigned long long)nodes);
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(&ctl->t_last, t);
        uh_stamp(p.tl_ns, c_self, 3, true);
    }
}

// number of levels of every chromosome = highest level with a node + 1 (depth[] zeroed by the host)
__global__ void uh_depth_kernel(const unsigned* __restrict__ lvlcnt, const long long* __restrict__ off, int* __restrict__ depth, int c) {
    const long long o = off[c];
    const int n = (int)(off[c + 1] - o);
    int mx = 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x)
        if (lvlcnt[o + l]) mx = l + 1;
    mx = (int)__reduce_max_sync(0xffffffffu, (unsigned)mx);
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(&depth[c], mx);
}

// seeds: one root per selected chromosome with more than min_size bins (one thread per chromosome)
__global__ void uh_seed_kernel(UhParams p, const unsigned char* __restrict__ selected, int n_chrom, int min_size) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMin(&p.ctl->t_first, t);
    }
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chrom; c += gridDim.x * blockDim.x) {
        const long long len = p.off[c + 1] - p.off[c];
        const bool run = selected[c] && len > (long long)min_size && len >= 2;
        const int tier = run ? uh_tier((int)len) : UH_TIER_NONE;
        // a chain kernel may have been launched for this chromosome on the strength of an upper bound of its length
        // (launch sequences are cached per input shape): without a big root it has nothing to wait for
        if (tier != UH_TIER_BIG) *(volatile int*)&p.cc[c].big_done_.v = 1;
        if (!run) continue;
        const int e = (int)len - 1;
        const long long p0 = p.off[c] + c;
        const double base = p.pz[p0], endv = p.pz[p0 + e + 1];
        if (tier == UH_TIER_BIG) {
            atomicAdd(&p.cc[c].outstanding_.v, 1);
            uh_push_big(p, c, 0, e, 0, base, endv);
        } else if (tier == UH_TIER_MID) {
            const int i = atomicAdd(&p.cc[c].mid_tail_.v, 1);
            if (i < p.cp[c].mid_cap) p.mid[p.cp[c].mid_base + i] = UhTask{c, 0, e, 0, base, endv};
        } else if (tier == UH_TIER_SMALL) {
            const int i = atomicAdd(&p.cc[c].small_tail_.v, 1);
            if (i < p.cp[c].small_cap) p.small[p.cp[c].small_base + i] = UhTask{c, 0, e, 0, base, endv};
        } else if (tier == UH_TIER_TINY) {
            const int i = atomicAdd(&p.cc[c].tiny_tail_.v, 1);
            if (i < p.cp[c].tiny_cap) p.tiny[p.cp[c].tiny_base + i] = UhTinyTask{c, 0, e, 0};
        }
    }
}
