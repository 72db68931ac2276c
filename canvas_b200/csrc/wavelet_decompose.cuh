// Unbalanced-Haar top-down decomposition (WaveletSegmentation.cs:264-379) as ONE persistent kernel.
//
// The reference walks the tree level by level and, for every node, evaluates the inner products
// with all n-1 Unbalanced-Haar vectors by a sequential recurrence (GetInnerProdIter :19-48), then
// takes the first arg-max of |ip| (:54-67).  The tree is ~100-130 levels deep on WGS data because
// noise splits are lopsided, so the work is a long chain of big nodes with small subtrees peeling
// off.  Here:
//   * closed form: with P the prefix sums of the node, a = m+1, b = n-a,
//       ip[m] = (P_m - a*T/n) * sqrt(n / (a*b)),  arg-max |ip| = arg-max (P_m - a*T/n)^2 / (a*b)
//     so a node costs one division per split point on the chromosome-wide prefix-sum array;
//   * big nodes (n > UH_SMALL_MAX) are walked by "chain worker" CTAs: a CTA takes a big node, finds
//     its split with all its threads (16 loads in flight per thread on the L2-resident prefix sums,
//     one __syncthreads per node), emits it, keeps going with the LARGER big child and hands the other
//     big child to an idle CTA through a task ring — no level barrier, no per-node global handshake:
//     the prefix values bounding each child are carried over from the arg-max winner, so the chain
//     never re-reads them;
//   * a node of <= UH_SMALL_MAX bins is handed to ONE warp that runs its whole subtree depth-first
//     (smaller child first: stack depth <= log2 n) without touching the global queue;
//   * nodes of <= UH_TINY_MAX bins are batched 32 at a time and done one per thread with the
//     reference's own recurrence, operation for operation, so exact ties in tiny nodes (symmetric
//     patterns on 2-decimal data) break exactly as in the reference.
// Nodes are not stored: only per-level node counts (HardThresh's germline weights need them) and
// the few nodes whose coefficient can survive the threshold ("candidates").
#pragma once
#include "wavelet.cuh"

struct UhParams {
    const double* x;         // coverage [N]
    const double* pz;        // prefix sums with leading zero per chromosome [N + n_chrom]
    const long long* off;    // [n_chrom + 1]
    const double* cand_thr;  // [n_chrom]
    unsigned* lvlcnt;        // [N]: node count of level l of chromosome c at off[c] + l
    int* depth;              // [n_chrom]: number of levels
    UhBigTask* big;               // ring [UH_QCAP]; c < 0 = empty
    UhSmallTask* small;            // c < 0 = not published yet
    int small_cap;
    UhCand* cand;
    int cand_cap;
    WvCtl* ctl;
};

__device__ inline void uh_emit_candidate(const UhParams& p, int c, int level, int s, int b, int e, double coef) {
    if (fabs(coef) <= p.cand_thr[c]) return;  // NaN falls through on purpose (never zeroed by HardThresh)
    const int i = atomicAdd(&p.ctl->cand_count_.v, 1);
    if (i >= p.cand_cap) { p.ctl->overflow_.v = 1; return; }
    UhCand k;
    k.key = ((unsigned long long)c << 56) | ((unsigned long long)level << 32) | (unsigned)s;
    k.s = s; k.b = b; k.e = e; k.level = level; k.c = c; k.pad = 0; k.coef = coef;
    p.cand[i] = k;
}

__device__ inline void uh_push_small(const UhParams& p, int c, int s, int e, int level) {
    const int i = atomicAdd(&p.ctl->small_tail_.v, 1);
    if (i >= p.small_cap) { p.ctl->overflow_.v = 1; return; }
    UhSmallTask* t = p.small + i;
    t->s = s; t->e = e; t->level = level;
    __threadfence();
    *(volatile int*)&t->c = c;  // publish
}

// Big-task ring: a producer reserves a position with one atomicAdd, fills the slot and publishes it
// by storing the chromosome id last; a consumer reserves a position the same way and waits for it.
__device__ inline void uh_push_big(const UhParams& p, int c, int s, int e, int level, double base, double endv) {
    const unsigned long long pos = atomicAdd(&p.ctl->q_tail_.v, 1ull);
    UhBigTask* t = p.big + (pos & (UH_QCAP - 1));
    while (*(volatile int*)&t->c >= 0) __nanosleep(64);  // slot still held by an unconsumed task (ring wrapped)
    t->s = s; t->e = e; t->level = level; t->base = base; t->endv = endv;
    __threadfence();
    *(volatile int*)&t->c = c;
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

// inner product and coefficient of the chosen split (closed form of :36-45, scaling of :282-283)
__device__ inline double uh_coef(const double* __restrict__ pz, long long p0, int s, int n, int m, double base, double T) {
    const double a = (double)(m + 1), b = (double)(n - m - 1), nn = (double)n;
    const double mu = T / nn;
    const double D = (pz[p0 + s + m + 1] - base) - a * mu;
    const double ip = D * sqrt(nn / (a * b));
    return ip / fmax(0.5, mu / 200.0);
}

// ---------------------------------------------------------------------------------------------
// Tiny subtree (n <= UH_TINY_MAX), one thread, the reference recurrence verbatim.
// ---------------------------------------------------------------------------------------------
__device__ void uh_tiny_subtree(const UhParams& p, int c, int s0, int e0, int level0, unsigned* s_lvl, int lvl_base,
                                unsigned long long& visits, unsigned long long& nodes) {
    const double* __restrict__ xc = p.x + p.off[c];
    int st_s[6], st_e[6], st_l[6];
    int sp = 0;
    st_s[0] = s0; st_e[0] = e0; st_l[0] = level0; sp = 1;
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
        double plus = __dmul_rn(sqrt(__dsub_rn(1.0, __ddiv_rn(1.0, nn))), xl[0]);
        double minus = __dmul_rn(__ddiv_rn(1.0, sqrt((double)((long long)n * (long long)(n - 1)))), sum_x);
        double best_ip = __dsub_rn(plus, minus);
        double best_abs = fabs(best_ip);
        if (!(best_abs >= 0.0)) best_abs = -1.0;
        int best_m = 0;
        for (int m = 1; m < n - 1; m++) {
            const double factor = sqrt(__ddiv_rn(__ddiv_rn(__dmul_rn((double)(n - m - 1), (double)m), (double)(m + 1)), (double)(n - m)));
            plus = __dadd_rn(__dmul_rn(plus, factor),
                             __dmul_rn(xl[m], sqrt(__dsub_rn(__ddiv_rn(1.0, (double)(m + 1)), __ddiv_rn(1.0, nn)))));
            minus = __dsub_rn(__ddiv_rn(minus, factor),
                              __ddiv_rn(xl[m], sqrt(__dsub_rn(__ddiv_rn(__dmul_rn(nn, nn), (double)(m + 1)), nn))));
            const double ip = __dsub_rn(plus, minus);
            const double a = fabs(ip);
            if (a > best_abs) { best_abs = a; best_m = m; best_ip = ip; }
        }
        const double coef = __ddiv_rn(best_ip, fmax(0.5, __ddiv_rn(mean, 200.0)));
        atomicAdd(&s_lvl[level - lvl_base], 1u);
        uh_emit_candidate(p, c, level, s, s + best_m, e, coef);
        visits += (unsigned long long)n;
        nodes++;
        // children: left [s, s+m] needs >= 2 bins, right [s+m+1, e] needs >= 2 bins; larger first
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

// ---------------------------------------------------------------------------------------------
// Small subtree (n <= UH_SMALL_MAX): one warp, depth first.
// ---------------------------------------------------------------------------------------------
constexpr int UH_WARP_STACK = 16;
constexpr int UH_TINY_BUF = 32;

struct UhWarpScratch {
    unsigned lvl[UH_SMALL_MAX];  // node count per level relative to the task's level
    int st_s[UH_WARP_STACK], st_e[UH_WARP_STACK], st_l[UH_WARP_STACK];
    double st_b[UH_WARP_STACK], st_v[UH_WARP_STACK];  // prefix sums bounding each stacked node
    int tn_s[UH_TINY_BUF], tn_e[UH_TINY_BUF], tn_l[UH_TINY_BUF];
};

__device__ void uh_small_subtree(const UhParams& p, UhWarpScratch& ws, int c, int S, int E, int L0,
                                 unsigned long long& visits_small, unsigned long long& visits_tiny,
                                 unsigned long long& nodes_small, unsigned long long& nodes_tiny) {
    const int lane = threadIdx.x & 31;
    const long long p0 = p.off[c] + c;
    const double* __restrict__ pz = p.pz;
    for (int t = lane; t < UH_SMALL_MAX; t += 32) ws.lvl[t] = 0u;
    int sp = 0, ntiny = 0;
    if (lane == 0) {
        ws.st_s[0] = S; ws.st_e[0] = E; ws.st_l[0] = L0;
        ws.st_b[0] = pz[p0 + S]; ws.st_v[0] = pz[p0 + E + 1];
    }
    sp = 1;
    __syncwarp();
    while (sp > 0 || ntiny > 0) {
        if (sp == 0 || ntiny == UH_TINY_BUF) {
            // flush the tiny batch: one subtree per lane
            if (lane < ntiny) uh_tiny_subtree(p, c, ws.tn_s[lane], ws.tn_e[lane], ws.tn_l[lane], ws.lvl, L0, visits_tiny, nodes_tiny);
            __syncwarp();
            ntiny = 0;
            continue;
        }
        sp--;
        const int s = ws.st_s[sp], e = ws.st_e[sp], level = ws.st_l[sp];
        const double base = ws.st_b[sp], endv = ws.st_v[sp];
        __syncwarp();
        const int n = e - s + 1;
        if (n <= UH_TINY_MAX) {
            if (lane == 0) { ws.tn_s[ntiny] = s; ws.tn_e[ntiny] = e; ws.tn_l[ntiny] = level; }
            ntiny++;
            __syncwarp();
            continue;
        }
        const double T = endv - base;
        const double nn = (double)n;
        const double mu = T / nn;
        const double* __restrict__ q = pz + p0 + s + 1;
        double bnum = -1.0, bden = 1.0, bv = 0.0;
        int bm = 0x7fffffff;
        int m = lane;
        for (; m + 96 < n - 1; m += 128) {  // four loads in flight per lane
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = __ldg(q + m + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double a = (double)(m + 32 * u + 1);
                const double D = fma(-a, mu, v[u] - base);
                const double num = D * D, den = a * (nn - a);
                if (num * bden > bnum * den) { bnum = num; bden = den; bm = m + 32 * u; bv = v[u]; }
            }
        }
        for (; m < n - 1; m += 32) {
            const double vv = __ldg(q + m);
            const double a = (double)(m + 1);
            const double D = fma(-a, mu, vv - base);
            const double num = D * D, den = a * (nn - a);
            if (num * bden > bnum * den) { bnum = num; bden = den; bm = m; bv = vv; }
        }
        double best = bnum >= 0.0 ? bnum / bden : -1.0;
        int best_m = bm;
        warp_argmax(best, best_m);
        const unsigned own = __ballot_sync(0xffffffffu, bm == best_m && best_m != 0x7fffffff);
        const double fv = __shfl_sync(0xffffffffu, bv, own ? __ffs(own) - 1 : 0);
        if (lane == 0) { visits_small += (unsigned long long)n; }
        if (best == 0.0 || best_m == 0x7fffffff) {
            // every inner product is exactly zero (a run of zeros): the reference peels one bin per
            // level with coefficient 0 — levels level .. level+n-2 get one node each
            for (int k = lane; k < n - 1; k += 32) atomicAdd(&ws.lvl[level - L0 + k], 1u);
            if (lane == 0) nodes_small += (unsigned long long)(n - 1);
            __syncwarp();
            continue;
        }
        if (best_m < 0 || best_m > n - 2) best_m = 0;
        if (lane == 0) {
            const double a = (double)(best_m + 1), b = (double)(n - best_m - 1);
            const double D = fma(-a, mu, fv - base);
            const double coef = D * sqrt(nn / (a * b)) / fmax(0.5, mu / 200.0);
            atomicAdd(&ws.lvl[level - L0], 1u);
            uh_emit_candidate(p, c, level, s, s + best_m, e, coef);
            nodes_small++;
            const int ls = s, le = s + best_m, rs = s + best_m + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            int qn = sp;
            // larger child first so that the smaller one is popped next (stack depth <= log2 n)
            if (ln >= rn) {
                if (ln >= 2) { ws.st_s[qn] = ls; ws.st_e[qn] = le; ws.st_l[qn] = level + 1; ws.st_b[qn] = base; ws.st_v[qn] = fv; qn++; }
                if (rn >= 2) { ws.st_s[qn] = rs; ws.st_e[qn] = re; ws.st_l[qn] = level + 1; ws.st_b[qn] = fv; ws.st_v[qn] = endv; qn++; }
            } else {
                if (rn >= 2) { ws.st_s[qn] = rs; ws.st_e[qn] = re; ws.st_l[qn] = level + 1; ws.st_b[qn] = fv; ws.st_v[qn] = endv; qn++; }
                if (ln >= 2) { ws.st_s[qn] = ls; ws.st_e[qn] = le; ws.st_l[qn] = level + 1; ws.st_b[qn] = base; ws.st_v[qn] = fv; qn++; }
            }
        }
        {
            const int ln = best_m + 1, rn = n - best_m - 1;
            sp += (ln >= 2) + (rn >= 2);
        }
        __syncwarp();
    }
    // flush level counts
    int maxrel = -1;
    for (int t = lane; t < UH_SMALL_MAX; t += 32) {
        const unsigned v = ws.lvl[t];
        if (v) { atomicAdd(&p.lvlcnt[p.off[c] + L0 + t], v); maxrel = t; }
    }
    maxrel = (int)__reduce_max_sync(0xffffffffu, (unsigned)(maxrel + 1));
    if (lane == 0 && maxrel > 0) atomicMax(&p.depth[c], L0 + maxrel);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// The persistent kernel.  blockIdx % 4 == 0: big worker (whole CTA per chunk ticket) until the big
// phase ends, then joins the others; the rest: 8 independent warp workers on small subtrees.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UH_THREADS, 2)
uh_decompose_kernel(UhParams p) {
    extern __shared__ __align__(16) unsigned char uh_smem[];
    UhWarpScratch* s_ws = reinterpret_cast<UhWarpScratch*>(uh_smem);
    WvCtl* ctl = p.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMin(&ctl->t_first, t);
    }
    unsigned long long v_big = 0, v_small = 0, v_tiny = 0, n_big = 0, n_small = 0, n_tiny = 0;

    if ((blockIdx.x & 1) == 0) {
        // ----------------------------------------------------------------- chain worker
        __shared__ int s_task[4];
        __shared__ double s_taskd[2];
        __shared__ double s_ws[2][UH_THREADS / 32];   // per-warp best score, double-buffered by node parity
        __shared__ double s_wv[2][UH_THREADS / 32];   // prefix value at the best split
        __shared__ int s_wm2[2][UH_THREADS / 32];
        const double* __restrict__ pz = p.pz;
        int parity = 0;
        for (;;) {
            // ---- take a big task
            if (threadIdx.x == 0) {
                int c = -1;
                const unsigned long long pos = atomicAdd(&ctl->q_head_.v, 1ull);
                UhBigTask* t = p.big + (pos & (UH_QCAP - 1));
                volatile int* ready = &t->c;
                unsigned polls = 0;
                for (;;) {
                    c = *ready;
                    if (c >= 0) break;
                    if ((++polls & 3u) == 0u && (*(volatile int*)&ctl->big_done_.v || *(volatile int*)&ctl->overflow_.v)) {
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
            __syncthreads();
            const int c = s_task[0];
            int s = s_task[1], e = s_task[2], level = s_task[3];
            double base = s_taskd[0], endv = s_taskd[1];
            __syncthreads();
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
                // arg-max of D^2 / (a b); inside a thread compared by cross-multiplication (no division)
                double bnum = -1.0, bden = 1.0, bv = 0.0;
                int bm = 0x7fffffff;
                int m = threadIdx.x;
                for (; m + 15 * UH_THREADS < n - 1; m += 16 * UH_THREADS) {
                    double v[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) v[u] = __ldg(q + m + u * UH_THREADS);
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const double a = (double)(m + u * UH_THREADS + 1);
                        const double D = fma(-a, mu, v[u] - base);
                        const double num = D * D, den = a * (nn - a);
                        if (num * bden > bnum * den) { bnum = num; bden = den; bm = m + u * UH_THREADS; bv = v[u]; }
                    }
                }
                for (; m + 3 * UH_THREADS < n - 1; m += 4 * UH_THREADS) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = __ldg(q + m + u * UH_THREADS);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double a = (double)(m + u * UH_THREADS + 1);
                        const double D = fma(-a, mu, v[u] - base);
                        const double num = D * D, den = a * (nn - a);
                        if (num * bden > bnum * den) { bnum = num; bden = den; bm = m + u * UH_THREADS; bv = v[u]; }
                    }
                }
                for (; m < n - 1; m += UH_THREADS) {
                    const double vv = __ldg(q + m);
                    const double a = (double)(m + 1);
                    const double D = fma(-a, mu, vv - base);
                    const double num = D * D, den = a * (nn - a);
                    if (num * bden > bnum * den) { bnum = num; bden = den; bm = m; bv = vv; }
                }
                double best = bnum >= 0.0 ? bnum / bden : -1.0;
                int best_m = bm;
                warp_argmax(best, best_m);
                {
                    // the lane that owns the winning split also owns its prefix value
                    const unsigned own = __ballot_sync(0xffffffffu, bm == best_m && best_m != 0x7fffffff);
                    const int src = own ? __ffs(own) - 1 : 0;
                    const double wv = __shfl_sync(0xffffffffu, bv, src);
                    if (lane == 0) { s_ws[parity][warp] = best; s_wm2[parity][warp] = best_m; s_wv[parity][warp] = wv; }
                }
                __syncthreads();
                double fbest = s_ws[parity][0], fv = s_wv[parity][0];
                int fm = s_wm2[parity][0];
#pragma unroll
                for (int w = 1; w < UH_THREADS / 32; w++) {
                    const double sc = s_ws[parity][w];
                    const int mm = s_wm2[parity][w];
                    if (sc > fbest || (sc == fbest && mm < fm)) { fbest = sc; fm = mm; fv = s_wv[parity][w]; }
                }
                parity ^= 1;
                if (threadIdx.x == 0) v_big += (unsigned long long)n;
                if (fbest == 0.0 || fm == 0x7fffffff) {
                    // run of exact zeros: comb of n-1 nodes with coefficient 0 (see uh_small_subtree)
                    for (int k = threadIdx.x; k < n - 1; k += UH_THREADS) atomicAdd(&p.lvlcnt[loff + level + k], 1u);
                    if (threadIdx.x == 0) { atomicMax(&p.depth[c], level + n - 1); n_big += (unsigned long long)(n - 1); }
                    break;  // chain ends
                }
                // ---- emit (one lane of warp 1; every thread derives the children itself)
                if (threadIdx.x == 32) {
                    const double a = (double)(fm + 1), b = (double)(n - fm - 1);
                    const double D = fma(-a, mu, fv - base);
                    const double ip = D * sqrt(nn / (a * b));
                    const double coef = ip / fmax(0.5, mu / 200.0);
                    atomicAdd(&p.lvlcnt[loff + level], 1u);
                    atomicMax(&p.depth[c], level + 1);
                    uh_emit_candidate(p, c, level, s, s + fm, e, coef);
                    n_big++;
                }
                const int ls = s, le = s + fm, rs = s + fm + 1, re = e;
                const int ln = le - ls + 1, rn = re - rs + 1;
                const bool lbig = ln > UH_SMALL_MAX, rbig = rn > UH_SMALL_MAX;
                // continue with the larger big child (left on ties)
                const bool cont_left = lbig && (!rbig || ln >= rn);
                const bool cont_right = rbig && !cont_left;
                if (threadIdx.x == 0) {
                    if (lbig && !cont_left) { atomicAdd(&ctl->outstanding_.v, 1); uh_push_big(p, c, ls, le, level + 1, base, fv); }
                    if (rbig && !cont_right) { atomicAdd(&ctl->outstanding_.v, 1); uh_push_big(p, c, rs, re, level + 1, fv, endv); }
                    if (!lbig && ln >= 2) uh_push_small(p, c, ls, le, level + 1);
                    if (!rbig && rn >= 2) uh_push_small(p, c, rs, re, level + 1);
                }
                if (cont_left) { e = le; endv = fv; level++; }
                else if (cont_right) { s = rs; base = fv; level++; }
                else break;  // no big child: chain ends
            }
            // ---- chain finished: one big task less
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                const int now = atomicSub(&ctl->outstanding_.v, 1) - 1;
                if (now == 0) {
                    __threadfence();
                    *(volatile int*)&ctl->big_done_.v = 1;
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                    ctl->t_big_done = t;
                }
            }
        }
        __syncthreads();
    }

    // --------------------------------------------------------------------- small workers (per warp)
    UhWarpScratch& ws = s_ws[warp];
    for (;;) {
        int idx = 0, c = -1, s = 0, e = 0, level = 0;
        if (lane == 0) {
            idx = atomicAdd(&ctl->small_head_.v, 1);
            if (idx < p.small_cap) {
                volatile int* ready = &p.small[idx].c;
                unsigned backoff = 100, polls = 0;
                for (;;) {
                    c = *ready;
                    if (c >= 0) break;
                    // the shared flags are looked at every 4th poll only (one line for all idle warps)
                    if ((++polls & 3u) == 0u && (*(volatile int*)&ctl->big_done_.v || *(volatile int*)&ctl->overflow_.v)) {
                        __threadfence();
                        c = *ready;  // every push happened before big_done was raised
                        break;
                    }
                    __nanosleep(backoff);
                    if (backoff < 1600) backoff <<= 1;
                }
                if (c >= 0) {
                    __threadfence();
                    s = *(volatile int*)&p.small[idx].s;
                    e = *(volatile int*)&p.small[idx].e;
                    level = *(volatile int*)&p.small[idx].level;
                }
            }
        }
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c < 0) break;
        s = __shfl_sync(0xffffffffu, s, 0);
        e = __shfl_sync(0xffffffffu, e, 0);
        level = __shfl_sync(0xffffffffu, level, 0);
        uh_small_subtree(p, ws, c, s, e, level, v_small, v_tiny, n_small, n_tiny);
    }
    // statistics
    if (v_big) atomicAdd(&ctl->visits_big, v_big);
    if (v_small) atomicAdd(&ctl->visits_small, v_small);
    if (v_tiny) atomicAdd(&ctl->visits_tiny, v_tiny);
    if (n_big) atomicAdd(&ctl->nodes_big, n_big);
    if (n_small) atomicAdd(&ctl->nodes_small, n_small);
    if (n_tiny) atomicAdd(&ctl->nodes_tiny, n_tiny);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(&ctl->t_last, t);
    }
}

// seeds: one root per selected chromosome with more than min_size bins (one thread per chromosome)
__global__ void uh_seed_kernel(UhParams p, const unsigned char* __restrict__ selected, int n_chrom, int min_size) {
    __shared__ int s_big;
    if (threadIdx.x == 0) s_big = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) {
        const long long len = p.off[c + 1] - p.off[c];
        if (!selected[c] || len <= (long long)min_size || len < 2) continue;
        const int e = (int)len - 1;
        if (e + 1 > UH_SMALL_MAX) {
            atomicAdd(&s_big, 1);
            atomicAdd(&p.ctl->outstanding_.v, 1);
            const long long p0 = p.off[c] + c;
            uh_push_big(p, c, 0, e, 0, p.pz[p0], p.pz[p0 + e + 1]);
        } else {
            uh_push_small(p, c, 0, e, 0);
        }
    }
    __syncthreads();
    // no big node at all: the big phase is over before it starts
    if (threadIdx.x == 0 && s_big == 0) { __threadfence(); *(volatile int*)&p.ctl->big_done_.v = 1; }
}
