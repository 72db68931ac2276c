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
//   * big nodes (n > UH_SMALL_MAX) are cut into chunk tickets served from a global ring by
//     "big worker" CTAs; the last chunk to finish reduces, emits the node and enqueues its children
//     — no level barrier, every chromosome's chain advances on its own;
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
    UhNode* bn;
    int bn_cap;
    unsigned long long* tickets;  // ring [UH_QCAP]; 0 = empty
    double* part_score;           // ring [UH_QCAP]: per-ticket partial arg-max of a multi-chunk node
    int* part_m;                  // ring [UH_QCAP]
    UhSmallTask* small;            // c < 0 = not published yet
    int small_cap;
    UhCand* cand;
    int cand_cap;
    WvCtl* ctl;
};

__device__ inline void uh_emit_candidate(const UhParams& p, int c, int level, int s, int b, int e, double coef) {
    if (fabs(coef) <= p.cand_thr[c]) return;  // NaN falls through on purpose (never zeroed by HardThresh)
    const int i = atomicAdd(&p.ctl->cand_count, 1);
    if (i >= p.cand_cap) { p.ctl->overflow = 1; return; }
    UhCand k;
    k.key = ((unsigned long long)c << 56) | ((unsigned long long)level << 32) | (unsigned)s;
    k.s = s; k.b = b; k.e = e; k.level = level; k.c = c; k.pad = 0; k.coef = coef;
    p.cand[i] = k;
}

__device__ inline void uh_push_small(const UhParams& p, int c, int s, int e, int level) {
    const int i = atomicAdd(&p.ctl->small_tail, 1);
    if (i >= p.small_cap) { p.ctl->overflow = 1; return; }
    UhSmallTask* t = p.small + i;
    t->s = s; t->e = e; t->level = level;
    __threadfence();
    *(volatile int*)&t->c = c;  // publish
}

// A big node owns `nch` consecutive positions of the ticket ring: ticket i of the node sits at
// position pos + i, its partial result at the same ring index, and the node record itself in slot
// pos & mask.  One atomicAdd on q_tail therefore allocates all three.  `tid`/`nthr` let a whole CTA
// write the tickets in parallel (the stores are independent); a single thread passes (0, 1).
// The caller must make the node record visible (threadfence) before tickets are written; this
// routine does that for the single-thread form.
__device__ inline unsigned long long uh_big_alloc(const UhParams& p, int c, int s, int e, int level) {
    const int nsplit = e - s;  // n - 1 split positions
    const int nch = (nsplit + UH_CHUNK - 1) / UH_CHUNK;
    const unsigned long long pos = atomicAdd(&p.ctl->q_tail, (unsigned long long)nch);
    UhNode* nd = p.bn + (pos & (UH_QCAP - 1));
    nd->c = c; nd->s = s; nd->e = e; nd->level = level;
    nd->nchunks = nch; nd->done = 0; nd->pos = pos;
    __threadfence();
    return pos;
}

__device__ inline void uh_big_publish(const UhParams& p, unsigned long long pos, int nch, int tid, int nthr) {
    const unsigned long long slot1 = (pos & (UH_QCAP - 1)) + 1ull;
    for (int i = tid; i < nch; i += nthr) {
        volatile unsigned long long* slot = p.tickets + ((pos + i) & (UH_QCAP - 1));
        while (*slot != 0ull) __nanosleep(64);
        *slot = (slot1 << 32) | (unsigned)i;
    }
}

__device__ inline int uh_nchunks(int s, int e) { return (e - s + UH_CHUNK - 1) / UH_CHUNK; }

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
    int tn_s[UH_TINY_BUF], tn_e[UH_TINY_BUF], tn_l[UH_TINY_BUF];
};

__device__ void uh_small_subtree(const UhParams& p, UhWarpScratch& ws, int c, int S, int E, int L0,
                                 unsigned long long& visits_small, unsigned long long& visits_tiny,
                                 unsigned long long& nodes_small, unsigned long long& nodes_tiny) {
    const int lane = threadIdx.x & 31;
    const long long p0 = p.off[c] + c;
    const double* __restrict__ pz = p.pz;
    for (int t = lane; t < UH_SMALL_MAX; t += 32) ws.lvl[t] = 0u;
    __syncwarp();
    int sp = 0, ntiny = 0;
    if (lane == 0) { ws.st_s[0] = S; ws.st_e[0] = E; ws.st_l[0] = L0; }
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
        __syncwarp();
        const int n = e - s + 1;
        if (n <= UH_TINY_MAX) {
            if (lane == 0) { ws.tn_s[ntiny] = s; ws.tn_e[ntiny] = e; ws.tn_l[ntiny] = level; }
            ntiny++;
            __syncwarp();
            continue;
        }
        const double base = pz[p0 + s];
        const double T = pz[p0 + e + 1] - base;
        const double nn = (double)n;
        const double mu = T / nn;
        double best = -1.0;
        int best_m = 0x7fffffff;
        for (int m = lane; m < n - 1; m += 32) {
            const double a = (double)(m + 1);
            const double D = (pz[p0 + s + m + 1] - base) - a * mu;
            const double sc = D * D / (a * (nn - a));
            if (sc > best) { best = sc; best_m = m; }
        }
        warp_argmax(best, best_m);
        if (lane == 0) { visits_small += (unsigned long long)n; }
        if (best == 0.0) {
            // every inner product is exactly zero (a run of zeros): the reference peels one bin per
            // level with coefficient 0 — levels level .. level+n-2 get one node each
            for (int k = lane; k < n - 1; k += 32) atomicAdd(&ws.lvl[level - L0 + k], 1u);
            if (lane == 0) nodes_small += (unsigned long long)(n - 1);
            __syncwarp();
            continue;
        }
        if (best_m == 0x7fffffff || best_m < 0 || best_m > n - 2) best_m = 0;
        if (lane == 0) {
            const double coef = uh_coef(pz, p0, s, n, best_m, base, T);
            atomicAdd(&ws.lvl[level - L0], 1u);
            uh_emit_candidate(p, c, level, s, s + best_m, e, coef);
            nodes_small++;
            const int ls = s, le = s + best_m, rs = s + best_m + 1, re = e;
            const int ln = le - ls + 1, rn = re - rs + 1;
            int q = sp;
            if (ln >= rn) {
                if (ln >= 2) { ws.st_s[q] = ls; ws.st_e[q] = le; ws.st_l[q] = level + 1; q++; }
                if (rn >= 2) { ws.st_s[q] = rs; ws.st_e[q] = re; ws.st_l[q] = level + 1; q++; }
            } else {
                if (rn >= 2) { ws.st_s[q] = rs; ws.st_e[q] = re; ws.st_l[q] = level + 1; q++; }
                if (ln >= 2) { ws.st_s[q] = ls; ws.st_e[q] = le; ws.st_l[q] = level + 1; q++; }
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
    __shared__ unsigned long long s_ticket;
    __shared__ double s_wscore[UH_THREADS / 32];
    __shared__ int s_wm[UH_THREADS / 32];
    __shared__ int s_flag;
    WvCtl* ctl = p.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long v_big = 0, v_small = 0, v_tiny = 0, n_big = 0, n_small = 0, n_tiny = 0;

    if ((blockIdx.x & 1) == 0) {
        // ----------------------------------------------------------------- big worker
        __shared__ int s_child[2][4];               // routed children {s, e, level, kind}
        __shared__ unsigned long long s_child_pos[2];
        bool have_local = false;                    // a single-chunk child continued without the queue
        int lc = 0, ls = 0, le = 0, ll = 0;
        for (;;) {
            int c, s, e, level, nchunks, chunk;
            unsigned long long npos = 0;
            UhNode* nd = nullptr;
            if (have_local) {
                c = lc; s = ls; e = le; level = ll; nchunks = 1; chunk = 0;
                have_local = false;
            } else {
                if (threadIdx.x == 0) {
                    unsigned long long t = 0;
                    const unsigned long long pos = atomicAdd(&ctl->q_head, 1ull);
                    volatile unsigned long long* slot = p.tickets + (pos & (UH_QCAP - 1));
                    for (;;) {
                        t = *slot;
                        if (t != 0ull) { *slot = 0ull; break; }
                        if (*(volatile int*)&ctl->big_done || *(volatile int*)&ctl->overflow) break;
                        __nanosleep(40);
                    }
                    s_ticket = t;
                }
                __syncthreads();
                const unsigned long long t = s_ticket;
                __syncthreads();
                if (t == 0ull) break;
                nd = p.bn + ((t >> 32) - 1ull);
                chunk = (int)(t & 0xffffffffu);
                c = __ldcg(&nd->c); s = __ldcg(&nd->s); e = __ldcg(&nd->e); level = __ldcg(&nd->level);
                nchunks = __ldcg(&nd->nchunks);
                npos = __ldcg(&nd->pos);
            }
            const int n = e - s + 1;
            const long long p0 = p.off[c] + c;
            const double* __restrict__ pz = p.pz;
            const double base = pz[p0 + s];
            const double T = pz[p0 + e + 1] - base;
            const double nn = (double)n;
            const double mu = T / nn;
            const int m0 = chunk * UH_CHUNK;
            const int m1 = min(m0 + UH_CHUNK, n - 1);
            double best = -1.0;
            int best_m = 0x7fffffff;
            {
                const double* __restrict__ q = pz + p0 + s + 1;
                int m = m0 + threadIdx.x;
                // eight independent loads in flight per thread (the prefix sums are L2-resident; a chunk
                // is two round trips for the CTA)
                for (; m + 7 * UH_THREADS < m1; m += 8 * UH_THREADS) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = q[m + u * UH_THREADS];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const double a = (double)(m + u * UH_THREADS + 1);
                        const double D = (v[u] - base) - a * mu;
                        const double sc = D * D / (a * (nn - a));
                        if (sc > best) { best = sc; best_m = m + u * UH_THREADS; }
                    }
                }
                for (; m < m1; m += UH_THREADS) {
                    const double a = (double)(m + 1);
                    const double D = (q[m] - base) - a * mu;
                    const double sc = D * D / (a * (nn - a));
                    if (sc > best) { best = sc; best_m = m; }
                }
            }
            warp_argmax(best, best_m);
            if (lane == 0) { s_wscore[warp] = best; s_wm[warp] = best_m; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < UH_THREADS / 32; w++)
                    if (s_wscore[w] > best || (s_wscore[w] == best && s_wm[w] < best_m)) { best = s_wscore[w]; best_m = s_wm[w]; }
                v_big += (unsigned long long)(m1 - m0);
                int fin = 1;
                if (nchunks > 1) {
                    const unsigned long long ri = (npos + (unsigned)chunk) & (UH_QCAP - 1);
                    *(volatile double*)&p.part_score[ri] = best;
                    *(volatile int*)&p.part_m[ri] = best_m;
                    __threadfence();
                    fin = (atomicAdd(&nd->done, 1) == nchunks - 1) ? 1 : 0;
                }
                s_flag = fin;
                s_wscore[0] = best;
                s_wm[0] = best_m;
            }
            __syncthreads();
            const int finalize = s_flag;
            if (!finalize) { __syncthreads(); continue; }
            // ---- last chunk: reduce the partials of a multi-chunk node (one ring read per thread)
            double fbest = s_wscore[0];
            int fm = s_wm[0];
            __syncthreads();
            if (nchunks > 1) {
                __threadfence();
                double ps = -1.0;
                int pm = 0x7fffffff;
                for (int i = threadIdx.x; i < nchunks; i += UH_THREADS) {
                    const unsigned long long ri = (npos + (unsigned)i) & (UH_QCAP - 1);
                    const double sc = __ldcg(&p.part_score[ri]);
                    const int mm = __ldcg(&p.part_m[ri]);
                    if (sc > ps || (sc == ps && mm < pm)) { ps = sc; pm = mm; }
                }
                warp_argmax(ps, pm);
                if (lane == 0) { s_wscore[warp] = ps; s_wm[warp] = pm; }
                __syncthreads();
                fbest = s_wscore[0]; fm = s_wm[0];
                for (int w = 1; w < UH_THREADS / 32; w++)
                    if (s_wscore[w] > fbest || (s_wscore[w] == fbest && s_wm[w] < fm)) { fbest = s_wscore[w]; fm = s_wm[w]; }
                __syncthreads();
            }
            int nbig_children = 0;
            if (fbest == 0.0) {
                // run of exact zeros: comb of n-1 nodes with coefficient 0 (see uh_small_subtree)
                for (int k = threadIdx.x; k < n - 1; k += UH_THREADS) atomicAdd(&p.lvlcnt[p.off[c] + level + k], 1u);
                if (threadIdx.x == 0) {
                    atomicMax(&p.depth[c], level + n - 1);
                    n_big += (unsigned long long)(n - 1);
                    s_child[0][3] = 0; s_child[1][3] = 0;
                }
                __syncthreads();
            } else {
                if (fm == 0x7fffffff || fm < 0 || fm > n - 2) fm = 0;
                // children: kind 0 none, 1 small, 2 big through the queue, 3 big continued locally
                if (threadIdx.x < 2) {
                    const int k = threadIdx.x;
                    const int cs = k == 0 ? s : s + fm + 1, ce = k == 0 ? s + fm : e;
                    const int cn = ce - cs + 1;
                    int kind = 0;
                    if (cn >= 2) {
                        if (cn <= UH_SMALL_MAX) { uh_push_small(p, c, cs, ce, level + 1); kind = 1; }
                        else {
                            const int other_n = n - cn;
                            const bool single = uh_nchunks(cs, ce) == 1;
                            // continue locally with a single-chunk child; if both qualify, the larger (left on ties)
                            const bool other_single_big = other_n > UH_SMALL_MAX && other_n - 1 <= UH_CHUNK;
                            const bool prefer = !other_single_big || cn > other_n || (cn == other_n && k == 0);
                            if (single && prefer) kind = 3;
                            else { s_child_pos[k] = uh_big_alloc(p, c, cs, ce, level + 1); kind = 2; }
                        }
                    }
                    s_child[k][0] = cs; s_child[k][1] = ce; s_child[k][2] = level + 1; s_child[k][3] = kind;
                }
                if (threadIdx.x == 2) {
                    const double coef = uh_coef(pz, p0, s, n, fm, base, T);
                    atomicAdd(&p.lvlcnt[p.off[c] + level], 1u);
                    atomicMax(&p.depth[c], level + 1);
                    uh_emit_candidate(p, c, level, s, s + fm, e, coef);
                    n_big++;
                }
                __syncthreads();
                for (int k = 0; k < 2; k++) {
                    const int kind = s_child[k][3];
                    if (kind == 2) {
                        uh_big_publish(p, s_child_pos[k], uh_nchunks(s_child[k][0], s_child[k][1]), threadIdx.x, UH_THREADS);
                        nbig_children++;
                    } else if (kind == 3) {
                        have_local = true;
                        lc = c; ls = s_child[k][0]; le = s_child[k][1]; ll = s_child[k][2];
                        nbig_children++;
                    }
                }
            }
            if (threadIdx.x == 0) {
                const int delta = nbig_children - 1;
                if (delta != 0) {
                    __threadfence();
                    const int now = atomicAdd(&ctl->outstanding, delta) + delta;
                    if (now == 0) { __threadfence(); *(volatile int*)&ctl->big_done = 1; }
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }

    // --------------------------------------------------------------------- small workers (per warp)
    UhWarpScratch& ws = s_ws[warp];
    for (;;) {
        int idx = 0, c = -1, s = 0, e = 0, level = 0;
        if (lane == 0) {
            idx = atomicAdd(&ctl->small_head, 1);
            if (idx < p.small_cap) {
                volatile int* ready = &p.small[idx].c;
                unsigned backoff = 100;
                for (;;) {
                    c = *ready;
                    if (c >= 0) break;
                    if (*(volatile int*)&ctl->big_done || *(volatile int*)&ctl->overflow) {
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
            atomicAdd(&p.ctl->outstanding, 1);
            const unsigned long long pos = uh_big_alloc(p, c, 0, e, 0);
            uh_big_publish(p, pos, uh_nchunks(0, e), 0, 1);
        } else {
            uh_push_small(p, c, 0, e, 0);
        }
    }
    __syncthreads();
    // no big node at all: the big phase is over before it starts
    if (threadIdx.x == 0 && s_big == 0) { __threadfence(); *(volatile int*)&p.ctl->big_done = 1; }
}
