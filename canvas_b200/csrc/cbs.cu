// Circular binary segmentation (CanvasPartition -m CBS) on the GPU.
//
// Reference: CBSRunner.Run CBSRunner.cs:40-151; ChangePoint.ChangePoints / FindChangePoints / XPerm
// ChangePoint.cs:44-153, :291-400, :407-421; CBSTStatistic.TMaxO / HTMaxP / TMaxP / TPermP
// CBSTStatistic.cs:19-341, :354-586, :599-934, :947-1024; TailProbability.TailP TailProbability.cs:21-44;
// GetBoundary.ComputeBoundary GetBoundary.cs:19-160.
//
// Shape of the work: chromosomes are independent (own MT19937 stream), the tests inside a chromosome
// form one sequential chain (the random stream is consumed in stack order), and inside a test the
// permutations are independent.  So: one thread-block cluster per chromosome walks the segment stack;
// per test the cluster finds the observed max-t arc in parallel, then runs permutations in batches,
// one thread per permutation, over a random stream that CTA 0 generates with a block-parallel MT19937.
// Every floating-point value that feeds a decision is produced with the reference's operation order
// (sequential sums where the reference sums sequentially), so the accepted change points are the
// reference's; TailP uses device exp/log/erfc and agrees to a few ulp.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace cgx = cooperative_groups;

namespace {

constexpr int CBS_THREADS = 256;      // threads per CTA
constexpr int CBS_PT = 64;            // permutation threads per CTA (batch = CBS_PT * cluster size)
constexpr int CBS_TILE = 2048;        // doubles staged per step of a sequential pass
constexpr int CBS_KMAX = 32;          // largest supported k_max (arc length of the hybrid statistic)
constexpr int CBS_NMIN_MAX = 256;     // largest supported n_min (segments up to this size use the full permuted search)
constexpr int CBS_NB_SMALL = 16;      // round(sqrt(256))

struct CbsOpts {
    double alpha;
    unsigned n_perm;
    int min_width, k_max;
    unsigned n_min;
};

struct CbsWork {
    long long off;
    int n;
    int chrom;
    unsigned seed;
    int pad;
};

struct CbsCand {
    double lim, key;
    int bi, bj, alen, idx;
};

struct CbsBest {
    double v;
    int seq, at, len, pad;
};

struct CbsCtl {
    int work, sp, nloc, ncp;
    int cp[2];
    int same, nb;
    unsigned long long gen_pos, use_pos;
    double avg, tss, bss0, ostat, p1;
    int ti, tj, search, ncand;
    unsigned long long bss_bits;
    int seg0, seg1, verdict, nrejc;
    int nrej, kk, np_done, consumed;
    // tpermp
    double xbar, t_ostat, rm1;
    int t_mode, t_m1, t_nrej, pad2;
    long long st_tests, st_perms, st_steps, st_edge;
    long long ph[8];  // ns per phase: 0 passes, 1 observed search, 2 tail p, 3 stream, 4 shuffle+statistic, 5 edge tests, 6 total
    double tailp_term[128];
};

struct CbsScratch {
    double *cur, *sx, *px, *pstat, *pmin, *pmax;
    int *stack, *locs, *bb, *imin, *imax, *arr;
    CbsCand* cand;
    CbsBest* cbest;
    unsigned* ring;
    CbsCtl* ctl;
    long long px_bytes;
    unsigned ring_mask;
    int n_alloc;
};

struct CbsParams {
    CbsOpts o;
    const double* cov;
    const CbsWork* work;
    int nwork;
    int* queue;
    const unsigned* sbdry;
    CbsScratch* scratch;
    int* n_seg;       // [n_chrom]
    int* seg_len;     // at chrom_off
    double* seg_mean;
    long long* stats;  // [n_chrom][4]
    long long* phase_ns;  // [n_chrom][8]
};

struct Grp {
    cgx::cluster_group cl;
    unsigned rank, size, cta, nctas;
    __device__ Grp() : cl(cgx::this_cluster()) {
        rank = cl.thread_rank();
        size = cl.num_threads();
        cta = cl.block_rank();
        nctas = cl.num_blocks();
    }
    __device__ void sync() {
        __threadfence();
        cl.sync();
        __threadfence();  // acquire side: later plain loads must not be served from a stale L1 line
    }
};

__device__ inline long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define CBS_PHASE(ctl, G, k, t0)                       \
    do {                                               \
        if ((G).rank == 0) {                           \
            const long long now__ = gtime();           \
            (ctl)->ph[k] += now__ - (t0);              \
            (t0) = now__;                              \
        }                                              \
    } while (0)

__device__ inline double ldd(const double* p) { return __ldcg(p); }
__device__ inline int ldi(const int* p) { return __ldcg(p); }
template <typename T>
__device__ inline T vol(const T* p) { return *(const volatile T*)p; }

__device__ inline int round_even_i(double v) { return (int)rint(v); }
__device__ inline double arc_scale(double rn, double r) { return rn / (r * (rn - r)); }

// ---------------------------------------------------------------------------------------------
// MT19937 (MathNet.Numerics MersenneTwister): CTA 0 keeps the state in shared memory and produces
// 624 numbers per three barriers; tempered words go to the ring at their absolute stream position.
// ---------------------------------------------------------------------------------------------
struct MtShared {
    unsigned s[2][624];
    int cur;
};

__device__ inline unsigned mt_mix(unsigned a, unsigned b) {
    const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ inline unsigned mt_temper(unsigned y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// all threads of CTA 0; generates until the stream reaches `need`
__device__ void mt_ensure(MtShared& m, const CbsScratch& S, unsigned long long need) {
    CbsCtl* ctl = S.ctl;
    unsigned long long pos = vol(&ctl->gen_pos);
    const int t = threadIdx.x;
    int cur = m.cur;
    __syncthreads();
    while (pos < need) {
        const unsigned* o = m.s[cur];
        unsigned* w = m.s[cur ^ 1];
        if (t < 227) {
            const unsigned v = o[t + 397] ^ mt_mix(o[t], o[t + 1]);
            w[t] = v;
            S.ring[(unsigned)((pos + t) & S.ring_mask)] = mt_temper(v);
        }
        __syncthreads();
        if (t < 227) {
            const int k = t + 227;
            const unsigned v = w[k - 227] ^ mt_mix(o[k], o[k + 1]);
            w[k] = v;
            S.ring[(unsigned)((pos + k) & S.ring_mask)] = mt_temper(v);
        }
        __syncthreads();
        if (t < 170) {
            const int k = t + 454;
            const unsigned v = w[k - 227] ^ mt_mix(o[k], k == 623 ? w[0] : o[k + 1]);
            w[k] = v;
            S.ring[(unsigned)((pos + k) & S.ring_mask)] = mt_temper(v);
        }
        __syncthreads();
        cur ^= 1;
        pos += 624;
    }
    if (t == 0) { ctl->gen_pos = pos; m.cur = cur; }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Sequential passes (CTA 0).  The reference adds in index order; one thread does the additions out of
// shared memory while the whole CTA moves the tiles.
// ---------------------------------------------------------------------------------------------
struct SeqShared {
    double a[CBS_TILE];
    double b[CBS_TILE];
    double red[2][CBS_THREADS / 32];
};

// sum, min, max of x[0..n): ctl->avg = sum / n (Enumerable.Average), ctl->same = (max == min)
__device__ void pass_average(SeqShared& sh, const double* __restrict__ x, int n, CbsCtl* ctl) {
    const int t = threadIdx.x;
    double mn = INFINITY, mx = -INFINITY, sum = 0.0;
    for (int base = 0; base < n; base += CBS_TILE) {
        const int len = min(CBS_TILE, n - base);
        for (int k = t; k < len; k += CBS_THREADS) {
            const double v = x[base + k];
            sh.a[k] = v;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
        __syncthreads();
        if (t == 0) {
            int k = 0;
            for (; k + 8 <= len; k += 8) {
                const double v0 = sh.a[k], v1 = sh.a[k + 1], v2 = sh.a[k + 2], v3 = sh.a[k + 3];
                const double v4 = sh.a[k + 4], v5 = sh.a[k + 5], v6 = sh.a[k + 6], v7 = sh.a[k + 7];
                sum += v0; sum += v1; sum += v2; sum += v3; sum += v4; sum += v5; sum += v6; sum += v7;
            }
            for (; k < len; k++) sum += sh.a[k];
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((t & 31) == 0) { sh.red[0][t >> 5] = mn; sh.red[1][t >> 5] = mx; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < CBS_THREADS / 32; k++) { mn = fmin(mn, sh.red[0][k]); mx = fmax(mx, sh.red[1][k]); }
        ctl->avg = sum / (double)n;
        ctl->same = (mx == mn) ? 1 : 0;
    }
    __syncthreads();
}

// cur = x - avg, sx = running sum of cur, ctl->tss = sum of cur^2 (Helper.InplaceSub, WeightedSumOfSquares, the sx loops)
__device__ void pass_centre(SeqShared& sh, const double* __restrict__ x, int n, double avg, double* cur, double* sx, CbsCtl* ctl) {
    const int t = threadIdx.x;
    double tss = 0.0, run = 0.0;
    for (int base = 0; base < n; base += CBS_TILE) {
        const int len = min(CBS_TILE, n - base);
        for (int k = t; k < len; k += CBS_THREADS) sh.a[k] = x[base + k] - avg;
        __syncthreads();
        if (t == 0) {
            for (int k = 0; k < len; k++) {
                const double v = sh.a[k];
                tss += (1.0 * v) * v;
                run += v;
                sh.b[k] = run;
            }
        }
        __syncthreads();
        for (int k = t; k < len; k += CBS_THREADS) {
            cur[base + k] = sh.a[k];
            sx[base + k] = sh.b[k];
        }
        __syncthreads();
    }
    if (t == 0) ctl->tss = tss;
}

// ---------------------------------------------------------------------------------------------
// Observed statistic (TMaxO).  Positions are 1-based like the reference: S_p = sx[p - 1].
// ---------------------------------------------------------------------------------------------
__device__ inline int blk_lo(const int* bb, int b) { return b == 1 ? 1 : ldi(bb + b - 2) + 1; }

__device__ void tmaxo(Grp& G, const CbsScratch& S, const CbsOpts& o, int n) {
    CbsCtl* ctl = S.ctl;
    const double rn = (double)n;
    const int nb = n >= 50 ? round_even_i(sqrt((double)n)) : 1;
    const int al0 = o.min_width;
    // block boundaries
    for (int i = G.rank; i < nb; i += G.size) S.bb[i] = round_even_i(rn * ((i + 1.0) / nb));
    if (G.rank == 0) { ctl->nb = nb; ctl->ncand = 0; }
    G.sync();
    // block extremes: first strict minimum / maximum of the running sum inside each block
    for (int b = G.rank; b < nb; b += G.size) {
        const int lo = b == 0 ? 1 : ldi(S.bb + b - 1) + 1, hi = ldi(S.bb + b);
        double mn = ldd(S.sx + lo - 1), mx = mn;
        int imn = lo, imx = lo;
        for (int p = lo + 1; p <= hi; p++) {
            const double v = ldd(S.sx + p - 1);
            if (v < mn) { mn = v; imn = p; }
            if (v > mx) { mx = v; imx = p; }
        }
        S.pmin[b] = mn; S.pmax[b] = mx; S.imin[b] = imn; S.imax[b] = imx;
    }
    G.sync();
    if (G.rank == 0) {
        double gmin = 0, gmax = 0;
        int igmin = n, igmax = n;
        for (int b = 0; b < nb; b++) {
            const double mn = ldd(S.pmin + b), mx = ldd(S.pmax + b);
            if (mn < gmin) { gmin = mn; igmin = ldi(S.imin + b); }
            if (mx > gmax) { gmax = mx; igmax = ldi(S.imax + b); }
        }
        const double psdiff = gmax - gmin;
        const double rj = (double)abs(igmax - igmin);
        double bss = arc_scale(rn, rj) * (psdiff * psdiff);
        int search = 1;
        if (psdiff <= 0) { bss = 0; search = 0; }
        ctl->bss0 = bss;
        ctl->ti = min(igmax, igmin);
        ctl->tj = max(igmax, igmin);
        ctl->search = search;
        ctl->bss_bits = (unsigned long long)__double_as_longlong(bss);
    }
    G.sync();
    const int search = vol(&ctl->search);
    const double bss0 = vol(&ctl->bss0);
    const int nal0 = n - al0;
    if (search) {
        // block pairs that can still reach bss0 (CBSTStatistic.cs:143-212)
        for (int bi = 1 + (int)G.rank; bi <= nb; bi += G.size) {
            const int ilo = blk_lo(S.bb, bi), ihi = ldi(S.bb + bi - 1);
            const double mni = ldd(S.pmin + bi - 1), mxi = ldd(S.pmax + bi - 1);
            for (int bj = bi; bj <= nb; bj++) {
                const int jlo = blk_lo(S.bb, bj), jhi = ldi(S.bb + bj - 1);
                int alenhi = jhi - ilo;
                if (alenhi > nal0) alenhi = nal0;
                int alenlo = bi == bj ? 1 : jlo - ihi;
                if (alenlo < al0) alenlo = al0;
                const double s1 = fabs(ldd(S.pmax + bj - 1) - mni);
                const double s2 = fabs(mxi - ldd(S.pmin + bj - 1));
                const double smx = fmax(s1, s2);
                const double rlo = (double)alenlo, rhi = (double)alenhi;
                const double lim = rn / fmin(rlo * (rn - rlo), rhi * (rn - rhi)) * (smx * smx);
                if (bss0 <= lim) {
                    CbsCand c;
                    double s;
                    if (s1 > s2) { c.alen = abs(ldi(S.imax + bj - 1) - ldi(S.imin + bi - 1)); s = s1; }
                    else { c.alen = abs(ldi(S.imin + bj - 1) - ldi(S.imax + bi - 1)); s = s2; }
                    const double r = (double)c.alen;
                    c.key = arc_scale(rn, r) * (s * s);
                    c.lim = lim; c.bi = bi; c.bj = bj;
                    c.idx = (bi - 1) * nb + (bj - 1);
                    S.cand[atomicAdd(&ctl->ncand, 1)] = c;
                }
            }
        }
    }
    G.sync();
    const int ncand = search ? vol(&ctl->ncand) : 0;
    // scan the arcs of every candidate pair: one warp per pair, one lane per arc length
    {
        const int lane = threadIdx.x & 31;
        const int warp = G.rank >> 5, nwarps = G.size >> 5;
        const double rnov2 = rn / 2;
        for (int c = warp; c < ncand; c += nwarps) {
            const CbsCand cd = S.cand[c];
            CbsBest best;
            best.v = -1.0; best.seq = 0x7fffffff; best.at = 0; best.len = 0; best.pad = 0;
            unsigned long long bound_bits = lane == 0 ? vol(&ctl->bss_bits) : 0ull;
            bound_bits = __shfl_sync(0xffffffffu, bound_bits, 0);
            const double bound = __longlong_as_double((long long)bound_bits);
            if (bound <= cd.lim) {
                const int ilo = blk_lo(S.bb, cd.bi), ihi = ldi(S.bb + cd.bi - 1);
                const int jlo = blk_lo(S.bb, cd.bj), jhi = ldi(S.bb + cd.bj - 1);
                int alenhi = jhi - ilo;
                if (alenhi > nal0) alenhi = nal0;
                int alenlo = cd.bi == cd.bj ? 1 : jlo - ihi;
                if (alenlo < al0) alenlo = al0;
                int alenmax = cd.alen;
                if (alenmax > n - alenmax) alenmax = n - alenmax;
                const int nup = ((double)alenlo <= rnov2 && alenlo <= alenmax) ? alenmax - alenlo + 1 : 0;
                const int alenmax2 = n - alenmax;
                const int ndn = ((double)alenhi >= rnov2 && alenhi >= alenmax2) ? alenhi - alenmax2 + 1 : 0;
                for (int q = lane; q < nup + ndn; q += 32) {
                    const int L = q < nup ? alenlo + q : alenhi - (q - nup);
                    const int ixlo = max(0, jlo - ilo - L), ixhi = max(0, ihi + L - jhi);
                    double mx = 0;
                    int at = ilo + ixlo - 1;
                    for (int i = ilo + ixlo; i <= ihi - ixhi; i++) {
                        const double a = fabs(ldd(S.sx + i + L - 1) - ldd(S.sx + i - 1));
                        if (mx < a) { mx = a; at = i; }
                    }
                    const double v = arc_scale(rn, (double)L) * (mx * mx);
                    if (v > best.v) { best.v = v; best.seq = q; best.at = at; best.len = L; }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best.v, d);
                    const int os = __shfl_xor_sync(0xffffffffu, best.seq, d);
                    const int oa = __shfl_xor_sync(0xffffffffu, best.at, d);
                    const int ol = __shfl_xor_sync(0xffffffffu, best.len, d);
                    if (ov > best.v || (ov == best.v && os < best.seq)) { best.v = ov; best.seq = os; best.at = oa; best.len = ol; }
                }
                if (lane == 0 && best.v > bound) atomicMax(&ctl->bss_bits, (unsigned long long)__double_as_longlong(best.v));
            }
            if (lane == 0) S.cbest[c] = best;
        }
    }
    G.sync();
    // the arc the reference ends on: the first one, in its visiting order, that attains the maximum
    if (G.cta == 0) {
        __shared__ double r_v[CBS_THREADS], r_key[CBS_THREADS];
        __shared__ int r_idx[CBS_THREADS], r_c[CBS_THREADS];
        const int t = threadIdx.x;
        double bv = -1.0, bk = -INFINITY;
        int bidx = -1, bc = -1;
        for (int c = t; c < ncand; c += CBS_THREADS) {
            const double v = __ldcg(&S.cbest[c].v);
            double k = __ldcg(&S.cand[c].key);
            if (k != k) k = -INFINITY;
            const int ix = __ldcg(&S.cand[c].idx);
            if (v > bv || (v == bv && (k > bk || (k == bk && ix > bidx)))) { bv = v; bk = k; bidx = ix; bc = c; }
        }
        r_v[t] = bv; r_key[t] = bk; r_idx[t] = bidx; r_c[t] = bc;
        __syncthreads();
        for (int s = CBS_THREADS / 2; s > 0; s >>= 1) {
            if (t < s) {
                const double v = r_v[t + s], k = r_key[t + s];
                const int ix = r_idx[t + s];
                if (v > r_v[t] || (v == r_v[t] && (k > r_key[t] || (k == r_key[t] && ix > r_idx[t])))) {
                    r_v[t] = v; r_key[t] = k; r_idx[t] = ix; r_c[t] = r_c[t + s];
                }
            }
            __syncthreads();
        }
        if (t == 0) {
            double bss = bss0;
            int ti = ctl->ti, tj = ctl->tj;
            if (r_c[0] >= 0 && r_v[0] > bss) {
                const CbsBest b = S.cbest[r_c[0]];
                bss = b.v; ti = b.at; tj = b.at + b.len;
            }
            double tss = ctl->tss;
            if (tss <= bss + 0.0001) tss = bss + 1.0;
            ctl->ostat = bss / ((tss - bss) / (rn - 2.0));
            ctl->seg0 = ti;
            ctl->seg1 = tj;
        }
    }
    G.sync();
}

// ---------------------------------------------------------------------------------------------
// TailProbability.TailP: the 100 grid terms in parallel, summed in order
// ---------------------------------------------------------------------------------------------
__device__ inline double dev_pnorm(double x) { return 0.5 * erfc(-x / sqrt(2.0)); }

// Nu (TailProbability.cs:46-82): the series is summed chunk by chunk (2, 2, 4, 8, ... terms, the reference's
// convergence checkpoints); a warp computes the terms of a chunk in parallel.
__device__ double warp_nu(double x, double tol, int lane) {
    double l1;
    if (x > 0.01) {
        l1 = log(2.0) - 2 * log(x);
        double l0;
        long long done = 0;
        int k = 2;
        bool first = true;
        while (true) {
            l0 = l1;
            double part = 0.0;
            for (long long q = lane; q < k; q += 32) {
                const double dk = (double)(done + q + 1);
                part += 2.0 * dev_pnorm(-x * sqrt(dk) / 2.0) / dk;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            l1 -= part;
            done += k;
            if (!first) k *= 2;
            if (!first && !(fabs((l1 - l0) / l1) > tol)) break;
            if (first) {
                first = false;
                // after the initial two terms the reference enters the while loop only if the test holds
                // against lnu0 = the value before them
                if (!(fabs((l1 - l0) / l1) > tol)) break;
            }
        }
    } else {
        l1 = -0.583 * x;
    }
    return exp(l1);
}

__device__ double dev_integral(double x, double a) {
    double y = x + a - 0.5;
    double v = (8.0 * y) / (1.0 - 4.0 * (y * y)) + 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    y = x - 0.5;
    v = v - (8.0 * y) / (1.0 - 4.0 * (y * y)) - 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    return v;
}

// whole group; the caller syncs afterwards and rank 0 adds the terms up in order
__device__ void tailp_terms(Grp& G, CbsCtl* ctl, double b, double delta, int m) {
    const int ngrid = 100;
    const int lane = threadIdx.x & 31;
    const int warp = G.rank >> 5, nwarps = G.size >> 5;
    const double dincr = (0.5 - delta) / ngrid, bsqrtm = b / sqrt((double)m);
    for (int g = warp; g < ngrid; g += nwarps) {
        double tl = 0.5 - dincr, tt = 0.5 - 0.5 * dincr;
        for (int i = 0; i <= g; i++) { tl += dincr; tt += dincr; }
        const double v = warp_nu(bsqrtm / sqrt(tt * (1 - tt)), 1e-6, lane);
        if (lane == 0) ctl->tailp_term[g] = (v * v) * dev_integral(tl, dincr);
    }
}

__device__ double tailp_finish(const CbsCtl* ctl, double b) {
    double acc = 0.0;
    for (int i = 0; i < 100; i++) acc += vol(&ctl->tailp_term[i]);
    acc = 9.973557E-2 * (b * b * b) * exp(-(b * b) / 2) * acc;
    return 2.0 * acc;
}

// ---------------------------------------------------------------------------------------------
// One permutation per thread.  The column of thread t lives at px[i * CBS_PT + t].
// ---------------------------------------------------------------------------------------------
__device__ inline unsigned ring_at(const CbsScratch& S, unsigned long long pos) { return __ldcg(S.ring + (unsigned)(pos & S.ring_mask)); }

// XPerm (ChangePoint.cs:407-421): j = (int)(NextDouble() * (i + 1)) = (u32 * (i + 1)) >> 32 exactly.
// CBS_U steps are in flight per thread: their loads are issued together (with the next group's random
// numbers), then the group is replayed in order on the loaded values, patching every value that an
// earlier step of the same group has moved.
__device__ void shuffle_column(const CbsScratch& S, double* col, int n, unsigned long long start) {
    constexpr int CBS_U = 16;
    unsigned rv[CBS_U];
    int t0 = 0;
    if (n >= CBS_U) {
#pragma unroll
        for (int k = 0; k < CBS_U; k++) rv[k] = ring_at(S, start + k);
    }
    for (; t0 + CBS_U <= n; t0 += CBS_U) {
        int iv[CBS_U], jv[CBS_U];
        double a[CBS_U], b[CBS_U];
#pragma unroll
        for (int k = 0; k < CBS_U; k++) {
            const int i = n - 1 - (t0 + k);
            iv[k] = i;
            jv[k] = (int)(((unsigned long long)rv[k] * (unsigned long long)(i + 1)) >> 32);
        }
#pragma unroll
        for (int k = 0; k < CBS_U; k++) {
            a[k] = col[(size_t)iv[k] * CBS_PT];
            b[k] = col[(size_t)jv[k] * CBS_PT];
        }
        if (t0 + 2 * CBS_U <= n) {
#pragma unroll
            for (int k = 0; k < CBS_U; k++) rv[k] = ring_at(S, start + t0 + CBS_U + k);
        }
#pragma unroll
        for (int k = 0; k < CBS_U; k++) {
            double av = a[k];
#pragma unroll
            for (int m = 0; m < k; m++)
                if (jv[m] == iv[k]) av = a[m];
            double bv = b[k];
#pragma unroll
            for (int m = 0; m < k; m++)
                if (jv[m] == jv[k]) bv = a[m];
            if (jv[k] == iv[k]) bv = av;
            a[k] = av;
            b[k] = bv;
        }
#pragma unroll
        for (int k = 0; k < CBS_U; k++) {
            col[(size_t)iv[k] * CBS_PT] = b[k];
            col[(size_t)jv[k] * CBS_PT] = a[k];
        }
    }
    for (; t0 < n; t0++) {
        const int i = n - 1 - t0;
        const int j = (int)(((unsigned long long)ring_at(S, start + t0) * (unsigned long long)(i + 1)) >> 32);
        const double av = col[(size_t)i * CBS_PT], bv = col[(size_t)j * CBS_PT];
        col[(size_t)i * CBS_PT] = bv;
        col[(size_t)j * CBS_PT] = av;
    }
}

// HTMaxP: max over arcs no longer than k, including the arcs that wrap around the end.  The block
// pruning of the reference never changes the maximum (its bounds hold exactly in floating point), so
// the arcs are simply enumerated: M[j] = max_p |S_p - S_{p-j}|, p - j >= 1.  The running sums of the last
// K positions rotate through registers; the ring starts as NaN so that arcs reaching before S_1 drop out.
// Four threads share one permutation: each of them runs the (identical) running sum and owns every
// fourth arc length.
constexpr int CBS_PARTS = CBS_THREADS / CBS_PT;

template <int K, int AL0, int PART, bool CHECK>
__device__ __forceinline__ void ht_group(const double* col, int p, int n, int k, int al0, double& run, double (&M)[K + 1], double (&h)[K],
                                         double* first, double* last) {
#pragma unroll
    for (int q = 0; q < K; q++) {
        if (!CHECK || p + q <= n) {
            run += col[(size_t)(p + q - 1) * CBS_PT];
#pragma unroll
            for (int j = 2; j <= K; j++) {
                if ((j & (CBS_PARTS - 1)) != PART) continue;
                const bool on = AL0 > 0 ? (j >= AL0) : (j >= al0 && j <= k);
                if (on) M[j] = fmax(M[j], fabs(run - h[(q - j + 2 * K) % K]));
            }
            h[q] = run;
            if (CHECK && PART == 0) {
                if (p + q <= k) first[(p + q) * CBS_PT] = run;
                if (p + q > n - k) last[(p + q - (n - k)) * CBS_PT] = run;
            }
        }
    }
}

// called by all threads of the CTA (barrier inside); returns the largest c_j * M_j^2 over the thread's arc lengths
template <int K, int AL0, int PART>
__device__ double htmaxp_part(const double* col, bool active, int n, int k, int al0, double* first, double* last) {
    const double rn = (double)n;
    double M[K + 1];
    double h[K];
#pragma unroll
    for (int j = 0; j <= K; j++) M[j] = 0.0;
#pragma unroll
    for (int q = 0; q < K; q++) h[q] = __longlong_as_double(0x7ff8000000000000ll);
    if (active) {
        double run = 0.0;
        int p = 1;
        ht_group<K, AL0, PART, true>(col, p, n, k, al0, run, M, h, first, last);
        p += K;
        for (; p + K - 1 <= n - k; p += K) ht_group<K, AL0, PART, false>(col, p, n, k, al0, run, M, h, first, last);
        for (; p <= n; p += K) ht_group<K, AL0, PART, true>(col, p, n, k, al0, run, M, h, first, last);
    }
    __syncthreads();
    double best = 0.0;
    if (active) {
#pragma unroll
        for (int j = 2; j <= K; j++) {
            if ((j & (CBS_PARTS - 1)) != PART) continue;
            const bool on = AL0 > 0 ? (j >= AL0) : (j >= al0 && j <= k);
            if (on) {
                double mx = M[j];
                // arcs through the end (CBSTStatistic.cs:459-485): |S_{n-j+i} - S_i|, i = 1..j
                for (int i = 1; i <= j; i++) mx = fmax(mx, fabs(last[(k - j + i) * CBS_PT] - first[i * CBS_PT]));
                best = fmax(best, arc_scale(rn, (double)j) * (mx * mx));
            }
        }
    }
    return best;
}

template <int K, int AL0>
__device__ double htmaxp_parts(int part, const double* col, bool active, int n, int k, int al0, double* first, double* last) {
    switch (part) {
        case 0: return htmaxp_part<K, AL0, 0>(col, active, n, k, al0, first, last);
        case 1: return htmaxp_part<K, AL0, 1>(col, active, n, k, al0, first, last);
        case 2: return htmaxp_part<K, AL0, 2>(col, active, n, k, al0, first, last);
        default: return htmaxp_part<K, AL0, 3>(col, active, n, k, al0, first, last);
    }
}

// TMaxP for the short segments (n <= n_min): the full search on the permuted column, which is turned into its
// running sums in place.  Pairs are visited in natural order with the running maximum as the filter; a
// pair the reference skips cannot hold the maximum, so the value is the same.
__device__ double tmaxp_column(double* col, int n, int al0, double tss) {
    const double rn = (double)n;
    const int nb = n >= 50 ? round_even_i(sqrt((double)n)) : 1;
    int bb[CBS_NB_SMALL], imin[CBS_NB_SMALL], imax[CBS_NB_SMALL];
    double pmin[CBS_NB_SMALL], pmax[CBS_NB_SMALL];
    for (int i = 0; i < nb; i++) bb[i] = round_even_i(rn * ((double)(i + 1) / nb));
    double gmin = 0, gmax = 0, run = 0;
    int igmin = n, igmax = n, lo = 1;
    for (int b = 0; b < nb; b++) {
        double mn = 0, mx = 0;
        int imn = lo, imx = lo;
        for (int p = lo; p <= bb[b]; p++) {
            run += col[(size_t)(p - 1) * CBS_PT];
            col[(size_t)(p - 1) * CBS_PT] = run;
            if (p == lo) { mn = mx = run; }
            else {
                if (run < mn) { mn = run; imn = p; }
                if (run > mx) { mx = run; imx = p; }
            }
        }
        imin[b] = imn; imax[b] = imx; pmin[b] = mn; pmax[b] = mx;
        if (mn < gmin) { gmin = mn; igmin = imn; }
        if (mx > gmax) { gmax = mx; igmax = imx; }
        lo = bb[b] + 1;
    }
    const double psdiff = gmax - gmin;
    double bss = arc_scale(rn, (double)abs(igmax - igmin)) * (psdiff * psdiff);
    const double rnov2 = rn / 2;
    const int nal0 = n - al0;
    for (int bi = 1; bi <= nb; bi++)
        for (int bj = bi; bj <= nb; bj++) {
            const int ilo = bi == 1 ? 1 : bb[bi - 2] + 1, ihi = bb[bi - 1];
            const int jlo = bj == 1 ? 1 : bb[bj - 2] + 1, jhi = bb[bj - 1];
            int alenhi = jhi - ilo;
            if (alenhi > nal0) alenhi = nal0;
            int alenlo = bi == bj ? 1 : jlo - ihi;
            if (alenlo < al0) alenlo = al0;
            const double s1 = fabs(pmax[bj - 1] - pmin[bi - 1]), s2 = fabs(pmax[bi - 1] - pmin[bj - 1]);
            const double smx = fmax(s1, s2);
            const double rlo = (double)alenlo, rhi = (double)alenhi;
            const double lim = rn / fmin(rlo * (rn - rlo), rhi * (rn - rhi)) * (smx * smx);
            if (!(bss <= lim)) continue;
            int alenmax = s1 > s2 ? abs(imax[bj - 1] - imin[bi - 1]) : abs(imin[bj - 1] - imax[bi - 1]);
            if (alenmax > n - alenmax) alenmax = n - alenmax;
            for (int phase = 0; phase < 2; phase++) {
                int La, Lb;
                if (phase == 0) {
                    if (!((double)alenlo <= rnov2 && alenlo <= alenmax)) continue;
                    La = alenlo; Lb = alenmax;
                } else {
                    const int a2 = n - alenmax;
                    if (!((double)alenhi >= rnov2 && alenhi >= a2)) continue;
                    La = a2; Lb = alenhi;
                }
                for (int L = La; L <= Lb; L++) {
                    const int ixlo = max(0, jlo - ilo - L), ixhi = max(0, ihi + L - jhi);
                    double mx = 0;
                    for (int i = ilo + ixlo; i <= ihi - ixhi; i++)
                        mx = fmax(mx, fabs(col[(size_t)(i + L - 1) * CBS_PT] - col[(size_t)(i - 1) * CBS_PT]));
                    bss = fmax(bss, arc_scale(rn, (double)L) * (mx * mx));
                }
            }
        }
    if (tss <= bss + 0.0001) tss = bss + 1.0;
    return bss / ((tss - bss) / (rn - 2.0));
}

// ---------------------------------------------------------------------------------------------
// TPermP (CBSTStatistic.cs:947-1024).  The 10 000 partial shuffles act on ONE array, one after the
// other.  They are cut into blocks of consecutive permutations: a block only touches the m1 top
// positions and the bottom positions its draws hit, so (1) every block works out, in parallel, the
// permutation of its touched positions; (2) the blocks are chained in order over the arrangement A
// (which element sits where), recording each block's start contents; (3) every block replays its
// shuffles on the real values, in parallel, adding up the top sum after each permutation.
// ---------------------------------------------------------------------------------------------
struct TpBlock {
    int* top;     // [m1]   content id (phase 1) / element index (phase 3) of the top slots
    int* hkey;    // [H]    bottom position or -1
    int* hval;    // [H]
    int* klist;   // [cap]  slots in insertion order
    int* startv;  // [m1 + H]
    int* cnt;     // [1]
};

__device__ inline TpBlock tp_block(char* base, long long stride, int b, int m1, int H, int cap) {
    TpBlock B;
    int* p = (int*)(base + stride * b);
    B.cnt = p;
    B.top = p + 4;
    B.hkey = B.top + m1;
    B.hval = B.hkey + H;
    B.klist = B.hval + H;
    B.startv = B.klist + cap;
    return B;
}
__host__ __device__ inline long long tp_stride(int m1, int H, int cap) { return (((long long)(4 + m1 + H + H + cap + m1 + H) * 4 + 15) / 16) * 16; }

__device__ inline int tp_slot(const TpBlock& B, int H, int key, bool insert, int m1) {
    unsigned h = ((unsigned)key * 2654435761u) & (unsigned)(H - 1);
    while (true) {
        const int k = B.hkey[h];
        if (k == key) return (int)h;
        if (k < 0) {
            if (!insert) return -1;
            B.hkey[h] = key;
            B.hval[h] = m1 + (int)h;
            B.klist[(*B.cnt)++] = (int)h;
            return (int)h;
        }
        h = (h + 1) & (unsigned)(H - 1);
    }
}

// phase 1 (replay == false): contents are ids; phase 3 (replay == true): contents are element indices and the
// top sums are compared with the observed statistic
__device__ int tp_run_block(const CbsScratch& S, const TpBlock& B, int H, int n12, int m1, int nperm_blk, unsigned long long start,
                            bool replay, const double* x, double rm1, double xbar, double ostat) {
    const int base = n12 - m1;
    int nrej = 0;
    unsigned long long pos = start;
    for (int p = 0; p < nperm_blk; p++) {
        double sum = 0;
        for (int i = n12 - 1; i >= base; i--, pos++) {
            const int j = (int)(((unsigned long long)ring_at(S, pos) * (unsigned long long)(i + 1)) >> 32);
            int* pi = B.top + (i - base);
            int* pj = j >= base ? B.top + (j - base) : B.hval + tp_slot(B, H, j, !replay, m1);
            const int vi = *pi, vj = *pj;
            *pi = vj;
            *pj = vi;
            if (replay) sum += ldd(x + vj);
        }
        if (replay && ostat <= fabs(sum / rm1 - xbar)) nrej++;
    }
    return nrej;
}

union CbsShared {
    SeqShared seq;
    struct { double first[(CBS_KMAX + 1) * CBS_PT]; double last[(CBS_KMAX + 1) * CBS_PT]; double part[CBS_THREADS]; } ht;
};

// group-wide; result in ctl->t_nrej (count of permutations at least as extreme), p = t_nrej / n_perm
__device__ void tpermp(Grp& G, const CbsScratch& S, const CbsOpts& o, CbsShared& sh, MtShared& mt, int n1, int n2, int n12, int xoff) {
    CbsCtl* ctl = S.ctl;
    const double* x = S.cur + xoff;
    const int t = threadIdx.x;
    if (n1 == 1 || n2 == 1) {
        if (G.rank == 0) ctl->t_nrej = (int)o.n_perm;
        G.sync();
        return;
    }
    if (G.cta == 0) {
        // xsum1, xsum2 and tss in index order
        double s1 = 0, s2 = 0, tss = 0;
        for (int base = 0; base < n12; base += CBS_TILE) {
            const int len = min(CBS_TILE, n12 - base);
            for (int k = t; k < len; k += CBS_THREADS) sh.seq.a[k] = ldd(x + base + k);
            __syncthreads();
            if (t == 0) {
                for (int k = 0; k < len; k++) {
                    const double v = sh.seq.a[k];
                    if (base + k < n1) s1 += v; else s2 += v;
                    tss += v * v;
                }
            }
            __syncthreads();
        }
        if (t == 0) {
            const double rn1 = n1, rn2 = n2, rn = rn1 + rn2;
            const double xbar = (s1 + s2) / rn;
            tss = tss - rn * (xbar * xbar);
            int m1;
            double rm1, ostat, tstat;
            if (n1 <= n2) { m1 = n1; rm1 = rn1; ostat = 0.99999 * fabs(s1 / rn1 - xbar); tstat = (ostat * ostat) * rn1 * rn / rn2; }
            else { m1 = n2; rm1 = rn2; ostat = 0.99999 * fabs(s2 / rn2 - xbar); tstat = (ostat * ostat) * rn2 * rn / rn1; }
            tstat = tstat / ((tss - tstat) / (rn - 2.0));
            ctl->xbar = xbar; ctl->t_ostat = ostat; ctl->rm1 = rm1; ctl->t_m1 = m1;
            ctl->t_mode = (tstat > 25 && m1 >= 10) ? 0 : 1;
            ctl->t_nrej = 0;
        }
    }
    G.sync();
    if (vol(&ctl->t_mode) == 0) return;
    const int m1 = vol(&ctl->t_m1);
    const double rm1 = vol(&ctl->rm1), xbar = vol(&ctl->xbar), ostat = vol(&ctl->t_ostat);
    const int P = (int)o.n_perm;
    const int nthr = CBS_PT * (int)G.nctas;
    // permutations per block: balance the chain (one round per block) against the per-block replay
    int pp = (int)sqrt((double)P * 10.0 / (2.0 * m1));
    pp = max(1, min(pp, P));
    int nblk, H, cap;
    long long stride;
    while (true) {
        cap = pp * m1;
        H = 16;
        while (H < 2 * cap) H <<= 1;
        stride = tp_stride(m1, H, cap);
        nblk = min(nthr, (P + pp - 1) / pp);
        const long long fit = S.px_bytes / stride;
        if (fit >= nblk || pp == 1) { nblk = (int)min((long long)nblk, max(1ll, fit)); break; }
        pp = max(1, pp / 2);
    }
    // keep one wave's draws inside the ring
    while ((long long)nblk * pp * m1 > (long long)S.ring_mask - 2 * 624 && nblk > 1) nblk = max(1, nblk / 2);
    char* pool = (char*)S.px;
    // arrangement A: identity
    for (int q = G.rank; q < n12; q += G.size) S.arr[q] = q;
    const int base = n12 - m1;
    const bool mine = t < CBS_PT;
    const int myblk = (int)G.cta * CBS_PT + t;
    for (int w0 = 0; w0 < P; w0 += nblk * pp) {
        const int wave = min(P - w0, nblk * pp);
        const int nb = (wave + pp - 1) / pp;
        const unsigned long long ubase = vol(&ctl->use_pos);
        if (G.cta == 0) mt_ensure(mt, S, ubase + (unsigned long long)wave * m1);
        // clear the tables
        for (long long q = G.rank; q < (long long)nb * (m1 + H); q += G.size) {
            const int b = (int)(q / (m1 + H)), r = (int)(q % (m1 + H));
            TpBlock B = tp_block(pool, stride, b, m1, H, cap);
            if (r < m1) B.top[r] = r; else B.hkey[r - m1] = -1;
            if (r == 0) *B.cnt = 0;
        }
        G.sync();
        if (mine && myblk < nb) {
            TpBlock B = tp_block(pool, stride, myblk, m1, H, cap);
            const int np = min(pp, wave - myblk * pp);
            tp_run_block(S, B, H, n12, m1, np, ubase + (unsigned long long)myblk * pp * m1, false, x, rm1, xbar, ostat);
        }
        G.sync();
        // chain (CTA 0): start contents of block b, then the arrangement after it
        if (G.cta == 0) {
            for (int b = 0; b < nb; b++) {
                TpBlock B = tp_block(pool, stride, b, m1, H, cap);
                const int touched = m1 + __ldcg(B.cnt);
                for (int u = t; u < touched; u += CBS_THREADS) {
                    int id, pos;
                    if (u < m1) { id = u; pos = base + u; }
                    else { const int slot = __ldcg(B.klist + u - m1); id = m1 + slot; pos = __ldcg(B.hkey + slot); }
                    B.startv[id] = S.arr[pos];
                }
                __syncthreads();
                for (int u = t; u < touched; u += CBS_THREADS) {
                    int c, pos;
                    if (u < m1) { c = __ldcg(B.top + u); pos = base + u; }
                    else { const int slot = __ldcg(B.klist + u - m1); c = __ldcg(B.hval + slot); pos = __ldcg(B.hkey + slot); }
                    S.arr[pos] = B.startv[c];
                }
                __syncthreads();
            }
        }
        G.sync();
        if (mine && myblk < nb) {
            TpBlock B = tp_block(pool, stride, myblk, m1, H, cap);
            const int cnt = *B.cnt;
            for (int s = 0; s < m1; s++) B.top[s] = __ldcg(B.startv + s);
            for (int u = 0; u < cnt; u++) { const int slot = B.klist[u]; B.hval[slot] = __ldcg(B.startv + m1 + slot); }
            const int np = min(pp, wave - myblk * pp);
            const int r = tp_run_block(S, B, H, n12, m1, np, ubase + (unsigned long long)myblk * pp * m1, true, x, rm1, xbar, ostat);
            if (r) atomicAdd(&ctl->t_nrej, r);
        }
        if (G.rank == 0) { ctl->use_pos = ubase + (unsigned long long)wave * m1; ctl->st_edge += (long long)wave * m1; }
        G.sync();
    }
}

// ---------------------------------------------------------------------------------------------
// fndcpt for the segment cur[0..n) (already centred): ctl->ncp, ctl->cp[]
// ---------------------------------------------------------------------------------------------
__device__ void find_change_points(Grp& G, const CbsScratch& S, const CbsParams& P, CbsShared& sh, MtShared& mt, int n) {
    CbsCtl* ctl = S.ctl;
    const CbsOpts& o = P.o;
    const int t = threadIdx.x;
    long long tph = gtime();
    tmaxo(G, S, o, n);
    CBS_PHASE(ctl, G, 1, tph);
    const bool hybrid = o.n_min < (unsigned)n;
    // 0 = no change point, 1 = permutations needed, 2 = split without permutations
    if (G.rank == 0) {
        const double ostat1 = sqrt(ctl->ostat);
        const int l = min(ctl->seg1 - ctl->seg0, n - ctl->seg1 + ctl->seg0);
        int verdict;
        if (ostat1 <= 0.1) verdict = 0;
        else if (ostat1 >= 7.0 && l >= 10) verdict = 2;
        else verdict = 1;
        ctl->verdict = verdict;
        ctl->st_tests++;
    }
    G.sync();
    if (vol(&ctl->verdict) == 1) {
        const double ostat1 = sqrt(vol(&ctl->ostat));
        if (hybrid) {
            tailp_terms(G, ctl, ostat1, (o.k_max + 1.0) / n, n);
            G.sync();
        }
        if (G.rank == 0) {
            int nrejc;
            if (hybrid) {
                const double p1 = tailp_finish(ctl, ostat1);
                ctl->p1 = p1;
                if (p1 > o.alpha) ctl->verdict = 0;
                nrejc = (int)((o.alpha - p1) * o.n_perm);
            } else {
                nrejc = (int)(o.alpha * o.n_perm);
            }
            ctl->nrejc = nrejc;
            ctl->kk = nrejc * (nrejc + 1) / 2 + 1;
            ctl->nrej = 0;
            ctl->np_done = 0;
        }
    }
    G.sync();
    CBS_PHASE(ctl, G, 2, tph);
    const double ostat = vol(&ctl->ostat) * 0.99999;
    const double tss = vol(&ctl->tss);
    const int batch = CBS_PT * (int)G.nctas;
    // Every batch is the cluster's full capacity: a batch costs the same time whatever its size (the shuffle is
    // bound by per-SM sector bandwidth and latency), so smaller first batches only add rounds (measured).
    int grow = batch;
    while (vol(&ctl->verdict) == 1) {
        const int np_done = vol(&ctl->np_done);
        const int nb = min(grow, (int)o.n_perm - np_done);
        grow = min(batch, grow * 2);
        const unsigned long long ubase = vol(&ctl->use_pos);
        if (G.cta == 0) mt_ensure(mt, S, ubase + (unsigned long long)nb * n);
        G.sync();
        CBS_PHASE(ctl, G, 3, tph);
        {
            const int pt = t & (CBS_PT - 1), part = t / CBS_PT;
            const int b = pt * (int)G.nctas + (int)G.cta;  // permutation b of the batch: spread over the CTAs
            const bool active = b < nb;
            double* col = S.px + (size_t)G.cta * CBS_PT * S.n_alloc + pt;
            if (active)
                for (int i = part; i < n; i += CBS_PARTS) col[(size_t)i * CBS_PT] = ldd(S.cur + i);
            __syncthreads();
            if (active && part == 0) {
                const long long ts0 = gtime();
                shuffle_column(S, col, n, ubase + (unsigned long long)b * n);
                if (G.rank == 0) ctl->ph[7] += gtime() - ts0;
            }
            __syncthreads();
            if (hybrid) {
                double best;
                if (o.k_max == 25 && o.min_width == 2) best = htmaxp_parts<25, 2>(part, col, active, n, 25, 2, sh.ht.first + pt, sh.ht.last + pt);
                else best = htmaxp_parts<CBS_KMAX, 0>(part, col, active, n, o.k_max, o.min_width, sh.ht.first + pt, sh.ht.last + pt);
                sh.ht.part[t] = best;
                __syncthreads();
                if (active && part == 0) {
                    for (int r = 1; r < CBS_PARTS; r++) best = fmax(best, sh.ht.part[r * CBS_PT + pt]);
                    double ts = tss;
                    if (ts <= best + 0.0001) ts = best + 1.0;
                    S.pstat[b] = best / ((ts - best) / ((double)n - 2.0));
                }
            } else if (active && part == 0) {
                S.pstat[b] = tmaxp_column(col, n, o.min_width, tss);
            }
        }
        G.sync();
        if (G.rank == 0) {
            int nrej = ctl->nrej, k = ctl->kk, verdict = 1, consumed = nb;
            const int nrejc = ctl->nrejc;
            for (int q = 0; q < nb; q++) {
                const int np = np_done + q + 1;
                if (ostat <= ldd(S.pstat + q)) { nrej++; k++; }
                if (nrej > nrejc) { verdict = 0; consumed = q + 1; break; }
                if ((unsigned)np >= __ldg(P.sbdry + k - 1)) { verdict = 2; consumed = q + 1; break; }
            }
            if (verdict == 1 && np_done + nb >= (int)o.n_perm) verdict = 2;
            ctl->nrej = nrej; ctl->kk = k; ctl->verdict = verdict;
            ctl->np_done = np_done + consumed;
            ctl->use_pos = ubase + (unsigned long long)consumed * n;
            ctl->st_perms += consumed;
            ctl->st_steps += (long long)consumed * n;
        }
        G.sync();
        CBS_PHASE(ctl, G, 4, tph);
    }
    if (vol(&ctl->verdict) == 0) {
        if (G.rank == 0) ctl->ncp = 0;
        G.sync();
        return;
    }
    const int seg0 = vol(&ctl->seg0), seg1 = vol(&ctl->seg1);
    if (seg1 == n || seg0 == 0) {
        if (G.rank == 0) { ctl->ncp = 1; ctl->cp[0] = seg1 == n ? seg0 : seg1; }
        G.sync();
        return;
    }
    int ncp = 0, cp0 = 0, cp1 = 0;
    tpermp(G, S, o, sh, mt, seg0, seg1 - seg0, seg1, 0);
    if ((double)vol(&ctl->t_nrej) / o.n_perm <= o.alpha) { ncp = 1; cp0 = seg0; }
    G.sync();
    tpermp(G, S, o, sh, mt, seg1 - seg0, n - seg1, n - seg0, seg0);
    if ((double)vol(&ctl->t_nrej) / o.n_perm <= o.alpha) {
        ncp++;
        if (ncp == 1) cp0 = seg1; else cp1 = seg1;
    }
    G.sync();
    if (G.rank == 0) { ctl->ncp = ncp; ctl->cp[0] = cp0; ctl->cp[1] = cp1; }
    G.sync();
    CBS_PHASE(ctl, G, 5, tph);
}

// ---------------------------------------------------------------------------------------------
// ChangePoints for one chromosome
// ---------------------------------------------------------------------------------------------
__device__ void run_chromosome(Grp& G, const CbsScratch& S, const CbsParams& P, CbsShared& sh, MtShared& mt, const CbsWork& w) {
    CbsCtl* ctl = S.ctl;
    const int t = threadIdx.x;
    const int n = w.n;
    const double* g = P.cov + w.off;
    if (G.cta == 0) {
        if (t == 0) {
            unsigned* s = mt.s[0];
            s[0] = w.seed;
            for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (unsigned)i;
            mt.cur = 0;
            ctl->gen_pos = 0;
            ctl->use_pos = 0;
            S.stack[0] = 0;
            S.stack[1] = n;
            ctl->sp = 2;
            ctl->nloc = 0;
            ctl->st_tests = ctl->st_perms = ctl->st_steps = ctl->st_edge = 0;
            for (int k = 0; k < 8; k++) ctl->ph[k] = 0;
        }
        __syncthreads();
    }
    G.sync();
    long long tph = gtime();
    const long long tstart = tph;
    while (true) {
        const int sp = vol(&ctl->sp);
        if (sp <= 1) break;
        const int a = ldi(S.stack + sp - 2), cn = ldi(S.stack + sp - 1) - a;
        bool tested = false;
        if (cn >= 2 * P.o.min_width) {
            if (G.cta == 0) pass_average(sh.seq, g + a, cn, ctl);
            G.sync();
            if (!vol(&ctl->same)) {
                if (G.cta == 0) pass_centre(sh.seq, g + a, cn, vol(&ctl->avg), S.cur, S.sx, ctl);
                G.sync();
                CBS_PHASE(ctl, G, 0, tph);
                find_change_points(G, S, P, sh, mt, cn);
                tph = gtime();
                tested = true;
            }
        }
        if (G.rank == 0) {
            const int ncp = tested ? ctl->ncp : 0;
            const int end = S.stack[sp - 1];
            if (ncp == 0) {
                S.locs[ctl->nloc++] = end;
                ctl->sp = sp - 1;
            } else {
                S.stack[sp - 1] = ctl->cp[0] + a;
                if (ncp == 2) S.stack[sp] = ctl->cp[1] + a;
                S.stack[sp - 1 + ncp] = end;
                ctl->sp = sp + ncp;
            }
        }
        G.sync();
    }
    // segment lengths (change locations were collected right to left) and means (Helper.WeightedAverage)
    const int nseg = vol(&ctl->nloc);
    for (int s = G.rank; s < nseg; s += G.size) {
        const int end = ldi(S.locs + nseg - 1 - s);
        const int beg = s == 0 ? 0 : ldi(S.locs + nseg - s);
        double sum = 0.0, wsum = 0.0;
        for (int p = beg; p < end; p++) { wsum += 1.0; sum += g[p] * 1.0; }
        P.seg_len[w.off + s] = end - beg;
        P.seg_mean[w.off + s] = sum / wsum;
    }
    if (G.rank == 0) {
        P.n_seg[w.chrom] = nseg;
        P.stats[w.chrom * 4 + 0] = ctl->st_tests;
        P.stats[w.chrom * 4 + 1] = ctl->st_perms;
        P.stats[w.chrom * 4 + 2] = ctl->st_steps;
        P.stats[w.chrom * 4 + 3] = ctl->st_edge;
        ctl->ph[6] = gtime() - tstart;
        for (int k = 0; k < 8; k++) P.phase_ns[w.chrom * 8 + k] = ctl->ph[k];
    }
    G.sync();
}

__global__ void __launch_bounds__(CBS_THREADS, 1) cbs_kernel(CbsParams P) {
    __shared__ CbsShared sh;
    __shared__ MtShared mt;
    Grp G;
    const int cid = blockIdx.x / G.nctas;
    const int nclusters = gridDim.x / G.nctas;
    const CbsScratch S = P.scratch[cid];
    // the first chromosome of a cluster is fixed (scratch is sized for it); later ones come from the queue
    int w = cid;
    while (w < P.nwork) {
        run_chromosome(G, S, P, sh, mt, P.work[w]);
        if (G.rank == 0) S.ctl->work = nclusters + atomicAdd(P.queue, 1);
        G.sync();
        w = vol(&S.ctl->work);
    }
}

// ---------------------------------------------------------------------------------------------
// Sequential boundary of the permutation test (GetBoundary.cs:19-160).  Host-side table, a function of
// (n_perm, alpha, eta) only; the reference computes it once per CanvasPartition run as well.
// ---------------------------------------------------------------------------------------------
double h_lchoose(int n, int k) {
    if (k < 0 || k > n) return -INFINITY;
    return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0);
}

// lower tail of the hypergeometric distribution: at most k of the n1s marked items among the first i of n_perm
double h_phyper(int k, int marked, int unmarked, int draws) {
    const int lo = std::max(0, draws - unmarked), hi = std::min(draws, marked);
    if (k < lo) return 0.0;
    if (k >= hi) return 1.0;
    double term = exp(h_lchoose(marked, lo) + h_lchoose(unmarked, draws - lo) - h_lchoose(marked + unmarked, draws));
    double sum = term;
    for (int x = lo; x < k; x++) {
        term *= (double)(marked - x) * (double)(draws - x) / ((double)(x + 1) * (double)(unmarked - draws + x + 1));
        sum += term;
    }
    return std::min(sum, 1.0);
}

void h_eta_row(unsigned n_perm, double eta0, unsigned ones, std::vector<unsigned>& sb, unsigned off) {
    unsigned k = 0;
    for (unsigned i = 1; i <= n_perm; i++)
        if (h_phyper((int)k, (int)ones, (int)(n_perm - ones), (int)i) <= eta0) sb[off + k++] = i;
}

double h_p_exceed(unsigned n_perm, unsigned ones, const std::vector<unsigned>& sb, unsigned off) {
    const double dl = h_lchoose((int)n_perm, (int)ones);
    auto at = [&](int i) { return (int)sb[off + i]; };
    double p = exp(h_lchoose((int)n_perm - at(0), (int)ones) - dl);
    if (ones >= 2) p += exp(log((double)at(0)) + h_lchoose((int)n_perm - at(1), (int)ones - 1) - dl);
    if (ones >= 3) {
        const double c = h_lchoose((int)n_perm - at(2), (int)ones - 2) - dl;
        p += exp(log((double)at(0)) + log(at(0) - 1.0) - log(2.0) + c) + exp(log((double)at(0)) + log((double)(at(1) - at(0))) + c);
    }
    for (int i = 4; i <= (int)ones; i++) {
        const int a = at(i - 4), b = at(i - 3), d = at(i - 2);
        const double c = h_lchoose((int)n_perm - at(i - 1), (int)ones - i + 1) - dl;
        p += exp(h_lchoose(a, i - 1) + c) + exp(h_lchoose(a, i - 2) + log((double)(d - a)) + c) +
             exp(h_lchoose(a, i - 3) + log((double)(b - a)) + log((double)(d - b)) + c) +
             exp(h_lchoose(a, i - 3) + log((double)(b - a)) - log(2.0) + log(b - a - 1.0) + c);
    }
    return p;
}

std::vector<unsigned> h_boundary_compute(unsigned n_perm, double alpha, double eta);

// the table only depends on its three parameters: computed once per process
const std::vector<unsigned>& h_boundary(unsigned n_perm, double alpha, double eta) {
    static std::mutex mu;
    static std::map<std::tuple<unsigned, double, double>, std::vector<unsigned>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(n_perm, alpha, eta);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, h_boundary_compute(n_perm, alpha, eta)).first;
    return it->second;
}

std::vector<unsigned> h_boundary_compute(unsigned n_perm, double alpha, double eta) {
    const unsigned max_ones = (unsigned)(floor(n_perm * alpha) + 1);
    std::vector<unsigned> sb((size_t)max_ones * (max_ones + 1) / 2, 0u);
    sb[0] = n_perm - (unsigned)(n_perm * eta);
    double eta0 = eta;
    unsigned l = 0;
    for (unsigned j = 2; j <= max_ones; j++) {
        double hi = eta0 * 1.1;
        h_eta_row(n_perm, hi, j, sb, l + 1);
        double p_hi = h_p_exceed(n_perm, j, sb, l + 1);
        double lo = eta0 * 0.25;
        h_eta_row(n_perm, lo, j, sb, l + 1);
        double p_lo = h_p_exceed(n_perm, j, sb, l + 1);
        while ((hi - lo) / lo > 1e-2) {
            eta0 = lo + (hi - lo) * (eta - p_lo) / (p_hi - p_lo);
            h_eta_row(n_perm, eta0, j, sb, l + 1);
            const double p = h_p_exceed(n_perm, j, sb, l + 1);
            if (p > eta) { hi = eta0; p_hi = p; } else { lo = eta0; p_lo = p; }
        }
        l += j;
    }
    return sb;
}

// ---------------------------------------------------------------------------------------------
// -s SDUndo (ChangePoint.cs:155-196, :423-474): host-side post-processing of the segment list, as in the
// reference (a handful of medians per chromosome; the trimmed SD is one sort of the genome's |differences|).
// ---------------------------------------------------------------------------------------------
double h_std_normal_quantile(double p) {
    double lo = -40.0, hi = 40.0;  // bisection on Phi, then it is exact to the last bits
    for (int it = 0; it < 200; it++) {
        const double mid = 0.5 * (lo + hi);
        if (0.5 * erfc(-mid / sqrt(2.0)) < p) lo = mid; else hi = mid;
    }
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 4; it++) x -= (0.5 * erfc(-x / sqrt(2.0)) - p) / (exp(-0.5 * x * x) / sqrt(2.0 * M_PI));
    return x;
}

double h_trimmed_sd(int n_chrom, const int64_t* off, const double* cov, double trim) {
    std::vector<double> d;
    d.reserve((size_t)std::max<int64_t>(off[n_chrom] - 1, 0));
    bool have = false;
    double last = 0.0;
    for (int c = 0; c < n_chrom; c++)
        for (int64_t i = off[c]; i < off[c + 1]; i++) {
            if (!std::isfinite(cov[i])) continue;
            if (have) d.push_back(fabs(cov[i] - last));
            last = cov[i];
            have = true;
        }
    const long keep = (long)nearbyint((1 - 2 * trim) * (double)d.size());
    std::sort(d.begin(), d.end());
    double ss = 0.0;
    for (long i = 0; i < keep && i < (long)d.size(); i++) ss += d[i] * d[i];
    // inflation factor: 1 / E[X^2] of N(0,1) truncated to its central 1 - 2 trim, midpoint rule on 10000 points
    const double a = h_std_normal_quantile(1 - trim);
    const double step = 2 * a / 10000, from = -a + step / 2, to = a - step / 2, inc = (to - from) / 9999;
    double e = 0.0, x = from;
    for (int i = 0; i < 10000; i++) {
        const double xi = i == 9999 ? to : x;
        e += (xi * xi) * (exp(-0.5 * xi * xi) / sqrt(2.0 * M_PI));
        x += inc;
    }
    e = e * step / (1 - 2 * trim);
    return sqrt((1 / e) * ss / (2 * keep));
}

double h_segment_median(const double* g, int a, int b) {
    std::vector<double> y(g + a, g + b);
    const size_t mid = y.size() / 2;
    std::nth_element(y.begin(), y.begin() + mid, y.end());
    if (y.size() & 1) return y[mid];
    return (y[mid] + *std::max_element(y.begin(), y.begin() + mid)) / 2;
}

// merge the two neighbours whose medians are closest while that distance is below the threshold
void h_sd_undo(const double* g, std::vector<int>& len, double threshold) {
    std::vector<int> end;
    int at = 0;
    for (int l : len) { at += l; end.push_back(at); }
    std::vector<double> med(end.size());
    for (size_t i = 0; i < end.size(); i++) med[i] = h_segment_median(g, i ? end[i - 1] : 0, end[i]);
    while (end.size() > 1) {
        size_t best = 0;
        double mn = fabs(med[1] - med[0]);
        for (size_t i = 1; i + 1 < end.size(); i++) {
            const double dv = fabs(med[i + 1] - med[i]);
            if (dv < mn) { mn = dv; best = i; }
        }
        if (!(mn < threshold)) break;
        end.erase(end.begin() + (long)best);
        med.erase(med.begin() + (long)best);
        med[best] = h_segment_median(g, best ? end[best - 1] : 0, end[best]);
    }
    len.clear();
    int prev = 0;
    for (int e : end) { len.push_back(e - prev); prev = e; }
}

// ---------------------------------------------------------------------------------------------
// -s Prune (ChangePoint.cs:205-271, Prune.cs:18-76): host-side post-processing like SDUndo.  The reference scores
// every j-subset of the K change points for j = K-1 .. 1 from scratch (O(K) per subset).  Here the term of every
// possible merged group [a, b) of segments, (sum of the segments' sums, added left to right)^2 / length, is
// tabulated once, and the subsets are walked depth-first in the same lexicographic order carrying the partial sum
// of terms, which is the reference's own left-to-right sum: the same doubles, O(1) amortised per subset.
// `<=` keeps the LAST of equally good subsets.  Returns false when more than `budget` subsets would be scored
// (the search is exponential in K; the reference does not terminate in practice there either).
// ---------------------------------------------------------------------------------------------
constexpr long long CG_PRUNE_BUDGET = 1LL << 28;
bool h_prune_undo(const double* g, int n, std::vector<int>& len, double cutoff, long long budget, long long* scored) {
    if (budget <= 0) budget = CG_PRUNE_BUDGET;
    const int S = (int)len.size(), K = S - 1;  // segments, change points
    if (K > 4096) return false;
    std::vector<double> sx(S);
    double ssq = 0.0;
    for (int i = 0; i < n; i++) ssq += pow(g[i], 2);
    int at = 0;
    for (int i = 0; i < S; i++) {
        double s = 0.0;
        for (int p2 = at; p2 < at + len[i]; p2++) s += g[p2];
        sx[i] = s;
        at += len[i];
    }
    // term[a * (S + 1) + b] for 0 <= a < b <= S
    std::vector<double> term((size_t)(S + 1) * (S + 1), 0.0);
    for (int a = 0; a < S; a++) {
        double s = 0.0;
        int cnt = 0;
        for (int b = a + 1; b <= S; b++) {
            s += sx[b - 1];
            cnt += len[b - 1];
            term[(size_t)a * (S + 1) + b] = pow(s, 2) / cnt;
        }
    }
    auto T = [&](int a, int b) { return term[(size_t)a * (S + 1) + b]; };
    std::vector<double> tail(S + 1);  // the last group [l, S)
    for (int l = 0; l < S; l++) tail[l] = T(l, S);
    double full = 0.0;  // all K change points kept
    for (int i = 0; i < S; i++) full += T(i, i + 1);
    const double wssqk = ssq - full;
    std::vector<int> loc(K > 0 ? K : 1), best(K > 0 ? K : 1), best_prev(K > 0 ? K : 1);
    std::vector<double> part(K + 1);
    for (int i = 0; i < K; i++) best_prev[i] = i + 1;
    int pruned = 0;
    long long count = 0;
    for (int j = K - 1; j > 0; j--) {
        // depth-first over loc[0] < loc[1] < ... < loc[j-1] in 1..K; loc[d] <= K - (j - 1 - d); the last level is
        // a plain loop over one row of the table
        double wssqj = 0.0;
        bool first = true;
        int d = 0;
        loc[0] = 1;
        while (d >= 0) {
            if (d < j - 1) {
                if (loc[d] > K - (j - 1 - d)) {  // level exhausted
                    d--;
                    if (d >= 0) loc[d]++;
                    continue;
                }
                part[d] = (d == 0 ? 0.0 + T(0, loc[0]) : part[d - 1] + T(loc[d - 1], loc[d]));
                loc[d + 1] = loc[d] + 1;
                d++;
                continue;
            }
            const int from = loc[d], prev = d == 0 ? 0 : loc[d - 1];
            const double base = d == 0 ? 0.0 : part[d - 1];
            const double* row = &term[(size_t)prev * (S + 1)];
            int arg = -1;
            for (int l = from; l <= K; l++) {
                const double w1 = ssq - ((base + row[l]) + tail[l]);
                if (first || w1 <= wssqj) { first = false; wssqj = w1; arg = l; }
            }
            if (arg >= 0) {
                for (int i = 0; i < d; i++) best[i] = loc[i];
                best[d] = arg;
            }
            count += K - from + 1;
            if (count > budget) return false;
            d--;
            if (d >= 0) loc[d]++;
        }
        if (wssqj / wssqk > 1 + cutoff) {
            pruned = j + 1;
            break;
        }
        for (int i = 0; i < j; i++) best_prev[i] = best[i];
    }
    if (scored) *scored = count;
    std::vector<int> cum(S);
    cum[0] = len[0];
    for (int i = 1; i < S; i++) cum[i] = cum[i - 1] + len[i];
    std::vector<int> out;
    int prev = 0;
    for (int i = 0; i < pruned; i++) {
        const int e = cum[best_prev[i] - 1];
        out.push_back(e - prev);
        prev = e;
    }
    out.push_back(n - prev);
    len.swap(out);
    return true;
}

unsigned mt_first_outputs(unsigned seed, int count, std::vector<unsigned>& out) {
    unsigned s[624];
    s[0] = seed;
    for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (unsigned)i;
    out.clear();
    int at = 624;
    for (int c = 0; c < count; c++) {
        if (at >= 624) {
            for (int k = 0; k < 624; k++) {
                const unsigned y = (s[k] & 0x80000000u) | (s[(k + 1) % 624] & 0x7fffffffu);
                s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            at = 0;
        }
        unsigned y = s[at++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        out.push_back(y);
    }
    return out.empty() ? 0u : out.back();
}

}  // namespace

extern "C" int cg_cbs_prune(const double* g, int64_t n, const int32_t* seg_len, int n_seg, double cutoff, int64_t max_subsets,
                            int32_t* seg_len_out, int64_t* subsets_scored) {
    if (!g || !seg_len || !seg_len_out || n_seg < 2 || n <= 0 || n > 0x3fffffffLL) return CG_ERR_ARG;
    int64_t total = 0;
    for (int i = 0; i < n_seg; i++) {
        if (seg_len[i] <= 0) return CG_ERR_ARG;
        total += seg_len[i];
    }
    if (total != n) return CG_ERR_ARG;
    std::vector<int> len(seg_len, seg_len + n_seg);
    long long scored = 0;
    if (!h_prune_undo(g, (int)n, len, cutoff, max_subsets, &scored)) return CG_ERR_UNSUPPORTED;
    for (size_t i = 0; i < len.size(); i++) seg_len_out[i] = len[i];
    if (subsets_scored) *subsets_scored = scored;
    return (int)len.size();
}

extern "C" int64_t cg_cbs_boundary(uint32_t n_perm, double alpha, double eta, uint32_t* out, int64_t cap) {
    if (n_perm == 0 || !(alpha > 0) || !(eta > 0)) return -1;
    const std::vector<unsigned>& sb = h_boundary(n_perm, alpha, eta);
    if (out)
        for (size_t i = 0; i < sb.size() && (int64_t)i < cap; i++) out[i] = sb[i];
    return (int64_t)sb.size();
}

static int partition_cbs_impl(cg_ctx* ctx, const cg_cbs_opts* o, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                              const int64_t* chrom_off, const double* coverage, const uint8_t* chrom_selected, int32_t* n_seg,
                              int32_t* seg_len, double* seg_mean, int64_t* stats) {
    if (!ctx) return CG_ERR_ARG;
    if (!o || n_chrom < 0 || (n_chrom > 0 && (!chrom_off || !n_seg))) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: bad argument");
    if (o->undo < 0 || o->undo > 2) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: undo must be 0 (none), 1 (prune) or 2 (sdundo)");
    if (!o->hybrid) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_cbs: only the hybrid p-value method (the one CanvasPartition uses) is supported");
    if (o->min_width < 2 || o->min_width > 5) return cg_fail(ctx, CG_ERR_ARG, "Minimum segment width should be between 2 and 5");
    if (o->n_min < 4u * (unsigned)o->k_max) return cg_fail(ctx, CG_ERR_ARG, "nMin should be >= 4 * kMax");
    if (o->k_max > CBS_KMAX || o->k_max < o->min_width || o->n_min > (unsigned)CBS_NMIN_MAX)
        return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_cbs: k_max <= 32 and n_min <= 256 are supported");
    if (o->n_perm == 0 || !(o->alpha > 0)) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: n_perm and alpha must be positive");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    std::vector<unsigned> table;
    if (!sbdry) {
        table = h_boundary(o->n_perm, o->alpha, o->eta > 0 ? o->eta : 0.05);
        sbdry = table.data();
        n_sbdry = (int64_t)table.size();
    }
    {
        const uint64_t max_ones = (uint64_t)(floor(o->n_perm * o->alpha) + 1);
        if ((uint64_t)n_sbdry < max_ones * (max_ones + 1) / 2) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: boundary table too short");
    }
    const int64_t N = n_chrom ? chrom_off[n_chrom] : 0;
    for (int c = 0; c < n_chrom; c++) {
        if (chrom_off[c + 1] < chrom_off[c]) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: chrom_off must be non-decreasing");
        if (chrom_off[c + 1] - chrom_off[c] > 0x3fffffffLL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_cbs: chromosome too long");
        n_seg[c] = 0;
    }
    if (N == 0) return CG_OK;
    if (!coverage || !seg_len || !seg_mean) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs: null array");
    for (int c = 0; c < n_chrom; c++)
        for (int64_t i = chrom_off[c]; i < chrom_off[c + 1] && (!chrom_selected || chrom_selected[c]); i++)
        if (!std::isfinite(coverage[i]))
            return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_cbs: non-finite coverage (the reference feeds it to ChangePoints unfiltered, CBSRunner.cs:118)");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));

    // per-chromosome seeds: MersenneTwister(0).NextFullRangeInt32() in CoverageByChr order (CBSRunner.cs:107-112)
    std::vector<unsigned> seeds;
    mt_first_outputs(o->seed, n_chrom, seeds);
    std::vector<CbsWork> work;
    for (int c = 0; c < n_chrom; c++) {
        const int n = (int)(chrom_off[c + 1] - chrom_off[c]);
        if (n > 0 && (!chrom_selected || chrom_selected[c])) work.push_back(CbsWork{chrom_off[c], n, c, seeds[c], 0});
    }
    std::stable_sort(work.begin(), work.end(), [](const CbsWork& a, const CbsWork& b) { return a.n > b.n; });
    const int nwork = (int)work.size();
    if (nwork == 0) return CG_OK;
    int G = 8;
    while (G > 1 && nwork * G > ctx->num_sms) G >>= 1;
    int nclusters = std::min(nwork, std::max(1, ctx->num_sms / G));
    const int batch = CBS_PT * G;

    // scratch per cluster, sized for its first (largest) chromosome
    struct Sizes { size_t n, nb, nb2, ring; };
    std::vector<Sizes> sz(nclusters);
    size_t need = arena_need(N, 8) + arena_need(nwork, sizeof(CbsWork)) + arena_need(n_sbdry, 4) + arena_need(nclusters, sizeof(CbsScratch)) +
                  arena_need(n_chrom, 4) + arena_need(N, 4) + arena_need(N, 8) + arena_need((size_t)n_chrom * 4, 8) + arena_need((size_t)n_chrom * 8, 8) + 4096;
    for (int k = 0; k < nclusters; k++) {
        const size_t n = (size_t)work[k].n;
        size_t nb = (size_t)std::llround(std::sqrt((double)n)) + 2;
        size_t ring = 1024;
        while (ring < (size_t)batch * n + 4 * 624) ring <<= 1;
        sz[k] = Sizes{n, nb, nb * (nb + 1) / 2 + 1, ring};
        need += arena_need(n, 8) * 2 + arena_need((size_t)batch * n, 8) + arena_need(batch, 8) + arena_need(nb, 8) * 2 + arena_need(n + 2, 4) * 3 +
                arena_need(nb, 4) * 3 + arena_need(sz[k].nb2, sizeof(CbsCand)) + arena_need(sz[k].nb2, sizeof(CbsBest)) + arena_need(ring, 4) +
                arena_need(1, sizeof(CbsCtl)) + 8192;
    }
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    double* d_cov = arena_take<double>(ctx, N);
    CbsWork* d_work = arena_take<CbsWork>(ctx, nwork);
    unsigned* d_sb = arena_take<unsigned>(ctx, n_sbdry);
    CbsScratch* d_scr = arena_take<CbsScratch>(ctx, nclusters);
    int* d_nseg = arena_take<int>(ctx, n_chrom);
    int* d_len = arena_take<int>(ctx, N);
    double* d_mean = arena_take<double>(ctx, N);
    long long* d_stats = arena_take<long long>(ctx, (size_t)n_chrom * 4);
    int* d_queue = arena_take<int>(ctx, 64);
    long long* d_phase = arena_take<long long>(ctx, (size_t)n_chrom * 8);
    std::vector<CbsScratch> scr(nclusters);
    bool ok = d_cov && d_work && d_sb && d_scr && d_nseg && d_len && d_mean && d_stats && d_queue && d_phase;
    for (int k = 0; k < nclusters && ok; k++) {
        CbsScratch& s = scr[k];
        const Sizes& z = sz[k];
        s.cur = arena_take<double>(ctx, z.n);
        s.sx = arena_take<double>(ctx, z.n);
        s.px = arena_take<double>(ctx, (size_t)batch * z.n);
        s.pstat = arena_take<double>(ctx, batch);
        s.pmin = arena_take<double>(ctx, z.nb);
        s.pmax = arena_take<double>(ctx, z.nb);
        s.stack = arena_take<int>(ctx, z.n + 2);
        s.locs = arena_take<int>(ctx, z.n + 2);
        s.arr = arena_take<int>(ctx, z.n + 2);
        s.bb = arena_take<int>(ctx, z.nb);
        s.imin = arena_take<int>(ctx, z.nb);
        s.imax = arena_take<int>(ctx, z.nb);
        s.cand = arena_take<CbsCand>(ctx, z.nb2);
        s.cbest = arena_take<CbsBest>(ctx, z.nb2);
        s.ring = arena_take<unsigned>(ctx, z.ring);
        s.ctl = arena_take<CbsCtl>(ctx, 1);
        s.px_bytes = (long long)batch * (long long)z.n * 8;
        s.ring_mask = (unsigned)(z.ring - 1);
        s.n_alloc = (int)z.n;
        ok = s.cur && s.sx && s.px && s.pstat && s.pmin && s.pmax && s.stack && s.locs && s.arr && s.bb && s.imin && s.imax && s.cand && s.cbest &&
             s.ring && s.ctl;
    }
    if (!ok) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_cbs: device arena exhausted");
    cudaStream_t st = ctx->stream;
    CG_CUDA(ctx, cudaMemcpyAsync(d_cov, coverage, N * 8, cudaMemcpyHostToDevice, st));
    CG_CUDA(ctx, cudaMemcpyAsync(d_work, work.data(), nwork * sizeof(CbsWork), cudaMemcpyHostToDevice, st));
    CG_CUDA(ctx, cudaMemcpyAsync(d_sb, sbdry, n_sbdry * 4, cudaMemcpyHostToDevice, st));
    CG_CUDA(ctx, cudaMemcpyAsync(d_scr, scr.data(), nclusters * sizeof(CbsScratch), cudaMemcpyHostToDevice, st));
    CG_CUDA(ctx, cudaMemsetAsync(d_nseg, 0, n_chrom * 4, st));
    CG_CUDA(ctx, cudaMemsetAsync(d_stats, 0, (size_t)n_chrom * 32, st));
    CG_CUDA(ctx, cudaMemsetAsync(d_queue, 0, 256, st));
    CG_CUDA(ctx, cudaMemsetAsync(d_phase, 0, (size_t)n_chrom * 64, st));
    CbsParams P;
    P.o = CbsOpts{o->alpha, o->n_perm, o->min_width, o->k_max, o->n_min};
    P.cov = d_cov; P.work = d_work; P.nwork = nwork; P.queue = d_queue; P.sbdry = d_sb; P.scratch = d_scr;
    P.n_seg = d_nseg; P.seg_len = d_len; P.seg_mean = d_mean; P.stats = d_stats; P.phase_ns = d_phase;
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nclusters * G);
        cfg.blockDim = dim3(CBS_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, cbs_kernel, P);
        ctx->launches++;
        if (le != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, std::string("cbs_kernel launch: ") + cudaGetErrorString(le));
    }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    std::vector<long long> h_stats((size_t)n_chrom * 4);
    CG_CUDA(ctx, cudaMemcpyAsync(n_seg, d_nseg, n_chrom * 4, cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaMemcpyAsync(seg_len, d_len, N * 4, cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaMemcpyAsync(seg_mean, d_mean, N * 8, cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaMemcpyAsync(h_stats.data(), d_stats, (size_t)n_chrom * 32, cudaMemcpyDeviceToHost, st));
    std::vector<long long> h_phase((size_t)n_chrom * 8);
    CG_CUDA(ctx, cudaMemcpyAsync(h_phase.data(), d_phase, (size_t)n_chrom * 64, cudaMemcpyDeviceToHost, st));
    CG_CUDA(ctx, cudaStreamSynchronize(st));
    CG_CUDA(ctx, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    {
        // phase times (ms) of the chromosome that took longest: cg_last_partition_stats slots 0..6
        int worst = 0;
        for (int c = 0; c < n_chrom; c++)
            if (h_phase[(size_t)c * 8 + 6] > h_phase[(size_t)worst * 8 + 6]) worst = c;
        for (int k = 0; k < 8; k++) ctx->stats[k] = (double)h_phase[(size_t)worst * 8 + k] * 1e-6;
        ctx->stats[8] = worst;
    }
    if (stats) {
        stats[0] = stats[1] = stats[2] = stats[3] = 0;
        for (int c = 0; c < n_chrom; c++)
            for (int k = 0; k < 4; k++) stats[k] += h_stats[(size_t)c * 4 + k];
    }
    if (o->undo == 1 || o->undo == 2) {
        const double threshold = o->undo == 2 ? o->undo_sd * h_trimmed_sd(n_chrom, chrom_off, coverage, o->trim) : 0.0;
        for (int c = 0; c < n_chrom; c++) {
            if (n_seg[c] <= 1) continue;
            const double* g = coverage + chrom_off[c];
            std::vector<int> len(seg_len + chrom_off[c], seg_len + chrom_off[c] + n_seg[c]);
            if (o->undo == 2) h_sd_undo(g, len, threshold);
            else if (!h_prune_undo(g, (int)(chrom_off[c + 1] - chrom_off[c]), len, o->undo_prune, CG_PRUNE_BUDGET, nullptr))
                return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_cbs: the prune undo would score more than 2^28 change-point subsets on one chromosome (the search is exponential in the number of change points)");
            int at = 0;
            for (size_t i = 0; i < len.size(); i++) {
                double sum = 0.0, w = 0.0;
                for (int p2 = at; p2 < at + len[i]; p2++) { w += 1.0; sum += g[p2] * 1.0; }
                seg_len[chrom_off[c] + (int64_t)i] = len[i];
                seg_mean[chrom_off[c] + (int64_t)i] = sum / w;
                at += len[i];
            }
            n_seg[c] = (int)len.size();
        }
    }
    return CG_OK;
}

extern "C" int cg_partition_cbs(cg_ctx* ctx, const cg_cbs_opts* o, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                                const int64_t* chrom_off, const double* coverage, int32_t* n_seg, int32_t* seg_len, double* seg_mean,
                                int64_t* stats) {
    return partition_cbs_impl(ctx, o, sbdry, n_sbdry, n_chrom, chrom_off, coverage, nullptr, n_seg, seg_len, seg_mean, stats);
}

extern "C" int cg_partition_cbs_shard(cg_ctx* ctx, const cg_cbs_opts* o, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                                      const int64_t* chrom_off, const double* coverage, const uint8_t* chrom_selected, int32_t* n_seg,
                                      int32_t* seg_len, double* seg_mean, int64_t* stats) {
    if (!chrom_selected && n_chrom > 0) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_cbs_shard: chrom_selected is required");
    return partition_cbs_impl(ctx, o, sbdry, n_sbdry, n_chrom, chrom_off, coverage, chrom_selected, n_seg, seg_len, seg_mean, stats);
}
