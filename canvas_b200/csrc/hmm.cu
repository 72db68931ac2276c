// cg_partition_hmm: CanvasPartition -m HMM / -m PerSampleHMM on the device (reference HiddenMarkovModelsRunner.cs:23-109,
// HMM.cs:62-130, Distributions.cs:206-323).
//
// The reference runs one sequential 5-state Viterbi pass per chromosome.  Here the pass is cut into blocks of
// HMM_BLOCK bins and made parallel in three steps:
//   A  every block folds its bins into one 5x5 (max,+) transfer matrix                 (thread per block)
//   B  per chromosome the matrices are applied in order to the start vector, which
//      gives the score vector at the start of every block                               (thread per chromosome)
//   C  every block replays the reference recurrence bin by bin from its start vector —
//      the same additions in the same order, first index on ties — and records the
//      back pointers and the block's end-state -> previous-block-end-state map          (thread per block)
// Back tracking is the same two-level walk (per chromosome over block maps in shared memory, then per block over
// its back pointers), and breakpoints are the bins whose state differs from the previous bin's.
// Step B re-associates the additions, so a block's start vector can differ from the sequential one in the last
// ulps; decisions inside a block only see differences of scores, which are ~1e-10 apart at most.  Exact-equal
// scores (states 0/1 and 3/4 share their emissions in joint mode) stay exactly equal.  Degenerate inputs (a score
// reaching Double.MinValue) are detected and rerun by the strictly sequential kernel, which is also selectable
// (cg_hmm_opts.exact_sequential) as a cross-check.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "clean.cuh"
#include "select.cuh"

namespace {

constexpr int HMM_NS = 5;
constexpr int HMM_BLOCK = 64;    // bins per block (thread granularity of steps A and C)
constexpr int HMM_GROUP = 64;    // blocks per group (second level of the scan and of the back tracking)
constexpr int HMM_MAX_SAMPLES = 4;
constexpr int HMM_MAX_CHROM = 256;

struct HmmBlk {
    long long t0, t1;  // global bin range of the block (the first bin of a chromosome belongs to no block)
    int chrom;
    int last;          // last block of its chromosome
};

struct HmmGrp {
    int b0, b1;        // block range of the group
    int chrom;
    int pad;
};

struct HmmChromInfo {
    long long a, b;    // global bin range
    int first_blk, n_blk;
    int first_grp, n_grp;
    int active;        // selected and longer than min_size
    int tab;           // emission table index
    double max_thr;    // RemoveOutliers threshold
};

struct HmmCtl {
    int degenerate;    // a score reached Double.MinValue / every state had zero emission
    int bad_value;     // coverage outside the emission table (negative)
};

// ---------------------------------------------------------------------------------------------------------------
// statistics for the emission model
// ---------------------------------------------------------------------------------------------------------------
struct QuartView {  // whole-genome quartiles per sample in single precision (HiddenMarkovModelsRunner.cs:38-50)
    const double* cov;
    long long N;
    int S;
    __device__ long long size() const { return N * S; }
    __device__ bool get(long long i, uint32_t& key, int& a, int& b) const {
        key = f32_key((float)cov[i]);
        a = (int)(i / N);
        b = -1;
        return true;
    }
};

struct ChromMedView {  // per-chromosome median per sample (:118)
    const double* cov;
    const uint8_t* chrom_id;
    long long N;
    int S, C;
    __device__ long long size() const { return N * S; }
    __device__ bool get(long long i, uint64_t& key, int& a, int& b) const {
        const long long t = i % N;
        key = f64_key(cov[i]);
        a = (int)(i / N) * C + chrom_id[t];
        b = -1;
        return true;
    }
};

// ranks of Utilities.Quartiles (Utilities.cs:361-419); slots: 0,1 = Q2 pair, 2,3 = Q1 pair, 4,5 = Q3 pair
__global__ void hmm_quartile_request_kernel(SelState<uint32_t> st, long long n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    unsigned long long* k = st.req_k + (size_t)s * SEL_G;
    if (n < 2) { st.nreq[s] = 0; return; }
    const long long mid = n / 2;
    if (n % 2 == 0) {
        k[0] = mid - 1; k[1] = mid;
        const long long mm = mid / 2;
        if (mid % 2 == 0) { k[2] = mm - 1; k[3] = mm; k[4] = mid + mm - 1; k[5] = mid + mm; }
        else { k[2] = mm; k[3] = mm; k[4] = mm + mid; k[5] = mm + mid; }
    } else {
        k[0] = mid; k[1] = mid;
        if ((n - 1) % 4 == 0) { const long long q = (n - 1) / 4; k[2] = q - 1; k[3] = q; k[4] = 3 * q; k[5] = 3 * q + 1; }
        else { const long long q = (n - 3) / 4; k[2] = q; k[3] = q + 1; k[4] = 3 * q + 1; k[5] = 3 * q + 2; }
    }
    st.nreq[s] = 6;
}

__global__ void hmm_median_request_kernel(SelState<uint64_t> st, const HmmChromInfo* __restrict__ ci, int C) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= st.nseg) return;
    const long long n = ci[seg % C].b - ci[seg % C].a;
    if (n <= 0) { st.nreq[seg] = 0; return; }
    st.req_k[(size_t)seg * SEL_G + 0] = (unsigned long long)((n - 1) / 2);
    st.req_k[(size_t)seg * SEL_G + 1] = (unsigned long long)(n / 2);
    st.nreq[seg] = 2;
}

// block and group tables of one chromosome from its range (one CTA per chromosome)
__global__ void hmm_tables_kernel(const HmmChromInfo* __restrict__ ci, HmmBlk* __restrict__ blk, HmmGrp* __restrict__ grp) {
    const int c = blockIdx.x;
    const HmmChromInfo h = ci[c];
    for (int k = threadIdx.x; k < h.n_blk; k += blockDim.x) {
        HmmBlk b;
        b.t0 = h.a + 1 + (long long)k * HMM_BLOCK;
        b.t1 = b.t0 + HMM_BLOCK < h.b ? b.t0 + HMM_BLOCK : h.b;
        b.chrom = c;
        b.last = b.t1 == h.b;
        blk[h.first_blk + k] = b;
    }
    for (int k = threadIdx.x; k < h.n_grp; k += blockDim.x) {
        HmmGrp g;
        g.b0 = h.first_blk + k * HMM_GROUP;
        g.b1 = g.b0 + HMM_GROUP < h.first_blk + h.n_blk ? g.b0 + HMM_GROUP : h.first_blk + h.n_blk;
        g.chrom = c;
        g.pad = 0;
        grp[h.first_grp + k] = g;
    }
}

__global__ void hmm_chrom_id_kernel(const HmmChromInfo* __restrict__ ci, int C, long long N, uint8_t* __restrict__ chrom_id) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = C - 1;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (t >= ci[m].b) lo = m + 1; else hi = m;
        }
        chrom_id[t] = (uint8_t)lo;
    }
}

// Utilities.Variance (Utilities.cs:290-302) per (sample, chromosome): mean, then sum of squared deviations.
// The reference adds left to right; a block-wide tree differs from that in the last ulps only.
__global__ void __launch_bounds__(1024) hmm_variance_kernel(const double* __restrict__ cov, long long N, const HmmChromInfo* __restrict__ ci,
                                                            int C, double* __restrict__ var_out) {
    __shared__ double red[32];
    __shared__ double s_mu;
    const int c = blockIdx.x % C, s = blockIdx.x / C;
    const long long a = ci[c].a, n = ci[c].b - ci[c].a;
    const double* x = cov + (size_t)s * N + a;
    auto block_sum = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        }
        return v;  // valid in thread 0
    };
    double acc = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) s_mu = n > 0 ? acc / (double)n : 0.0;
    __syncthreads();
    const double mu = s_mu;
    acc = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double d = x[i] - mu; acc += d * d; }
    acc = block_sum(acc);
    if (threadIdx.x == 0) var_out[blockIdx.x] = acc / (double)(n - 1);
}

// ---------------------------------------------------------------------------------------------------------------
// emissions: le[t][j] = log(max over genotype arrangements of the product over samples)   (Distributions.cs:257-296)
// S == 1: the host has taken the logarithm already (tab = log E[tab][j][x]); S > 1: tab = P[tab][g][s][x]
// ---------------------------------------------------------------------------------------------------------------
__global__ void hmm_emission_kernel(const double* __restrict__ cov, long long N, int S, const uint8_t* __restrict__ chrom_id,
                                    const HmmChromInfo* __restrict__ ci, const double* __restrict__ tab, int L, int use_all_states,
                                    double* __restrict__ le, HmmCtl* ctl) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const HmmChromInfo c = ci[chrom_id[t]];
        if (!c.active) continue;
        int x[HMM_MAX_SAMPLES];
        bool bad = false;
        for (int s = 0; s < S; s++) {
            double v = cov[(size_t)s * N + t];
            if (!(v >= 0.0 && v < INFINITY)) bad = true;  // NaN / negative: Convert.ToInt32 or the table index throw in the reference
            v = v > c.max_thr ? c.max_thr : v;          // RemoveOutliers (HiddenMarkovModelsRunner.cs:155-163)
            x[s] = __double2int_rn(v);                   // Convert.ToInt32: half to even
            if (x[s] < 0 || x[s] >= L) { bad = true; x[s] = 0; }
        }
        if (bad) ctl->bad_value = 1;
        double out[HMM_NS];
        if (S == 1) {
            const double* tb = tab + (size_t)c.tab * HMM_NS * L;
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) out[j] = tb[(size_t)j * L + x[0]];
        } else {
            const double* tb = tab + (size_t)c.tab * HMM_NS * S * L;
            double p[HMM_NS][HMM_MAX_SAMPLES];  // factor of genotype g for sample s
            for (int s = 0; s < S; s++) {
                double q[HMM_NS];
#pragma unroll
                for (int g = 0; g < HMM_NS; g++) q[g] = tb[((size_t)g * S + s) * L + x[s]];
                if (use_all_states) {
#pragma unroll
                    for (int g = 0; g < HMM_NS; g++) p[g][s] = q[g];
                } else {
                    const double lo = fmax(q[0], q[1]), hi = fmax(q[3], q[4]);
                    p[0][s] = lo; p[1][s] = lo; p[2][s] = q[2]; p[3][s] = hi; p[4][s] = hi;
                }
            }
            const unsigned n_assign = 1u << S;
            for (int j = 0; j < HMM_NS; j++) {
                double best = -DBL_MAX;
                for (unsigned mask = 0; mask < n_assign; mask++) {  // bit s: sample s is diploid
                    if (j != 2 && mask == n_assign - 1) continue;
                    if (j == 2 && mask != n_assign - 1) continue;
                    double l = 1.0;
                    for (int s = 0; s < S; s++) l = __dmul_rn(l, ((mask >> s) & 1u) ? p[2][s] : p[j][s]);
                    if (isnan(l) || isinf(l)) l = 0;
                    if (best < l) best = l;
                }
                out[j] = log(best);
            }
        }
        bool all_zero = true;
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) {
            le[(size_t)t * HMM_NS + j] = out[j];
            if (out[j] > -INFINITY) all_zero = false;
        }
        if (all_zero) ctl->degenerate = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// step A: block transfer matrices.  A[i][j] = best score of a path that enters the block from state i (at the bin
// before the block) and is in state j at the block's last bin.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) hmm_block_kernel(const double* __restrict__ le, const HmmBlk* __restrict__ blk, int n_blk,
                                                        double ls, double lo, double* __restrict__ mats) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const HmmBlk bk = blk[b];
    double A[HMM_NS][HMM_NS];
    {
        const double* e = le + (size_t)bk.t0 * HMM_NS;
#pragma unroll
        for (int i = 0; i < HMM_NS; i++)
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) A[i][j] = fmax(-DBL_MAX, __dadd_rn(e[j], i == j ? ls : lo));
    }
    double en[HMM_NS];  // emissions of the next bin, loaded one step ahead
    if (bk.t0 + 1 < bk.t1) {
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) en[j] = le[(size_t)(bk.t0 + 1) * HMM_NS + j];
    }
    for (long long t = bk.t0 + 1; t < bk.t1; t++) {
        double e[HMM_NS];
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) e[j] = en[j];
        if (t + 1 < bk.t1) {
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) en[j] = le[(size_t)(t + 1) * HMM_NS + j];
        }
        double cs[HMM_NS], co[HMM_NS];
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) { cs[j] = __dadd_rn(e[j], ls); co[j] = __dadd_rn(e[j], lo); }
#pragma unroll
        for (int i = 0; i < HMM_NS; i++) {
            // largest and second largest entry of row i: max over k != j of A[i][k] in O(1); adding the same
            // constant is monotone, so the maximum of the sums is the sum of the maximum
            double m1 = A[i][0], m2 = -INFINITY;
            int k1 = 0;
#pragma unroll
            for (int k = 1; k < HMM_NS; k++) {
                const double v = A[i][k];
                if (v > m1) { m2 = m1; m1 = v; k1 = k; } else if (v > m2) m2 = v;
            }
            double nw[HMM_NS];
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) {
                const double other = (k1 == j) ? m2 : m1;
                nw[j] = fmax(-DBL_MAX, fmax(__dadd_rn(A[i][j], cs[j]), __dadd_rn(other, co[j])));
            }
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) A[i][j] = nw[j];
        }
    }
    double* out = mats + (size_t)b * 25;
#pragma unroll
    for (int i = 0; i < HMM_NS; i++)
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) out[i * HMM_NS + j] = A[i][j];
}

// first-bin scores (HMM.cs:75-82): log(pi_j) + (le_j + log T[0][j]) - log T[0][j]
__device__ inline void hmm_init_scores(const double* e, double log_start, double ls, double lo, double* s) {
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) {
        const double lt = j == 0 ? ls : lo;
        s[j] = __dsub_rn(__dadd_rn(log_start, __dadd_rn(e[j], lt)), lt);
    }
}

// first index of the maximum with the reference's strict comparison against Double.MinValue (HMM.cs:105-117)
__device__ inline int hmm_best_final(const double* s) {
    int best = -1;
    double mx = -DBL_MAX;
#pragma unroll
    for (int i = 0; i < HMM_NS; i++)
        if (s[i] > mx) { best = i; mx = s[i]; }
    return best < 0 ? 0 : best;
}

// ---------------------------------------------------------------------------------------------------------------
// step B: start vector of every block, in two levels.  B1: one warp per group multiplies the group's block
// matrices ((max,+) product, lane (i,j) owns one entry); B2: per chromosome the group matrices are applied in order
// from the first-bin scores (staged in shared memory, one thread walks them); B3: one thread per group walks its
// blocks from the group's start vector.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) hmm_group_kernel(const HmmGrp* __restrict__ grp, int n_grp, const double* __restrict__ mats,
                                                        double* __restrict__ gmats) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n_grp) return;
    const HmmGrp gr = grp[g];
    const int L = lane < 25 ? lane : 0;
    const int i = L / HMM_NS, j = L % HMM_NS;
    double P = mats[(size_t)gr.b0 * 25 + L];
    double nxt = gr.b0 + 1 < gr.b1 ? mats[(size_t)(gr.b0 + 1) * 25 + L] : 0.0;
    for (int b = gr.b0 + 1; b < gr.b1; b++) {
        const double M = nxt;
        if (b + 1 < gr.b1) nxt = mats[(size_t)(b + 1) * 25 + L];
        double acc = -DBL_MAX;
#pragma unroll
        for (int k = 0; k < HMM_NS; k++) {
            const double pk = __shfl_sync(0xffffffffu, P, i * HMM_NS + k);
            const double mk = __shfl_sync(0xffffffffu, M, k * HMM_NS + j);
            acc = fmax(acc, __dadd_rn(pk, mk));
        }
        P = acc;
    }
    if (lane < 25) gmats[(size_t)g * 25 + lane] = P;
}

__device__ inline void hmm_apply(double* s, const double* A) {  // s <- s (x) A with the reference's Double.MinValue floor
    double ns[HMM_NS];
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) {
        double mx = -DBL_MAX;
#pragma unroll
        for (int i = 0; i < HMM_NS; i++) mx = fmax(mx, __dadd_rn(s[i], A[i * HMM_NS + j]));
        ns[j] = mx;
    }
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) s[j] = ns[j];
}

__global__ void __launch_bounds__(128) hmm_scan_groups_kernel(const double* __restrict__ le, const HmmChromInfo* __restrict__ ci,
                                                              const double* __restrict__ gmats, double log_start, double ls, double lo,
                                                              double* __restrict__ gvec, int* __restrict__ end_state) {
    extern __shared__ double s_gm[];
    const HmmChromInfo ch = ci[blockIdx.x];
    if (!ch.active) return;
    for (int q = threadIdx.x; q < ch.n_grp * 25; q += blockDim.x) s_gm[q] = gmats[(size_t)ch.first_grp * 25 + q];
    __syncthreads();
    if (threadIdx.x != 0) return;
    double s[HMM_NS];
    hmm_init_scores(le + (size_t)ch.a * HMM_NS, log_start, ls, lo, s);
    if (ch.n_blk == 0) { end_state[blockIdx.x] = hmm_best_final(s); return; }
    for (int g = 0; g < ch.n_grp; g++) {
        double* gv = gvec + (size_t)(ch.first_grp + g) * HMM_NS;
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) gv[j] = s[j];
        hmm_apply(s, s_gm + g * 25);
    }
}

__global__ void __launch_bounds__(64) hmm_scan_blocks_kernel(const HmmGrp* __restrict__ grp, int n_grp, const double* __restrict__ mats,
                                                             const double* __restrict__ gvec, double* __restrict__ svec, HmmCtl* ctl) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grp) return;
    const HmmGrp gr = grp[g];
    double s[HMM_NS];
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) s[j] = gvec[(size_t)g * HMM_NS + j];
    double nxt[25];
#pragma unroll
    for (int q = 0; q < 25; q++) nxt[q] = mats[(size_t)gr.b0 * 25 + q];
    bool degenerate = false;
    for (int b = gr.b0; b < gr.b1; b++) {
        double A[25];
#pragma unroll
        for (int q = 0; q < 25; q++) A[q] = nxt[q];
        if (b + 1 < gr.b1) {
#pragma unroll
            for (int q = 0; q < 25; q++) nxt[q] = mats[(size_t)(b + 1) * 25 + q];
        }
        bool any_alive = false;
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) { svec[(size_t)b * HMM_NS + j] = s[j]; if (s[j] > -DBL_MAX) any_alive = true; }
        if (!any_alive) degenerate = true;  // every state at Double.MinValue: only the sequential order is defined
        hmm_apply(s, A);
    }
    if (degenerate) ctl->degenerate = 1;
}

// one step of BestPathViterbi's induction (HMM.cs:85-103): returns the packed back pointers (3 bits per state)
__device__ inline unsigned hmm_step(const double* e, double ls, double lo, double* s) {
    double ns[HMM_NS];
    unsigned packed = 0;
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) {
        const double cs = __dadd_rn(e[j], ls), co = __dadd_rn(e[j], lo);  // vitLogL = log E + log T[i][j]
        int state = 0;
        double mx = -DBL_MAX;
#pragma unroll
        for (int i = 0; i < HMM_NS; i++) {
            const double v = __dadd_rn(s[i], i == j ? cs : co);
            if (v > mx) { state = i; mx = v; }
        }
        ns[j] = mx;
        packed |= (unsigned)state << (3 * j);
    }
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) s[j] = ns[j];
    return packed;
}

// ---------------------------------------------------------------------------------------------------------------
// step C: exact replay inside every block
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) hmm_replay_kernel(const double* __restrict__ le, const HmmBlk* __restrict__ blk, int n_blk,
                                                         const double* __restrict__ svec, double ls, double lo,
                                                         unsigned short* __restrict__ back, unsigned short* __restrict__ map,
                                                         int* __restrict__ end_state) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const HmmBlk bk = blk[b];
    double s[HMM_NS];
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) s[j] = svec[(size_t)b * HMM_NS + j];
    unsigned origin = 0;  // origin[j] = state at the bin before the block on the best path into state j
    double en[HMM_NS];  // emissions of the next bin, loaded one step ahead
#pragma unroll
    for (int j = 0; j < HMM_NS; j++) en[j] = le[(size_t)bk.t0 * HMM_NS + j];
    for (long long t = bk.t0; t < bk.t1; t++) {
        double e[HMM_NS];
#pragma unroll
        for (int j = 0; j < HMM_NS; j++) e[j] = en[j];
        if (t + 1 < bk.t1) {
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) en[j] = le[(size_t)(t + 1) * HMM_NS + j];
        }
        const unsigned p = hmm_step(e, ls, lo, s);
        back[t] = (unsigned short)p;
        if (t == bk.t0) origin = p;
        else {
            unsigned o2 = 0;
#pragma unroll
            for (int j = 0; j < HMM_NS; j++) o2 |= ((origin >> (3 * ((p >> (3 * j)) & 7u))) & 7u) << (3 * j);
            origin = o2;
        }
    }
    map[b] = (unsigned short)origin;
    if (bk.last) end_state[bk.chrom] = hmm_best_final(s);
}

// back tracking over blocks, in two levels: D1 composes the maps of a group's blocks, D2 walks the groups of a
// chromosome from its end state (state at the last bin of every group, and at the chromosome's first bin), D3 walks
// the blocks of every group
__global__ void __launch_bounds__(64) hmm_group_map_kernel(const HmmGrp* __restrict__ grp, int n_grp, const unsigned short* __restrict__ map,
                                                           unsigned short* __restrict__ gmap) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grp) return;
    const HmmGrp gr = grp[g];
    unsigned cur = 0;
#pragma unroll
    for (int e = 0; e < HMM_NS; e++) cur |= (unsigned)e << (3 * e);
    for (int b = gr.b1 - 1; b >= gr.b0; b--) {
        const unsigned m = map[b];
        unsigned nx = 0;
#pragma unroll
        for (int e = 0; e < HMM_NS; e++) nx |= ((m >> (3 * ((cur >> (3 * e)) & 7u))) & 7u) << (3 * e);
        cur = nx;
    }
    gmap[g] = (unsigned short)cur;
}

__global__ void __launch_bounds__(128) hmm_backtrack_groups_kernel(const HmmChromInfo* __restrict__ ci, const unsigned short* __restrict__ gmap,
                                                                   const int* __restrict__ end_state, uint8_t* __restrict__ grp_end,
                                                                   uint8_t* __restrict__ states) {
    extern __shared__ unsigned short s_map[];
    const HmmChromInfo ch = ci[blockIdx.x];
    if (!ch.active) return;
    for (int g = threadIdx.x; g < ch.n_grp; g += blockDim.x) s_map[g] = gmap[ch.first_grp + g];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned e = (unsigned)end_state[blockIdx.x];
        for (int g = ch.n_grp - 1; g >= 0; g--) {
            grp_end[ch.first_grp + g] = (uint8_t)e;
            e = (s_map[g] >> (3 * e)) & 7u;
        }
        states[ch.a] = (uint8_t)e;
    }
}

__global__ void __launch_bounds__(64) hmm_backtrack_blocks_kernel(const HmmGrp* __restrict__ grp, int n_grp, const unsigned short* __restrict__ map,
                                                                  const uint8_t* __restrict__ grp_end, uint8_t* __restrict__ blk_end) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grp) return;
    const HmmGrp gr = grp[g];
    unsigned e = grp_end[g];
    for (int b = gr.b1 - 1; b >= gr.b0; b--) {
        blk_end[b] = (uint8_t)e;
        e = (map[b] >> (3 * e)) & 7u;
    }
}

// states of every bin of a block from its end state; number of state changes inside it
__global__ void __launch_bounds__(128) hmm_states_kernel(const HmmBlk* __restrict__ blk, int n_blk, const unsigned short* __restrict__ back,
                                                         const uint8_t* __restrict__ blk_end, uint8_t* __restrict__ states,
                                                         int* __restrict__ blk_cnt) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const HmmBlk bk = blk[b];
    unsigned e = blk_end[b];
    int cnt = 0;
    for (long long t = bk.t1 - 1; t >= bk.t0; t--) {
        states[t] = (uint8_t)e;
        const unsigned p = (back[t] >> (3 * e)) & 7u;
        cnt += p != e;
        e = p;
    }
    blk_cnt[b] = cnt;
}

// per chromosome: exclusive scan of the block counts (+1: breakpoint 0), total = n_bp
__global__ void __launch_bounds__(256) hmm_bp_scan_kernel(const HmmChromInfo* __restrict__ ci, const int* __restrict__ blk_cnt,
                                                          int* __restrict__ blk_off, int* __restrict__ n_bp, int* __restrict__ bp) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    const HmmChromInfo ch = ci[blockIdx.x];
    if (!ch.active) { if (threadIdx.x == 0) n_bp[blockIdx.x] = 0; return; }
    if (threadIdx.x == 0) s_carry = 1;
    __syncthreads();
    for (int base = 0; base < ch.n_blk; base += blockDim.x) {
        const int b = base + threadIdx.x;
        const int v = b < ch.n_blk ? blk_cnt[ch.first_blk + b] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += s_warp[w];
        const int carry = s_carry;
        if (b < ch.n_blk) blk_off[ch.first_blk + b] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { n_bp[blockIdx.x] = s_carry; bp[ch.a] = 0; }
}

__global__ void __launch_bounds__(128) hmm_bp_write_kernel(const HmmBlk* __restrict__ blk, int n_blk, const HmmChromInfo* __restrict__ ci,
                                                           const uint8_t* __restrict__ states, const int* __restrict__ blk_off,
                                                           int* __restrict__ bp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const HmmBlk bk = blk[b];
    const long long a = ci[bk.chrom].a;
    int k = blk_off[b];
    unsigned prev = states[bk.t0 - 1];
    for (long long t = bk.t0; t < bk.t1; t++) {
        const unsigned e = states[t];
        if (e != prev) bp[a + k++] = (int)(t - a);
        prev = e;
    }
}

// The reference loop as it is: one thread per chromosome (cross-check and degenerate inputs)
__global__ void hmm_sequential_kernel(const double* __restrict__ le, const HmmChromInfo* __restrict__ ci, int C, double log_start, double ls,
                                      double lo, unsigned short* __restrict__ back, uint8_t* __restrict__ states, int* __restrict__ n_bp,
                                      int* __restrict__ bp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const HmmChromInfo ch = ci[c];
    if (!ch.active) { n_bp[c] = 0; return; }
    double s[HMM_NS];
    hmm_init_scores(le + (size_t)ch.a * HMM_NS, log_start, ls, lo, s);
    for (long long t = ch.a + 1; t < ch.b; t++) back[t] = (unsigned short)hmm_step(le + (size_t)t * HMM_NS, ls, lo, s);
    unsigned e = (unsigned)hmm_best_final(s);
    for (long long t = ch.b - 1; t > ch.a; t--) {
        states[t] = (uint8_t)e;
        e = (back[t] >> (3 * e)) & 7u;
    }
    states[ch.a] = (uint8_t)e;
    int k = 0;
    bp[ch.a + k++] = 0;
    for (long long t = ch.a + 1; t < ch.b; t++)
        if (states[t] != states[t - 1]) bp[ch.a + k++] = (int)(t - ch.a);
    n_bp[c] = k;
}

// ---------------------------------------------------------------------------------------------------------------
// host side of the emission model
// ---------------------------------------------------------------------------------------------------------------
// MathNet.Numerics SpecialFunctions.GammaLn (Lanczos, g = 10.900511, 11 terms) and FactorialLn [EXT, restated]
double gamma_ln(double z) {
    static const double dk[11] = {2.48574089138753565546e-5, 1.05142378581721974210,   -3.45687097222016235469,
                                  4.51227709466894823700,    -2.98285225323576655721,  1.05639711577126713077,
                                  -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
                                  4.63399473359905636708e-6, -2.71994908488607703910e-9};
    const double r = 10.900511, ln_pi = 1.1447298858494001741434273513530587116472948129153,
                 log_2_sqrt_e_over_pi = 0.6207822376352452223455184457816472122518527279025978;
    if (z < 0.5) {
        double s = dk[0];
        for (int i = 1; i <= 10; i++) s += dk[i] / ((double)i - z);
        return ln_pi - std::log(std::sin(M_PI * z)) - std::log(s) - log_2_sqrt_e_over_pi - ((0.5 - z) * std::log((0.5 - z + r) / M_E));
    }
    double s = dk[0];
    for (int i = 1; i <= 10; i++) s += dk[i] / (z + (double)i - 1.0);
    return std::log(s) + log_2_sqrt_e_over_pi + ((z - 0.5) * std::log((z - 0.5 + r) / M_E));
}

double factorial_ln(int x) {
    if (x <= 1) return 0.0;
    if (x < 171) {
        double f = 1.0;
        for (int i = 2; i <= x; i++) f *= (double)i;
        return std::log(f);
    }
    return gamma_ln((double)x + 1.0);
}

// DistributionUtilities.NegativeBinomialWrapper (CanvasCommon/DistributionUtilities.cs:51-69, called by
// MultivariateNegativeBinomial, Distributions.cs:33): the clumping parameter has a floor of 2 on this path
void negative_binomial(double mean, double variance, int len, double* density) {
    const double r = std::max(2.0, std::pow(std::max(mean, 0.1), 2) / (std::max(variance, mean * 1.2) - mean));
    for (int x = 0; x < len; x++) {
        const double t = std::exp(std::log(std::pow(1 + mean / r, -r)) + std::log(std::pow(mean / (mean + r), x)) + gamma_ln(r + x) -
                                  factorial_ln(x) - gamma_ln(r));
        density[x] = (std::isnan(t) || std::isinf(t)) ? 0.0 : t;
    }
}

float f32_from_key(uint32_t k) { return f32_unkey(k); }

}  // namespace

// float counts -> the doubles CanvasPartition would parse from the text file that carries them (text_mode 1: "{F2}" of the
// .cleaned file, 2: float.ToString() of the pedigree workflow's merged file, 0: plain widening)
__global__ void hmm_counts_to_coverage_kernel(const float* __restrict__ count, long long n, int text_mode, double* __restrict__ cov) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = count[i];
        cov[i] = text_mode == 1 ? dotnet_f2_roundtrip(v) : (text_mode == 2 ? dotnet_g7_roundtrip(v) : (double)v);
    }
}

// count32_on_device: the float counts already sit in this GPU's memory (the pedigree chain, pedigree.cu)
static int partition_hmm_impl(cg_ctx* ctx, const cg_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off,
                              const double* coverage, const float* count32, int text_mode, const uint8_t* chrom_selected, int32_t* n_bp,
                              int32_t* bp, uint8_t* states_out, bool count32_on_device = false) {
    if (!ctx) return CG_ERR_ARG;
    if (n_chrom > HMM_MAX_CHROM) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (see DESIGN.md, Limits)");
    if (!o || !chrom_off || !n_bp || n_chrom < 0 || n_samples < 1)
        return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: bad argument");
    if (n_samples > HMM_MAX_SAMPLES) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: at most 4 samples are segmented jointly");
    if (o->n_states != HMM_NS) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: the reference model has 5 hidden states");
    if (o->per_sample && n_samples != 1)
        return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: PerSampleHMM segments one sample per call (CanvasPartition.cs:165-170)");
    const int C = n_chrom, S = n_samples;
    for (int c = 0; c < C; c++)
        if (chrom_off[c + 1] < chrom_off[c]) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: chromosome offsets must not decrease");
    const long long N = C > 0 ? chrom_off[C] - chrom_off[0] : 0;
    if (C > 0 && chrom_off[0] != 0) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: chrom_off[0] must be 0");
    for (int c = 0; c < C; c++) n_bp[c] = 0;
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    if (N == 0) return CG_OK;
    if ((!coverage && !count32) || !bp) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: null array");
    if (N > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: too many bins");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool debug = getenv("CANVAS_DEBUG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto dbg = [&](const char* what) {
        if (!debug) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[hmm] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    dbg("validate");

    // ---- chromosomes and blocks
    // (the block and group tables themselves — 47 000 blocks for a 3 M-bin genome — are filled on the device from these
    // per-chromosome ranges: hmm_tables_kernel)
    std::vector<HmmChromInfo> ci((size_t)C);
    long long n_blk_ll = 0;
    int n_grp = 0, max_chrom_grp = 1;
    for (int c = 0; c < C; c++) {
        HmmChromInfo& h = ci[(size_t)c];
        h.a = chrom_off[c]; h.b = chrom_off[c + 1];
        const long long n = h.b - h.a;
        h.active = (n > o->min_size && n >= 1 && (!chrom_selected || chrom_selected[c])) ? 1 : 0;
        h.first_blk = (int)n_blk_ll;
        h.tab = 0; h.max_thr = 0;
        h.n_blk = (h.active && n > 1) ? (int)((n - 1 + HMM_BLOCK - 1) / HMM_BLOCK) : 0;  // blocks of bins a+1 .. b-1
        n_blk_ll += h.n_blk;
        h.first_grp = n_grp;
        h.n_grp = (h.n_blk + HMM_GROUP - 1) / HMM_GROUP;
        n_grp += h.n_grp;
        max_chrom_grp = std::max(max_chrom_grp, h.n_grp);
    }
    const int n_blk = (int)n_blk_ll;
    if ((size_t)max_chrom_grp * 200 > 200 * 1024) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: chromosome too long");

    // ---- workspace
    const int nseg_q = S, nseg_m = S * C;
    size_t need = arena_need((size_t)S * N, 8) + arena_need((size_t)N * HMM_NS, 8) + arena_need(N, 1) * 2 + arena_need(N, 2) + arena_need(N, 4) +
                  arena_need(C + 1, sizeof(HmmChromInfo)) + arena_need(n_blk + 1, sizeof(HmmBlk)) + arena_need((size_t)(n_blk + 1) * 25, 8) +
                  arena_need((size_t)(n_blk + 1) * HMM_NS, 8) + arena_need(n_blk + 1, 2) + arena_need(n_blk + 1, 1) + arena_need(n_blk + 1, 4) * 2 +
                  arena_need(n_grp + 1, sizeof(HmmGrp)) + arena_need((size_t)(n_grp + 1) * 25, 8) + arena_need((size_t)(n_grp + 1) * HMM_NS, 8) +
                  arena_need(n_grp + 1, 2) + arena_need(n_grp + 1, 1) + arena_need(C + 1, 4) * 2 + arena_need(1, sizeof(HmmCtl)) + sel_state_bytes<uint32_t>(nseg_q) + sel_state_bytes<uint64_t>(nseg_m) +
                  arena_need((size_t)nseg_m + 1, 8) + (8u << 20);
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    double* d_cov = arena_take<double>(ctx, (size_t)S * N);
    double* d_le = arena_take<double>(ctx, (size_t)N * HMM_NS);
    uint8_t* d_chrom_id = arena_take<uint8_t>(ctx, N);
    uint8_t* d_states = arena_take<uint8_t>(ctx, N);
    unsigned short* d_back = arena_take<unsigned short>(ctx, N);
    int* d_bp = arena_take<int>(ctx, N);
    HmmChromInfo* d_ci = arena_take<HmmChromInfo>(ctx, C + 1);
    HmmBlk* d_blk = arena_take<HmmBlk>(ctx, n_blk + 1);
    double* d_mats = arena_take<double>(ctx, (size_t)(n_blk + 1) * 25);
    double* d_svec = arena_take<double>(ctx, (size_t)(n_blk + 1) * HMM_NS);
    unsigned short* d_map = arena_take<unsigned short>(ctx, n_blk + 1);
    uint8_t* d_blk_end = arena_take<uint8_t>(ctx, n_blk + 1);
    int* d_blk_cnt = arena_take<int>(ctx, n_blk + 1);
    int* d_blk_off = arena_take<int>(ctx, n_blk + 1);
    HmmGrp* d_grp = arena_take<HmmGrp>(ctx, n_grp + 1);
    double* d_gmats = arena_take<double>(ctx, (size_t)(n_grp + 1) * 25);
    double* d_gvec = arena_take<double>(ctx, (size_t)(n_grp + 1) * HMM_NS);
    unsigned short* d_gmap = arena_take<unsigned short>(ctx, n_grp + 1);
    uint8_t* d_grp_end = arena_take<uint8_t>(ctx, n_grp + 1);
    int* d_end_state = arena_take<int>(ctx, C + 1);
    int* d_nbp = arena_take<int>(ctx, C + 1);
    HmmCtl* d_ctl = arena_take<HmmCtl>(ctx, 1);
    double* d_var = arena_take<double>(ctx, (size_t)nseg_m + 1);
    SelState<uint32_t> sel_q;
    SelState<uint64_t> sel_m;
    bool ok = sel_state_alloc<uint32_t>(ctx, nseg_q, sel_q) && sel_state_alloc<uint64_t>(ctx, nseg_m, sel_m);
    if (!ok || !d_cov || !d_le || !d_chrom_id || !d_states || !d_back || !d_bp || !d_ci || !d_blk || !d_mats || !d_svec || !d_map ||
        !d_grp || !d_gmats || !d_gvec || !d_gmap || !d_grp_end || !d_blk_end || !d_blk_cnt || !d_blk_off || !d_end_state || !d_nbp || !d_ctl || !d_var)
        return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_hmm: device arena exhausted");
    cudaStream_t s = ctx->stream;
    if (coverage) {
        CG_CUDA(ctx, cudaMemcpyAsync(d_cov, coverage, (size_t)S * N * 8, cudaMemcpyHostToDevice, s));
    } else {
        // float counts: half the bytes over PCIe, the text round trip happens here (staged in the emission table's space)
        float* d_cnt = reinterpret_cast<float*>(d_le);
        static_assert(HMM_NS * 8 >= HMM_MAX_SAMPLES * 4, "the float staging area fits the emission table");
        if (!count32_on_device) CG_CUDA(ctx, cudaMemcpyAsync(d_cnt, count32, (size_t)S * N * 4, cudaMemcpyHostToDevice, s));
        hmm_counts_to_coverage_kernel<<<std::max(1, std::min(div_up((long long)S * N, 256), ctx->num_sms * 8)), 256, 0, s>>>(count32_on_device ? count32 : d_cnt, (long long)S * N, text_mode, d_cov);
        ctx->launches++;
    }
    CG_CUDA(ctx, cudaMemcpyAsync(d_ci, ci.data(), (size_t)C * sizeof(HmmChromInfo), cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_ctl, 0, sizeof(HmmCtl), s));
    CG_CUDA(ctx, cudaMemsetAsync(d_states, 0, N, s));
    CG_CUDA(ctx, cudaMemsetAsync(d_nbp, 0, (size_t)(C + 1) * 4, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    cudaEventRecord(ctx->stage_ev[0], s);
    ctx->stage_used[0] = true;
    const int grid_stream = std::max(1, std::min(div_up(N, 256), ctx->num_sms * 8));
    CG_LAUNCH(ctx, hmm_chrom_id_kernel, grid_stream, 256, 0, d_ci, C, N, d_chrom_id);

    dbg("upload");
    // ---- statistics of the emission model
    std::vector<double> haploid((size_t)S * C, 0.0), variance((size_t)S * C, 0.0);
    if (o->per_sample) {
        CG_CUDA(ctx, cudaMemsetAsync(sel_q.hist, 0, (size_t)nseg_q * SEL_G * SEL_BINS * sizeof(unsigned), s));
        CG_LAUNCH(ctx, hmm_quartile_request_kernel, div_up(nseg_q, 32), 32, 0, sel_q, N);
        QuartView qv{d_cov, N, S};
        sel_run_scatter<uint32_t, QuartView>(ctx, qv, sel_q, (long long)S * N);
        static_assert(HMM_MAX_SAMPLES * SEL_G * 4 <= 1024, "quartile keys fit their pinned slot");
        const uint32_t* keys = reinterpret_cast<const uint32_t*>(ctx->pinned + CG_PINNED_SMALL_AT);
        CG_CUDA(ctx, cg_readback_small(s, ctx->pinned + CG_PINNED_SMALL_AT, sel_q.req_key, (size_t)nseg_q * SEL_G * 4));
        ctx->launches++;
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        for (int sm = 0; sm < S; sm++) {
            float v[6] = {0, 0, 0, 0, 0, 0};
            if (N >= 2)
                for (int q = 0; q < 6; q++) v[q] = f32_from_key(keys[(size_t)sm * SEL_G + q]);
            float q1, q2, q3;
            // Utilities.Quartiles in single precision (Utilities.cs:361-419)
            if (N % 2 == 0) {
                q2 = (v[0] + v[1]) / 2;
                if ((N / 2) % 2 == 0) { q1 = (v[2] + v[3]) / 2; q3 = (v[4] + v[5]) / 2; }
                else { q1 = v[2]; q3 = v[4]; }
            } else {
                q2 = v[0];
                if ((N - 1) % 4 == 0) { q1 = (v[2] * 0.25f) + (v[3] * 0.75f); q3 = (v[4] * 0.75f) + (v[5] * 0.25f); }
                else { q1 = (v[2] * 0.75f) + (v[3] * 0.25f); q3 = (v[4] * 0.25f) + (v[5] * 0.75f); }
            }
            if (N < 2) {
                double only = 0;  // a one-bin genome: its own value is the median (read back: the coverage may have been built on the device)
                if (N == 1) { CG_CUDA(ctx, cudaMemcpy(&only, d_cov + (size_t)sm * N, 8, cudaMemcpyDeviceToHost)); }
                q1 = q3 = 0; q2 = (float)only;
            }
            const float iqr = q3 - q1;
            for (int c = 0; c < C; c++) {
                haploid[(size_t)sm * C + c] = (double)q2 / 2.0;
                variance[(size_t)sm * C + c] = (double)(iqr * iqr);
            }
        }
    } else {
        CG_CUDA(ctx, cudaMemsetAsync(sel_m.hist, 0, (size_t)nseg_m * SEL_G * SEL_BINS * sizeof(unsigned), s));
        CG_LAUNCH(ctx, hmm_median_request_kernel, div_up(nseg_m, 64), 64, 0, sel_m, d_ci, C);
        ChromMedView mv{d_cov, d_chrom_id, N, S, C};
        sel_run_scatter<uint64_t, ChromMedView>(ctx, mv, sel_m, (long long)S * N);
        CG_LAUNCH(ctx, hmm_variance_kernel, nseg_m, 1024, 0, d_cov, N, d_ci, C, d_var);
        std::vector<uint64_t> keys((size_t)nseg_m * SEL_G);
        std::vector<double> var((size_t)nseg_m);
        CG_CUDA(ctx, cudaMemcpyAsync(keys.data(), sel_m.req_key, keys.size() * 8, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaMemcpyAsync(var.data(), d_var, var.size() * 8, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        for (int sm = 0; sm < S; sm++)
            for (int c = 0; c < C; c++) {
                const size_t seg = (size_t)sm * C + c;
                if (ci[(size_t)c].b - ci[(size_t)c].a <= 0) continue;
                const double lo_v = f64_unkey(keys[seg * SEL_G + 0]), hi_v = f64_unkey(keys[seg * SEL_G + 1]);
                const double med = ((ci[(size_t)c].b - ci[(size_t)c].a) & 1) ? lo_v : (lo_v + hi_v) / 2.0;  // SortedList<double>.Median
                haploid[seg] = std::max(1.0, med) / 2.0;
                variance[seg] = var[seg];
            }
    }
    CG_CHECK_LAUNCHES(ctx);
    dbg("statistics");

    // ---- emission tables (host: the same libm calls as the reference's Math.*; one table per distinct model)
    int L = 1;
    std::vector<int> tab_of((size_t)C, -1);
    int n_tab = 0;
    for (int c = 0; c < C; c++) {
        if (!ci[(size_t)c].active) continue;
        double hmax = haploid[c];
        for (int sm = 1; sm < S; sm++) hmax = std::max(hmax, haploid[(size_t)sm * C + c]);
        ci[(size_t)c].max_thr = hmax * HMM_NS;
        if (!(ci[(size_t)c].max_thr < 2e5)) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_hmm: coverage scale too large for the emission table");
        L = std::max(L, (int)std::nearbyint(ci[(size_t)c].max_thr) + 2);
        if (o->per_sample && n_tab > 0) tab_of[(size_t)c] = 0;  // whole-genome statistics: one model for all chromosomes
        else tab_of[(size_t)c] = n_tab++;
        ci[(size_t)c].tab = tab_of[(size_t)c];
    }
    const size_t per_tab = (size_t)HMM_NS * S * L;
    std::vector<double> tab(std::max<size_t>(1, per_tab * (size_t)std::max(n_tab, 1)), 0.0);
    {
        std::vector<char> built((size_t)std::max(n_tab, 1), 0);
        std::vector<double> dens((size_t)L);
        for (int c = 0; c < C; c++) {
            const int tb = tab_of[(size_t)c];
            if (tb < 0 || built[(size_t)tb]) continue;
            built[(size_t)tb] = 1;
            double* base = tab.data() + (size_t)tb * per_tab;
            for (int cn = 0; cn < HMM_NS; cn++)
                for (int sm = 0; sm < S; sm++) {
                    negative_binomial(std::max((double)cn, 0.1) * haploid[(size_t)sm * C + c], variance[(size_t)sm * C + c], L, dens.data());
                    std::copy(dens.begin(), dens.end(), base + ((size_t)cn * S + sm) * L);
                }
            if (S == 1) {
                // one sample: the arrangement list of state j is {[j]}, so E_j = 1.0 * P_j (or the merged pairs when
                // the states are not all distinct, Distributions.cs:276-284); store log E
                for (int x = 0; x < L; x++) {
                    double p[HMM_NS];
                    for (int g = 0; g < HMM_NS; g++) p[g] = base[(size_t)g * L + x];
                    if (!o->per_sample) {
                        const double lo_p = std::max(p[0], p[1]), hi_p = std::max(p[3], p[4]);
                        p[0] = p[1] = lo_p; p[3] = p[4] = hi_p;
                    }
                    for (int g = 0; g < HMM_NS; g++) {
                        double l = 1.0 * p[g];
                        if (std::isnan(l) || std::isinf(l)) l = 0;
                        base[(size_t)g * L + x] = std::log(l);
                    }
                }
            }
        }
    }
    rc = aux_reserve(ctx, tab.size() * 8);
    if (rc) return rc;
    double* d_tab = (double*)ctx->aux;
    dbg("tables");
    CG_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemcpyAsync(d_ci, ci.data(), (size_t)C * sizeof(HmmChromInfo), cudaMemcpyHostToDevice, s));
    if (n_blk > 0) CG_LAUNCH(ctx, hmm_tables_kernel, C, 256, 0, d_ci, d_blk, d_grp);
    const double self_t = 0.99;
    const double ls = std::log(self_t), lo = std::log((1.0 - self_t) / (HMM_NS - 1));
    const double log_start = std::log((double)(1.0f / HMM_NS));

    cudaEventRecord(ctx->stage_ev[1], s);
    cudaEventRecord(ctx->stage_ev[2], s);
    ctx->stage_used[1] = true;
    CG_LAUNCH(ctx, hmm_emission_kernel, grid_stream, 256, 0, d_cov, N, S, d_chrom_id, d_ci, d_tab, L, o->per_sample ? 1 : 0, d_le, d_ctl);
    cudaEventRecord(ctx->stage_ev[3], s);
    cudaEventRecord(ctx->stage_ev[4], s);
    ctx->stage_used[2] = true;
    dbg("emission");
    bool sequential = o->exact_sequential != 0;
    if ((size_t)max_chrom_grp * 200 > 48 * 1024)
        CG_CUDA(ctx, cudaFuncSetAttribute(hmm_scan_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_chrom_grp * 200));
    HmmCtl h_ctl = {0, 0};
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!sequential) {
            if (n_blk > 0) {
                CG_LAUNCH(ctx, hmm_block_kernel, div_up(n_blk, 128), 128, 0, d_le, d_blk, n_blk, ls, lo, d_mats);
                CG_LAUNCH(ctx, hmm_group_kernel, div_up(n_grp * 32, 128), 128, 0, d_grp, n_grp, d_mats, d_gmats);
            }
            CG_LAUNCH(ctx, hmm_scan_groups_kernel, C, 128, (size_t)max_chrom_grp * 200, d_le, d_ci, d_gmats, log_start, ls, lo, d_gvec, d_end_state);
            if (n_blk > 0) {
                CG_LAUNCH(ctx, hmm_scan_blocks_kernel, div_up(n_grp, 64), 64, 0, d_grp, n_grp, d_mats, d_gvec, d_svec, d_ctl);
                CG_LAUNCH(ctx, hmm_replay_kernel, div_up(n_blk, 128), 128, 0, d_le, d_blk, n_blk, d_svec, ls, lo, d_back, d_map, d_end_state);
                CG_LAUNCH(ctx, hmm_group_map_kernel, div_up(n_grp, 64), 64, 0, d_grp, n_grp, d_map, d_gmap);
            }
            CG_LAUNCH(ctx, hmm_backtrack_groups_kernel, C, 128, (size_t)max_chrom_grp * 2, d_ci, d_gmap, d_end_state, d_grp_end, d_states);
            if (n_blk > 0) CG_LAUNCH(ctx, hmm_backtrack_blocks_kernel, div_up(n_grp, 64), 64, 0, d_grp, n_grp, d_map, d_grp_end, d_blk_end);
            if (n_blk > 0) CG_LAUNCH(ctx, hmm_states_kernel, div_up(n_blk, 128), 128, 0, d_blk, n_blk, d_back, d_blk_end, d_states, d_blk_cnt);
            CG_LAUNCH(ctx, hmm_bp_scan_kernel, C, 256, 0, d_ci, d_blk_cnt, d_blk_off, d_nbp, d_bp);
            if (n_blk > 0) CG_LAUNCH(ctx, hmm_bp_write_kernel, div_up(n_blk, 128), 128, 0, d_blk, n_blk, d_ci, d_states, d_blk_off, d_bp);
        } else {
            CG_LAUNCH(ctx, hmm_sequential_kernel, C, 1, 0, d_le, d_ci, C, log_start, ls, lo, d_back, d_states, d_nbp, d_bp);
        }
        static_assert(sizeof(HmmCtl) % 4 == 0 && sizeof(HmmCtl) <= 256, "control block fits its pinned slot");
        CG_CUDA(ctx, cg_readback_small(s, ctx->pinned + CG_PINNED_SMALL_AT + 1024, d_ctl, sizeof(HmmCtl)));
        ctx->launches++;
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(&h_ctl, ctx->pinned + CG_PINNED_SMALL_AT + 1024, sizeof(HmmCtl));
        if (sequential || !h_ctl.degenerate) break;
        sequential = true;  // a score reached Double.MinValue: only the strictly sequential order reproduces the reference
    }
    cudaEventRecord(ctx->stage_ev[5], s);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    dbg("viterbi");
    // ---- results
    static_assert(HMM_MAX_CHROM * 4 <= 2048, "breakpoint counts fit their pinned slot");
    const int* h_nbp = reinterpret_cast<const int*>(ctx->pinned + CG_PINNED_SMALL_AT + 2048);
    CG_CUDA(ctx, cg_readback_small(s, ctx->pinned + CG_PINNED_SMALL_AT + 2048, d_nbp, (size_t)C * 4));
    ctx->launches++;
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    long long total_bp = 0;
    for (int c = 0; c < C; c++) { n_bp[c] = h_nbp[c]; total_bp += n_bp[c] > 0 ? n_bp[c] : 0; }
    if (total_bp * 4 <= (long long)CG_PINNED_LIST_BYTES) {
        // short lists (the usual case): through the page-locked block, written by kernels (see cg_readback_small)
        int* stage = reinterpret_cast<int*>(ctx->pinned + CG_PINNED_LIST_AT);
        long long at = 0;
        for (int c = 0; c < C; c++)
            if (n_bp[c] > 0) {
                CG_CUDA(ctx, cg_readback_small(s, stage + at, d_bp + chrom_off[c], (size_t)n_bp[c] * 4));
                ctx->launches++;
                at += n_bp[c];
            }
        if (states_out) CG_CUDA(ctx, cudaMemcpyAsync(states_out, d_states, (size_t)N, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        at = 0;
        for (int c = 0; c < C; c++)
            if (n_bp[c] > 0) { memcpy(bp + chrom_off[c], stage + at, (size_t)n_bp[c] * 4); at += n_bp[c]; }
    } else {
        for (int c = 0; c < C; c++)
            if (n_bp[c] > 0)
                CG_CUDA(ctx, cudaMemcpyAsync(bp + chrom_off[c], d_bp + chrom_off[c], (size_t)n_bp[c] * 4, cudaMemcpyDeviceToHost, s));
        if (states_out) CG_CUDA(ctx, cudaMemcpyAsync(states_out, d_states, (size_t)N, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    if (h_ctl.bad_value) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: coverage must be finite and non-negative (Convert.ToInt32 / the emission table index throw in the reference)");
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    ctx->stats[0] = sequential ? 1.0 : 0.0;
    ctx->stats[1] = (double)n_blk;
    ctx->stats[2] = (double)L;
    ctx->stats[3] = (double)N;
    return CG_OK;
}

extern "C" int cg_partition_hmm_shard(cg_ctx* ctx, const cg_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off,
                                      const double* coverage, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                                      uint8_t* states_out) {
    if (ctx && !coverage && n_chrom > 0 && chrom_off && chrom_off[n_chrom] > 0) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: null array");
    return partition_hmm_impl(ctx, o, n_samples, n_chrom, chrom_off, coverage, nullptr, 0, chrom_selected, n_bp, bp, states_out);
}

extern "C" int cg_partition_hmm_counts(cg_ctx* ctx, const cg_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off,
                                       const float* count, int text_mode, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                                       uint8_t* states_out) {
    if (ctx && (text_mode < 0 || text_mode > 2)) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm_counts: text_mode must be 0, 1 or 2");
    if (ctx && !count && n_chrom > 0 && chrom_off && chrom_off[n_chrom] > 0) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_hmm: null array");
    return partition_hmm_impl(ctx, o, n_samples, n_chrom, chrom_off, nullptr, count, text_mode, chrom_selected, n_bp, bp, states_out);
}

// One sample of the pedigree chain: float counts in device memory, results to the host (pedigree.cu).
int hmm_partition_device_counts(cg_ctx* ctx, const cg_hmm_opts* o, int n_chrom, const int64_t* chrom_off, const float* d_count,
                                int text_mode, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp) {
    return partition_hmm_impl(ctx, o, 1, n_chrom, chrom_off, nullptr, d_count, text_mode, chrom_selected, n_bp, bp, nullptr, true);
}

extern "C" int cg_partition_hmm(cg_ctx* ctx, const cg_hmm_opts* o, int n_samples, int n_chrom, const int64_t* chrom_off,
                                const double* coverage, int32_t* n_bp, int32_t* bp, uint8_t* states_out) {
    return cg_partition_hmm_shard(ctx, o, n_samples, n_chrom, chrom_off, coverage, nullptr, n_bp, bp, states_out);
}
