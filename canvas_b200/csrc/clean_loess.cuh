// LOESS GC normalisation (CanvasClean -m LOESS).
//
// Reference: LoessGCNormalizer.cs:35-131 (log transform, bandwidth by golden-section search on the bins
// outside chrY, final fit on all bins), LoessInterpolator.cs:61-315 (degree-1 local regression with
// tricube weights over the ceil(bandwidth * n) nearest points; window edges follow the STABLE sort
// of x, i.e. they cut through a GC bucket in genomic order), Utilities.cs:1014-1044 (golden section).
//
// x only takes the values 0..100, so a window is a handful of whole GC buckets plus two partial
// ones, and every weighted sum of the fit is  sum_g w_g * (sum of y over the part of bucket g inside
// the window).  One stable partition of log(count) by GC plus a prefix sum turns each fit into <= 101
// terms; the whole bandwidth search (2 fits x ~25 objective evaluations x 101 query points) then runs
// in one small kernel.  The sums are associated bucket by bucket instead of point by point, so fitted
// values agree with the reference to ~1e-13 relative (the contract for floating point is 1e-5).
#pragma once
#include "clean.cuh"
#include "select.cuh"

constexpr int LO_TILE = 1024;
constexpr int LO_KEYS = GC_BINS + 1;  // GC 0..100, LO_KEYS - 1 = not in the data set
constexpr int LO_SKIP = GC_BINS;

struct LoessCtl {
    unsigned cnt[2][LO_KEYS];
    unsigned P[2][LO_KEYS + 1];  // bucket starts in the GC-sorted order; P[.][GC_BINS] = points in the set
    double med[2];               // median of log(count): all points, points outside chrY
    double f[2 * GC_BINS + 2];   // final fitted curve at minGC, minGC + 1, ...
    int flen, min_gc, ok, evals;
    double best_bw;
};

struct LoessDev {
    double* y;            // [n] log(count)
    uint8_t* key[2];      // [n] GC bucket in set 0 (finite log) / set 1 (finite log, not chrY), LO_SKIP otherwise
    unsigned* thist[2];   // [ntiles][LO_KEYS] per-tile bucket counts, then offsets inside the bucket
    double* sorted[2];    // [n] y in stable GC order, then its running sum inside each scan tile
    double* tbase[2];     // [ntiles + 1] sums of the earlier scan tiles
    LoessCtl* lc;
    SelState<uint64_t> sel;
    const uint8_t* is_chry;
    int ntiles;
};

struct LoessYView {  // medians of log(count): segment 0 = all points, 1 = points outside chrY
    const double* y;
    const uint8_t* key0;
    const uint8_t* key1;
    const CleanCtl* ctl;
    const int* enabled;
    __device__ long long size() const { return *enabled ? ctl->n2 : 0; }
    __device__ bool get(long long i, uint64_t& key, int& a, int& b) const {
        if (key0[i] == LO_SKIP) return false;
        key = f64_key(y[i]);
        a = 0;
        b = key1[i] == LO_SKIP ? -1 : 1;
        return true;
    }
};

// y = log(count); +-Infinity is dropped (LoessGCNormalizer.cs:45-47), chrY is left out of the search set (:52-56)
__global__ void loess_prep_kernel(const float* __restrict__ count, const uint8_t* __restrict__ gc, const uint8_t* __restrict__ chrom,
                                  const uint8_t* __restrict__ alive, const uint8_t* __restrict__ is_chry, const CleanCtl* ctl,
                                  const int* enabled, LoessDev d) {
    if (!*enabled) return;
    const int n = ctl->n2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = log((double)count[i]);
        const bool in = alive[i] && !isinf(v);
        d.y[i] = v;
        d.key[0][i] = in ? gc[i] : LO_SKIP;
        d.key[1][i] = (in && !is_chry[chrom[i]]) ? gc[i] : LO_SKIP;
    }
}

__global__ void __launch_bounds__(LO_TILE) loess_tile_hist_kernel(const CleanCtl* ctl, const int* enabled, LoessDev d) {
    if (!*enabled) return;
    __shared__ unsigned h[2][LO_KEYS];
    const int n = ctl->n2;
    const long long i = (long long)blockIdx.x * LO_TILE + threadIdx.x;
    if ((long long)blockIdx.x * LO_TILE >= n) return;
    for (int k = threadIdx.x; k < 2 * LO_KEYS; k += LO_TILE) (&h[0][0])[k] = 0;
    __syncthreads();
    if (i < n) {
        atomicAdd(&h[0][d.key[0][i]], 1u);
        atomicAdd(&h[1][d.key[1][i]], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * LO_KEYS; k += LO_TILE) d.thist[k / LO_KEYS][(size_t)blockIdx.x * LO_KEYS + k % LO_KEYS] = h[k / LO_KEYS][k % LO_KEYS];
}

// per bucket: offsets of every tile inside the bucket; then bucket starts
__global__ void loess_offsets_kernel(const CleanCtl* ctl, const int* enabled, LoessDev d) {
    if (!*enabled) return;
    const int t = threadIdx.x;
    const int ntiles = (ctl->n2 + LO_TILE - 1) / LO_TILE;
    if (t < 2 * LO_KEYS) {
        const int s = t / LO_KEYS, g = t % LO_KEYS;
        unsigned run = 0;
        for (int tile = 0; tile < ntiles; tile++) {
            const unsigned v = d.thist[s][(size_t)tile * LO_KEYS + g];
            d.thist[s][(size_t)tile * LO_KEYS + g] = run;
            run += v;
        }
        d.lc->cnt[s][g] = run;
    }
    __syncthreads();
    if (t < 2) {
        unsigned run = 0;
        for (int g = 0; g < GC_BINS; g++) { d.lc->P[t][g] = run; run += d.lc->cnt[t][g]; }
        d.lc->P[t][GC_BINS] = run;
        d.lc->P[t][GC_BINS + 1] = run;
    }
}

// stable scatter into GC order: rank inside the tile = warp offset + rank among the lanes with the same key
__global__ void __launch_bounds__(LO_TILE) loess_scatter_kernel(const CleanCtl* ctl, const int* enabled, LoessDev d) {
    if (!*enabled) return;
    __shared__ unsigned short wcnt[LO_TILE / 32][LO_KEYS];
    const int n = ctl->n2;
    if ((long long)blockIdx.x * LO_TILE >= n) return;
    const long long i = (long long)blockIdx.x * LO_TILE + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double v = i < n ? d.y[i] : 0.0;
    for (int s = 0; s < 2; s++) {
        for (int k = threadIdx.x; k < (LO_TILE / 32) * LO_KEYS; k += LO_TILE) (&wcnt[0][0])[k] = 0;
        __syncthreads();
        const int key = i < n ? d.key[s][i] : LO_SKIP;
        const unsigned same = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (rank == 0) wcnt[w][key] = (unsigned short)__popc(same);
        __syncthreads();
        if (threadIdx.x < LO_KEYS) {
            unsigned run = 0;
            for (int q = 0; q < LO_TILE / 32; q++) { const unsigned c = wcnt[q][threadIdx.x]; wcnt[q][threadIdx.x] = (unsigned short)run; run += c; }
        }
        __syncthreads();
        if (key != LO_SKIP)
            d.sorted[s][d.lc->P[s][key] + d.thist[s][(size_t)blockIdx.x * LO_KEYS + key] + wcnt[w][key] + rank] = v;
        __syncthreads();
    }
}

// running sums inside tiles of LO_TILE sorted values (in place) + tile totals
__global__ void __launch_bounds__(LO_TILE) loess_scan_tiles_kernel(const int* enabled, LoessDev d) {
    if (!*enabled) return;
    __shared__ double wsum[LO_TILE / 32];
    const int s = blockIdx.y;
    const unsigned n = d.lc->P[s][GC_BINS];
    const long long i = (long long)blockIdx.x * LO_TILE + threadIdx.x;
    if ((long long)blockIdx.x * LO_TILE >= n) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double v = i < n ? d.sorted[s][i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
        wsum[lane] = x;
    }
    __syncthreads();
    if (w > 0) v += wsum[w - 1];
    if (i < n) d.sorted[s][i] = v;
    if (threadIdx.x == LO_TILE - 1) d.tbase[s][blockIdx.x + 1] = v;  // inclusive total of this tile (zero padded)
}

__global__ void loess_scan_totals_kernel(const int* enabled, LoessDev d) {
    if (!*enabled) return;
    const int s = threadIdx.x;
    if (s >= 2) return;
    const unsigned n = d.lc->P[s][GC_BINS];
    const int nt = (int)((n + LO_TILE - 1) / LO_TILE);
    double run = 0.0;
    d.tbase[s][0] = 0.0;
    for (int t = 1; t <= nt; t++) { run += d.tbase[s][t]; d.tbase[s][t] = run; }
}

__global__ void loess_median_request_kernel(SelState<uint64_t> st, const LoessCtl* lc, const int* enabled) {
    const int s = threadIdx.x;
    if (s >= 2) return;
    const unsigned n = *enabled ? lc->P[s][GC_BINS] : 0u;
    if (n == 0) { st.nreq[s] = 0; return; }
    st.nreq[s] = 2;
    st.req_k[s * SEL_G + 0] = (n & 1u) ? n / 2 : n / 2 - 1;
    st.req_k[s * SEL_G + 1] = n / 2;
}

__global__ void loess_median_finish_kernel(SelState<uint64_t> st, LoessCtl* lc, const int* enabled) {
    const int s = threadIdx.x;
    if (s >= 2 || !*enabled) return;
    if (st.nreq[s] == 0) { lc->med[s] = 0.0; return; }
    const double a = f64_unkey(st.req_key[s * SEL_G + 0]), b = f64_unkey(st.req_key[s * SEL_G + 1]);
    lc->med[s] = st.req_key[s * SEL_G + 0] == st.req_key[s * SEL_G + 1] ? a : (a + b) / 2.0;
}

// ---------------------------------------------------------------------------------------------
// The model on one data set
// ---------------------------------------------------------------------------------------------
struct LoSet {
    const unsigned* P;    // [GC_BINS + 2]
    const double* ps;     // tile-local running sums of the sorted y
    const double* tbase;  // sums of earlier tiles
    unsigned n;
    int min_gc, max_gc;
};

__device__ inline double lo_prefix(const LoSet& D, unsigned k) {  // sum of the first k sorted values
    if (k == 0) return 0.0;
    return D.tbase[(k - 1) / LO_TILE] + D.ps[k - 1];
}
__device__ inline int lo_gc_at(const LoSet& D, unsigned k) {  // x of the k-th sorted point
    int lo = 0, hi = GC_BINS - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (D.P[mid] <= k) lo = mid; else hi = mid - 1; }
    return lo;
}

// LoessInterpolator.updateBandwidthInterval (:271-301), moving whole runs of equal x at a time
__device__ bool lo_update_window(const LoSet& D, double x, long long& left, long long& right) {
    const long long n = D.n;
    bool moved = false;
    if (right < n - 1 && x > (double)lo_gc_at(D, (unsigned)right)) {
        // advance until xv[right] >= x
        int g = (int)ceil(x);
        if (g > GC_BINS) g = GC_BINS;
        long long target = g < 0 ? 0 : (long long)D.P[g];
        if (target > n - 1) target = n - 1;
        if (target > right) { left += target - right; right = target; moved = true; }
    }
    while (right < n - 1) {
        const int gr = lo_gc_at(D, (unsigned)(right + 1)), gl = lo_gc_at(D, (unsigned)left);
        if (!((double)gr - x < x - (double)gl)) break;
        long long d = min((long long)D.P[gr + 1] - (right + 1), (long long)D.P[gl + 1] - left);
        d = min(d, n - 1 - right);
        if (d <= 0) break;
        left += d;
        right += d;
        moved = true;
    }
    return moved;
}

__device__ inline double lo_tricube(double x) {
    const double t = 1 - x * x * x;
    return t * t * t;
}

// computeCoefficients + predict (:199-262) at x over the window [left, right]; y' = y - sub[g] + add
__device__ double lo_fit_at(const LoSet& D, double x, long long left, long long right, const double* sub, double add) {
    const int gl = lo_gc_at(D, (unsigned)left), gr = lo_gc_at(D, (unsigned)right);
    const double xe = (x - (double)gl > (double)gr - x) ? (double)gl : (double)gr;
    const double denom = fabs(1.0 / (xe - x));
    double sw = 0, sx = 0, sxx = 0, sy = 0, sxy = 0;
    for (int g = gl; g <= gr; g++) {
        const long long a = max((long long)D.P[g], left), b = min((long long)D.P[g + 1], right + 1);
        if (b <= a) continue;
        const double m = (double)(b - a);
        double ysum = lo_prefix(D, (unsigned)b) - lo_prefix(D, (unsigned)a);
        if (sub) ysum = ysum - m * sub[g] + m * add;
        const double xk = (double)g;
        const double w = lo_tricube(fabs(x - xk) * denom) * 1.0;
        const double xkw = xk * w;
        sw += m * w;
        sx += m * xkw;
        sxx += m * (xk * xkw);
        sy += ysum * w;
        sxy += ysum * xkw;
    }
    const double mx = sx / sw, my = sy / sw, mxy = sxy / sw, mxx = sxx / sw;
    const double beta = (mxx == mx * mx) ? 0 : (mxy - mx * my) / (mxx - mx * mx);
    const double alpha = my - beta * mx;
    double y = 0;
    y += 1.0 * alpha;
    y += x * beta;
    return y;
}

struct LoScratch {
    double ivmax[GC_BINS + 4];
    long long ivl[GC_BINS + 4], ivr[GC_BINS + 4];
    int niv;
    double f1[2 * GC_BINS + 2], f2[2 * GC_BINS + 2], sub[GC_BINS + 1];
    double fc, fd, a, b, c, d;
    int go;
};

// Train (intervals with xStep = 1, :177-197) + Predict at minGC .. minGC + maxGC - 1 (LoessGCNormalizer.cs:74,111:
// Enumerable.Range(minGC, maxGC) takes maxGC as a COUNT).  All threads of the block; result in out[0..max_gc).
__device__ void lo_model(const LoSet& D, double bandwidth, const double* sub, double add, LoScratch& sc, double* out) {
    const int t = threadIdx.x;
    __syncthreads();
    if (t == 0) {
        const long long bw = (long long)ceil(bandwidth * (double)D.n);
        long long left = 0, right = bw - 1;
        if (right > (long long)D.n - 1) right = (long long)D.n - 1;
        int niv = 0;
        for (double x = (double)D.min_gc; x <= (double)D.max_gc; x += 1.0) {
            long long nl = left, nr = right;
            if (lo_update_window(D, x, nl, nr)) {
                sc.ivmax[niv] = x; sc.ivl[niv] = left; sc.ivr[niv] = right; niv++;
                left = nl; right = nr;
            }
        }
        sc.ivmax[niv] = INFINITY; sc.ivl[niv] = left; sc.ivr[niv] = right; niv++;
        sc.niv = niv;
    }
    __syncthreads();
    for (int q = t; q < D.max_gc; q += blockDim.x) {
        const double x = (double)(D.min_gc + q);
        int ii = 0;
        while (ii + 1 < sc.niv && sc.ivmax[ii] <= x) ii++;
        out[q] = lo_fit_at(D, x, sc.ivl[ii], sc.ivr[ii], sub, add);
    }
    __syncthreads();
}

// LoessGCNormalizer objective (:97-131): SD of the second-pass fitted values
__device__ double lo_objective(const LoSet& D, double bandwidth, double median_y, LoScratch& sc) {
    const int t = threadIdx.x;
    lo_model(D, bandwidth, nullptr, 0.0, sc, sc.f1);
    for (int g = t; g <= GC_BINS; g += blockDim.x) {
        int idx = g - D.min_gc;
        if (idx > D.max_gc - 1) idx = D.max_gc - 1;
        if (idx < 0) idx = 0;
        sc.sub[g] = sc.f1[idx];
    }
    __syncthreads();
    lo_model(D, bandwidth, sc.sub, median_y, sc, sc.f2);
    // Utilities.StandardDeviation over the per-point fitted values, bucket by bucket
    double mean = 0, ss = 0;
    for (int g = D.min_gc; g <= D.max_gc; g++) {
        const double m = (double)(D.P[g + 1] - D.P[g]);
        int idx = g - D.min_gc;
        if (idx > D.max_gc - 1) idx = D.max_gc - 1;
        mean += m * sc.f2[idx];
    }
    mean /= (double)D.n;
    for (int g = D.min_gc; g <= D.max_gc; g++) {
        const double m = (double)(D.P[g + 1] - D.P[g]);
        int idx = g - D.min_gc;
        if (idx > D.max_gc - 1) idx = D.max_gc - 1;
        const double dv = sc.f2[idx] - mean;
        ss += m * (dv * dv);
    }
    return sqrt(ss / ((double)D.n - 1.0));
}

__device__ inline void lo_make_set(const LoessDev& d, int s, LoSet& D) {
    D.P = d.lc->P[s];
    D.ps = d.sorted[s];
    D.tbase = d.tbase[s];
    D.n = d.lc->P[s][GC_BINS];
    D.min_gc = 0;
    D.max_gc = 0;
    for (int g = 0; g < GC_BINS; g++)
        if (d.lc->cnt[s][g] > 0) { D.min_gc = g; break; }
    for (int g = GC_BINS - 1; g >= 0; g--)
        if (d.lc->cnt[s][g] > 0) { D.max_gc = g; break; }
}

// bandwidth search on the set without chrY, final curve on all points (LoessGCNormalizer.cs:61-95)
__global__ void __launch_bounds__(128) loess_search_kernel(const int* enabled, LoessDev d) {
    if (!*enabled) return;
    __shared__ LoScratch sc;
    LoSet A, Y;
    lo_make_set(d, 0, A);
    lo_make_set(d, 1, Y);
    LoessCtl* lc = d.lc;
    if (A.n == 0 || Y.n < 2 || A.max_gc < 1) {  // nothing to fit: counts stay as they are
        if (threadIdx.x == 0) { lc->ok = 0; lc->flen = 0; }
        return;
    }
    double lo = fmax(2.0 / (double)Y.n, 0.3), hi = 0.75;
    if (hi < lo) hi = lo;
    // Utilities.GoldenSectionSearch (Utilities.cs:1014-1044), tol 1e-5
    const double gr = 0.618034;
    double a = lo, b = hi;
    double c = b - gr * (b - a), dd = a + gr * (b - a);
    double fc = lo_objective(Y, c, lc->med[1], sc);
    double fd = lo_objective(Y, dd, lc->med[1], sc);
    int evals = 2;
    while (fabs(dd - c) > 1e-5) {
        if (fc < fd) {
            b = dd; dd = c; fd = fc;
            c = b - gr * (b - a);
            fc = lo_objective(Y, c, lc->med[1], sc);
        } else {
            a = c; c = dd; fc = fd;
            dd = a + gr * (b - a);
            fd = lo_objective(Y, dd, lc->med[1], sc);
        }
        evals++;
    }
    const double best = (b + a) / 2;
    lo_model(A, best, nullptr, 0.0, sc, sc.f1);
    for (int q = threadIdx.x; q < A.max_gc; q += blockDim.x) lc->f[q] = sc.f1[q];
    if (threadIdx.x == 0) { lc->flen = A.max_gc; lc->min_gc = A.min_gc; lc->best_bw = best; lc->ok = 1; lc->evals = evals; }
}

// count = (float) exp(log(count) - fitted[gc - minGC] + median) (LoessGCNormalizer.cs:76-81)
__global__ void loess_apply_kernel(float* __restrict__ count, const uint8_t* __restrict__ gc, const uint8_t* __restrict__ alive,
                                   const CleanCtl* ctl, const int* enabled, LoessDev d) {
    if (!*enabled || !d.lc->ok) return;
    __shared__ double f[2 * GC_BINS + 2];
    const int flen = d.lc->flen, min_gc = d.lc->min_gc;
    for (int k = threadIdx.x; k < flen; k += blockDim.x) f[k] = d.lc->f[k];
    __syncthreads();
    const double med = d.lc->med[0];
    const int n = ctl->n2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!alive[i]) continue;
        int idx = (int)gc[i] - min_gc;
        if (idx > flen - 1) idx = flen - 1;
        if (idx < 0) idx = 0;
        const double smoothed = log((double)count[i]) - f[idx] + med;
        count[i] = (float)exp(smoothed);
    }
}

inline size_t loess_workspace_bytes(int64_t n) {
    const int64_t nt = n / LO_TILE + 2;
    return arena_need(n, 8) * 3 + arena_need(n, 1) * 2 + arena_need(nt * LO_KEYS, 4) * 2 + arena_need(nt + 2, 8) * 2 +
           arena_need(1, sizeof(LoessCtl)) + arena_need(256, 1) + (1 << 20);
}

inline bool loess_alloc(cg_ctx* ctx, int64_t n, LoessDev& d) {
    const int64_t nt = n / LO_TILE + 2;
    d.ntiles = (int)((n + LO_TILE - 1) / LO_TILE);
    d.y = arena_take<double>(ctx, n);
    bool ok = d.y != nullptr;
    for (int s = 0; s < 2; s++) {
        d.key[s] = arena_take<uint8_t>(ctx, n);
        d.thist[s] = arena_take<unsigned>(ctx, nt * LO_KEYS);
        d.sorted[s] = arena_take<double>(ctx, n);
        d.tbase[s] = arena_take<double>(ctx, nt + 2);
        ok = ok && d.key[s] && d.thist[s] && d.sorted[s] && d.tbase[s];
    }
    d.lc = arena_take<LoessCtl>(ctx, 1);
    ok = ok && d.lc && sel_state_alloc<uint64_t>(ctx, 2, d.sel);
    return ok;
}

// one LoessGCNormalizer.Normalize() over the alive bins of count2, gated by *enabled
inline void loess_enqueue(cg_ctx* ctx, LoessDev& d, float* count2, const uint8_t* gc2, const uint8_t* chrom2, const uint8_t* alive,
                          const CleanCtl* ctl, const int* enabled, int n, int grid_stream) {
    const int nt = std::max(1, d.ntiles);
    cudaMemsetAsync(d.sel.hist, 0, (size_t)2 * SEL_G * SEL_BINS * sizeof(unsigned), ctx->stream);
    CG_LAUNCH(ctx, loess_prep_kernel, grid_stream, 256, 0, count2, gc2, chrom2, alive, d.is_chry, ctl, enabled, d);
    CG_LAUNCH(ctx, loess_tile_hist_kernel, nt, LO_TILE, 0, ctl, enabled, d);
    CG_LAUNCH(ctx, loess_offsets_kernel, 1, 256, 0, ctl, enabled, d);
    CG_LAUNCH(ctx, loess_scatter_kernel, nt, LO_TILE, 0, ctl, enabled, d);
    CG_LAUNCH(ctx, loess_scan_tiles_kernel, dim3(nt, 2), LO_TILE, 0, enabled, d);
    CG_LAUNCH(ctx, loess_scan_totals_kernel, 1, 32, 0, enabled, d);
    LoessYView yv{d.y, d.key[0], d.key[1], ctl, enabled};
    CG_LAUNCH(ctx, loess_median_request_kernel, 1, 32, 0, d.sel, d.lc, enabled);
    sel_run_scatter<uint64_t, LoessYView>(ctx, yv, d.sel, n);
    CG_LAUNCH(ctx, loess_median_finish_kernel, 1, 32, 0, d.sel, d.lc, enabled);
    CG_LAUNCH(ctx, loess_search_kernel, 1, 128, 0, enabled, d);
    CG_LAUNCH(ctx, loess_apply_kernel, grid_stream, 256, 0, count2, gc2, alive, ctl, enabled, d);
}
