// Genome-wide scalars of CanvasPartition's wavelet branch and the per-chromosome prefix sums.
//   GetCoverageVariability / reportVariabilityByWindow   Segmentation.cs:309-347
//   FactorOfThreeCoverageVariabilities / triplets        Segmentation.cs:364-429
//   GetEvennessScore / reportScoresByWindow              Segmentation.cs:260-297
//   per-chromosome median / MAD and the threshold        WaveletSegmentation.cs:406-417
#pragma once
#include "wavelet.cuh"

// One view for every order statistic of the partition stage; the segment index selects the array.
struct PartView {
    const double* cov;
    const double* cmad;
    const double* ev10;
    const double* ev100;
    const float* r10;
    const float* r100;
    const double* center;  // per-segment centre for the MAD wave (nullptr otherwise)
    WvSegTable t;
    __device__ bool get(long long i, int seg, uint64_t& key) const {
        double x;
        if (seg < t.base_f3) {
            x = cov[i];
            if (center) x = fabs(x - center[seg]);
        } else if (seg < t.base_ev10) {
            x = cmad[i];
        } else if (seg == t.base_ev10) {
            x = ev10[i];
            if (isnan(x) || isinf(x)) return false;  // Segmentation.cs:291
            x = (double)(float)x;                     // Select(Convert.ToSingle), :265
        } else if (seg == t.base_ev100) {
            x = ev100[i];
            if (isnan(x) || isinf(x)) return false;
        } else if (seg == t.base_r10) {
            x = (double)r10[i];
        } else {
            x = (double)r100[i];
        }
        key = f64_key(x);
        return true;
    }
    // coverage windows / chromosomes and the factor-of-three pools are plain double arrays
    __device__ bool plain(int seg, const double*& base, double& centre, bool& has_centre) const {
        if (seg < t.base_f3) {
            base = cov;
            has_centre = center != nullptr;
            centre = has_centre ? center[seg] : 0.0;
            return true;
        }
        if (seg < t.base_ev10) { base = cmad; has_centre = false; centre = 0.0; return true; }
        return false;
    }
};

// ---------------------------------------------------------------------------------------------
// Integer keys.  The coverage CanvasPartition reads is hundredths / 100 by construction of the .cleaned text (IO.cs:21), so
// the medians and MADs of the coverage windows and chromosomes (Segmentation.cs:309-347, WaveletSegmentation.cs:406-417) are
// selected on the integer hundredths: keys below 2^24, three 8-bit digit passes instead of eight over the doubles, and the
// doubles are rebuilt from the selected integers (h / 100.0 is the very double the text parses to).  MAD wave: the integer
// distance |2h - m2| (m2 = the sum of the two middle hundredths = twice the median) orders the doubles |x - median| except
// INSIDE one distance class, where the bins above and below the median can give two different doubles; the key carries the
// side in its lowest bit and the class that holds the requested rank is resolved from the populations of its two sides
// (CPU prototype: tools/hundredths_select_study.py, profiles/r03r_*).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t WV_HQ_SAT = (1u << 22) - 1;  // hundredths from here on (coverage >= 41943.03) do not take the integer path

struct CovView32 {
    const uint32_t* hq;  // hundredths per bin
    const int* m2;       // per segment: twice the median in hundredths (MAD wave); nullptr in the median wave
    __device__ bool get(long long i, int seg, uint32_t& key) const {
        const uint32_t h = hq[i];
        if (!m2) { key = h; return true; }
        const int t = 2 * (int)h - m2[seg];
        key = ((unsigned)abs(t) << 1) | (t > 0 ? 1u : 0u);
        return true;
    }
    __device__ bool plain32(int seg, const uint32_t*& base, int& twice_centre, bool& has_centre) const {
        base = hq;
        has_centre = m2 != nullptr;
        twice_centre = has_centre ? m2[seg] : 0;
        return true;
    }
};

// coverage (already two-decimal doubles) -> hundredths; anything that is not exactly a non-negative hundredth below the
// saturation bound raises the flag and the caller keeps the double keys
__global__ void wv_hundredths_kernel(const double* __restrict__ cov, long long n, uint32_t* __restrict__ hq, unsigned* __restrict__ bad) {
    bool any_bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = cov[i];
        const double h = rint(__dmul_rn(x, 100.0));
        const bool ok = h >= 0.0 && h < (double)WV_HQ_SAT && __ddiv_rn(h, 100.0) == x && !(x == 0.0 && signbit(x));
        hq[i] = ok ? (uint32_t)h : WV_HQ_SAT;
        any_bad = any_bad || !ok;
    }
    if (__any_sync(0xffffffffu, any_bad) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

struct WvScalarParams {
    WvSegTable t;
    const long long* seg_len;  // [nseg] host-known lengths (0 for the ev/ratio kinds)
    long long f3_cnt[WV_F3_LEVELS];
    long long n_total;
    int window;        // EvennessScoreWindow
    int cv_possible;   // N >= 10 * window  (Segmentation.cs:311)
    int is_germline;
    double mad_factor, thr_lower, thr_upper;
};

__device__ inline void median_pair(unsigned long long n, unsigned long long* k) {
    k[0] = (n & 1ull) ? n / 2 : n / 2 - 1;
    k[1] = n / 2;
}

// Quartile ranks of Utilities.Quartiles (Utilities.cs:361-419): {Q1a, Q1b, Q2a, Q2b, Q3a, Q3b}
__device__ inline void wv_quartile_ranks(unsigned long long n, unsigned long long* k) {
    const unsigned long long mid = n / 2;
    if ((n & 1ull) == 0) {
        k[2] = mid - 1; k[3] = mid;
        const unsigned long long mm = mid / 2;
        if ((mid & 1ull) == 0) { k[0] = mm - 1; k[1] = mm; k[4] = mid + mm - 1; k[5] = mid + mm; }
        else { k[0] = k[1] = mm; k[4] = k[5] = mm + mid; }
    } else {
        k[2] = k[3] = mid;
        if ((n - 1) % 4 == 0) { const unsigned long long q = (n - 1) / 4; k[0] = q - 1; k[1] = q; k[4] = 3 * q; k[5] = 3 * q + 1; }
        else { const unsigned long long q = (n - 3) / 4; k[0] = q; k[1] = q + 1; k[4] = 3 * q + 1; k[5] = 3 * q + 2; }
    }
}

__device__ inline void wv_quartile_values(unsigned long long n, const float* v, float* q) {
    const unsigned long long mid = n / 2;
    if ((n & 1ull) == 0) {
        q[1] = __fdiv_rn(__fadd_rn(v[2], v[3]), 2.0f);
        if ((mid & 1ull) == 0) {
            q[0] = __fdiv_rn(__fadd_rn(v[0], v[1]), 2.0f);
            q[2] = __fdiv_rn(__fadd_rn(v[4], v[5]), 2.0f);
        } else { q[0] = v[0]; q[2] = v[4]; }
    } else {
        q[1] = v[2];
        if ((n - 1) % 4 == 0) {
            q[0] = __fadd_rn(__fmul_rn(v[0], 0.25f), __fmul_rn(v[1], 0.75f));
            q[2] = __fadd_rn(__fmul_rn(v[4], 0.75f), __fmul_rn(v[5], 0.25f));
        } else {
            q[0] = __fadd_rn(__fmul_rn(v[0], 0.75f), __fmul_rn(v[1], 0.25f));
            q[2] = __fadd_rn(__fmul_rn(v[4], 0.25f), __fmul_rn(v[5], 0.75f));
        }
    }
}

// wave 1: medians of coverage windows + chromosomes, of the factor-of-three CMAD pools, and the
// evenness order statistics; wave 2: MADs (centre = wave-1 median); wave 3: the float statistics of
// the per-window MAD/median ratios.
// skip_cov: the coverage segments (windows, chromosomes) are served by the integer-key select (wv_request32_kernel)
__global__ void wv_request_kernel(SelState<uint64_t> st, WvScalarParams p, const WvCtl* ctl, int wave, int skip_cov) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    const WvSegTable& t = p.t;
    st.nreq[s] = 0;
    if (skip_cov && s < t.base_f3) return;
    unsigned long long* k = st.req_k + (size_t)s * SEL_G;
    if (wave == 1) {
        if (s < t.base_f3) {
            long long n = p.seg_len[s];
            if (n > 0) { st.nreq[s] = 2; median_pair((unsigned long long)n, k); }
        } else if (s < t.base_ev10) {
            long long n = p.f3_cnt[s - t.base_f3];
            if (n >= 50) { st.nreq[s] = 2; median_pair((unsigned long long)n, k); }  // Segmentation.cs:393
        } else if (s == t.base_ev10) {
            unsigned n = ctl->ev10_valid;
            if (n >= 2) { st.nreq[s] = 6; wv_quartile_ranks(n, k); }
        } else if (s == t.base_ev100) {
            unsigned n = ctl->ev100_valid;
            if (n >= 1) { st.nreq[s] = 2; median_pair(n, k); }
        }
    } else if (wave == 2) {
        const bool window_seg = s < t.base_chrom;
        const bool chrom_seg = s >= t.base_chrom && s < t.base_f3;
        if ((window_seg && p.cv_possible) || (chrom_seg && !p.cv_possible)) {
            long long n = p.seg_len[s];
            if (n > 0) { st.nreq[s] = 2; median_pair((unsigned long long)n, k); }
        }
    } else {
        if (!p.cv_possible) return;
        if (s == t.base_r10 && p.window > WV_WINDOW_IQR && t.n_w10 >= 2) { st.nreq[s] = 6; wv_quartile_ranks(t.n_w10, k); }
        if (s == t.base_r100 && t.n_w100 >= 1) { st.nreq[s] = 2; median_pair(t.n_w100, k); }
    }
}

// SortedList<double>.Median() of every segment that had a median request
__global__ void wv_median_finish_kernel(SelState<uint64_t> st, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    if (st.nreq[s] != 2) return;
    const uint64_t ka = st.req_key[s * SEL_G + 0], kb = st.req_key[s * SEL_G + 1];
    const double a = f64_unkey(ka), b = f64_unkey(kb);
    out[s] = ka == kb ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
}

// integer path: median requests for the coverage segments only (windows + chromosomes); wave as in wv_request_kernel
__global__ void wv_request32_kernel(SelState<uint32_t> st, WvScalarParams p, int wave) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    const WvSegTable& t = p.t;
    st.nreq[s] = 0;
    if (s >= t.base_f3) return;
    if (wave == 2) {
        const bool window_seg = s < t.base_chrom;
        if (window_seg != (p.cv_possible != 0)) return;
    }
    const long long n = p.seg_len[s];
    if (n > 0) { st.nreq[s] = 2; median_pair((unsigned long long)n, st.req_k + (size_t)s * SEL_G); }
}

// medians from the selected hundredths: SortedList<double>.Median() of the doubles h / 100
__global__ void wv_median_finish32_kernel(SelState<uint32_t> st, double* __restrict__ med, int* __restrict__ m2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    if (st.nreq[s] != 2) return;
    const uint32_t ha = st.req_key[s * SEL_G + 0], hb = st.req_key[s * SEL_G + 1];
    const double a = __ddiv_rn((double)ha, 100.0), b = __ddiv_rn((double)hb, 100.0);
    med[s] = ha == hb ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
    m2[s] = (int)(ha + hb);
}

// MADs from the selected distance classes.  A request ended on key (d << 1 | side); its class holds c0 bins below the median
// (hundredths (m2 - d) / 2) and c1 above ((m2 + d) / 2), and the requested rank is r inside the class: the smaller of the two
// doubles |h / 100 - median| comes first.
__global__ void wv_mad_finish32_kernel(SelState<uint32_t> st, const double* __restrict__ med, const int* __restrict__ m2,
                                       double* __restrict__ mad) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.nseg) return;
    if (st.nreq[s] != 2) return;
    const double centre = med[s];
    const int mm = m2[s];
    double v[2];
    for (int r = 0; r < 2; r++) {
        const uint32_t key = st.req_key[s * SEL_G + r];
        const unsigned* aux = st.req_aux + ((size_t)s * SEL_G + r) * 4;
        const int d = (int)(key >> 1);
        const unsigned c0 = aux[0], c1 = aux[1], rin = aux[2];
        const double v_dn = fabs(__dsub_rn(__ddiv_rn((double)((mm - d) / 2), 100.0), centre));
        const double v_up = fabs(__dsub_rn(__ddiv_rn((double)((mm + d) / 2), 100.0), centre));
        if (c0 == 0u) v[r] = v_up;
        else if (c1 == 0u) v[r] = v_dn;
        else if (v_dn <= v_up) v[r] = rin < c0 ? v_dn : v_up;
        else v[r] = rin < c1 ? v_up : v_dn;
    }
    mad[s] = v[0] == v[1] ? v[0] : __ddiv_rn(__dadd_rn(v[0], v[1]), 2.0);
}

// evenness order statistics are consumed right after wave 1 (the select state is reused by wave 2)
__global__ void wv_evenness_finish_kernel(SelState<uint64_t> st, WvSegTable t, WvCtl* ctl) {
    ctl->evenness_ok = 0;
    ctl->evenness = 0.0;
    const unsigned n10 = ctl->ev10_valid, n100 = ctl->ev100_valid;
    if (n10 < 2 || n100 < 1) return;  // Quartiles()/Median() of an empty list throw; swallowed by WaveletsRunner.cs:56-66
    float v[6], q[3];
    for (int r = 0; r < 6; r++) v[r] = (float)f64_unkey(st.req_key[t.base_ev10 * SEL_G + r]);
    wv_quartile_values(n10, v, q);
    const uint64_t ka = st.req_key[t.base_ev100 * SEL_G + 0], kb = st.req_key[t.base_ev100 * SEL_G + 1];
    const double a = f64_unkey(ka), b = f64_unkey(kb);
    const double median = ka == kb ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
    // Segmentation.cs:268
    ctl->evenness = ((double)__fsub_rn(q[2], q[0]) > 0.015) ? __dmul_rn((double)q[2], 100.0) : __dmul_rn(median, 100.0);
    ctl->evenness_ok = 1;
}

// factor-of-three list (Segmentation.cs:379-401) from the wave-1 medians
__global__ void wv_f3_finish_kernel(WvScalarParams p, const double* __restrict__ med, WvCtl* ctl) {
    double last = 0.0;
    ctl->f3[0] = 0.0;
    int r = 1;
    for (; r <= WV_F3_LEVELS; r++) {
        if (p.f3_cnt[r - 1] < 50) break;
        last = med[p.t.base_f3 + r - 1];
        ctl->f3[r] = last;
    }
    for (; r <= WV_F3_LEVELS; r++) ctl->f3[r] = last;
}

// per-window MAD / median as float (Segmentation.cs:340-346)
__global__ void wv_ratio_kernel(WvSegTable t, const double* __restrict__ med, const double* __restrict__ mad,
                                float* __restrict__ r10, float* __restrict__ r100) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t.base_chrom) return;
    const float r = (float)__ddiv_rn(mad[s], med[s]);
    if (s < t.base_w100) r10[s - t.base_w10] = r;
    else r100[s - t.base_w100] = r;
}

// Wave 3 without the radix passes: the per-window ratio lists are a few hundred floats, so one CTA per list sorts
// the keys in shared memory (bitonic) and writes the requested order statistics where the select engine would
// have left them (quartile ranks of the 10 000-bin windows, median pair of the evenness-size windows).
constexpr int WV_RATIO_SORT_MAX = 4096;
__global__ void __launch_bounds__(1024) wv_ratio_stats_kernel(unsigned long long* __restrict__ ratio_keys, WvScalarParams p,
                                                              const float* __restrict__ r10, const float* __restrict__ r100) {
    __shared__ unsigned long long s_k[WV_RATIO_SORT_MAX];
    const WvSegTable& t = p.t;
    const bool ten = blockIdx.x == 0;
    const int n = ten ? t.n_w10 : t.n_w100;
    const float* src = ten ? r10 : r100;
    if (!p.cv_possible) return;
    if (ten && !(p.window > WV_WINDOW_IQR && n >= 2)) return;
    if (!ten && n < 1) return;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) s_k[i] = i < n ? f64_key((double)src[i]) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = s_k[i], b = s_k[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_k[i] = b; s_k[l] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        unsigned long long rk[SEL_G];
        int nr;
        if (ten) { nr = 6; wv_quartile_ranks((unsigned long long)n, rk); }
        else { nr = 2; median_pair((unsigned long long)n, rk); }
        for (int r = 0; r < nr; r++) ratio_keys[(ten ? 0 : 6) + r] = s_k[rk[r] < (unsigned long long)n ? rk[r] : n - 1];
    }
}

// wave 3 through the select engine (more windows than wv_ratio_stats_kernel sorts): its keys to where the CV step reads them
__global__ void wv_ratio_keys_from_select_kernel(SelState<uint64_t> st, WvSegTable t, unsigned long long* __restrict__ ratio_keys) {
    const int r = threadIdx.x;
    if (r < 6) ratio_keys[r] = st.req_key[(size_t)t.base_r10 * SEL_G + r];
    else if (r < 8) ratio_keys[r] = st.req_key[(size_t)t.base_r100 * SEL_G + (r - 6)];
}

// CV decision (Segmentation.cs:309-327) and per-chromosome thresholds (WaveletSegmentation.cs:406-417)
__global__ void wv_cv_sigma_kernel(const unsigned long long* __restrict__ ratio_keys, WvScalarParams p, const double* __restrict__ med,
                                   const double* __restrict__ mad, const long long* __restrict__ off, WvCtl* ctl,
                                   double* __restrict__ sigma, double* __restrict__ cand_thr) {
    __shared__ double s_cv;
    __shared__ int s_has;
    const WvSegTable& t = p.t;
    if (threadIdx.x == 0) {
        int has = 0;
        double cv = 0.0;
        if (p.cv_possible) {
            has = 1;
            bool done = false;
            if (p.window > WV_WINDOW_IQR && t.n_w10 >= 2) {
                float v[6], q[3];
                for (int r = 0; r < 6; r++) v[r] = (float)f64_unkey(ratio_keys[r]);
                wv_quartile_values(t.n_w10, v, q);
                if ((double)__fdiv_rn(__fsub_rn(q[2], q[0]), q[1]) > 0.015) { cv = (double)q[0]; done = true; }
            }
            if (!done) {
                if (t.n_w100 >= 1) {
                    const float a = (float)f64_unkey(ratio_keys[6]);
                    const float b = (float)f64_unkey(ratio_keys[7]);
                    cv = (t.n_w100 & 1) ? (double)a : (double)__fdiv_rn(__fadd_rn(a, b), 2.0f);
                } else {
                    cv = 0.0;  // Median of an empty list: unreachable for WGS-shaped input
                }
            }
        }
        ctl->cv = cv;
        ctl->cv_has_value = has;
        s_cv = cv;
        s_has = has;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < t.n_chrom; c += blockDim.x) {
        const double median = med[t.base_chrom + c];
        const double variability = s_has ? __dmul_rn(median, s_cv) : mad[t.base_chrom + c];
        double thr = __dmul_rn(p.mad_factor, variability);
        if (thr < p.thr_lower) thr = p.thr_lower;
        if (thr > p.thr_upper) thr = p.thr_upper;
        sigma[c] = thr;
        const double n = (double)(off[c + 1] - off[c]);
        const double wmin = p.is_germline ? 0.8 : 1.0;
        // superset of the final test |coef| <= 2*sigma*w*sqrt(2 ln n), w > 0.8 (germline) or 1
        cand_thr[c] = n > 1 ? 2.0 * thr * wmin * sqrt(2.0 * log(n)) * (1.0 - 1e-9) : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Factor-of-three cascade (Segmentation.cs:404-429): level r holds the middle of every triplet of
// level r-1; the scaled spread (c - a) / 2 / b goes to the level's CMAD pool.
// ---------------------------------------------------------------------------------------------
struct WvF3Level {
    long long src_off[WV_MAX_CHROM];  // offset of chromosome c in the source array
    long long dst_off[WV_MAX_CHROM];  // offset in tmed / cmad
    int cnt[WV_MAX_CHROM];            // triplets of chromosome c at this level
};

__global__ void wv_triplet_kernel(const double* __restrict__ src, double* __restrict__ tmed,
                                  double* __restrict__ cmad, const WvF3Level* __restrict__ lv) {
    const int c = blockIdx.y;
    const int n = lv->cnt[c];
    const double* x = src + lv->src_off[c];
    const long long d0 = lv->dst_off[c];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a = x[3 * (long long)i], b = x[3 * (long long)i + 1], cc = x[3 * (long long)i + 2];
        double t;
        if (a > b) { t = a; a = b; b = t; }
        if (a > cc) { t = a; a = cc; cc = t; }
        if (b > cc) { t = b; b = cc; cc = t; }
        tmed[d0 + i] = b;
        cmad[d0 + i] = __ddiv_rn(__ddiv_rn(__dsub_rn(cc, a), 2.0), b);
    }
}

// ---------------------------------------------------------------------------------------------
// Evenness of one window (Segmentation.cs:281-292): sum over integer depths c = 0..floor(mean) of
// #{bin >= c} / sum(bins).  One block per window: block sum, histogram of floor(bin) clipped to the
// mean, then the same left-to-right accumulation of quotients as the reference.
// ---------------------------------------------------------------------------------------------
struct WvEvWork {
    long long lo;  // first bin (genome index)
    int cnt;       // window - 1 bins (Take(windowSize - 1))
    int out;       // index in ev10 (kind 0) or ev100 (kind 1)
    int kind, pad;
};
constexpr int WV_EV_BINS = 8192;

__device__ inline double block_sum_double(double v) {
    __shared__ double s_w[32];
    __shared__ double s_tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double x = lane < nw ? s_w[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) s_tot = x;
    }
    __syncthreads();
    const double r = s_tot;
    __syncthreads();
    return r;
}

__global__ void wv_evenness_kernel(const double* __restrict__ cov, const WvEvWork* __restrict__ work,
                                   double* __restrict__ ev10, double* __restrict__ ev100, WvCtl* ctl) {
    __shared__ unsigned s_hist[WV_EV_BINS];
    const WvEvWork w = work[blockIdx.x];
    const double* x = cov + w.lo;
    // four loads in flight per thread: a 100 000-bin window is one CTA's work and would otherwise be a chain of
    // dependent L2 round trips
    double part = 0.0;
    {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        const int bd = blockDim.x;
        int i = threadIdx.x;
        for (; i + 3 * bd < w.cnt; i += 4 * bd) { p0 += x[i]; p1 += x[i + bd]; p2 += x[i + 2 * bd]; p3 += x[i + 3 * bd]; }
        for (; i < w.cnt; i += bd) p0 += x[i];
        part = (p0 + p1) + (p2 + p3);
    }
    const double sum = block_sum_double(part);
    const double average = sum / (double)w.cnt;
    double ev;
    if (!(average >= 0.0)) {
        ev = 0.0;  // the loop `for (c = 0; c <= average; c++)` runs zero times (also for NaN)
    } else if (isinf(average)) {
        ev = average;  // the reference never terminates here; reported as a non-finite score
    } else {
        const long long cmax = (long long)floor(average);
        if (cmax < WV_EV_BINS) {
            for (int t = threadIdx.x; t <= (int)cmax; t += blockDim.x) s_hist[t] = 0u;
            __syncthreads();
#pragma unroll 4
            for (int i = threadIdx.x; i < w.cnt; i += blockDim.x) {
                const double v = x[i];
                if (v >= 0.0) {
                    const double f = floor(v);
                    const int t = f >= (double)cmax ? (int)cmax : (int)f;
                    atomicAdd(&s_hist[t], 1u);
                }
            }
            __syncthreads();
            ev = 0.0;
            if (threadIdx.x == 0) {
                unsigned run = 0;
                for (int t = (int)cmax; t >= 0; t--) { run += s_hist[t]; s_hist[t] = run; }  // #{bin >= t}
                for (int t = 0; t <= (int)cmax; t++) ev = __dadd_rn(ev, __ddiv_rn((double)s_hist[t], sum));
            }
        } else {
            // mean depth beyond the shared-memory histogram: closed form (differs from the
            // reference's running sum of quotients only in the last bits)
            double acc = 0.0;
            for (int i = threadIdx.x; i < w.cnt; i += blockDim.x) {
                const double v = x[i];
                if (v >= 0.0) { const double f = floor(v); acc += (f >= (double)cmax ? (double)cmax : f) + 1.0; }
            }
            ev = block_sum_double(acc) / sum;
        }
    }
    if (threadIdx.x == 0) {
        if (w.kind == 0) ev10[w.out] = ev; else ev100[w.out] = ev;
        if (!(isnan(ev) || isinf(ev))) atomicAdd(w.kind == 0 ? &ctl->ev10_valid : &ctl->ev100_valid, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// Per-chromosome inclusive prefix sums with a leading zero: Pz[off[c] + c] = 0,
// Pz[off[c] + c + 1 + i] = x[off[c]] + ... + x[off[c] + i].  Tiles never cross chromosomes.
// ---------------------------------------------------------------------------------------------
struct WvScanTile {
    long long lo;  // genome index of the first element
    int len;       // <= WV_SCAN_TILE
    int c;         // chromosome
};

// Where a scan tile lies: either the host's table (tiles of the final plan), or — fused call, before the host knows the
// cleaned lengths — tile k of chromosome c by position: the tile grid is laid out for the INPUT lengths (tile_c, tile_first)
// and clipped against the device-side offsets of the cleaned list.  Tile boundaries relative to a chromosome's first bin are
// the same in both forms, so the sums are grouped identically.
struct WvScanSrc {
    const WvScanTile* tiles;
    const int* tile_c;
    const int* tile_first;
    const long long* off;
};

__device__ inline WvScanTile wv_scan_tile_get(const WvScanSrc& s, int b) {
    if (s.tiles) return s.tiles[b];
    const int c = s.tile_c[b];
    const long long lo = s.off[c] + (long long)(b - s.tile_first[c]) * WV_SCAN_TILE;
    const long long len = s.off[c + 1] - lo;
    WvScanTile t;
    t.lo = lo;
    t.len = (int)(len < 0 ? 0 : (len > WV_SCAN_TILE ? WV_SCAN_TILE : len));
    t.c = c;
    return t;
}

__global__ void wv_scan_tile_sum_kernel(const double* __restrict__ cov, WvScanSrc src,
                                        double* __restrict__ tsum) {
    const WvScanTile t = wv_scan_tile_get(src, blockIdx.x);
    double part = 0.0;
    for (int i = threadIdx.x; i < t.len; i += blockDim.x) part += cov[t.lo + i];
    const double s = block_sum_double(part);
    if (threadIdx.x == 0) tsum[blockIdx.x] = s;
}

// exclusive scan of the tile sums inside each chromosome (one thread per chromosome; <= a few
// hundred tiles each)
__global__ void wv_scan_tile_offsets_kernel(double* __restrict__ tsum, const int* __restrict__ tile_first, int n_chrom) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chrom) return;
    double run = 0.0;
    for (int t = tile_first[c]; t < tile_first[c + 1]; t++) {
        const double v = tsum[t];
        tsum[t] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(256)
wv_scan_apply_kernel(const double* __restrict__ cov, WvScanSrc src,
                     const double* __restrict__ toff, const long long* __restrict__ off, double* __restrict__ pz) {
    __shared__ double s_w[8];
    const WvScanTile t = wv_scan_tile_get(src, blockIdx.x);
    constexpr int ITEMS = WV_SCAN_TILE / 256;
    const int first = threadIdx.x * ITEMS;
    double v[ITEMS];
    double local = 0.0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int i = first + k;
        v[k] = i < t.len ? cov[t.lo + i] : 0.0;
        local += v[k];
        v[k] = local;
    }
    // block exclusive scan of `local`
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    double wbase = 0.0;
    for (int k = 0; k < w; k++) wbase += s_w[k];
    const double base = toff[blockIdx.x] + wbase + (incl - local);
    const long long p0 = t.lo + t.c + 1;  // index of element t.lo in pz
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int i = first + k;
        if (i < t.len) pz[p0 + i] = base + v[k];
    }
    if (threadIdx.x == 0 && t.lo == off[t.c]) pz[off[t.c] + t.c] = 0.0;
}
