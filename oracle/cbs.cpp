// ORACLE — test infrastructure only (never linked into or called by the product path).
// CPU restatement of CanvasPartition's circular binary segmentation (a port of DNAcopy in the reference):
//   CBSRunner.Run                    CBSRunner.cs:40-151
//   ChangePoints / FindChangePoints  ChangePoint.cs:44-153, :291-400   XPerm :407-421   TrimmedVariance :423-452
//   TMaxO / HTMaxP / TMaxP / TPermP  CBSTStatistic.cs:19-341, :354-586, :599-934, :947-1024
//   TailP / Nu                       TailProbability.cs:21-100
//   ComputeBoundary                  GetBoundary.cs:19-160
// dataType is always "logratio" on this path (CBSRunner.cs:118), so the binary branches are left out.
//
// PARITY UNPINNED: the reference has no CBS test.  Third-party pieces restated from their published
// algorithms: MathNet.Numerics 3.17 MersenneTwister (MT19937 init_genrand / genrand_int32, NextDouble =
// int32 * 2^-32, NextFullRangeInt32 = the raw 32-bit draw), Normal CDF (0.5 erfc(-x/sqrt 2)), BinomialLn
// (lgamma), R's phyper (direct summation of the mass function).
#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "oracle.h"

namespace {

struct Mt {
    uint32_t s[624];
    int at;
    explicit Mt(uint32_t seed) {
        s[0] = seed;
        for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        at = 624;
    }
    uint32_t u32() {
        if (at >= 624) {
            for (int k = 0; k < 624; k++) {
                uint32_t y = (s[k] & 0x80000000u) | (s[(k + 1) % 624] & 0x7fffffffu);
                s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            at = 0;
        }
        uint32_t y = s[at++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double unit() { return (double)u32() * (1.0 / 4294967296.0); }
};

inline int round_even(double v) { return (int)std::nearbyint(v); }
inline double pnorm(double x) { return 0.5 * std::erfc(-x / std::sqrt(2.0)); }

// ---------------------------------------------------------------- TailProbability.cs
double nu(double x, double tol) {
    double l1;
    if (x > 0.01) {
        l1 = std::log(2.0) - 2 * std::log(x);
        double l0 = l1, dk = 0;
        int k = 2;
        for (int i = 0; i < k; i++) { dk += 1; l1 -= 2.0 * pnorm(-x * std::sqrt(dk) / 2.0) / dk; }
        while (std::fabs((l1 - l0) / l1) > tol) {
            l0 = l1;
            for (int i = 0; i < k; i++) { dk += 1; l1 -= 2.0 * pnorm(-x * std::sqrt(dk) / 2.0) / dk; }
            k *= 2;
        }
    } else {
        l1 = -0.583 * x;
    }
    return std::exp(l1);
}

double integral_inv_t1t_sq(double x, double a) {
    double y = x + a - 0.5;
    double v = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * std::log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    y = x - 0.5;
    v = v - (8.0 * y) / (1.0 - 4.0 * y * y) - 2.0 * std::log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    return v;
}

double tail_p(double b, double delta, int m, int ngrid, double tol) {
    const double dincr = (0.5 - delta) / ngrid, bsqrtm = b / std::sqrt((double)m);
    double tl = 0.5 - dincr, t = 0.5 - 0.5 * dincr, acc = 0.0;
    for (int i = 0; i < ngrid; i++) {
        tl += dincr;
        t += dincr;
        const double v = nu(bsqrtm / std::sqrt(t * (1 - t)), tol);
        acc += v * v * integral_inv_t1t_sq(tl, dincr);
    }
    acc = 9.973557E-2 * (b * b * b) * std::exp(-(b * b) / 2) * acc;
    return 2.0 * acc;
}

// ---------------------------------------------------------------- GetBoundary.cs
inline double lchoose(int n, int k) {
    if (k < 0 || k > n) return -INFINITY;
    return std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0);
}

// P(X <= k), X ~ hypergeometric(white = n1s, black = dn, draws = i)
double phyper_lower(int k, int white, int black, int draws) {
    const int lo = std::max(0, draws - black), hi = std::min(draws, white);
    if (k < lo) return 0.0;
    if (k >= hi) return 1.0;
    double term = std::exp(lchoose(white, lo) + lchoose(black, draws - lo) - lchoose(white + black, draws));
    double sum = term;
    for (int x = lo; x < k; x++) {
        term *= (double)(white - x) * (double)(draws - x) / ((double)(x + 1) * (double)(black - draws + x + 1));
        sum += term;
    }
    return std::min(sum, 1.0);
}

void eta_boundary(uint32_t nperm, double eta0, uint32_t n1s, std::vector<uint32_t>& sb, uint32_t off) {
    uint32_t k = 0;
    for (uint32_t i = 1; i <= nperm; i++) {
        if (phyper_lower((int)k, (int)n1s, (int)(nperm - n1s), (int)i) <= eta0) {
            sb[off + k] = i;
            k++;
        }
    }
}

double p_exceed(uint32_t nperm, uint32_t n1s, const std::vector<uint32_t>& sb, uint32_t off) {
    const int N = (int)nperm;
    const double dl = lchoose(N, (int)n1s);
    double p = std::exp(lchoose((int)(nperm - sb[off]), (int)n1s) - dl);
    if (n1s >= 2) {
        const int n1 = (int)sb[off], n = (int)(nperm - sb[off + 1]), k = (int)n1s - 1;
        p += std::exp(std::log((double)n1) + lchoose(n, k) - dl);
    }
    if (n1s >= 3) {
        const int n1 = (int)sb[off], n2 = (int)sb[off + 1], n = (int)(nperm - sb[off + 2]), k = (int)n1s - 2;
        p += std::exp(std::log((double)n1) + std::log(n1 - 1.0) - std::log(2.0) + lchoose(n, k) - dl) +
             std::exp(std::log((double)n1) + std::log((double)(n2 - n1)) + lchoose(n, k) - dl);
    }
    for (int i = 4; i <= (int)n1s; i++) {
        const int n1 = (int)sb[off + i - 4], n2 = (int)sb[off + i - 3], n3 = (int)sb[off + i - 2];
        const int n = (int)(nperm - sb[off + i - 1]), k = (int)n1s - i + 1;
        const double c = lchoose(n, k) - dl;
        p += std::exp(lchoose(n1, i - 1) + c) + std::exp(lchoose(n1, i - 2) + std::log((double)(n3 - n1)) + c) +
             std::exp(lchoose(n1, i - 3) + std::log((double)(n2 - n1)) + std::log((double)(n3 - n2)) + c) +
             std::exp(lchoose(n1, i - 3) + std::log((double)(n2 - n1)) - std::log(2.0) + std::log(n2 - n1 - 1.0) + c);
    }
    return p;
}

std::vector<uint32_t> compute_boundary(uint32_t nperm, double alpha, double eta, double tol = 1e-2) {
    const uint32_t max_ones = (uint32_t)(std::floor(nperm * alpha) + 1);
    std::vector<uint32_t> sb((size_t)max_ones * (max_ones + 1) / 2, 0u);
    uint32_t l = 0;
    sb[0] = nperm - (uint32_t)(nperm * eta);
    double eta0 = eta;
    for (uint32_t j = 2; j <= max_ones; j++) {
        double hi = eta0 * 1.1;
        eta_boundary(nperm, hi, j, sb, l + 1);
        double p_hi = p_exceed(nperm, j, sb, l + 1);
        double lo = eta0 * 0.25;
        eta_boundary(nperm, lo, j, sb, l + 1);
        double p_lo = p_exceed(nperm, j, sb, l + 1);
        while ((hi - lo) / lo > tol) {
            eta0 = lo + (hi - lo) * (eta - p_lo) / (p_hi - p_lo);
            eta_boundary(nperm, eta0, j, sb, l + 1);
            const double p = p_exceed(nperm, j, sb, l + 1);
            if (p > eta) { hi = eta0; p_hi = p; } else { lo = eta0; p_lo = p; }
        }
        l += j;
    }
    return sb;
}

// ---------------------------------------------------------------- CBSTStatistic.cs
struct Blocks {
    int nb;
    std::vector<int> bb, imin, imax;     // 1-based positions like the reference
    std::vector<double> pmin, pmax;
    double gmin = 0, gmax = 0;
    int igmin, igmax;
};

// prefix sums S_p (sx[p-1]) in sqrt(n) blocks with block / global extremes (TMaxO :45-113, TMaxP :622-690)
void block_prefix(const double* x, int n, double* sx, Blocks& B) {
    const double rn = (double)n;
    B.nb = n >= 50 ? round_even(std::sqrt((double)n)) : 1;
    B.bb.resize(B.nb); B.imin.resize(B.nb); B.imax.resize(B.nb); B.pmin.resize(B.nb); B.pmax.resize(B.nb);
    for (int i = 0; i < B.nb; i++) B.bb[i] = round_even(rn * ((i + 1.0) / B.nb));
    B.gmin = B.gmax = 0;
    B.igmin = B.igmax = n;
    int ilo = 1;
    double psum = 0;
    for (int b = 0; b < B.nb; b++) {
        sx[ilo - 1] = psum + x[ilo - 1];
        double mn = sx[ilo - 1], mx = mn;
        int imn = ilo, imx = ilo;
        for (int p = ilo + 1; p <= B.bb[b]; p++) {
            sx[p - 1] = sx[p - 2] + x[p - 1];
            if (sx[p - 1] < mn) { mn = sx[p - 1]; imn = p; }
            if (sx[p - 1] > mx) { mx = sx[p - 1]; imx = p; }
        }
        B.imin[b] = imn; B.imax[b] = imx; B.pmin[b] = mn; B.pmax[b] = mx;
        if (mn < B.gmin) { B.gmin = mn; B.igmin = imn; }
        if (mx > B.gmax) { B.gmax = mx; B.igmax = imx; }
        psum = sx[B.bb[b] - 1];
        ilo = B.bb[b] + 1;
    }
}

// max over arcs of rn / (L (rn - L)) (S_j - S_i)^2 with the block pruning of tmaxo/tmaxp; `seg` gets the arc
// (observed data only).  Returns the t statistic.
double tmax_search(const double* x, int n, double tss, double* sx, int al0, int* seg) {
    const double rn = (double)n;
    Blocks B;
    block_prefix(x, n, sx, B);
    const double psdiff = B.gmax - B.gmin;
    double rj = (double)std::abs(B.igmax - B.igmin);
    double bssmax = rn / (rj * (rn - rj)) * (psdiff * psdiff);
    int ti = std::min(B.igmax, B.igmin), tj = std::max(B.igmax, B.igmin);
    if (seg && psdiff <= 0) {
        bssmax = 0;
    } else {
        const double rnov2 = rn / 2;
        const int nal0 = n - al0;
        struct Cand { int bi, bj, alen; double lim; };
        std::vector<Cand> cand;
        std::vector<double> key;
        auto lo_of = [&](int b) { return b == 1 ? 1 : B.bb[b - 2] + 1; };
        for (int bi = 1; bi <= B.nb; bi++)
            for (int bj = bi; bj <= B.nb; bj++) {
                const int ilo = lo_of(bi), ihi = B.bb[bi - 1], jlo = lo_of(bj), jhi = B.bb[bj - 1];
                int alenhi = jhi - ilo;
                if (alenhi > nal0) alenhi = nal0;
                int alenlo = bi == bj ? 1 : jlo - ihi;
                if (alenlo < al0) alenlo = al0;
                const double s1 = std::fabs(B.pmax[bj - 1] - B.pmin[bi - 1]);
                const double s2 = std::fabs(B.pmax[bi - 1] - B.pmin[bj - 1]);
                const double smx = std::max(s1, s2);
                const double rlo = (double)alenlo, rhi = (double)alenhi;
                const double lim = rn / std::min(rlo * (rn - rlo), rhi * (rn - rhi)) * (smx * smx);
                if (bssmax <= lim) {
                    Cand c{bi, bj, 0, lim};
                    double s;
                    if (s1 > s2) { c.alen = std::abs(B.imax[bj - 1] - B.imin[bi - 1]); s = s1; }
                    else { c.alen = std::abs(B.imin[bj - 1] - B.imax[bi - 1]); s = s2; }
                    const double r = (double)c.alen;
                    key.push_back(rn / (r * (rn - r)) * (s * s));
                    cand.push_back(c);
                }
            }
        // Array.Sort(keys, items): ascending; equal keys only change the visiting order, not the maximum
        std::vector<int> order(cand.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const double ka = key[a], kb = key[b];
            if (std::isnan(ka)) return !std::isnan(kb);
            return ka < kb;
        });
        for (int l = (int)order.size() - 1; l >= 0; l--) {
            const Cand& c = cand[order[l]];
            if (!(bssmax <= c.lim)) continue;
            const int ilo = lo_of(c.bi), ihi = B.bb[c.bi - 1], jlo = lo_of(c.bj), jhi = B.bb[c.bj - 1];
            int alenhi = jhi - ilo;
            if (alenhi > nal0) alenhi = nal0;
            int alenlo = c.bi == c.bj ? 1 : jlo - ihi;
            if (alenlo < al0) alenlo = al0;
            int alenmax = c.alen;
            if (alenmax > n - alenmax) alenmax = n - alenmax;
            auto scan = [&](int L) {
                const int ixlo = std::max(0, jlo - ilo - L), ixhi = std::max(0, ihi + L - jhi);
                double mx = 0;
                int at = ilo + ixlo - 1;
                for (int i = ilo + ixlo; i <= ihi - ixhi; i++) {
                    const double a = std::fabs(sx[i + L - 1] - sx[i - 1]);
                    if (mx < a) { mx = a; at = i; }
                }
                const double r = (double)L;
                const double v = rn / (r * (rn - r)) * (mx * mx);
                if (v > bssmax) { bssmax = v; ti = at; tj = at + L; }
            };
            if ((double)alenlo <= rnov2 && alenlo <= alenmax)
                for (int L = alenlo; L <= alenmax; L++) scan(L);
            alenmax = n - alenmax;
            if ((double)alenhi >= rnov2 && alenhi >= alenmax)
                for (int L = alenhi; L >= alenmax; L--) scan(L);
        }
    }
    if (tss <= bssmax + 0.0001) tss = bssmax + 1.0;
    if (seg) { seg[0] = ti; seg[1] = tj; }
    return bssmax / ((tss - bssmax) / (rn - 2.0));
}

// max t over arcs no longer than k on permuted data (CBSTStatistic.cs:354-586)
double htmaxp(int k, double tss, const double* px, int n, double* sx, int al0) {
    const double rn = (double)n;
    const int nb = (int)(rn / k);
    std::vector<double> bmax(nb), bmin(nb);
    std::vector<int> bb(nb);
    for (int i = 0; i < nb; i++) bb[i] = round_even(rn * ((double)(i + 1) / nb));
    int ilo = 1;
    double psum = 0, best = 0.0;
    for (int b = 0; b < nb; b++) {
        sx[ilo - 1] = psum + px[ilo - 1];
        double mn = sx[ilo - 1], mx = mn;
        int imn = ilo, imx = ilo;
        for (int i = ilo; i < bb[b]; i++) {
            sx[i] = sx[i - 1] + px[i];
            if (sx[i] < mn) { mn = sx[i]; imn = i + 1; }
            if (sx[i] > mx) { mx = sx[i]; imx = i + 1; }
        }
        bmin[b] = mn; bmax[b] = mx;
        psum = sx[bb[b] - 1];
        ilo = bb[b] + 1;
        const int d = std::abs(imn - imx);
        if (d <= k && d >= al0) {
            const double r = (double)d;
            const double v = rn / (r * (rn - r)) * ((mx - mn) * (mx - mn));
            if (best < v) best = v;
        }
    }
    auto sweep = [&](double psdiff, auto&& arcs) {
        const double sq = psdiff * psdiff;
        for (int j = al0; j <= k; j++) {
            const double r = (double)j, c = rn / (r * (rn - r));
            if (c * sq < best) break;
            const double mx = arcs(j);
            const double v = c * (mx * mx);
            if (best < v) best = v;
        }
    };
    auto within = [&](int lo, int hi) {
        return [&, lo, hi](int j) {
            double mx = 0.0;
            for (int i = lo; i <= hi - j; i++) { const double a = std::fabs(sx[i + j - 1] - sx[i - 1]); if (mx < a) mx = a; }
            return mx;
        };
    };
    sweep(bmax[0] - bmin[0], within(1, bb[0]));
    sweep(std::max(std::fabs(bmax[0] - bmin[nb - 1]), std::fabs(bmax[nb - 1] - bmin[0])), [&](int j) {
        double mx = 0.0;
        const int nmj = n - j;
        for (int i = 0; i < j; i++) { const double a = std::fabs(sx[i + nmj] - sx[i]); if (mx < a) mx = a; }
        return mx;
    });
    for (int l = 1; l < nb; l++) {
        const int lo = bb[l - 1] + 1, hi = bb[l];
        sweep(bmax[l] - bmin[l], within(lo, hi));
        sweep(std::max(std::fabs(bmax[l] - bmin[l - 1]), std::fabs(bmax[l - 1] - bmin[l])), [&](int j) {
            double mx = 0.0;
            for (int i = lo - j; i <= lo - 1; i++) { const double a = std::fabs(sx[i + j - 1] - sx[i - 1]); if (mx < a) mx = a; }
            return mx;
        });
    }
    if (tss <= best + 0.0001) tss = best + 1.0;
    return best / ((tss - best) / (rn - 2.0));
}

void xperm(const double* x, double* px, int n, Mt& rnd) {
    std::memcpy(px, x, sizeof(double) * (size_t)n);
    for (int i = n - 1; i >= 0; i--) {
        int j = (int)(rnd.unit() * (i + 1));
        if (j > i) j = i;
        std::swap(px[i], px[j]);
    }
}

// p-value of the two-sample t statistic used to trim the edges of a ternary split (:947-1024)
double tpermp(int n1, int n2, int n, const double* x, double* px, uint32_t nperm, Mt& rnd, int64_t* perm_steps) {
    const double rn1 = n1, rn2 = n2, rn = rn1 + rn2;
    int nrej;
    if (n1 == 1 || n2 == 1) {
        nrej = (int)nperm;
    } else {
        double s1 = 0, s2 = 0, tss = 0;
        for (int i = 0; i < n1; i++) { px[i] = x[i]; s1 += x[i]; tss += x[i] * x[i]; }
        for (int i = n1; i < n; i++) { px[i] = x[i]; s2 += x[i]; tss += x[i] * x[i]; }
        const double xbar = (s1 + s2) / rn;
        tss = tss - rn * (xbar * xbar);
        int m1;
        double rm1, ostat, tstat;
        if (n1 <= n2) { m1 = n1; rm1 = rn1; ostat = 0.99999 * std::fabs(s1 / rn1 - xbar); tstat = (ostat * ostat) * rn1 * rn / rn2; }
        else { m1 = n2; rm1 = rn2; ostat = 0.99999 * std::fabs(s2 / rn2 - xbar); tstat = (ostat * ostat) * rn2 * rn / rn1; }
        nrej = 0;
        tstat = tstat / ((tss - tstat) / (rn - 2.0));
        if (!(tstat > 25 && m1 >= 10)) {
            for (uint32_t np = 0; np < nperm; np++) {
                double s = 0;
                for (int i = n - 1; i >= n - m1; i--) {
                    int j = (int)(rnd.unit() * (i + 1));
                    if (j > i) j = i;
                    std::swap(px[i], px[j]);
                    s += px[i];
                }
                if (ostat <= std::fabs(s / rm1 - xbar)) nrej++;
            }
            if (perm_steps) *perm_steps += (int64_t)nperm * m1;
        }
    }
    return (double)nrej / nperm;
}

struct CbsStats { int64_t tests = 0, perms = 0, perm_steps = 0, edge_steps = 0; };

// fndcpt (ChangePoint.cs:291-400): 0, 1 or 2 change points of the centred segment x[0..n)
int find_change_points(const double* x, int n, double tss, const ora_cbs_opts& o, bool hybrid, double delta,
                       const std::vector<uint32_t>& sbdry, Mt& rnd, int* out, std::vector<double>& px, std::vector<double>& sx,
                       CbsStats& st) {
    int seg[2];
    st.tests++;
    double ostat = tmax_search(x, n, tss, sx.data(), o.min_width, seg);
    const double ostat1 = std::sqrt(ostat);
    ostat *= 0.99999;
    if (ostat1 <= 0.1) return 0;
    const int l = std::min(seg[1] - seg[0], n - seg[1] + seg[0]);
    if (!(ostat1 >= 7.0 && l >= 10)) {
        int nrej = 0, nrejc, k;
        if (hybrid) {
            const double p1 = tail_p(ostat1, delta, n, 100, 1e-6);
            if (p1 > o.alpha) return 0;
            nrejc = (int)((o.alpha - p1) * o.n_perm);
        } else {
            nrejc = (int)(o.alpha * o.n_perm);
        }
        k = nrejc * (nrejc + 1) / 2 + 1;
        for (uint32_t np = 1; np <= o.n_perm; np++) {
            xperm(x, px.data(), n, rnd);
            st.perms++;
            st.perm_steps += n;
            const double pstat = hybrid ? htmaxp(o.k_max, tss, px.data(), n, sx.data(), o.min_width)
                                        : tmax_search(px.data(), n, tss, sx.data(), o.min_width, nullptr);
            if (ostat <= pstat) { nrej++; k++; }
            if (nrej > nrejc) return 0;
            if (np >= sbdry[k - 1]) break;
        }
    }
    if (seg[1] == n) { out[0] = seg[0]; return 1; }
    if (seg[0] == 0) { out[0] = seg[1]; return 1; }
    int ncp = 0;
    {
        const int n1 = seg[0], n12 = seg[1], n2 = n12 - n1;
        if (tpermp(n1, n2, n12, x, px.data(), o.n_perm, rnd, &st.edge_steps) <= o.alpha) out[ncp++] = seg[0];
    }
    {
        const int n12 = n - seg[0], n2 = n - seg[1], n1 = n12 - n2;
        if (tpermp(n1, n2, n12, x + seg[0], px.data(), o.n_perm, rnd, &st.edge_steps) <= o.alpha) { ncp++; out[ncp - 1] = seg[1]; }
    }
    return ncp;
}

// ChangePoints (ChangePoint.cs:44-153), undo = none: segment lengths of one chromosome
std::vector<int> change_points(const double* g, int n, const ora_cbs_opts& o, const std::vector<uint32_t>& sbdry, Mt& rnd, CbsStats& st) {
    std::vector<int> seg_end{0, n}, locs;
    std::vector<double> cur, px, sx;
    while (seg_end.size() > 1) {
        const size_t k = seg_end.size();
        const int a = seg_end[k - 2], cn = seg_end[k - 1] - a;
        int ncp = 0, cp[2] = {0, 0};
        if (cn >= 2 * o.min_width) {
            cur.assign(g + a, g + a + cn);
            const bool hybrid = o.hybrid && o.n_min < (uint32_t)cn;
            const double delta = hybrid ? (o.k_max + 1.0) / cn : 0.0;
            double mx = cur[0], mn = cur[0];
            for (double v : cur) { if (v > mx) mx = v; if (v < mn) mn = v; }  // NaN never wins, as Enumerable.Max/Min skip it... see header
            if (mx != mn) {
                double s = 0;
                for (double v : cur) s += v;
                const double avg = s / cn;
                double tss = 0;
                for (double& v : cur) { v -= avg; }
                for (double v : cur) tss += 1.0 * v * v;
                px.resize(cn); sx.resize(cn);
                ncp = find_change_points(cur.data(), cn, tss, o, hybrid, delta, sbdry, rnd, cp, px, sx, st);
            }
        }
        if (ncp == 0) { locs.push_back(seg_end[k - 1]); seg_end.pop_back(); }
        else if (ncp == 1) seg_end.insert(seg_end.end() - 1, cp[0] + a);
        else { const int v[2] = {cp[0] + a, cp[1] + a}; seg_end.insert(seg_end.end() - 1, v, v + 2); }
    }
    std::reverse(locs.begin(), locs.end());
    std::vector<int> len;
    int prev = 0;
    for (int e : locs) { len.push_back(e - prev); prev = e; }
    return len;
}

// Normal(0,1).InverseCumulativeDistribution: Newton steps on the erfc form of the distribution function
double qnorm(double p) {
    double x = 0.0;
    for (int it = 0; it < 200; it++) {
        const double f = pnorm(x) - p;
        const double d = std::exp(-0.5 * x * x) / std::sqrt(2.0 * M_PI);
        const double nx = x - f / d;
        if (nx == x) break;
        x = nx;
    }
    return x;
}

// ChangePoint.InflationFactor (:460-474): midpoint rule with 10000 points on the truncated normal
double inflation_factor(double trim) {
    const double a = qnorm(1 - trim);
    const double step = 2 * a / 10000;
    const double from = -a + step / 2, to = a - step / 2;
    std::vector<double> xs(10000);
    const double inc = (to - from) / (10000 - 1);  // Helper.Seq
    xs[0] = from;
    xs[9999] = to;
    for (int i = 1; i < 9999; i++) xs[i] = xs[i - 1] + inc;
    double e = 0.0;
    for (double x : xs) e += (x * x) * (std::exp(-0.5 * x * x) / std::sqrt(2.0 * M_PI));
    e = e * step / (1 - 2 * trim);
    return 1 / e;
}

// ChangePoint.TrimmedVariance (:423-452) over the concatenated finite scores
double trimmed_variance(const std::vector<double>& all, double trim) {
    const long n = (long)all.size();
    std::vector<double> d((size_t)std::max<long>(n - 1, 0));
    for (long i = 0; i + 1 < n; i++) d[i] = std::fabs(all[i + 1] - all[i]);
    const int keep = round_even((1 - 2 * trim) * (double)(n - 1));
    std::sort(d.begin(), d.end());
    double s = 0.0;
    for (int i = 0; i < keep; i++) s += d[i] * d[i];
    return inflation_factor(trim) * s / (2 * keep);
}

double helper_median(const double* x, int a, int b) {  // Helper.Median (Helper.cs:31-44)
    std::vector<double> y(x + a, x + b);
    const int mid = (int)y.size() / 2;
    std::nth_element(y.begin(), y.begin() + mid, y.end());
    double m = y[mid];
    if (y.size() % 2 == 0) {
        const double lo = *std::max_element(y.begin(), y.begin() + mid);
        m = (m + lo) / 2;
    }
    return m;
}

// ChangePointsSDUndo (:155-196)
std::vector<int> sd_undo(const double* g, const std::vector<int>& len, double trimmed_sd, double change_sd) {
    change_sd *= trimmed_sd;
    std::vector<int> locs;
    int at = 0;
    for (int l : len) { at += l; locs.push_back(at); }
    while (locs.size() > 1) {
        const size_t k = locs.size();
        std::vector<double> med(k);
        for (size_t i = 0; i < k; i++) med[i] = helper_median(g, i == 0 ? 0 : locs[i - 1], locs[i]);
        size_t imin = 0;
        double mn = std::fabs(med[1] - med[0]);
        for (size_t i = 1; i + 1 < k; i++) {
            const double dv = std::fabs(med[i + 1] - med[i]);
            if (dv < mn) { mn = dv; imin = i; }
        }
        if (mn < change_sd) locs.erase(locs.begin() + (long)imin); else break;
    }
    std::vector<int> out;
    int prev = 0;
    for (int e : locs) { out.push_back(e - prev); prev = e; }
    return out;
}

// ChangePointsPrune (ChangePoint.cs:205-271) with Prune.ErrorSumOfSquares / Prune.Combination (Prune.cs:18-76):
// for j = K-1 .. 1 change points, every j-subset of the K found ones is scored (AS 88 enumeration, the last of
// equally good subsets wins through `<=`); stops at the first j whose best residual exceeds (1 + cutoff) times the
// full model's and returns the best subset of size j + 1.  Falling through the loop returns NO change point (:264).
double prune_errssq(const std::vector<int>& len, const std::vector<double>& sx, int k, const std::vector<int>& loc) {
    double ess = 0.0, segsx = 0.0;
    int segnx = 0;
    for (int i = 0; i < loc[0]; i++) { segsx += sx[i]; segnx += len[i]; }
    ess += std::pow(segsx, 2) / segnx;
    for (int j = 1; j < k; j++) {
        segsx = 0.0;
        segnx = 0;
        for (int i = loc[j - 1]; i < loc[j]; i++) { segsx += sx[i]; segnx += len[i]; }
        ess += std::pow(segsx, 2) / segnx;
    }
    segsx = 0.0;
    segnx = 0;
    for (int i = loc[k - 1]; i < (int)len.size(); i++) { segsx += sx[i]; segnx += len[i]; }
    ess += std::pow(segsx, 2) / segnx;
    return ess;
}

void prune_combination(int r, int nmr, std::vector<int>& loc, bool& rleft) {
    int i = r - 1;
    while (loc[i] == nmr + i + 1) i--;
    loc[i]++;
    for (int j = i + 1; j < r; j++) loc[j] = loc[j - 1] + 1;
    if (loc[0] == nmr + 1) rleft = false;
}

std::vector<int> prune_undo(const double* g, int n, const std::vector<int>& len, double cutoff) {
    const int nseg = (int)len.size(), K = nseg - 1;
    std::vector<double> sx(nseg);
    std::vector<int> loc(K), best_j(K), best_prev(K);
    double ssq = 0.0;
    for (int i = 0; i < n; i++) ssq += std::pow(g[i], 2);
    int at = 0;
    for (int i = 0; i < nseg; i++) {
        double s = 0.0;
        for (int p = at; p < at + len[i]; p++) s += std::pow(g[p], 1);
        sx[i] = s;
        at += len[i];
    }
    for (int i = 0; i < K; i++) { loc[i] = i + 1; best_prev[i] = i + 1; }
    const double wssqk = ssq - prune_errssq(len, sx, K, loc);
    int pruned = 0;
    for (int j = K - 1; j > 0; j--) {
        const int kmj = K - j;
        bool jleft = true;
        for (int i = 0; i < j; i++) { loc[i] = i + 1; best_j[i] = i + 1; }
        double wssqj = ssq - prune_errssq(len, sx, j, loc);
        while (jleft) {
            prune_combination(j, kmj, loc, jleft);
            const double w1 = ssq - prune_errssq(len, sx, j, loc);
            if (w1 <= wssqj) {
                wssqj = w1;
                for (int i = 0; i < j; i++) best_j[i] = loc[i];
            }
        }
        if (wssqj / wssqk > 1 + cutoff) {
            pruned = j + 1;
            for (int i = 0; i < pruned; i++) loc[i] = best_prev[i];
            break;
        }
        for (int i = 0; i < j; i++) best_prev[i] = best_j[i];
    }
    std::vector<int> cum(nseg);
    cum[0] = len[0];
    for (int i = 1; i < nseg; i++) cum[i] = cum[i - 1] + len[i];
    std::vector<int> pts(pruned + 2);
    pts[0] = 0;
    for (int i = 0; i < pruned; i++) pts[i + 1] = cum[loc[i] - 1];
    pts[pruned + 1] = n;
    std::vector<int> out(pruned + 1);
    for (int i = 0; i <= pruned; i++) out[i] = pts[i + 1] - pts[i];
    return out;
}

}  // namespace

extern "C" int ora_cbs_prune(const double* g, int n, const int32_t* len, int n_seg, double cutoff, int32_t* len_out) {
    if (n_seg < 2) return -2;
    std::vector<int> out = prune_undo(g, n, std::vector<int>(len, len + n_seg), cutoff);
    for (size_t i = 0; i < out.size(); i++) len_out[i] = out[i];
    return (int)out.size();
}

extern "C" double ora_cbs_inflation_factor(double trim) { return inflation_factor(trim); }
extern "C" double ora_cbs_trimmed_variance(const double* x, int64_t n, double trim) {
    return trimmed_variance(std::vector<double>(x, x + n), trim);
}

extern "C" int64_t ora_cbs_boundary(uint32_t n_perm, double alpha, double eta, uint32_t* out, int64_t cap) {
    auto sb = compute_boundary(n_perm, alpha, eta);
    for (size_t i = 0; i < sb.size() && (int64_t)i < cap; i++) out[i] = sb[i];
    return (int64_t)sb.size();
}

extern "C" double ora_cbs_tailp(double b, double delta, int m) { return tail_p(b, delta, m, 100, 1e-6); }

extern "C" void ora_mt19937(uint32_t seed, int64_t n, uint32_t* out) {
    Mt m(seed);
    for (int64_t i = 0; i < n; i++) out[i] = m.u32();
}

extern "C" double ora_cbs_tmaxo(const double* x, int n, int al0, int* seg) {
    std::vector<double> c(x, x + n), sx(n);
    double s = 0;
    for (double v : c) s += v;
    const double avg = s / n;
    double tss = 0;
    for (double& v : c) v -= avg;
    for (double v : c) tss += 1.0 * v * v;
    return tmax_search(c.data(), n, tss, sx.data(), al0, seg);
}

extern "C" double ora_cbs_htmaxp(const double* px, int n, int k, double tss, int al0) {
    std::vector<double> sx(n);
    return htmaxp(k, tss, px, n, sx.data(), al0);
}

// CBSRunner.Run (:40-151).  seg_len / seg_mean are written per chromosome at chrom_off[c]; seg_first / seg_last
// are the bin indices (inside the chromosome) of each segment's first and last bin after mapping through the
// finite-value index list (:122-138).  stats[0..3] = tests, permutations, permuted bins, edge-test bins.
extern "C" int ora_partition_cbs(const ora_cbs_opts* o, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom, const int64_t* chrom_off,
                                 const double* coverage, int32_t* n_seg, int32_t* seg_len, double* seg_mean, int32_t* seg_first,
                                 int32_t* seg_last, int64_t* stats, int n_threads) {
    if (o->undo < 0 || o->undo > 2) return -2;
    double trimmed_sd = 0.0;
    if (o->undo == 2) {
        std::vector<double> all;
        for (int c = 0; c < n_chrom; c++)
            for (int64_t i = chrom_off[c]; i < chrom_off[c + 1]; i++)
                if (std::isfinite(coverage[i])) all.push_back(coverage[i]);
        trimmed_sd = std::sqrt(trimmed_variance(all, o->trim));
    }
    std::vector<uint32_t> sb(sbdry, sbdry + n_sbdry);
    Mt seeder(o->seed);
    std::vector<uint32_t> seeds(n_chrom);
    for (int c = 0; c < n_chrom; c++) seeds[c] = seeder.u32();  // NextFullRangeInt32 -> MersenneTwister(int)
    std::vector<CbsStats> st(n_chrom);
    auto one = [&](int c) {
        const double* g = coverage + chrom_off[c];
        const int n = (int)(chrom_off[c + 1] - chrom_off[c]);
        n_seg[c] = 0;
        std::vector<int> ina;
        for (int i = 0; i < n; i++)
            if (std::isfinite(g[i])) ina.push_back(i);
        if (n == 0) return;
        Mt rnd(seeds[c]);
        std::vector<int> len = change_points(g, n, *o, sb, rnd, st[c]);
        if (o->undo == 1 && len.size() > 1) len = prune_undo(g, n, len, o->undo_prune);
        if (o->undo == 2 && len.size() > 1) len = sd_undo(g, len, trimmed_sd, o->undo_sd);
        int lo = 0, cs1 = 0, cs2 = -1;
        for (size_t i = 0; i < len.size(); i++) {
            cs2 += len[i];
            const int64_t at = chrom_off[c] + (int64_t)i;
            seg_len[at] = len[i];
            double s = 0, w = 0;
            for (int p = lo; p < lo + len[i]; p++) { w += 1.0; s += g[p] * 1.0; }
            seg_mean[at] = s / w;
            seg_first[at] = cs1 < (int)ina.size() ? ina[cs1] : -1;
            seg_last[at] = cs2 < (int)ina.size() ? ina[cs2] : -1;
            cs1 += len[i];
            lo += len[i];
        }
        n_seg[c] = (int)len.size();
    };
    // chromosomes are independent tasks (Parallel.ForEach, CBSRunner.cs:149); largest first
    std::vector<int> order(n_chrom);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return chrom_off[a + 1] - chrom_off[a] > chrom_off[b + 1] - chrom_off[b]; });
    std::atomic<int> next{0};
    auto worker = [&]() { for (int k; (k = next.fetch_add(1)) < n_chrom;) one(order[k]); };
    if (n_threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    if (stats) {
        stats[0] = stats[1] = stats[2] = stats[3] = 0;
        for (auto& s : st) { stats[0] += s.tests; stats[1] += s.perms; stats[2] += s.perm_steps; stats[3] += s.edge_steps; }
    }
    return 0;
}
