// cg_partition_wavelet / cg_partition_wavelet_shard / cg_clean_partition_wavelet
// (reference WaveletsRunner.Run, WaveletsRunner.cs:52-139).
#include <cmath>
#include <cstdlib>

#include "clean.cuh"
#include "comm.cuh"
#include "wavelet.cuh"
#include <chrono>
#include "wavelet_decompose.cuh"
#include "wavelet_finish.cuh"
#include "wavelet_rqindex.cuh"
#include "wavelet_scalars.cuh"

namespace {

struct WvPlan {
    int n_chrom = 0;
    long long N = 0;
    std::vector<long long> off;
    WvSegTable t{};
    std::vector<long long> seg_len;
    std::vector<SelWork> work;       // select work list (coverage windows, chromosomes, f3 pools, evenness / ratio lists)
    std::vector<int> seg_nwork;      // work items of every segment
    // the same items as two lists: coverage pieces (integer-key select on the main stream) and the rest (factor-of-three
    // pools, evenness and ratio lists: double keys, side stream).  The combined list serves the all-double fallback.
    size_t n_work_cov = 0;           // work[0 .. n_work_cov) are the coverage pieces
    std::vector<WvEvWork> ev_work;
    std::vector<WvScanTile> tiles;
    std::vector<int> tile_first;
    std::vector<int> rq_tfirst;   // first 256-bin index tile of each chromosome
    int rq_ntiles = 0;
    WvF3Level f3lv[WV_F3_LEVELS];
    long long f3_cnt[WV_F3_LEVELS];
    long long f3_off[WV_F3_LEVELS];  // offset of level r's pool in cmad/tmed
    long long f3_total = 0;
    int f3_levels_run = 0;
    int window = 100000;
    int cv_possible = 0;
    std::vector<UhChromPlan> cplan;  // per-chromosome slices of the decomposition's rings and lists
};

constexpr long long SEL_CHUNK = 8192;

void add_work(std::vector<SelWork>& w, int seg, long long lo, long long hi, int seg_b = -1, int seg_c = -1) {
    for (long long a = lo; a < hi; a += SEL_CHUNK) w.push_back(SelWork{{seg, seg_b, seg_c}, 0, a, std::min(hi, a + SEL_CHUNK)});
}

void make_plan(WvPlan& pl, int n_chrom, const int64_t* chrom_off, int window) {
    pl.n_chrom = n_chrom;
    pl.off.assign(chrom_off, chrom_off + n_chrom + 1);
    pl.N = pl.off[n_chrom];
    pl.window = window;
    pl.cv_possible = (pl.N >= 10LL * window) ? 1 : 0;  // Segmentation.cs:311
    // windows: for (idx = 0; idx < len - w; idx += w)  (Segmentation.cs:281, :338)
    auto nwin = [](long long len, long long w) { return len >= 1 ? (len - 1) / w : 0LL; };
    long long n10 = 0, n100 = 0;
    for (int c = 0; c < n_chrom; c++) {
        long long len = pl.off[c + 1] - pl.off[c];
        n10 += nwin(len, WV_WINDOW_IQR);
        n100 += nwin(len, window);
    }
    WvSegTable& t = pl.t;
    t.n_w10 = (int)n10; t.n_w100 = (int)n100; t.n_chrom = n_chrom;
    t.base_w10 = 0; t.base_w100 = t.n_w10; t.base_chrom = t.base_w100 + t.n_w100;
    t.base_f3 = t.base_chrom + n_chrom; t.base_ev10 = t.base_f3 + WV_F3_LEVELS; t.base_ev100 = t.base_ev10 + 1;
    t.base_r10 = t.base_ev100 + 1; t.base_r100 = t.base_r10 + 1; t.nseg = t.base_r100 + 1;
    pl.seg_len.assign(t.nseg, 0);
    pl.work.clear();
    pl.ev_work.clear();
    int s10 = 0, s100 = 0;
    std::vector<long long> cuts;
    for (int c = 0; c < n_chrom; c++) {
        long long o = pl.off[c], len = pl.off[c + 1] - o;
        const int first10 = s10, first100 = s100;
        for (long long idx = 0; idx < len - WV_WINDOW_IQR; idx += WV_WINDOW_IQR) {
            pl.seg_len[t.base_w10 + s10] = WV_WINDOW_IQR;
            pl.ev_work.push_back(WvEvWork{o + idx, WV_WINDOW_IQR - 1, s10, 0, 0});
            s10++;
        }
        for (long long idx = 0; idx < len - window; idx += window) {
            pl.seg_len[t.base_w100 + s100] = window;
            pl.ev_work.push_back(WvEvWork{o + idx, window - 1, s100, 1, 0});
            s100++;
        }
        pl.seg_len[t.base_chrom + c] = len;
        // One pass over the chromosome feeds all three families: cut it at every window boundary, so that a piece lies
        // in at most one window of each size (pieces past the last full window belong to the chromosome only).
        const long long cov10 = (long long)(s10 - first10) * WV_WINDOW_IQR, cov100 = (long long)(s100 - first100) * window;
        cuts.clear();
        cuts.push_back(0);
        cuts.push_back(len);
        for (long long b = WV_WINDOW_IQR; b <= cov10; b += WV_WINDOW_IQR) cuts.push_back(b);
        for (long long b = window; b <= cov100; b += window) cuts.push_back(b);
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        for (size_t q = 0; q + 1 < cuts.size(); q++) {
            const long long a = cuts[q], b = cuts[q + 1];
            if (b <= a || a >= len) continue;
            const int seg10 = a < cov10 ? t.base_w10 + first10 + (int)(a / WV_WINDOW_IQR) : -1;
            const int seg100 = a < cov100 ? t.base_w100 + first100 + (int)(a / window) : -1;
            add_work(pl.work, t.base_chrom + c, o + a, o + std::min(b, len), seg10, seg100);
        }
    }
    pl.n_work_cov = pl.work.size();
    // factor-of-three cascade
    std::vector<long long> cur_len(n_chrom), cur_off(n_chrom);
    for (int c = 0; c < n_chrom; c++) { cur_len[c] = pl.off[c + 1] - pl.off[c]; cur_off[c] = pl.off[c]; }
    long long pool = 0;
    pl.f3_levels_run = 0;
    for (int r = 0; r < WV_F3_LEVELS; r++) {
        long long cnt = 0;
        pl.f3_off[r] = pool;
        for (int c = 0; c < n_chrom; c++) {
            long long m = cur_len[c] / 3;
            pl.f3lv[r].src_off[c] = cur_off[c];
            pl.f3lv[r].dst_off[c] = pool + cnt;
            pl.f3lv[r].cnt[c] = (int)m;
            cur_off[c] = pool + cnt;  // next level reads this level's medians (tmed array)
            cur_len[c] = m;
            cnt += m;
        }
        pl.f3_cnt[r] = cnt;
        pl.seg_len[t.base_f3 + r] = cnt;
        if (cnt > 0) add_work(pl.work, t.base_f3 + r, pool, pool + cnt);
        pool += cnt;
        pl.f3_levels_run = r + 1;
        if (cnt < 50) {  // the reference stops here (Segmentation.cs:393-397)
            for (int q = r + 1; q < WV_F3_LEVELS; q++) { pl.f3_cnt[q] = 0; pl.f3_off[q] = pool; }
            break;
        }
    }
    pl.f3_total = pool;
    // evenness / ratio segments: whole arrays
    if (t.n_w10 > 0) add_work(pl.work, t.base_ev10, 0, t.n_w10);
    if (t.n_w100 > 0) add_work(pl.work, t.base_ev100, 0, t.n_w100);
    if (t.n_w10 > 0) add_work(pl.work, t.base_r10, 0, t.n_w10);
    if (t.n_w100 > 0) add_work(pl.work, t.base_r100, 0, t.n_w100);
    pl.seg_nwork.assign(t.nseg, 0);
    for (const SelWork& wk : pl.work)
        for (int a = 0; a < SEL_WSEG; a++)
            if (wk.seg[a] >= 0) pl.seg_nwork[wk.seg[a]]++;
    // scan tiles
    pl.tiles.clear();
    pl.tile_first.assign(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++) {
        pl.tile_first[c] = (int)pl.tiles.size();
        long long o = pl.off[c], len = pl.off[c + 1] - o;
        for (long long a = 0; a < len; a += WV_SCAN_TILE)
            pl.tiles.push_back(WvScanTile{o + a, (int)std::min<long long>(WV_SCAN_TILE, len - a), c});
    }
    pl.tile_first[n_chrom] = (int)pl.tiles.size();
    pl.rq_tfirst.assign(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++)
        pl.rq_tfirst[c + 1] = pl.rq_tfirst[c] + (int)((pl.off[c + 1] - pl.off[c] + RQ_TILE - 1) / RQ_TILE);
    pl.rq_ntiles = pl.rq_tfirst[n_chrom];
    // decomposition: every chromosome owns a slice of the ring and of each list.  A node of stage M / S / T has more than
    // UH_SMALL_MAX / UH_TINY_MAX / 1 bins and the roots of one stage are disjoint, which bounds the list lengths; big nodes
    // (> UH_MID_MAX bins) are disjoint too, so a small ring holds every big node that can be pending at one time.
    pl.cplan.assign(std::max(n_chrom, 1), UhChromPlan{});
    int ring = 0, mid = 0, small = 0, tiny = 0, cand = 0;
    for (int c = 0; c < n_chrom; c++) {
        const long long len = pl.off[c + 1] - pl.off[c];
        UhChromPlan& q = pl.cplan[c];
        int rc = 8;
        while (rc < 2 * (len / UH_MID_MAX + 2)) rc <<= 1;
        q.ring_base = ring; q.ring_cap = rc; ring += rc;
        q.mid_base = mid; q.mid_cap = (int)(len / UH_SMALL_MAX + 64); mid += q.mid_cap;
        q.small_base = small; q.small_cap = (int)(len / UH_TINY_MAX + 64); small += q.small_cap;
        q.tiny_base = tiny; q.tiny_cap = (int)(len / 2 + 64); tiny += q.tiny_cap;
        q.cand_base = cand; q.cand_cap = (int)(len / 4 + 4096); cand += q.cand_cap;
    }
}

struct WvDev {
    double* cov;
    long long* off;
    unsigned char* selected;
    double* pz;
    // scalars
    SelState<uint64_t> sel;
    SelState<uint32_t> sel32;   // integer-key select over the coverage segments (hundredths)
    uint32_t* hq;               // hundredths of the coverage
    int* m2;                    // per segment: twice the median in hundredths
    unsigned* hq_bad;           // != 0: some coverage value is not a plain non-negative hundredth -> double keys
    unsigned long long* ratio_keys;  // order statistics of the per-window MAD / median ratios: [0..5] 10 000-bin windows, [6..7] evenness-size windows
    long long* seg_len;
    SelWork* work;
    int* seg_nwork;
    WvEvWork* ev_work;
    WvScanTile* tiles;
    int* tile_first;
    double* tsum;
    WvF3Level* f3lv;
    double *tmed, *cmad, *ev10, *ev100;
    float *r10, *r100;
    double *med, *mad, *sigma, *cand_thr, *log3;
    WvCtl* ctl;
    char* plan_end;  // end of the contiguous block of uploaded plan tables that starts at `off`
    // decomposition
    unsigned* lvlcnt;
    int* depth;
    UhBigTask* big;
    UhTask *mid, *small;
    UhTinyTask* tiny;
    int mid_cap, small_cap, tiny_cap;
    UhChromCtl* cc;     // [C] queue counters of the per-chromosome pipelines
    UhChromPlan* cp;    // [C] (plan table, uploaded with the others)
    UhCand* cand;
    int cand_cap;
    // finish scratch
    int *lvl_idx, *sv, *piece, *prelim, *lvl_first;
    unsigned long long* svkey;
    unsigned* bitmap;
    double* rec;
    int *n_bp, *bp;
    // range-quantile index
    unsigned long long *rq_spl, *rq_sorted;
    unsigned short *rq_hist, *rq_tstart;
    unsigned* rq_cum;
    int* rq_tfirst;
    // fused call: the index is built on the side stream from device-side offsets, before the plan tables are uploaded
    long long* rq_off2;
    int* rq_tfirst2;
    unsigned char* rq_sel2;
    // fused call: the prefix sums are computed before the host knows the cleaned lengths (tile grid of the input lengths)
    int* e_tile_c;
    int* e_tile_first;
    unsigned long long* phase_ns;
    unsigned long long* tl_ns;  // debug timeline of the decomposition stages, [C][16]
    unsigned long long* task_dbg;  // debug: per mid task (chromosome, bins, nodes, ns)
    UhTinyTab* tiny_tab;
    int* pack;  // results packed for one download: n_bp[C], depth[C], total, then the breakpoint lists back to back
    // sizes the arrays were allocated for (wv_alloc); a plan run on them must not exceed any
    const long long* alloc_off;  // host: chromosome offsets of the plan the workspace was allocated for (an upper bound of every later plan)
    long long cap_N;
    size_t cap_work, cap_ev, cap_tiles;
    int cap_nseg, cap_f3, cap_w10, cap_w100, cap_rq, cap_C;
};

// Results of one partition run packed for a single download (the host used to issue one copy per chromosome):
// [0, C) n_bp, [C, 2C) tree depth, [2C] total breakpoints, then the lists in chromosome order while they fit.
constexpr int WV_PACK_INTS = 16384;
__global__ void __launch_bounds__(256) wv_pack_kernel(const int* __restrict__ n_bp, const int* __restrict__ depth, const int* __restrict__ bp,
                                                      const long long* __restrict__ off, int C, int* __restrict__ pack,
                                                      const UhChromCtl* __restrict__ cc, WvCtl* ctl) {
    __shared__ int s_at[WV_MAX_CHROM + 1];
    if (threadIdx.x == 32) {
        unsigned long long tot = 0;
        for (int c = 0; c < C; c++) tot += (unsigned long long)cc[c].cand_count_.v;
        ctl->cand_total = tot;
    }
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < C; c++) { s_at[c] = run; run += n_bp[c]; }
        s_at[C] = run;
        pack[2 * C] = run;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) { pack[c] = n_bp[c]; pack[C + c] = depth[c]; }
    __syncthreads();
    const int room = WV_PACK_INTS - (2 * C + 1);
    for (int c = 0; c < C; c++) {
        const int at = s_at[c], n = n_bp[c];
        for (int i = threadIdx.x; i < n && at + i < room; i += blockDim.x) pack[2 * C + 1 + at + i] = bp[off[c] + i];
    }
}

// Sharded runs: this rank's results packed for the all-gather as [length, n_bp[0..C), lists of the chromosomes in order]
// (length = C + total breakpoints; lists beyond the capacity are left out and travel in the exact second round).
__global__ void __launch_bounds__(256) wv_pack_comm_kernel(const int* __restrict__ n_bp, const int* __restrict__ bp,
                                                           const long long* __restrict__ off, int C, int* __restrict__ pack, int cap) {
    __shared__ int s_at[WV_MAX_CHROM + 1];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < C; c++) { s_at[c] = run; run += n_bp[c]; }
        s_at[C] = run;
        pack[0] = C + run;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) pack[1 + c] = n_bp[c];
    __syncthreads();
    const int room = cap - (1 + C);
    for (int c = 0; c < C; c++) {
        const int at = s_at[c], n = n_bp[c];
        for (int i = threadIdx.x; i < n && at + i < room; i += blockDim.x) pack[1 + C + at + i] = bp[off[c] + i];
    }
}

size_t wv_workspace_bytes(const WvPlan& pl) {
    const size_t N = (size_t)pl.N, C = (size_t)pl.n_chrom;
    size_t s = 0;
    s += arena_need(N, 8) + arena_need(C + 1, 8) + arena_need(C + 1, 1) + arena_need(N + C + 1, 8);
    s += sel_state_bytes<uint64_t>(pl.t.nseg) + arena_need(pl.t.nseg, 8);
    s += sel_state_bytes<uint32_t>(pl.t.nseg) + arena_need(N + 1, 4) + arena_need(pl.t.nseg, 4) + arena_need(64, 4) + arena_need(16, 8);
    s += arena_need(pl.t.nseg + 1, 4) + arena_need(pl.work.size() + pl.work.size() / 8 + 64, sizeof(SelWork)) + arena_need(pl.ev_work.size() + 1, sizeof(WvEvWork));
    s += arena_need(pl.tiles.size() + 1, sizeof(WvScanTile)) + arena_need(C + 2, 4) + arena_need(pl.tiles.size() + 1, 8);
    s += arena_need(WV_F3_LEVELS, sizeof(WvF3Level));
    s += arena_need(pl.f3_total + 1, 8) * 2 + arena_need(pl.t.n_w10 + 1, 8) + arena_need(pl.t.n_w100 + 1, 8);
    s += arena_need(pl.t.n_w10 + 1, 4) + arena_need(pl.t.n_w100 + 1, 4);
    s += arena_need(pl.t.nseg, 8) * 2 + arena_need(C + 1, 8) * 2 + arena_need(32, 8) + arena_need(1, sizeof(WvCtl));
    s += arena_need(N + 1, 4) + arena_need(C + 1, 4);
    s += arena_need(UH_QCAP, sizeof(UhBigTask)) + arena_need(N / UH_SMALL_MAX + 64 * C + 64, sizeof(UhTask)) +
         arena_need(N / UH_TINY_MAX + 64 * C + 64, sizeof(UhTask)) + arena_need(N / 2 + 64 * C + 64, sizeof(UhTinyTask));
    s += arena_need(N / 4 + 4096 * C + 4096, sizeof(UhCand));
    s += arena_need(C + 1, sizeof(UhChromCtl)) + arena_need(C + 1, sizeof(UhChromPlan));
    s += arena_need(N + 1, 4) * 5 + arena_need(N + 1, 8) * 2 + arena_need(N / 32 + C + 2, 4);
    s += arena_need(C + 1, 4) + arena_need(N + 1, 4);
    s += arena_need((C + 1) * RQ_BUCKETS, 8) + arena_need(N + 1, 8) + arena_need((size_t)(pl.rq_ntiles + 1) * RQ_BUCKETS, 2) * 2;
    s += arena_need((size_t)(pl.rq_ntiles + C + 2) * RQ_BUCKETS, 4) + arena_need(C + 2, 4) + arena_need((C + 1) * 8, 8) + arena_need((C + 1) * 16, 8) + arena_need(2 * UH_TASK_DBG_CAP + 2, 8) + arena_need(1, sizeof(UhTinyTab)) + arena_need(WV_PACK_INTS, 4) + arena_need(C + 2, 8) + arena_need(C + 2, 4) + arena_need(C + 2, 1);
    s += arena_need(pl.tiles.size() + 1, 4) + arena_need(C + 2, 4);
    return s + (1 << 16);
}

int wv_alloc(cg_ctx* ctx, const WvPlan& pl, WvDev& d, double* cov_dev_existing, uint32_t* hq_existing = nullptr) {
    const size_t N = (size_t)pl.N, C = (size_t)pl.n_chrom;
    d.cov = cov_dev_existing ? cov_dev_existing : arena_take<double>(ctx, N);
    // plan tables uploaded by wv_enqueue: one contiguous block [d.off, d.plan_end) so that a single copy from a
    // pinned staging buffer fills all of them
    d.off = arena_take<long long>(ctx, C + 1);
    d.selected = arena_take<unsigned char>(ctx, C + 1);
    d.seg_len = arena_take<long long>(ctx, pl.t.nseg);
    d.work = arena_take<SelWork>(ctx, pl.work.size() + pl.work.size() / 8 + 64);
    d.seg_nwork = arena_take<int>(ctx, pl.t.nseg + 1);
    d.ev_work = arena_take<WvEvWork>(ctx, pl.ev_work.size() + 1);
    d.tiles = arena_take<WvScanTile>(ctx, pl.tiles.size() + 1);
    d.tile_first = arena_take<int>(ctx, C + 2);
    d.f3lv = arena_take<WvF3Level>(ctx, WV_F3_LEVELS);
    d.rq_tfirst = arena_take<int>(ctx, C + 2);
    d.log3 = arena_take<double>(ctx, 32);
    d.cp = arena_take<UhChromPlan>(ctx, C + 1);
    d.plan_end = ctx->arena + ctx->arena_off;
    d.pz = arena_take<double>(ctx, N + C + 1);
    bool ok = sel_state_alloc<uint64_t>(ctx, pl.t.nseg, d.sel);
    ok = sel_state_alloc<uint32_t>(ctx, pl.t.nseg, d.sel32) && ok;
    d.hq = hq_existing ? hq_existing : arena_take<uint32_t>(ctx, N + 1);
    d.m2 = arena_take<int>(ctx, pl.t.nseg);
    d.hq_bad = arena_take<unsigned>(ctx, 64);
    d.ratio_keys = arena_take<unsigned long long>(ctx, 16);
    ok = ok && d.hq && d.m2 && d.hq_bad && d.ratio_keys;
    d.tsum = arena_take<double>(ctx, pl.tiles.size() + 1);
    d.tmed = arena_take<double>(ctx, pl.f3_total + 1);
    d.cmad = arena_take<double>(ctx, pl.f3_total + 1);
    d.ev10 = arena_take<double>(ctx, pl.t.n_w10 + 1);
    d.ev100 = arena_take<double>(ctx, pl.t.n_w100 + 1);
    d.r10 = arena_take<float>(ctx, pl.t.n_w10 + 1);
    d.r100 = arena_take<float>(ctx, pl.t.n_w100 + 1);
    d.med = arena_take<double>(ctx, pl.t.nseg);
    d.mad = arena_take<double>(ctx, pl.t.nseg);
    d.sigma = arena_take<double>(ctx, C + 1);
    d.cand_thr = arena_take<double>(ctx, C + 1);
    d.ctl = arena_take<WvCtl>(ctx, 1);
    d.lvlcnt = arena_take<unsigned>(ctx, N + 1);
    d.depth = arena_take<int>(ctx, C + 1);
    d.big = arena_take<UhBigTask>(ctx, UH_QCAP);
    // a node of stage M / S / T has more than UH_SMALL_MAX / UH_TINY_MAX / 1 bins and the roots of one
    // stage are disjoint, which bounds the list lengths
    d.mid_cap = (int)(N / UH_SMALL_MAX + 64 * C + 64);
    d.mid = arena_take<UhTask>(ctx, d.mid_cap);
    d.small_cap = (int)(N / UH_TINY_MAX + 64 * C + 64);
    d.small = arena_take<UhTask>(ctx, d.small_cap);
    d.tiny_cap = (int)(N / 2 + 64 * C + 64);
    d.tiny = arena_take<UhTinyTask>(ctx, d.tiny_cap);
    d.cand_cap = (int)(N / 4 + 4096 * C + 4096);
    d.cand = arena_take<UhCand>(ctx, d.cand_cap);
    d.cc = arena_take<UhChromCtl>(ctx, C + 1);
    d.lvl_idx = arena_take<int>(ctx, N + 1);
    d.sv = arena_take<int>(ctx, N + 1);
    d.piece = arena_take<int>(ctx, N + 1);
    d.prelim = arena_take<int>(ctx, N + 1);
    d.lvl_first = arena_take<int>(ctx, N + 1);
    d.svkey = arena_take<unsigned long long>(ctx, N + 1);
    d.rec = arena_take<double>(ctx, N + 1);
    d.bitmap = arena_take<unsigned>(ctx, N / 32 + C + 2);
    d.n_bp = arena_take<int>(ctx, C + 1);
    d.bp = arena_take<int>(ctx, N + 1);
    d.rq_spl = arena_take<unsigned long long>(ctx, (C + 1) * RQ_BUCKETS);
    d.rq_sorted = arena_take<unsigned long long>(ctx, N + 1);
    d.rq_hist = arena_take<unsigned short>(ctx, (size_t)(pl.rq_ntiles + 1) * RQ_BUCKETS);
    d.rq_tstart = arena_take<unsigned short>(ctx, (size_t)(pl.rq_ntiles + 1) * RQ_BUCKETS);
    d.rq_cum = arena_take<unsigned>(ctx, (size_t)(pl.rq_ntiles + C + 2) * RQ_BUCKETS);
    d.phase_ns = arena_take<unsigned long long>(ctx, (C + 1) * 8);
    d.tl_ns = arena_take<unsigned long long>(ctx, (C + 1) * 16);
    d.task_dbg = arena_take<unsigned long long>(ctx, 2 * UH_TASK_DBG_CAP + 2);
    d.pack = arena_take<int>(ctx, WV_PACK_INTS);
    d.rq_off2 = arena_take<long long>(ctx, C + 2);
    d.rq_tfirst2 = arena_take<int>(ctx, C + 2);
    d.rq_sel2 = arena_take<unsigned char>(ctx, C + 2);
    ok = ok && d.rq_off2 && d.rq_tfirst2 && d.rq_sel2;
    d.e_tile_c = arena_take<int>(ctx, pl.tiles.size() + 1);
    d.e_tile_first = arena_take<int>(ctx, C + 2);
    ok = ok && d.e_tile_c && d.e_tile_first;
    d.tiny_tab = arena_take<UhTinyTab>(ctx, 1);
    ok = ok && d.rq_spl && d.rq_sorted && d.rq_hist && d.rq_tstart && d.rq_cum && d.rq_tfirst;
    ok = ok && d.cov && d.off && d.selected && d.pz && d.seg_len && d.work && d.seg_nwork && d.ev_work && d.tiles && d.tile_first &&
         d.tsum && d.f3lv && d.tmed && d.cmad && d.ev10 && d.ev100 && d.r10 && d.r100 && d.med && d.mad && d.sigma &&
         d.cand_thr && d.log3 && d.ctl && d.pack && d.lvlcnt && d.depth && d.big && d.mid && d.small && d.tiny && d.cand && d.cc && d.cp && d.lvl_idx &&
         d.sv && d.piece && d.prelim && d.lvl_first && d.svkey && d.rec && d.bitmap && d.n_bp && d.bp;
    d.alloc_off = pl.off.data();
    d.cap_N = pl.N; d.cap_nseg = pl.t.nseg; d.cap_work = pl.work.size() + pl.work.size() / 8 + 63; d.cap_ev = pl.ev_work.size(); d.cap_tiles = pl.tiles.size();
    d.cap_f3 = pl.f3_total; d.cap_w10 = pl.t.n_w10; d.cap_w100 = pl.t.n_w100; d.cap_rq = pl.rq_ntiles; d.cap_C = pl.n_chrom;
    return ok ? CG_OK : cg_fail(ctx, CG_ERR_CUDA, "partition: device arena exhausted");
}

// Clear the accumulators of one partition run.  Sized by the plan the arrays were allocated for, so that it can
// be enqueued before the final plan is known (fused call: while the host still waits for the Clean results).
// One kernel fills all nine regions (separate memsets cost a few microseconds of stream latency each).
struct WvFillTable {
    void* ptr[12];
    unsigned long long words[12];  // 32-bit words
    unsigned value[12];
    int n;
};

__global__ void __launch_bounds__(256) wv_fill_kernel(WvFillTable t) {
    for (int e = 0; e < t.n; e++) {
        unsigned* p = static_cast<unsigned*>(t.ptr[e]);
        const unsigned long long w = t.words[e];
        const unsigned v = t.value[e];
        const unsigned long long w4 = w >> 2;  // regions start 256-byte aligned: 128-bit stores for the bulk
        uint4* p4 = reinterpret_cast<uint4*>(p);
        const uint4 v4 = make_uint4(v, v, v, v);
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < w4; i += (unsigned long long)gridDim.x * blockDim.x)
            p4[i] = v4;
        if (blockIdx.x == 0 && threadIdx.x < (w & 3ull)) p[(w4 << 2) + threadIdx.x] = v;
    }
}

int wv_clear(cg_ctx* ctx, WvDev& d) {
    const size_t C = (size_t)d.cap_C;
    WvFillTable t;
    t.n = 0;
    auto add = [&](void* p, size_t bytes, unsigned value) {
        t.ptr[t.n] = p; t.words[t.n] = (bytes + 3) / 4; t.value[t.n] = value; t.n++;
    };
    add(d.ctl, sizeof(WvCtl), 0u);
    add(d.sel.hist, (size_t)d.cap_nseg * SEL_G * SEL_BINS * sizeof(unsigned), 0u);
    add(d.sel32.hist, (size_t)d.cap_nseg * SEL_G * SEL_BINS * sizeof(unsigned), 0u);
    add(d.lvlcnt, (size_t)(d.cap_N + 1) * sizeof(unsigned), 0u);
    add(d.depth, (C + 1) * sizeof(int), 0u);
    add(d.big, (size_t)UH_QCAP * sizeof(UhBigTask), 0xffffffffu);
    add(d.cc, (C + 1) * sizeof(UhChromCtl), 0u);
    add(d.n_bp, (C + 1) * sizeof(int), 0u);
    add(d.med, (size_t)d.cap_nseg * 8, 0u);
    add(d.mad, (size_t)d.cap_nseg * 8, 0u);
    static_assert(sizeof(WvCtl) % 4 == 0 && sizeof(UhBigTask) % 4 == 0 && sizeof(UhChromCtl) % 4 == 0, "fill regions are whole 32-bit words");
    CG_LAUNCH(ctx, wv_fill_kernel, ctx->num_sms * 4, 256, 0, t);
    // t_first = all ones (a minimum is taken over it); it lies inside the control block cleared above
    CG_CUDA(ctx, cudaMemsetAsync(&d.ctl->t_first, 0xff, sizeof(unsigned long long), ctx->stream));
    return CG_OK;
}

// The range-quantile index of the finish stage needs only the coverage and the chromosome offsets.
static void wv_enqueue_rq_index(cg_ctx* ctx, WvDev& d, int C, int rq_ntiles_upper, const long long* off, const int* tfirst,
                                const unsigned char* selected) {
    if (rq_ntiles_upper <= 0 || C <= 0) return;
    CG_LAUNCH(ctx, rq_splitter_kernel, C, 1024, 0, d.cov, off, selected, d.rq_spl);
    CG_LAUNCH(ctx, rq_tile_kernel, rq_ntiles_upper, RQ_TILE, 0, d.cov, off, tfirst, C, selected, d.rq_spl, d.rq_hist,
              d.rq_tstart, d.rq_sorted);
    CG_LAUNCH(ctx, rq_cumulate_kernel, C, RQ_BUCKETS, 0, tfirst, selected, d.rq_hist, d.rq_cum);
}

// chromosome offsets and first index tiles from the per-chromosome bin counts, on the device (fused call: lets the index
// be built while the host is still waiting for those counts)
__global__ void wv_device_offsets_kernel(const unsigned* __restrict__ chrom_cnt, int C, long long* __restrict__ off, int* __restrict__ rq_tfirst) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long o = 0;
    int t = 0;
    for (int c = 0; c < C; c++) {
        off[c] = o;
        rq_tfirst[c] = t;
        const long long n = chrom_cnt[c];
        o += n;
        t += (int)((n + RQ_TILE - 1) / RQ_TILE);
    }
    off[C] = o;
    rq_tfirst[C] = t;
}

static inline double wv_now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int wv_enqueue(cg_ctx* ctx, const cg_wavelet_opts* o, const WvPlan& pl, WvDev& d, const unsigned char* selected_host,
               bool rq_index_done = false, int* comm_pack = nullptr, bool use_int = false, bool scan_done = false) {
    cudaStream_t s = ctx->stream;
    const int C = pl.n_chrom;
    const WvSegTable& t = pl.t;
    if (pl.N > d.cap_N || t.nseg > d.cap_nseg || pl.work.size() > d.cap_work || pl.ev_work.size() > d.cap_ev ||
        pl.tiles.size() > d.cap_tiles || pl.f3_total > d.cap_f3 || t.n_w10 > d.cap_w10 || t.n_w100 > d.cap_w100 ||
        pl.rq_ntiles > d.cap_rq || C != d.cap_C)
        return cg_fail(ctx, CG_ERR_CAPACITY, "partition: plan exceeds the workspace it was allocated for");
    d.sel.nseg = t.nseg;
    d.sel32.nseg = t.nseg;
    // ---- plan tables: packed into the pinned staging block with the device layout, one copy
    {
        char* base = (char*)d.off;
        const size_t bytes = (size_t)(d.plan_end - base);
        if (bytes > ctx->plan_pinned_cap) {
            if (ctx->plan_pinned) cudaFreeHost(ctx->plan_pinned);
            ctx->plan_pinned = nullptr;
            ctx->plan_pinned_cap = 0;
            CG_CUDA(ctx, cudaMallocHost((void**)&ctx->plan_pinned, bytes + (bytes >> 2)));
            ctx->plan_pinned_cap = bytes + (bytes >> 2);
        }
        char* h = ctx->plan_pinned;
        auto put = [&](const void* dev_ptr, const void* src, size_t nbytes) {
            if (nbytes) memcpy(h + ((const char*)dev_ptr - base), src, nbytes);
        };
        put(d.off, pl.off.data(), (size_t)(C + 1) * 8);
        put(d.selected, selected_host, (size_t)C);
        put(d.seg_len, pl.seg_len.data(), (size_t)t.nseg * 8);
        put(d.seg_nwork, pl.seg_nwork.data(), (size_t)t.nseg * 4);
        put(d.work, pl.work.data(), pl.work.size() * sizeof(SelWork));
        put(d.ev_work, pl.ev_work.data(), pl.ev_work.size() * sizeof(WvEvWork));
        put(d.tiles, pl.tiles.data(), pl.tiles.size() * sizeof(WvScanTile));
        put(d.tile_first, pl.tile_first.data(), (size_t)(C + 1) * 4);
        put(d.f3lv, pl.f3lv, sizeof(WvF3Level) * WV_F3_LEVELS);
        put(d.rq_tfirst, pl.rq_tfirst.data(), (size_t)(C + 1) * 4);
        put(d.cp, pl.cplan.data(), (size_t)C * sizeof(UhChromPlan));
        // ceil(log(3^k) / log(3)) as the host's libm evaluates it (WaveletSegmentation.cs:224)
        double log3_tab[32] = {0};
        double pw = 1.0;
        for (int k = 0; k < 20; k++) { log3_tab[k] = std::ceil(std::log(pw) / std::log(3.0)); pw *= 3.0; }
        put(d.log3, log3_tab, 20 * 8);
        // on its own stream: in the fused call the main stream is still computing the prefix sums when the host gets here, and
        // a copy queued behind those kernels would hold the order statistics and the pipelines back by its own latency
        static const bool plan_inline = getenv("CANVAS_PLAN_INLINE") != nullptr;
        cudaStream_t ps = plan_inline ? s : ctx->plan_stream;
        CG_CUDA(ctx, cudaMemcpyAsync(base, h, bytes, cudaMemcpyHostToDevice, ps));
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_plan, ps));  // offsets, slices and masks are on the device: the pipelines read them
        if (!plan_inline) CG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_plan, 0));
    }
    if (pl.N == 0) {
        if (comm_pack) CG_LAUNCH(ctx, wv_pack_comm_kernel, 1, 256, 0, d.n_bp, d.bp, d.off, C, comm_pack, ctx->comm->pack_ints);
        return CG_OK;
    }

    WvScalarParams sp;
    sp.t = t;
    sp.seg_len = d.seg_len;
    for (int r = 0; r < WV_F3_LEVELS; r++) sp.f3_cnt[r] = pl.f3_cnt[r];
    sp.n_total = pl.N;
    sp.window = pl.window;
    sp.cv_possible = pl.cv_possible;
    sp.is_germline = o->is_germline;
    sp.mad_factor = o->mad_factor;
    sp.thr_lower = o->thr_lower;
    sp.thr_upper = o->thr_upper;

    CG_TL(ctx, "host wait + plan");
    ctx->host_ts[0] = wv_now_us();
    cudaEventRecord(ctx->stage_ev[2], s);
    ctx->stage_used[1] = true;
    const bool dbg_on = getenv("CANVAS_DEBUG") != nullptr;
    if (dbg_on) {  // debug timelines of the pipelines: cleared before anything that may stamp them
        cudaMemsetAsync(d.phase_ns, 0, (size_t)(C + 1) * 64, s);
        if (d.tl_ns) cudaMemsetAsync(d.tl_ns, 0, (size_t)(C + 1) * 128, s);
        if (d.task_dbg) cudaMemsetAsync(d.task_dbg, 0, 8, s);
    }
    // ---- range-quantile index for the medians of the finish stage: needs only the coverage, and is enqueued first so
    // that the device has work while the host is still launching the many small kernels of the scalars
    if (!rq_index_done) {
        wv_enqueue_rq_index(ctx, d, C, pl.rq_ntiles, d.off, d.rq_tfirst, d.selected);
        CG_TL(ctx, "rq index");
    }
    const int ntiles = (int)pl.tiles.size();
    int max_len = 1;
    for (int c = 0; c < C; c++) max_len = std::max<long long>(max_len, pl.off[c + 1] - pl.off[c]);
    const int nwork = (int)pl.work.size();
    const int rq_grid = div_up(t.nseg, 128);
    PartView pv{d.cov, d.cmad, d.ev10, d.ev100, d.r10, d.r100, nullptr, t};
    auto enqueue_evenness = [&]() {
        if (!pl.ev_work.empty())
            CG_LAUNCH(ctx, wv_evenness_kernel, (int)pl.ev_work.size(), 1024, 0, d.cov, d.ev_work, d.ev10, d.ev100, d.ctl);
    };
    auto enqueue_scan = [&]() {  // (fused call: already done, from device-side offsets, while the host waited for the counts)
        if (ntiles > 0 && !scan_done) {
            const WvScanSrc src{d.tiles, nullptr, nullptr, nullptr};
            CG_LAUNCH(ctx, wv_scan_tile_sum_kernel, ntiles, 256, 0, d.cov, src, d.tsum);
            CG_LAUNCH(ctx, wv_scan_tile_offsets_kernel, div_up(C, 64), 64, 0, d.tsum, d.tile_first, C);
            CG_LAUNCH(ctx, wv_scan_apply_kernel, ntiles, 256, 0, d.cov, src, d.tsum, d.off, d.pz);
        }
    };
    auto enqueue_triplets = [&]() {
        for (int r = 0; r < pl.f3_levels_run; r++) {
            if (pl.f3_cnt[r] == 0) break;
            long long per = std::max<long long>(1, max_len / 3);
            for (int q = 0; q < r; q++) per = std::max<long long>(1, per / 3);
            dim3 grid((unsigned)std::min<long long>(256, (per + 255) / 256), (unsigned)C);
            CG_LAUNCH(ctx, wv_triplet_kernel, grid, 256, 0, r == 0 ? d.cov : d.tmed, d.tmed, d.cmad, d.f3lv + r);
        }
    };
    cudaStream_t pipe = ctx->pipe_stream;
    // ---- decomposition + finish: one pipeline per chromosome, each on its own stream (largest chromosomes first)
    UhParams up;
    up.x = d.cov; up.pz = d.pz; up.off = d.off; up.cand_thr = d.cand_thr; up.lvlcnt = d.lvlcnt;
    up.big = d.big; up.mid = d.mid; up.small = d.small; up.tiny = d.tiny; up.cand = d.cand; up.cc = d.cc; up.cp = d.cp; up.ctl = d.ctl;
    FinParams fp;
    fp.x = d.cov; fp.pz = d.pz; fp.off = d.off; fp.selected = d.selected; fp.cand = d.cand; fp.cc = d.cc; fp.cp = d.cp; fp.ctl = d.ctl;
    fp.lvlcnt = d.lvlcnt; fp.depth = d.depth; fp.sigma = d.sigma; fp.chrom_median = d.med + t.base_chrom;
    fp.log3_scale_tab = d.log3;
    fp.rq.spl = d.rq_spl; fp.rq.hist = d.rq_hist; fp.rq.tstart = d.rq_tstart; fp.rq.cum = d.rq_cum; fp.rq.sorted = d.rq_sorted;
    fp.rq.tfirst = d.rq_tfirst;
    fp.phase_ns = dbg_on ? d.phase_ns : nullptr;
    up.tl_ns = fp.phase_ns && d.tl_ns ? d.tl_ns : nullptr;
    up.task_dbg = up.tl_ns && d.task_dbg ? d.task_dbg : nullptr;
    fp.is_germline = o->is_germline; fp.min_size = o->min_size; fp.n_chrom = C; fp.pad = 0;
    fp.lvl_idx = d.lvl_idx; fp.sv = d.sv; fp.svkey = d.svkey; fp.bitmap = d.bitmap; fp.piece = d.piece; fp.rec = d.rec;
    fp.prelim = d.prelim; fp.lvl_first = d.lvl_first; fp.n_bp = d.n_bp; fp.bp = d.bp;
    {
        const UhChromPlan& last = pl.cplan[C - 1];
        if (last.ring_base + last.ring_cap > UH_QCAP || last.mid_base + last.mid_cap > d.mid_cap || last.small_base + last.small_cap > d.small_cap ||
            last.tiny_base + last.tiny_cap > d.tiny_cap || last.cand_base + last.cand_cap > d.cand_cap)
            return cg_fail(ctx, CG_ERR_CAPACITY, "partition: per-chromosome queues exceed the workspace");
    }
    const size_t uh_smem = sizeof(UhWarpScratch) * (UH_SMALL_THREADS / 32);
    if (!ctx->uh_attrs_set) {
        cudaFuncSetAttribute(uh_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)uh_smem);
        cudaFuncSetAttribute(uh_mid_kernel<1024, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((UH_MID_MAX + 2) * sizeof(double)));
        cudaFuncSetAttribute(uh_mid_kernel<512, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((UH_MID_MAX + 2) * sizeof(double)));
        cudaFuncSetAttribute(uh_mid_kernel<256, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((UH_MID_MAX + 2) * sizeof(double)));
        cudaFuncSetAttribute(uh_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem_bytes());
        ctx->uh_attrs_set = true;
    }
    // The launch sequence of the pipelines depends on the chromosome lengths only through grid sizes and through which
    // stages a chromosome needs; with the lengths the workspace was allocated for (the INPUT lengths in the fused call: the
    // same for every sample binned on the same reference) it is the same for every call of that shape, so it is captured
    // once into a CUDA graph and replayed: one launch instead of ~150, every chromosome's first kernel starts at once, and
    // ranks that share a host no longer queue behind each other's launches.  A shape is captured the second time it is seen.
    static const int mid_threads = getenv("CANVAS_MID_THREADS") ? atoi(getenv("CANVAS_MID_THREADS")) : UH_MID_THREADS;
    const size_t mid_smem = (size_t)(UH_MID_MAX + 2) * sizeof(double);
    // part 0: a chromosome's whole pipeline; 1: seed + chains + mid stage only (need the prefix sums, nothing else);
    // 2: small / tiny stages, depth and finish only (stream-ordered after part 1 by the caller, thresholds known)
    auto enqueue_pipelines = [&](const long long* loff, cudaStream_t root, bool capturing, int part) -> int {
        std::vector<int> order;
        for (int c = 0; c < C; c++) {
            const long long len = loff[c + 1] - loff[c];
            if (selected_host[c] && len > o->min_size && len >= 2) order.push_back(c);
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return loff[a + 1] - loff[a] > loff[b + 1] - loff[b]; });
        const int n_streams = (int)std::min<size_t>(order.size(), CG_CHROM_STREAMS);
        int rc_streams = cg_chrom_streams(ctx, n_streams);
        if (rc_streams) return rc_streams;
        ctx->stream = root;
        if (part != 2) {
            CG_LAUNCH(ctx, uh_seed_kernel, div_up(C, 128), 128, 0, up, d.selected, C, o->min_size);
            CG_LAUNCH(ctx, uh_tiny_table_kernel, UH_TINY_MAX - 1, UH_TINY_MAX, 0, d.tiny_tab);
        }
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_fork2, root));
        for (size_t k = 0; k < order.size(); k++) {
            const int c = order[k];
            const long long len = loff[c + 1] - loff[c];
            cudaStream_t cs = ctx->chrom_streams[k % n_streams];
            if (k < (size_t)n_streams) CG_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_fork2, 0));
            ctx->stream = cs;  // CG_LAUNCH enqueues on ctx->stream
            if (part != 2 && len > UH_MID_MAX) {
                // chains of big nodes: clusters wait on the chromosome's ring, any one of them can finish the work alone
                cudaLaunchConfig_t cfg = {};
                cfg.blockDim = dim3(UH_THREADS);
                cfg.dynamicSmemBytes = 0;
                cfg.stream = cs;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = UH_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                cfg.gridDim = dim3(UH_CLUSTER * (len > 100000 ? 2 : 1));
                cudaError_t le = cudaLaunchKernelEx(&cfg, uh_chain_kernel, up, c);
                if (le != cudaSuccess) { ctx->stream = s; return cg_fail(ctx, CG_ERR_CUDA, std::string("uh_chain_kernel launch: ") + cudaGetErrorString(le)); }
                ctx->launches++;
            }
            if (part != 2 && len > UH_SMALL_MAX)
            {
                if (mid_threads > 0) {  // subtree staged in shared memory (default)
                    const int mid_grid = (int)std::min<long long>(32, std::max<long long>(1, len / 8192));
                    if (mid_threads >= 1024) CG_LAUNCH(ctx, (uh_mid_kernel<1024, 4, true>), mid_grid, 1024, mid_smem, up, c);
                    else if (mid_threads >= 512) CG_LAUNCH(ctx, (uh_mid_kernel<512, 4, true>), mid_grid, 512, mid_smem, up, c);
                    else CG_LAUNCH(ctx, (uh_mid_kernel<256, 4, true>), mid_grid, 256, mid_smem, up, c);
                } else {                // CANVAS_MID_THREADS=0: every node straight from L2 (the earlier form, kept for A/B runs)
                    const int mid_grid = (int)std::min<long long>(64, std::max<long long>(1, len / 2048));
                    CG_LAUNCH(ctx, (uh_mid_kernel<256, 8, false>), mid_grid, 256, 0, up, c);
                }
            }
            if (part == 1) continue;
            // thresholds: recorded on the main stream before this enqueue started (an external event for a captured sequence;
            // part 2 is launched behind stream-level waits instead)
            if (part == 0) CG_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_thr, capturing ? cudaEventWaitExternal : 0));
            if (len > UH_TINY_MAX)
                CG_LAUNCH(ctx, uh_small_kernel, (int)std::min<long long>(ctx->num_sms, std::max<long long>(1, len / 1024)), UH_SMALL_THREADS, uh_smem, up, c);
            CG_LAUNCH(ctx, uh_tiny_kernel, (int)std::min<long long>(ctx->num_sms, std::max<long long>(1, len / 1024)), 128, 0, up, d.tiny_tab, c);
            CG_LAUNCH(ctx, uh_depth_kernel, (int)std::min<long long>(32, std::max<long long>(1, len / 4096)), 256, 0, d.lvlcnt, d.off, d.depth, c);
            // the factor-of-three list comes from the side stream (recorded before this enqueue started)
            if (part == 0 && use_int) CG_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_join, capturing ? cudaEventWaitExternal : 0));
            CG_LAUNCH(ctx, uh_finish_kernel, 1, FIN_THREADS, fin_smem_bytes(), fp, c);
        }
        ctx->stream = s;
        for (int k = 0; k < n_streams; k++) {
            CG_CUDA(ctx, cudaEventRecord(ctx->chrom_ev[k], ctx->chrom_streams[k]));
            CG_CUDA(ctx, cudaStreamWaitEvent(root, ctx->chrom_ev[k], 0));
        }
        return CG_OK;
    };
    auto launch_part = [&](int part) -> int {
        bool replayed = false;
        if (!fp.phase_ns && !getenv("CANVAS_NO_GRAPH")) {
            CgGraphEntry want{};
            unsigned long long hl = 1469598103934665603ull, hs = 1469598103934665603ull;  // FNV-1a of the launch lengths / the mask
            for (int c = 0; c <= C; c++) { hl ^= (unsigned long long)d.alloc_off[c]; hl *= 1099511628211ull; }
            for (int c = 0; c < C; c++) { hs ^= (unsigned long long)(selected_host[c] ? 1 : 0) + 2; hs *= 1099511628211ull; }
            // the finish kernels take a pointer into the median table that depends on the number of windows of THIS plan
            hs ^= (unsigned long long)t.base_chrom + 0x9e3779b97f4a7c15ull; hs *= 1099511628211ull;
            hs ^= (use_int ? 0x51ull : 0x15ull) + 0x100ull * (unsigned)part; hs *= 1099511628211ull;  // the integer-key form has one more event wait per chromosome
            const long long key[12] = {(long long)(uintptr_t)ctx->arena, (long long)(uintptr_t)d.cov, (long long)(uintptr_t)d.ctl,
                                       (long long)(uintptr_t)d.cc, (long long)(uintptr_t)d.bp, C, o->min_size, o->is_germline, (long long)hl,
                                       (long long)hs, (long long)(uintptr_t)d.rq_sorted, (long long)(uintptr_t)d.tiny_tab};
            for (int i = 0; i < 12; i++) want.key[i] = key[i];
            CgGraphEntry* hit = nullptr;
            for (auto& g : ctx->part_graphs)
                if (!memcmp(g.key, want.key, sizeof(want.key))) { hit = &g; break; }
            if (!hit) {  // first sighting of this shape: remember it, launch directly below
                if (ctx->part_graphs.size() >= 16) {
                    for (auto& g : ctx->part_graphs)
                        if (g.exec) cudaGraphExecDestroy(g.exec);
                    ctx->part_graphs.clear();
                }
                want.exec = nullptr;
                ctx->part_graphs.push_back(want);
            } else {
                if (!hit->exec) {
                    const int launches_before = ctx->launches;
                    cudaGraph_t graph = nullptr;
                    if (cudaStreamBeginCapture(pipe, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                        const int rc_cap = enqueue_pipelines(d.alloc_off, pipe, true, part);
                        ctx->stream = s;
                        const cudaError_t ce = cudaStreamEndCapture(pipe, &graph);
                        hit->launches = ctx->launches - launches_before;
                        ctx->launches = launches_before;
                        if (rc_cap == CG_OK && ce == cudaSuccess && graph) {
                            if (cudaGraphInstantiate(&hit->exec, graph, 0) != cudaSuccess) hit->exec = nullptr;
                        }
                        if (graph) cudaGraphDestroy(graph);
                        cudaGetLastError();
                        ctx->launch_err = cudaSuccess;
                    } else {
                        cudaGetLastError();
                    }
                }
                if (hit->exec) {
                    const cudaError_t le = cudaGraphLaunch(hit->exec, pipe);
                    if (le != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, std::string("partition: graph launch failed: ") + cudaGetErrorString(le));
                    ctx->launches += hit->launches;
                    replayed = true;
                }
            }
        }
        if (!replayed) {
            const int rc_p = enqueue_pipelines(pl.off.data(), pipe, false, part);
            ctx->stream = s;
            if (rc_p) return rc_p;
        }
        return CG_OK;
    };
    // Experiment, off by default (CANVAS_SPLIT_PIPE=1): with the prefix sums already there (fused call) the first half of every
    // pipeline — seed, chains, mid stage — is launched BEFORE the host enqueues the ~60 kernels of the order statistics (0.2 ms
    // of host time) and the second half once the thresholds' event is recorded, stream-ordered behind the first.  Measured
    // 3.44 ms against 3.10 (profiles/rd2ad_split_pipelines_ab.txt): the barrier between the halves costs more than the earlier
    // start gains — a short chromosome's small stage must not wait for the longest chromosome's chain.
    static const bool split_env = getenv("CANVAS_SPLIT_PIPE") && atoi(getenv("CANVAS_SPLIT_PIPE")) != 0;
    const bool split_pipe = split_env && scan_done && !dbg_on;
    if (split_pipe) {
        CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_scan, 0));
        CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_plan, 0));
        cudaEventRecord(ctx->stage_ev[4], pipe);
        const int rc_l = launch_part(1);
        if (rc_l) return rc_l;
    }
    // ---- side stream: the evenness score and the factor-of-three list (per-window evenness, triplet cascade, their order
    // statistics on double keys).  The finish stage of every chromosome reads the factor-of-three list (germline refinement),
    // so the pipelines wait for ev_join before their finish kernels; the results are packed after it as well.
    // (Enqueueing this block AFTER the pipelines' launch was measured — the pipelines start 0.1 ms earlier — and lost: 3.20 ms
    // against 3.03, the side kernels then compete with the chains instead of running in the host's shadow; profiles/rd2z_*.)
    const bool side_first = use_int;
    auto enqueue_side = [&]() -> int {
        cudaStream_t main_s = ctx->stream;
        CG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        ctx->stream = ctx->side_stream;  // CG_LAUNCH and the select drivers enqueue on ctx->stream
        enqueue_evenness();
        enqueue_triplets();
        CG_LAUNCH(ctx, wv_request_kernel, rq_grid, 128, 0, d.sel, sp, d.ctl, 1, 1);
        sel_run_contig<uint64_t, PartView>(ctx, pv, d.work + pl.n_work_cov, d.seg_nwork, nwork - (int)pl.n_work_cov, d.sel);
        CG_LAUNCH(ctx, wv_median_finish_kernel, rq_grid, 128, 0, d.sel, d.med);
        CG_LAUNCH(ctx, wv_evenness_finish_kernel, 1, 1, 0, d.sel, t, d.ctl);
        CG_LAUNCH(ctx, wv_f3_finish_kernel, 1, 1, 0, sp, d.med, d.ctl);
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
        ctx->stream = main_s;
        return CG_OK;
    };
    if (use_int) {
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));  // plan tables are up: the side stream may start from here
        if (side_first) { const int rc_s = enqueue_side(); if (rc_s) return rc_s; }
        // ---- main stream: prefix sums, then medians and MADs of the coverage windows / chromosomes on integer hundredths
        enqueue_scan();
        if (!scan_done) CG_CUDA(ctx, cudaEventRecord(ctx->ev_scan, ctx->stream));  // the decomposition's chains and mid stage need nothing else
        CG_TL(ctx, "scan");
        CovView32 cv1{d.hq, nullptr}, cv2{d.hq, d.m2};
        CG_LAUNCH(ctx, wv_request32_kernel, rq_grid, 128, 0, d.sel32, sp, 1);
        sel_run_contig<uint32_t, CovView32>(ctx, cv1, d.work, d.seg_nwork, (int)pl.n_work_cov, d.sel32, 24, 0);
        CG_LAUNCH(ctx, wv_median_finish32_kernel, rq_grid, 128, 0, d.sel32, d.med, d.m2);
        CG_TL(ctx, "wave1 medians");
        CG_LAUNCH(ctx, wv_request32_kernel, rq_grid, 128, 0, d.sel32, sp, 2);
        sel_run_contig<uint32_t, CovView32>(ctx, cv2, d.work, d.seg_nwork, (int)pl.n_work_cov, d.sel32, 24, 1);
        CG_LAUNCH(ctx, wv_mad_finish32_kernel, rq_grid, 128, 0, d.sel32, d.med, d.m2, d.mad);
        CG_TL(ctx, "wave2 mads");
    } else {
        // ---- evenness per window (coverage only: early, so that the device has work while the host launches the rest)
        enqueue_evenness();
        CG_TL(ctx, "evenness");
        enqueue_scan();
        if (!scan_done) CG_CUDA(ctx, cudaEventRecord(ctx->ev_scan, ctx->stream));
        CG_TL(ctx, "scan");
        enqueue_triplets();
        CG_TL(ctx, "triplets");
        // ---- order statistics on double keys, dependent waves
        CG_LAUNCH(ctx, wv_request_kernel, rq_grid, 128, 0, d.sel, sp, d.ctl, 1, 0);
        sel_run_contig<uint64_t, PartView>(ctx, pv, d.work, d.seg_nwork, nwork, d.sel);
        CG_TL(ctx, "wave1 medians");
        CG_LAUNCH(ctx, wv_median_finish_kernel, rq_grid, 128, 0, d.sel, d.med);
        CG_LAUNCH(ctx, wv_evenness_finish_kernel, 1, 1, 0, d.sel, t, d.ctl);
        CG_LAUNCH(ctx, wv_f3_finish_kernel, 1, 1, 0, sp, d.med, d.ctl);
        PartView pv2 = pv;
        pv2.center = d.med;
        CG_LAUNCH(ctx, wv_request_kernel, rq_grid, 128, 0, d.sel, sp, d.ctl, 2, 0);
        sel_run_contig<uint64_t, PartView>(ctx, pv2, d.work, d.seg_nwork, nwork, d.sel);
        CG_TL(ctx, "wave2 mads");
        CG_LAUNCH(ctx, wv_median_finish_kernel, rq_grid, 128, 0, d.sel, d.mad);
    }
    if (pl.cv_possible) {
        if (t.base_chrom > 0) CG_LAUNCH(ctx, wv_ratio_kernel, div_up(t.base_chrom, 128), 128, 0, t, d.med, d.mad, d.r10, d.r100);
        if (t.n_w10 <= WV_RATIO_SORT_MAX && t.n_w100 <= WV_RATIO_SORT_MAX) {
            CG_LAUNCH(ctx, wv_ratio_stats_kernel, 2, 1024, 0, d.ratio_keys, sp, d.r10, d.r100);
        } else {
            // more windows than one CTA sorts: the select engine on double keys (after the side stream is done with its state)
            if (use_int) CG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
            CG_LAUNCH(ctx, wv_request_kernel, rq_grid, 128, 0, d.sel, sp, d.ctl, 3, 0);
            sel_run_contig<uint64_t, PartView>(ctx, pv, d.work, d.seg_nwork, nwork, d.sel);
            CG_LAUNCH(ctx, wv_ratio_keys_from_select_kernel, 1, 32, 0, d.sel, t, d.ratio_keys);
        }
    }
    CG_LAUNCH(ctx, wv_cv_sigma_kernel, 1, 128, 0, d.ratio_keys, sp, d.med, d.mad, d.off, d.ctl, d.sigma, d.cand_thr);

    CG_TL(ctx, "wave3 + cv");
    ctx->host_ts[1] = wv_now_us();
    cudaEventRecord(ctx->stage_ev[3], s);
    CG_CUDA(ctx, cudaEventRecord(ctx->ev_thr, s));  // thresholds (sigma, cand_thr) exist: the small / tiny stages and the finish may run
    ctx->stage_used[2] = true;
    // The chromosome pipelines run on their own root stream from the moment the prefix sums exist: chains and the mid stage
    // record EVERY node as a candidate (a few hundred per chromosome; the finish stage applies the real threshold anyway),
    // so they need no threshold and overlap the order statistics above; a chromosome's small stage waits for ev_thr.
    if (!split_pipe) {
        CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_scan, 0));
        CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_plan, 0));  // (the prefix sums of the fused call are older than the plan upload)
        if (rq_index_done) CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_rq, 0));  // built on the side stream (fused call)
        cudaEventRecord(ctx->stage_ev[4], pipe);
        const int rc_l = launch_part(0);
        if (rc_l) return rc_l;
    } else {
        // second half: thresholds and the factor-of-three list exist (events recorded above), the index is built
        CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_thr, 0));
        if (use_int) CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_join, 0));
        if (rq_index_done) CG_CUDA(ctx, cudaStreamWaitEvent(pipe, ctx->ev_rq, 0));
        const int rc_l = launch_part(2);
        if (rc_l) return rc_l;
    }
    ctx->host_ts[2] = wv_now_us();
    if (use_int && !side_first) { const int rc_s = enqueue_side(); if (rc_s) return rc_s; }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev_pipe, pipe));
    CG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_pipe, 0));
    CG_TL(ctx, "decompose + finish");
    cudaEventRecord(ctx->stage_ev[5], s);
    cudaEventRecord(ctx->stage_ev[6], s);
    ctx->stage_used[3] = true;
    cudaEventRecord(ctx->stage_ev[7], s);
    if (use_int) CG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // evenness / factor-of-three scalars are in the control block
    CG_LAUNCH(ctx, wv_pack_kernel, 1, 256, 0, d.n_bp, d.depth, d.bp, d.off, C, d.pack, d.cc, d.ctl);
    if (comm_pack) CG_LAUNCH(ctx, wv_pack_comm_kernel, 1, 256, 0, d.n_bp, d.bp, d.off, C, comm_pack, ctx->comm->pack_ints);
    return CG_OK;
}

int wv_collect(cg_ctx* ctx, const WvPlan& pl, WvDev& d, int32_t* n_bp, int32_t* bp, double* evenness, int* evenness_ok,
               double* cv, int* cv_has_value, double* factor_of_three, bool exchange = false) {
    cudaStream_t s = ctx->stream;
    const int C = pl.n_chrom;
    WvCtl* h = (WvCtl*)ctx->pinned;
    int* h_pack = (int*)(ctx->pinned + 8192);
    static_assert(sizeof(WvCtl) <= 8192, "control block does not fit its pinned slot");
    // one download: control block + packed results (counts, depths and the first breakpoints; the rest, if any, follows)
    const int head = 2 * C + 1;
    const int first = std::min<int>(WV_PACK_INTS, head + 4096);
    CG_CUDA(ctx, cudaMemcpyAsync(h, d.ctl, sizeof(WvCtl), cudaMemcpyDeviceToHost, s));
    if (pl.N > 0) CG_CUDA(ctx, cudaMemcpyAsync(h_pack, d.pack, (size_t)first * 4, cudaMemcpyDeviceToHost, s));
    else memset(h_pack, 0, (size_t)first * 4);  // nothing was enqueued for an empty genome
    // sharded run: the all-gather of every rank's packed lists rides on the same stream (its first wait covers the copies above)
    std::vector<int64_t> x_counts;
    std::vector<int32_t> x_all, x_full;
    if (exchange) {
        auto fetch = [&]() -> const int32_t* {  // this rank's lists did not fit the fixed-capacity round: all of them, from the device
            x_full.assign((size_t)C, 0);
            if (C > 0) cudaMemcpy(x_full.data(), d.n_bp, (size_t)C * 4, cudaMemcpyDeviceToHost);
            for (int c = 0; c < C; c++) {
                const size_t at = x_full.size();
                const int n = x_full[c];
                x_full.resize(at + (size_t)n);
                if (n > 0) cudaMemcpy(x_full.data() + at, d.bp + pl.off[c], (size_t)n * 4, cudaMemcpyDeviceToHost);
            }
            return x_full.data();
        };
        int rc = comm_allgatherv(ctx, nullptr, 0, fetch, x_counts, x_all);
        if (rc) return rc;
    } else {
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    if (h->overflow_.v) return cg_fail(ctx, CG_ERR_CAPACITY, "partition: internal queue capacity exceeded");
    ctx->stats[0] = (double)(h->visits_big + h->visits_small + h->visits_tiny);
    ctx->stats[1] = (double)(h->nodes_big + h->nodes_small + h->nodes_tiny);
    ctx->stats[2] = (double)h->cand_total;
    ctx->stats[3] = (double)pl.N;
    const int* h_nbp = h_pack;
    for (int c = 0; c < C; c++) n_bp[c] = h_nbp[c];
    if (getenv("CANVAS_DEBUG")) {
        std::vector<unsigned long long> ph((size_t)(pl.n_chrom + 1) * 8, 0), tn((size_t)(pl.n_chrom + 1) * 16, 0);
        cudaMemcpy(ph.data(), d.phase_ns, ph.size() * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(tn.data(), d.tl_ns, tn.size() * 8, cudaMemcpyDeviceToHost);
        if (d.task_dbg) {
            // the slowest mid-stage subtrees: a stage ends with its slowest CTA
            std::vector<unsigned long long> td(2 * UH_TASK_DBG_CAP + 2, 0);
            cudaMemcpy(td.data(), d.task_dbg, td.size() * 8, cudaMemcpyDeviceToHost);
            const size_t cnt = (size_t)std::min<unsigned long long>(td[0], UH_TASK_DBG_CAP);
            std::vector<size_t> idx(cnt);
            for (size_t i = 0; i < cnt; i++) idx[i] = i;
            std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return (td[2 + 2 * a] & 0xffffffffull) > (td[2 + 2 * b] & 0xffffffffull); });
            double tot = 0;
            for (size_t i = 0; i < cnt; i++) tot += (double)(td[2 + 2 * i] & 0xffffffffull);
            fprintf(stderr, "[mid] %zu subtrees, %.1f us of CTA time in total; slowest:\n", cnt, tot * 1e-3);
            for (size_t k = 0; k < std::min<size_t>(cnt, 12); k++) {
                const size_t i = idx[k];
                fprintf(stderr, "[mid]   chr %2llu bins %6llu nodes %5llu %8.1f us\n", td[1 + 2 * i] >> 32, td[1 + 2 * i] & 0xffffffffull,
                        td[2 + 2 * i] >> 32, (double)(td[2 + 2 * i] & 0xffffffffull) * 1e-3);
            }
        }
        {
            // pipeline of every chromosome: start (relative to the first decomposition kernel) and duration of each stage, us
            const unsigned long long t0 = h->t_first;
            for (int c = 0; c < pl.n_chrom; c++) {
                const unsigned long long* q = tn.data() + (size_t)c * 16;
                const unsigned long long* f = ph.data() + (size_t)c * 8;
                if (!q[1] && !q[3] && !q[5] && !q[7]) continue;
                auto rel = [&](unsigned long long t) { return t > t0 ? (double)(t - t0) * 1e-3 : 0.0; };
                fprintf(stderr, "[pipe] chr %2d n=%7lld | chain %7.1f +%7.1f | mid %7.1f +%7.1f | small %7.1f +%7.1f | tiny %7.1f +%7.1f | finish %7.1f +%7.1f us\n", c,
                        (long long)(pl.off[c + 1] - pl.off[c]),
                        q[1] ? rel(~q[0]) : 0.0, q[1] ? (double)(q[1] - ~q[0]) * 1e-3 : 0.0, q[3] ? rel(~q[2]) : 0.0, q[3] ? (double)(q[3] - ~q[2]) * 1e-3 : 0.0,
                        q[5] ? rel(~q[4]) : 0.0, q[5] ? (double)(q[5] - ~q[4]) * 1e-3 : 0.0, q[7] ? rel(~q[6]) : 0.0, q[7] ? (double)(q[7] - ~q[6]) * 1e-3 : 0.0,
                        f[0] ? rel(f[0]) : 0.0, f[0] && f[5] > f[0] ? (double)(f[5] - f[0]) * 1e-3 : 0.0);
            }
        }
        for (int c = 0; c < pl.n_chrom; c++) {
            const unsigned long long* q = ph.data() + (size_t)c * 8;
            if (!q[0]) continue;
            fprintf(stderr, "[fin] chr %2d n=%7lld bp=%3d | sort %7.1f thr %7.1f sortsv %7.1f pieces+rec %7.1f heal %7.1f refine %7.1f us | idx ok %llu fallback %llu na_sum %llu na_max %llu\n", c,
                    (long long)(pl.off[c + 1] - pl.off[c]), n_bp[c], (q[1] - q[0]) * 1e-3, (q[2] - q[1]) * 1e-3, 0.0, (q[3] - q[2]) * 1e-3,
                    (q[4] - q[3]) * 1e-3, (q[5] - q[4]) * 1e-3, q[6] & 0xffffffffull, q[6] >> 32, q[7] & 0xffffffffull, q[7] >> 32);
        }
    }
    ctx->stats[4] = (double)h->visits_big; ctx->stats[5] = (double)h->visits_small; ctx->stats[6] = (double)h->visits_tiny;
    ctx->stats[7] = (double)h->nodes_big; ctx->stats[8] = (double)h->nodes_small; ctx->stats[9] = (double)h->nodes_tiny;
    ctx->stats[10] = h->t_big_done > h->t_first ? (double)(h->t_big_done - h->t_first) * 1e-6 : 0.0;  // ms
    ctx->stats[11] = h->t_last > h->t_first ? (double)(h->t_last - h->t_first) * 1e-6 : 0.0;
    ctx->stats[12] = (double)h->multi_chunk_nodes; ctx->stats[13] = (double)h->queue_hops;
    {
        int maxd = 0;
        for (int c = 0; c < C; c++) maxd = std::max(maxd, h_pack[C + c]);
        ctx->stats[14] = (double)maxd;
    }
    // breakpoints: from the packed block when they all arrived with it, else one copy per chromosome
    const long long total = h_pack[2 * C];
    if (exchange) {
        // every rank's block is [n_bp[0..C), lists in chromosome order]; a chromosome has breakpoints on its owner only
        for (int c = 0; c < C; c++) n_bp[c] = 0;
        size_t at = 0;
        for (size_t r = 0; r < x_counts.size(); r++) {
            const int32_t* blk = x_all.data() + at;
            if (x_counts[r] < C) return cg_fail(ctx, CG_ERR_CUDA, "partition: malformed block in the all-gather");
            const int32_t* src = blk + C;
            for (int c = 0; c < C; c++) {
                const int n = blk[c];
                if (n < 0 || (src - blk) + n > x_counts[r] || n > pl.off[c + 1] - pl.off[c])
                    return cg_fail(ctx, CG_ERR_CUDA, "partition: malformed block in the all-gather");
                if (n > 0) { n_bp[c] = n; memcpy(bp + pl.off[c], src, (size_t)n * 4); }
                src += n;
            }
            at += (size_t)x_counts[r];
        }
    } else if (head + total <= first) {
        const int* src = h_pack + head;
        for (int c = 0; c < C; c++) {
            if (h_nbp[c] > 0) memcpy(bp + pl.off[c], src, (size_t)h_nbp[c] * 4);
            src += h_nbp[c];
        }
    } else {
        for (int c = 0; c < C; c++)
            if (h_nbp[c] > 0)
                CG_CUDA(ctx, cudaMemcpyAsync(bp + pl.off[c], d.bp + pl.off[c], (size_t)h_nbp[c] * 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    *evenness = h->evenness;
    *evenness_ok = h->evenness_ok;
    *cv = h->cv;
    *cv_has_value = h->cv_has_value;
    for (int i = 0; i <= WV_F3_LEVELS; i++) factor_of_three[i] = h->f3[i];
    return CG_OK;
}

int check_args(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom, const int64_t* chrom_off) {
    if (n_chrom > WV_MAX_CHROM) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_wavelet: more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (see DESIGN.md, Limits)");
    if (!opts || n_chrom < 0 || !chrom_off) return cg_fail(ctx, CG_ERR_ARG, "partition: bad argument");
    if (opts->evenness_window <= 0) return cg_fail(ctx, CG_ERR_ARG, "partition: evenness_window must be positive");
    if (chrom_off[0] != 0) return cg_fail(ctx, CG_ERR_ARG, "partition: chrom_off[0] must be 0");
    for (int c = 0; c < n_chrom; c++)
        if (chrom_off[c + 1] < chrom_off[c]) return cg_fail(ctx, CG_ERR_ARG, "partition: chrom_off must be non-decreasing");
    if (chrom_off[n_chrom] > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_ARG, "partition: too many bins");
    return CG_OK;
}

}  // namespace

// exchange: this rank segments the chromosomes the LPT assignment gives it and the all-gather completes the lists
static int partition_wavelet_impl(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom,
                                  const int64_t* chrom_off, const double* coverage,
                                  const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                                  double* evenness, int* evenness_ok, double* cv, int* cv_has_value,
                                  double* factor_of_three, bool exchange, int32_t* owner) {
    if (!ctx) return CG_ERR_ARG;
    int rc = check_args(ctx, opts, n_chrom, chrom_off);
    if (rc) return rc;
    std::vector<uint8_t> x_mask;
    if (exchange) {
        if (!ctx->comm) return cg_fail(ctx, CG_ERR_ARG, "cg_partition_wavelet_sharded: no communicator (cg_comm_init)");
        std::vector<int64_t> w(n_chrom);
        std::vector<int32_t> own(n_chrom);
        for (int c = 0; c < n_chrom; c++) w[c] = chrom_off[c + 1] - chrom_off[c];
        comm_assign_lpt(n_chrom, w.data(), ctx->comm->size, own.data());
        x_mask.assign(n_chrom + 1, 0);
        for (int c = 0; c < n_chrom; c++) { x_mask[c] = own[c] == ctx->comm->rank; if (owner) owner[c] = own[c]; }
        chrom_selected = x_mask.data();
    }
    if (!n_bp || !bp || !evenness || !evenness_ok || !cv || !cv_has_value || !factor_of_three || (chrom_off[n_chrom] > 0 && !coverage))
        return cg_fail(ctx, CG_ERR_ARG, "partition: null output");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    WvPlan pl;
    make_plan(pl, n_chrom, chrom_off, opts->evenness_window);
    rc = arena_reserve(ctx, wv_workspace_bytes(pl));
    if (rc) return rc;
    WvDev d;
    rc = wv_alloc(ctx, pl, d, nullptr);
    if (rc) return rc;
    std::vector<unsigned char> sel(n_chrom + 1, 1);
    if (chrom_selected)
        for (int c = 0; c < n_chrom; c++) sel[c] = chrom_selected[c] ? 1 : 0;
    cudaStream_t s = ctx->stream;
    if (pl.N > 0) CG_CUDA(ctx, cudaMemcpyAsync(d.cov, coverage, (size_t)pl.N * 8, cudaMemcpyHostToDevice, s));
    CgTimeline tl;
    tl.begin(s);
    ctx->tl = tl.on ? &tl : nullptr;
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    rc = wv_clear(ctx, d);
    if (rc) return rc;
    // integer keys for the order statistics when every coverage value is a plain two-decimal number (what the .cleaned
    // text holds); the host needs that verdict before it enqueues the select passes
    bool use_int = false;
    if (pl.N > 0 && !getenv("CANVAS_NO_INT_KEYS")) {
        CG_CUDA(ctx, cudaMemsetAsync(d.hq_bad, 0, 4, s));
        CG_LAUNCH(ctx, wv_hundredths_kernel, std::max(1, std::min(div_up(pl.N, 256), ctx->num_sms * 8)), 256, 0, d.cov, pl.N, d.hq, d.hq_bad);
        unsigned* h_bad = (unsigned*)(ctx->pinned + 8192);
        CG_CUDA(ctx, cudaMemcpyAsync(h_bad, d.hq_bad, 4, cudaMemcpyDeviceToHost, s));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        use_int = *h_bad == 0u;
    }
    ctx->stats[15] = use_int ? 1.0 : 0.0;
    rc = wv_enqueue(ctx, opts, pl, d, sel.data(), false, exchange ? ctx->comm->d_send : nullptr, use_int);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    rc = wv_collect(ctx, pl, d, n_bp, bp, evenness, evenness_ok, cv, cv_has_value, factor_of_three, exchange);
    tl.print("partition");
    ctx->tl = nullptr;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    return rc;
}

extern "C" int cg_partition_wavelet_shard(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom,
                                          const int64_t* chrom_off, const double* coverage,
                                          const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp,
                                          double* evenness, int* evenness_ok, double* cv, int* cv_has_value,
                                          double* factor_of_three) {
    return partition_wavelet_impl(ctx, opts, n_chrom, chrom_off, coverage, chrom_selected, n_bp, bp, evenness, evenness_ok, cv,
                                  cv_has_value, factor_of_three, false, nullptr);
}

extern "C" int cg_partition_wavelet_sharded(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom, const int64_t* chrom_off,
                                            const double* coverage, int32_t* n_bp, int32_t* bp, double* evenness,
                                            int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three,
                                            int32_t* owner) {
    return partition_wavelet_impl(ctx, opts, n_chrom, chrom_off, coverage, nullptr, n_bp, bp, evenness, evenness_ok, cv,
                                  cv_has_value, factor_of_three, true, owner);
}

extern "C" int cg_partition_wavelet(cg_ctx* ctx, const cg_wavelet_opts* opts, int n_chrom, const int64_t* chrom_off,
                                    const double* coverage, int32_t* n_bp, int32_t* bp, double* evenness,
                                    int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three) {
    return cg_partition_wavelet_shard(ctx, opts, n_chrom, chrom_off, coverage, nullptr, n_bp, bp, evenness, evenness_ok,
                                      cv, cv_has_value, factor_of_three);
}

// coverage for the partition stage = .cleaned round trip of the cleaned counts; per-chromosome bin counts
// ... and the integer hundredths themselves (keys of the partition's order statistics); chrom_cnt[256] is raised when a
// value cannot take the integer path (negative, not finite, or beyond WV_HQ_SAT)
__global__ void fused_coverage_kernel(const float* __restrict__ count_out, const int32_t* __restrict__ kept,
                                      const uint8_t* __restrict__ chrom_in, const CleanCtl* __restrict__ ctl,
                                      double* __restrict__ cov, uint32_t* __restrict__ hq, unsigned* __restrict__ chrom_cnt) {
    const int n = ctl->n_out;
    const int n_round = ((n + 31) / 32) * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool ok = i < n;
        int c = 0;
        bool bad = false;
        if (ok) {
            const float v = count_out[i];
            double hundredths;
            const bool finite = dotnet_f2_hundredths(v, hundredths);
            const double r = hundredths / 100.0;
            cov[i] = !finite ? (double)v : (v < 0 ? -r : r);
            bad = !finite || v < 0 || !(hundredths < (double)WV_HQ_SAT);
            hq[i] = bad ? WV_HQ_SAT : (uint32_t)hundredths;
            c = chrom_in[kept[i]];
        }
        if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&chrom_cnt[256], 1u);
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned m = __match_any_sync(act, c);
            if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&chrom_cnt[c], (unsigned)__popc(m));
        }
    }
}

static int clean_partition_wavelet_impl(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts,
                                        int64_t n, const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                                        const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                                        const int32_t* stop, const float* count, const uint8_t* gc,
                                        const uint8_t* chrom_selected, int64_t* n_out, int32_t* kept_index,
                                        float* count_out, double* local_sd, int* gc_norm_skipped,
                                        int64_t* chrom_off_out, int32_t* n_bp, int32_t* bp, double* evenness,
                                        int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three,
                                        bool exchange, int32_t* owner) {
    if (!ctx) return CG_ERR_ARG;
    if (exchange && !ctx->comm) return cg_fail(ctx, CG_ERR_ARG, "cg_clean_partition_wavelet_sharded: no communicator (cg_comm_init)");
    if (n_chrom > WV_MAX_CHROM) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean_partition_wavelet: more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (see DESIGN.md, Limits)");
    if (!copts || !wopts || n < 0 || n > 0x7fff0000LL || n_chrom < 0 || !n_out || !local_sd ||
        !gc_norm_skipped || !chrom_off_out || !n_bp || !evenness || !evenness_ok || !cv || !cv_has_value || !factor_of_three)
        return cg_fail(ctx, CG_ERR_ARG, "cg_clean_partition_wavelet: bad argument");
    if (wopts->evenness_window <= 0) return cg_fail(ctx, CG_ERR_ARG, "partition: evenness_window must be positive");
    ctx->launches = 0;
    ctx->tl = nullptr;
    ctx->launch_err = cudaSuccess;
    ctx->last_kernel_ms = 0;
    for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
    ctx->gap_used = false;
    *n_out = 0; *local_sd = -1.0; *gc_norm_skipped = 0; *evenness = 0; *evenness_ok = 0; *cv = 0; *cv_has_value = 0;
    for (int i = 0; i <= WV_F3_LEVELS; i++) factor_of_three[i] = 0;
    for (int c = 0; c <= n_chrom; c++) chrom_off_out[c] = 0;
    for (int c = 0; c < n_chrom; c++) n_bp[c] = 0;
    if (n == 0 && !exchange) { *gc_norm_skipped = copts->gc_norm ? 1 : 0; return CG_OK; }
    if (n == 0) {  // still a collective: every rank enters the all-gather
        *gc_norm_skipped = copts->gc_norm ? 1 : 0;
        std::vector<int32_t> none((size_t)n_chrom + 1, 0), all;
        std::vector<int64_t> counts;
        if (owner) for (int c = 0; c < n_chrom; c++) owner[c] = 0;
        return comm_allgatherv(ctx, none.data(), n_chrom, nullptr, counts, all);
    }
    if (!chrom || !start || !stop || !count || !gc || !kept_index || !count_out || !bp || !chrom_is_autosome)
        return cg_fail(ctx, CG_ERR_ARG, "cg_clean_partition_wavelet: null array");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    static const bool host_times = getenv("CANVAS_HOST_TIMES") != nullptr;  // where the HOST thread spends the call (stderr)
    double ht[8] = {0};
    auto now_us = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    ht[0] = now_us();
    // chromosome runs of the input (ids are non-decreasing; the device validates that): binary search
    std::vector<int64_t> in_off(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++) {
        int64_t lo = in_off[c], hi = n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (chrom[mid] <= (uint8_t)c) lo = mid + 1; else hi = mid; }
        in_off[c + 1] = lo;
    }
    // workspace for both stages; the partition plan of the input lengths bounds the plan after cleaning
    WvPlan worst;
    make_plan(worst, n_chrom, in_off.data(), wopts->evenness_window);
    const bool loess = copts->gc_norm && copts->gc_mode != 0;
    size_t need = clean_workspace_bytes(n, n_chrom, loess) + arena_need(n, 8) + arena_need(n + 1, 4) + arena_need(264, 4) + wv_workspace_bytes(worst);
    need += need / 16;
    int rc = arena_reserve(ctx, need);
    if (rc) return rc;
    CleanDev d;
    rc = clean_alloc(ctx, n, n_chrom, d, loess);
    if (rc) return rc;
    d.max_chrom_bins = -1;
    if (in_off[n_chrom] == n) {
        d.max_chrom_bins = 0;
        for (int c = 0; c < n_chrom; c++) d.max_chrom_bins = std::max<int64_t>(d.max_chrom_bins, in_off[c + 1] - in_off[c]);
    }
    double* cov = arena_take<double>(ctx, n);
    uint32_t* hq = arena_take<uint32_t>(ctx, n + 1);
    unsigned* chrom_cnt = arena_take<unsigned>(ctx, 264);  // [256]: a coverage value that cannot take the integer-key path
    if (!cov || !hq || !chrom_cnt) return cg_fail(ctx, CG_ERR_CUDA, "arena exhausted");
    cudaStream_t s = ctx->stream;
    if (CgStageSlot* sl = cg_stage_take(ctx, n, chrom, start, stop, count, gc)) {
        // staged by cg_prefetch_bins while the previous call ran: read the columns where they are
        CG_CUDA(ctx, cudaStreamWaitEvent(s, sl->ready, 0));
        d.chrom = sl->chrom; d.gc = sl->gc; d.start = sl->start; d.stop = sl->stop; d.count = sl->count;
    } else {
        CG_CUDA(ctx, cudaMemcpyAsync(d.chrom, chrom, n, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.gc, gc, n, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.start, start, n * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.stop, stop, n * 4, cudaMemcpyHostToDevice, s));
        CG_CUDA(ctx, cudaMemcpyAsync(d.count, count, n * 4, cudaMemcpyHostToDevice, s));
    }
    CG_CUDA(ctx, cudaMemsetAsync(d.is_auto, 0, 256, s));
    if (n_chrom > 0) CG_CUDA(ctx, cudaMemcpyAsync(d.is_auto, chrom_is_autosome, n_chrom, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(d.is_chry, 0, 256, s));
    if (n_chrom > 0 && chrom_is_chrY) CG_CUDA(ctx, cudaMemcpyAsync(d.is_chry, chrom_is_chrY, n_chrom, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaMemsetAsync(chrom_cnt, 0, 264 * 4, s));
    CgTimeline tl;
    tl.begin(s);
    ctx->tl = tl.on ? &tl : nullptr;
    CG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    rc = clean_enqueue(ctx, copts, d);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    CG_LAUNCH(ctx, fused_coverage_kernel, std::max(1, std::min(div_up(n, 256), ctx->num_sms * 8)), 256, 0, d.count_out,
              d.kept, d.chrom, d.ctl, cov, hq, chrom_cnt);
    CG_TL(ctx, "coverage f2");
    CleanCtl* h = (CleanCtl*)ctx->pinned;
    unsigned* h_cnt = (unsigned*)(ctx->pinned + 8192);
    CG_CUDA(ctx, cudaMemcpyAsync(h, d.ctl, sizeof(CleanCtl), cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaMemcpyAsync(h_cnt, chrom_cnt, 264 * 4, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaEventRecord(ctx->ev_mid, s));
    // partition workspace sized for the input lengths (an upper bound of the cleaned ones); its accumulators are
    // cleared on the device while the host waits for the per-chromosome survivor counts
    WvDev wd;
    rc = wv_alloc(ctx, worst, wd, cov, hq);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    rc = wv_clear(ctx, wd);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    // ... and the range-quantile index of the finish stage is built there too: offsets come from the device-side counts
    std::vector<unsigned char> sel(n_chrom + 1, 1);
    if (exchange) {  // LPT over the chromosome run lengths of the input: known before Clean, identical on every rank
        std::vector<int64_t> w(n_chrom);
        std::vector<int32_t> own(n_chrom);
        for (int c = 0; c < n_chrom; c++) w[c] = in_off[c + 1] - in_off[c];
        comm_assign_lpt(n_chrom, w.data(), ctx->comm->size, own.data());
        for (int c = 0; c < n_chrom; c++) { sel[c] = own[c] == ctx->comm->rank; if (owner) owner[c] = own[c]; }
    } else if (chrom_selected)
        for (int c = 0; c < n_chrom; c++) sel[c] = chrom_selected[c] ? 1 : 0;
    bool scan_done = false;
    if (n_chrom > 0) {
        // chromosome offsets of the cleaned list, on the device: the index and the prefix sums need nothing else from the plan
        CG_LAUNCH(ctx, wv_device_offsets_kernel, 1, 32, 0, chrom_cnt, n_chrom, wd.rq_off2, wd.rq_tfirst2);
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_off, s));
        // ... the index on the side stream: the finish stage needs it a millisecond from now
        CG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_off, 0));
        ctx->stream = ctx->side_stream;
        cudaMemcpyAsync(wd.rq_sel2, sel.data(), n_chrom, cudaMemcpyHostToDevice, ctx->side_stream);
        wv_enqueue_rq_index(ctx, wd, n_chrom, worst.rq_ntiles, wd.rq_off2, wd.rq_tfirst2, wd.rq_sel2);
        ctx->stream = s;
        CG_CUDA(ctx, cudaEventRecord(ctx->ev_rq, ctx->side_stream));
        CG_TL(ctx, "clears");
        // ... and the prefix sums on this stream, while the host waits for the counts and plans: the tile grid is laid out for
        // the input lengths and every tile clips itself against the device-side offsets (same tile boundaries relative to a
        // chromosome's first bin as the host's table, hence the same sums)
        const size_t nt = worst.tiles.size();
        const size_t early_bytes = (nt + (size_t)n_chrom + 2) * 4;
        constexpr size_t EARLY_AT = 128u << 10;  // second half of the pinned scalar block (the first holds the downloads)
        if (nt > 0 && EARLY_AT + early_bytes <= ctx->pinned_cap && !getenv("CANVAS_NO_EARLY_SCAN")) {
            int* h_tc = (int*)(ctx->pinned + EARLY_AT);
            int* h_tf = h_tc + nt;
            for (size_t k = 0; k < nt; k++) h_tc[k] = worst.tiles[k].c;
            for (int c = 0; c <= n_chrom; c++) h_tf[c] = worst.tile_first[c];
            CG_CUDA(ctx, cudaMemcpyAsync(wd.e_tile_c, h_tc, nt * 4, cudaMemcpyHostToDevice, s));
            CG_CUDA(ctx, cudaMemcpyAsync(wd.e_tile_first, h_tf, (size_t)(n_chrom + 1) * 4, cudaMemcpyHostToDevice, s));
            const WvScanSrc src{nullptr, wd.e_tile_c, wd.e_tile_first, wd.rq_off2};
            CG_LAUNCH(ctx, wv_scan_tile_sum_kernel, (int)nt, 256, 0, wd.cov, src, wd.tsum);
            CG_LAUNCH(ctx, wv_scan_tile_offsets_kernel, div_up(n_chrom, 64), 64, 0, wd.tsum, wd.e_tile_first, n_chrom);
            CG_LAUNCH(ctx, wv_scan_apply_kernel, (int)nt, 256, 0, wd.cov, src, wd.tsum, wd.rq_off2, wd.pz);
            CG_CUDA(ctx, cudaEventRecord(ctx->ev_scan, s));
            scan_done = true;
            CG_TL(ctx, "prefix sums (early)");
        }
    } else {
        CG_TL(ctx, "clears");
    }
    ht[1] = now_us();
    cudaEventRecord(ctx->gap_ev, s);
    ctx->gap_used = true;
    {
        // the counts have landed when ev_mid completes; the index kernels keep the device busy meanwhile.  Polling the
        // event returns as soon as it fires (a blocking wait on an event in the middle of a busy stream does not).
        cudaError_t q;
        while ((q = cudaEventQuery(ctx->ev_mid)) == cudaErrorNotReady) {}
        if (q != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, std::string("cg_clean_partition_wavelet: ") + cudaGetErrorString(q));
    }
    ht[2] = now_us();
    CG_CUDA(ctx, cudaGetLastError());
    CG_CHECK_LAUNCHES(ctx);
    if (h->unsorted) return cg_fail(ctx, CG_ERR_UNSORTED, "cg_clean: chromosome ids must form non-decreasing runs and GC must be 0..100");
    if (h->need_weighted)
        return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean: a GC bucket with < 100 autosomal bins needs more than 16384 neighbouring values for its weighted quantiles");
    const int64_t m = h->n_out;
    const double lsd = h->local_sd;
    const int skipped = h->gc_skipped;
    // clean outputs go back on the copy stream while the partition kernels run
    if (m > 0) {
        CG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_mid, 0));
        CG_CUDA(ctx, cudaMemcpyAsync(kept_index, d.kept, m * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CG_CUDA(ctx, cudaMemcpyAsync(count_out, d.count_out, m * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    std::vector<int64_t> off(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++) off[c + 1] = off[c] + h_cnt[c];
    for (int c = 0; c <= n_chrom; c++) chrom_off_out[c] = off[c];
    WvPlan pl;
    make_plan(pl, n_chrom, off.data(), wopts->evenness_window);
    const bool use_int = h_cnt[256] == 0u && !getenv("CANVAS_NO_INT_KEYS");
    ctx->stats[15] = use_int ? 1.0 : 0.0;
    ht[3] = now_us();
    rc = wv_enqueue(ctx, wopts, pl, wd, sel.data(), n_chrom > 0, exchange ? ctx->comm->d_send : nullptr, use_int, scan_done);
    if (rc) { cudaStreamSynchronize(s); cudaStreamSynchronize(ctx->copy_stream); return rc; }
    CG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    ht[4] = now_us();
    rc = wv_collect(ctx, pl, wd, n_bp, bp, evenness, evenness_ok, cv, cv_has_value, factor_of_three, exchange);
    ht[5] = now_us();
    if (host_times)
        fprintf(stderr, "[host] enqueue copies+clean+index %.0f us | wait for counts %.0f | plan %.0f | enqueue partition %.0f (plan tables up %.0f, "
                        "order statistics enqueued %.0f, pipelines launched %.0f) | wait+collect %.0f\n",
                ht[1] - ht[0], ht[2] - ht[1], ht[3] - ht[2], ht[4] - ht[3], ctx->host_ts[0] - ht[3], ctx->host_ts[1] - ht[3],
                ctx->host_ts[2] - ht[3], ht[5] - ht[4]);
    tl.print("fused");
    ctx->tl = nullptr;
    CG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_kernel_ms = ms;
    *n_out = m;
    *local_sd = lsd;
    *gc_norm_skipped = skipped;
    return rc;
}

extern "C" int cg_clean_partition_wavelet_shard(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts,
                                                int64_t n, const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                                                const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                                                const int32_t* stop, const float* count, const uint8_t* gc,
                                                const uint8_t* chrom_selected, int64_t* n_out, int32_t* kept_index,
                                                float* count_out, double* local_sd, int* gc_norm_skipped,
                                                int64_t* chrom_off_out, int32_t* n_bp, int32_t* bp, double* evenness,
                                                int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three) {
    return clean_partition_wavelet_impl(ctx, copts, wopts, n, chrom, chrom_is_autosome, chrom_is_chrY, n_chrom, start, stop, count, gc,
                                        chrom_selected, n_out, kept_index, count_out, local_sd, gc_norm_skipped, chrom_off_out, n_bp,
                                        bp, evenness, evenness_ok, cv, cv_has_value, factor_of_three, false, nullptr);
}

extern "C" int cg_clean_partition_wavelet_sharded(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts, int64_t n,
                                                  const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                                                  const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                                                  const int32_t* stop, const float* count, const uint8_t* gc, int64_t* n_out,
                                                  int32_t* kept_index, float* count_out, double* local_sd, int* gc_norm_skipped,
                                                  int64_t* chrom_off_out, int32_t* n_bp, int32_t* bp, double* evenness,
                                                  int* evenness_ok, double* cv, int* cv_has_value, double* factor_of_three,
                                                  int32_t* owner) {
    return clean_partition_wavelet_impl(ctx, copts, wopts, n, chrom, chrom_is_autosome, chrom_is_chrY, n_chrom, start, stop, count, gc,
                                        nullptr, n_out, kept_index, count_out, local_sd, gc_norm_skipped, chrom_off_out, n_bp, bp,
                                        evenness, evenness_ok, cv, cv_has_value, factor_of_three, true, owner);
}

extern "C" int cg_clean_partition_wavelet(cg_ctx* ctx, const cg_clean_opts* copts, const cg_wavelet_opts* wopts,
                                          int64_t n, const uint8_t* chrom, const uint8_t* chrom_is_autosome,
                                          const uint8_t* chrom_is_chrY, int n_chrom, const int32_t* start,
                                          const int32_t* stop, const float* count, const uint8_t* gc,
                                          int64_t* n_out, int32_t* kept_index, float* count_out,
                                          double* local_sd, int* gc_norm_skipped, int64_t* chrom_off_out,
                                          int32_t* n_bp, int32_t* bp, double* evenness, int* evenness_ok,
                                          double* cv, int* cv_has_value, double* factor_of_three) {
    return cg_clean_partition_wavelet_shard(ctx, copts, wopts, n, chrom, chrom_is_autosome, chrom_is_chrY, n_chrom, start, stop,
                                            count, gc, nullptr, n_out, kept_index, count_out, local_sd, gc_norm_skipped,
                                            chrom_off_out, n_bp, bp, evenness, evenness_ok, cv, cv_has_value, factor_of_three);
}
