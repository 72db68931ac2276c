// The SmallPedigree chain between CanvasBin and the callers, device resident (include/canvasgpu.h: cg_pedigree_hmm):
//   CanvasClean per sample            Canvas/CanvasRunner.cs:883-893
//   bins common to every sample       CanvasRunner.cs:895-903, CanvasCommon/Utilities.cs:834-920 (MergeMultiSampleCleanedBedFile)
//   CanvasPartition -m PerSampleHMM   CanvasRunner.cs:927, CanvasPartition/HiddenMarkovModelsRunner.cs:23-109
// The reference runs these as 2 S + 1 processes that talk through gzip text files.  Here the cleaned lists of every sample
// stay in HBM between the stages: the shared bin layout is uploaded once (10 B/bin) and each sample adds its counts (4 B/bin);
// nothing but results crosses PCIe afterwards.  With a communicator of R ranks sample s is cleaned on rank s mod R, the cleaned
// lists travel GPU to GPU (grouped ncclBroadcast on device buffers), every rank merges (cheap), the S x C (sample, chromosome)
// units of the HMM are spread longest-first over the ranks and one all-gather of the packed breakpoint lists completes the
// result on every rank.
#include <chrono>
#include <cstring>
#include <vector>

#include "clean.cuh"
#include "comm.cuh"
#include "common.cuh"

int merge_kept_lists(cg_ctx* ctx, int64_t n_bins, int n_samples, const int64_t* n_kept, const int32_t* const* kept,
                     const float* const* count, int64_t* n_out, int32_t* common_index, float* count_out, bool on_device,
                     size_t out_stride);  // merge.cu
int hmm_partition_device_counts(cg_ctx* ctx, const cg_hmm_opts* o, int n_chrom, const int64_t* chrom_off, const float* d_count,
                                int text_mode, const uint8_t* chrom_selected, int32_t* n_bp, int32_t* bp);  // hmm.cu

namespace {

constexpr int PED_MAX_SAMPLES = 8;  // = MERGE_MAX_SAMPLES

// offsets of the chromosomes in the list of common bins: first common bin whose layout index is >= the chromosome's first bin
__global__ void ped_offsets_kernel(const int32_t* __restrict__ common, int n_common, const long long* __restrict__ layout_off,
                                   int C, long long* __restrict__ off_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > C) return;
    const long long key = layout_off[c];
    int lo = 0, hi = n_common;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (common[mid] < key) lo = mid + 1; else hi = mid;
    }
    off_out[c] = lo;
}

inline int ped_reserve(cg_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ped_cap) return CG_OK;
    if (ctx->ped) {
        CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CG_CUDA(ctx, cudaFree(ctx->ped));
        ctx->ped = nullptr;
        ctx->ped_cap = 0;
    }
    const size_t cap = bytes + (bytes >> 4) + (1u << 16);
    CG_CUDA(ctx, cudaMalloc((void**)&ctx->ped, cap));
    ctx->ped_cap = cap;
    return CG_OK;
}

struct PedSampleInfo { int64_t m; int skipped; double local_sd; };

double ms_since(std::chrono::steady_clock::time_point& t) {
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - t).count();
    t = now;
    return ms;
}

}  // namespace

extern "C" int cg_pedigree_hmm(cg_ctx* ctx, const cg_clean_opts* copts, const cg_hmm_opts* hopts, int n_samples, int64_t n,
                               const uint8_t* chrom, const uint8_t* chrom_is_autosome, const uint8_t* chrom_is_chrY, int n_chrom,
                               const int32_t* start, const int32_t* stop, const float* const* count, const uint8_t* gc, int sharded,
                               int64_t* n_kept, double* local_sd, int* gc_norm_skipped, int64_t* n_common, int32_t* common_index,
                               float* count_out, int64_t* chrom_off_out, int32_t* n_bp, int32_t* bp, int32_t* owner) {
    if (!ctx) return CG_ERR_ARG;
    if (n_chrom > 256) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_pedigree_hmm: more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (see DESIGN.md, Limits)");
    if (!copts || !hopts || n_samples < 1 || n < 0 || n > 0x7fff0000LL || n_chrom < 0 || !count || !n_kept || !local_sd || !gc_norm_skipped ||
        !n_common || !chrom_off_out || !n_bp)
        return cg_fail(ctx, CG_ERR_ARG, "cg_pedigree_hmm: bad argument");
    if (n_samples > PED_MAX_SAMPLES) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_pedigree_hmm: at most 8 samples");
    if (!hopts->per_sample) return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_pedigree_hmm: the chain segments every sample on its own (-m PerSampleHMM)");
    if (sharded && !ctx->comm) return cg_fail(ctx, CG_ERR_ARG, "cg_pedigree_hmm: no communicator (cg_comm_init)");
    const int S = n_samples, C = n_chrom;
    const int R = sharded ? ctx->comm->size : 1, me = sharded ? ctx->comm->rank : 0;
    const bool exchange = sharded && R > 1;
    for (int s = 0; s < S; s++) { n_kept[s] = 0; local_sd[s] = -1.0; gc_norm_skipped[s] = 0; }
    for (int i = 0; i < S * C; i++) n_bp[i] = 0;
    for (int c = 0; c <= C; c++) chrom_off_out[c] = 0;
    *n_common = 0;
    if (owner) for (int i = 0; i < S * C; i++) owner[i] = 0;
    double phase[9] = {0};  // the stages below reuse ctx->stats; the chain's own figures are written at the end
    bool any_mine = false;
    for (int s = 0; s < S; s++) any_mine = any_mine || (s % R == me);
    int rc_local = CG_OK;
    if (n > 0 && (!chrom || !start || !stop || !gc || !bp || (C > 0 && !chrom_is_autosome) || (!common_index != !count_out)))
        rc_local = cg_fail(ctx, CG_ERR_ARG, "cg_pedigree_hmm: null array");
    const bool want_tables = common_index && count_out;  // a rank that does not write the merged .cleaned files skips 4 (S + 1) B/bin of download
    for (int s = 0; s < S && rc_local == CG_OK && n > 0; s++)
        if (s % R == me && !count[s]) rc_local = cg_fail(ctx, CG_ERR_ARG, "cg_pedigree_hmm: null counts of a sample this rank cleans");
    if (rc_local != CG_OK && !exchange) return rc_local;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return cg_fail(ctx, CG_ERR_CUDA, "cudaSetDevice failed");
    auto t_phase = std::chrono::steady_clock::now();
    double kernel_ms = 0;
    int launches = 0;
    cudaStream_t st = ctx->stream;

    // ---- the block that outlives the stages
    const size_t n_al = ((size_t)n + 63) & ~(size_t)63;
    size_t ped_bytes = n_al * (4 * (size_t)S * 4 + 4 + 1 + 1 + 4 + 4) + (size_t)(C + 1) * 16 + 4096;
    int32_t* p_kept[PED_MAX_SAMPLES];
    float* p_cnt[PED_MAX_SAMPLES];
    float* p_raw[PED_MAX_SAMPLES];  // the samples' input counts, uploaded ahead of their Clean
    int32_t* p_common = nullptr;
    float* p_cnt_m = nullptr;
    int32_t *p_start = nullptr, *p_stop = nullptr;
    uint8_t *p_chrom = nullptr, *p_gc = nullptr;
    long long *p_layout_off = nullptr, *p_off = nullptr;
    std::vector<int64_t> layout_off((size_t)C + 1, 0);
    if (rc_local == CG_OK && n > 0) {
        rc_local = ped_reserve(ctx, ped_bytes);
        if (rc_local == CG_OK) {
            char* b = ctx->ped;
            for (int s = 0; s < S; s++) { p_kept[s] = (int32_t*)b; b += n_al * 4; }
            for (int s = 0; s < S; s++) { p_cnt[s] = (float*)b; b += n_al * 4; }
            p_cnt_m = (float*)b; b += n_al * 4 * S;
            for (int s = 0; s < S; s++) { p_raw[s] = (float*)b; b += n_al * 4; }
            p_common = (int32_t*)b; b += n_al * 4;
            p_start = (int32_t*)b; b += n_al * 4;
            p_stop = (int32_t*)b; b += n_al * 4;
            p_layout_off = (long long*)b; b += (size_t)(C + 1) * 8;
            p_off = (long long*)b; b += (size_t)(C + 1) * 8;
            p_chrom = (uint8_t*)b; b += n_al;
            p_gc = (uint8_t*)b; b += n_al;
        }
    }
    // chromosome runs of the layout (ids are non-decreasing; the device validates that during Clean)
    if (rc_local == CG_OK && n > 0) {
        for (int c = 0; c < C; c++) {
            int64_t lo = layout_off[c], hi = n;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (chrom[mid] <= (uint8_t)c) lo = mid + 1; else hi = mid; }
            layout_off[c + 1] = lo;
        }
    }

    // ---- CanvasClean: every sample once, on rank s mod R
    std::vector<PedSampleInfo> info((size_t)S, PedSampleInfo{0, 0, -1.0});
    auto clean_mine = [&]() -> int {
        if (n == 0) {  // an empty list is "every bin GC-filtered" for the reference (CanvasClean.cs:502-505)
            for (int s = 0; s < S; s++) info[s].skipped = copts->gc_norm ? 1 : 0;
            return CG_OK;
        }
        if (!any_mine) return CG_OK;
        CG_CUDA(ctx, cudaMemcpyAsync(p_chrom, chrom, (size_t)n, cudaMemcpyHostToDevice, st));
        CG_CUDA(ctx, cudaMemcpyAsync(p_gc, gc, (size_t)n, cudaMemcpyHostToDevice, st));
        CG_CUDA(ctx, cudaMemcpyAsync(p_start, start, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CG_CUDA(ctx, cudaMemcpyAsync(p_stop, stop, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        // every count column of this rank goes up front, on the copy stream: sample k's upload overlaps the Clean of sample k-1
        if (ctx->ped_ev.size() < (size_t)S) {
            const size_t have = ctx->ped_ev.size();
            ctx->ped_ev.resize((size_t)S, nullptr);
            for (size_t k = have; k < (size_t)S; k++) CG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ped_ev[k], cudaEventDisableTiming));
        }
        for (int s = 0; s < S; s++) {
            if (s % R != me) continue;
            CG_CUDA(ctx, cudaMemcpyAsync(p_raw[s], count[s], (size_t)n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
            CG_CUDA(ctx, cudaEventRecord(ctx->ped_ev[s], ctx->copy_stream));
        }
        const bool loess = copts->gc_norm && copts->gc_mode != 0;
        for (int s = 0; s < S; s++) {
            if (s % R != me) continue;
            ctx->launches = 0;
            ctx->tl = nullptr;
            ctx->launch_err = cudaSuccess;
            for (int i = 0; i < 4; i++) ctx->stage_used[i] = false;
            ctx->gap_used = false;
            int rc = arena_reserve(ctx, clean_workspace_bytes(n, C, loess));
            if (rc) return rc;
            CleanDev d;
            rc = clean_alloc(ctx, n, C, d, loess);
            if (rc) return rc;
            d.max_chrom_bins = -1;
            if (layout_off[C] == n) {
                d.max_chrom_bins = 0;
                for (int c = 0; c < C; c++) d.max_chrom_bins = std::max<int64_t>(d.max_chrom_bins, layout_off[c + 1] - layout_off[c]);
            }
            CG_CUDA(ctx, cudaMemcpyAsync(d.chrom, p_chrom, (size_t)n, cudaMemcpyDeviceToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d.gc, p_gc, (size_t)n, cudaMemcpyDeviceToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d.start, p_start, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            CG_CUDA(ctx, cudaMemcpyAsync(d.stop, p_stop, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            CG_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ped_ev[s], 0));
            CG_CUDA(ctx, cudaMemcpyAsync(d.count, p_raw[s], (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            CG_CUDA(ctx, cudaMemsetAsync(d.is_auto, 0, 256, st));
            if (C > 0) CG_CUDA(ctx, cudaMemcpyAsync(d.is_auto, chrom_is_autosome, C, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaMemsetAsync(d.is_chry, 0, 256, st));
            if (C > 0 && chrom_is_chrY) CG_CUDA(ctx, cudaMemcpyAsync(d.is_chry, chrom_is_chrY, C, cudaMemcpyHostToDevice, st));
            CG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
            rc = clean_enqueue(ctx, copts, d);
            if (rc) { cudaStreamSynchronize(st); return rc; }
            CG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
            CleanCtl* h = (CleanCtl*)ctx->pinned;
            CG_CUDA(ctx, cudaMemcpyAsync(h, d.ctl, sizeof(CleanCtl), cudaMemcpyDeviceToHost, st));
            CG_CUDA(ctx, cudaStreamSynchronize(st));
            CG_CUDA(ctx, cudaGetLastError());
            CG_CHECK_LAUNCHES(ctx);
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            kernel_ms += ms;
            launches += ctx->launches;
            if (h->unsorted) return cg_fail(ctx, CG_ERR_UNSORTED, "cg_clean: chromosome ids must form non-decreasing runs and GC must be 0..100");
            if (h->need_weighted)
                return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean: a GC bucket with < 100 autosomal bins needs more than 16384 neighbouring values for its weighted quantiles");
            info[s].m = h->n_out;
            info[s].skipped = h->gc_skipped;
            info[s].local_sd = h->local_sd;
            if (info[s].m > 0) {  // out of the arena before the next sample reuses it
                CG_CUDA(ctx, cudaMemcpyAsync(p_kept[s], d.kept, (size_t)info[s].m * 4, cudaMemcpyDeviceToDevice, st));
                CG_CUDA(ctx, cudaMemcpyAsync(p_cnt[s], d.count_out, (size_t)info[s].m * 4, cudaMemcpyDeviceToDevice, st));
            }
        }
        return CG_OK;
    };
    if (rc_local == CG_OK) rc_local = clean_mine();
    if (rc_local != CG_OK && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);  // uploads of the caller's columns may still be in flight
    phase[0] = ms_since(t_phase);

    // ---- every sample's cleaned list on every rank
    if (exchange) {
        // lengths and the per-sample scalars first; a rank that failed still enters (marker -1) so that nobody waits for it
        std::vector<int32_t> mine;
        mine.push_back(rc_local == CG_OK ? 0 : -1);
        if (rc_local == CG_OK)
            for (int s = 0; s < S; s++)
                if (s % R == me) {
                    int32_t w[2];
                    memcpy(w, &info[s].local_sd, 8);
                    mine.insert(mine.end(), {(int32_t)s, (int32_t)info[s].m, (int32_t)info[s].skipped, w[0], w[1]});
                }
        std::vector<int64_t> counts;
        std::vector<int32_t> all;
        const std::string first_err = ctx->err;
        int rc = comm_allgatherv(ctx, mine.data(), (int64_t)mine.size(), nullptr, counts, all);
        if (rc_local != CG_OK) { ctx->err = first_err; return rc_local; }
        if (rc) return rc;
        size_t at = 0;
        for (int r = 0; r < R; r++) {
            const int32_t* p = all.data() + at;
            if (counts[r] < 1 || p[0] != 0) return cg_fail(ctx, CG_ERR_CUDA, "cg_pedigree_hmm: rank " + std::to_string(r) + " failed in its Clean stage");
            for (int64_t k = 1; k + 5 <= counts[r]; k += 5) {
                const int s = p[k];
                if (s < 0 || s >= S) return cg_fail(ctx, CG_ERR_CUDA, "cg_pedigree_hmm: corrupt exchange");
                info[s].m = p[k + 1];
                info[s].skipped = p[k + 2];
                memcpy(&info[s].local_sd, p + k + 3, 8);
            }
            at += (size_t)counts[r];
        }
        if (n > 0) {
            const CgNccl* nc = cg_nccl(&ctx->err);
            if (!nc) return CG_ERR_CUDA;
            CG_CUDA(ctx, cudaEventRecord(ctx->comm->ev0, st));
            CG_NCCL(ctx, nc, nc->GroupStart());
            for (int s = 0; s < S; s++)
                if (info[s].m > 0) {
                    CG_NCCL(ctx, nc, nc->Broadcast(p_kept[s], p_kept[s], (size_t)info[s].m, ncclInt32, s % R, ctx->comm->comm, st));
                    CG_NCCL(ctx, nc, nc->Broadcast(p_cnt[s], p_cnt[s], (size_t)info[s].m, ncclFloat32, s % R, ctx->comm->comm, st));
                }
            CG_NCCL(ctx, nc, nc->GroupEnd());
            CG_CUDA(ctx, cudaEventRecord(ctx->comm->ev1, st));
            CG_CUDA(ctx, cudaStreamSynchronize(st));
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->comm->ev0, ctx->comm->ev1);
            ctx->comm->last_exchange_ms = ms;
            phase[7] = ms;
        }
    }
    if (rc_local != CG_OK) return rc_local;  // (with an exchange every rank has left above, together)
    for (int s = 0; s < S; s++) { n_kept[s] = info[s].m; local_sd[s] = info[s].local_sd; gc_norm_skipped[s] = info[s].skipped; }
    phase[1] = ms_since(t_phase);
    if (n == 0) return CG_OK;

    // ---- bins common to every sample (every rank)
    int64_t m_common = 0;
    {
        int64_t nk[PED_MAX_SAMPLES];
        for (int s = 0; s < S; s++) nk[s] = info[s].m;
        int rc = merge_kept_lists(ctx, n, S, nk, p_kept, p_cnt, &m_common, p_common, p_cnt_m, true, n_al);
        if (rc) return rc;
        kernel_ms += ctx->last_kernel_ms;
        launches += ctx->launches;
    }
    *n_common = m_common;
    std::vector<int64_t> off((size_t)C + 1, 0);
    if (m_common > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(p_layout_off, layout_off.data(), (size_t)(C + 1) * 8, cudaMemcpyHostToDevice, st));
        ped_offsets_kernel<<<div_up(C + 1, 64), 64, 0, st>>>(p_common, (int)m_common, p_layout_off, C, p_off);
        launches++;
        CG_CUDA(ctx, cudaMemcpyAsync(off.data(), p_off, (size_t)(C + 1) * 8, cudaMemcpyDeviceToHost, st));
        CG_CUDA(ctx, cudaStreamSynchronize(st));
        // The merged table (4 (S + 1) bytes per common bin, 48 MB for a trio) goes home on the copy stream while the HMM runs.
        // The HMM and the gather read their scalars and short lists back with kernel stores into page-locked memory
        // (cg_readback_small): small copies would queue behind this download on the copy engine of that direction
        // (measured on an 8-rank host: the HMM stage of a rank took 4.6 ms instead of 1.2).
        if (want_tables) {
            CG_CUDA(ctx, cudaMemcpyAsync(common_index, p_common, (size_t)m_common * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
            for (int s = 0; s < S; s++)
                CG_CUDA(ctx, cudaMemcpyAsync(count_out + (size_t)s * n, p_cnt_m + (size_t)s * n_al, (size_t)m_common * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    auto download_merged = [&]() -> int {  // wait for it
        CG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        return CG_OK;
    };
    for (int c = 0; c <= C; c++) chrom_off_out[c] = off[c];
    phase[2] = ms_since(t_phase);

    // ---- PerSampleHMM: the S x C (sample, chromosome) units, longest first over the ranks
    // Every HMM call carries a fixed cost (whole-genome quartiles, emission table, three host syncs: ~1 ms) next to ~0.5 ms of
    // per-chromosome work for a whole sample, so a rank should touch as few samples as possible: with R >= S the ranks r with
    // r mod S == s form sample s's group and share its chromosomes longest-first; with R < S whole samples go round robin.
    std::vector<int32_t> own((size_t)S * C, 0);
    if (C > 0) {
        std::vector<int64_t> w((size_t)C);
        for (int c = 0; c < C; c++) w[c] = off[c + 1] - off[c];
        for (int s = 0; s < S; s++) {
            if (R < S) {
                for (int c = 0; c < C; c++) own[(size_t)s * C + c] = s % R;
                continue;
            }
            const int group = (R - s + S - 1) / S;  // ranks s, s + S, s + 2 S, ... below R
            std::vector<int32_t> in_group((size_t)C);
            comm_assign_lpt(C, w.data(), group, in_group.data());
            for (int c = 0; c < C; c++) own[(size_t)s * C + c] = s + S * in_group[c];
        }
    }
    if (owner) for (int i = 0; i < S * C; i++) owner[i] = own[i];
    int rc_hmm = CG_OK;
    std::string hmm_err;
    for (int s = 0; s < S && rc_hmm == CG_OK && m_common > 0; s++) {
        std::vector<uint8_t> mask((size_t)C + 1, 0);
        bool any = false;
        for (int c = 0; c < C; c++) { mask[c] = own[(size_t)s * C + c] == me; any = any || mask[c]; }
        if (!any) continue;
        rc_hmm = hmm_partition_device_counts(ctx, hopts, C, off.data(), p_cnt_m + (size_t)s * n_al, 2, R > 1 ? mask.data() : nullptr,
                                             n_bp + (size_t)s * C, bp + (size_t)s * n);
        if (rc_hmm != CG_OK) { hmm_err = ctx->err; break; }
        kernel_ms += ctx->last_kernel_ms;
        launches += ctx->launches;
    }
    phase[3] = ms_since(t_phase);
    if (rc_hmm != CG_OK && !exchange) { cudaStreamSynchronize(ctx->copy_stream); return rc_hmm; }

    // ---- one all-gather of the packed lists: [sample, chromosome, count, breakpoints ...] per unit
    if (exchange) {
        std::vector<int32_t> mine;
        mine.push_back(rc_hmm == CG_OK ? 0 : -1);
        if (rc_hmm == CG_OK)
            for (int s = 0; s < S; s++)
                for (int c = 0; c < C; c++) {
                    const int k = n_bp[(size_t)s * C + c];
                    if (own[(size_t)s * C + c] != me || k <= 0) continue;
                    mine.insert(mine.end(), {(int32_t)s, (int32_t)c, (int32_t)k});
                    const int32_t* src = bp + (size_t)s * n + off[c];
                    mine.insert(mine.end(), src, src + k);
                }
        std::vector<int64_t> counts;
        std::vector<int32_t> all;
        int rc = comm_allgatherv(ctx, mine.data(), (int64_t)mine.size(), nullptr, counts, all);
        if (rc_hmm != CG_OK) { ctx->err = hmm_err; cudaStreamSynchronize(ctx->copy_stream); return rc_hmm; }
        if (rc) { cudaStreamSynchronize(ctx->copy_stream); return rc; }
        size_t at = 0;
        for (int r = 0; r < R; r++) {
            const int32_t* p = all.data() + at;
            if (counts[r] < 1 || p[0] != 0) { cudaStreamSynchronize(ctx->copy_stream); return cg_fail(ctx, CG_ERR_CUDA, "cg_pedigree_hmm: rank " + std::to_string(r) + " failed in its HMM stage"); }
            int64_t k = 1;
            while (k + 3 <= counts[r]) {
                const int s = p[k], c = p[k + 1], cnt = p[k + 2];
                if (s < 0 || s >= S || c < 0 || c >= C || cnt < 0 || k + 3 + cnt > counts[r] || cnt > off[c + 1] - off[c]) {
                    cudaStreamSynchronize(ctx->copy_stream);
                    return cg_fail(ctx, CG_ERR_CUDA, "cg_pedigree_hmm: corrupt exchange");
                }
                n_bp[(size_t)s * C + c] = cnt;
                memcpy(bp + (size_t)s * n + off[c], p + k + 3, (size_t)cnt * 4);
                k += 3 + cnt;
            }
            at += (size_t)counts[r];
        }
        phase[7] += ctx->comm->last_exchange_ms;
    }
    phase[4] = ms_since(t_phase);
    {
        const int rc_d = download_merged();
        if (rc_d) return rc_d;
    }
    phase[8] = ms_since(t_phase);
    phase[5] = kernel_ms;
    phase[6] = (double)launches;
    for (int i = 0; i < 16; i++) ctx->stats[i] = i < 9 ? phase[i] : 0.0;
    ctx->last_kernel_ms = kernel_ms;
    ctx->launches = launches;
    return CG_OK;
}
