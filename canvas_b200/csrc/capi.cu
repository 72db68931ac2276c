// Context management of libcanvasgpu.
#include "common.cuh"

extern "C" int cg_create(int device, cg_ctx** out) {
    if (!out) return CG_ERR_ARG;
    *out = nullptr;
    // one hardware queue per chromosome stream of the partition (only effective if no CUDA context exists yet)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
        return CG_ERR_CUDA;  // no CPU fallback: the engine needs a CUDA device
    cg_ctx* ctx = new cg_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CG_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return CG_ERR_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    char buf[512];
    snprintf(buf, sizeof buf, "canvasgpu 0.1 sm_%d%d %s %d SMs", prop.major, prop.minor, prop.name, prop.multiProcessorCount);
    ctx->desc = buf;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_rq, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->pipe_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->plan_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_thr, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_pipe, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_off, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_plan, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_mid, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->pinned, 1 << 18) != cudaSuccess) {
        cg_destroy(ctx);
        return CG_ERR_CUDA;
    }
    ctx->pinned_cap = 1 << 18;
    for (int i = 0; i < 8; i++)
        if (cudaEventCreate(&ctx->stage_ev[i]) != cudaSuccess) { cg_destroy(ctx); return CG_ERR_CUDA; }
    if (cudaEventCreate(&ctx->gap_ev) != cudaSuccess) { cg_destroy(ctx); return CG_ERR_CUDA; }
    *out = ctx;
    return CG_OK;
}

extern "C" void cg_destroy(cg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cg_comm_destroy(ctx);
    if (ctx->arena) cudaFree(ctx->arena);
    cg_graphs_clear(ctx);
    if (ctx->gap_ev) cudaEventDestroy(ctx->gap_ev);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->plan_pinned) cudaFreeHost(ctx->plan_pinned);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->ped) cudaFree(ctx->ped);
    for (cudaEvent_t ev : ctx->ped_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->prefetch_stream) { cudaStreamSynchronize(ctx->prefetch_stream); cudaStreamDestroy(ctx->prefetch_stream); }
    for (CgStageSlot& sl : ctx->stage) {
        if (sl.base) cudaFree(sl.base);
        if (sl.ready) cudaEventDestroy(sl.ready);
    }
    for (int i = 0; i < 8; i++)
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
    for (cudaStream_t st : ctx->chrom_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (cudaEvent_t ev : ctx->chrom_ev) cudaEventDestroy(ev);
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_rq) cudaEventDestroy(ctx->ev_rq);
    if (ctx->ev_scan) cudaEventDestroy(ctx->ev_scan);
    if (ctx->ev_thr) cudaEventDestroy(ctx->ev_thr);
    if (ctx->ev_pipe) cudaEventDestroy(ctx->ev_pipe);
    if (ctx->ev_off) cudaEventDestroy(ctx->ev_off);
    if (ctx->ev_plan) cudaEventDestroy(ctx->ev_plan);
    if (ctx->pipe_stream) { cudaStreamSynchronize(ctx->pipe_stream); cudaStreamDestroy(ctx->pipe_stream); }
    if (ctx->plan_stream) { cudaStreamSynchronize(ctx->plan_stream); cudaStreamDestroy(ctx->plan_stream); }
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* cg_last_error(cg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" const char* cg_describe(cg_ctx* ctx) { return ctx ? ctx->desc.c_str() : ""; }
extern "C" double cg_last_kernel_ms(cg_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.0; }
extern "C" int cg_last_launches(cg_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" double cg_last_stage_ms(cg_ctx* ctx, int stage) {
    if (!ctx) return -1.0;
    float ms = 0;
    if (stage == 4 || stage == 5) {  // fused call: 4 = Clean end -> end of the pre-wait work, 5 = that -> start of the scalars
        if (!ctx->gap_used || !ctx->stage_used[0] || !ctx->stage_used[1]) return -1.0;
        const cudaError_t e = stage == 4 ? cudaEventElapsedTime(&ms, ctx->stage_ev[1], ctx->gap_ev)
                                         : cudaEventElapsedTime(&ms, ctx->gap_ev, ctx->stage_ev[2]);
        return e == cudaSuccess ? ms : -1.0;
    }
    if (stage < 0 || stage > 3 || !ctx->stage_used[stage]) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->stage_ev[2 * stage], ctx->stage_ev[2 * stage + 1]) != cudaSuccess) return -1.0;
    return ms;
}
extern "C" int cg_last_partition_stats(cg_ctx* ctx, double* out, int n) {
    if (!ctx || !out) return 0;
    int k = n < 16 ? n : 16;
    for (int i = 0; i < k; i++) out[i] = ctx->stats[i];
    return k;
}

extern "C" void* cg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void cg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------
// cg_prefetch_bins: the next sample's columns cross PCIe while the current call's kernels run
// ---------------------------------------------------------------------------------------------
extern "C" int cg_prefetch_bins(cg_ctx* ctx, int64_t n, const uint8_t* chrom, const int32_t* start, const int32_t* stop,
                                const float* count, const uint8_t* gc) {
    if (!ctx) return CG_ERR_ARG;
    if (n < 0 || n > 0x7fff0000LL) return cg_fail(ctx, CG_ERR_ARG, "cg_prefetch_bins: bad length");
    if (n == 0) return CG_OK;
    if (!chrom || !start || !stop || !count || !gc) return cg_fail(ctx, CG_ERR_ARG, "cg_prefetch_bins: null array");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->prefetch_stream) CG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->prefetch_stream, cudaStreamNonBlocking));
    // a free slot if there is one, else the older staged copy is given up: with one prefetch per call the slots alternate, and
    // the slot the next call reads is never the one being filled (calls are synchronous: nothing reads a slot right now)
    CgStageSlot* sl = !ctx->stage[0].staged ? &ctx->stage[0] : !ctx->stage[1].staged ? &ctx->stage[1]
                      : (ctx->stage[0].seq < ctx->stage[1].seq ? &ctx->stage[0] : &ctx->stage[1]);
    const size_t n_al = ((size_t)n + 255) & ~(size_t)255;
    const size_t need = n_al * 14;
    if (need > sl->cap) {
        if (sl->base) { CG_CUDA(ctx, cudaStreamSynchronize(ctx->prefetch_stream)); CG_CUDA(ctx, cudaFree(sl->base)); sl->base = nullptr; sl->cap = 0; }
        CG_CUDA(ctx, cudaMalloc((void**)&sl->base, need + (need >> 4)));
        sl->cap = need + (need >> 4);
    }
    if (!sl->ready) CG_CUDA(ctx, cudaEventCreateWithFlags(&sl->ready, cudaEventDisableTiming));
    sl->start = (int32_t*)sl->base;
    sl->stop = (int32_t*)(sl->base + n_al * 4);
    sl->count = (float*)(sl->base + n_al * 8);
    sl->chrom = (uint8_t*)(sl->base + n_al * 12);
    sl->gc = (uint8_t*)(sl->base + n_al * 13);
    cudaStream_t ps = ctx->prefetch_stream;
    sl->staged = false;
    CG_CUDA(ctx, cudaMemcpyAsync(sl->start, start, (size_t)n * 4, cudaMemcpyHostToDevice, ps));
    CG_CUDA(ctx, cudaMemcpyAsync(sl->stop, stop, (size_t)n * 4, cudaMemcpyHostToDevice, ps));
    CG_CUDA(ctx, cudaMemcpyAsync(sl->count, count, (size_t)n * 4, cudaMemcpyHostToDevice, ps));
    CG_CUDA(ctx, cudaMemcpyAsync(sl->chrom, chrom, (size_t)n, cudaMemcpyHostToDevice, ps));
    CG_CUDA(ctx, cudaMemcpyAsync(sl->gc, gc, (size_t)n, cudaMemcpyHostToDevice, ps));
    CG_CUDA(ctx, cudaEventRecord(sl->ready, ps));
    sl->n = n;
    sl->key[0] = chrom; sl->key[1] = start; sl->key[2] = stop; sl->key[3] = count; sl->key[4] = gc;
    sl->seq = ++ctx->stage_seq;
    sl->age = 0;
    sl->staged = true;
    return CG_OK;
}
