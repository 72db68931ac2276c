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
        cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_thr, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_pipe, cudaEventDisableTiming) != cudaSuccess ||
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
    if (ctx->pipe_stream) { cudaStreamSynchronize(ctx->pipe_stream); cudaStreamDestroy(ctx->pipe_stream); }
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
