// Shared plumbing of libcanvasgpu: context, device arena, error handling, key mappings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <string>
#include <vector>

#include "canvasgpu.h"

#define CG_NUM_SMS_FALLBACK 148

struct CgTimeline;
struct CgComm;  // NCCL communicator + exchange buffers of this context (comm.cuh); nullptr until cg_comm_init
// an instantiated CUDA graph of one launch sequence, valid for one exact problem shape and arena placement
struct CgGraphEntry {
    long long key[12];
    cudaGraphExec_t exec;
    int launches;
};
// Input columns of a sample staged ahead of the call that consumes them (cg_prefetch_bins)
struct CgStageSlot {
    char* base = nullptr;
    size_t cap = 0;
    int64_t n = 0;
    const void* key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // the host arrays the columns were copied from
    cudaEvent_t ready = nullptr;
    bool staged = false;
    unsigned long long seq = 0;  // order of staging: the oldest matching copy is consumed first
    int age = 0;                 // calls that passed it by: a copy nobody asks for is dropped after the second one
    uint8_t *chrom = nullptr, *gc = nullptr;
    int32_t *start = nullptr, *stop = nullptr;
    float* count = nullptr;
};

struct cg_ctx {
    int device = 0;
    int num_sms = CG_NUM_SMS_FALLBACK;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // overlaps result downloads with later kernels
    cudaStream_t side_stream = nullptr;  // kernels off the critical path (partition: evenness / factor-of-three statistics)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_rq = nullptr;
    cudaStream_t pipe_stream = nullptr;  // root of the per-chromosome pipelines of the partition
    cudaEvent_t ev_scan = nullptr, ev_thr = nullptr, ev_pipe = nullptr, ev_off = nullptr, ev_plan = nullptr;
    // partition: one stream per chromosome pipeline (decomposition stages + finish), created on first use
    std::vector<cudaStream_t> chrom_streams;
    std::vector<cudaEvent_t> chrom_ev;
    cudaEvent_t ev_fork2 = nullptr;
    bool uh_attrs_set = false;
    cudaEvent_t ev_mid = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    std::string desc;
    double last_kernel_ms = 0;
    int launches = 0;
    cudaError_t launch_err = cudaSuccess;  // first failed kernel launch of the current call
    const char* launch_err_kernel = "";
    // stage timers: events [2*i], [2*i+1] bracket stage i
    cudaEvent_t stage_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool stage_used[4] = {false, false, false, false};
    cudaEvent_t gap_ev = nullptr;  // fused call: end of the work enqueued before the mid-call wait (see cg_last_stage_ms 4, 5)
    bool gap_used = false;
    double stats[16] = {0};
    // device arena, grown on demand and reused across calls
    char* arena = nullptr;
    size_t arena_cap = 0;
    size_t arena_off = 0;
    // small pinned staging block for scalars coming back from the device
    char* pinned = nullptr;
    size_t pinned_cap = 0;
    // pinned staging block for the partition plan tables (one host-to-device copy per call)
    char* plan_pinned = nullptr;
    size_t plan_pinned_cap = 0;
    // second device block for tables whose size is only known in the middle of a call (HMM emission tables)
    char* aux = nullptr;
    size_t aux_cap = 0;
    // third device block: what the pedigree chain keeps between its stages (cleaned lists of every sample, merged counts)
    char* ped = nullptr;
    size_t ped_cap = 0;
    std::vector<cudaEvent_t> ped_ev;  // per sample: its input counts are on the device
    // two staging slots for prefetched inputs: one is read by the running call while the other fills
    CgStageSlot stage[2];
    unsigned long long stage_seq = 0;
    cudaStream_t prefetch_stream = nullptr;
    cudaStream_t plan_stream = nullptr;  // upload of the partition's plan tables: not queued behind kernels or the result downloads
    CgTimeline* tl = nullptr;  // debug timeline of the current call (CANVAS_DEBUG)
    std::vector<CgGraphEntry> clean_graphs;  // Clean pipeline graphs (clean.cu), dropped when the arena moves
    std::vector<CgGraphEntry> part_graphs;   // partition: per-chromosome pipelines (wavelet.cu); exec == nullptr: shape seen once
    CgComm* comm = nullptr;
    double host_ts[8] = {0};  // CANVAS_HOST_TIMES: host clock (us) at points inside the partition enqueue
};

constexpr size_t CG_CHROM_STREAMS = 32;
inline int cg_chrom_streams(cg_ctx* ctx, int n) {
    int prio_lo = 0, prio_hi = 0;  // numerically lower = more urgent
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    while ((int)ctx->chrom_streams.size() < n) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        // streams are handed out to the chromosomes in order of decreasing length: the longest chromosomes (whose finish
        // stage ends the partition) get their blocks scheduled first when pipelines compete for the SMs
        const int k = (int)ctx->chrom_streams.size();
        const int prio = std::min(prio_lo, prio_hi + k / 4);
        if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            ctx->err = "cannot create the per-chromosome streams";
            return CG_ERR_CUDA;
        }
        ctx->chrom_streams.push_back(st);
        ctx->chrom_ev.push_back(ev);
    }
    if (!ctx->ev_fork2 && cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming) != cudaSuccess) {
        ctx->err = "cannot create an event";
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

inline void cg_graphs_clear(cg_ctx* ctx) {
    for (auto& g : ctx->clean_graphs) cudaGraphExecDestroy(g.exec);
    ctx->clean_graphs.clear();
    for (auto& g : ctx->part_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->part_graphs.clear();
}

inline int cg_fail(cg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define CG_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return cg_fail(ctx, CG_ERR_CUDA,                                                 \
                           std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +      \
                               __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    } while (0)

// The staged copy of exactly these host arrays, if cg_prefetch_bins was called for them (consumed by the caller)
inline CgStageSlot* cg_stage_find(cg_ctx* ctx, int64_t n, const void* chrom, const void* start, const void* stop, const void* count,
                                  const void* gc) {
    CgStageSlot* hit = nullptr;
    for (CgStageSlot& sl : ctx->stage)
        if (sl.staged && sl.n == n && sl.key[0] == chrom && sl.key[1] == start && sl.key[2] == stop && sl.key[3] == count && sl.key[4] == gc &&
            (!hit || sl.seq < hit->seq))
            hit = &sl;
    return hit;
}

// The call that consumes staged columns: takes the oldest matching copy (nullptr: none, copy as usual).  Every other staged
// copy grows older; one that two calls in a row did not ask for is dropped, so that a forgotten prefetch cannot be matched much
// later by a host buffer that has been refilled in the meantime.  (In a pipeline — prefetch next, call current — a copy is
// passed by exactly once.)
inline CgStageSlot* cg_stage_take(cg_ctx* ctx, int64_t n, const void* chrom, const void* start, const void* stop, const void* count,
                                  const void* gc) {
    CgStageSlot* hit = cg_stage_find(ctx, n, chrom, start, stop, count, gc);
    for (CgStageSlot& sl : ctx->stage)
        if (sl.staged && &sl != hit && ++sl.age >= 2) sl.staged = false;
    if (hit) hit->staged = false;
    return hit;
}

// Arena: one cudaMalloc, bump allocation, 256-byte aligned.  reserve() is called once per API call
// with the worst-case footprint, take() hands out slices.
inline int arena_reserve(cg_ctx* ctx, size_t bytes) {
    ctx->arena_off = 0;
    if (bytes <= ctx->arena_cap) return CG_OK;
    if (ctx->arena) {
        CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cg_graphs_clear(ctx);  // they hold pointers into the old arena
        CG_CUDA(ctx, cudaFree(ctx->arena));
        ctx->arena = nullptr;
        ctx->arena_cap = 0;
    }
    size_t cap = bytes + (bytes >> 3) + (1u << 20);
    CG_CUDA(ctx, cudaMalloc((void**)&ctx->arena, cap));
    ctx->arena_cap = cap;
    return CG_OK;
}

// grow-only side block; contents do not survive the call that filled them
inline int aux_reserve(cg_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->aux_cap) return CG_OK;
    if (ctx->aux) {
        CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CG_CUDA(ctx, cudaFree(ctx->aux));
        ctx->aux = nullptr;
        ctx->aux_cap = 0;
    }
    const size_t cap = bytes + (bytes >> 2) + (1u << 16);
    CG_CUDA(ctx, cudaMalloc((void**)&ctx->aux, cap));
    ctx->aux_cap = cap;
    return CG_OK;
}

template <typename T>
inline T* arena_take(cg_ctx* ctx, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (ctx->arena_off + bytes > ctx->arena_cap) return nullptr;
    T* p = (T*)(ctx->arena + ctx->arena_off);
    ctx->arena_off += bytes;
    return p;
}

inline size_t arena_need(size_t count, size_t elem) { return ((count * elem + 255) & ~(size_t)255); }

// A few words from device memory into PAGE-LOCKED host memory, written by a kernel instead of the copy engine: a small
// device-to-host cudaMemcpyAsync queues behind whatever large download is in flight in that direction (the pedigree chain
// sends 48 MB home while its HMM stage reads scalars back: 4.6 ms instead of 1.2 on an 8-rank host), a store over PCIe from an
// SM does not.  dst must come from cudaMallocHost / cg_host_alloc; both pointers 4-byte aligned.
static __global__ void cg_small_copy_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
inline cudaError_t cg_readback_small(cudaStream_t s, void* pinned_dst, const void* dev_src, size_t bytes) {
    if (bytes == 0) return cudaSuccess;
    cg_small_copy_kernel<<<1, 256, 0, s>>>(static_cast<const unsigned*>(dev_src), static_cast<unsigned*>(pinned_dst), (int)((bytes + 3) / 4));
    return cudaGetLastError();
}
constexpr size_t CG_PINNED_SMALL_AT = 72u << 10;   // [72 KiB, 80 KiB) of ctx->pinned: scalars read back by cg_readback_small
constexpr size_t CG_PINNED_LIST_AT = 80u << 10;    // [80 KiB, 128 KiB): short result lists read back the same way
constexpr size_t CG_PINNED_LIST_BYTES = 48u << 10;

// ---------------------------------------------------------------------------------------------
// Order-preserving key maps.  .NET orders NaN below every number (Double.CompareTo), so NaN maps
// to key 0, which no other value produces.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t f32_key(float x) {
    if (x != x) return 0u;
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(x);
#else
    memcpy(&u, &x, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float f32_unkey(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    if (k == 0u) u = 0x7fc00000u;
    float x;
#ifdef __CUDA_ARCH__
    x = __uint_as_float(u);
#else
    memcpy(&x, &u, 4);
#endif
    return x;
}
__host__ __device__ inline uint64_t f64_key(double x) {
    if (x != x) return 0ull;
    uint64_t u;
#ifdef __CUDA_ARCH__
    u = (uint64_t)__double_as_longlong(x);
#else
    memcpy(&u, &x, 8);
#endif
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double f64_unkey(uint64_t k) {
    uint64_t u = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    if (k == 0ull) u = 0x7ff8000000000000ull;
    double x;
#ifdef __CUDA_ARCH__
    x = __longlong_as_double((long long)u);
#else
    memcpy(&x, &u, 8);
#endif
    return x;
}
__host__ __device__ inline uint32_t i32_key(int32_t x) { return (uint32_t)x ^ 0x80000000u; }
__host__ __device__ inline int32_t i32_unkey(uint32_t k) { return (int32_t)(k ^ 0x80000000u); }

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Debug timeline (CANVAS_DEBUG): named CUDA events on the ctx stream, printed after the call's last sync.
struct CgTimeline {
    bool on = false;
    cudaStream_t s = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> ev;
    void begin(cudaStream_t stream) { on = getenv("CANVAS_DEBUG") != nullptr; s = stream; mark("start"); }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.emplace_back(name, e);
    }
    void print(const char* tag) {
        if (!on) return;
        cudaStreamSynchronize(s);
        for (size_t i = 1; i < ev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            fprintf(stderr, "[%s] %-16s %8.1f us\n", tag, ev[i].first, ms * 1e3);
        }
        for (auto& e : ev) cudaEventDestroy(e.second);
        ev.clear();
    }
};

#define CG_TL(ctx, name) do { if ((ctx)->tl) (ctx)->tl->mark(name); } while (0)

// counted launch helper
#define CG_LAUNCH(ctx, kernel, grid, block, smem, ...)                                 \
    do {                                                                               \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);               \
        (ctx)->launches++;                                                             \
        cudaError_t le__ = cudaGetLastError();                                         \
        if (le__ != cudaSuccess && (ctx)->launch_err == cudaSuccess) {                 \
            (ctx)->launch_err = le__;                                                  \
            (ctx)->launch_err_kernel = #kernel;                                        \
        }                                                                              \
    } while (0)

// to be called after the stream has been synchronised: a launch that failed is an error of the call
#define CG_CHECK_LAUNCHES(ctx)                                                                         \
    do {                                                                                               \
        if ((ctx)->launch_err != cudaSuccess) {                                                        \
            cudaError_t e__ = (ctx)->launch_err;                                                       \
            (ctx)->launch_err = cudaSuccess;                                                           \
            return cg_fail(ctx, CG_ERR_CUDA, std::string("kernel launch failed: ") + (ctx)->launch_err_kernel + ": " + \
                                                 cudaGetErrorString(e__));                             \
        }                                                                                              \
    } while (0)
