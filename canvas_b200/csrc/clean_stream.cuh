// K8, bulk-copy form — normalise apply (CanvasClean.cs:190-195) as a persistent stream kernel.
//
// count = (float)(gMed * (double)count / med[gc]) is 9 algorithmic bytes per bin (4 B count + 1 B GC in,
// 4 B out): an HBM stream, once the FP64 divide is kept off the common path (k8_norm).  Each CTA walks tiles of K8_TILE bins; one elected
// thread feeds a ring of K8_STAGES shared-memory stages with 1-D bulk copies (cp.async.bulk, the TMA
// engine: SASS UBLKCP) that complete on an mbarrier per stage, so the loads of the next K8_STAGES-1 tiles
// are in flight while the CTA divides the current one; results leave as coalesced 128-bit streaming stores.
// The 101-entry median table sits in shared memory.  The ragged end of a sample (n % K8_TILE bins, and
// arrays whose base is not 16-byte aligned) goes through plain loads.
#pragma once
#include "clean.cuh"

constexpr int K8_TILE = 2048;   // bins per tile: 8 KB counts + 2 KB GC (+ 2 KB alive mask)
constexpr int K8_STAGES = 4;
constexpr int K8_THREADS = 256;
constexpr int K8_CTAS_PER_SM = 4;  // 4 x 50 KB of stages per SM: measured 72 % (8 samples) / 80 % (16) of HBM peak, 3 CTAs 60 %

struct K8Stage {
    alignas(128) float cnt[K8_TILE];
    alignas(128) uint8_t gc[K8_TILE];
    alignas(128) uint8_t alive[K8_TILE];
};
struct K8Smem {
    K8Stage st[K8_STAGES];
    alignas(8) unsigned long long full[K8_STAGES];
    double med[GC_BINS];
    double rcp[GC_BINS];  // RN(1 / med)
};

__device__ __forceinline__ uint32_t k8_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k8_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void k8_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k8_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void k8_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// (float)(gmed * c / m) without the FP64 divide in the common case.  With r = RN(1/m), q' = RN(p * r) is within 1.5 ulp
// of the reference quotient RN(p / m) (p = RN(gmed * c) as in the reference), so both round to the same float unless a
// float rounding boundary — a double whose low 29 mantissa bits are 1000...0 — lies within a few ulp of q'.  Those
// elements (one in ~2^25), non-normal float ranges and non-finite values take the exact divide.
__device__ __forceinline__ float k8_norm(float c, unsigned g, unsigned a, const double* med, const double* rcp, double gmed) {
    const double m = med[g];
    if (!(a && m > 0)) return c;
    const double p = __dmul_rn(gmed, (double)c);
    if (rcp == nullptr) return (float)__ddiv_rn(p, m);  // experiment switch: always the exact divide
    const double q = __dmul_rn(p, rcp[g]);
    const int low = __double2loint(q) & 0x1fffffff;
    const double aq = fabs(q);
    const bool safe = (unsigned)abs(low - 0x10000000) > 8u && aq > 1e-30 && aq < 1e30;
    return safe ? (float)q : (float)__ddiv_rn(p, m);
}

// grid = (CTAs per sample, samples).  n comes from the device (pipeline: survivors of the filters) or is fixed.
__global__ void __launch_bounds__(K8_THREADS)
normalize_apply_bulk_kernel(const float* in, const uint8_t* __restrict__ gc, const uint8_t* __restrict__ alive,
                            float* out, const int* __restrict__ n_ptr, long long n_fixed,
                            const double* __restrict__ med_tab, const double* __restrict__ gmed_tab,
                            const int* __restrict__ enabled, long long sample_stride, int exact_divide = 0) {
    if (enabled && !*enabled) return;
    extern __shared__ unsigned char k8_raw[];
    K8Smem& sm = *reinterpret_cast<K8Smem*>(((uintptr_t)k8_raw + 127) & ~(uintptr_t)127);
    const int sample = blockIdx.y;
    const int tid = threadIdx.x;
    for (int t = tid; t < GC_BINS; t += K8_THREADS) {
        const double m = med_tab[(size_t)sample * GC_BINS + t];
        sm.med[t] = m;
        sm.rcp[t] = m > 0 ? __ddiv_rn(1.0, m) : 0.0;
    }
    const double gmed = gmed_tab[sample];
    const double* rcp_tab = exact_divide ? nullptr : sm.rcp;
    const long long n = n_ptr ? (long long)*n_ptr : n_fixed;
    in += sample * sample_stride;
    out += sample * sample_stride;
    gc += sample * sample_stride;
    if (alive) alive += sample * sample_stride;
    // bulk copies need 16-byte aligned global addresses; otherwise everything takes the plain path
    const bool bulk_ok = ((((uintptr_t)in) | ((uintptr_t)gc) | ((uintptr_t)alive)) & 15) == 0;
    const int tiles = bulk_ok ? (int)(n / K8_TILE) : 0;
    const uint32_t tile_bytes = K8_TILE * 4 + K8_TILE + (alive ? K8_TILE : 0);

    if (tid == 0) {
        for (int s = 0; s < K8_STAGES; s++) k8_mbar_init(k8_smem_addr(&sm.full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int stage, int tile) {
        const uint32_t bar = k8_smem_addr(&sm.full[stage]);
        const long long base = (long long)tile * K8_TILE;
        k8_mbar_expect(bar, tile_bytes);
        k8_bulk_load(k8_smem_addr(sm.st[stage].cnt), in + base, K8_TILE * 4, bar);
        k8_bulk_load(k8_smem_addr(sm.st[stage].gc), gc + base, K8_TILE, bar);
        if (alive) k8_bulk_load(k8_smem_addr(sm.st[stage].alive), alive + base, K8_TILE, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < K8_STAGES; s++) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < tiles) issue(s, t);
        }
    }
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
        const int s = it % K8_STAGES;
        k8_mbar_wait(k8_smem_addr(&sm.full[s]), (it / K8_STAGES) & 1);
        const float4* c4 = reinterpret_cast<const float4*>(sm.st[s].cnt);
        const uchar4* g4 = reinterpret_cast<const uchar4*>(sm.st[s].gc);
        const uchar4* a4 = reinterpret_cast<const uchar4*>(sm.st[s].alive);
        float4* o4 = reinterpret_cast<float4*>(out + (long long)t * K8_TILE);
        constexpr int PER = K8_TILE / 4 / K8_THREADS;
        float4 c[PER];
        uchar4 g[PER], a[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            c[u] = c4[tid + u * K8_THREADS];
            g[u] = g4[tid + u * K8_THREADS];
            a[u] = alive ? a4[tid + u * K8_THREADS] : make_uchar4(1, 1, 1, 1);
        }
        __syncthreads();  // every thread holds its part of stage s in registers: the stage can be refilled
        if (tid == 0) {
            const int tn = t + K8_STAGES * gridDim.x;
            if (tn < tiles) issue(s, tn);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            c[u].x = k8_norm(c[u].x, g[u].x, a[u].x, sm.med, rcp_tab, gmed);
            c[u].y = k8_norm(c[u].y, g[u].y, a[u].y, sm.med, rcp_tab, gmed);
            c[u].z = k8_norm(c[u].z, g[u].z, a[u].z, sm.med, rcp_tab, gmed);
            c[u].w = k8_norm(c[u].w, g[u].w, a[u].w, sm.med, rcp_tab, gmed);
            __stcs(o4 + tid + u * K8_THREADS, c[u]);
        }
    }
    // ragged end (and unaligned arrays): plain loads, spread over the sample's CTAs
    for (long long i = (long long)tiles * K8_TILE + (long long)blockIdx.x * K8_THREADS + tid; i < n;
         i += (long long)gridDim.x * K8_THREADS)
        out[i] = k8_norm(in[i], gc[i], alive ? alive[i] : 1u, sm.med, rcp_tab, gmed);
}

inline size_t k8_smem_bytes() { return sizeof(K8Smem) + 128; }
