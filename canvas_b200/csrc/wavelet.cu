// cg_partition_wavelet — placeholder until the wavelet kernels land.
#include "common.cuh"

extern "C" int cg_partition_wavelet(cg_ctx* ctx, const cg_wavelet_opts*, int, const int64_t*, const double*, int32_t*,
                                    int32_t*, double*, int*, double*, int*, double*) {
    return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_wavelet: not built yet");
}
extern "C" int cg_partition_wavelet_shard(cg_ctx* ctx, const cg_wavelet_opts*, int, const int64_t*, const double*,
                                          const uint8_t*, int32_t*, int32_t*, double*, int*, double*, int*, double*) {
    return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_partition_wavelet_shard: not built yet");
}
extern "C" int cg_clean_partition_wavelet(cg_ctx* ctx, const cg_clean_opts*, const cg_wavelet_opts*, int64_t,
                                          const uint8_t*, const uint8_t*, const uint8_t*, int, const int32_t*,
                                          const int32_t*, const float*, const uint8_t*, int64_t*, int32_t*, float*,
                                          double*, int*, int64_t*, int32_t*, int32_t*, double*, int*, double*, int*,
                                          double*) {
    return cg_fail(ctx, CG_ERR_UNSUPPORTED, "cg_clean_partition_wavelet: not built yet");
}
