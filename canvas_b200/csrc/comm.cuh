// Multi-GPU exchange of libcanvasgpu: one NCCL communicator per cg_ctx, the LPT chromosome assignment, and the
// all-gather of variable-length result lists (replaces the lock(dict) merges of the per-chromosome
// results, WaveletsRunner.cs:128-131, CBSRunner.cs:140-143, HiddenMarkovModelsRunner.cs:90-105).
//
// NCCL is bound at run time (dlopen "libnccl.so.2" on the first cg_comm_* call): a host that already carries an NCCL
// (a torch process) keeps using that one library, and a single-GPU host needs none.
#pragma once
#include <nccl.h>

#include <functional>

#include "common.cuh"

constexpr int CG_COMM_PACK_INTS = 16384;  // first-round capacity of a rank's packed list: [length, payload ...]
constexpr int CG_COMM_PACK_MIN = 512;     // smallest capacity CANVAS_COMM_PACK_INTS may ask for (tests of the second round)

struct CgNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

struct CgComm {
    ncclComm_t comm = nullptr;  // nullptr with size == 1: loopback (no NCCL call is made)
    int rank = 0, size = 1;
    int pack_ints = CG_COMM_PACK_INTS;  // first-round capacity; CANVAS_COMM_PACK_INTS (same value on every rank) overrides it
    // first-round exchange buffers (device: send [cap], recv [size * cap]; pinned host mirrors), allocated once
    int32_t* d_send = nullptr;
    int32_t* d_recv = nullptr;
    int32_t* h_recv = nullptr;
    int32_t* h_send = nullptr;
    size_t cap_ints = 0;
    // second (exact) round: grow-only device buffers; replaced ones wait in `retired` for cg_comm_destroy
    int32_t* d2_send = nullptr;
    int32_t* d2_recv = nullptr;
    size_t cap2_ints = 0;
    std::vector<int32_t*> retired;
    double last_exchange_ms = 0;  // device time of the last all-gather (events on the ctx stream)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define CG_NCCL(ctx, nc, call)                                                                                     \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess)                                                                                    \
            return cg_fail(ctx, CG_ERR_CUDA, std::string(#call) + ": " + (nc)->GetErrorString(r__));               \
    } while (0)

const CgNccl* cg_nccl(std::string* err);                       // nullptr + message when the library cannot be bound
int comm_reserve(cg_ctx* ctx, size_t cap_ints);                // first-round buffers; second-round buffers for cap_ints ints per rank when > 0
// Greedy longest-processing-time-first: heaviest unit first onto the least loaded rank (ties: lower rank, earlier unit).
void comm_assign_lpt(int n_units, const int64_t* weight, int n_ranks, int32_t* owner);
// All-gather of one int32 list per rank.  The local list is either on the host (`local`) or already packed on the device
// in ctx->comm->d_send as [n_local, payload ...] (local == nullptr; if that payload turns out longer than
// CG_COMM_PACK_INTS - 1 ints, `fetch_local_full` is asked for a host copy of all of it).  One fixed-capacity NCCL all-gather; lists that did not fit travel in
// a second, exactly sized round (grouped broadcasts) that every rank enters or skips together, because the decision is
// taken from the gathered lengths.  Result: counts[size] and the lists back to back in `all`.
int comm_allgatherv(cg_ctx* ctx, const int32_t* local, int64_t n_local, const std::function<const int32_t*()>& fetch_local_full,
                    std::vector<int64_t>& counts, std::vector<int32_t>& all);
