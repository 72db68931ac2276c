// cg_comm_*: NCCL communicator per context, LPT assignment, variable-length all-gather (see comm.cuh).
#include "comm.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <numeric>

namespace {

CgNccl g_nccl;
std::string g_nccl_err;
std::once_flag g_nccl_once;

template <typename F>
bool bind(void* lib, const char* name, F& fn) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    if (!fn) g_nccl_err = std::string("libnccl: symbol ") + name + " not found";
    return fn != nullptr;
}

void load_nccl() {
    const char* names[] = {getenv("CANVAS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) { g_nccl_err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return; }
    CgNccl n;
    n.lib = lib;
    if (bind(lib, "ncclGetUniqueId", n.GetUniqueId) && bind(lib, "ncclCommInitRank", n.CommInitRank) &&
        bind(lib, "ncclCommInitAll", n.CommInitAll) && bind(lib, "ncclCommDestroy", n.CommDestroy) &&
        bind(lib, "ncclAllGather", n.AllGather) && bind(lib, "ncclBroadcast", n.Broadcast) &&
        bind(lib, "ncclGroupStart", n.GroupStart) && bind(lib, "ncclGroupEnd", n.GroupEnd) &&
        bind(lib, "ncclGetErrorString", n.GetErrorString) && bind(lib, "ncclGetVersion", n.GetVersion))
        g_nccl = n;
}



void comm_free_buffers(CgComm* c) {
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->h_recv) cudaFreeHost(c->h_recv);
    if (c->h_send) cudaFreeHost(c->h_send);
    c->d_send = c->d_recv = c->h_recv = c->h_send = nullptr;
    c->cap_ints = 0;
    if (c->d2_send) cudaFree(c->d2_send);
    if (c->d2_recv) cudaFree(c->d2_recv);
    for (int32_t* p : c->retired) cudaFree(p);
    c->retired.clear();
    c->d2_send = c->d2_recv = nullptr;
    c->cap2_ints = 0;
}

int comm_attach(cg_ctx* ctx, ncclComm_t comm, int rank, int size) {
    CgComm* c = new CgComm();
    c->comm = comm;
    c->rank = rank;
    c->size = size;
    if (const char* e = getenv("CANVAS_COMM_PACK_INTS")) {
        const long v = atol(e);
        if (v >= CG_COMM_PACK_MIN && v <= (1 << 24)) c->pack_ints = (int)v;
    }
    ctx->comm = c;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    CG_CUDA(ctx, cudaEventCreate(&c->ev0));
    CG_CUDA(ctx, cudaEventCreate(&c->ev1));
    return comm_reserve(ctx, 0);
}

}  // namespace

const CgNccl* cg_nccl(std::string* err) {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.lib) {
        if (err) *err = g_nccl_err;
        return nullptr;
    }
    return &g_nccl;
}

// First-round buffers are allocated once per communicator and never move: page-locked memory must not be allocated
// or freed while another rank of the same process may sit in a collective waiting for this one (cudaFreeHost waits for
// every device of the process).  The rare second round uses device buffers that only grow (old ones are retired until
// cg_comm_destroy) and pageable host memory.
int comm_reserve(cg_ctx* ctx, size_t cap_ints) {
    CgComm* c = ctx->comm;
    if (!c) return cg_fail(ctx, CG_ERR_ARG, "no communicator: call cg_comm_init first");
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!c->d_send) {
        const size_t cap = (size_t)c->pack_ints;
        CG_CUDA(ctx, cudaMalloc((void**)&c->d_send, cap * 4));
        CG_CUDA(ctx, cudaMalloc((void**)&c->d_recv, cap * 4 * (size_t)c->size));
        CG_CUDA(ctx, cudaMallocHost((void**)&c->h_recv, cap * 4 * (size_t)c->size));
        CG_CUDA(ctx, cudaMallocHost((void**)&c->h_send, cap * 4));
        CG_CUDA(ctx, cudaMemset(c->d_send, 0, cap * 4));
        c->cap_ints = cap;
    }
    if (cap_ints == 0 || cap_ints <= c->cap2_ints) return CG_OK;  // cap_ints: what the second round has to hold per rank
    if (c->d2_send) c->retired.push_back(c->d2_send);
    if (c->d2_recv) c->retired.push_back(c->d2_recv);
    c->d2_send = c->d2_recv = nullptr;
    c->cap2_ints = 0;
    const size_t cap = cap_ints + (cap_ints >> 2);
    CG_CUDA(ctx, cudaMalloc((void**)&c->d2_send, cap * 4));
    CG_CUDA(ctx, cudaMalloc((void**)&c->d2_recv, cap * 4 * (size_t)c->size));
    c->cap2_ints = cap;
    return CG_OK;
}

void comm_assign_lpt(int n_units, const int64_t* weight, int n_ranks, int32_t* owner) {
    std::vector<int> order(n_units);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<int64_t> load(std::max(n_ranks, 1), 0);
    for (int u : order) {
        int best = 0;
        for (int r = 1; r < n_ranks; r++)
            if (load[r] < load[best]) best = r;
        owner[u] = best;
        load[best] += weight[u];
    }
}

static __global__ void comm_lengths_kernel(const int32_t* __restrict__ d_recv, int32_t* __restrict__ h_recv, int cap, int n_ranks) {
    for (int r = threadIdx.x; r < n_ranks; r += blockDim.x) h_recv[(size_t)r * cap] = d_recv[(size_t)r * cap];
    __threadfence_system();
}

int comm_allgatherv(cg_ctx* ctx, const int32_t* local, int64_t n_local, const std::function<const int32_t*()>& fetch_local_full,
                    std::vector<int64_t>& counts, std::vector<int32_t>& all) {
    CgComm* c = ctx->comm;
    if (!c) return cg_fail(ctx, CG_ERR_ARG, "no communicator: call cg_comm_init first");
    const CgNccl* nc = nullptr;
    if (c->comm) {
        nc = cg_nccl(&ctx->err);
        if (!nc) return CG_ERR_CUDA;
    }
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int R = c->size;
    const int CAP = c->pack_ints;
    int rc = comm_reserve(ctx, 0);
    if (rc) return rc;
    if (n_local < 0 || n_local > 0x7ffffff0LL) return cg_fail(ctx, CG_ERR_ARG, "comm_allgatherv: bad list length");
    if (local) {
        c->h_send[0] = (int32_t)n_local;
        const int64_t first = std::min<int64_t>(n_local, CAP - 1);
        if (first > 0) memcpy(c->h_send + 1, local, (size_t)first * 4);
        CG_CUDA(ctx, cudaMemcpyAsync(c->d_send, c->h_send, (size_t)(first + 1) * 4, cudaMemcpyHostToDevice, s));
    }
    // ---- round 1: fixed capacity
    CG_CUDA(ctx, cudaEventRecord(c->ev0, s));
    if (c->comm) CG_NCCL(ctx, nc, nc->AllGather(c->d_send, c->d_recv, (size_t)CAP, ncclInt32, c->comm, s));
    else CG_CUDA(ctx, cudaMemcpyAsync(c->d_recv, c->d_send, (size_t)CAP * 4, cudaMemcpyDeviceToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(c->ev1, s));
    // lengths first (one strided copy), then exactly the used part of every rank's block
    // (read back by kernel stores into the page-locked mirror, not by the copy engine: see cg_readback_small)
    comm_lengths_kernel<<<1, 64, 0, s>>>(c->d_recv, c->h_recv, (int)CAP, R);
    CG_CUDA(ctx, cudaGetLastError());
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    counts.assign(R, 0);
    int64_t maxn = 0, total = 0;
    for (int r = 0; r < R; r++) {
        counts[r] = c->h_recv[(size_t)r * CAP];
        if (counts[r] < 0) return cg_fail(ctx, CG_ERR_CUDA, "comm_allgatherv: corrupt length received");
        maxn = std::max(maxn, counts[r]);
        total += counts[r];
    }
    all.resize((size_t)total);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_exchange_ms = ms;
    if (maxn <= CAP - 1) {
        for (int r = 0; r < R; r++)
            if (counts[r] > 0)
                CG_CUDA(ctx, cg_readback_small(s, c->h_recv + (size_t)r * CAP + 1, c->d_recv + (size_t)r * CAP + 1, (size_t)counts[r] * 4));
        CG_CUDA(ctx, cudaStreamSynchronize(s));
        size_t at = 0;
        for (int r = 0; r < R; r++) {
            if (counts[r] > 0) memcpy(all.data() + at, c->h_recv + (size_t)r * CAP + 1, (size_t)counts[r] * 4);
            at += (size_t)counts[r];
        }
        return CG_OK;
    }
    // ---- round 2 (every rank sees the same lengths, so every rank gets here together): exact sizes, one broadcast per rank
    const int64_t mine = counts[c->rank];
    rc = comm_reserve(ctx, (size_t)maxn);
    if (rc) return rc;
    const size_t stride = c->cap2_ints;
    if (local || mine > CAP - 1) {
        const int32_t* src = local ? local : (fetch_local_full ? fetch_local_full() : nullptr);
        if (!src) return cg_fail(ctx, CG_ERR_CAPACITY, "comm_allgatherv: the packed list overflowed and no full copy was provided");
        if (mine > 0) CG_CUDA(ctx, cudaMemcpyAsync(c->d2_send, src, (size_t)mine * 4, cudaMemcpyHostToDevice, s));
    } else if (mine > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(c->d2_send, c->d_send + 1, (size_t)mine * 4, cudaMemcpyDeviceToDevice, s));
    }
    CG_CUDA(ctx, cudaEventRecord(c->ev0, s));
    if (c->comm) {
        CG_NCCL(ctx, nc, nc->GroupStart());
        for (int r = 0; r < R; r++)
            if (counts[r] > 0)
                CG_NCCL(ctx, nc, nc->Broadcast(c->d2_send, c->d2_recv + (size_t)r * stride, (size_t)counts[r], ncclInt32, r, c->comm, s));
        CG_NCCL(ctx, nc, nc->GroupEnd());
    } else if (mine > 0) {
        CG_CUDA(ctx, cudaMemcpyAsync(c->d2_recv, c->d2_send, (size_t)mine * 4, cudaMemcpyDeviceToDevice, s));
    }
    CG_CUDA(ctx, cudaEventRecord(c->ev1, s));
    size_t at = 0;
    for (int r = 0; r < R; r++) {  // straight into the (pageable) result
        if (counts[r] > 0)
            CG_CUDA(ctx, cudaMemcpyAsync(all.data() + at, c->d2_recv + (size_t)r * stride, (size_t)counts[r] * 4, cudaMemcpyDeviceToHost, s));
        at += (size_t)counts[r];
    }
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_exchange_ms += ms;
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int cg_comm_unique_id(uint8_t* id) {
    if (!id) return CG_ERR_ARG;
    std::string err;
    const CgNccl* nc = cg_nccl(&err);
    if (!nc) return CG_ERR_CUDA;
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) return CG_ERR_CUDA;
    static_assert(sizeof(u) == CG_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof u);
    return CG_OK;
}

extern "C" int cg_comm_destroy(cg_ctx* ctx) {
    if (!ctx) return CG_ERR_ARG;
    CgComm* c = ctx->comm;
    if (!c) return CG_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (c->comm) {
        const CgNccl* nc = cg_nccl(nullptr);
        if (nc) nc->CommDestroy(c->comm);
    }
    comm_free_buffers(c);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
    ctx->comm = nullptr;
    return CG_OK;
}

extern "C" int cg_comm_init(cg_ctx* ctx, int n_ranks, int rank, const uint8_t* id) {
    if (!ctx) return CG_ERR_ARG;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id)) return cg_fail(ctx, CG_ERR_ARG, "cg_comm_init: bad argument");
    cg_comm_destroy(ctx);
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclComm_t comm = nullptr;
    if (n_ranks > 1) {
        const CgNccl* nc = cg_nccl(&ctx->err);
        if (!nc) return CG_ERR_CUDA;
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        CG_NCCL(ctx, nc, nc->CommInitRank(&comm, n_ranks, u, rank));
    }
    return comm_attach(ctx, comm, rank, n_ranks);
}

extern "C" int cg_comm_init_all(int n, cg_ctx* const* ctxs) {
    if (n < 1 || !ctxs) return CG_ERR_ARG;
    for (int i = 0; i < n; i++)
        if (!ctxs[i]) return CG_ERR_ARG;
    cg_ctx* c0 = ctxs[0];
    std::vector<ncclComm_t> comms(n, nullptr);
    if (n > 1) {
        const CgNccl* nc = cg_nccl(&c0->err);
        if (!nc) return CG_ERR_CUDA;
        std::vector<int> devs(n);
        for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
        CG_NCCL(c0, nc, nc->CommInitAll(comms.data(), n, devs.data()));
    }
    for (int i = 0; i < n; i++) {
        cg_comm_destroy(ctxs[i]);
        int rc = comm_attach(ctxs[i], comms[i], i, n);
        if (rc) return rc;
    }
    return CG_OK;
}

extern "C" int cg_comm_rank(cg_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->rank : -1; }
extern "C" int cg_comm_size(cg_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->size : 0; }
extern "C" double cg_comm_last_exchange_ms(cg_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->last_exchange_ms : -1.0; }

extern "C" int cg_comm_nccl_version(void) {
    const CgNccl* nc = cg_nccl(nullptr);
    int v = 0;
    if (!nc || nc->GetVersion(&v) != ncclSuccess) return -1;
    return v;
}

extern "C" int cg_shard_assign(int n_units, const int64_t* weight, int n_ranks, int32_t* owner) {
    if (n_units < 0 || n_ranks < 1 || (n_units > 0 && (!weight || !owner))) return CG_ERR_ARG;
    comm_assign_lpt(n_units, weight, n_ranks, owner);
    return CG_OK;
}

extern "C" int cg_comm_allgather_lists(cg_ctx* ctx, int64_t n_local, const int32_t* local, int64_t* counts, int32_t* all, int64_t cap,
                                       int64_t* n_total) {
    if (!ctx) return CG_ERR_ARG;
    if (n_local < 0 || (n_local > 0 && !local) || !counts || !n_total || cap < 0 || (cap > 0 && !all))
        return cg_fail(ctx, CG_ERR_ARG, "cg_comm_allgather_lists: bad argument");
    std::vector<int64_t> cnt;
    std::vector<int32_t> buf;
    static const int32_t none = 0;
    int rc = comm_allgatherv(ctx, local ? local : &none, n_local, nullptr, cnt, buf);
    if (rc) return rc;
    for (size_t r = 0; r < cnt.size(); r++) counts[r] = cnt[r];
    *n_total = (int64_t)buf.size();
    if ((int64_t)buf.size() > cap) return cg_fail(ctx, CG_ERR_CAPACITY, "cg_comm_allgather_lists: output capacity too small (see *n_total)");
    if (!buf.empty()) memcpy(all, buf.data(), buf.size() * 4);
    return CG_OK;
}

// Broadcast of a host buffer from `root` to every rank, staged through the device (pedigree mode: the rank that cleaned a
// sample hands the cleaned bins to the ranks that segment its chromosomes).
extern "C" int cg_comm_broadcast(cg_ctx* ctx, void* buf, int64_t bytes, int root) {
    if (!ctx) return CG_ERR_ARG;
    CgComm* c = ctx->comm;
    if (!c) return cg_fail(ctx, CG_ERR_ARG, "no communicator: call cg_comm_init first");
    if (bytes < 0 || (bytes > 0 && !buf) || root < 0 || root >= c->size) return cg_fail(ctx, CG_ERR_ARG, "cg_comm_broadcast: bad argument");
    if (bytes == 0 || c->size == 1) return CG_OK;
    const CgNccl* nc = cg_nccl(&ctx->err);
    if (!nc) return CG_ERR_CUDA;
    CG_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = arena_reserve(ctx, (size_t)bytes + 512);
    if (rc) return rc;
    char* d = arena_take<char>(ctx, (size_t)bytes);
    if (!d) return cg_fail(ctx, CG_ERR_CUDA, "cg_comm_broadcast: device arena exhausted");
    cudaStream_t s = ctx->stream;
    if (c->rank == root) CG_CUDA(ctx, cudaMemcpyAsync(d, buf, (size_t)bytes, cudaMemcpyHostToDevice, s));
    CG_CUDA(ctx, cudaEventRecord(c->ev0, s));
    CG_NCCL(ctx, nc, nc->Broadcast(d, d, (size_t)bytes, ncclUint8, root, c->comm, s));
    CG_CUDA(ctx, cudaEventRecord(c->ev1, s));
    if (c->rank != root) CG_CUDA(ctx, cudaMemcpyAsync(buf, d, (size_t)bytes, cudaMemcpyDeviceToHost, s));
    CG_CUDA(ctx, cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_exchange_ms = ms;
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------
// Sharded CBS / HMM: the single-GPU shard call on this rank's chromosomes, then the all-gather of the packed results.
// Their segment lists are finished on the host (undo methods, breakpoint lists), so the packed list starts from host
// memory; the wavelet path packs on the device (wavelet.cu).
// ---------------------------------------------------------------------------------------------
namespace {
int shard_mask(cg_ctx* ctx, const char* who, int n_chrom, const int64_t* chrom_off, std::vector<uint8_t>& mask, int32_t* owner) {
    if (!ctx->comm) return cg_fail(ctx, CG_ERR_ARG, std::string(who) + ": no communicator (cg_comm_init)");
    if (n_chrom < 0 || !chrom_off) return cg_fail(ctx, CG_ERR_ARG, std::string(who) + ": bad argument");
    std::vector<int64_t> w(n_chrom);
    std::vector<int32_t> own(n_chrom);
    for (int c = 0; c < n_chrom; c++) w[c] = chrom_off[c + 1] - chrom_off[c];
    comm_assign_lpt(n_chrom, w.data(), ctx->comm->size, own.data());
    mask.assign((size_t)n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++) { mask[c] = own[c] == ctx->comm->rank; if (owner) owner[c] = own[c]; }
    return CG_OK;
}
}  // namespace

extern "C" int cg_partition_cbs_sharded(cg_ctx* ctx, const cg_cbs_opts* opts, const uint32_t* sbdry, int64_t n_sbdry, int n_chrom,
                                        const int64_t* chrom_off, const double* coverage, int32_t* n_seg, int32_t* seg_len,
                                        double* seg_mean, int64_t* stats, int32_t* owner) {
    if (!ctx) return CG_ERR_ARG;
    std::vector<uint8_t> mask;
    int rc = shard_mask(ctx, "cg_partition_cbs_sharded", n_chrom, chrom_off, mask, owner);
    if (rc) return rc;
    rc = cg_partition_cbs_shard(ctx, opts, sbdry, n_sbdry, n_chrom, chrom_off, coverage, mask.data(), n_seg, seg_len, seg_mean, stats);
    // a rank that failed still enters the collective (with an empty block marked -1), so that nobody waits for it forever
    const int C = n_chrom;
    std::vector<int32_t> pack;
    if (rc == CG_OK) {
        pack.assign(n_seg, n_seg + C);
        for (int c = 0; c < C; c++) {
            const int n = n_seg[c];
            if (n <= 0) continue;
            const size_t at = pack.size();
            pack.resize(at + (size_t)n * 3);
            memcpy(pack.data() + at, seg_len + chrom_off[c], (size_t)n * 4);
            memcpy(pack.data() + at + n, seg_mean + chrom_off[c], (size_t)n * 8);
        }
    } else {
        pack.assign(1, -1);
    }
    const std::string first_err = ctx->err;
    std::vector<int64_t> counts;
    std::vector<int32_t> all;
    int rc2 = comm_allgatherv(ctx, pack.data(), (int64_t)pack.size(), nullptr, counts, all);
    if (rc) { ctx->err = first_err; return rc; }
    if (rc2) return rc2;
    for (int c = 0; c < C; c++) n_seg[c] = 0;
    size_t at = 0;
    for (size_t r = 0; r < counts.size(); r++) {
        const int32_t* blk = all.data() + at;
        at += (size_t)counts[r];
        if (counts[r] == 1 && blk[0] == -1) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_cbs_sharded: rank " + std::to_string(r) + " failed");
        if (counts[r] < C) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_cbs_sharded: malformed block in the all-gather");
        const int32_t* src = blk + C;
        for (int c = 0; c < C; c++) {
            const int n = blk[c];
            if (n <= 0) continue;
            if ((src - blk) + (int64_t)n * 3 > counts[r] || n > chrom_off[c + 1] - chrom_off[c])
                return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_cbs_sharded: malformed block in the all-gather");
            n_seg[c] = n;
            memcpy(seg_len + chrom_off[c], src, (size_t)n * 4);
            memcpy(seg_mean + chrom_off[c], src + n, (size_t)n * 8);
            src += (size_t)n * 3;
        }
    }
    return CG_OK;
}

extern "C" int cg_partition_hmm_sharded(cg_ctx* ctx, const cg_hmm_opts* opts, int n_samples, int n_chrom, const int64_t* chrom_off,
                                        const double* coverage, int32_t* n_bp, int32_t* bp, uint8_t* states, int32_t* owner) {
    if (!ctx) return CG_ERR_ARG;
    std::vector<uint8_t> mask;
    int rc = shard_mask(ctx, "cg_partition_hmm_sharded", n_chrom, chrom_off, mask, owner);
    if (rc) return rc;
    rc = cg_partition_hmm_shard(ctx, opts, n_samples, n_chrom, chrom_off, coverage, mask.data(), n_bp, bp, states);
    const int C = n_chrom;
    std::vector<int32_t> pack;
    if (rc == CG_OK) {
        pack.assign(n_bp, n_bp + C);
        for (int c = 0; c < C; c++) {
            const int n = n_bp[c];
            if (n > 0) pack.insert(pack.end(), bp + chrom_off[c], bp + chrom_off[c] + n);
        }
        if (states)  // the Viterbi path of this rank's chromosomes, four states per int
            for (int c = 0; c < C; c++) {
                if (!mask[c]) continue;
                const int64_t len = chrom_off[c + 1] - chrom_off[c];
                const size_t at = pack.size();
                pack.resize(at + (size_t)((len + 3) / 4), 0);
                memcpy(pack.data() + at, states + chrom_off[c], (size_t)len);
            }
    } else {
        pack.assign(1, -1);
    }
    const std::string first_err = ctx->err;
    std::vector<int64_t> counts;
    std::vector<int32_t> all;
    int rc2 = comm_allgatherv(ctx, pack.data(), (int64_t)pack.size(), nullptr, counts, all);
    if (rc) { ctx->err = first_err; return rc; }
    if (rc2) return rc2;
    std::vector<int64_t> w(C);
    std::vector<int32_t> own(C);
    for (int c = 0; c < C; c++) w[c] = chrom_off[c + 1] - chrom_off[c];
    comm_assign_lpt(C, w.data(), ctx->comm->size, own.data());
    for (int c = 0; c < C; c++) n_bp[c] = 0;
    size_t at = 0;
    for (size_t r = 0; r < counts.size(); r++) {
        const int32_t* blk = all.data() + at;
        const int32_t* end = blk + counts[r];
        at += (size_t)counts[r];
        if (counts[r] == 1 && blk[0] == -1) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_hmm_sharded: rank " + std::to_string(r) + " failed");
        if (counts[r] < C) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_hmm_sharded: malformed block in the all-gather");
        const int32_t* src = blk + C;
        for (int c = 0; c < C; c++) {
            const int n = blk[c];
            if (n <= 0) continue;
            if (src + n > end || n > chrom_off[c + 1] - chrom_off[c])
                return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_hmm_sharded: malformed block in the all-gather");
            n_bp[c] = n;
            memcpy(bp + chrom_off[c], src, (size_t)n * 4);
            src += n;
        }
        if (states)
            for (int c = 0; c < C; c++) {
                if (own[c] != (int)r) continue;
                const int64_t len = chrom_off[c + 1] - chrom_off[c];
                const int64_t ints = (len + 3) / 4;
                if (src + ints > end) return cg_fail(ctx, CG_ERR_CUDA, "cg_partition_hmm_sharded: malformed block in the all-gather");
                memcpy(states + chrom_off[c], src, (size_t)len);
                src += ints;
            }
    }
    return CG_OK;
}
