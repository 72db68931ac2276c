"""ctypes binding of libcanvasgpu.so — the only way the Python host layer reaches the GPU.

There is no CPU fallback: importing works anywhere (so the CPU test tier can check the exported
symbols), but creating an Engine without the library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

# the partition runs one stream per chromosome pipeline: let the driver give every one of them its own hardware queue
# (read when the CUDA context is created, so it has to be in the environment before the first CUDA call of the process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcanvasgpu.so")

CG_OK, CG_ERR_CUDA, CG_ERR_ARG, CG_ERR_UNSORTED, CG_ERR_UNSUPPORTED, CG_ERR_CAPACITY = 0, -1, -2, -3, -4, -5


class CanvasGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libcanvasgpu error {code}: {message}")
        self.code = code


class CleanOpts(C.Structure):
    _fields_ = [("size_filter", C.c_int), ("outlier_filter", C.c_int), ("gc_norm", C.c_int),
                ("gc_mode", C.c_int), ("want_local_sd", C.c_int), ("min_bins_per_gc", C.c_int)]


class WaveletOpts(C.Structure):
    _fields_ = [("is_germline", C.c_int), ("mad_factor", C.c_double), ("thr_lower", C.c_double),
                ("thr_upper", C.c_double), ("min_size", C.c_int), ("evenness_window", C.c_int)]


class HmmOpts(C.Structure):
    _fields_ = [("n_states", C.c_int), ("per_sample", C.c_int), ("min_size", C.c_int), ("exact_sequential", C.c_int)]


_P = C.POINTER
_u8, _i32, _i64, _f32, _f64 = C.c_uint8, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol include/canvasgpu.h declares
SIGNATURES = {
    "cg_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "cg_destroy": (None, [C.c_void_p]),
    "cg_last_error": (C.c_char_p, [C.c_void_p]),
    "cg_describe": (C.c_char_p, [C.c_void_p]),
    "cg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "cg_host_free": (None, [C.c_void_p]),
    "cg_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "cg_last_launches": (C.c_int, [C.c_void_p]),
    "cg_last_stage_ms": (C.c_double, [C.c_void_p, C.c_int]),
    "cg_last_partition_stats": (C.c_int, [C.c_void_p, _P(_f64), C.c_int]),
    "cg_clean": (C.c_int, [C.c_void_p, _P(CleanOpts), _i64, _P(_u8), _P(_u8), _P(_u8), C.c_int,
                           _P(_i32), _P(_i32), _P(_f32), _P(_u8), _P(_i64), _P(_i32), _P(_f32),
                           _P(_f64), _P(C.c_int)]),
    "cg_partition_wavelet": (C.c_int, [C.c_void_p, _P(WaveletOpts), C.c_int, _P(_i64), _P(_f64),
                                       _P(_i32), _P(_i32), _P(_f64), _P(C.c_int), _P(_f64),
                                       _P(C.c_int), _P(_f64)]),
    "cg_partition_wavelet_shard": (C.c_int, [C.c_void_p, _P(WaveletOpts), C.c_int, _P(_i64),
                                             _P(_f64), _P(_u8), _P(_i32), _P(_i32), _P(_f64),
                                             _P(C.c_int), _P(_f64), _P(C.c_int), _P(_f64)]),
    "cg_clean_partition_wavelet": (C.c_int, [C.c_void_p, _P(CleanOpts), _P(WaveletOpts), _i64,
                                             _P(_u8), _P(_u8), _P(_u8), C.c_int, _P(_i32), _P(_i32),
                                             _P(_f32), _P(_u8), _P(_i64), _P(_i32), _P(_f32),
                                             _P(_f64), _P(C.c_int), _P(_i64), _P(_i32), _P(_i32),
                                             _P(_f64), _P(C.c_int), _P(_f64), _P(C.c_int), _P(_f64)]),
    "cg_clean_partition_wavelet_shard": (C.c_int, [C.c_void_p, _P(CleanOpts), _P(WaveletOpts), _i64,
                                                   _P(_u8), _P(_u8), _P(_u8), C.c_int, _P(_i32), _P(_i32),
                                                   _P(_f32), _P(_u8), _P(_u8), _P(_i64), _P(_i32), _P(_f32),
                                                   _P(_f64), _P(C.c_int), _P(_i64), _P(_i32), _P(_i32),
                                                   _P(_f64), _P(C.c_int), _P(_f64), _P(C.c_int), _P(_f64)]),
    "cg_partition_cbs": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_uint32), _i64, C.c_int, _P(_i64), _P(C.c_double), _P(_i32),
                                   _P(_i32), _P(C.c_double), _P(_i64)]),
    "cg_partition_cbs_shard": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_uint32), _i64, C.c_int, _P(_i64), _P(C.c_double), _P(_u8),
                                         _P(_i32), _P(_i32), _P(C.c_double), _P(_i64)]),
    "cg_partition_hmm": (C.c_int, [C.c_void_p, _P(HmmOpts), C.c_int, C.c_int, _P(_i64), _P(_f64), _P(_i32), _P(_i32), _P(_u8)]),
    "cg_partition_hmm_shard": (C.c_int, [C.c_void_p, _P(HmmOpts), C.c_int, C.c_int, _P(_i64), _P(_f64), _P(_u8), _P(_i32),
                                         _P(_i32), _P(_u8)]),
    "cg_partition_hmm_counts": (C.c_int, [C.c_void_p, _P(HmmOpts), C.c_int, C.c_int, _P(_i64), _P(_f32), C.c_int, _P(_u8), _P(_i32),
                                          _P(_i32), _P(_u8)]),
    "cg_merge_common_bins": (C.c_int, [C.c_void_p, C.c_int, _P(_i64), _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p),
                                       _P(C.c_void_p), _P(_i64), _P(_i32), _P(_i32), _P(_f32)]),
    "cg_merge_kept_indices": (C.c_int, [C.c_void_p, _i64, C.c_int, _P(_i64), _P(C.c_void_p), _P(C.c_void_p), _P(_i64), _P(_i32), _P(_f32)]),
    "cg_prefetch_bins": (C.c_int, [C.c_void_p, _i64, _P(_u8), _P(_i32), _P(_i32), _P(_f32), _P(_u8)]),
    "cg_pedigree_hmm": (C.c_int, [C.c_void_p, _P(CleanOpts), _P(HmmOpts), C.c_int, _i64, _P(_u8), _P(_u8), _P(_u8), C.c_int, _P(_i32), _P(_i32),
                                  _P(C.c_void_p), _P(_u8), C.c_int, _P(_i64), _P(_f64), _P(C.c_int), _P(_i64), _P(_i32), _P(_f32), _P(_i64),
                                  _P(_i32), _P(_i32), _P(_i32)]),
    "cg_smooth": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(_i64), _P(_f32), _P(_i64), _P(_f32)]),
    "cg_format_bins": (_i64, [_i64, C.c_int, _P(C.c_char_p), _P(_u8), _P(_i32), _P(_i32), _P(_f32), _P(_u8), C.c_int,
                              C.c_char_p, _i64, C.c_int]),
    "cg_parse_bins": (_i64, [C.c_char_p, _i64, _i64, _P(_u8), _P(_i32), _P(_i32), _P(_f32), _P(_u8), _P(C.c_int),
                             C.c_char_p, _i64, C.c_int]),
    "cg_normalize_reference": (C.c_int, [C.c_void_p, C.c_int, _i64, _P(_f64), _P(_u8), _P(_f64), _P(_f64), _P(_f64)]),
    "cg_normalize_best_lr2": (C.c_int, [C.c_void_p, C.c_int, _i64, _P(_f64), _P(_f64), _P(_u8), _P(C.c_int), _P(_f64), _P(_i64)]),
    "cg_normalize_pca_reference": (C.c_int, [C.c_void_p, _i64, C.c_int, _P(_f32), _P(_f32), _P(_f64), _P(_u8), C.c_double, C.c_double,
                                             _P(_f32), _P(_f64)]),
    "cg_normalize_ratio": (C.c_int, [C.c_void_p, _i64, _P(_f32), _P(_f32), _P(_u8), C.c_int, C.c_double, C.c_double, _P(_i32),
                                     _P(_i64), _P(_i32), _P(_f32), _P(_f32), _P(_f64)]),
    "cg_cbs_boundary": (_i64, [C.c_uint32, C.c_double, C.c_double, _P(C.c_uint32), _i64]),
    "cg_cbs_prune": (C.c_int, [_P(_f64), _i64, _P(_i32), C.c_int, C.c_double, _i64, _P(_i32), _P(_i64)]),
    "cg_bin_screen": (C.c_int, [C.c_void_p, _i64, _P(_u8), _P(C.c_uint64), _i64, _P(_i32), _P(_i32), _P(_i64), _P(_i64)]),
    "cg_bin_fragment_stats": (C.c_int, [C.c_void_p, _i64, _P(C.c_int16), _P(_i64), _P(_i64)]),
    "cg_bin_read_gc": (C.c_int, [C.c_void_p, _i64, C.c_char_p, _P(C.c_int16), C.c_int, _P(_u8), _P(_u8), _P(_i64), _P(_i64)]),
    "cg_bin_hits": (C.c_int, [C.c_void_p, _i64, _P(_u8), _P(C.c_uint64), C.c_char_p, C.c_int, C.c_int, _P(_u8),
                              _P(_f32), _i64, _P(_i64), _P(_i32), _P(_i32), _P(_i32), _P(_u8)]),
    "cg_bin_fragments": (C.c_int, [C.c_void_p, _i64, _P(_i32), _P(_i32), _i64, _P(_i32), _i64, _P(_i32), _P(_i32),
                                   _P(_i32), _P(_i32)]),
    "cg_normalize_apply": (C.c_int, [C.c_void_p, C.c_int, _i64, _P(_f32), _P(_u8), _P(_f64),
                                     _P(_f64), _P(_f32), C.c_int, _P(_f64)]),
    # multi-GPU (NCCL communicator per context)
    "cg_comm_unique_id": (C.c_int, [_P(_u8)]),
    "cg_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(_u8)]),
    "cg_comm_init_all": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "cg_comm_destroy": (C.c_int, [C.c_void_p]),
    "cg_comm_rank": (C.c_int, [C.c_void_p]),
    "cg_comm_size": (C.c_int, [C.c_void_p]),
    "cg_comm_nccl_version": (C.c_int, []),
    "cg_comm_last_exchange_ms": (C.c_double, [C.c_void_p]),
    "cg_shard_assign": (C.c_int, [C.c_int, _P(_i64), C.c_int, _P(_i32)]),
    "cg_comm_allgather_lists": (C.c_int, [C.c_void_p, _i64, _P(_i32), _P(_i64), _P(_i32), _i64, _P(_i64)]),
    "cg_comm_broadcast": (C.c_int, [C.c_void_p, C.c_void_p, _i64, C.c_int]),
    "cg_partition_wavelet_sharded": (C.c_int, [C.c_void_p, _P(WaveletOpts), C.c_int, _P(_i64), _P(_f64), _P(_i32), _P(_i32),
                                               _P(_f64), _P(C.c_int), _P(_f64), _P(C.c_int), _P(_f64), _P(_i32)]),
    "cg_clean_partition_wavelet_sharded": (C.c_int, [C.c_void_p, _P(CleanOpts), _P(WaveletOpts), _i64,
                                                     _P(_u8), _P(_u8), _P(_u8), C.c_int, _P(_i32), _P(_i32),
                                                     _P(_f32), _P(_u8), _P(_i64), _P(_i32), _P(_f32),
                                                     _P(_f64), _P(C.c_int), _P(_i64), _P(_i32), _P(_i32),
                                                     _P(_f64), _P(C.c_int), _P(_f64), _P(C.c_int), _P(_f64), _P(_i32)]),
    "cg_partition_cbs_sharded": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_uint32), _i64, C.c_int, _P(_i64), _P(C.c_double),
                                           _P(_i32), _P(_i32), _P(C.c_double), _P(_i64), _P(_i32)]),
    "cg_partition_hmm_sharded": (C.c_int, [C.c_void_p, _P(HmmOpts), C.c_int, C.c_int, _P(_i64), _P(_f64), _P(_i32), _P(_i32),
                                           _P(_u8), _P(_i32)]),
}

_lib = None


def load():
    """Load the shared library and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CanvasGpuError(CG_ERR_CUDA, f"{LIB_PATH} is missing: run `python -m canvas_b200.build` "
                                 "(__graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(_P(t))


def shard_assign(weights, n_ranks):
    """cg_shard_assign: longest-processing-time-first owner of every unit (host code; no device needed)."""
    w = np.ascontiguousarray(weights, np.int64)
    owner = np.zeros(max(len(w), 1), np.int32)
    rc = load().cg_shard_assign(len(w), _ptr(w, _i64), int(n_ranks), _ptr(owner, _i32))
    if rc != CG_OK:
        raise CanvasGpuError(rc, "cg_shard_assign: bad argument")
    return owner[:len(w)]


def comm_unique_id():
    """cg_comm_unique_id (ncclGetUniqueId): 128 bytes rank 0 creates and the host hands to every rank."""
    ident = np.zeros(128, np.uint8)
    rc = load().cg_comm_unique_id(_ptr(ident, _u8))
    if rc != CG_OK:
        raise CanvasGpuError(rc, "cg_comm_unique_id: libnccl.so.2 could not be bound")
    return ident


def format_bins(names, chrom, start, stop, count, gc=None, four_columns=False, n_threads=0):
    """cg_format_bins: the text of a .binned / .cleaned file (bytes)."""
    lib = load()
    chrom = np.ascontiguousarray(chrom, np.uint8)
    start = np.ascontiguousarray(start, np.int32)
    stop = np.ascontiguousarray(stop, np.int32)
    count = np.ascontiguousarray(count, np.float32)
    n = len(count)
    g = None if gc is None else np.ascontiguousarray(gc, np.uint8)
    if g is None and not four_columns:
        raise ValueError("the five-column layout needs GC")
    arr = (C.c_char_p * max(len(names), 1))(*[str(x).encode() for x in names])
    args = (n, len(names), arr, _ptr(chrom, _u8), _ptr(start, _i32), _ptr(stop, _i32), _ptr(count, _f32),
            None if g is None else _ptr(g, _u8), int(four_columns))
    cap = n * (max([len(str(x)) for x in names] + [1]) + 64) + 1  # a line is at most name + 2 x 11 + 48 + 3 + 5 characters
    buf = np.empty(cap, np.uint8)  # not zero-filled
    got = lib.cg_format_bins(*args, buf.ctypes.data_as(C.c_char_p), cap, n_threads)
    if got < 0 or got > cap:
        raise CanvasGpuError(CG_ERR_ARG, "cg_format_bins: bad argument (chromosome id outside the name table?)")
    return buf[:got].tobytes()


def parse_bins(text, n_threads=0):
    """cg_parse_bins: (run names, chrom run ids u8, start i32, stop i32, count f32, gc u8) of a .binned / .cleaned text."""
    lib = load()
    data = bytes(text)
    cap = data.count(b"\n") + 1
    chrom = np.zeros(cap, np.uint8)
    start = np.zeros(cap, np.int32)
    stop = np.zeros(cap, np.int32)
    count = np.zeros(cap, np.float32)
    gc = np.zeros(cap, np.uint8)
    n_names = C.c_int(0)
    names = C.create_string_buffer(1 << 16)
    n = lib.cg_parse_bins(data, len(data), cap, _ptr(chrom, _u8), _ptr(start, _i32), _ptr(stop, _i32), _ptr(count, _f32),
                          _ptr(gc, _u8), C.byref(n_names), names, len(names), n_threads)
    if n == -2:
        raise ValueError("malformed line in bin file")
    if n == -3:
        raise ValueError("more than 256 chromosome runs (contigs, or chromosomes that reappear later in the file): this build "
                         "addresses chromosomes with 8-bit ids (DESIGN.md, Limits)")
    if n < 0 or n > cap:
        raise CanvasGpuError(CG_ERR_ARG, "cg_parse_bins failed")
    parts = names.raw.split(b"\0")[:n_names.value]
    return [p.decode() for p in parts], chrom[:n], start[:n], stop[:n], count[:n], gc[:n]


def cbs_prune(g, seg_len, cutoff=0.05, max_subsets=0):
    """ChangePointsPrune (ChangePoint.cs:205-271) on one chromosome; host code in the library, no device needed.
    Returns (new segment lengths, subsets scored)."""
    g = np.ascontiguousarray(g, np.float64)
    ln = np.ascontiguousarray(seg_len, np.int32)
    out = np.zeros(len(ln), np.int32)
    scored = C.c_int64(0)
    k = load().cg_cbs_prune(_ptr(g, _f64), len(g), _ptr(ln, _i32), len(ln), cutoff, max_subsets,
                            _ptr(out, _i32), C.byref(scored))
    if k < 1:
        raise CanvasGpuError(k, "cg_cbs_prune: bad arguments" if k != CG_ERR_UNSUPPORTED else "cg_cbs_prune: more change-point subsets to score than allowed")
    return out[:k].copy(), int(scored.value)


class PinnedPool:
    """numpy arrays over page-locked memory from cg_host_alloc."""

    def __init__(self, lib):
        self._lib = lib
        self._blocks = []

    def empty(self, n, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(1, int(n) * dtype.itemsize)
        p = self._lib.cg_host_alloc(nbytes)
        if not p:
            raise CanvasGpuError(CG_ERR_CUDA, "cg_host_alloc failed")
        self._blocks.append(p)
        buf = (C.c_char * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def array(self, src, dtype=None):
        src = np.asarray(src, dtype=dtype)
        a = self.empty(src.size, src.dtype)
        a[...] = src.ravel()
        return a

    def close(self):
        for p in self._blocks:
            self._lib.cg_host_free(p)
        self._blocks = []


class CbsOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("n_perm", C.c_uint32), ("hybrid", C.c_int), ("min_width", C.c_int), ("k_max", C.c_int),
                ("n_min", C.c_uint32), ("eta", C.c_double), ("trim", C.c_double), ("undo", C.c_int), ("undo_prune", C.c_double),
                ("undo_sd", C.c_double), ("seed", C.c_uint32)]


def bind_host_near_gpu(device=0):
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that the page-locked host
    buffers allocated afterwards and the launch path sit on the socket the GPU hangs off: with one process per GPU on a
    two-socket host, eight 43 MB uploads per step otherwise share one socket's memory controllers.  A placement hint only:
    returns the CPU list it set, or None when NVML / the affinity call is unavailable (nothing else changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        index = int(device)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                index = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class Engine:
    """One cg_ctx (one GPU)."""

    def __init__(self, device=0, bind_numa=False):
        self.host_cpus = bind_host_near_gpu(device) if bind_numa else None
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.cg_create(device, C.byref(h))
        if rc != CG_OK or not h:
            raise CanvasGpuError(rc, f"cg_create(device={device}) failed: no usable CUDA device "
                                 "(libcanvasgpu has no CPU fallback)")
        self.h = h
        self.device = device
        self.pinned = PinnedPool(self.lib)

    def close(self):
        if getattr(self, "h", None):
            self.pinned.close()
            self.lib.cg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def describe(self):
        return self.lib.cg_describe(self.h).decode()

    # ------------------------------------------------------------------ multi-GPU communicator
    def comm_init(self, n_ranks, rank, unique_id=None):
        """cg_comm_init: this context joins an NCCL communicator of n_ranks (unique_id from comm_unique_id() on rank 0,
        handed over by the host; n_ranks == 1 is a loopback communicator that never calls NCCL)."""
        ident = None if unique_id is None else np.ascontiguousarray(unique_id, np.uint8)
        self._check(self.lib.cg_comm_init(self.h, int(n_ranks), int(rank), None if ident is None else _ptr(ident, _u8)))

    def comm_init_torch(self):
        """One process per GPU under torchrun: rank 0 creates the id, torch.distributed only carries its 128 bytes."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        if world == 1:
            return self.comm_init(1, 0)
        ident = comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
        dev = f"cuda:{self.device}" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(ident).to(dev)
        dist.broadcast(t, 0)
        self.comm_init(world, rank, t.cpu().numpy())

    @property
    def comm_size(self):
        return self.lib.cg_comm_size(self.h)

    @property
    def comm_rank(self):
        return self.lib.cg_comm_rank(self.h)

    @property
    def last_exchange_ms(self):
        return self.lib.cg_comm_last_exchange_ms(self.h)

    def allgather_lists(self, local):
        """cg_comm_allgather_lists: one int32 list per rank -> the list of every rank, on every rank."""
        loc = np.ascontiguousarray(local, np.int32)
        size = max(self.comm_size, 1)
        counts = np.zeros(size, np.int64)
        total = _i64(0)

        def gather(lst, cap):
            out = np.zeros(max(cap, 1), np.int32)
            self._check(self.lib.cg_comm_allgather_lists(self.h, len(lst), _ptr(lst, _i32), _ptr(counts, _i64), _ptr(out, _i32),
                                                         cap, C.byref(total)))
            return out

        # the output capacity has to be right on every rank at once (a retry would be a collective of its own), so the
        # lengths travel first
        lens = gather(np.array([len(loc)], np.int32), size)[:size]
        out = gather(loc, int(lens.sum()))
        ends = np.cumsum(counts)
        return [out[e - c:e].copy() for c, e in zip(counts, ends)]

    def broadcast(self, array, root):
        """cg_comm_broadcast of a contiguous numpy array, in place."""
        a = np.ascontiguousarray(array)
        self._check(self.lib.cg_comm_broadcast(self.h, a.ctypes.data_as(C.c_void_p), a.nbytes, int(root)))
        return a

    def _check(self, rc):
        if rc != CG_OK:
            raise CanvasGpuError(rc, self.lib.cg_last_error(self.h).decode())

    @property
    def last_kernel_ms(self):
        return self.lib.cg_last_kernel_ms(self.h)

    @property
    def last_launches(self):
        return self.lib.cg_last_launches(self.h)

    def last_stage_ms(self):
        names = ("clean", "scalars", "decompose", "finish", "between_clean_and_wait", "host_wait_gap")
        out = {k: self.lib.cg_last_stage_ms(self.h, i) for i, k in enumerate(names)}
        return {k: v for k, v in out.items() if v >= 0 or k in names[:4]}

    def last_partition_stats(self):
        out = np.zeros(16, np.float64)
        self.lib.cg_last_partition_stats(self.h, _ptr(out, _f64), 16)
        return {"visits": out[0], "nodes": out[1], "candidates": out[2], "bins": out[3],
                "visits_big": out[4], "visits_small": out[5], "visits_tiny": out[6],
                "nodes_big": out[7], "nodes_small": out[8], "nodes_tiny": out[9],
                "big_phase_ms": out[10], "decompose_span_ms": out[11], "multi_chunk_nodes": out[12],
                "queue_hops": out[13], "max_depth": out[14], "integer_keys": bool(out[15])}

    # ------------------------------------------------------------------ CanvasClean
    def clean(self, chrom, is_autosome, is_chr_y, start, stop, count, gc, size_filter=True,
              outlier_filter=True, gc_norm=True, gc_mode=0, want_local_sd=True, min_bins_per_gc=100,
              out=None):
        n = len(count)
        chrom = np.ascontiguousarray(chrom, np.uint8)
        is_autosome = np.ascontiguousarray(is_autosome, np.uint8)
        is_chr_y = np.ascontiguousarray(is_chr_y, np.uint8)
        start = np.ascontiguousarray(start, np.int32)
        stop = np.ascontiguousarray(stop, np.int32)
        count = np.ascontiguousarray(count, np.float32)
        gc = np.ascontiguousarray(gc, np.uint8)
        o = CleanOpts(int(size_filter), int(outlier_filter), int(gc_norm), int(gc_mode),
                      int(want_local_sd), int(min_bins_per_gc))
        if out is None:
            kept = np.empty(max(n, 1), np.int32)
            cnt = np.empty(max(n, 1), np.float32)
        else:
            kept, cnt = out
        n_out = _i64(0)
        lsd = _f64(0)
        skipped = C.c_int(0)
        rc = self.lib.cg_clean(self.h, C.byref(o), n, _ptr(chrom, _u8), _ptr(is_autosome, _u8),
                               _ptr(is_chr_y, _u8), len(is_autosome), _ptr(start, _i32),
                               _ptr(stop, _i32), _ptr(count, _f32), _ptr(gc, _u8), C.byref(n_out),
                               _ptr(kept, _i32), _ptr(cnt, _f32), C.byref(lsd), C.byref(skipped))
        self._check(rc)
        k = n_out.value
        return {"kept_index": kept[:k], "count": cnt[:k], "local_sd": lsd.value,
                "gc_norm_skipped": bool(skipped.value)}

    # ------------------------------------------------------------------ CanvasPartition (wavelets)
    def partition_wavelet(self, chrom_off, coverage, is_germline=True, mad_factor=5.0,
                          thr_lower=0.05, thr_upper=80.0, min_size=10, evenness_window=100000,
                          chrom_selected=None, out=None, sharded=False):
        """sharded=True: cg_partition_wavelet_sharded (LPT over the communicator's ranks + NCCL all-gather inside the call)."""
        chrom_off = np.ascontiguousarray(chrom_off, np.int64)
        coverage = np.ascontiguousarray(coverage, np.float64)
        nc = len(chrom_off) - 1
        n = int(chrom_off[-1])
        o = WaveletOpts(int(is_germline), mad_factor, thr_lower, thr_upper, min_size, evenness_window)
        if out is None:
            n_bp = np.zeros(max(nc, 1), np.int32)
            bp = np.zeros(max(n, 1), np.int32)
        else:
            n_bp, bp = out
        ev, cv = _f64(0), _f64(0)
        ev_ok, cv_has = C.c_int(0), C.c_int(0)
        f3 = np.zeros(9, np.float64)
        owner = np.zeros(max(nc, 1), np.int32)
        if sharded:
            rc = self.lib.cg_partition_wavelet_sharded(self.h, C.byref(o), nc, _ptr(chrom_off, _i64), _ptr(coverage, _f64),
                                                       _ptr(n_bp, _i32), _ptr(bp, _i32), C.byref(ev), C.byref(ev_ok), C.byref(cv),
                                                       C.byref(cv_has), _ptr(f3, _f64), _ptr(owner, _i32))
        elif chrom_selected is None:
            rc = self.lib.cg_partition_wavelet(self.h, C.byref(o), nc, _ptr(chrom_off, _i64),
                                               _ptr(coverage, _f64), _ptr(n_bp, _i32), _ptr(bp, _i32),
                                               C.byref(ev), C.byref(ev_ok), C.byref(cv),
                                               C.byref(cv_has), _ptr(f3, _f64))
        else:
            sel = np.ascontiguousarray(chrom_selected, np.uint8)
            rc = self.lib.cg_partition_wavelet_shard(self.h, C.byref(o), nc, _ptr(chrom_off, _i64),
                                                     _ptr(coverage, _f64), _ptr(sel, _u8),
                                                     _ptr(n_bp, _i32), _ptr(bp, _i32), C.byref(ev),
                                                     C.byref(ev_ok), C.byref(cv), C.byref(cv_has),
                                                     _ptr(f3, _f64))
        self._check(rc)
        bps = [bp[chrom_off[c]:chrom_off[c] + n_bp[c]].copy() for c in range(nc)]
        r = {"breakpoints": bps, "evenness": ev.value if ev_ok.value else None,
             "cv": cv.value if cv_has.value else None, "factor_of_three": f3}
        if sharded:
            r["owner"] = owner[:nc]
        return r

    # ------------------------------------------------------------------ Clean + Partition, fused
    def clean_partition_wavelet(self, chrom, is_autosome, is_chr_y, start, stop, count, gc,
                                size_filter=True, outlier_filter=True, gc_norm=True, gc_mode=0,
                                want_local_sd=True, min_bins_per_gc=100, is_germline=True,
                                mad_factor=5.0, thr_lower=0.05, thr_upper=80.0, min_size=10,
                                evenness_window=100000, out=None, chrom_selected=None, sharded=False):
        """cg_clean_partition_wavelet: both stages without leaving the device in between.
        `out` = (kept_index, count_out, n_bp, bp) preallocated (e.g. pinned) arrays; `chrom_selected` = 0/1 mask of
        the chromosomes this rank segments (multi-GPU)."""
        n = len(count)
        chrom = np.ascontiguousarray(chrom, np.uint8)
        is_autosome = np.ascontiguousarray(is_autosome, np.uint8)
        is_chr_y = np.ascontiguousarray(is_chr_y, np.uint8)
        start = np.ascontiguousarray(start, np.int32)
        stop = np.ascontiguousarray(stop, np.int32)
        count = np.ascontiguousarray(count, np.float32)
        gc = np.ascontiguousarray(gc, np.uint8)
        nc = len(is_autosome)
        co = CleanOpts(int(size_filter), int(outlier_filter), int(gc_norm), int(gc_mode),
                       int(want_local_sd), int(min_bins_per_gc))
        wo = WaveletOpts(int(is_germline), mad_factor, thr_lower, thr_upper, min_size, evenness_window)
        if out is None:
            kept = np.empty(max(n, 1), np.int32)
            cnt = np.empty(max(n, 1), np.float32)
            n_bp = np.zeros(max(nc, 1), np.int32)
            bp = np.zeros(max(n, 1), np.int32)
        else:
            kept, cnt, n_bp, bp = out
        n_out, lsd, skipped = _i64(0), _f64(0), C.c_int(0)
        off = np.zeros(nc + 1, np.int64)
        ev, cv = _f64(0), _f64(0)
        ev_ok, cv_has = C.c_int(0), C.c_int(0)
        f3 = np.zeros(9, np.float64)
        mask = None if chrom_selected is None else np.ascontiguousarray(chrom_selected, np.uint8)
        owner = np.zeros(max(nc, 1), np.int32)
        if sharded:
            rc = self.lib.cg_clean_partition_wavelet_sharded(
                self.h, C.byref(co), C.byref(wo), n, _ptr(chrom, _u8), _ptr(is_autosome, _u8),
                _ptr(is_chr_y, _u8), nc, _ptr(start, _i32), _ptr(stop, _i32), _ptr(count, _f32),
                _ptr(gc, _u8), C.byref(n_out), _ptr(kept, _i32), _ptr(cnt, _f32), C.byref(lsd),
                C.byref(skipped), _ptr(off, _i64), _ptr(n_bp, _i32), _ptr(bp, _i32), C.byref(ev),
                C.byref(ev_ok), C.byref(cv), C.byref(cv_has), _ptr(f3, _f64), _ptr(owner, _i32))
        else:
          rc = self.lib.cg_clean_partition_wavelet_shard(
            self.h, C.byref(co), C.byref(wo), n, _ptr(chrom, _u8), _ptr(is_autosome, _u8),
            _ptr(is_chr_y, _u8), nc, _ptr(start, _i32), _ptr(stop, _i32), _ptr(count, _f32),
            _ptr(gc, _u8), None if mask is None else _ptr(mask, _u8), C.byref(n_out), _ptr(kept, _i32), _ptr(cnt, _f32), C.byref(lsd),
            C.byref(skipped), _ptr(off, _i64), _ptr(n_bp, _i32), _ptr(bp, _i32), C.byref(ev),
            C.byref(ev_ok), C.byref(cv), C.byref(cv_has), _ptr(f3, _f64))
        self._check(rc)
        k = n_out.value
        bps = [bp[off[c]:off[c] + n_bp[c]].copy() for c in range(nc)]
        r = {"kept_index": kept[:k], "count": cnt[:k], "local_sd": lsd.value,
             "gc_norm_skipped": bool(skipped.value), "chrom_off": off, "breakpoints": bps,
             "evenness": ev.value if ev_ok.value else None,
             "cv": cv.value if cv_has.value else None, "factor_of_three": f3}
        if sharded:
            r["owner"] = owner[:nc]
        return r

    # ------------------------------------------------------------------ CBS segmentation
    def cbs_boundary(self, n_perm=10000, alpha=0.01, eta=0.05):
        key = (n_perm, alpha, eta)
        cache = self.__dict__.setdefault("_bdry", {})
        if key not in cache:
            k = self.lib.cg_cbs_boundary(n_perm, alpha, eta, None, 0)
            if k < 0:
                raise ValueError("cg_cbs_boundary: bad arguments")
            out = np.zeros(k, np.uint32)
            self.lib.cg_cbs_boundary(n_perm, alpha, eta, _ptr(out, C.c_uint32), k)
            cache[key] = out
        return cache[key]

    def partition_cbs(self, chrom_off, coverage, alpha=0.01, n_perm=10000, hybrid=True, min_width=2, k_max=25, n_min=200,
                      eta=0.05, undo=0, seed=0, sbdry=None, chrom_selected=None, trim=0.025, undo_sd=3.0, undo_prune=0.05, sharded=False):
        """CBSRunner.Run: per chromosome the segment lengths (bins) and means."""
        off = np.ascontiguousarray(chrom_off, np.int64)
        cov = np.ascontiguousarray(coverage, np.float64)
        nc = len(off) - 1
        if sbdry is None:
            sbdry = self.cbs_boundary(n_perm, alpha, eta)
        sbdry = np.ascontiguousarray(sbdry, np.uint32)
        o = CbsOpts(alpha, n_perm, int(hybrid), min_width, k_max, n_min, eta, trim, undo, undo_prune, undo_sd, seed)
        n = max(len(cov), 1)
        n_seg = np.zeros(max(nc, 1), np.int32)
        seg_len = np.zeros(n, np.int32)
        seg_mean = np.zeros(n, np.float64)
        stats = np.zeros(4, np.int64)
        owner = np.zeros(max(nc, 1), np.int32)
        if sharded:
            rc = self.lib.cg_partition_cbs_sharded(self.h, C.byref(o), _ptr(sbdry, C.c_uint32), len(sbdry), nc, _ptr(off, _i64),
                                                   _ptr(cov, C.c_double), _ptr(n_seg, _i32), _ptr(seg_len, _i32),
                                                   _ptr(seg_mean, C.c_double), _ptr(stats, _i64), _ptr(owner, _i32))
        elif chrom_selected is None:
            rc = self.lib.cg_partition_cbs(self.h, C.byref(o), _ptr(sbdry, C.c_uint32), len(sbdry), nc, _ptr(off, _i64),
                                           _ptr(cov, C.c_double), _ptr(n_seg, _i32), _ptr(seg_len, _i32),
                                           _ptr(seg_mean, C.c_double), _ptr(stats, _i64))
        else:
            mask = np.ascontiguousarray(chrom_selected, np.uint8)
            rc = self.lib.cg_partition_cbs_shard(self.h, C.byref(o), _ptr(sbdry, C.c_uint32), len(sbdry), nc, _ptr(off, _i64),
                                                 _ptr(cov, C.c_double), _ptr(mask, _u8), _ptr(n_seg, _i32), _ptr(seg_len, _i32),
                                                 _ptr(seg_mean, C.c_double), _ptr(stats, _i64))
        self._check(rc)
        segs = []
        for c in range(nc):
            a, k = int(off[c]), int(n_seg[c])
            segs.append({"len": seg_len[a:a + k].copy(), "mean": seg_mean[a:a + k].copy()})
        return {"segments": segs, "tests": int(stats[0]), "perms": int(stats[1]), "perm_steps": int(stats[2]),
                "edge_steps": int(stats[3]), "kernel_ms": self.lib.cg_last_kernel_ms(self.h), "phase_ms": self._cbs_phases()}

    def partition_hmm(self, chrom_off, coverage, per_sample=True, min_size=10, exact_sequential=False, chrom_selected=None, sharded=False):
        """HiddenMarkovModelsRunner.Run: breakpoints and Viterbi states.  coverage: [N] or [n_samples, N]."""
        off = np.ascontiguousarray(chrom_off, np.int64)
        cov = np.ascontiguousarray(np.atleast_2d(np.asarray(coverage, np.float64)))
        ns, n = cov.shape
        nc = len(off) - 1
        if nc > 0 and n != int(off[-1]):
            raise ValueError("coverage length does not match the chromosome offsets")
        o = HmmOpts(5, int(per_sample), min_size, int(exact_sequential))
        n_bp = np.zeros(max(nc, 1), np.int32)
        bp = np.zeros(max(n, 1), np.int32)
        states = np.zeros(max(n, 1), np.uint8)
        if sharded:
            owner = np.zeros(max(nc, 1), np.int32)
            rc = self.lib.cg_partition_hmm_sharded(self.h, C.byref(o), ns, nc, _ptr(off, _i64), _ptr(cov, _f64), _ptr(n_bp, _i32),
                                                   _ptr(bp, _i32), _ptr(states, _u8), _ptr(owner, _i32))
        elif chrom_selected is None:
            rc = self.lib.cg_partition_hmm(self.h, C.byref(o), ns, nc, _ptr(off, _i64), _ptr(cov, _f64), _ptr(n_bp, _i32),
                                           _ptr(bp, _i32), _ptr(states, _u8))
        else:
            mask = np.ascontiguousarray(chrom_selected, np.uint8)
            rc = self.lib.cg_partition_hmm_shard(self.h, C.byref(o), ns, nc, _ptr(off, _i64), _ptr(cov, _f64), _ptr(mask, _u8),
                                                 _ptr(n_bp, _i32), _ptr(bp, _i32), _ptr(states, _u8))
        self._check(rc)
        return {"breakpoints": [bp[off[c]:off[c] + n_bp[c]].copy() for c in range(nc)], "states": states[:n],
                "kernel_ms": self.lib.cg_last_kernel_ms(self.h), "launches": self.lib.cg_last_launches(self.h)}

    def partition_hmm_counts(self, chrom_off, count, text_mode=2, per_sample=True, min_size=10, chrom_selected=None):
        """cg_partition_hmm_counts: float counts in, the .cleaned text round trip (1 = F2, 2 = float.ToString(), 0 = none) on
        the device.  count: [N] or [n_samples, N] float32."""
        off = np.ascontiguousarray(chrom_off, np.int64)
        cnt = np.ascontiguousarray(np.atleast_2d(np.asarray(count, np.float32)))
        ns, n = cnt.shape
        nc = len(off) - 1
        if nc > 0 and n != int(off[-1]):
            raise ValueError("count length does not match the chromosome offsets")
        o = HmmOpts(5, int(per_sample), min_size, 0)
        n_bp = np.zeros(max(nc, 1), np.int32)
        bp = np.zeros(max(n, 1), np.int32)
        states = np.zeros(max(n, 1), np.uint8)
        mask = None if chrom_selected is None else np.ascontiguousarray(chrom_selected, np.uint8)
        rc = self.lib.cg_partition_hmm_counts(self.h, C.byref(o), ns, nc, _ptr(off, _i64), _ptr(cnt, _f32), int(text_mode),
                                              None if mask is None else _ptr(mask, _u8), _ptr(n_bp, _i32), _ptr(bp, _i32), _ptr(states, _u8))
        self._check(rc)
        return {"breakpoints": [bp[off[c]:off[c] + n_bp[c]].copy() for c in range(nc)], "states": states[:n],
                "kernel_ms": self.lib.cg_last_kernel_ms(self.h), "launches": self.lib.cg_last_launches(self.h)}

    def prefetch_bins(self, chrom, start, stop, count, gc):
        """cg_prefetch_bins: stage these columns on the device while the current call computes; the next clean /
        clean_partition_wavelet call given the SAME arrays (contiguous, right dtypes, ideally page-locked) consumes them."""
        for a, t in ((chrom, np.uint8), (start, np.int32), (stop, np.int32), (count, np.float32), (gc, np.uint8)):
            if not (isinstance(a, np.ndarray) and a.dtype == t and a.flags.c_contiguous):
                raise ValueError("prefetch_bins needs the contiguous typed arrays the later call will be given")
        self._check(self.lib.cg_prefetch_bins(self.h, len(count), _ptr(chrom, _u8), _ptr(start, _i32), _ptr(stop, _i32),
                                              _ptr(count, _f32), _ptr(gc, _u8)))

    def pedigree_hmm(self, chrom, is_autosome, is_chr_y, start, stop, counts, gc, sharded=False, min_size=10, out=None, want_tables=True,
                     size_filter=True, outlier_filter=True, gc_norm=True, gc_mode=0, want_local_sd=True, min_bins_per_gc=100):
        """cg_pedigree_hmm: CanvasClean per sample -> common bins -> PerSampleHMM per sample, device resident (one call).
        counts: one float32 array per sample over the shared layout (None allowed for samples another rank cleans when
        sharded).  out: optional (common_index i32[n], count f32[S, n], bp i32[S, n]) buffers, e.g. page-locked.
        want_tables=False: this rank does not need the merged table (common_index / count stay None, nothing is downloaded)."""
        chrom = np.ascontiguousarray(chrom, np.uint8)
        n = len(chrom)
        is_autosome = np.ascontiguousarray(is_autosome, np.uint8)
        is_chr_y = np.ascontiguousarray(is_chr_y, np.uint8)
        start = np.ascontiguousarray(start, np.int32)
        stop = np.ascontiguousarray(stop, np.int32)
        gc = np.ascontiguousarray(gc, np.uint8)
        cols = [None if c is None else np.ascontiguousarray(c, np.float32) for c in counts]
        S, nc = len(cols), len(is_autosome)
        for c in cols:
            if c is not None and len(c) != n:
                raise ValueError("every sample's counts must cover the shared bin layout")
        ptrs = (C.c_void_p * S)(*[None if c is None else c.ctypes.data for c in cols])
        co = CleanOpts(int(size_filter), int(outlier_filter), int(gc_norm), int(gc_mode), int(want_local_sd), int(min_bins_per_gc))
        ho = HmmOpts(5, 1, min_size, 0)
        if out is None:
            common = np.empty(max(n, 1), np.int32)
            cnt = np.empty((S, max(n, 1)), np.float32)
            bp = np.empty((S, max(n, 1)), np.int32)
        else:
            common, cnt, bp = out
        n_kept = np.zeros(S, np.int64)
        lsd = np.zeros(S, np.float64)
        skipped = np.zeros(S, np.int32)
        n_common = _i64(0)
        off = np.zeros(nc + 1, np.int64)
        n_bp = np.zeros((S, max(nc, 1)), np.int32)
        owner = np.zeros((S, max(nc, 1)), np.int32)
        rc = self.lib.cg_pedigree_hmm(self.h, C.byref(co), C.byref(ho), S, n, _ptr(chrom, _u8), _ptr(is_autosome, _u8), _ptr(is_chr_y, _u8),
                                      nc, _ptr(start, _i32), _ptr(stop, _i32), ptrs, _ptr(gc, _u8), int(bool(sharded)), _ptr(n_kept, _i64),
                                      _ptr(lsd, _f64), skipped.ctypes.data_as(_P(C.c_int)), C.byref(n_common),
                                      _ptr(common, _i32) if want_tables else None, _ptr(cnt, _f32) if want_tables else None,
                                      _ptr(off, _i64), _ptr(n_bp, _i32), _ptr(bp, _i32), _ptr(owner, _i32))
        self._check(rc)
        m = n_common.value
        st = self.last_partition_stats_raw()
        return {"breakpoints": [[bp[s, off[c]:off[c] + n_bp[s, c]].copy() for c in range(nc)] for s in range(S)],
                "chrom_off": off, "n_common": m, "common_index": common[:m] if want_tables else None,
                "count": cnt[:, :m] if want_tables else None, "n_kept": n_kept,
                "local_sd": lsd, "gc_norm_skipped": skipped.astype(bool), "owner": owner[:, :nc],
                "phases_ms": {"clean": st[0], "broadcast": st[1], "merge": st[2], "hmm": st[3], "gather": st[4], "download": st[8]},
                "kernel_ms": st[5], "launches": int(st[6]), "nccl_ms": st[7]}

    def last_partition_stats_raw(self):
        buf = (C.c_double * 16)()
        self.lib.cg_last_partition_stats(self.h, buf, 16)
        return list(buf)

    def merge_common_bins(self, samples):
        """MergeMultiSampleCleanedBedFile: samples = [(chrom_id u8, start i32, stop i32, count f32), ...] ordered by
        (chromosome id, start).  Returns kept_index (into sample 0), stop and counts [n_samples, n_common]."""
        ns = len(samples)
        cols = [[np.ascontiguousarray(a, t) for a, t in zip(smp, (np.uint8, np.int32, np.int32, np.float32))] for smp in samples]
        n = np.array([len(c[0]) for c in cols], np.int64)
        n0 = int(n[0]) if ns else 0
        arr = lambda k: (C.c_void_p * ns)(*[c[k].ctypes.data for c in cols])  # noqa: E731
        kept = np.zeros(max(n0, 1), np.int32)
        stop = np.zeros(max(n0, 1), np.int32)
        cnt = np.zeros((max(ns, 1), max(n0, 1)), np.float32)
        m = _i64(0)
        rc = self.lib.cg_merge_common_bins(self.h, ns, _ptr(n, _i64), arr(0), arr(1), arr(2), arr(3), C.byref(m), _ptr(kept, _i32),
                                           _ptr(stop, _i32), _ptr(cnt, _f32))
        self._check(rc)
        k = m.value
        return {"kept_index": kept[:k].copy(), "stop": stop[:k].copy(), "count": cnt[:ns, :k].copy(),
                "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def merge_kept_indices(self, n_bins, kept_lists, count_lists):
        """cg_merge_kept_indices: samples cleaned from one bin layout — kept_index / count of every sample in, the indices
        (into the layout) of the bins every sample kept and counts [n_samples, n_common] out."""
        ns = len(kept_lists)
        kept = [np.ascontiguousarray(k, np.int32) for k in kept_lists]
        cnt = [np.ascontiguousarray(c, np.float32) for c in count_lists]
        n = np.array([len(k) for k in kept], np.int64)
        n0 = max(int(n[0]), 1)
        kp = (C.c_void_p * ns)(*[k.ctypes.data for k in kept])
        cp = (C.c_void_p * ns)(*[c.ctypes.data for c in cnt])
        common = np.zeros(n0, np.int32)
        out = np.zeros((ns, n0), np.float32)
        n_out = _i64(0)
        self._check(self.lib.cg_merge_kept_indices(self.h, int(n_bins), ns, _ptr(n, _i64), kp, cp, C.byref(n_out), _ptr(common, _i32),
                                                   _ptr(out, _f32)))
        m = n_out.value
        return {"common_index": common[:m], "count": out[:, :m], "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def smooth(self, chrom_off, count, max_half_window):
        """RepeatedMedianSmoother.Smooth per chromosome: list of smoothed count arrays (possibly shorter than the input)."""
        off = np.ascontiguousarray(chrom_off, np.int64)
        cnt = np.ascontiguousarray(count, np.float32)
        nc = len(off) - 1
        n_out = np.zeros(max(nc, 1), np.int64)
        out = np.zeros(max(len(cnt), 1), np.float32)
        rc = self.lib.cg_smooth(self.h, int(max_half_window), nc, _ptr(off, _i64), _ptr(cnt, _f32), _ptr(n_out, _i64), _ptr(out, _f32))
        self._check(rc)
        return [out[off[c]:off[c] + n_out[c]].copy() for c in range(nc)]

    # ------------------------------------------------------------------ CanvasNormalize
    def normalize_reference(self, counts, on_target=None):
        """WeightedAverageReferenceGenerator.Run: counts [n_samples, n] (doubles) -> medians, weights, weighted bin counts."""
        c = np.ascontiguousarray(np.atleast_2d(np.asarray(counts, np.float64)))
        s, n = c.shape
        on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
        if on is not None and len(on) != n:
            raise ValueError("on_target must have one flag per bin")
        med, w, ref = np.zeros(s), np.zeros(s), np.zeros(max(n, 1))
        rc = self.lib.cg_normalize_reference(self.h, s, n, _ptr(c, _f64), _ptr(on, _u8) if on is not None else None,
                                             _ptr(med, _f64), _ptr(w, _f64), _ptr(ref, _f64))
        self._check(rc)
        return {"median": med, "weight": w, "reference": ref[:n], "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def normalize_best_lr2(self, sample, controls, on_target=None):
        """BestLR2ReferenceGenerator.Run: the control closest to the sample in mean squared log ratio."""
        t = np.ascontiguousarray(sample, np.float64)
        c = np.ascontiguousarray(np.atleast_2d(np.asarray(controls, np.float64)))
        s, n = c.shape
        if len(t) != n:
            raise ValueError("sample and controls must have the same number of bins")
        on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
        mean, ign, best = np.zeros(s), np.zeros(s, np.int64), C.c_int(-1)
        rc = self.lib.cg_normalize_best_lr2(self.h, s, n, _ptr(t, _f64), _ptr(c, _f64), _ptr(on, _u8) if on is not None else None,
                                            C.byref(best), _ptr(mean, _f64), _ptr(ign, _i64))
        self._check(rc)
        return {"best": int(best.value), "mean_sq_log_ratio": mean, "ignored": ign, "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def normalize_pca_reference(self, sample, mu, axes, on_target=None, min_ref=1.0, max_ref=float("inf")):
        """PCAReferenceGenerator.Run: reference counts (float) and the median sample / reference ratio."""
        a = np.ascontiguousarray(sample, np.float32)
        m = np.ascontiguousarray(mu, np.float32)
        ax = np.ascontiguousarray(np.atleast_2d(np.asarray(axes, np.float64)))
        k, n = ax.shape
        if len(a) != n or len(m) != n:
            raise ValueError("sample, mean and axes must have the same number of bins")
        on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
        ref = np.zeros(max(n, 1), np.float32)
        med = C.c_double(0)
        rc = self.lib.cg_normalize_pca_reference(self.h, n, k, _ptr(a, _f32), _ptr(m, _f32), _ptr(ax, _f64),
                                                 _ptr(on, _u8) if on is not None else None, min_ref, max_ref, _ptr(ref, _f32),
                                                 C.byref(med))
        self._check(rc)
        return {"reference": ref[:n], "median_ratio": med.value, "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def normalize_ratio(self, sample, reference, on_target=None, mode="lsnorm", min_ref=1.0, max_ref=float("inf"), ploidy=None):
        """LSNormRatioCalculator.Run ("lsnorm") or RawRatioCalculator.Run ("raw") + RatiosToCounts on the bins both lists
        share (the enumeration stops at the shorter one): kept bin indices, ratios, counts."""
        a = np.ascontiguousarray(sample, np.float32)
        b = np.ascontiguousarray(reference, np.float32)
        if len(a) != len(b) and mode == "lsnorm":
            # the reference takes each median over its whole file and only then enumerates up to the shorter one
            # (LSNormRatioCalculator.cs:28-36); medians over truncated lists would silently differ
            raise CanvasGpuError(CG_ERR_UNSUPPORTED, "normalize_ratio (lsnorm): sample and reference bin lists of different length")
        n = min(len(a), len(b))
        on = None if on_target is None else np.ascontiguousarray(on_target, np.uint8)
        pl = None if ploidy is None else np.ascontiguousarray(ploidy, np.int32)
        if (on is not None and len(on) < n) or (pl is not None and len(pl) < n):
            raise ValueError("on_target / ploidy must cover every bin")
        idx = np.zeros(max(n, 1), np.int32)
        ratio = np.zeros(max(n, 1), np.float32)
        count = np.zeros(max(n, 1), np.float32)
        k = C.c_int64(0)
        lsf = C.c_double(0)
        rc = self.lib.cg_normalize_ratio(self.h, n, _ptr(a, _f32), _ptr(b, _f32), _ptr(on, _u8) if on is not None else None,
                                         {"raw": 0, "lsnorm": 1}[mode], min_ref, max_ref,
                                         _ptr(pl, _i32) if pl is not None else None, C.byref(k), _ptr(idx, _i32),
                                         _ptr(ratio, _f32), _ptr(count, _f32), C.byref(lsf))
        self._check(rc)
        k = int(k.value)
        return {"kept_index": idx[:k].copy(), "ratio": ratio[:k].copy(), "count": count[:k].copy(),
                "library_size_factor": lsf.value, "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def _cbs_phases(self):
        """Phase times (ms) of the slowest chromosome of the last cg_partition_cbs call."""
        out = np.zeros(16, np.float64)
        self.lib.cg_last_partition_stats(self.h, _ptr(out, _f64), 16)
        names = ("passes", "observed", "tailp", "stream", "perms", "edge", "total")
        d = {k: round(float(out[i]), 3) for i, k in enumerate(names)}
        d["shuffle"] = round(float(out[7]), 3)
        d["chrom"] = int(out[8])
        return d

    # ------------------------------------------------------------------ CanvasBin counting
    def bin_screen(self, hits, possible, filter_start=(), filter_stop=()):
        """ExcludeTagsOverlappingFilterFile + ScreenObservedTags + the counts of GetRates on one chromosome.
        hits: uint8[len]; possible: bool[len].  Returns the screened copies and the two counts."""
        h = np.array(hits, np.uint8)
        n = len(h)
        pos = np.asarray(possible, bool)
        if len(pos) != n:
            raise ValueError("hits and possible must have one entry per position")
        words = np.packbits(np.concatenate([pos, np.zeros((-n) % 64, bool)]), bitorder="little").view(np.uint64).copy() if n else np.zeros(1, np.uint64)
        fs = np.ascontiguousarray(filter_start, np.int32)
        fe = np.ascontiguousarray(filter_stop, np.int32)
        obs, npos = C.c_int64(0), C.c_int64(0)
        rc = self.lib.cg_bin_screen(self.h, n, _ptr(h, _u8), _ptr(words, C.c_uint64), len(fs), _ptr(fs, _i32), _ptr(fe, _i32),
                                    C.byref(obs), C.byref(npos))
        self._check(rc)
        out_pos = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        return {"hits": h, "possible": out_pos, "observed": obs.value, "n_possible": npos.value,
                "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def bin_fragment_stats(self, frag_len):
        """Sum and number of the positive fragment lengths of one chromosome (Utilities.NonZeroMean)."""
        f = np.ascontiguousarray(frag_len, np.int16)
        s, c = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.cg_bin_fragment_stats(self.h, len(f), _ptr(f, C.c_int16), C.byref(s), C.byref(c)))
        return s.value, c.value

    def bin_read_gc(self, bases, frag_len, mean_frag, hits, expected=None, observed=None):
        """Read GC content per position and the chromosome's expected / observed read counts per GC bin, added to the
        int64[101] arrays passed in (new ones when omitted)."""
        b = bytes(bases)
        f = np.ascontiguousarray(frag_len, np.int16)
        h = np.ascontiguousarray(hits, np.uint8)
        n = len(b)
        if len(f) != n or len(h) != n:
            raise ValueError("bases, fragment lengths and hits must have one entry per position")
        gc = np.zeros(max(n, 1), np.uint8)
        exp = np.zeros(101, np.int64) if expected is None else expected
        obs = np.zeros(101, np.int64) if observed is None else observed
        rc = self.lib.cg_bin_read_gc(self.h, n, b, _ptr(f, C.c_int16), int(mean_frag), _ptr(h, _u8), _ptr(gc, _u8), _ptr(exp, _i64),
                                     _ptr(obs, _i64))
        self._check(rc)
        return {"read_gc": gc[:n], "expected": exp, "observed": obs, "kernel_ms": self.lib.cg_last_kernel_ms(self.h)}

    def bin_hits(self, hits, possible, bases, bin_size, mode=0, read_gc=None, obs_vs_exp_gc=None):
        """hits: uint8[len]; possible: bool[len]; bases: bytes of length len."""
        hits = np.ascontiguousarray(hits, np.uint8)
        n = len(hits)
        bits = np.packbits(np.asarray(possible, bool), bitorder="little")
        pad = (-len(bits)) % 8
        words = np.frombuffer(np.concatenate([bits, np.zeros(pad, np.uint8)]).tobytes(), dtype=np.uint64).copy()
        if len(words) == 0:
            words = np.zeros(1, np.uint64)
        cap = n // max(bin_size, 1) + 1
        start = np.zeros(cap, np.int32); stop = np.zeros(cap, np.int32); count = np.zeros(cap, np.int32)
        gc = np.zeros(cap, np.uint8)
        nb = _i64(0)
        rgc = np.ascontiguousarray(read_gc if read_gc is not None else np.zeros(1), np.uint8)
        ratio = np.ascontiguousarray(obs_vs_exp_gc if obs_vs_exp_gc is not None else np.ones(101), np.float32)
        rc = self.lib.cg_bin_hits(self.h, n, _ptr(hits, _u8), _ptr(words, C.c_uint64), bytes(bases), int(bin_size), int(mode),
                                  _ptr(rgc, _u8), _ptr(ratio, _f32), cap, C.byref(nb), _ptr(start, _i32), _ptr(stop, _i32),
                                  _ptr(count, _i32), _ptr(gc, _u8))
        self._check(rc)
        k = nb.value
        return {"start": start[:k], "stop": stop[:k], "count": count[:k], "gc": gc[:k]}

    def bin_fragments(self, frag_start, frag_stop, bin_start, bin_stop, undo_index=None):
        fs = np.ascontiguousarray(frag_start, np.int32); fe = np.ascontiguousarray(frag_stop, np.int32)
        bs = np.ascontiguousarray(bin_start, np.int32); be = np.ascontiguousarray(bin_stop, np.int32)
        undo = np.ascontiguousarray(undo_index if undo_index is not None else [], np.int32)
        best = np.full(max(len(fs), 1), -1, np.int32)
        count = np.zeros(max(len(bs), 1), np.int32)
        rc = self.lib.cg_bin_fragments(self.h, len(fs), _ptr(fs, _i32), _ptr(fe, _i32), len(undo), _ptr(undo, _i32), len(bs),
                                       _ptr(bs, _i32), _ptr(be, _i32), _ptr(best, _i32), _ptr(count, _i32))
        self._check(rc)
        return {"best_bin": best[:len(fs)], "count": count[:len(bs)]}

    # ------------------------------------------------------------------ stand-alone K8
    def normalize_apply(self, count, gc, median_by_gc, global_median, repeats=1):
        count = np.ascontiguousarray(count, np.float32)
        gc = np.ascontiguousarray(gc, np.uint8)
        batch, n = count.shape
        med = np.ascontiguousarray(median_by_gc, np.float64).reshape(batch, 101)
        gmed = np.ascontiguousarray(global_median, np.float64).reshape(batch)
        out = np.empty_like(count)
        ms = _f64(0)
        rc = self.lib.cg_normalize_apply(self.h, batch, n, _ptr(count, _f32), _ptr(gc, _u8),
                                         _ptr(med, _f64), _ptr(gmed, _f64), _ptr(out, _f32),
                                         repeats, C.byref(ms))
        self._check(rc)
        return out, ms.value
