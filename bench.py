#!/usr/bin/env python
"""Headline benchmark: Mbins/s through CanvasClean + CanvasPartition on synthetic WGS coverage arrays (BASELINE.json).

  python bench.py --gpus 1 --steps 5 --warmup 3              # config 2: one 3.1 M-bin germline sample on one B200
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W               # config 5: one independent sample per GPU (lists gathered once afterwards)
  python bench.py --config 3 | --config 4 [...]               # tumour/normal pair on one GPU | trio, (sample, chromosome) units over N GPUs
  python bench.py --impl reference [...]                      # the same workload on the host cores (C++ restatement of the reference)

One JSON line on stdout (rank 0).  `value` = bins of all ranks / device time of the kernels with the inputs already resident
in HBM (CUDA events on the library's launch stream, max over ranks); `e2e` = the same metric through the C-ABI call with
pinned HOST buffers (H2D + kernels + D2H, wall clock around the synchronous calls, max over ranks); `e2e_pipelined` = the
same steps with cg_prefetch_bins issued before every call (the next step's upload overlaps this step's kernels);
`roofline` = the Unbalanced-Haar decomposition against the measured HBM copy bandwidth; `cpu_baseline` = the oracle on this
box's cores.  At N > 1 the line also carries `per_rank` (stage times of every rank), `strong_scaling_single_sample` (ONE
sample, chromosomes LPT-sharded over the ranks inside cg_clean_partition_wavelet_sharded) and `config4` (the trio chain).
Every exchange goes through the library's own NCCL communicator (cg_comm_*); torch.distributed only carries the 128-byte
NCCL id, the barriers and the max-over-ranks of the timings.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before torch creates the CUDA context (see canvas_b200/native.py)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mbins/s through Clean+Partition on 3M-bin WGS array"
UNIT = "Mbins/s"
WORKLOADS = {
    2: ("config2 germline-WGS 30x synthetic, ~3.1M x 1kb bins: CanvasClean (-g -s -r, local-SD metric, MedianByGC) + "
        "CanvasPartition wavelets (-g); one sample per GPU"),
    3: ("config3 somatic-WGS tumour/normal pair, 2 x ~3.1M bins on one GPU: CanvasClean + CanvasPartition wavelets "
        "(somatic thresholds) per sample"),
    4: ("config4 SmallPedigree-WGS trio, 3 x ~3.1M bins, one device-resident call: CanvasClean per sample (once, on rank s mod N) -> NCCL broadcast -> "
        "common-bin merge -> PerSampleHMM over (sample, chromosome) units LPT-sharded across the GPUs -> NCCL gather"),
    5: ("config5 batch of independent 30x WGS samples (config-2 pipeline), one per GPU; the per-sample segment lists are "
        "gathered once after the timed steps"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------- CPU arm
def oracle_clean(s):
    from oracle import pyoracle as ora
    return ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)


def oracle_partition_input(s, cleaned):
    """The .cleaned text round trip between the two modules (Python glue here, file I/O in the reference: timed on neither arm)."""
    from canvas_b200 import synth
    from oracle import pyoracle as ora
    return synth.chrom_offsets(s.chrom[cleaned["kept_index"]], len(s.names)), ora.f2_roundtrip(cleaned["count"])


def oracle_samples(samples, threads, germline=True):
    """The reference path of a batch of samples on the CPU, side by side: CanvasClean (single thread per sample, as the
    reference) -> [untimed: .cleaned text round trip] -> CanvasPartition wavelets with one thread per chromosome.  Returns the
    wall seconds of the two timed phases."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as ora
    workers = min(len(samples), threads)
    per = max(1, threads // workers)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        t0 = time.perf_counter()
        cl = list(ex.map(oracle_clean, samples))
        t1 = time.perf_counter()
        mid = [oracle_partition_input(s, c) for s, c in zip(samples, cl)]
        t2 = time.perf_counter()
        ps = list(ex.map(lambda oc: ora.partition_wavelet(oc[0], oc[1], is_germline=germline, n_threads=per), mid))
        t3 = time.perf_counter()
    return {"clean_s": t1 - t0, "partition_s": t3 - t2, "timed_s": (t1 - t0) + (t3 - t2),
            "breakpoints": sum(len(b) for p in ps for b in p["breakpoints"])}


def oracle_trio(samples, threads):
    """Config 4 on the CPU: Clean per sample (one thread each, side by side as separate CanvasClean processes would run),
    dictionary merge of the common bins, PerSampleHMM per sample with one thread per chromosome."""
    from concurrent.futures import ThreadPoolExecutor
    from canvas_b200 import pedigree, synth, textcodec
    from oracle import pyoracle as ora
    pos = pedigree.bin_positions(samples[0])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(len(samples), threads)) as ex:
        cl = list(ex.map(lambda s: ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc), samples))
    t1 = time.perf_counter()
    cleaned = [(s.chrom[c["kept_index"]], pos[c["kept_index"]], (pos[c["kept_index"]] + 1000).astype(np.int32), c["count"])
               for s, c in zip(samples, cl)]
    m = ora.merge_common_bins_np(cleaned)
    ch0 = cleaned[0][0][m["kept_index"]]
    off = synth.chrom_offsets(ch0, len(samples[0].names))
    t2 = time.perf_counter()
    covs = [textcodec.float_default_roundtrip(m["count"][k]) for k in range(len(samples))]  # merged-file text round trip: untimed glue
    t2b = time.perf_counter()
    nbp = 0
    for cov in covs:
        nbp += sum(len(b) for b in ora.partition_hmm(off, cov, per_sample=True, n_threads=threads)["breakpoints"])
    t3 = time.perf_counter()
    return {"clean_s": t1 - t0, "merge_s": t2 - t1, "partition_s": t3 - t2b, "timed_s": (t2 - t0) + (t3 - t2b), "breakpoints": nbp}


def run_reference(args, config, n_samples, K, W):
    """`--impl reference`: the C++ restatement of the reference on every host core, same workload as the GPU arm."""
    from concurrent.futures import ThreadPoolExecutor
    from canvas_b200 import synth
    threads = os.cpu_count() or 1
    if config == 4:
        samples = [synth.make_sample(config=4, sample=k, scale=args.scale, n_events=60) for k in range(3)]
        step = lambda: oracle_trio(samples, threads)  # noqa: E731
        how = ("the trio per step: Clean of the three samples side by side (one thread each), dictionary merge, "
               "PerSampleHMM with one thread per chromosome")
    else:
        if config == 3:
            samples = [synth.make_sample(config=3, sample=k, scale=args.scale, n_events=150, tumour=(k == 0)) for k in range(2)]
            how = "tumour and normal per step, side by side: Clean 1 thread per sample, Partition 1 thread per chromosome"
        else:
            samples = [synth.make_sample(config=2, sample=k, scale=args.scale) for k in range(n_samples)]
            how = (f"{n_samples} full config-2 sample(s) per step, side by side (Clean on 1 thread per sample as the reference, "
                   "Partition one thread per chromosome)")
        step = lambda: oracle_samples(samples, threads, germline=(config != 3))  # noqa: E731
    nb = sum(len(s) for s in samples)
    for _ in range(W):
        step()
    parts = [step() for _ in range(K)]
    sec = sum(x["timed_s"] for x in parts) / K  # the modules' numeric work; the text round trips between them are not timed (nor on the GPU arm)
    v = nb / sec / 1e6
    return {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if config == 4 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[config], "bins_per_sample": len(samples[0]), "samples": len(samples)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": how + "; C++ restatement of the reference (the C# build needs private NuGet feeds)"},
            "clean_ms": 1e3 * sum(x["clean_s"] for x in parts) / K,
            "partition_ms": 1e3 * sum(x["partition_s"] for x in parts) / K,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


_REAL_STDOUT = None


def _claim_stdout():
    """Everything a library prints on fd 1 (NCCL's version banner, for one) goes to stderr from here on; the JSON
    line is written to the original stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def pack_segments(breakpoints):
    """(chromosome, breakpoint) pairs of one sample as a flat int32 list."""
    if not any(len(b) for b in breakpoints):
        return np.zeros(0, np.int32)
    return np.concatenate([np.stack([np.full(len(b), c, np.int32), b.astype(np.int32)], 1).ravel()
                           for c, b in enumerate(breakpoints) if len(b)])


def effective_config(args, world):
    return args.config or (2 if max(world, args.gpus) == 1 else 5)


# ---------------------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="canvas_b200", choices=["canvas_b200", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json configuration (default: 2 on one GPU, 5 = one config-2 sample per GPU on several)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the genome (debugging only; 1.0 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the strong-scaling and config-4 side measurements")
    ap.add_argument("--method", default="wavelets", choices=["wavelets", "cbs"],
                    help="segmentation method of the Partition stage (cbs: one GPU, config 2; Clean and Partition as two calls)")
    args = ap.parse_args()
    W = max(args.warmup, 0)
    K = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = effective_config(args, world)
    from canvas_b200 import synth
    _claim_stdout()

    if args.impl == "reference":
        if rank != 0:
            return 0
        n_samples = max(world, args.gpus) if config in (2, 5) else 1
        _emit(run_reference(args, config, n_samples, K, W))
        return 0

    import torch
    import torch.distributed as dist
    from canvas_b200 import native, pedigree
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not mix with the JSON line on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    eng = native.Engine(local_rank, bind_numa=True)  # CPU cores and page-locked buffers on the GPU's own socket
    if world > 1:
        eng.comm_init_torch()  # the library's own NCCL communicator; torch only hands the id around
    else:
        eng.comm_init(1, 0)
    pin = eng.pinned
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    W = max(W, 3)  # timing rules: at least three warm-up steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(v):
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    def pinned_sample(s):
        return dict(chrom=pin.array(s.chrom), start=pin.array(s.start), stop=pin.array(s.stop), count=pin.array(s.count),
                    gc=pin.array(s.gc))

    def timed(step, n_steps, step_barrier=True):
        """n_steps timed steps between barriers; L2 flushed before every step; per-step wall seconds (the flush is outside).
        step_barrier: the ranks also start every step together (steps that contain a collective)."""
        walls = []
        barrier()
        for _ in range(n_steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            if world > 1 and step_barrier:
                dist.barrier()
            t = time.perf_counter()
            step()
            walls.append(time.perf_counter() - t)
        barrier()
        return walls

    line_extra = {}

    # ------------------------------------------------------------------ config 4: the trio chain
    def run_config4(n_warm, n_steps):
        trio = [synth.make_sample(config=4, sample=k, scale=args.scale, n_events=60) for k in range(3)]
        t0 = trio[0]
        n4, S4 = len(t0), len(trio)
        # page-locked layout columns and counts in, page-locked result tables out, as in the other configurations; a rank
        # only holds the counts of the samples it cleans (sample s on rank s mod N)
        lay = dict(chrom=pin.array(t0.chrom), start=pin.array(t0.start), stop=pin.array(t0.stop), gc=pin.array(t0.gc))
        cols = [pin.array(t.count) if k % world == rank else None for k, t in enumerate(trio)]
        out = (pin.empty(n4, np.int32), pin.empty(S4 * n4, np.float32).reshape(S4, n4), pin.empty(S4 * n4, np.int32).reshape(S4, n4))
        tm, res = {}, {}

        def step():
            # rank 0 writes the merged .cleaned files: it alone downloads the merged table; every rank gets all breakpoints
            r = eng.pedigree_hmm(lay["chrom"], t0.is_autosome, t0.is_chr_y, lay["start"], lay["stop"], cols, lay["gc"],
                                 sharded=world > 1, out=out, want_tables=rank == 0)
            res["r"] = r
            if res.get("timed"):
                for k, v in r["phases_ms"].items():
                    tm[k] = tm.get(k, 0.0) + v
                for k in ("kernel_ms", "launches", "nccl_ms"):
                    tm[k] = tm.get(k, 0.0) + r[k]
        for _ in range(n_warm):
            step()
        res["timed"] = True
        walls = timed(step, n_steps)
        wall_ms, kern_ms = max_over_ranks([1e3 * sum(walls) / n_steps, tm.get("kernel_ms", 0.0) / n_steps])
        bins = S4 * n4
        r = res["r"]
        mine = sum(1 for k in range(S4) if k % world == rank)
        return {"workload": WORKLOADS[4], "bins": bins, "common_bins": r["n_common"],
                "breakpoints": int(sum(len(b) for per in r["breakpoints"] for b in per)),
                "ms_per_step": wall_ms, "Mbins_per_s": bins / wall_ms / 1e3,
                "kernel_ms_max_rank": kern_ms, "Mbins_per_s_kernels": bins / kern_ms / 1e3 if kern_ms > 0 else None,
                "phases_ms_rank0": {k: v / n_steps for k, v in tm.items() if k not in ("kernel_ms", "launches", "nccl_ms")},
                "nccl_ms_rank0": tm.get("nccl_ms", 0.0) / n_steps,
                "launches_rank0": int(tm.get("launches", 0)) // max(n_steps, 1),
                "units_per_rank": np.bincount(r["owner"].ravel(), minlength=world).tolist(),
                "h2d_bytes": (10 * n4 + 4 * n4 * mine) if mine else 0,
                "d2h_bytes": (4 * r["n_common"] * (S4 + 1) if rank == 0 else 0) + 4 * int(sum(len(b) for per in r["breakpoints"] for b in per)),
                "call": "cg_pedigree_hmm (one device-resident call: Clean per sample, common bins, PerSampleHMM per sample)",
                "timing": "wall clock around the C-ABI call, page-locked host buffers in and out, max over ranks"}

    # ------------------------------------------------------------------ CanvasPartition -m CBS on the config-2 sample
    if args.method == "cbs":
        s = synth.make_sample(config=2, sample=0, scale=args.scale)
        inp = pinned_sample(s)
        acc = {"clean": 0.0, "cbs": 0.0, "launches": 0}
        res = {}

        def step():
            c = eng.clean(inp["chrom"], s.is_autosome, s.is_chr_y, inp["start"], inp["stop"], inp["count"], inp["gc"])
            acc["clean"] += eng.last_kernel_ms
            acc["launches"] += eng.last_launches
            off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
            from canvas_b200 import textcodec
            r = eng.partition_cbs(off, textcodec.f2_roundtrip(c["count"]))
            acc["cbs"] += r["kernel_ms"]
            acc["launches"] += 1
            res["r"], res["bins"] = r, len(c["kept_index"])
        eng.cbs_boundary()
        for _ in range(W):
            step()
        acc.update(clean=0.0, cbs=0.0, launches=0)
        sampler = ClockSampler(local_rank)
        sampler.start()
        walls = timed(step, K)
        clocks = sampler.summary()
        r = res["r"]
        kern_ms = (acc["clean"] + acc["cbs"]) / K
        wall_ms = 1e3 * sum(walls) / K
        hbm, peak_src = peaks()
        # a test on n bins reads the partial sums once and shuffles + scans them once per executed permutation
        alg = 16.0 * (r["perm_steps"] + r["edge_steps"] + res["bins"])
        line = {"metric": METRIC, "value": len(s) / kern_ms / 1e3, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": kern_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOADS[2].replace("wavelets (-g)", "-m CBS (hybrid p-values, 10000 permutations)"), "bins_per_sample": len(s),
                           "samples": 1, "l2": "flushed between steps (256 MiB device write)", "method": "cbs"},
                "e2e": {"value": len(s) / wall_ms / 1e3, "unit": UNIT, "ms_per_step": wall_ms, "h2d_bytes_per_step": 14 * len(s) + 8 * res["bins"],
                        "d2h_bytes_per_step": 8 * res["bins"] + 12 * sum(len(x["len"]) for x in r["segments"]),
                        "timing": "wall clock around cg_clean + the host's .cleaned round trip + cg_partition_cbs, pinned inputs"},
                "gpu_launches": acc["launches"], "stages_ms": {"clean": acc["clean"] / K, "cbs": acc["cbs"] / K},
                "cbs": {"tests": r["tests"], "permutations": r["perms"], "permuted_bins": r["perm_steps"], "edge_draws": r["edge_steps"],
                        "segments": int(sum(len(x["len"]) for x in r["segments"])), "phase_ms_slowest_chromosome": r["phase_ms"]},
                "roofline": {"kernel": "cbs_kernel (one cluster per chromosome)", "bound": "hbm", "achieved": alg / (acc["cbs"] / K * 1e-3) / 1e9,
                             "peak": hbm, "unit": "GB/s", "frac": alg / (acc["cbs"] / K * 1e-3) / 1e9 / hbm, "traffic": None,
                             "algorithmic_bytes": alg, "peak_source": peak_src,
                             "note": "16 B x (permuted bins + edge draws + bins): dependent Fisher-Yates chains and the MT19937 stream bound it, not HBM"},
                "clocks": clocks, "device": eng.describe()}
        if not args.no_cpu_baseline:
            from oracle import pyoracle as ora
            threads = os.cpu_count() or 1
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every core of the box
            c = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
            off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
            t0 = time.perf_counter()
            ora.partition_cbs(off, ora.f2_roundtrip(c["count"]), n_threads=threads)
            sec = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": res["bins"] / sec / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "CBS of the full cleaned config-2 sample, one thread per chromosome", "partition_ms": sec * 1e3}
        _emit(line)
        return 0

    if config == 4:
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        c4 = run_config4(W, K)
        clocks = sampler.summary() if rank == 0 else None
        if rank == 0:
            line = {"metric": METRIC, "value": c4["Mbins_per_s_kernels"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                    "ms_per_step": c4["kernel_ms_max_rank"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic",
                    "config": {"workload": WORKLOADS[4], "bins_per_sample": c4["bins"] // 3, "samples": 3,
                               "l2": "flushed between steps (256 MiB device write)", "parallelism": f"(sample, chromosome) units x{world}"},
                    "e2e": {"value": c4["Mbins_per_s"], "unit": UNIT, "ms_per_step": c4["ms_per_step"],
                            "h2d_bytes_per_step": c4["h2d_bytes"], "d2h_bytes_per_step": c4["d2h_bytes"], "timing": c4["timing"]},
                    "gpu_launches": c4["launches_rank0"] * K, "config4": c4, "clocks": clocks, "device": eng.describe()}
            _emit(line)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ configs 2 / 5 (one germline sample per GPU) and 3 (pair)
    if config == 3:
        samples = [synth.make_sample(config=3, sample=k, scale=args.scale, n_events=150, tumour=(k == 0)) for k in range(2)]
        germline = False
    else:
        samples = [synth.make_sample(config=2, sample=rank, scale=args.scale)]
        germline = True
    nb = sum(len(s) for s in samples)
    nc = len(samples[0].names)
    inps = [pinned_sample(s) for s in samples]
    outs = [(pin.empty(len(s), np.int32), pin.empty(len(s), np.float32), pin.empty(max(nc, 1), np.int32), pin.empty(len(s), np.int32))
            for s in samples]
    h2d = nb * (1 + 4 + 4 + 4 + 1)
    acc = {"dev_ms": [], "launches": 0, "stages": {}, "pstats": None, "d2h": 0, "x_ms": 0.0, "last": None}

    def step():
        dev_ms, d2h = 0.0, 0
        r = None
        for s, inp, out in zip(samples, inps, outs):
            r = eng.clean_partition_wavelet(inp["chrom"], s.is_autosome, s.is_chr_y, inp["start"], inp["stop"], inp["count"],
                                            inp["gc"], is_germline=germline, out=out)
            dev_ms += eng.last_kernel_ms
            acc["launches"] += eng.last_launches
            for k, v in eng.last_stage_ms().items():
                acc["stages"][k] = acc["stages"].get(k, 0.0) + v
            acc["pstats"] = eng.last_partition_stats()
            nbp = sum(len(b) for b in r["breakpoints"])
            d2h += len(r["kept_index"]) * 8 + nbp * 4 + nc * 4 + 4096
            acc["last"] = r
        acc["dev_ms"].append(dev_ms)
        acc["d2h"] = d2h

    for _ in range(W):
        step()
    acc.update(dev_ms=[], launches=0, stages={}, x_ms=0.0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # independent samples: no collective and no barrier inside the timed region (the samples of a cohort share nothing)
    walls = timed(step, K, step_barrier=False)
    clocks = sampler.summary() if rank == 0 else None
    my_wall, my_kern = 1e3 * sum(walls) / K, sum(acc["dev_ms"]) / K
    # the same K steps software-pipelined across calls (cg_prefetch_bins): the upload of the next step's sample is started
    # before the current step's call and crosses PCIe while that call's kernels run; every step still uploads one sample
    # from page-locked memory and downloads its own results
    def step_pipelined():
        for s, inp, out in zip(samples, inps, outs):
            eng.prefetch_bins(inp["chrom"], inp["start"], inp["stop"], inp["count"], inp["gc"])
            eng.clean_partition_wavelet(inp["chrom"], s.is_autosome, s.is_chr_y, inp["start"], inp["stop"], inp["count"],
                                        inp["gc"], is_germline=germline, out=out)
    for s, inp in zip(samples, inps):  # what the (untimed) previous step would have started
        eng.prefetch_bins(inp["chrom"], inp["start"], inp["stop"], inp["count"], inp["gc"])
    for _ in range(2):
        step_pipelined()
    walls_p = timed(step_pipelined, K, step_barrier=False)
    (wall_p_ms,) = max_over_ranks([1e3 * sum(walls_p) / K])
    if world > 1:
        # after the timed region: every rank's segment lists on every rank through the library's communicator (cohort report)
        acc["all"] = eng.allgather_lists(pack_segments(acc["last"]["breakpoints"]))
        acc["x_ms"] = eng.last_exchange_ms * K
    wall_ms, kern_ms = max_over_ranks([my_wall, my_kern])
    total_bins = sum_over_ranks(nb)
    r = acc["last"]
    pstats = acc["pstats"]
    stages = {k: v / K for k, v in acc["stages"].items()}
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "device_ms": my_kern, "e2e_ms": my_wall, "stages_ms": stages, "exchange_ms": acc["x_ms"] / K,
                "l_eff": pstats["visits"] / max(1, len(r["kept_index"])), "launches_per_step": acc["launches"] // K, "host_cpus": len(eng.host_cpus) if eng.host_cpus else None,
                "breakpoints": int(sum(len(b) for b in r["breakpoints"]))}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    strong = None
    if world > 1 and not args.no_extras:
        # ONE sample with its chromosomes sharded over the ranks, entirely inside the C-ABI: Clean and the
        # genome-wide scalars on every rank, each rank segments its LPT share, one NCCL all-gather of breakpoints
        s0 = samples[0] if rank == 0 else synth.make_sample(config=2, sample=0, scale=args.scale)
        in0 = inps[0] if rank == 0 else pinned_sample(s0)
        res = {"k": 0.0}

        def sstep():
            res["r"] = eng.clean_partition_wavelet(in0["chrom"], s0.is_autosome, s0.is_chr_y, in0["start"], in0["stop"], in0["count"],
                                                   in0["gc"], is_germline=True, out=outs[0], sharded=True)
            res["k"] += eng.last_kernel_ms
        for _ in range(W):
            sstep()
        res["k"] = 0.0
        sw = timed(sstep, K)
        s_wall, s_kern = max_over_ranks([1e3 * sum(sw) / K, res["k"] / K])
        strong = {"what": "ONE config-2 sample through cg_clean_partition_wavelet_sharded: Clean + scalars on every rank, chromosomes "
                          "LPT-sharded, one NCCL all-gather of breakpoints inside the call; wall clock incl. H2D/D2H, max over ranks",
                  "ms_per_sample": s_wall, "Mbins_per_s": len(s0) / s_wall / 1e3, "device_ms": s_kern,
                  "breakpoints": int(sum(len(b) for b in res["r"]["breakpoints"])), "exchange_ms": eng.last_exchange_ms}
        line_extra["config4"] = run_config4(2, 3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm, peak_src = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full captures
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    visits = pstats["visits"]
    # the chromosomes' pipelines (decomposition stages + finish) overlap: the decomposition's own span is the device-side
    # timestamp from its first kernel to its last tiny-stage thread (partition_stats), not a stage bracket
    dec_ms = pstats.get("decompose_span_ms", 0.0) or stages.get("decompose", -1.0)
    alg_bytes = 8.0 * visits  # one f64 prefix sum read per bin visit (8 * L_eff B/bin)
    achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else None
    line = {"metric": METRIC, "value": total_bins / (kern_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[config], "bins_per_sample": len(samples[0]), "samples": world * len(samples),
                       "l2": "flushed between steps (256 MiB device write)", "parallelism": f"sample-per-gpu x{world}",
                       "exchange": "none in the timed region (independent samples); the per-sample segment lists are gathered once afterwards with cg_comm_allgather_lists" if world > 1 else "none"},
            "e2e": {"value": total_bins / (wall_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wall_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": acc["d2h"],
                    "timing": "wall clock around the synchronous C-ABI call(s) with pinned host buffers, max over ranks"},
            "e2e_pipelined": {"value": total_bins / (wall_p_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wall_p_ms,
                              "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": acc["d2h"],
                              "timing": "as e2e, with cg_prefetch_bins issued before every call: the upload of the next step's sample "
                                        "overlaps the current call's kernels (two device staging slots); K uploads and K downloads in K steps"},
            "gpu_launches": acc["launches"],
            "stages_ms": stages, "partition_stats": pstats,
            "roofline": {"kernel": "Unbalanced-Haar decomposition (uh_chain + uh_mid + uh_small + uh_tiny kernels, one pipeline per chromosome)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": (achieved / hbm) if achieved else None,
                         "traffic": traffic.get("uh_decompose"), "traffic_source": traffic.get("source"),
                         "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                         "l_eff": visits / max(1, len(r["kept_index"])), "ms": dec_ms,
                         "note": "the prefix sums (24 MB) stay in L2: DRAM traffic (cold-cache ncu) is ~1/%d of the algorithmic bytes; the stage "
                                 "is bound by dependent chains of tree nodes, not by HBM" % max(1, round(alg_bytes / max(1, traffic.get("uh_decompose", 1))))},
            "clocks": clocks, "device": eng.describe()}
    # Clean / Partition totals against the HBM roofline with SURVEY.md 8(d)'s accounting: Clean = 101 B per input bin,
    # wavelet Partition = 8 * L_eff + 72 B per cleaned bin
    kept = max(1, len(r["kept_index"]))
    l_eff = visits / kept
    n_s = len(samples)
    part_ms = sum(stages.get(k, 0.0) for k in ("scalars", "decompose", "finish"))
    totals = {}
    for name, nbytes, ms in (("clean", 101.0 * nb, stages.get("clean", 0.0)), ("partition_wavelet", (8.0 * l_eff + 72.0) * kept * n_s, part_ms)):
        if ms > 0:
            ach = nbytes / (ms * 1e-3) / 1e9
            totals[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved": ach, "unit": "GB/s", "frac": ach / hbm}
    line["roofline_totals"] = totals
    if per_rank is not None:
        line["per_rank"] = per_rank
    if strong is not None:
        line["strong_scaling_single_sample"] = strong
    line.update(line_extra)
    # K8 normalise stream on batches larger than L2 (the kernel BASELINE.json's roofline target names):
    # 8 samples = config 5 (223 MB), 16 samples (446 MB) and 32 samples (892 MB)
    try:
        s = samples[0]
        n16 = (len(s) // 16) * 16
        rng = np.random.default_rng(1)
        for batch, key in ((8, "roofline_normalize"), (16, "roofline_normalize_16"), (32, "roofline_normalize_32")):
            cnt = np.tile(s.count[:n16], (batch, 1))
            gcb = np.tile(s.gc[:n16], (batch, 1))
            med = rng.uniform(80, 120, (batch, 101))
            _, k8_ms = eng.normalize_apply(cnt, gcb, med, np.full(batch, 100.0), repeats=20)
            k8_bytes = 9.0 * batch * n16
            line[key] = {"kernel": "normalize_apply_bulk_kernel", "bound": "hbm",
                         "achieved": k8_bytes / (k8_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": k8_bytes / (k8_ms * 1e-3) / 1e9 / hbm,
                         "traffic": traffic.get("normalize_apply_bulk_kernel") if batch == 8 else None,
                         "algorithmic_bytes": k8_bytes, "batch_samples": batch, "batch_bins": batch * n16, "ms": k8_ms,
                         "timing": "CUDA events on the library stream around 20 back-to-back launches after 1 warm-up"}
    except Exception as e:  # noqa
        line["roofline_normalize"] = {"error": str(e)}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every core of the box, not only the GPU's socket
        o = oracle_samples(samples, threads, germline=germline)
        line["cpu_baseline"] = {"value": nb / o["timed_s"] / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{len(samples)} full sample(s) of this workload (Clean 1 thread per sample, Partition 1 thread per chromosome; "
                                          "the text round trip between the modules is not timed)",
                                "clean_ms": o["clean_s"] * 1e3, "partition_ms": o["partition_s"] * 1e3}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
