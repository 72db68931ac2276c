#!/usr/bin/env python
"""Headline benchmark: Mbins/s through CanvasClean + CanvasPartition (wavelets) on a synthetic
3.1 M-bin germline WGS coverage array (BASELINE.json config 2), 1..8 GPUs, one sample per GPU.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # CPU restatement of the reference on the host cores

One JSON line on stdout (rank 0).  `value` = bins of all ranks / device time of the kernels with the
inputs already resident in HBM (CUDA events on the library's launch stream); `e2e` = the same metric
through the C-ABI call with pinned HOST buffers (H2D + kernels + D2H, wall clock around the
synchronous call, max over ranks); `roofline` = Unbalanced-Haar decomposition kernel against the
measured HBM copy bandwidth; `cpu_baseline` = oracle (C++ restatement of the reference) on this box.
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
WORKLOAD = ("config2 germline-WGS 30x synthetic, ~3.1M x 1kb bins: CanvasClean (-g -s -r, local-SD metric, "
            "MedianByGC) + CanvasPartition wavelets (-g); one sample per GPU")


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


def run_oracle_once(s, threads):
    """The reference path on the CPU: CanvasClean (single thread, as the reference) -> .cleaned text
    round trip -> CanvasPartition wavelets with one thread per chromosome up to `threads`."""
    from canvas_b200 import synth
    from oracle import pyoracle as ora
    t0 = time.perf_counter()
    r = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    t1 = time.perf_counter()
    off = synth.chrom_offsets(s.chrom[r["kept_index"]], len(s.names))
    cov = ora.f2_roundtrip(r["count"])
    t2 = time.perf_counter()
    p = ora.partition_wavelet(off, cov, is_germline=True, n_threads=threads)
    t3 = time.perf_counter()
    return {"clean_s": t1 - t0, "partition_s": t3 - t2, "total_s": (t1 - t0) + (t3 - t2),
            "breakpoints": sum(len(b) for b in p["breakpoints"])}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="canvas_b200", choices=["canvas_b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the genome (debugging only; 1.0 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    W = max(args.warmup, 0)
    K = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from canvas_b200 import synth
    _claim_stdout()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        s = synth.make_sample(config=2, sample=0, scale=args.scale)
        nb = len(s)
        for _ in range(max(W, 0)):
            run_oracle_once(s, threads)
        t = []
        for _ in range(K):
            t.append(run_oracle_once(s, threads))
        sec = sum(x["total_s"] for x in t) / K
        v = nb / sec / 1e6
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD, "bins_per_sample": nb, "samples": 1},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": "one full config-2 sample per step (Clean on 1 thread as the reference, "
                                           "Partition one thread per chromosome); C++ restatement, the C# build "
                                           "needs private NuGet feeds"},
                "clean_ms": 1e3 * sum(x["clean_s"] for x in t) / K,
                "partition_ms": 1e3 * sum(x["partition_s"] for x in t) / K,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    from canvas_b200 import native
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not mix with the JSON line on stdout
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = native.Engine(local_rank)
    s = synth.make_sample(config=2, sample=rank, scale=args.scale)
    nb = len(s)
    nc = len(s.names)
    pin = eng.pinned
    inp = dict(chrom=pin.array(s.chrom), start=pin.array(s.start), stop=pin.array(s.stop), count=pin.array(s.count),
               gc=pin.array(s.gc))
    out = (pin.empty(nb, np.int32), pin.empty(nb, np.float32), pin.empty(max(nc, 1), np.int32), pin.empty(nb, np.int32))
    h2d = nb * (1 + 4 + 4 + 4 + 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    SEG_CAP = 8192
    seg_local = torch.zeros(SEG_CAP, dtype=torch.int32, device=dev)
    seg_all = torch.zeros(SEG_CAP * world, dtype=torch.int32, device=dev) if world > 1 else None

    e2e_s = []

    def step():
        flush.fill_(1)
        torch.cuda.synchronize()
        t_call = time.perf_counter()
        r = eng.clean_partition_wavelet(inp["chrom"], s.is_autosome, s.is_chr_y, inp["start"], inp["stop"],
                                        inp["count"], inp["gc"], is_germline=True, out=out)
        nbp = sum(len(b) for b in r["breakpoints"])
        if world > 1:
            # config 5: gather the per-sample segment lists (chromosome, breakpoint) on every rank
            flat = np.zeros(SEG_CAP, np.int32)
            flat[0] = nbp
            k = 1
            for c, b in enumerate(r["breakpoints"]):
                m = min(len(b), (SEG_CAP - k) // 2)
                flat[k:k + 2 * m:2] = c
                flat[k + 1:k + 2 * m:2] = b[:m]
                k += 2 * m
            seg_local.copy_(torch.from_numpy(flat))
            dist.all_gather_into_tensor(seg_all, seg_local)
            torch.cuda.synchronize()
        e2e_s.append(time.perf_counter() - t_call)  # the synchronous call: H2D + kernels + D2H (+ the gather)
        return r, nbp

    W = max(W, 3)  # timing rules: at least three warm-up steps
    for _ in range(W):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, launches, stage_acc, visits, d2h = [], 0, {}, 0.0, 0
    e2e_s.clear()
    t0 = time.perf_counter()
    for _ in range(K):
        r, nbp = step()
        dev_ms.append(eng.last_kernel_ms)
        launches += eng.last_launches
        for k, v in eng.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        pstats = eng.last_partition_stats()
        visits = pstats["visits"]
        d2h = len(r["kept_index"]) * 8 + nbp * 4 + nc * 4 + 4096
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    clocks = sampler.summary() if rank == 0 else None
    loop_ms = (t1 - t0) * 1e3 / K  # includes the L2 flush between steps
    wall_ms = sum(e2e_s) * 1e3 / K
    kern_ms = sum(dev_ms) / K
    del loop_ms
    tt = torch.tensor([wall_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_ms, kern_ms = tt.tolist()
    total_bins = nb * world if world == 1 else None
    if world > 1:
        nbt = torch.tensor([nb], dtype=torch.int64, device=dev)
        dist.all_reduce(nbt)
        total_bins = int(nbt.item())
    strong = None
    if world > 1:
        # the same path on ONE sample with its chromosomes sharded over the ranks: Clean and the
        # genome-wide scalars replicated, each rank segments its LPT share, one all-gather of breakpoints
        from canvas_b200 import multi
        s0 = synth.make_sample(config=2, sample=0, scale=args.scale)
        in0 = dict(chrom=pin.array(s0.chrom), start=pin.array(s0.start), stop=pin.array(s0.stop), count=pin.array(s0.count),
                   gc=pin.array(s0.gc))
        lens0 = np.bincount(s0.chrom, minlength=len(s0.names))
        ts = []
        for it in range(W + K):
            flush.fill_(1)
            torch.cuda.synchronize()
            dist.barrier()
            ta = time.perf_counter()
            p0 = multi.clean_partition_wavelet_sharded(
                eng, (in0["chrom"], s0.is_autosome, s0.is_chr_y, in0["start"], in0["stop"], in0["count"], in0["gc"]), lens0,
                is_germline=True, out=out)
            torch.cuda.synchronize()
            if it >= W:
                ts.append(time.perf_counter() - ta)
        tst = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64, device=dev)
        dist.all_reduce(tst, op=dist.ReduceOp.MAX)
        strong = {"what": "ONE config-2 sample: Clean + scalars replicated, chromosomes LPT-sharded over the ranks, one "
                          "all-gather of breakpoints; wall clock of the fused C-ABI call incl. H2D/D2H, max over ranks",
                  "ms_per_sample": tst.item() * 1e3, "Mbins_per_s": len(s0) / tst.item() / 1e6,
                  "breakpoints": sum(len(b) for b in p0["breakpoints"])}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm, peak_src = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full captures
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    stages = {k: v / K for k, v in stage_acc.items()}
    # the chromosomes' pipelines (decomposition stages + finish) overlap: the decomposition's own span is the device-side
    # timestamp from its first kernel to its last tiny-stage thread (partition_stats), not a stage bracket
    dec_ms = pstats.get("decompose_span_ms", 0.0) or stages.get("decompose", -1.0)
    alg_bytes = 8.0 * visits  # one f64 prefix sum read per bin visit (8 * L_eff B/bin)
    achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else None
    line = {"metric": METRIC, "value": total_bins / (kern_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bins_per_sample": nb, "samples": world,
                       "l2": "flushed between steps (256 MiB device write)", "parallelism": f"sample-per-gpu x{world}",
                       "exchange": "NCCL all-gather of per-sample segment lists" if world > 1 else "none"},
            "e2e": {"value": total_bins / (wall_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wall_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timing": "wall clock around the synchronous C-ABI call with pinned host buffers, max over ranks"},
            "gpu_launches": launches,
            "stages_ms": stages, "partition_stats": pstats,
            "roofline": {"kernel": "Unbalanced-Haar decomposition (uh_chain + uh_mid + uh_small + uh_tiny kernels, one pipeline per chromosome)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": (achieved / hbm) if achieved else None,
                         "traffic": traffic.get("uh_decompose"), "traffic_source": traffic.get("source"),
                         "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                         "l_eff": visits / max(1, len(r["kept_index"])),
                         "note": "the prefix sums (24 MB) stay in L2: DRAM traffic (cold-cache ncu, each of the four stage kernels "
                                 "re-reading them once) is ~1/%d of the algorithmic bytes; the stage is bound by the dependent "
                                 "chain of big nodes, not by HBM" % max(1, round(alg_bytes / max(1, traffic.get("uh_decompose", 1))))},
            "clocks": clocks, "device": eng.describe()}
    # Clean / Partition totals against the HBM roofline with SURVEY.md 8(d)'s accounting: Clean = 101 B per input bin,
    # wavelet Partition = 8 * L_eff + 72 B per cleaned bin
    kept = max(1, len(r["kept_index"]))
    l_eff = visits / kept
    part_ms = sum(stages.get(k, 0.0) for k in ("scalars", "decompose", "finish"))
    totals = {}
    for name, nbytes, ms in (("clean", 101.0 * nb, stages.get("clean", 0.0)), ("partition_wavelet", (8.0 * l_eff + 72.0) * kept, part_ms)):
        if ms > 0:
            ach = nbytes / (ms * 1e-3) / 1e9
            totals[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved": ach, "unit": "GB/s", "frac": ach / hbm}
    line["roofline_totals"] = totals
    # K8 normalise stream on batches larger than L2 (the kernel BASELINE.json's roofline target names):
    # 8 samples = config 5 (223 MB), 16 samples (446 MB) and 32 samples (892 MB)
    if strong is not None:
        line["strong_scaling_single_sample"] = strong
    try:
        n16 = (nb // 16) * 16
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
        o = run_oracle_once(s, threads)
        line["cpu_baseline"] = {"value": nb / o["total_s"] / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "one full config-2 sample (Clean 1 thread, Partition 1 thread per chromosome)",
                                "clean_ms": o["clean_s"] * 1e3, "partition_ms": o["partition_s"] * 1e3}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
