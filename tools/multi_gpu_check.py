"""Multi-GPU parity on real hardware (run under torchrun, one rank per GPU, NCCL):
chromosome-sharded wavelet and CBS partition + the single all-gather must equal the single-GPU call.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [scale]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from canvas_b200 import multi, native, synth
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = native.Engine(local)
    s = synth.make_sample(config=2, sample=3, scale=scale, n_events=60)
    c = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)  # replicated on every rank
    off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
    from canvas_b200 import textcodec
    cov = textcodec.f2_roundtrip(c["count"])
    ew = max(2000, int(100000 * scale))
    full = eng.partition_wavelet(off, cov, is_germline=True, evenness_window=ew)
    shard = multi.partition_wavelet_sharded(eng, off, cov, is_germline=True, evenness_window=ew)
    ok_w = all(a.tolist() == b.tolist() for a, b in zip(full["breakpoints"], shard["breakpoints"]))
    ok_w = ok_w and full["cv"] == shard["cv"] and np.array_equal(full["factor_of_three"], shard["factor_of_three"])
    # the fused Clean + partition call, sharded: equals the single-GPU fused call
    fused = eng.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, is_germline=True,
                                        evenness_window=ew)
    fs = multi.clean_partition_wavelet_sharded(eng, (s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc),
                                               np.bincount(s.chrom, minlength=len(s.names)), is_germline=True, evenness_window=ew)
    ok_w = ok_w and all(a.tolist() == b.tolist() for a, b in zip(fused["breakpoints"], fs["breakpoints"]))
    ok_w = ok_w and all(a.tolist() == b.tolist() for a, b in zip(fused["breakpoints"], full["breakpoints"]))
    # CBS on a few chromosomes' worth of bins (permutation tests are heavier)
    ncb = min(len(off) - 1, 6)
    off_c = off[:ncb + 1]
    cov_c = cov[:off_c[-1]]
    full_c = eng.partition_cbs(off_c, cov_c)
    shard_c = multi.partition_cbs_sharded(eng, off_c, cov_c)
    ok_c = all(a["len"].tolist() == b["len"].tolist() for a, b in zip(full_c["segments"], shard_c["segments"]))
    full_h = eng.partition_hmm(off, cov, per_sample=True)
    shard_h = multi.partition_hmm_sharded(eng, off, cov, per_sample=True)
    ok_h = all(a.tolist() == b.tolist() for a, b in zip(full_h["breakpoints"], shard_h["breakpoints"]))
    flag = torch.tensor([int(ok_w), int(ok_c and ok_h)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-gpu check: world", world, "bins", len(cov), "wavelet", bool(flag[0].item()), "cbs+hmm", bool(flag[1].item()),
              "breakpoints", sum(len(b) for b in full["breakpoints"]), "cbs segments", sum(len(x["len"]) for x in full_c["segments"]),
              "owners", shard["owner"].tolist())
    dist.destroy_process_group()
    eng.close()
    return 0 if (ok_w and ok_c and ok_h) else 1


if __name__ == "__main__":
    sys.exit(main())
