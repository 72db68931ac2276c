#!/bin/bash
# 8-GPU session: sharded parity over NCCL, config-5 bench (+ strong scaling + config 4 side measurements), config 4 alone
tag=${1:-rd2n8}
N=${2:-8}
out=gpurun_out
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun.log | cut -c1-700 || tail -5 $out/${tag}_check_torchrun.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_${N}gpu.json"))
    print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    for r in d.get("per_rank", []):
        print("  rank", r["rank"], round(r["device_ms"], 3), round(r["e2e_ms"], 3), {k: round(v, 3) for k, v in r["stages_ms"].items()}, "x", round(r["exchange_ms"], 3), r["launches_per_step"])
    print("  strong", d.get("strong_scaling_single_sample"))
    c4 = d.get("config4") or {}
    print("  config4", c4.get("ms_per_step"), c4.get("Mbins_per_s"), c4.get("kernel_ms_max_rank"), c4.get("phases_ms_rank0"))
except Exception as e:
    print("bench failed", e); print(open("$out/${tag}_bench_${N}gpu.err").read()[-2500:])
PY
