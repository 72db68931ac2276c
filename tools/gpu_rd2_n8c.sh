#!/bin/bash
# 8-GPU closing session: sharded parity (incl. the pedigree chain), config-5 bench with side measurements, config 4 over N ranks
tag=${1:-rd2n8c}
N=${2:-8}
out=gpurun_out
mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun.log | cut -c1-900 || tail -5 $out/${tag}_check_torchrun.log
timeout 300 python tools/multi_gpu_check.py --single-process $N 0.1 > $out/${tag}_check_single_process.log 2>&1
grep "multi-gpu check" $out/${tag}_check_single_process.log | cut -c1-500 || tail -5 $out/${tag}_check_single_process.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_${N}gpu.json"))
    print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pipelined", d["e2e_pipelined"]["value"], d["e2e_pipelined"]["ms_per_step"])
    for r in d.get("per_rank", []):
        print("  rank", r["rank"], round(r["device_ms"], 3), round(r["e2e_ms"], 3), {k: round(v, 3) for k, v in r["stages_ms"].items()})
    print("  strong", d.get("strong_scaling_single_sample"))
    c4 = d.get("config4") or {}
    print("  config4", c4.get("ms_per_step"), c4.get("Mbins_per_s"), c4.get("kernel_ms_max_rank"), c4.get("phases_ms_rank0"), c4.get("units_per_rank"))
except Exception as e:
    print("bench failed", e); print(open("$out/${tag}_bench_${N}gpu.err").read()[-2500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --config 4 --steps 6 --warmup 3 > $out/${tag}_bench_c4_${N}gpu.json 2> $out/${tag}_bench_c4_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_${N}gpu.json")); c4 = d["config4"]
    print("config4 N=$N", c4["ms_per_step"], c4["Mbins_per_s"], c4["kernel_ms_max_rank"], c4["phases_ms_rank0"], c4["nccl_ms_rank0"], c4["units_per_rank"])
except Exception as e:
    print("c4 failed", e); print(open("$out/${tag}_bench_c4_${N}gpu.err").read()[-2000:])
PY
