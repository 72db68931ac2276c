#!/bin/bash
# 8-GPU: sharded parity (one process per GPU) and config 4 through the device-resident chain
tag=${1:-rd2n8d}
N=${2:-8}
out=gpurun_out
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun.log | cut -c1-600 || tail -5 $out/${tag}_check_torchrun.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --config 4 --steps 6 --warmup 3 > $out/${tag}_bench_c4_${N}gpu.json 2> $out/${tag}_bench_c4_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_${N}gpu.json")); c4 = d["config4"]
    print("config4 N=$N", c4["ms_per_step"], c4["Mbins_per_s"], c4["kernel_ms_max_rank"], c4["phases_ms_rank0"], c4["nccl_ms_rank0"], c4["units_per_rank"])
except Exception as e:
    print("c4 failed", e); print(open("$out/${tag}_bench_c4_${N}gpu.err").read()[-2000:])
PY
