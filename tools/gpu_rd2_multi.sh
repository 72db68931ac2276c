#!/bin/bash
# N-GPU session (gpurun --gpus N): parity of the sharded C-ABI entry points, bench at N (both arms).  Usage: bash tools/gpu_rd2_multi.sh <tag> <N>
tag=${1:-rd2m}
N=${2:-2}
out=gpurun_out
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun.log | cut -c1-900 || tail -5 $out/${tag}_check_torchrun.log
CANVAS_COMM_PACK_INTS=512 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun_smallpack.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun_smallpack.log | cut -c1-600 || tail -5 $out/${tag}_check_torchrun_smallpack.log
timeout 300 python tools/multi_gpu_check.py --single-process $N 0.25 > $out/${tag}_check_single_process.log 2>&1
grep "multi-gpu check" $out/${tag}_check_single_process.log | cut -c1-600 || tail -5 $out/${tag}_check_single_process.log
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
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --config 4 --steps 5 --warmup 3 > $out/${tag}_bench_c4_${N}gpu.json 2> $out/${tag}_bench_c4_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_${N}gpu.json"))
    print("config4 N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config4"]["phases_ms_rank0"])
except Exception as e:
    print("bench c4 failed", e); print(open("$out/${tag}_bench_c4_${N}gpu.err").read()[-2500:])
PY
if [ -n "$3" ]; then
timeout 600 python -m pytest tests/test_clean_gpu.py tests/test_bin_gpu.py -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 600 python tools/bin_bench.py > $out/${tag}_bin_bench.jsonl 2> $out/${tag}_bin_bench.err; cut -c1-200 $out/${tag}_bin_bench.jsonl
fi
