#!/bin/bash
# 8-GPU session B: host topology, config-5 bench without side measurements, host-thread times of one rank, config 4
tag=${1:-rd2n8b}
N=${2:-8}
out=gpurun_out
mkdir -p $out
{ nvidia-smi topo -m; lscpu | grep -i -E "^CPU\(s\)|socket|NUMA|Model name|Thread"; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2; } > $out/${tag}_topology.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_${N}gpu.json"))
    print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    for r in d.get("per_rank", []):
        print("  rank", r["rank"], round(r["device_ms"], 3), round(r["e2e_ms"], 3), {k: round(v, 3) for k, v in r["stages_ms"].items()}, r.get("host_cpus"))
except Exception as e:
    print("bench failed", e); print(open("$out/${tag}_bench_${N}gpu.err").read()[-2500:])
PY
CANVAS_HOST_TIMES=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_1gpu_hosttimes.json 2> $out/${tag}_hosttimes.err
grep "\[host\]" $out/${tag}_hosttimes.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --config 4 --steps 5 --warmup 3 > $out/${tag}_bench_c4_${N}gpu.json 2> $out/${tag}_bench_c4_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_${N}gpu.json")); c4 = d["config4"]
    print("config4 N=$N", c4["ms_per_step"], c4["Mbins_per_s"], c4["kernel_ms_max_rank"], c4["phases_ms_rank0"])
except Exception as e:
    print("c4 failed", e)
PY
