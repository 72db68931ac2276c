#!/bin/bash
# pedigree chain: tests on one GPU, config-4 bench; with N > 1 also the sharded parity check and config 4 over N ranks
tag=${1:-rd2ped}
N=${2:-1}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_hmm_gpu.py tests/test_merge_gpu.py "tests/test_full_size_gpu.py::test_config4_trio_clean_merge_hmm" -x -q > $out/${tag}_pytest.log 2>&1
tail -5 $out/${tag}_pytest.log
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 > $out/${tag}_bench_c4_1gpu.json 2> $out/${tag}_bench_c4_1gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_1gpu.json")); c4 = d["config4"]
    print("config4 N=1", c4["ms_per_step"], c4["Mbins_per_s"], c4["kernel_ms_max_rank"], c4["phases_ms_rank0"], c4["launches_rank0"])
except Exception as e:
    print("c4 failed", e); print(open("$out/${tag}_bench_c4_1gpu.err").read()[-2000:])
PY
if [ "$N" -gt 1 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 0.25 > $out/${tag}_check_torchrun.log 2>&1
grep "multi-gpu check" $out/${tag}_check_torchrun.log | cut -c1-900 || tail -5 $out/${tag}_check_torchrun.log
timeout 300 python tools/multi_gpu_check.py --single-process $N 0.1 > $out/${tag}_check_single_process.log 2>&1
grep "multi-gpu check" $out/${tag}_check_single_process.log | cut -c1-600 || tail -5 $out/${tag}_check_single_process.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --config 4 --steps 5 --warmup 3 > $out/${tag}_bench_c4_${N}gpu.json 2> $out/${tag}_bench_c4_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4_${N}gpu.json")); c4 = d["config4"]
    print("config4 N=$N", c4["ms_per_step"], c4["Mbins_per_s"], c4["kernel_ms_max_rank"], c4["phases_ms_rank0"], c4["nccl_ms_rank0"], c4["units_per_rank"])
except Exception as e:
    print("c4 failed", e); print(open("$out/${tag}_bench_c4_${N}gpu.err").read()[-2000:])
PY
fi
