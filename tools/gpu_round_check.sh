#!/bin/bash
# One GPU-box session: GPU tests, bench (both arms), ncu launch list and full captures of the main kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round_check.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
tail -c 300 gpurun_out/${tag}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"uh_chain_kernel|uh_mid_kernel|uh_small_kernel|uh_finish_kernel" -c 8 \
    -o gpurun_out/${tag}_uh -f python tools/profile_driver.py 1.0 2 both > gpurun_out/${tag}_ncu_uh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normalize_apply_kernel" -c 2 \
    -o gpurun_out/${tag}_k8 -f python tools/profile_driver.py 1.0 1 clean > gpurun_out/${tag}_ncu_k8.log 2>&1
ls -la gpurun_out | tail -12
