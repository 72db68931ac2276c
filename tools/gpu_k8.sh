#!/bin/bash
# K8 experiments on the GPU box: parity of both variants, the sweep, one full ncu capture of each variant.
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_clean_gpu.py -m gpu -x -q -k normalize > gpurun_out/${tag}_k8_pytest.log 2>&1
tail -3 gpurun_out/${tag}_k8_pytest.log
timeout 900 python tools/k8_sweep.py > gpurun_out/${tag}_k8_sweep.jsonl 2> gpurun_out/${tag}_k8_sweep.err
cat gpurun_out/${tag}_k8_sweep.jsonl; tail -3 gpurun_out/${tag}_k8_sweep.err
K8_REPEATS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normalize_apply" -c 4 \
    -o gpurun_out/${tag}_k8_batch -f python tools/k8_sweep.py quick > gpurun_out/${tag}_ncu_k8_batch.log 2>&1
tail -5 gpurun_out/${tag}_ncu_k8_batch.log
