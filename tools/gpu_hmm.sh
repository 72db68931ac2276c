#!/bin/bash
tag=${1:-hmm}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hmm_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest_hmm.log 2>&1
tail -25 gpurun_out/${tag}_pytest_hmm.log
timeout 600 python tools/hmm_bench.py 1.0 > gpurun_out/${tag}_hmm_bench.json 2> gpurun_out/${tag}_hmm_bench.err
cat gpurun_out/${tag}_hmm_bench.json; tail -5 gpurun_out/${tag}_hmm_bench.err
bash tools/gpu_quick.sh ${tag}
