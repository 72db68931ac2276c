#!/bin/bash
tag=${1:-step}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merge_gpu.py tests/test_hmm_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest_new.log 2>&1
tail -25 gpurun_out/${tag}_pytest_new.log
CANVAS_DEBUG=1 timeout 600 python tools/hmm_bench.py 1.0 > gpurun_out/${tag}_hmm_bench.json 2> gpurun_out/${tag}_hmm_bench.err
cat gpurun_out/${tag}_hmm_bench.json; tail -40 gpurun_out/${tag}_hmm_bench.err
