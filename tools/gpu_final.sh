#!/bin/bash
# Round-end evidence: GPU tests, both bench arms, launch list of the bench, timeline, HMM bench.
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-200 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
cut -c1-200 gpurun_out/${tag}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
bash tools/gpu_timeline.sh ${tag} > /dev/null
timeout 600 python tools/hmm_bench.py 1.0 > gpurun_out/${tag}_hmm_bench.json 2> gpurun_out/${tag}_hmm_bench.err
cat gpurun_out/${tag}_hmm_bench.json
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/aux_bench.py > gpurun_out/${tag}_aux_kernels.jsonl 2> gpurun_out/${tag}_aux.err
cut -c1-160 gpurun_out/${tag}_aux_kernels.jsonl
