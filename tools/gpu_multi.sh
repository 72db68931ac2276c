#!/bin/bash
# N-GPU session (gpurun --gpus N): sharded-partition parity over NCCL + the N-GPU bench line.
n=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    tools/multi_gpu_check.py 0.25 > gpurun_out/${tag}_check_n$n.log 2>&1
tail -4 gpurun_out/${tag}_check_n$n.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
cat gpurun_out/${tag}_bench_n$n.json | cut -c1-2500; tail -3 gpurun_out/${tag}_bench_n$n.err
