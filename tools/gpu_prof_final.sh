#!/bin/bash
# `ncu --set full` captures of the kernels of the final build: decomposition + finish, both selection kernels, K8.
tag=${1:-prof}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"uh_chain_kernel|uh_mid_kernel|uh_small_kernel|uh_tiny_kernel|uh_finish_kernel" --launch-skip 5 -c 5 \
    -o gpurun_out/${tag}_uh -f python tools/profile_driver.py 1.0 2 both > gpurun_out/${tag}_ncu_uh.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_contig_kernel" --launch-skip 16 -c 3 \
    -o gpurun_out/${tag}_sel_contig -f python tools/profile_driver.py 1.0 2 both > gpurun_out/${tag}_ncu_sel_contig.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_scatter_kernel" --launch-skip 20 -c 4 \
    -o gpurun_out/${tag}_sel_scatter -f python tools/profile_driver.py 1.0 2 both > gpurun_out/${tag}_ncu_sel_scatter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normalize_apply_bulk_kernel" --launch-skip 3 -c 2 \
    -o gpurun_out/${tag}_k8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_k8.log 2>&1
ls -la gpurun_out/${tag}*
