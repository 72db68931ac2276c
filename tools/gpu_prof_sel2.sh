#!/bin/bash
tag=${1:-sel}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_scatter_kernel" --launch-skip 20 -c 5 \
    -o gpurun_out/${tag}_sel_scatter -f python tools/profile_driver.py 1.0 1 clean > gpurun_out/${tag}_ncu_sel_scatter.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_contig_kernel" -c 2 \
    -o gpurun_out/${tag}_sel_contig -f python tools/profile_driver.py 1.0 1 both > gpurun_out/${tag}_ncu_sel_contig.log 2>&1
ls -la gpurun_out/${tag}*
