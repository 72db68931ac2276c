#!/bin/bash
# ncu --set full of the selection passes (Clean's GC-bucket select, the partition scalars' window select)
tag=${1:-sel}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_scatter_kernel" --launch-skip 20 -c 6 \
    -o gpurun_out/${tag}_sel_scatter -f python tools/profile_driver.py 1.0 1 clean > gpurun_out/${tag}_ncu_sel_scatter.log 2>&1
tail -3 gpurun_out/${tag}_ncu_sel_scatter.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sel_hist_contig_kernel|wv_evenness_kernel|fused_coverage" -c 4 \
    -o gpurun_out/${tag}_sel_contig -f python tools/profile_driver.py 1.0 1 both > gpurun_out/${tag}_ncu_sel_contig.log 2>&1
tail -3 gpurun_out/${tag}_ncu_sel_contig.log
CANVAS_DEBUG=1 timeout 300 python tools/profile_driver.py 1.0 3 both > gpurun_out/${tag}_debug.log 2>&1
tail -40 gpurun_out/${tag}_debug.log
