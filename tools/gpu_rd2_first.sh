#!/bin/bash
# Round-2 opening session on the GPU box: baseline bench of the inherited build, compute-sanitizer over every kernel family,
# CanvasBin at chr1 scale, and `ncu --set full` captures of the kernel families round 1 left without one (bin_*, lo_*, cbs_*, hmm_*).
# Usage (under gpurun, from the repo root): bash tools/gpu_rd2_first.sh <tag>
tag=${1:-rd2a}
out=gpurun_out
mkdir -p $out
timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cut -c1-300 $out/${tag}_bench.json
san() {  # tool, what, limit
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_driver.py $2 0.02 > $out/${tag}_san_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_san_$1_$2.log | tail -1)"
}
for what in fused k8 cbs hmm bin loess; do san memcheck $what 240; done
for what in fused k8 cbs hmm; do san racecheck $what 300; done
for what in fused k8 cbs; do san synccheck $what 240; done
san initcheck fused 240
timeout 900 python tools/bin_bench.py check > $out/${tag}_bin_bench.jsonl 2> $out/${tag}_bin_bench.err
cut -c1-220 $out/${tag}_bin_bench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bin_" -c 24 \
    -o $out/${tag}_ncu_bin -f python tools/bin_bench.py 64e6 > $out/${tag}_ncu_bin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lo_" -c 12 \
    -o $out/${tag}_ncu_loess -f python tools/sanitize_driver.py loess 1.0 > $out/${tag}_ncu_loess.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hmm_" -c 16 \
    -o $out/${tag}_ncu_hmm -f python tools/hmm_bench.py 0.3 > $out/${tag}_ncu_hmm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cbs_" -c 6 \
    -o $out/${tag}_ncu_cbs -f python tools/cbs_scale.py 0.03 > $out/${tag}_ncu_cbs.log 2>&1
timeout 600 python tools/cbs_scale.py 1.0 > $out/${tag}_cbs_config2.txt 2>&1
tail -3 $out/${tag}_cbs_config2.txt
ls -la $out | grep ${tag}
