#!/bin/bash
# compute-sanitizer over the final kernels of the round (fused call with early prefix sums and deferred candidate records,
# pedigree chain with kernel-store read-backs, prefetched call)
tag=${1:-rd2san}
out=gpurun_out
mkdir -p $out
san() {  # tool, what, limit
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_driver.py $2 0.02 > $out/${tag}_san_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_san_$1_$2.log | tail -1)"
}
san memcheck fused 200
san memcheck pedigree 200
san racecheck fused 300
san racecheck pedigree 300
san synccheck pedigree 200
