#!/bin/bash
# Round-2 session b: new comm tests, racecheck/initcheck logs, and the ncu captures of the kernel families round 1 left without
# one -- summarised on the box (the .ncu-rep files stay there: gpurun_out/ comes back only while it is under 64 MiB).
tag=${1:-rd2b}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_comm_gpu.py -x -q > $out/${tag}_pytest_comm.log 2>&1; tail -3 $out/${tag}_pytest_comm.log
san() {  # tool, what, limit
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_driver.py $2 0.02 > $out/${tag}_san_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_san_$1_$2.log | tail -1)"
}
for what in fused k8 cbs hmm bin loess; do san memcheck $what 240; done
for what in fused k8 cbs hmm; do san racecheck $what 300; done
for what in fused k8 cbs; do san synccheck $what 240; done
cap() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/${tag}_$name -f "$@" > $out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_full_$name.txt 2>&1
  wc -l $out/${tag}_ncu_full_$name.txt
}
cap bin "bin_|read_gc|gc_tile|gc_prefix" 24 python tools/bin_bench.py 64e6
cap loess "lo_" 12 python tools/sanitize_driver.py loess 1.0
cap hmm "hmm_" 16 python tools/hmm_bench.py 0.3
cap cbs "cbs_" 6 python tools/cbs_scale.py 0.03
ls -la $out | grep ${tag} | head -50
