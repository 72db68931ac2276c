#!/bin/bash
# Round-2 session: full GPU test tier, bench, debug timeline of one warm fused call, leftovers of the sanitizer / ncu list.
tag=${1:-rd2c}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -5 $out/${tag}_pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
PY
CANVAS_DEBUG=1 timeout 300 python tools/profile_driver.py 1.0 3 fused > $out/${tag}_timeline.txt 2>&1
tail -32 $out/${tag}_timeline.txt | cut -c1-150
san() {  # tool, what, limit
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_driver.py $2 0.02 > $out/${tag}_san_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_san_$1_$2.log | tail -1)"
}
san racecheck fused 300
san memcheck fused 240
for t in memcheck racecheck synccheck; do san $t k8 240; done
if [ -n "$2" ]; then
cap() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/${tag}_$name -f "$@" > $out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_full_$name.txt 2>&1
  wc -l $out/${tag}_ncu_full_$name.txt
}
cap loess "loess_" 12 python tools/sanitize_driver.py loess 1.0
cap sel "sel_hist_contig_kernel" 14 python tools/profile_driver.py 1.0 1 fused
fi
