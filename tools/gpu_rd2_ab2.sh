#!/bin/bash
# A/B of the two ordering waits added to the partition (measurement only)
tag=$1
out=gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$label.json 2> $out/${tag}_bench_$label.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$label.json"))
print("$label", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["stages_ms"].items()}, round(d["e2e"]["ms_per_step"], 3))
PY
}
for rep in 1 2 3; do
run split$rep A=1
run whole$rep CANVAS_SPLIT_PIPE=0
done
