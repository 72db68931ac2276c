#!/bin/bash
# A/B over an environment switch.  Usage: bash tools/gpu_ab.sh <tag> <ENVVAR> "<values>" ["pytest args"]
tag=$1; var=$2; vals=$3
out=gpurun_out
mkdir -p $out
if [ -n "$4" ]; then timeout 900 python -m pytest $4 -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log; fi
for v in $vals; do
  env $var=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$v.json 2> $out/${tag}_bench_$v.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$v.json"))
print("$var=$v", round(d["value"], 1), round(d["ms_per_step"], 4), {k: round(x, 3) for k, x in d["stages_ms"].items()}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 3), d["partition_stats"]["decompose_span_ms"])
PY
done
CANVAS_DEBUG=1 python tools/sample_variance.py 8 0 2> $out/${tag}_sample0_timeline.txt > /dev/null; grep "\[pipe\]" $out/${tag}_sample0_timeline.txt | tail -24 | cut -c1-200
