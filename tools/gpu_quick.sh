#!/bin/bash
# quick session: selected tests, bench, debug timeline.  Usage: bash tools/gpu_quick.sh <tag> "<pytest args>"
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest $2 -x -q > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["partition_stats"]["decompose_span_ms"])
PY
CANVAS_DEBUG=1 timeout 300 python tools/profile_driver.py 1.0 3 fused > $out/${tag}_timeline.txt 2>&1
grep "\[pipe\]" $out/${tag}_timeline.txt | tail -24 | cut -c1-200
if [ -n "$3" ]; then bash -c "$3"; fi
