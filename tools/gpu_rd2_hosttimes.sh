#!/bin/bash
tag=$1
out=gpurun_out
CANVAS_HOST_TIMES=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_ht.json 2> $out/${tag}_ht.err
grep "\[host\]" $out/${tag}_ht.err | tail -8
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_ht.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms")}, d["e2e"]["value"], d["e2e_pipelined"]["value"])
PY
