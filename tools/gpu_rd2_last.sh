#!/bin/bash
# last validation of the round on one GPU: full GPU test tier, smoke(), both bench arms, config 4
tag=${1:-rd2last}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_pipelined"]["value"], d["e2e_pipelined"]["ms_per_step"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 > $out/${tag}_bench_c4_1gpu.json 2> $out/${tag}_bench_c4_1gpu.err
python -c "
import json; d=json.load(open('$out/${tag}_bench_c4_1gpu.json')); print('c4', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config4']['phases_ms_rank0'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; cut -c1-200 $out/${tag}_bench_reference.json
