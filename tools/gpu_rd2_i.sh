#!/bin/bash
tag=${1:-rd2i}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_bin_gpu.py tests/test_hmm_gpu.py tests/test_full_size_gpu.py -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 600 python tools/bin_bench.py check > $out/${tag}_bin_bench.jsonl 2> $out/${tag}_bin_bench.err; cut -c1-200 $out/${tag}_bin_bench.jsonl; grep -o '"parity_bit_exact": [a-z]*' $out/${tag}_bin_bench.jsonl | sort | uniq -c
for cfg in 4; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > $out/${tag}_bench_c$cfg.json 2> $out/${tag}_bench_c$cfg.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c$cfg.json"))
    print("config $cfg", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("stages_ms"), d.get("config4", {}).get("phases_ms_rank0"), d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("config $cfg failed", e); print(open("$out/${tag}_bench_c$cfg.err").read()[-1500:])
PY
done
