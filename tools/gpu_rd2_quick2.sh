#!/bin/bash
tag=$1
out=gpurun_out
timeout 900 python -m pytest tests/test_wavelet_gpu.py tests/test_full_size_gpu.py tests/test_comm_gpu.py tests/test_modules_gpu.py -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$i.json 2> $out/${tag}_bench_$i.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$i.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_pipelined"]["value"], d["e2e_pipelined"]["ms_per_step"], d["roofline"]["frac"], d["partition_stats"]["big_phase_ms"], d["partition_stats"]["decompose_span_ms"])
PY
done
CANVAS_NO_EARLY_SCAN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_noearly.json 2> $out/${tag}_bench_noearly.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_noearly.json"))
print("no early scan", {k: d[k] for k in ("value", "ms_per_step", "stages_ms")}, d["e2e"]["value"])
PY
CANVAS_DEBUG=1 timeout 300 python tools/profile_driver.py 1.0 3 fused > $out/${tag}_timeline.txt 2>&1
grep "fused\]" $out/${tag}_timeline.txt | tail -17
