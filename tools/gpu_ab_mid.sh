#!/bin/bash
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_wavelet_gpu.py tests/test_full_size_gpu.py tests/test_modules_gpu.py -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
for mt in 0 256 512 1024; do
  CANVAS_MID_THREADS=$mt timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_mt$mt.json 2> $out/${tag}_bench_mt$mt.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_mt$mt.json"))
print("mid threads $mt", round(d["value"], 1), round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["value"], 1), d["partition_stats"]["decompose_span_ms"])
PY
done
CANVAS_DEBUG=1 python tools/sample_variance.py 8 0 2> $out/${tag}_sample0_timeline.txt > /dev/null; grep "\[mid\]" $out/${tag}_sample0_timeline.txt | tail -6; grep "\[pipe\]" $out/${tag}_sample0_timeline.txt | tail -24 | cut -c1-200
