#!/bin/bash
tag=${1:-rd2j}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_merge_gpu.py tests/test_wavelet_gpu.py -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
for mt in 256 512 1024; do
  CANVAS_MID_THREADS=$mt timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_mt$mt.json 2> $out/${tag}_bench_mt$mt.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_mt$mt.json"))
print("mid threads $mt", d["value"], d["ms_per_step"], d["stages_ms"], "e2e", d["e2e"]["value"], d["partition_stats"]["decompose_span_ms"])
PY
done
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench_c4.json"))
    print("config 4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("config4", {}).get("phases_ms_rank0"))
except Exception as e:
    print("config 4 failed", e); print(open("$out/${tag}_bench_c4.err").read()[-1500:])
PY
cap() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/${tag}_$name -f "$@" > $out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_full_$name.txt 2>&1
  wc -l $out/${tag}_ncu_full_$name.txt
}
cap bin "bin_accum|read_gc_tile|bin_sum_weighted|bin_fragments_kernel|bin_screen_kernel" 6 python tools/bin_bench.py
