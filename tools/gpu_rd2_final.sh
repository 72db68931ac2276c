#!/bin/bash
# Round-2 closing session on one GPU: full GPU test tier, bench lines (configs 2, 3, 4, CBS, reference arm), debug timeline,
# launch list of the bench under ncu, ncu --set full of the kernels changed this round, sanitizers over the final code.
tag=${1:-rd2w}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -4 $out/${tag}_pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "stages_ms", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_pipelined"]["value"], d["e2e_pipelined"]["ms_per_step"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"))
PY
timeout 300 python bench.py --config 3 --steps 6 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 > $out/${tag}_bench_c4_1gpu.json 2> $out/${tag}_bench_c4_1gpu.err
python - <<PY
import json
for f in ("c3", "c4_1gpu"):
    try:
        d = json.load(open("$out/${tag}_bench_%s.json" % f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], (d.get("config4") or {}).get("phases_ms_rank0"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; cut -c1-400 $out/${tag}_bench_reference.json
CANVAS_DEBUG=1 timeout 300 python tools/profile_driver.py 1.0 3 fused > $out/${tag}_timeline.txt 2>&1
CANVAS_HOST_TIMES=1 timeout 300 python tools/profile_driver.py 1.0 4 fused 2>&1 | grep "\[host\]" > $out/${tag}_host_times.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_launches_bench.log 2>&1
wc -l $out/${tag}_launches.csv
cap() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/${tag}_$name -f "$@" > $out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_full_$name.txt 2>&1
  wc -l $out/${tag}_ncu_full_$name.txt
}
CANVAS_NO_GRAPH=1 cap uh "uh_" 60 python tools/profile_driver.py 1.0 1 fused
CANVAS_NO_GRAPH=1 cap scalars "fused_coverage|sel_hist_contig|wv_scan_apply|rq_" 24 python tools/profile_driver.py 1.0 1 fused
cap bin "bin_accum|read_gc_tile|bin_sum_weighted|bin_screen" 10 python tools/bin_bench.py 64e6
timeout 600 python tools/bin_bench.py check > $out/${tag}_bin_bench.jsonl 2> $out/${tag}_bin_bench.err
san() {  # tool, what, limit
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_driver.py $2 0.02 > $out/${tag}_san_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_san_$1_$2.log | tail -1)"
}
for what in fused pedigree bin; do san memcheck $what 300; done
for what in fused pedigree; do san racecheck $what 400; done
san synccheck fused 300
ls -la $out | grep ${tag} | awk '{print $5, $9}'
