"""Summarise an `ncu --set full` report (.ncu-rep) into the few per-launch metrics the profiles/ notes cite.
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__cluster_size"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (per launch; caches flushed before each launch by ncu)")
    for r in rows[2:]:
        print(r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                print(f"    {w:70s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")


if __name__ == "__main__":
    main()
