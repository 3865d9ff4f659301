"""Compact per-kernel table of the key metrics of an `ncu --set full` report (reads `ncu -i REP --page raw --csv`).
usage: python scripts/ncu_summary.py REP.ncu-rep [REP2 ...] > profiles/xxx.csv"""
import csv, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "inst_executed", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")][:60] for r in body]
    print(f"# {rep}")
    print("metric,unit," + ",".join('"' + n + '"' for n in names))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(",".join([k, units[i]] + [r[i] for r in body]))
    print()
