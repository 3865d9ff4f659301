"""Compact per-kernel table of the key metrics from a raw ncu CSV (`ncu -i REP --page raw --csv`, exported on the GPU box
because the reports themselves are too large to bring back).
usage: python scripts/distill_ncu.py RAW.csv "header comment" [--traffic profiles/roofline_traffic.json] > profiles/xxx.csv
--traffic: also store dram__bytes_read + dram__bytes_write per launch of every kernel (first launch of each name) there."""
import csv, json, re, sys

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

raw, comment = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(raw)) if r]
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units, body = rows[h], rows[h + 1], rows[h + 2:]
short = lambda n: re.sub(r"^(void )?(.*::)?", "", n.split("(")[0])[:60]
names = [short(r[hdr.index("Kernel Name")]) for r in body]
if "--first" in sys.argv:  # one column per kernel name (its first captured launch)
    keep, seen0 = [], set()
    for n, r in zip(names, body):
        if n not in seen0:
            seen0.add(n)
            keep.append((n, r))
    names, body = [k[0] for k in keep], [k[1] for k in keep]
print(f"# {comment}")
print("metric,unit," + ",".join('"' + n + '"' for n in names))
for k in KEYS + ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                 "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]:
    if k in hdr:
        i = hdr.index(k)
        print(",".join([k, units[i]] + [r[i] for r in body]))
if "--traffic" in sys.argv:
    path = sys.argv[sys.argv.index("--traffic") + 1]
    tj = json.load(open(path))
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    seen = set()
    for n, r in zip(names, body):
        if n in seen:
            continue
        seen.add(n)
        t = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        tj[n] = t
        tj.setdefault(n.split("<")[0], t)
        if n.split("<")[0] in ("adcensus_cost_kernel",):
            tj[n.split("<")[0]] = t
    json.dump(tj, open(path, "w"), indent=1)
