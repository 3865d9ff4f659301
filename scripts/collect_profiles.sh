#!/bin/bash
# Runs on the GPU box: the evidence under profiles/.  Raw-metric CSVs are exported on the box (gpurun_out/ is limited to
# 64 MiB, the reports are not).  Every ncu command is preceded by the same command exiting 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
BL="python bench.py --steps 2 --warmup 3 --stream-frames 4 --no-cpu-baseline"
$BL > $O/r2_bench_short.json 2> $O/r2_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_bench.csv $BL > /dev/null 2>&1
python scripts/one_frame.py 1080x1920x256 fast 2 > $O/r2_one_frame.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_pfull_1080p_d256.csv python scripts/one_frame.py 1080x1920x256 fast 2 > /dev/null 2>&1
if [ "$1" = "all" ]; then
  ncu --set full --clock-control none --import-source on -k regex:'agg_fx' -s 10 -c 5 -o $O/r2_ncu_agg_fx python scripts/one_frame.py 1080x1920x256 fast 2 > /dev/null 2>&1
  ncu -i $O/r2_ncu_agg_fx.ncu-rep --page raw --csv > $O/r2_ncu_agg_fx_raw.csv 2>/dev/null
fi
ncu --set full --clock-control none -k regex:'scan_path|adcensus_cost|wta_first' -s 7 -c 7 -o /tmp/r2_ncu_rest python scripts/one_frame.py 1080x1920x256 fast 2 > /dev/null 2>&1
ncu -i /tmp/r2_ncu_rest.ncu-rep --page raw --csv > $O/r2_ncu_scan_cost_wta_raw.csv 2>/dev/null
python scripts/matchers_once.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'asw_kernel|box_sum|ncc_|sad_select' -c 12 -o /tmp/r2_ncu_matchers python scripts/matchers_once.py > /dev/null 2>&1
ncu -i /tmp/r2_ncu_matchers.ncu-rep --page raw --csv > $O/r2_ncu_matchers_raw.csv 2>/dev/null
python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference --steps 4 --warmup 1 > $O/r2_bench_reference_arm.json 2>/dev/null
for w in sad_teddy pfull_teddy ncc_asw_720p; do python bench.py --workload $w --steps 10 --warmup 3 > $O/r2_bench_$w.json 2>/dev/null; done
ls -la $O/r2_*
