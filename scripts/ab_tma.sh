#!/bin/bash
# A/B on the GPU box: the bulk-copy (TMA) staged non-dividing pass against the LDG/STG one, same library build otherwise.
cd "$(dirname "$0")/.."
L=$PWD/stereo_match_traditional_b200/csrc/build_var/tma/libsmt_b200.so
echo "== parity with the TMA pass"; SMT_B200_LIB=$L SMT_FX1_TMA=1 python -m pytest tests/test_gpu_agg_fixedpoint.py -q -x -m gpu 2>&1 | tail -2
echo "== timing: dev build, LDG/STG"; SMT_B200_LIB=$L python scripts/agg_time.py 1080x1920x256 10 | grep aggregate
echo "== timing: dev build, TMA";     SMT_B200_LIB=$L SMT_FX1_TMA=1 python scripts/agg_time.py 1080x1920x256 10 | grep aggregate
echo "== per-kernel (ncu, 1 frame)"
SMT_B200_LIB=$L SMT_FX1_TMA=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum --clock-control none -k regex:agg_fx1 --csv --log-file gpurun_out/r2_tma_ab.csv python scripts/one_frame.py 1080x1920x256 fast 1 > /dev/null 2>&1
grep -E "agg_fx1" gpurun_out/r2_tma_ab.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | head -40
