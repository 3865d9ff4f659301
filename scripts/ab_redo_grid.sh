#!/bin/bash
# A/B of library variants in the escalating regime (integer AD costs) and the listed regime: scripts/ab_redo_grid.sh name ...
cd "$(dirname "$0")/.."
python scripts/agg_time_escalated.py 1080x1920x256 5
python scripts/agg_time.py 1080x1920x256 10
for v in "$@"; do
  export SMT_B200_LIB=$PWD/stereo_match_traditional_b200/csrc/build_var/$v/libsmt_b200.so
  python scripts/agg_time_escalated.py 1080x1920x256 5
  python scripts/agg_time.py 1080x1920x256 10
done
