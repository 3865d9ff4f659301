#!/bin/bash
# A/B timing of library variants on the GPU box: scripts/ab_agg.sh name1 name2 ...   (variants built by `make -C csrc variant`)
cd "$(dirname "$0")/.."
python scripts/agg_time.py 1080x1920x256 10
for v in "$@"; do
  SMT_B200_LIB=$PWD/stereo_match_traditional_b200/csrc/build_var/$v/libsmt_b200.so python scripts/agg_time.py 1080x1920x256 10
done
