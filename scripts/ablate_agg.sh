#!/bin/bash
# Development aid: rebuild libsmt_b200.so with the SMT_ABL_* switches of cross_agg_seq.cu and time one aggregation
# (scripts/time_agg.py).  Run on the GPU box; restores the product build at the end.
set -u
cd "$(dirname "$0")/.."
CS=stereo_match_traditional_b200/csrc
build() { touch $CS/cross_agg_seq.cu; make -C $CS NVCC="/usr/local/cuda/bin/nvcc $1" 2>&1 | grep -E "error" ; }
for v in "" "-DSMT_ABL_NOLOAD -DSMT_ABL_NOSTORE" "-DSMT_ABL_NOLOAD" "-DSMT_ABL_NOSTORE" "-DSMT_ABL_OCC16" "-DSMT_ABL_OCC16 -DSMT_ABL_NOLOAD -DSMT_ABL_NOSTORE" "-DSMT_ABL_OCC16 -DSMT_ABL_NOLOAD" "-DSMT_ABL_OCC16 -DSMT_ABL_NOSTORE"; do
  build "$v"
  echo "variant [$v]: $(python scripts/time_agg.py 2>&1 | tail -1)"
  if [ "${NCU:-0}" = "1" ] && { [ "$v" = "-DSMT_ABL_NOLOAD -DSMT_ABL_NOSTORE" ] || [ "$v" = "-DSMT_ABL_OCC16 -DSMT_ABL_NOLOAD -DSMT_ABL_NOSTORE" ]; }; then
    tag=$(echo "$v" | tr -d ' -' | tr 'A-Z' 'a-z')
    ncu --set full --clock-control none --import-source on -k regex:agg_seq_kernel -s 18 -c 2 -o gpurun_out/prof_abl_$tag -f python scripts/time_agg.py > /dev/null 2>&1
  fi
done
build ""
