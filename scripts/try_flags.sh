#!/bin/bash
# Development aid: rebuild cross_agg_seq.cu with extra nvcc flags, time one aggregation, restore the product build.
# usage: scripts/try_flags.sh "-DSMT_SEQ_AHEAD=1" ["more flags" ...]
set -u
cd "$(dirname "$0")/.."
CS=stereo_match_traditional_b200/csrc
for v in "$@"; do
  touch $CS/cross_agg_seq.cu $CS/scanline.cu $CS/adcensus.cu $CS/cross_agg.cu; make -C $CS NVCC="/usr/local/cuda/bin/nvcc $v" 2>&1 | grep -E "error"
  echo "flags [$v]: $(python ${TIMER:-scripts/time_agg.py} 2>&1 | tail -1 | cut -c1-420)"
done
touch $CS/cross_agg_seq.cu $CS/scanline.cu $CS/adcensus.cu $CS/cross_agg.cu; make -C $CS 2>&1 | grep -E "error"
