#!/bin/bash
# A/B of library variants through bench.py's stage timers: scripts/ab_bench.sh name1 name2 ...  (make -C csrc variant NAME=...)
cd "$(dirname "$0")/.."
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1', round(d['value'],2), round(d['e2e']['value'],2), {k:round(v['ms_per_frame'],3) for k,v in r['all_stages'].items()}, {k:round(v,3) for k,v in r['other_stage_ms_per_frame'].items()})"; }
python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | show default
for v in "$@"; do
  SMT_B200_LIB=$PWD/stereo_match_traditional_b200/csrc/build_var/$v/libsmt_b200.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | show $v
done
