#!/bin/bash
# 8 GPUs of one box: the frame-parallel bench line and the row-band split of one 7680x4320x512 frame (BASELINE configs 3, 4)
cd "$(dirname "$0")/.."
N=${1:-8}
WHAT=${2:-all}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$WHAT" != "band" ]; then
$TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 600 gpurun_out/r2_bench_n$N.json | head -c 300; echo
fi
$TR --master-port 29522 bench.py --gpus $N --workload bandsplit --steps 5 --warmup 3 > gpurun_out/r2_bandsplit_${N}gpu.json 2> gpurun_out/r2_bandsplit_${N}gpu.err
grep -o '"value": [0-9.]*' gpurun_out/r2_bandsplit_${N}gpu.json | head -1; grep -o '"share_of_step": [0-9.]*' gpurun_out/r2_bandsplit_${N}gpu.json; grep -o '"parity": {[^}]*}' gpurun_out/r2_bandsplit_${N}gpu.json
