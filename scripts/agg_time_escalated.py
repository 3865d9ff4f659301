"""Time SMT_SUM_FAST cross aggregation x4 of an integer AD volume (CBLSM's costs): the fixed-point sweeps escalate after the
first iteration and the float-pair passes take whole sweeps -- the regime the row-band split of config 5 runs in."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine, SUM_FAST
from stereo_match_traditional_b200.synth import make_pair

H, W, D = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1080x1920x256").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
p = make_pair(H, W, D, 1004, "scene")
st = torch.cuda.Stream(device=dev)
with Engine(0, st.cuda_stream) as e, torch.cuda.stream(st):
    gl = torch.as_tensor(p["gray_left"], device=dev)
    gr = torch.as_tensor(p["gray_right"], device=dev)
    bgr = torch.as_tensor(p["bgr_left"], device=dev)
    cost = e.ad_cost(gl, gr, D)
    cost = cost[0] if isinstance(cost, (tuple, list)) else cost
    for _ in range(2):
        e.aggregate_cross(bgr, cost, mode=SUM_FAST)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        e.aggregate_cross(bgr, cost, mode=SUM_FAST)
    e1.record(st)
    torch.cuda.synchronize()
    print(f"{os.environ.get('SMT_B200_LIB', 'default'):60s} escalating aggregate x4, one view: {e0.elapsed_time(e1) / reps:.3f} ms per call, status {e.aggregate_fast_status()}")
