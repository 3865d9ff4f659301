"""Time SMT_SUM_FAST cross aggregation x4 of one view (device-resident, CUDA events) -- the A/B harness behind the kernel
choices quoted in DESIGN.md.  SMT_B200_LIB selects an alternative build of the library (make -C csrc variant ...)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine, SUM_FAST
from stereo_match_traditional_b200.synth import make_pair

H, W, D = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1080x1920x256").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
p = make_pair(H, W, D, 1004, "scene")
st = torch.cuda.Stream(device=dev)
with Engine(0, st.cuda_stream) as e, torch.cuda.stream(st):
    gl = torch.as_tensor(p["gray_left"].astype(np.float32), device=dev)
    gr = torch.as_tensor(p["gray_right"].astype(np.float32), device=dev)
    bgr = torch.as_tensor(p["bgr_left"], device=dev)
    cost = e.adcensus_cost(gl, gr, D)["cost"]
    for _ in range(3):
        e.aggregate_cross(bgr, cost, mode=SUM_FAST)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        e.aggregate_cross(bgr, cost, mode=SUM_FAST)
    e1.record(st)
    torch.cuda.synchronize()
    # aggregate_cross also builds the arms, copies the input volume and measures its maximum: time those alone and subtract
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(st)
    for _ in range(reps):
        e.aggregate_cross(bgr, cost, iters=0, mode=SUM_FAST)
    e3.record(st)
    torch.cuda.synchronize()
    tot, base = e0.elapsed_time(e1) / reps, e2.elapsed_time(e3) / reps
    print(f"{os.environ.get('SMT_B200_LIB', 'default'):60s} aggregate x4, one view: {tot - base:.3f} ms  (call {tot:.3f}, overhead {base:.3f}) redone {e.aggregate_fast_status()}")
