"""Ad-hoc timing of P_full with per-stage CUDA-event times (development aid, not the benchmark)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from stereo_match_traditional_b200.api import Engine, SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.synth import make_pair

def run(H, W, D, mode, reps=3):
    e = Engine(0)
    p = make_pair(H, W, D, 1004, "scene")
    dev = "cuda:0"
    bl = torch.as_tensor(p["bgr_left"], device=dev); br = torch.as_tensor(p["bgr_right"], device=dev)
    gl = torch.as_tensor(p["gray_left"], device=dev); gr = torch.as_tensor(p["gray_right"], device=dev)
    e.pipeline(bl, br, D, gl, gr, sum_mode=mode); e.sync()
    e.stage_timing(True)
    t0 = time.time()
    for _ in range(reps):
        e.pipeline(bl, br, D, gl, gr, sum_mode=mode)
    e.sync()
    dt = (time.time() - t0) / reps
    st = e.read_stage_timing()
    print(json.dumps({"shape": [H, W, D], "mode": mode, "ms": dt * 1e3, "fps": 1 / dt,
                      "stages_ms": {k: v / reps for k, v in st.items()}, "launches": e.kernel_launches}))
    e.close()

if __name__ == "__main__":
    modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1]
    for m in modes:
        run(375, 450, 64, m)
        run(1080, 1920, 256, m, reps=2)
