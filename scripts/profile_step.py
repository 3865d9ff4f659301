"""One warm-up frame + one profiled frame of P_full (the command ncu wraps; see profiles/README.md)."""
import sys
sys.path.insert(0, ".")
import torch
from stereo_match_traditional_b200.api import Engine, SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.synth import make_pair

H, W, D = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1080x1920x256").split("x"))
mode = SUM_FAST if (len(sys.argv) < 3 or sys.argv[2] == "fast") else SUM_REFERENCE_ORDER
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
e = Engine(0)
p = make_pair(H, W, D, 1004, "scene")
t = {k: torch.as_tensor(p[k], device="cuda:0") for k in ("bgr_left", "bgr_right", "gray_left", "gray_right")}
for _ in range(frames):
    e.pipeline(t["bgr_left"], t["bgr_right"], D, t["gray_left"], t["gray_right"], sum_mode=mode)
e.sync()
print("launches", e.kernel_launches)
e.close()
