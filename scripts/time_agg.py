"""Ad-hoc: time CrossAggregator::Aggregate(4) alone at 1080p x 256 (development aid)."""
import sys, json
sys.path.insert(0, ".")
import torch
from stereo_match_traditional_b200.api import Engine, SUM_FAST
from stereo_match_traditional_b200.synth import make_pair

H, W, D = 1080, 1920, 256
e = Engine(0)
p = make_pair(H, W, D, 1004, "scene")
bl = torch.as_tensor(p["bgr_left"], device="cuda:0")
gl = torch.as_tensor(p["gray_left"], device="cuda:0").float()
gr = torch.as_tensor(p["gray_right"], device="cuda:0").float()
cost = e.adcensus_cost(gl, gr, D)["cost"]
for _ in range(2):
    e.aggregate_cross(bl, cost, mode=SUM_FAST)
e.sync()
e.stage_timing(True)
reps = 5
for _ in range(reps):
    e.aggregate_cross(bl, cost, mode=SUM_FAST)
st = e.read_stage_timing()
print(json.dumps({"aggregate_ms_per_view": st["aggregate"] / reps, "per_pass_ms": st["aggregate"] / reps / 8}))
