"""One SAD / NCC / ASW run at the BASELINE.json config-3 shape (1280x720, D=128; SAD at 450x375x64) -- the command profiled
with ncu for the matcher kernels' summaries under profiles/."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine
from stereo_match_traditional_b200.synth import make_pair, replicate_pad

dev = "cuda:0"
with Engine(0) as e:
    p = make_pair(375, 450, 64, 1001, "scene")
    lp, rp = (torch.as_tensor(replicate_pad(p[k], 4), device=dev) for k in ("gray_left", "gray_right"))
    for _ in range(2):
        e.sad(lp, rp, 64, 3)
    p = make_pair(720, 1280, 128, 1003, "scene")
    gl, gr = (torch.as_tensor(p[k], device=dev) for k in ("gray_left", "gray_right"))
    lp, rp = (torch.as_tensor(replicate_pad(p[k], 12), device=dev) for k in ("gray_left", "gray_right"))
    for _ in range(2):
        e.ncc(gl, gr, 128, 10)
        e.asw(lp, rp, 128, 11, 50, 30, 40)
    e.sync()
    print("done")
