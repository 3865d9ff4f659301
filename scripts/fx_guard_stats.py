"""Where does the domain guard of the fixed-point prefixes trip on the benchmark frames?  (GPU, torch float64 emulation of
the pass chain on a subset of disparities; not product code)"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine
from stereo_match_traditional_b200.synth import make_pair

H, W, D = 1080, 1920, 256
dev = "cuda:0"
dsel = torch.tensor([0, 5, 19, 20, 21, 43, 44, 45, 60, 100, 150, 200, 255], device=dev)
q = 2.0 ** -24

def wsum(t, lo, hi, axis):
    P = torch.cumsum(t, dim=axis)
    z = torch.zeros_like(P.narrow(axis, 0, 1))
    E = torch.cat([z, P], dim=axis)
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    if axis == 1:
        return E[yy, xx + hi + 1] - E[yy, xx - lo]
    return E[yy + hi + 1, xx] - E[yy - lo, xx]

with Engine(0) as e:
    for frame in (0, 1, 2):
        p = make_pair(H, W, D, 1004 + frame, "scene", frame=frame)
        gl = torch.as_tensor(p["gray_left"].astype(np.float32), device=dev)
        gr = torch.as_tensor(p["gray_right"].astype(np.float32), device=dev)
        for view in (0, 1):
            bgr = torch.as_tensor(p["bgr_left" if view == 0 else "bgr_right"], device=dev)
            cost = e.adcensus_cost(gl, gr, D, 10.0, 30.0, view)["cost"][:, :, dsel].double().contiguous()
            m = e.cross_arms(bgr)
            arms = m["arms"].long()
            L, R, T, B = arms[..., 0], arms[..., 1], arms[..., 2], arms[..., 3]
            cur = cost
            for k in range(4):
                hf = k % 2 == 0
                sup = (m["sup0"] if hf else m["sup1"]).long()
                lo1, hi1, ax1 = (L, R, 1) if hf else (T, B, 0)
                lo2, hi2, ax2 = (T, B, 0) if hf else (L, R, 1)
                s1 = wsum(cur, lo1, hi1, ax1)
                n1 = (lo1 + hi1 + 1)[..., None].double()
                r1 = torch.where(s1 > 0, s1 / q / (n1 * 65536.0), torch.full_like(s1, 1e30))
                s2 = wsum(s1, lo2, hi2, ax2)
                n2 = (lo2 + hi2 + 1)[..., None].double()
                coarse = (sup >= 128)[..., None]
                r2 = torch.where(coarse, s2 / (64 * q) / 131072.0, s2 / q / (n2 * 65536.0))
                r2 = torch.where(s2 > 0, r2, torch.full_like(s2, 1e30))
                cur = s2 / sup[..., None].double()
                def where(r):
                    i = int(torch.argmin(r)); y, x, dd = np.unravel_index(i, r.shape)
                    return f"min ratio {float(r.min()):.3f} at (y={y},x={x},d={int(dsel[dd])}) below1: {int((r < 1).sum())} below .5: {int((r < .5).sum())} below .25: {int((r < .25).sum())}"
                print(f"frame {frame} view {view} iter {k}: unprimed {where(r1)} n={int(n1.flatten()[0])}")
                print(f"frame {frame} view {view} iter {k}:   primed {where(r2)}  | min nonzero mean {float(cur[cur > 0].min()):.3e}", flush=True)
            del cost, cur
