"""GPU: the opt-in fused pass-pair kernels (SMT_AGG_FUSE = pair | hw | rings, read once per process) give the same
aggregated volume as the default unfused passes.  Each variant runs in its own interpreter."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from stereo_match_traditional_b200.api import Engine, SUM_FAST
from stereo_match_traditional_b200.synth import make_pair
out = {}
with Engine(0) as e:
    for (H, W, D, seed) in [(70, 300, 36, 13), (300, 41, 33, 14), (48, 200, 64, 7)]:
        p = make_pair(H, W, D, seed, "scene")
        cost = e.adcensus_cost(p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32), D)["cost"]
        out[f"{H}x{W}x{D}"] = e.aggregate_cross(p["bgr_left"], cost, mode=SUM_FAST)["cost"]
np.savez(sys.argv[1], **out)
"""


def run_variant(tmp_path, name):
    env = dict(os.environ)
    env.pop("SMT_AGG_FUSE", None)
    env.pop("SMT_AGG_FUSE_V", None)
    if name:
        env["SMT_AGG_FUSE"] = name
    f = tmp_path / f"agg_{name or 'default'}.npz"
    r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT), str(f)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(f)


@pytest.mark.parametrize("variant", ["pair", "hw", "rings"])
def test_fused_pair_kernels_match_unfused(tmp_path, variant):
    base = run_variant(tmp_path, "")
    got = run_variant(tmp_path, variant)
    for k in base.files:
        a, b = base[k], got[k]
        denom = np.maximum(np.abs(a), 1e-30)
        rel = np.where(a == b, 0.0, np.abs(a - b) / denom)
        assert rel.max() <= 2e-6, f"{variant} {k}: max rel diff {rel.max():.3e}"
