"""ORACLE tooling (test infrastructure only): generate the golden fixtures under tests/golden/.

The reference ships no tests and no golden vectors, so parity is pinned on outputs of the
reference ITSELF: this script drives oracle/_ref/libsmt_ref.so -- the unmodified reference sources compiled with g++
against the OpenCV stand-in -- on seeded synthetic inputs and stores

  * tests/golden/<case>.npz          full input/output arrays of small cases
  * tests/golden/digests.json        sha256 digests of the outputs of larger cases (inputs are regenerated from the
                                     seed by stereo_match_traditional_b200.synth.make_pair)

It needs /root/reference (to build oracle/_ref) and therefore only runs in the build container:

    python -m oracle.make_golden
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402
from stereo_match_traditional_b200.synth import make_pair, replicate_pad  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# (name, H, W, D, kind, seed, store_full_arrays)
ADCENSUS_CASES = [
    ("adc_scene_20x56x12", 20, 56, 12, "scene", 101, True),
    ("adc_noise_16x40x10", 16, 40, 10, "noise", 102, True),
    ("adc_scene_48x200x24", 48, 200, 24, "scene", 7, False),
    ("adc_scene_33x257x40", 33, 257, 40, "scene", 8, False),
    ("adc_const_30x70x32", 30, 70, 32, "constant", 12, False),
    # degenerate shapes: smaller than the census window / one arm, D = 1, D > W, one row, one column
    ("adc_noise_9x9x1", 9, 9, 1, "noise", 118, False),
    ("adc_noise_3x40x6", 3, 40, 6, "noise", 143, False),
    ("adc_noise_20x8x12", 20, 8, 12, "noise", 128, False),
    ("adc_noise_1x50x4", 1, 50, 4, "noise", 151, False),
    ("adc_noise_50x1x1", 50, 1, 1, "noise", 151, False),
    ("adc_noise_35x36x65", 35, 36, 65, "noise", 171, False),
]


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def adcensus_path(o, H, W, D, kind, seed):
    """Every stage of P_full, each fed with the reference's own previous stage."""
    p = make_pair(H, W, D, seed, kind)
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    out = {}
    c = o.adcensus(L, R, D, want=("costL", "costR", "adL", "adR", "censusL", "censusR", "wtaL", "wtaR"))
    out.update({f"adcensus_{k}": v for k, v in c.items()})
    for view, own, oth in (("L", "bgr_left", "bgr_right"), ("R", "bgr_right", "bgr_left")):
        a = o.cross_aggregate(p[own], p[oth], c["cost" + view])
        out[f"agg{view}_cost"] = a["cost"]
        out[f"agg{view}_arms"] = a["arms"]
        out[f"agg{view}_sup0"] = a["sup0"]
        out[f"agg{view}_sup1"] = a["sup1"]
    s = o.scanline(out["aggL_cost"], L, 10, 150, want=("left", "right", "up", "down", "sum", "disp"))
    out.update({f"scan_{k}": v for k, v in s.items()})
    out["dispR"] = o.wta(out["aggR_cost"])
    lrc, occ, mis = o.lrcheck(s["disp"], out["dispR"], 2)
    out["lrc_disp"], out["lrc_occ"], out["lrc_mis"] = lrc, occ, mis
    last, occ2, mis2 = o.lrcheck_and(s["disp"], out["dispR"], 2.0)
    out["lrand_last"], out["lrand_occ"], out["lrand_mis"] = last, occ2, mis2
    out["speckle_inf"] = o.remove_speckles(lrc, 1, 30, None)
    out["speckle_m1"] = o.remove_speckles(s["disp"], 1, 12, -1)
    out["median3"] = o.median(out["speckle_inf"], 3)
    out["median5"] = o.median(s["disp"], 5)
    return out


def main():
    if not orc.REFERENCE_ROOT.is_dir():
        raise SystemExit("needs /root/reference to build oracle/_ref (build container only)")
    orc.build("ref")
    o = orc.Oracle("ref")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    digests = {}
    for name, H, W, D, kind, seed, full in ADCENSUS_CASES:
        out = adcensus_path(o, H, W, D, kind, seed)
        digests[name] = {"shape": [H, W, D], "kind": kind, "seed": seed, "outputs": {k: digest(v) for k, v in out.items()}}
        if full:
            np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    extra = ROOT / "oracle" / "make_golden_matchers.py"
    if extra.exists():
        from oracle import make_golden_matchers

        make_golden_matchers.generate(o, GOLDEN, digests, digest)
    (GOLDEN / "digests.json").write_text(json.dumps(digests, indent=1, sort_keys=True) + "\n")
    print("wrote", sorted(x.name for x in GOLDEN.iterdir()))


if __name__ == "__main__":
    main()
