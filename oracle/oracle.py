"""ORACLE loader -- test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module, and only as the checker / reported CPU baseline.  The product package
(``stereo_match_traditional_b200``) never imports it.

Two interchangeable back ends with the same Python surface:

* ``kind="ref"``  -> ``oracle/_ref/libsmt_ref.so``: the UNMODIFIED reference sources from /root/reference
  compiled with g++ against the OpenCV stand-in (``oracle/ref_harness.cpp``, symbols ``ref_*``).
* ``kind="port"`` -> ``oracle/libsmt_oracle.so``: the plain-C restatement (``oracle/port/oracle_port.c``,
  symbols ``orc_*``), pinned against the compiled reference by ``tests/test_oracle_port_vs_ref.py`` and by the
  golden fixtures under ``tests/golden/`` (generated from the compiled reference by ``oracle/make_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libsmt_ref.so"
REF_OMP_SO = HERE / "_ref" / "libsmt_ref_omp.so"
PORT_SO = HERE / "libsmt_oracle.so"
REFERENCE_ROOT = Path(os.environ.get("SMT_REFERENCE_ROOT", "/root/reference"))

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i32p = C.POINTER(C.c_int32)


def build(what: str = "all") -> None:
    """Build the oracle libraries.  ``ref`` needs /root/reference and is skipped when it is absent."""
    targets = []
    if what in ("all", "port"):
        targets.append("port")
    if what in ("all", "ref") and REFERENCE_ROOT.is_dir():
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-C", str(HERE), f"REF_ROOT={REFERENCE_ROOT}"] + targets, check=True,
                       stdout=subprocess.DEVNULL)


def have(kind: str) -> bool:
    return {"ref": REF_SO, "ref_omp": REF_OMP_SO, "port": PORT_SO}[kind].exists()


def _p(a, typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Oracle:
    """numpy-in / numpy-out wrapper over one oracle library."""

    def __init__(self, kind: str = "ref"):
        path = {"ref": REF_SO, "ref_omp": REF_OMP_SO, "port": PORT_SO}[kind]
        if not path.exists():
            raise FileNotFoundError(f"oracle library {path} not built; run oracle.build() / make -C oracle")
        self.kind = kind
        self.prefix = "orc_" if kind == "port" else "ref_"
        self.lib = C.CDLL(str(path))
        # symbols are attached only when the library exports them (the C port grows stage by stage)
        def sig(name, argtypes=None, restype=None):
            try:
                fn = self._fn(name)
            except AttributeError:
                return
            if argtypes is not None:
                fn.argtypes = argtypes
            if restype is not None:
                fn.restype = restype

        sig("remove_speckles", [_f32p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int])
        sig("ncc_cost", restype=C.c_double)
        sig("sad_optimal_disparity", restype=C.c_float)
        sig("asw", [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _f32p])
        sig("asw_masks", [C.c_int, C.c_double, C.c_double, _f64p, _f64p])
        sig("adcensus", [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int] + [_f32p] * 8)
        sig("lrcheck_and", [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, _i32p, _i32p, _i32p, _i32p])

    def has(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def num_threads(self) -> int:
        return int(self._fn("num_threads")())

    # ------------------------------------------------------------------ AD_Census
    def adcensus(self, left, right, D, sigmaC=10.0, sigmaS=30.0, want=("costL", "costR")):
        """AD_Census::{ComputeADcensus, ComputeADcensusRight, WTA} (AD-Census.h:271,296,346)."""
        left = _c(left, np.float32)
        right = _c(right, np.float32)
        H, W = left.shape
        names = ["costL", "costR", "adL", "adR", "censusL", "censusR", "wtaL", "wtaR"]
        out = {}
        for n in names:
            if n in want:
                out[n] = np.empty((H, W) if n.startswith("wta") else (H, W, D), np.float32)
        if ("wtaL" in want) != ("wtaR" in want):
            raise ValueError("wtaL and wtaR come together (AD_Census::WTA computes both)")
        do_left = int(any(n in want for n in ("costL", "adL", "censusL", "wtaL")))
        do_right = int(any(n in want for n in ("costR", "adR", "censusR", "wtaR")))
        self._fn("adcensus")(_p(left, _f32p), _p(right, _f32p), H, W, D, sigmaC, sigmaS, do_left, do_right,
                             *[_p(out.get(n), _f32p) for n in names])
        return out

    def wta(self, vol):
        vol = _c(vol, np.float32)
        H, W, D = vol.shape
        disp = np.empty((H, W), np.float32)
        self._fn("wta")(_p(vol, _f32p), H, W, D, _p(disp, _f32p))
        return disp

    # ------------------------------------------------------------------ CrossArmAggregation
    def crossarm(self, img, D=1, tao=30, mode=0, vol_in=None):
        """Initialize + the four arm passes (CrossArm.cpp:147-598) [+ AggregationVertical (mode 1) / Aggregation (2)]."""
        img = _c(img, np.uint8)
        ch = 3 if img.ndim == 3 else 1
        H, W = img.shape[:2]
        arms = [np.empty((H, W), np.int32) for _ in range(4)]
        vol_out = None
        if mode:
            vol_in = _c(vol_in, np.float32)
            D = vol_in.shape[2]
            vol_out = np.empty_like(vol_in)
        f = self._fn("crossarm")
        f.restype = C.c_int
        tao_final = f(_p(img, _u8p), ch, H, W, D, tao, *[_p(a, _i32p) for a in arms], mode, _p(vol_in, _f32p),
                      _p(vol_out, _f32p))
        return {"left": arms[0], "right": arms[1], "top": arms[2], "bottom": arms[3], "vol": vol_out,
                "tao_final": int(tao_final)}

    # ------------------------------------------------------------------ ScanlineOptimizer
    def scanline(self, cost, gray, p1=10, p2=150, want=("sum", "disp")):
        cost = _c(cost, np.float32)
        gray = _c(gray, np.float32)
        H, W, D = cost.shape
        names = ["left", "right", "up", "down", "sum", "disp"]
        out = {n: np.empty((H, W) if n == "disp" else (H, W, D), np.float32) for n in names if n in want}
        self._fn("scanline")(_p(cost, _f32p), _p(gray, _f32p), H, W, D, int(p1), int(p2),
                             *[_p(out.get(n), _f32p) for n in names])
        return out

    # ------------------------------------------------------------------ PostProcessing
    def lrcheck(self, dispL, dispR, gate=2):
        """LeftRightConsistency (PostProcessing.h:72). Returns (dispL_out, occlusions[n,2], mismatches[m,2])."""
        d = np.array(dispL, np.float32, order="C", copy=True)
        r = _c(dispR, np.float32)
        H, W = d.shape
        occ = np.empty((H * W, 2), np.int32)
        mis = np.empty((H * W, 2), np.int32)
        no, nm = C.c_int32(0), C.c_int32(0)
        self._fn("lrcheck")(_p(d, _f32p), _p(r, _f32p), W, H, int(gate), _p(occ, _i32p), C.byref(no), _p(mis, _i32p),
                            C.byref(nm))
        return d, occ[: no.value].copy(), mis[: nm.value].copy()

    def lrcheck_and(self, dispL, dispR, gate=2.0):
        """LeftAndRightConsistency (PostProcessing.h:10). Returns (lastDisp, occlusions, mismatches)."""
        l = _c(dispL, np.float32)
        r = _c(dispR, np.float32)
        H, W = l.shape
        last = np.zeros((H, W), np.float32)
        occ = np.empty((H * W, 2), np.int32)
        mis = np.empty((H * W, 2), np.int32)
        no, nm = C.c_int32(0), C.c_int32(0)
        self._fn("lrcheck_and")(_p(l, _f32p), _p(r, _f32p), _p(last, _f32p), W, H, float(gate), _p(occ, _i32p),
                                C.byref(no), _p(mis, _i32p), C.byref(nm))
        return last, occ[: no.value].copy(), mis[: nm.value].copy()

    def remove_speckles(self, disp, diff=1, min_area=30, invalid=None):
        """RemoveSpeckles (PostProcessing.h:250). invalid=None reproduces main.cpp:93 (passes +inf to an int param)."""
        d = np.array(disp, np.float32, order="C", copy=True)
        H, W = d.shape
        if invalid is None:
            self._fn("remove_speckles_inf")(_p(d, _f32p), W, H, int(diff), C.c_uint(min_area))
        else:
            self._fn("remove_speckles")(_p(d, _f32p), W, H, int(diff), C.c_uint(min_area), int(invalid))
        return d

    def median(self, disp, k=3):
        d = _c(disp, np.float32)
        H, W = d.shape
        out = np.empty_like(d)
        self._fn("median")(_p(d, _f32p), _p(out, _f32p), W, H, int(k))
        return out

    def fill_the_hole(self, disp, disp_range, occ, mis, fixed=False):
        """FillTheHole (PostProcessing.h:156) as written.  Returns (disp_out, mismatch_list_after).
        fixed=True (port only): axes not swapped, invalid = +inf -- not a reference behaviour."""
        d = np.array(disp, np.float32, order="C", copy=True)
        row, col = d.shape
        o = np.ascontiguousarray(occ, np.int32).reshape(-1, 2)
        cap = max(row * col, len(mis)) + 1
        m = np.zeros((cap, 2), np.int32)
        m[: len(mis)] = np.asarray(mis, np.int32).reshape(-1, 2)
        nm = C.c_int32(len(mis))
        self._fn("fill_the_hole_fixed" if fixed else "fill_the_hole")(row, col, int(disp_range), _p(d, _f32p), _p(o, _i32p), len(o), _p(m, _i32p), C.byref(nm), cap)
        return d, m[: nm.value].copy()

    # ------------------------------------------------------------------ ASW display chain (ASWeight.cpp:66-78)
    def normalize_minmax_u8_from_u8(self, img):
        """cv::normalize(img, img, 0, 255, NORM_MINMAX) on CV_8U (port only; pinned to cv2 by tests/golden/display.npz)."""
        a = _c(img, np.uint8)
        out = np.empty_like(a)
        f = self._fn("normalize_minmax_u8_from_u8")
        f.restype = None
        f(_p(a, _u8p), a.shape[0], a.shape[1], _p(out, _u8p))
        return out

    def filter_speckles_u8(self, img, new_val=0, max_size=40, max_diff=2):
        """cv::filterSpeckles on CV_8UC1 (port only; pinned to cv2)."""
        a = np.array(img, np.uint8, order="C")
        f = self._fn("filter_speckles_u8")
        f.restype = None
        f(_p(a, _u8p), a.shape[0], a.shape[1], int(new_val), int(max_size), int(max_diff))
        return a

    def median_blur_u8(self, img, k):
        """cv::medianBlur on CV_8UC1, ksize 3 / 5 (port only; pinned to cv2)."""
        a = _c(img, np.uint8)
        out = np.empty_like(a)
        f = self._fn("median_blur_u8")
        f.restype = None
        f(_p(a, _u8p), _p(out, _u8p), a.shape[0], a.shape[1], int(k))
        return out

    def fill_image_new(self, img, past_end=0):
        """FillImageNew (ASW.h:434) as written.  Returns (image, n_unassigned): the last n_unassigned zero pixels of the
        raster order read past the reference's value vector and must not be compared (the port leaves them 0).
        past_end: the byte the reference finds one past the image (it reads it)."""
        a = np.array(img, np.uint8, order="C")
        H, W = a.shape
        if self.prefix == "ref_":
            buf = np.zeros(H * W + 1, np.uint8)
            buf[:-1] = a.ravel()
            buf[-1] = past_end
            f = self._fn("fill_image_new")
            f.restype = None
            f(_p(buf, _u8p), H, W)
            return buf[:-1].reshape(H, W).copy(), None
        n = C.c_int32(0)
        f = self._fn("fill_image_new")
        f.restype = None
        f(_p(a, _u8p), H, W, int(past_end), C.byref(n))
        return a, int(n.value)

    # ------------------------------------------------------------------ refinement (fixed specification, port only)
    @staticmethod
    def refine_directions():
        """sin / cos tables of the 16 interpolation directions, evaluated in float32 (shared by the port and the kernels)."""
        libm = C.CDLL("libm.so.6")   # the same sinf / cosf the library's host code calls
        for fn in (libm.sinf, libm.cosf):
            fn.restype, fn.argtypes = C.c_float, [C.c_float]
        ang = [np.float32(k) * np.float32(np.pi / 8) for k in range(16)]
        return (np.array([libm.sinf(C.c_float(a)) for a in ang], np.float32), np.array([libm.cosf(C.c_float(a)) for a in ang], np.float32))

    def refine_region_voting(self, disp, cls, arms, D, ts=20, th=0.4, iters=5):
        d = np.array(disp, np.float32, order="C")
        c = np.array(cls, np.uint8, order="C")
        arms = _c(arms, np.uint8)
        H, W = d.shape
        f = self._fn("refine_region_voting")
        f.restype = None
        f(_p(d, _f32p), _p(c, _u8p), _p(arms, _u8p), H, W, int(D), int(ts), C.c_float(th), int(iters))
        return d, c

    def refine_interpolate(self, disp, cls, bgr, max_len):
        d = np.array(disp, np.float32, order="C")
        c = np.array(cls, np.uint8, order="C")
        bgr = _c(bgr, np.uint8)
        H, W = d.shape
        sina, cosa = self.refine_directions()
        f = self._fn("refine_interpolate")
        f.restype = None
        f(_p(d, _f32p), _p(c, _u8p), _p(bgr, _u8p), H, W, int(max_len), _p(sina, _f32p), _p(cosa, _f32p))
        return d, c

    def refine_discontinuity(self, disp, cost, thres=5.0):
        d = np.array(disp, np.float32, order="C")
        cost = _c(cost, np.float32)
        H, W, D = cost.shape
        f = self._fn("refine_discontinuity")
        f.restype = None
        f(_p(d, _f32p), _p(cost, _f32p), H, W, D, C.c_float(thres))
        return d

    def transform_to_show(self, disp):
        """TransformToShow (PostProcessing.h:137)."""
        d = _c(disp, np.float32)
        H, W = d.shape
        out = np.zeros((H, W), np.uint8)
        self._fn("transform_to_show")(_p(d, _f32p), H, W, _p(out, _u8p))
        return out

    # ------------------------------------------------------------------ CrossAggregator
    def cross_aggregate(self, bgrL, bgrR, cost, L1=34, L2=17, t1=20, t2=6, iters=4):
        bgrL = _c(bgrL, np.uint8)
        bgrR = _c(bgrR, np.uint8)
        cost = _c(cost, np.float32)
        H, W, D = cost.shape
        out = np.empty_like(cost)
        arms = np.empty((H, W, 4), np.uint8)
        sup0 = np.empty((H, W), np.uint16)
        sup1 = np.empty((H, W), np.uint16)
        f = self._fn("cross_aggregate")
        f.restype = C.c_int
        ok = f(_p(bgrL, _u8p), _p(bgrR, _u8p), _p(cost, _f32p), W, H, D, L1, L2, t1, t2, iters, _p(out, _f32p),
               _p(arms, _u8p), _p(sup0, _u16p), _p(sup1, _u16p))
        if not ok:
            raise RuntimeError("CrossAggregator::Initialize returned false")
        return {"cost": out, "arms": arms, "sup0": sup0, "sup1": sup1}

    # ------------------------------------------------------------------ CBLSM
    def cblsm_ad(self, left, right, D):
        left = _c(left, np.uint8)
        right = _c(right, np.uint8)
        H, W = left.shape
        adL = np.zeros((H, W, D), np.float32)
        adR = np.zeros((H, W, D), np.float32)
        self._fn("cblsm_ad")(_p(left, _u8p), _p(right, _u8p), H, W, D, _p(adL, _f32p), _p(adR, _f32p))
        return adL, adR

    def cblsm_arms(self, img, tao=25, max_length=34, sec_length=17):
        img = _c(img, np.uint8)
        ch = 3 if img.ndim == 3 else 1
        H, W = img.shape[:2]
        arms = [np.zeros((H, W), np.int32) for _ in range(4)]
        self._fn("cblsm_arms")(_p(img, _u8p), ch, H, W, tao, max_length, sec_length, *[_p(a, _i32p) for a in arms])
        return {"left": arms[0], "right": arms[1], "up": arms[2], "down": arms[3]}

    def cblsm_aggregate_v5(self, vol, arms):
        vol = _c(vol, np.float32)
        H, W, D = vol.shape
        out = np.empty_like(vol)
        a = [_c(arms[k], np.int32) for k in ("left", "right", "up", "down")]
        self._fn("cblsm_aggregate_v5")(_p(vol, _f32p), _p(out, _f32p), *[_p(x, _i32p) for x in a], D, H, W)
        return out

    def cblsm_aggregate_variant(self, vol, arms, which):
        """which = "vertical" (costAggregationVertical, per-pixel arm maps) or "v4" (costAggregationV4, arm VOLUMES [H][W][D])."""
        vol = _c(vol, np.float32)
        H, W, D = vol.shape
        out = np.empty_like(vol)
        a = [_c(arms[k], np.int32) for k in ("left", "right", "up", "down")]
        self._fn("cblsm_aggregate_" + which)(_p(vol, _f32p), _p(out, _f32p), *[_p(x, _i32p) for x in a], D, H, W)
        return out

    def cblsm_wta(self, vol):
        vol = _c(vol, np.float32)
        H, W, D = vol.shape
        disp = np.empty((H, W), np.float32)
        self._fn("cblsm_wta")(_p(vol, _f32p), _p(disp, _f32p), D, H, W)
        return disp

    # ------------------------------------------------------------------ SAD
    def sad(self, left_padded, right_padded, D, winsize=3, right_view=False):
        lp = _c(left_padded, np.uint8)
        rp = _c(right_padded, np.uint8)
        Hp, Wp = lp.shape
        pad = winsize + 1
        disp = np.zeros((Hp - 2 * pad, Wp - 2 * pad), np.int32)
        self._fn("sad_right" if right_view else "sad_left")(_p(lp, _u8p), _p(rp, _u8p), Hp, Wp, D, winsize,
                                                             _p(disp, _i32p))
        return disp

    def sad_costs_left(self, left_padded, right_padded, D, winsize=3):
        lp = _c(left_padded, np.uint8)
        rp = _c(right_padded, np.uint8)
        Hp, Wp = lp.shape
        pad = winsize + 1
        costs = np.zeros((Hp - 2 * pad, Wp - 2 * pad, D), np.float32)
        self._fn("sad_costs_left")(_p(lp, _u8p), _p(rp, _u8p), Hp, Wp, D, winsize, _p(costs, _f32p))
        return costs

    def sad_optimal_disparity(self, sad):
        s = _c(sad, np.float32)
        return float(self._fn("sad_optimal_disparity")(_p(s, _f32p), s.shape[0]))

    # ------------------------------------------------------------------ NCC
    def sad_crosscheck(self, dispL, dispR):
        l, r = _c(dispL, np.int32), _c(dispR, np.int32)
        H, W = l.shape
        out = np.empty((H, W), np.int32)
        occ, mis = np.empty((H * W, 2), np.int32), np.empty((H * W, 2), np.int32)
        no, nm = C.c_int32(0), C.c_int32(0)
        f = self._fn("sad_crosscheck")
        f.restype = None
        f(_p(l, _i32p), _p(r, _i32p), H, W, _p(out, _i32p), _p(occ, _i32p), C.byref(no), _p(mis, _i32p), C.byref(nm))
        return out, occ[: no.value].copy(), mis[: nm.value].copy()

    def sad_subpixel(self, sad):
        """Port only: the value OptimalDisparity computes and discards (Sad.h:76-84)."""
        sad = _c(sad, np.float32)
        f = self._fn("sad_subpixel")
        f.restype = C.c_float
        return float(f(_p(sad, _f32p), int(sad.size)))

    def ncc_shift(self, left, right, right_view=False, add_constant=False):
        left, right = _c(left, np.uint8), _c(right, np.uint8)
        H, W = left.shape
        out = np.empty((H, W), np.uint8)
        f = self._fn("ncc_shift")
        f.restype = None
        f(_p(left, _u8p), _p(right, _u8p), H, W, int(bool(right_view)), int(bool(add_constant)), _p(out, _u8p))
        return out

    def ncc(self, left, right, D, win_size=10):
        left = _c(left, np.uint8)
        right = _c(right, np.uint8)
        H, W = left.shape
        disp = np.zeros((H, W), np.int32)
        self._fn("ncc")(_p(left, _u8p), _p(right, _u8p), H, W, win_size, D, _p(disp, _i32p))
        return disp

    def ncc_costs(self, left, right, D, win_size=10, rows=None):
        """ComputeCost per (pixel, d) for output rows [rows[0], rows[1]) -- float64 [n][W][D] (compiled reference only)."""
        left = _c(left, np.uint8)
        right = _c(right, np.uint8)
        H, W = left.shape
        i0, i1 = rows if rows is not None else (0, H)
        out = np.zeros((i1 - i0, W, D), np.float64)
        f = self._fn("ncc_costs")
        f.restype = None
        f(_p(left, _u8p), _p(right, _u8p), H, W, int(win_size), int(D), int(i0), int(i1), _p(out, _f64p))
        return out

    def asw_costs(self, left_padded, right_padded, D, win_size=11, space_sigma=50.0, color_sigma=30.0, T=40, rows=None):
        """costVolume rows of AdaptiveSupportWeight for output rows [rows[0], rows[1]) -- float32 [n][Wo][D]."""
        lp = _c(left_padded, np.uint8)
        rp = _c(right_padded, np.uint8)
        Hp, Wp = lp.shape
        pad = win_size + 1
        i0, i1 = rows if rows is not None else (0, Hp - 2 * pad)
        out = np.zeros((i1 - i0, Wp - 2 * pad, D), np.float32)
        f = self._fn("asw_costs")
        f.restype = None
        f(_p(lp, _u8p), _p(rp, _u8p), Hp, Wp, int(win_size), int(D), C.c_double(space_sigma), C.c_double(color_sigma), int(T),
          int(i0), int(i1), _p(out, _f32p))
        return out

    def ncc_cost(self, left, right, win_size, i, j, d):
        left = _c(left, np.uint8)
        right = _c(right, np.uint8)
        H, W = left.shape
        return float(self._fn("ncc_cost")(_p(left, _u8p), _p(right, _u8p), H, W, win_size, i, j, d))

    # ------------------------------------------------------------------ ASW
    def asw(self, left_padded, right_padded, D, win_size=11, space_sigma=50.0, color_sigma=30.0, T=40,
            right_view=False):
        lp = _c(left_padded, np.uint8)
        rp = _c(right_padded, np.uint8)
        Hp, Wp = lp.shape
        pad = win_size + 1
        disp = np.zeros((Hp - 2 * pad, Wp - 2 * pad), np.float32)
        self._fn("asw")(_p(lp, _u8p), _p(rp, _u8p), Hp, Wp, win_size, D, float(space_sigma), float(color_sigma),
                        int(T), int(bool(right_view)), _p(disp, _f32p))
        return disp

    def asw_masks(self, win_size=11, space_sigma=50.0, color_sigma=30.0):
        k = 2 * win_size + 3
        space = np.empty((k, k), np.float64)
        color = np.empty(256, np.float64)
        self._fn("asw_masks")(win_size, float(space_sigma), float(color_sigma), _p(space, _f64p), _p(color, _f64p))
        return space, color

    def asw_crosscheck(self, dispL, dispR):
        l = _c(dispL, np.float32)
        r = _c(dispR, np.float32)
        H, W = l.shape
        out = np.zeros((H, W), np.uint8)
        self._fn("asw_crosscheck")(_p(l, _f32p), _p(r, _f32p), H, W, _p(out, _u8p))
        return out
