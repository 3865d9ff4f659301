"""Host-side mirror of the reference's per-method entry points, on top of the C ABI.

``Engine`` wraps one ``smt_ctx`` (one per GPU / stream) and exposes every C-ABI call with numpy arrays (host
memspace: the reference's own convention -- caller-owned host arrays) or torch CUDA tensors (device memspace:
nothing crosses PCIe).  The classes and functions below it keep the reference's names, argument order and
meaning so call sites and tests read like the reference's:

    AD_Census            AD-CensusV1/AD-Census.h:9-43
    CrossAggregator      CBLSM/cross_aggregator.h:27-113
    ScanlineOptimizer    AD-CensusV1/ScanlineOptimizer.h:8-34
    LeftRightConsistency, LeftAndRightConsistency, RemoveSpeckles, MedianFilter   AD-CensusV1/PostProcessing.h

PyTorch is used only for device memory and streams; all compute is in libsmt_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import SMT_DEVICE, SMT_HOST, SUM_FAST, SUM_REFERENCE_ORDER, AdCensusParams

__all__ = [
    "Engine", "SmtError", "default_engine", "AD_Census", "CrossAggregator", "ScanlineOptimizer",
    "LeftRightConsistency", "LeftAndRightConsistency", "RemoveSpeckles", "MedianFilter", "Invalid_Float",
    "SUM_FAST", "SUM_REFERENCE_ORDER",
]

Invalid_Float = float("inf")  # PostProcessing.h:7
INVALID_INT_FROM_INF = 2147483647  # what `const int& invalid_val = Invalid_Float` becomes under g++ (main.cpp:93)


class SmtError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"smt status {status}: {text}")
        self.status = status


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Args:
    """Collects arguments of one call, decides the memspace, keeps temporaries alive."""

    def __init__(self, engine: "Engine", *inputs):
        self.engine = engine
        self.device = any(_is_torch(a) and a.is_cuda for a in inputs if a is not None)
        self.keep = []
        engine._device_call = self.device   # the next library call is ordered against torch's current stream

    @property
    def space(self) -> int:
        return SMT_DEVICE if self.device else SMT_HOST

    def inp(self, a, dtype, shape=None):
        if a is None:
            return None
        if self.device:
            import torch

            tdt = getattr(torch, np.dtype(dtype).name)
            if not _is_torch(a):
                a = torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=f"cuda:{self.engine.device}")
            if a.dtype != tdt:
                raise TypeError(f"expected {tdt}, got {a.dtype}")
            a = a.contiguous()
            self.keep.append(a)
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise ValueError(f"expected shape {shape}, got {tuple(a.shape)}")
            return C.c_void_p(a.data_ptr())
        a = np.ascontiguousarray(a, dtype=dtype)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"expected shape {shape}, got {a.shape}")
        self.keep.append(a)
        return C.c_void_p(a.ctypes.data)

    def out(self, shape, dtype, want=True, init=None):
        """Allocate an output (numpy or torch).  Returns (array, pointer)."""
        if not want:
            return None, None
        if self.device:
            import torch

            tdt = {np.uint64: torch.int64}.get(dtype, None) or getattr(torch, np.dtype(dtype).name)
            t = torch.empty(tuple(shape), dtype=tdt, device=f"cuda:{self.engine.device}")
            if init is not None:
                t.copy_(init if _is_torch(init) else torch.as_tensor(np.asarray(init)))
            self.keep.append(t)
            return t, C.c_void_p(t.data_ptr())
        a = np.empty(tuple(shape), dtype=dtype)
        if init is not None:
            a[...] = init
        self.keep.append(a)
        return a, C.c_void_p(a.ctypes.data)


class _OrderedLib:
    """The ctypes library as the Engine's methods see it.  A call that takes torch CUDA tensors runs on the context's
    stream, which is in general not the stream torch is using: the call is queued behind the work torch's current stream
    already holds (the producers of the inputs) and torch's current stream is then made to wait for it (the consumers of the
    outputs, and the caching allocator's reuse of freed blocks).  Everything else is passed through untouched."""

    def __init__(self, engine: "Engine", lib):
        self._engine, self._lib = engine, lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        eng = self._engine

        def call(*args):
            if not eng._device_call:      # set per method call by _Args; stays set for every library call of that method
                return fn(*args)
            import torch

            cur = torch.cuda.current_stream(eng.device)
            ext = eng.torch_stream
            if cur.cuda_stream == ext.cuda_stream:
                return fn(*args)
            ext.wait_stream(cur)
            try:
                return fn(*args)
            finally:
                cur.wait_stream(ext)

        setattr(self, name, call)
        return call


class Engine:
    """One ``smt_ctx``: a device, a stream, a cached device workspace."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._device_call = False
        self._torch_stream = None
        self._raw = _lib.load()
        self.lib = _OrderedLib(self, self._raw)
        self.device = int(device)
        h = C.c_void_p()
        st = self.lib.smt_create(self.device, C.c_void_p(stream) if stream else None, C.byref(h))
        if st != 0:
            raise SmtError(st, "smt_create failed (no CUDA device? this library has no CPU fallback): "
                           + self.lib.smt_status_string(st).decode())
        self.h = h

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self._device_call = False
            self._torch_stream = None
            self._raw.smt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, st: int):
        if st != 0:
            raise SmtError(st, self.lib.smt_last_error_string(self.h).decode() or self.lib.smt_status_string(st).decode())

    def sync(self):
        self._ck(self.lib.smt_sync(self.h))

    @property
    def torch_stream(self):
        """The context's stream as a torch stream (for wait_stream / events)."""
        if self._torch_stream is None:
            import torch

            self._torch_stream = torch.cuda.ExternalStream(int(self._raw.smt_get_stream(self.h)), device=f"cuda:{self.device}")
        return self._torch_stream

    def order_after_torch(self):
        """Queue the context's stream behind the work torch's current stream already holds."""
        import torch

        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.torch_stream.cuda_stream:
            self.torch_stream.wait_stream(cur)

    def order_torch_after(self):
        """Make torch's current stream wait for what the context's stream holds."""
        import torch

        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.torch_stream.cuda_stream:
            cur.wait_stream(self.torch_stream)

    def trim(self):
        self._ck(self.lib.smt_trim(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.smt_kernel_launches(self.h))

    def stage_timing(self, enable: bool):
        self._ck(self.lib.smt_stage_timing_enable(self.h, int(enable)))

    def read_stage_timing(self, reset: bool = True) -> dict:
        n = self.lib.smt_num_stages()
        buf = (C.c_double * n)()
        self._ck(self.lib.smt_stage_timing_read(self.h, buf, int(reset)))
        return {self.lib.smt_stage_name(k).decode(): float(buf[k]) for k in range(n)}

    # ------------------------------------------------------------------ AD-Census
    def adcensus_cost(self, left, right, D: int, sigmaC: float = 10.0, sigmaS: float = 30.0, view: int = 0,
                      want_ad: bool = False, want_census: bool = False):
        """AD_Census::ComputeADcensus (view 0) / ComputeADcensusRight (view 1).  Returns dict(cost[, ad, census])."""
        a = _Args(self, left, right)
        H, W = left.shape
        pl = a.inp(left, np.float32, (H, W))
        pr = a.inp(right, np.float32, (H, W))
        cost, pc = a.out((H, W, D), np.float32)
        ad, pa = a.out((H, W, D), np.float32, want_ad)
        cen, pn = a.out((H, W, D), np.float32, want_census)
        self._ck(self.lib.smt_adcensus_cost(self.h, pl, pr, H, W, D, sigmaC, sigmaS, view, pc, pa, pn, a.space))
        out = {"cost": cost}
        if want_ad:
            out["ad"] = ad
        if want_census:
            out["census"] = cen
        return out

    def census9x7(self, img):
        a = _Args(self, img)
        H, W = img.shape
        pi = a.inp(img, np.float32, (H, W))
        out, po = a.out((H, W), np.uint64)
        self._ck(self.lib.smt_census9x7_u64(self.h, pi, H, W, po, a.space))
        return out

    def wta(self, vol):
        a = _Args(self, vol)
        H, W, D = vol.shape
        pv = a.inp(vol, np.float32)
        disp, pd = a.out((H, W), np.float32)
        self._ck(self.lib.smt_wta_first_min(self.h, pv, H, W, D, pd, a.space))
        return disp

    # ------------------------------------------------------------------ CrossAggregator
    def cross_arms(self, bgr, L1=34, L2=17, t1=20, t2=6):
        a = _Args(self, bgr)
        H, W = bgr.shape[:2]
        pb = a.inp(bgr, np.uint8, (H, W, 3))
        arms, pa = a.out((H, W, 4), np.uint8)
        s0, p0 = a.out((H, W), np.uint16)
        s1, p1 = a.out((H, W), np.uint16)
        self._ck(self.lib.smt_cross_arms(self.h, pb, H, W, L1, L2, t1, t2, pa, p0, p1, a.space))
        return {"arms": arms, "sup0": s0, "sup1": s1}

    def aggregate_cross(self, bgr, cost, L1=34, L2=17, t1=20, t2=6, iters=4, mode=SUM_REFERENCE_ORDER,
                        want_arms: bool = False):
        a = _Args(self, bgr, cost)
        H, W, D = cost.shape
        pb = a.inp(bgr, np.uint8, (H, W, 3))
        pc = a.inp(cost, np.float32)
        out, po = a.out((H, W, D), np.float32)
        arms, pa = a.out((H, W, 4), np.uint8, want_arms)
        s0, p0 = a.out((H, W), np.uint16, want_arms)
        s1, p1 = a.out((H, W), np.uint16, want_arms)
        self._ck(self.lib.smt_aggregate_cross(self.h, pb, pc, H, W, D, L1, L2, t1, t2, iters, mode, po, pa, p0, p1, a.space))
        res = {"cost": out}
        if want_arms:
            res.update(arms=arms, sup0=s0, sup1=s1)
        return res

    def aggregate_fast_status(self) -> int:
        """Route of the last SUM_FAST aggregation on this engine: the number of (line, disparity chunk) tasks the float-pair
        passes redid behind the fixed-point sweeps (0 = none), or -1 when they took over whole sweeps."""
        v = C.c_int32(0)
        self._ck(self.lib.smt_aggregate_fast_status(self.h, C.byref(v)))
        return int(v.value)

    # ------------------------------------------------------------------ CrossArm / CBLSM
    def ad_cost(self, left, right, D, view=0):
        """CBLSM ComputeAD (view 0) / ComputeADRight (view 1) from uint8 grey images."""
        a = _Args(self, left, right)
        H, W = left.shape
        pl = a.inp(left, np.uint8, (H, W))
        pr = a.inp(right, np.uint8, (H, W))
        out, po = a.out((H, W, D), np.float32)
        self._ck(self.lib.smt_ad_cost(self.h, pl, pr, H, W, D, view, po, a.space))
        return out

    def arms_gray(self, img, tao, max_length=34, sec_length=17, variant=0):
        """variant 0: CrossArmAggregation's four passes (sticky member threshold, Right-pass quirk); 1: CBLSM ArmLength*."""
        a = _Args(self, img)
        ch = 3 if img.ndim == 3 else 1
        H, W = img.shape[:2]
        pi = a.inp(img, np.uint8)
        outs = [a.out((H, W), np.int32) for _ in range(4)]
        tf = C.c_int32(0)
        self._ck(self.lib.smt_arms_gray(self.h, pi, ch, H, W, int(tao), int(max_length), int(sec_length), int(variant),
                                        *[o[1] for o in outs], C.byref(tf), a.space))
        names = ("left", "right", "top", "bottom") if variant == 0 else ("left", "right", "up", "down")
        res = {n: o[0] for n, o in zip(names, outs)}
        res["tao_final"] = int(tf.value)
        return res

    def aggregate_rect(self, vol, arm_left, arm_right, arm_top, arm_bottom, variant=0, mode=SUM_REFERENCE_ORDER):
        """variant 0: AggregationVertical, 1: Aggregation (CrossArm.cpp:60,104), 2: costAggregationV5 (CBLSM.h:1179),
        3: costAggregationV4 (CBLSM.h:1128; the four arm arrays are then [H][W][D] volumes)."""
        a = _Args(self, vol, arm_left)
        H, W, D = vol.shape
        pv = a.inp(vol, np.float32)
        ashape = (H, W, D) if int(variant) == 3 else (H, W)
        ps = [a.inp(x, np.int32, ashape) for x in (arm_left, arm_right, arm_top, arm_bottom)]
        out, po = a.out((H, W, D), np.float32)
        self._ck(self.lib.smt_aggregate_rect(self.h, pv, *ps, H, W, D, int(variant), int(mode), po, a.space))
        return out

    # ------------------------------------------------------------------ ScanlineOptimizer
    def scanline(self, cost, gray, p1=10, p2=150, want: Sequence[str] = ("sum", "disp")):
        a = _Args(self, cost, gray)
        H, W, D = cost.shape
        pc = a.inp(cost, np.float32)
        pg = a.inp(gray, np.float32, (H, W))
        names = ["left", "right", "up", "down", "sum", "disp"]
        outs, ptrs = {}, []
        for n in names:
            arr, p = a.out((H, W) if n == "disp" else (H, W, D), np.float32, n in want)
            if n in want:
                outs[n] = arr
            ptrs.append(p)
        self._ck(self.lib.smt_scanline(self.h, pc, pg, H, W, D, int(p1), int(p2), *ptrs, a.space))
        return outs

    # ------------------------------------------------------------------ post-processing
    def lrcheck(self, dispL, dispR, gate=2, want_lists: bool = True):
        """LeftRightConsistency.  Returns (dispL_out, occlusions[n,2], mismatches[m,2])."""
        a = _Args(self, dispL, dispR)
        H, W = dispL.shape
        d, pd = a.out((H, W), np.float32, init=dispL)
        pr = a.inp(dispR, np.float32, (H, W))
        occ, po = a.out((H * W, 2), np.int32, want_lists)
        mis, pm = a.out((H * W, 2), np.int32, want_lists)
        no, nm = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.smt_lrcheck(self.h, pd, pr, W, H, int(gate), po, C.byref(no), pm, C.byref(nm), a.space))
        if not want_lists:
            return d, no.value, nm.value
        return d, occ[: no.value], mis[: nm.value]

    def lrcheck_and(self, dispL, dispR, gate=2.0, want_lists: bool = True):
        """LeftAndRightConsistency.  Returns (lastDisp, occlusions, mismatches)."""
        a = _Args(self, dispL, dispR)
        H, W = dispL.shape
        pl = a.inp(dispL, np.float32, (H, W))
        pr = a.inp(dispR, np.float32, (H, W))
        last, plast = a.out((H, W), np.float32)
        occ, po = a.out((H * W, 2), np.int32, want_lists)
        mis, pm = a.out((H * W, 2), np.int32, want_lists)
        no, nm = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.smt_lrcheck_and(self.h, pl, pr, plast, W, H, float(gate), po, C.byref(no), pm, C.byref(nm), a.space))
        if not want_lists:
            return last, no.value, nm.value
        return last, occ[: no.value], mis[: nm.value]

    def remove_speckles(self, disp, diff=1, min_area=30, invalid=INVALID_INT_FROM_INF):
        a = _Args(self, disp)
        H, W = disp.shape
        d, pd = a.out((H, W), np.float32, init=disp)
        self._ck(self.lib.smt_remove_speckles(self.h, pd, W, H, int(diff), int(min_area), int(invalid), a.space))
        return d

    def fill_the_hole(self, disp, disp_range, occ, mis, fixed: bool = False):
        """FillTheHole (PostProcessing.h:156).  As written (default): list entries (y, x) address disp.flat[y*rows + x]
        (the function swaps the axes), invalid = 65535.  fixed=True: entries are (row, col), invalid = +inf.
        Returns (disp_out, mismatch_list_after)."""
        a = _Args(self, disp)
        H, W = disp.shape
        d, pd = a.out((H, W), np.float32, init=disp)
        occ = np.ascontiguousarray(np.asarray(occ.cpu() if hasattr(occ, "cpu") else occ, np.int32).reshape(-1, 2))
        mis = np.asarray(mis.cpu() if hasattr(mis, "cpu") else mis, np.int32).reshape(-1, 2)
        if len(mis) > H * W or len(occ) > H * W:
            raise ValueError("list longer than the image")
        m0 = np.zeros((H * W, 2), np.int32)
        m0[: len(mis)] = mis
        po = a.inp(occ, np.int32) if len(occ) else None
        m, pm = a.out((H * W, 2), np.int32, init=m0)
        nm = C.c_int32(len(mis))
        self._ck(self.lib.smt_fill_the_hole(self.h, pd, H, W, int(disp_range), po, len(occ), pm, C.byref(nm), 1 if fixed else 0, a.space))
        return d, m[: nm.value]

    # ------------------------------------------------------------------ multi-step refinement (fixed specification)
    def invalid_class_map(self, occ, mis, shape):
        """uint8 [H][W]: 0 valid, 1 occlusion, 2 mismatch, from LeftRightConsistency's two (row, col) lists."""
        H, W = shape
        occ = np.ascontiguousarray(np.asarray(occ.cpu() if hasattr(occ, "cpu") else occ, np.int32).reshape(-1, 2))
        mis = np.ascontiguousarray(np.asarray(mis.cpu() if hasattr(mis, "cpu") else mis, np.int32).reshape(-1, 2))
        a = _Args(self, occ)
        po = a.inp(occ, np.int32) if len(occ) else None
        pm = a.inp(mis, np.int32) if len(mis) else None
        cls, pc = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_invalid_class_map(self.h, po, len(occ), pm, len(mis), H, W, pc, a.space))
        return cls

    def refine_region_voting(self, disp, cls, arms, D, ts=20, th=0.4, iters=5):
        a = _Args(self, disp, cls, arms)
        H, W = disp.shape
        d, pd = a.out((H, W), np.float32, init=disp)
        c, pc = a.out((H, W), np.uint8, init=cls)
        pa = a.inp(arms, np.uint8, (H, W, 4))
        self._ck(self.lib.smt_refine_region_voting(self.h, pd, pc, pa, H, W, int(D), int(ts), float(th), int(iters), a.space))
        return d, c

    def refine_interpolate(self, disp, cls, bgr, max_len):
        a = _Args(self, disp, cls, bgr)
        H, W = disp.shape
        d, pd = a.out((H, W), np.float32, init=disp)
        c, pc = a.out((H, W), np.uint8, init=cls)
        pb = a.inp(bgr, np.uint8, (H, W, 3))
        self._ck(self.lib.smt_refine_interpolate(self.h, pd, pc, pb, H, W, int(max_len), a.space))
        return d, c

    def refine_discontinuity(self, disp, cost, thres=5.0):
        a = _Args(self, disp, cost)
        H, W, D = cost.shape
        d, pd = a.out((H, W), np.float32, init=disp)
        pc = a.inp(cost, np.float32, (H, W, D))
        self._ck(self.lib.smt_refine_discontinuity(self.h, pd, pc, H, W, D, float(thres), a.space))
        return d

    def median(self, disp, k=3):
        a = _Args(self, disp)
        H, W = disp.shape
        pi = a.inp(disp, np.float32, (H, W))
        out, po = a.out((H, W), np.float32)
        self._ck(self.lib.smt_median(self.h, pi, po, W, H, int(k), a.space))
        return out

    # ------------------------------------------------------------------ window matchers
    def sad(self, left_padded, right_padded, D, winsize=3, right_view=False, want_costs=False, want_subpixel=False):
        """GetPointDepthLeft / GetPointDepthRight (SAD/Sad.h:96,141) on replicate-padded grey images (pad = winsize+1).
        want_subpixel (left view): also the parabola-refined value OptimalDisparity computes and discards (Sad.h:76-84);
        the result is then a dict."""
        a = _Args(self, left_padded, right_padded)
        Hp, Wp = left_padded.shape
        pad = winsize + 1
        pl = a.inp(left_padded, np.uint8, (Hp, Wp))
        pr = a.inp(right_padded, np.uint8, (Hp, Wp))
        disp, pd = a.out((Hp - 2 * pad, Wp - 2 * pad), np.int32)
        costs, pc = a.out((Hp - 2 * pad, Wp - 2 * pad, D), np.float32, want_costs)
        sub, psub = a.out((Hp - 2 * pad, Wp - 2 * pad), np.float32, want_subpixel)
        self._ck(self.lib.smt_sad(self.h, pl, pr, Hp, Wp, D, winsize, int(bool(right_view)), pd, pc, psub, a.space))
        if want_subpixel:
            return {"disp": disp, "costs": costs, "subpixel": sub}
        return (disp, costs) if want_costs else disp

    def sad_crosscheck(self, disp_left, disp_right):
        """CrossCheckDiaparity on the CV_32S maps (SAD/Sad.h:184).  Returns (out, occlusions[n,2], mismatches[m,2])."""
        a = _Args(self, disp_left, disp_right)
        H, W = disp_left.shape
        pl = a.inp(disp_left, np.int32, (H, W))
        pr = a.inp(disp_right, np.int32, (H, W))
        out, po = a.out((H, W), np.int32)
        occ, pocc = a.out((H * W, 2), np.int32)
        mis, pmis = a.out((H * W, 2), np.int32)
        no, nm = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.smt_sad_crosscheck(self.h, pl, pr, H, W, po, pocc, C.byref(no), pmis, C.byref(nm), a.space))
        return out, occ[: no.value], mis[: nm.value]

    def ncc_shift(self, left, right, right_view=False, add_constant=False):
        """ncc() (NCC/NCC.h:117): shifted-image NCC, clipped 11x11 window, offsets 1..79; uint8 depth = 3 * best offset."""
        a = _Args(self, left, right)
        H, W = left.shape
        pl = a.inp(left, np.uint8, (H, W))
        pr = a.inp(right, np.uint8, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_ncc_shift(self.h, pl, pr, H, W, int(bool(right_view)), int(bool(add_constant)), po, a.space))
        return out

    def pipeline_cblsm(self, gray_left, gray_right, D, tao=25, max_length=34, sec_length=17, mode=SUM_REFERENCE_ORDER):
        """CBLSM's main() (CBLSM.cpp:64-153) as one call.  Returns (disp_left, disp_right)."""
        a = _Args(self, gray_left, gray_right)
        H, W = gray_left.shape
        pl = a.inp(gray_left, np.uint8, (H, W))
        pr = a.inp(gray_right, np.uint8, (H, W))
        dl, pdl = a.out((H, W), np.float32)
        dr, pdr = a.out((H, W), np.float32)
        self._ck(self.lib.smt_pipeline_cblsm(self.h, pl, pr, H, W, int(D), int(tao), int(max_length), int(sec_length), int(mode), pdl, pdr, a.space))
        return dl, dr

    def ncc(self, left, right, D, win_size=10, want_costs: bool = False):
        """NCC_algorithem (NCC/NCC.h:69); want_costs adds ComputeCost's value per (pixel, d) as float64 [H][W][D]."""
        a = _Args(self, left, right)
        H, W = left.shape
        pl = a.inp(left, np.uint8, (H, W))
        pr = a.inp(right, np.uint8, (H, W))
        disp, pd = a.out((H, W), np.int32)
        costs, pc = a.out((H, W, D), np.float64, want_costs)
        self._ck(self.lib.smt_ncc(self.h, pl, pr, H, W, win_size, D, pd, pc, a.space))
        return (disp, costs) if want_costs else disp

    def asw(self, left_padded, right_padded, D, win_size=11, space_sigma=50.0, color_sigma=30.0, T=40, right_view=False,
            want_costs: bool = False):
        """AdaptiveSupportWeight / AdaptiveSupportWeightRight (ASW/ASW.h:329,382) on replicate-padded grey images;
        want_costs adds the per-pixel costVolume rows as float32 [H][W][D]."""
        a = _Args(self, left_padded, right_padded)
        Hp, Wp = left_padded.shape
        pad = win_size + 1
        pl = a.inp(left_padded, np.uint8, (Hp, Wp))
        pr = a.inp(right_padded, np.uint8, (Hp, Wp))
        disp, pd = a.out((Hp - 2 * pad, Wp - 2 * pad), np.float32)
        costs, pc = a.out((Hp - 2 * pad, Wp - 2 * pad, D), np.float32, want_costs)
        self._ck(self.lib.smt_asw(self.h, pl, pr, Hp, Wp, win_size, D, float(space_sigma), float(color_sigma), int(T),
                                  int(bool(right_view)), pd, pc, a.space))
        return (disp, costs) if want_costs else disp

    def asw_crosscheck(self, disp_left, disp_right):
        a = _Args(self, disp_left, disp_right)
        H, W = disp_left.shape
        pl = a.inp(disp_left, np.float32, (H, W))
        pr = a.inp(disp_right, np.float32, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_asw_crosscheck(self.h, pl, pr, H, W, po, a.space))
        return out

    # ------------------------------------------------------------------ pre / post (what the drivers do with OpenCV)
    def bgr_to_gray(self, bgr):
        a = _Args(self, bgr)
        H, W = bgr.shape[:2]
        pb = a.inp(bgr, np.uint8, (H, W, 3))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_bgr_to_gray(self.h, pb, H, W, po, a.space))
        return out

    def replicate_pad(self, img, pad):
        a = _Args(self, img)
        H, W = img.shape
        pi = a.inp(img, np.uint8, (H, W))
        out, po = a.out((H + 2 * pad, W + 2 * pad), np.uint8)
        self._ck(self.lib.smt_replicate_pad(self.h, pi, H, W, int(pad), po, a.space))
        return out

    def u8_to_f32(self, img):
        a = _Args(self, img)
        pi = a.inp(img, np.uint8)
        out, po = a.out(tuple(img.shape), np.float32)
        self._ck(self.lib.smt_u8_to_f32(self.h, pi, int(np.prod(img.shape)), po, a.space))
        return out

    def transform_to_show(self, disp):
        a = _Args(self, disp)
        H, W = disp.shape
        pi = a.inp(disp, np.float32, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_transform_to_show(self.h, pi, H, W, po, a.space))
        return out

    def normalize_minmax_u8(self, src):
        a = _Args(self, src)
        H, W = src.shape
        pi = a.inp(src, np.float32, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_normalize_minmax_u8(self.h, pi, H, W, po, a.space))
        return out

    # ------------------------------------------------------------------ the ASW driver's display chain (ASWeight.cpp:66-78)
    def normalize_minmax_u8_from_u8(self, src):
        a = _Args(self, src)
        H, W = src.shape
        pi = a.inp(src, np.uint8, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_normalize_minmax_u8_from_u8(self.h, pi, H, W, po, a.space))
        return out

    def filter_speckles_u8(self, img, new_val=0, max_speckle_size=40, max_diff=2):
        """cv::filterSpeckles on an 8-bit map; returns the filtered copy."""
        a = _Args(self, img)
        H, W = img.shape
        out, po = a.out((H, W), np.uint8, init=img)
        self._ck(self.lib.smt_filter_speckles_u8(self.h, po, H, W, int(new_val), int(max_speckle_size), int(max_diff), a.space))
        return out

    def median_blur_u8(self, img, ksize):
        a = _Args(self, img)
        H, W = img.shape
        pi = a.inp(img, np.uint8, (H, W))
        out, po = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_median_blur_u8(self.h, pi, po, H, W, int(ksize), a.space))
        return out

    def fill_image_new(self, img, past_end=0):
        """FillImageNew (ASW.h:434) as written.  Returns (image, n_unassigned)."""
        a = _Args(self, img)
        H, W = img.shape
        out, po = a.out((H, W), np.uint8, init=img)
        n = C.c_int32(0)
        self._ck(self.lib.smt_fill_image_new(self.h, po, H, W, int(past_end), C.byref(n), a.space))
        return out, int(n.value)

    def asw_display_chain(self, disp_left, disp_right):
        """Cross-check, normalize, filterSpeckles(0, 40, 2), medianBlur 5, FillImageNew, medianBlur 3.
        Returns the three images the driver writes: (speckle, filled, median)."""
        a = _Args(self, disp_left, disp_right)
        H, W = disp_left.shape
        pl = a.inp(disp_left, np.float32, (H, W))
        pr = a.inp(disp_right, np.float32, (H, W))
        o1, p1 = a.out((H, W), np.uint8)
        o2, p2 = a.out((H, W), np.uint8)
        o3, p3 = a.out((H, W), np.uint8)
        self._ck(self.lib.smt_asw_display_chain(self.h, pl, pr, H, W, p1, p2, p3, a.space))
        return o1, o2, o3

    def write_png(self, path, img):
        """imwrite(path, 8-bit image): grey [H][W] or BGR [H][W][3]."""
        a = _Args(self, img)
        H, W = img.shape[:2]
        ch = 1 if img.ndim == 2 else int(img.shape[2])
        pi = a.inp(img, np.uint8, (H, W) if ch == 1 else (H, W, ch))
        self._ck(self.lib.smt_write_png_u8(self.h, str(path).encode(), pi, H, W, ch, a.space))

    # ------------------------------------------------------------------ full pipeline
    def default_params(self, H, W, D) -> AdCensusParams:
        p = AdCensusParams()
        self.lib.smt_adcensus_default_params(C.byref(p), H, W, D)
        return p

    def pipeline(self, bgr_left, bgr_right, D, gray_left=None, gray_right=None, params: Optional[AdCensusParams] = None,
                 want: Sequence[str] = ("disp",), out=None, **overrides):
        """P_full.  ``want`` from {"disp", "left_raw", "right_raw", "lrc"}; keyword overrides set params fields.
        ``out``: optional caller-owned float32 [H, W] array for the final disparity (host memspace: e.g. a pinned buffer,
        like the reference's caller-owned output arrays)."""
        a = _Args(self, bgr_left, bgr_right)
        H, W = bgr_left.shape[:2]
        p = params or self.default_params(H, W, D)
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown pipeline parameter {k!r}")
            setattr(p, k, v)
        pbl = a.inp(bgr_left, np.uint8, (H, W, 3))
        pbr = a.inp(bgr_right, np.uint8, (H, W, 3))
        pgl = a.inp(gray_left, np.uint8, (H, W))
        pgr = a.inp(gray_right, np.uint8, (H, W))
        if out is not None and not a.device:
            if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (H, W) and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous float32 [H, W] numpy array")
            disp, pd = out, C.c_void_p(out.ctypes.data)
        else:
            disp, pd = a.out((H, W), np.float32)
        lraw, p1 = a.out((H, W), np.float32, "left_raw" in want)
        rraw, p2 = a.out((H, W), np.float32, "right_raw" in want)
        lrc, p3 = a.out((H, W), np.float32, "lrc" in want)
        self._ck(self.lib.smt_pipeline_adcensus_full(self.h, C.byref(p), pbl, pbr, pgl, pgr, pd, p1, p2, p3, a.space))
        out = {"disp": disp}
        for n, v in (("left_raw", lraw), ("right_raw", rraw), ("lrc", lrc)):
            if n in want:
                out[n] = v
        return out


_default: dict[int, Engine] = {}


def default_engine(device: int = 0) -> Engine:
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]


# ====================================================================================== reference-named facade
class AD_Census:
    """AD-CensusV1/AD-Census.h:9-43 -- same method names and argument order (cv::Mat arguments are accepted and
    ignored, exactly as the reference only stores them, :329-330)."""

    def __init__(self, engine: Optional[Engine] = None):
        self._e = engine
        self._costL = self._costR = None

    def Initialize(self, leftImage, rightImage, dispRange, row, col, LImage=None, RImage=None, sigmaC=10.0, sigmaS=30.0):
        self._e = self._e or default_engine()
        self._l = np.ascontiguousarray(leftImage, np.float32).reshape(row, col)
        self._r = np.ascontiguousarray(rightImage, np.float32).reshape(row, col)
        self._D, self._sc, self._ss = int(dispRange), float(sigmaC), float(sigmaS)
        self._costL = self._costR = None

    def ComputeADcensus(self):
        self._costL = self._e.adcensus_cost(self._l, self._r, self._D, self._sc, self._ss, 0)["cost"]

    def ComputeADcensusRight(self):
        self._costR = self._e.adcensus_cost(self._l, self._r, self._D, self._sc, self._ss, 1)["cost"]

    def GetPtrLeft(self):
        return self._costL  # None until computed (the reference returns nullptr, :50-60)

    def GetPtrRight(self):
        return self._costR

    def WTA(self, leftdisp, rightDisp):
        leftdisp[...] = self._e.wta(self._costL).reshape(leftdisp.shape)
        rightDisp[...] = self._e.wta(self._costR).reshape(rightDisp.shape)


class CrossAggregator:
    """CBLSM/cross_aggregator.h:27-113."""

    def __init__(self, engine: Optional[Engine] = None, mode: int = SUM_REFERENCE_ORDER):
        self._e = engine
        self._mode = mode
        self._init = False
        self._arms = self._cost = None
        self._params = (34, 17, 20, 6)

    def Initialize(self, width, height, min_disparity, max_disparity) -> bool:
        self._e = self._e or default_engine()
        self._w, self._h = int(width), int(height)
        self._D = int(max_disparity) - int(min_disparity)
        self._init = self._w * self._h > 0 and self._D > 0  # cross_aggregator.cpp:26-31
        return self._init

    def SetData(self, img_left, img_right, cost_init):
        self._img = np.ascontiguousarray(img_left, np.uint8).reshape(self._h, self._w, 3)
        self._cost_init = np.ascontiguousarray(cost_init, np.float32).reshape(self._h, self._w, self._D)

    def SetParams(self, cross_L1, cross_L2, cross_t1, cross_t2):
        self._params = (int(cross_L1), int(cross_L2), int(cross_t1), int(cross_t2))

    def Aggregate(self, num_iters):
        if not self._init:  # silently no-op (:91-93)
            return
        r = self._e.aggregate_cross(self._img, self._cost_init, *self._params, iters=int(num_iters), mode=self._mode,
                                    want_arms=True)
        self._cost, self._arms = r["cost"], r["arms"]
        self._sup = (r["sup0"], r["sup1"])

    def get_arms_ptr(self):
        return self._arms

    def get_cost_ptr(self):
        return self._cost


class ScanlineOptimizer:
    """AD-CensusV1/ScanlineOptimizer.h:8-34."""

    def __init__(self, engine: Optional[Engine] = None):
        self._e = engine
        self._out = {}

    def Initialize(self, row, col, dispRange, costVolume, p1, p2):
        self._e = self._e or default_engine()
        self._shape = (int(row), int(col), int(dispRange))
        self._p1, self._p2 = int(p1), int(p2)  # int parameters (:66)

    def ScanLine(self, costVolume, Image, want=("left", "right", "up", "down", "sum", "disp")):
        H, W, D = self._shape
        cost = np.ascontiguousarray(costVolume, np.float32).reshape(H, W, D)
        gray = np.ascontiguousarray(Image, np.float32).reshape(H, W)
        self._out = self._e.scanline(cost, gray, self._p1, self._p2, want=want)

    def WTA(self, disp):
        disp[...] = self._out["disp"].reshape(disp.shape)

    @property
    def volumes(self):
        return self._out


def LeftRightConsistency(col, row, gate, leftDisp, rightDisp, occlusion: list, mismatch: list, engine=None):
    """PostProcessing.h:72 -- modifies leftDisp in place, fills the two (row, col) lists."""
    e = engine or default_engine()
    d, occ, mis = e.lrcheck(np.asarray(leftDisp, np.float32).reshape(row, col),
                            np.asarray(rightDisp, np.float32).reshape(row, col), gate)
    leftDisp[...] = d.reshape(leftDisp.shape)
    occlusion[:] = [tuple(x) for x in occ.tolist()]
    mismatch[:] = [tuple(x) for x in mis.tolist()]


def LeftAndRightConsistency(leftDisp, rightDisp, lastDisp, col, row, gate, occlusion: list, mismatch: list, engine=None):
    """PostProcessing.h:10."""
    e = engine or default_engine()
    last, occ, mis = e.lrcheck_and(np.asarray(leftDisp, np.float32).reshape(row, col),
                                   np.asarray(rightDisp, np.float32).reshape(row, col), gate)
    lastDisp[...] = last.reshape(lastDisp.shape)
    occlusion.extend(tuple(x) for x in occ.tolist())  # the reference appends without clearing (:44-66)
    mismatch.extend(tuple(x) for x in mis.tolist())


def RemoveSpeckles(disparity_map, width, height, diff_insame, min_speckle_aera, invalid_val, engine=None):
    """PostProcessing.h:250 -- invalid_val is an int; +inf becomes 2147483647 as under g++."""
    e = engine or default_engine()
    if isinstance(invalid_val, float) and np.isinf(invalid_val):
        invalid_val = INVALID_INT_FROM_INF
    d = e.remove_speckles(np.asarray(disparity_map, np.float32).reshape(height, width), diff_insame, min_speckle_aera,
                          int(invalid_val))
    disparity_map[...] = d.reshape(disparity_map.shape)


def MedianFilter(in_, out, width, height, wnd_size, engine=None):
    """PostProcessing.h:314."""
    e = engine or default_engine()
    out[...] = e.median(np.asarray(in_, np.float32).reshape(height, width), wnd_size).reshape(out.shape)
