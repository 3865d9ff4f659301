"""ctypes binding of libsmt_b200.so (the C ABI declared in include/smt_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and if no CUDA device is
present ``smt_create`` fails -- the product path never routes through ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
# SMT_B200_LIB: developer override used by scripts/ab_agg.sh to time alternative builds of the same library (A/B runs)
LIB_PATH = Path(os.environ.get("SMT_B200_LIB") or HERE / "libsmt_b200.so")
CSRC = HERE / "csrc"
HEADER = HERE.parent / "include" / "smt_b200.h"

SMT_HOST, SMT_DEVICE = 0, 1
SUM_REFERENCE_ORDER, SUM_FAST = 0, 1

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p
_i = C.c_int32
_f = C.c_float


class AdCensusParams(C.Structure):
    """Mirror of ``smt_adcensus_params``."""

    _fields_ = [
        ("rows", _i), ("cols", _i), ("D", _i),
        ("sigmaC", _f), ("sigmaS", _f),
        ("L1", _i), ("L2", _i), ("t1", _i), ("t2", _i),
        ("agg_iters", _i),
        ("p1", _i), ("p2", _i),
        ("gate", _i),
        ("speckle_diff", _i), ("speckle_min_area", C.c_uint32), ("speckle_invalid", _i),
        ("median_size", _i),
        ("sum_mode", _i),
        ("irv_iters", _i), ("irv_ts", _i), ("irv_th", _f),
        ("do_interpolation", _i), ("do_discontinuity_adjustment", _i), ("discontinuity_thres", _f),
    ]


# name -> (restype, argtypes); every symbol include/smt_b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "smt_abi_version": (_i, []),
    "smt_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "smt_destroy": (None, [_vp]),
    "smt_sync": (_i, [_vp]),
    "smt_get_stream": (_vp, [_vp]),
    "smt_last_error_string": (C.c_char_p, [_vp]),
    "smt_status_string": (C.c_char_p, [_i]),
    "smt_kernel_launches": (C.c_uint64, [_vp]),
    "smt_num_stages": (_i, []),
    "smt_stage_name": (C.c_char_p, [_i]),
    "smt_stage_timing_enable": (_i, [_vp, _i]),
    "smt_stage_timing_read": (_i, [_vp, _f64p, _i]),
    "smt_trim": (_i, [_vp]),
    "smt_device_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "smt_device_free": (_i, [_vp, _vp]),
    "smt_memcpy_h2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "smt_memcpy_d2h": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "smt_adcensus_cost": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _i]),
    "smt_census9x7_u64": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "smt_wta_first_min": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "smt_cross_arms": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    "smt_aggregate_cross": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "smt_cross_sweep_count": (_i, [_i]),
    "smt_cross_sweep_info": (_i, [_i, _i, _i32p, _i32p]),
    "smt_aggregate_cross_sweep": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _i]),
    "smt_aggregate_fast_status": (_i, [_vp, _i32p]),
    "smt_invalid_class_map": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _i]),
    "smt_refine_region_voting": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i]),
    "smt_refine_interpolate": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "smt_refine_discontinuity": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _i]),
    "smt_band_create": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "smt_band_destroy": (None, [_vp]),
    "smt_band_geometry": (_i, [_vp, _i32p, _i32p, _i32p, _i32p]),
    "smt_band_volume": (_i, [_vp, C.POINTER(_vp)]),
    "smt_band_export": (_i, [_vp, _vp]),
    "smt_band_connect": (_i, [_vp, _vp, _vp]),
    "smt_band_connect_local": (_i, [_vp, _vp, _vp]),
    "smt_halo_exchange": (_i, [_vp]),
    "smt_band_aggregate_cross": (_i, [_vp, _vp, _vp, _vp, _i, _f, _i]),
    "smt_band_sync": (_i, [_vp, _f64p]),
    "smt_ad_cost": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i]),
    "smt_arms_gray": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i32p, _i]),
    "smt_aggregate_rect": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i]),
    "smt_scanline": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "smt_lrcheck": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i32p, _vp, _i32p, _i]),
    "smt_lrcheck_and": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp, _i32p, _vp, _i32p, _i]),
    "smt_remove_speckles": (_i, [_vp, _vp, _i, _i, _i, C.c_uint32, _i, _i]),
    "smt_median": (_i, [_vp, _vp, _vp, _i, _i, _i, _i]),
    "smt_fill_the_hole": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i]),
    "smt_sad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    "smt_sad_crosscheck": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i32p, _vp, _i32p, _i]),
    "smt_ncc_shift": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i]),
    "smt_pipeline_cblsm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i]),
    "smt_ncc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i]),
    "smt_asw": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_double, C.c_double, _i, _i, _vp, _vp, _i]),
    "smt_asw_crosscheck": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i]),
    "smt_bgr_to_gray": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "smt_replicate_pad": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "smt_u8_to_f32": (_i, [_vp, _vp, C.c_size_t, _vp, _i]),
    "smt_transform_to_show": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "smt_normalize_minmax_u8": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "smt_normalize_minmax_u8_from_u8": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "smt_filter_speckles_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i]),
    "smt_median_blur_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _i]),
    "smt_fill_image_new": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int32), _i]),
    "smt_asw_display_chain": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "smt_write_png_u8": (_i, [_vp, C.c_char_p, _vp, _i, _i, _i, _i]),
    "smt_adcensus_default_params": (None, [C.POINTER(AdCensusParams), _i, _i, _i]),
    "smt_pipeline_adcensus_full": (_i, [_vp, C.POINTER(AdCensusParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
}


def header_symbols() -> list[str]:
    """Every SMT_API function name declared in include/smt_b200.h."""
    text = HEADER.read_text()
    return sorted(set(re.findall(r"SMT_API\s+[\w\s\*]+?\b(smt_\w+)\s*\(", text)))


def build(verbose: bool = False) -> Path:
    """Compile libsmt_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    jobs = str(max(1, min(8, os.cpu_count() or 1)))
    proc = subprocess.run(["make", "-C", str(CSRC), "-j", jobs], capture_output=not verbose, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libsmt_b200.so failed:\n" + (proc.stdout or "") + (proc.stderr or ""))
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the library and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C stereo_match_traditional_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            if os.environ.get("SMT_B200_LIB"):   # an A/B build of an older revision may lack newer entry points
                continue
            raise ImportError(f"{LIB_PATH} does not export {name}: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
