"""ctypes binding of libglasdi_b200.so (the C-ABI declared in include/glasdi_b200.h).

There is NO CPU fallback: if the shared library is missing it is built with nvcc; if that is not
possible, or no sm_100 GPU is present when a handle is requested, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOT_READY, ERR_NUMERIC, ERR_OOM = range(7)

ACT_IDS = {"sigmoid": 0, "tanh": 1, "softplus": 2, "ReLU": 3, "ELU": 4, "SiLU": 5, "GELU": 6, "identity": 7,
           "hardshrink": 8, "hardsigmoid": 9, "hardtanh": 10, "hardswish": 11, "leakyReLU": 12, "logsigmoid": 13,
           "ReLU6": 14, "SELU": 15, "CELU": 16, "mish": 17, "softshrink": 18, "tanhshrink": 19}
FD_IDS = {"sbp12": 0, "sbp24": 1, "sbp36": 2, "sbp48": 3}
OPT_SPLIT_PASSES, OPT_INTEGRATOR, OPT_RK4_SUBSTEPS, OPT_MAX_CTAS, OPT_STAGE_TIMING, OPT_DECODE_KERNEL = 0, 1, 2, 3, 4, 5
OPT_MODE, OPT_SCREEN_MARGIN, OPT_WIDE_TENSOR, OPT_GRAM_KERNEL, OPT_TAYLOR_RADIUS, OPT_SINDY_LIBRARY, OPT_SCREEN_BOUND, OPT_MT_SEGMENT_LOG2 = 6, 7, 8, 9, 10, 11, 12, 13
LIB_POLY1, LIB_POLY1_SINE, LIB_POLY2, LIB_POLY2_SINE = 0, 1, 2, 3
MODE_DIRECT, MODE_SCREENED, MODE_SCREENED_GRAM = 0, 1, 2
INT_EXPM, INT_RK4 = 0, 1

# name -> (restype, argtypes); must list EVERY symbol of include/glasdi_b200.h (tests check this)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SIGNATURES = {
    "glasdi_abi_version": (_i, []),
    "glasdi_source_hash": (C.c_char_p, []),
    "glasdi_library_terms": (_i, [_i, _i]),
    "glasdi_create": (_i, [C.POINTER(_vp), _i]),
    "glasdi_destroy": (_i, [_vp]),
    "glasdi_last_error": (C.c_char_p, [_vp]),
    "glasdi_set_option": (_i, [_vp, _i, _i]),
    "glasdi_get_option": (_i, [_vp, _i, C.POINTER(_i)]),
    "glasdi_launch_count": (_i64, [_vp]),
    "glasdi_stage_times": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i)]),
    "glasdi_screen_stats": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i64), C.POINTER(C.c_float), _vp]),
    "glasdi_screen_bound_count": (_i, [_vp, C.POINTER(_i64), _vp]),
    "glasdi_set_decoder": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp), _i]),
    "glasdi_rollout": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp]),
    "glasdi_rollout_grid": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_d), _vp, _vp]),
    "glasdi_greedy_step": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp]),
    "glasdi_decode_max_std": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "glasdi_decode_stat_fields": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "glasdi_decode": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "glasdi_gp_predict": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "glasdi_normal_mt19937": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "glasdi_set_encoder": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp), _i]),
    "glasdi_encode": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "glasdi_check_numeric": (_i, [_vp, _vp]),
    "glasdi_argmax_first": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "glasdi_pack_maxloc": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "glasdi_fd_dzdt": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp]),
    "glasdi_sindy_fit": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "glasdi_mlp_forward_train": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp), _i, _vp, _i64, _vp, _vp, _vp, _vp]),
    "glasdi_mlp_backward": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), _i, _vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(_vp),
                                 C.POINTER(_vp), _vp]),
    "glasdi_mse_loss_grad": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "glasdi_sindy_fit_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _i, _vp, _vp]),
}


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Loads (building first if needed) the shared library and declares every signature."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: run `python -m gplasdi_b200.build` (no CPU fallback exists)")
        _build.build()
    # a library older than the sources beside it is rebuilt when a compiler is there and refused otherwise -- decided from
    # the hash file the build writes next to the library, BEFORE the library is opened (dlopen would keep a stale mapping)
    if os.path.isdir(_build.CSRC) and os.environ.get("GLASDI_LIBDIR") is None:
        built_from = _build.built_hash()
        if built_from != _build.source_hash():
            import shutil
            if not build_if_missing or shutil.which(_build._nvcc()) is None:
                raise RuntimeError(f"{path} was built from other sources (hash {built_from[:12]}, sources "
                                   f"{_build.source_hash()[:12]}): run `python -m gplasdi_b200.build`")
            _build.build(force=True)       # file times cannot be trusted once the hashes disagree
    lib = _open(path)
    _LIB = lib
    return lib


def _open(path):
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


class Mt19937State(C.Structure):
    """glasdi_mt19937_state of include/glasdi_b200.h = np.random.get_state() of the legacy global generator."""
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("cached_gaussian", C.c_double)]


class GlasdiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"glasdi_b200 status {code}: {msg}")
        self.code = code
