"""ctypes binding of lib/libaarmvs_sm100.so (the C ABI declared in include/aarmvs.h).

There is no fallback: if the library is missing or a call fails, RuntimeError is raised.
"""
import ctypes
import os

from . import build as _build

ERR_ARG = 1000001
ABI_VERSION = 5
SRC_PLAIN, SRC_GN_RELU, SRC_MAXPOOL2, SRC_NCHW = 0, 1, 2, 3

c_float_p = ctypes.POINTER(ctypes.c_float)


class ConvSrc(ctypes.Structure):
    """struct aarmvs_conv_src (include/aarmvs.h)."""
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("channels", ctypes.c_int),
        ("mode", ctypes.c_int),
        ("planes", ctypes.c_int),
        ("gn_groups", ctypes.c_int),
        ("gn_sums", ctypes.c_void_p),
        ("gn_weight", ctypes.c_void_p),
        ("gn_bias", ctypes.c_void_p),
        ("gn_count", ctypes.c_double),
        ("pp_stride", ctypes.c_size_t),
        ("pp_phase", ctypes.c_int),
    ]


class PlaneSrc(ctypes.Structure):
    """struct aarmvs_plane_src (include/aarmvs.h): fp16 (hi, lo) operand planes [pp][2][channels/8][H][W][8]."""
    _fields_ = [
        ("planes", ctypes.c_void_p),
        ("channels", ctypes.c_int),
        ("pp", ctypes.c_int),
        ("pp_phase", ctypes.c_int),
    ]


class Seg(ctypes.Structure):
    """struct aarmvs_seg (include/aarmvs.h): one channel segment of a concatenated / split NHWC operand."""
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("channels", ctypes.c_int),
        ("accumulate", ctypes.c_int),
    ]


_P, _I, _SZ, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
_SRCP = ctypes.POINTER(ConvSrc)
_PLSP = ctypes.POINTER(PlaneSrc)
_SEGP = ctypes.POINTER(Seg)

# name -> (restype, argtypes); must list every symbol include/aarmvs.h declares (tests/test_capi_symbols.py)
PROTOTYPES = {
    "aarmvs_version": (_I, []),
    "aarmvs_sm_count": (_I, []),
    "aarmvs_conv2d": (_I, [_SRCP, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "aarmvs_deform_conv": (_I, [_SRCP, _I, _I, _P, _P, _P, _P, _P, _P]),
    "aarmvs_deform_conv_tc": (_I, [_SRCP, _I, _I, _P, _P, _P, _P, _P, _P]),
    "aarmvs_feat_assemble": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "aarmvs_warp": (_I, [_P, _P, _I, _I, c_float_p, _F, _I, _P]),
    "aarmvs_omega_stats_bytes": (_SZ, [_I]),
    "aarmvs_omega_tail_floats": (_I, []),
    "aarmvs_cost_stem": (_I, [_P, _P, _SZ, _I, _I, _I, _P, c_float_p, c_float_p, _P, _I, _P, _I, _I, _P, _P, _P]),
    "aarmvs_cost_stem_tc": (_I, [_P, _P, _SZ, _I, _I, _I, _P, _P, c_float_p, _P, _I, _P, _I, _I, _P, _P, _P, _I, _P]),
    "aarmvs_cost_stem_tc_max_slots": (_I, []),
    "aarmvs_warp_cost": (_I, [_P, _P, _SZ, _I, _I, _I, _P, _P, _I, _P, _I, _I, _P, _P]),
    "aarmvs_omega_stats": (_I, [_I, _P, _I, _I, _I, _I, c_float_p, _P, _P]),
    "aarmvs_aggregate": (_I, [_P, _P, _SZ, _I, _I, _I, _P, c_float_p, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "aarmvs_split_planes": (_I, [_SRCP, _I, _I, _P, _P]),
    "aarmvs_cell_state_floats": (_SZ, [_I, _I, _I]),
    "aarmvs_convlstm_cell": (_I, [_SRCP, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "aarmvs_convlstm_cell_tc": (_I, [_PLSP, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "aarmvs_deconv2x_tc": (_I, [_PLSP, _I, _I, _P, c_float_p, _P, _P, _P, _P, _I, _P, _P]),
    "aarmvs_deconv2x": (_I, [_SRCP, _I, _I, c_float_p, c_float_p, _P, _P, _I, _P, _P]),
    "aarmvs_head_wta": (_I, [_SRCP, _I, _I, c_float_p, c_float_p, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "aarmvs_fuse_view_mats": (_I, []),
    "aarmvs_fuse_view": (_I, [_P, _P, _P, _I, _I, _I, _P, _F, _P, _P, _P, _P, _P]),
    "aarmvs_head_wta_tc": (_I, [_P, _I, _I, _P, c_float_p, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "aarmvs_next_plane": (_I, [_P, _P, _I, _P]),
    "aarmvs_loss_grad": (_I, [_P, _P, _P, _F, _I, _SZ, _P, _P]),
    # backward pass (fp32 NHWC primitives, csrc/backward*.cu)
    "aarmvs_conv_nhwc": (_I, [_SEGP, _I, _I, _I, _SEGP, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "aarmvs_conv_bwd_weight": (_I, [_SEGP, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "aarmvs_gn_sums": (_I, [_P, _SZ, _I, _I, _P, _P]),
    "aarmvs_gn_relu_apply": (_I, [_P, _SZ, _I, _I, _P, _P, _P, _I, _P, _P]),
    "aarmvs_gn_relu_bwd": (_I, [_I, _P, _P, _SZ, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "aarmvs_lstm_gates_bwd": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _I, _P]),
    "aarmvs_lstm_gates_bwd_act": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _SZ, _I, _P, _P, _P]),
    "aarmvs_split_planes_cat": (_I, [_SEGP, _I, _I, _I, _P, _P]),
    "aarmvs_conv_dw_tc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "aarmvs_split_planes_scaled": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "aarmvs_conv_dx_tc": (_I, [_P, _I, _I, _I, _SEGP, _I, _I, _P, _P, _P]),
    "aarmvs_maxpool2": (_I, [_P, _I, _I, _I, _P, _P]),
    "aarmvs_maxpool2_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "aarmvs_aggregate_bwd": (_I, [_P, _P, _P, _I, _SZ, _P, _P, _P]),
    "aarmvs_omega_bwd": (_I, [_I, _P, _P, _I, _I, _I, c_float_p, _P, _P, _P, _P, _P]),
    "aarmvs_warp_bwd": (_I, [_P, _P, _SZ, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "aarmvs_feat_assemble_bwd": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "aarmvs_deform_sample": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "aarmvs_deform_bwd": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "aarmvs_snapshot_state": (_I, [_P, _P, _I, _I, _P, _SZ, _I, _P, _P]),
    "aarmvs_adam_step": (_I, [_P, _P, _P, _P, _SZ, _F, _F, _F, _F, _F, _I, _P]),
    "aarmvs_tc_mma_bench": (_I, [_I, _I, _I, _I, _I, _P, _P]),
    "aarmvs_tc_selftest": (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
}

# exported only by -DAARMVS_DEBUG builds (scripts/cell_timeline.py, scripts/stem_ablation.py); bound when present
DEBUG_PROTOTYPES = {
    "aarmvs_debug_profile_buffer": (None, [_P]),
    "aarmvs_debug_flags": (None, [_I]),
    "aarmvs_debug_cell_flags": (None, [_I]),
    "aarmvs_debug_stem_profile_buffer": (None, [_P]),
}

_lib = None


def library_path():
    """In-tree library; AARMVS_LIB overrides it (kernel tuning experiments with differently compiled variants)."""
    return os.environ.get("AARMVS_LIB") or _build.LIB


def load():
    """Make sure the in-tree library matches csrc/ (digest stamp; rebuilds with nvcc when it does not, raises when that is
    impossible -- a stale library never runs silently), load it and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if path == _build.LIB:
        _build.build_library()
    elif not os.path.exists(path):
        raise RuntimeError(f"AARMVS_LIB={path} does not exist")
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI drift between the header and the library
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in DEBUG_PROTOTYPES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    if lib.aarmvs_version() != ABI_VERSION:
        raise RuntimeError("libaarmvs_sm100.so: ABI version mismatch")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        kind = "rejected argument" if code == ERR_ARG else f"cudaError {code}"
        raise RuntimeError(f"{what} failed: {kind}")


def host_floats(values):
    """A ctypes float array (host memory) from a flat python/torch sequence."""
    values = [float(v) for v in values]
    return (ctypes.c_float * len(values))(*values)
