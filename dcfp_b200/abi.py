"""ctypes view of the C ABI in include/dcfp_b200.h (libdcfp_b200.so).

Used by the CPU test-suite to check that the library loads and exports every declared symbol, and
by non-torch callers.  The torch product path goes through `dcfp_b200.ops`.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdcfp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "dcfp_b200.h")

F32, BF16 = 0, 1
LABEL_U8, LABEL_I32, LABEL_I64 = 0, 1, 2
NCHW, NHWC = 0, 1


class LayerDesc(ctypes.Structure):
    """struct dcfp_layer_desc"""
    _fields_ = [("x", ctypes.c_void_p), ("dy", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("keys", ctypes.c_void_p), ("S1", ctypes.c_void_p), ("S2", ctypes.c_void_p),
                ("N", ctypes.c_int32), ("C", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("K", ctypes.c_int32), ("dtype", ctypes.c_int32), ("layout", ctypes.c_int32), ("ld", ctypes.c_int32),
                ("affine_mode", ctypes.c_int32), ("hints", ctypes.c_int32)]


class BnDesc(ctypes.Structure):
    """struct dcfp_bn_desc"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("x", "y", "dy", "dx", "gamma", "beta", "mean", "invstd", "running_mean", "running_var",
                                               "scratch", "keys", "S1", "S2", "dgamma", "dbeta")] + \
               [(n, ctypes.c_int32) for n in ("N", "C", "h", "w", "dtype", "relu", "K", "ld")] + \
               [("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("phases", ctypes.c_int32), ("arena_f32", ctypes.c_int32),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64), ("residual", ctypes.c_void_p)]


class GatherDesc(ctypes.Structure):
    """struct dcfp_gather_desc"""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("out_idx", ctypes.c_void_p), ("in_idx", ctypes.c_void_p),
                ("n_out", ctypes.c_int32), ("n_in", ctypes.c_int32), ("I", ctypes.c_int32), ("khw", ctypes.c_int32)]


def declared_symbols(header_path=HEADER_PATH):
    """Names of all functions the public header declares."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcfp_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load(path=LIB_PATH):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError("libdcfp_b200.so is not built (%s); run `python -m dcfp_b200.build`" % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.dcfp_last_error.restype = ctypes.c_char_p
    lib.dcfp_abi_version.restype = i32
    lib.dcfp_launch_count.restype = i64
    lib.dcfp_launch_count.argtypes = [i32]
    lib.dcfp_label_keys.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.dcfp_class_stats.argtypes = [ctypes.POINTER(LayerDesc), vp]
    lib.dcfp_class_stats_grouped.argtypes = [ctypes.POINTER(LayerDesc), i32, vp]
    lib.dcfp_bn_supported.argtypes = [i32, i32, i32, i32, i32]
    lib.dcfp_bn_scratch_bytes.restype = ctypes.c_size_t
    lib.dcfp_bn_scratch_bytes.argtypes = [i32]
    lib.dcfp_bn_workspace_bytes.restype = ctypes.c_size_t
    lib.dcfp_bn_workspace_bytes.argtypes = [i32]
    lib.dcfp_bn_forward.argtypes = [ctypes.POINTER(BnDesc), vp]
    lib.dcfp_bn_backward.argtypes = [ctypes.POINTER(BnDesc), vp]
    lib.dcfp_relu_grad.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    lib.dcfp_eic_update.argtypes = [vp, vp, vp, i32, vp, ctypes.c_float, ctypes.c_float, i32, vp]
    lib.dcfp_eic_update_flat.argtypes = [vp, vp, vp, i32, ctypes.c_float, ctypes.c_float, i32, vp]
    lib.dcfp_reduce_classes.argtypes = [vp, i32, i32, vp, vp]
    lib.dcfp_fold_step.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.dcfp_fold_step2.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.dcfp_thresh_mask.argtypes = [vp, vp, vp, vp, i32, i32, ctypes.POINTER(i64), vp, vp, vp, vp]
    lib.dcfp_channel_gather.argtypes = [vp, vp, vp, i32, vp, i32, i32, i32, i32, vp]
    lib.dcfp_channel_gather_workspace.restype = ctypes.c_size_t
    lib.dcfp_channel_gather_workspace.argtypes = [i32]
    lib.dcfp_channel_gather_grouped.argtypes = [ctypes.POINTER(GatherDesc), i32, i32, vp, ctypes.c_size_t, vp]
    lib.dcfp_bias_comp.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.dcfp_class_balance_weights.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32, ctypes.c_double, vp, vp, vp]
    _lib = lib
    return lib


def last_error():
    return load().dcfp_last_error().decode()
