"""ctypes binding of the C-ABI in include/dmmfods_b200.h (libdmmfods_b200.so).

The library is built in-tree by `__graft_entry__.build()` / `python -m dmmfods_b200.build`.
There is NO fallback: if the shared library is missing, or a call returns an error code, a
RuntimeError is raised.  Nothing here touches oracle/ or any CPU implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMM_B200_LIB: an experiment build of the SAME sources (dmmfods_b200/build.py, DMM_BUILD_TAG); never a different implementation
LIB_PATH = os.environ.get("DMM_B200_LIB") or os.path.join(_HERE, "libdmmfods_b200.so")

MAX_SRC = 4
MAX_TAPS = 32
STATS_SLOTS = 8

c_void_p, c_int32, c_int64, c_int8, c_float, c_double = (C.c_void_p, C.c_int32, C.c_int64, C.c_int8, C.c_float,
                                                         C.c_double)


class View(C.Structure):
    _fields_ = [("ptr", c_void_p), ("C", c_int32), ("W", c_int32), ("H", c_int32), ("B", c_int32),
                ("sw", c_int64), ("sh", c_int64), ("sb", c_int64)]


class Bn(C.Structure):
    _fields_ = [("stats", c_void_p), ("stats_ld", c_int32), ("stats_off", c_int32), ("count", c_double),
                ("rep", c_double), ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p),
                ("running_var", c_void_p), ("save_mean", c_void_p), ("save_invstd", c_void_p),
                ("eps", c_float), ("momentum", c_float), ("training", c_int32)]


class Igemm(C.Structure):
    _fields_ = [("src", View * MAX_SRC), ("num_src", c_int32), ("num_taps", c_int32),
                ("tap_src", c_int8 * MAX_TAPS), ("tap_dy", c_int8 * MAX_TAPS), ("tap_dx", c_int8 * MAX_TAPS),
                ("weights", c_void_p), ("ktot", c_int64), ("n_rows", c_int32), ("kwidth", c_int32),
                ("W", c_int32), ("H", c_int32), ("B", c_int32), ("tile_w", c_int32), ("N", c_int32),
                ("n_tile", c_int32), ("out", c_void_p), ("out_mode", c_int32), ("ldo", c_int64),
                ("coff", c_int32), ("out_sy", c_int32), ("out_sx", c_int32), ("out_py", c_int32),
                ("out_px", c_int32), ("OH", c_int32), ("OW", c_int32), ("stats", c_void_p),
                ("stats_ld", c_int32), ("stats_off", c_int32),
                ("bnb_x", c_void_p), ("bnb_ldx", c_int64), ("bnb_gamma", c_void_p), ("bnb_beta", c_void_p),
                ("bnb_mean", c_void_p), ("bnb_invstd", c_void_p), ("bnb_sums", c_void_p), ("bnb_sums_ld", c_int32),
                ("bnb_sums_off", c_int32), ("pro_enable", c_int32), ("fold_kw", c_int32), ("pro_bn", Bn),
                ("dtype", c_int32), ("epi_relu", c_int32), ("epi_bias", c_void_p)]


WG_MAX_A = 8
WG_MAX_B = 25


class WgSlot(C.Structure):
    _fields_ = [("src", c_int8), ("dy", c_int8), ("dx", c_int8), ("pad_", c_int8), ("ch0", c_int32), ("out0", c_int32)]


class Wgrad(C.Structure):
    _fields_ = [("a_src", View * MAX_SRC), ("b_src", View * MAX_SRC), ("num_a_src", c_int32), ("num_b_src", c_int32),
                ("a", WgSlot * WG_MAX_A), ("b", WgSlot * WG_MAX_B), ("num_a", c_int32), ("num_b", c_int32),
                ("n_tile", c_int32), ("ya", c_int32), ("yb", c_int32), ("a_step", c_int32), ("b_step", c_int32),
                ("W", c_int32), ("H", c_int32), ("B", c_int32), ("kpx", c_int32), ("tile_w", c_int32),
                ("splits", c_int32), ("dw", c_void_p), ("ld", c_int64), ("pro_enable", c_int32), ("pad_", c_int32),
                ("pro_gamma", c_void_p), ("pro_beta", c_void_p), ("pro_mean", c_void_p), ("pro_invstd", c_void_p)]


class BnApply(C.Structure):
    _fields_ = [("x", c_void_p), ("ldx", c_int64), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("C", c_int32), ("bn", Bn), ("pool", c_int32), ("y", c_void_p), ("ldy", c_int64),
                ("ystats", c_void_p), ("ystats_ld", c_int32), ("ystats_off", c_int32), ("argmax", c_void_p), ("ldarg", c_int64)]


class BnBwd(C.Structure):
    _fields_ = [("sums", c_void_p), ("sums_ld", c_int32), ("sums_off", c_int32), ("count", c_double),
                ("gamma", c_void_p), ("beta", c_void_p), ("save_mean", c_void_p), ("save_invstd", c_void_p),
                ("dgamma", c_void_p), ("dbeta", c_void_p)]


class BnBwdArgs(C.Structure):
    _fields_ = [("x", c_void_p), ("ldx", c_int64), ("g", c_void_p), ("ldg", c_int64), ("g_is_f32", c_int32),
                ("gmode", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32),
                ("bn", BnBwd), ("out", c_void_p), ("ldo", c_int64), ("out_mode", c_int32), ("dz_out", c_void_p),
                ("lddz", c_int64), ("argmax", c_void_p), ("ldarg", c_int64), ("out_gw", c_int32), ("pad_", c_int32),
                ("out_plane", c_int64), ("fin_k", c_void_p), ("fin_ctr", c_void_p)]


class Head(C.Structure):
    _fields_ = [("u", c_void_p), ("ldu", c_int64), ("Cu", c_int32), ("x1", c_void_p), ("C1", c_int32),
                ("x2", c_void_p), ("C2", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("bn_u", Bn), ("bn_x", Bn), ("out", c_void_p), ("ldo", c_int64)]


class HeadBwd(C.Structure):
    _fields_ = [("u", c_void_p), ("ldu", c_int64), ("Cu", c_int32), ("x1", c_void_p), ("C1", c_int32),
                ("x2", c_void_p), ("C2", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("g", c_void_p), ("ldg", c_int64), ("bn_u", BnBwd), ("bn_x", BnBwd), ("du", c_void_p),
                ("lddu", c_int64)]


GATHER_MAX = 56


class GradGather(C.Structure):
    _fields_ = [("src", c_void_p * GATHER_MAX), ("ld", c_int64 * GATHER_MAX), ("plane", c_int64 * GATHER_MAX),
                ("gw", c_int32), ("pad0_", c_int32), ("nsrc", c_int32), ("nk", c_int32),
                ("k1", c_void_p * GATHER_MAX), ("k2", c_void_p * GATHER_MAX), ("x", c_void_p), ("ldx", c_int64),
                ("mean", c_void_p), ("rows", c_int64), ("C", c_int32), ("pad_", c_int32), ("out", c_void_p), ("ldo", c_int64)]


# name -> (restype, argtypes); every symbol declared in include/dmmfods_b200.h
SIGNATURES = {
    "dmm_last_error": (C.c_char_p, []),
    "dmm_version": (C.c_int, []),
    "dmm_device_ok": (C.c_int, []),
    "dmm_sizeof": (C.c_int, [C.c_int]),
    "dmm_conv_igemm": (C.c_int, [C.POINTER(Igemm), c_void_p]),
    "dmm_conv_wgrad": (C.c_int, [C.POINTER(Wgrad), c_void_p]),
    "dmm_pack_weights": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                   C.POINTER(c_int32), c_int64, c_int64, c_void_p]),
    "dmm_unpack_wgrad": (C.c_int, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p, c_int32,
                                   C.POINTER(c_int32), c_int64, c_int64, c_int32, c_void_p]),
    "dmm_pack_weights_batched": (C.c_int, [c_void_p, c_int32, c_void_p]),
    "dmm_unpack_wgrad_batched": (C.c_int, [c_void_p, c_int32, c_void_p]),
    "dmm_bn_fold_batched": (C.c_int, [c_void_p, c_int32, c_void_p]),
    "dmm_pack_weights_work": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "dmm_unpack_wgrad_work": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "dmm_bn_relu_apply": (C.c_int, [C.POINTER(BnApply), c_void_p]),
    "dmm_bn_relu_bwd_reduce": (C.c_int, [C.POINTER(BnBwdArgs), c_void_p]),
    "dmm_bn_relu_bwd_apply": (C.c_int, [C.POINTER(BnBwdArgs), c_void_p]),
    "dmm_bn_relu_bwd_contrib": (C.c_int, [C.POINTER(BnBwdArgs), c_void_p]),
    "dmm_bn_bwd_finalize": (C.c_int, [C.POINTER(BnBwd), c_int32, c_void_p, c_void_p]),
    "dmm_grad_gather": (C.c_int, [C.POINTER(GradGather), c_void_p]),
    "dmm_im2col_7x7s2": (C.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_int32, c_void_p]),
    "dmm_nchw_stats": (C.c_int, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int32, c_int32, c_void_p]),
    "dmm_head_input": (C.c_int, [C.POINTER(Head), c_void_p]),
    "dmm_head_input_bwd_reduce": (C.c_int, [C.POINTER(HeadBwd), c_void_p]),
    "dmm_head_input_bwd_apply": (C.c_int, [C.POINTER(HeadBwd), c_void_p]),
    "dmm_nchw_to_nhwc_bf16": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dmm_unfold_w7s2": (C.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dmm_dlogits_im2col": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dmm_dlogits_unfold_w": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dmm_rows_f32_to_bf16": (C.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "dmm_bce_logits": (C.c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "dmm_bn_relu_apply_f32": (C.c_int, [C.POINTER(BnApply), c_void_p]),
    "dmm_head_input_f32": (C.c_int, [C.POINTER(Head), c_void_p]),
    "dmm_im2col_7x7s2_f32": (C.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "dmm_pack_weights_work_f32": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "dmm_lidar_splat": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dmm_lidar_splat_batched": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dmm_heatmap_boxes_batched": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dmm_lidar_pool": (C.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "dmm_heatmap_boxes": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dmm_pool_kxk": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dmm_step_metrics": (C.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_float, c_void_p, c_void_p]),
    "dmm_adam_flat": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float,
                                c_float, c_float, c_int32, c_void_p]),
}

_lib = None


def load():
    """dlopen the in-tree library and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "dmmfods_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().dmm_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("dmmfods_b200: %s failed (rc=%d): %s" % (what, rc, last_error()))
