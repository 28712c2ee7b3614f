"""ctypes binding of libhfr.so (include/hfr.h).  No CPU fallback: if the library is missing this module fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhfr.so")

PREC = {"fp32": 0, "tf32": 1, "bf16": 2}
IN_F32, IN_U8 = 0, 1
FLAG_BGR, FLAG_MEAN_IMAGENET, FLAG_MEAN_VGGFACE2, FLAG_SCALE_PM1, FLAG_L2NORM, FLAG_CUDA_GRAPH = 1, 2, 4, 8, 16, 32

# every symbol include/hfr.h declares: (restype, argtypes)
_vp, _i, _i64, _f, _cp = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_char_p
SIGNATURES = {
    "hfr_last_error": (_cp, []),
    "hfr_version": (_i, []),
    "hfr_launch_count": (_i64, []),
    "hfr_model_load": (_i, [_cp, _cp, _cp, _cp, _f, _i, _i, _i, C.POINTER(_vp)]),
    "hfr_model_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "hfr_model_plan_json": (_i64, [_vp, _cp, _i64]),
    "hfr_model_layer_weights": (_i64, [_vp, _i, _vp, _i64, _vp, _i64]),
    "hfr_model_forward": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_vp), _vp]),
    "hfr_model_forward_host": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_vp), _vp]),
    "hfr_model_submit_host": (_i, [_vp, _i, _vp, _i, _i, _i, C.POINTER(_vp), _vp]),
    "hfr_model_wait_host": (_i, [_vp, _i]),
    "hfr_model_set_keep_activations": (_i, [_vp, _i]),
    "hfr_model_debug_layer": (_i64, [_vp, _i, _i, _vp, _vp]),
    "hfr_model_set_layer_timing": (_i, [_vp, _i]),
    "hfr_model_get_layer_times": (_i, [_vp, _vp, C.POINTER(_i)]),
    "hfr_model_free": (None, [_vp]),
    "hfr_crop_resize_u8": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp]),
    "hfr_resize_pil_u8": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "hfr_debug_gemm_tile_choice": (_i, [_i64, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "hfr_debug_knn_plan": (_i, [_i64, _i64, C.POINTER(_i), C.POINTER(_i)]),
    "hfr_debug_gemm_pair_config": (_i, [_i64, _i, _i, _i, _i, _i, _i] + [C.POINTER(_i)] * 6),
    "hfr_pairwise_dist": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp, _vp, _vp, C.c_float, _vp, _i, _vp]),
    "hfr_age_gender_post": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "hfr_l2_normalize": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "hfr_knn_create": (_i, [_i, _i, _i, C.POINTER(_vp)]),
    "hfr_knn_set_gallery": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "hfr_knn_query": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "hfr_knn_query_host": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "hfr_knn_merge": (_i, [_vp, _i, _i64, _i, _vp, _i, _vp]),
    "hfr_knn_query_partial": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "hfr_knn_merge_certify": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "hfr_knn_query_exact": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _vp]),
    "hfr_knn_merge_listed": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "hfr_knn_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "hfr_knn_debug_candidates": (_i64, [_vp, _vp, _vp, _i64]),
    "hfr_knn_free": (None, [_vp]),
    "hfr_mtcnn_load": (_i, [_cp, _i, C.POINTER(_vp)]),
    "hfr_mtcnn_out_shape": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "hfr_mtcnn_run": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hfr_mtcnn_free": (None, [_vp]),
    "hfr_op_dwconv3x3": (_i, [_vp, _vp, _vp, _vp] + [_i] * 12 + [_vp]),
    "hfr_op_gemm_bias_act": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "hfr_op_gemm_pair": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hfr_op_stem_conv": (_i, [_vp, _i, _vp, _vp, _vp] + [_i] * 15 + [_vp]),
    "hfr_op_stem_conv_tc": (_i, [_vp, _vp, _vp, _vp] + [_i] * 13 + [_vp]),
    "hfr_op_conv2d_window": (_i, [_vp, _vp, _vp, _vp] + [_i] * 13 + [_vp]),
    "hfr_op_conv2d": (_i, [_vp, _vp, _vp, _vp, _vp] + [_i] * 15 + [_vp]),
    "hfr_op_maxpool": (_i, [_vp, _vp] + [_i] * 13 + [_vp]),
}


class HfrError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python hse_facerec_tf_b200/build.py`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return lib.hfr_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Maps hfr_status to the exceptions the reference's callers would see from TF / sklearn."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc == -4:
        raise KeyError(msg)          # graph.get_tensor_by_name on an unknown name
    if rc in (-1, -3, -5):
        raise ValueError(msg)
    if rc == -2:
        raise FileNotFoundError(msg)
    raise HfrError(f"[hfr status {rc}] {msg}")
