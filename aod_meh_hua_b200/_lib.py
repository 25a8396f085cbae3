"""ctypes binding of libmehhua.so (the C ABI declared in include/mehhua.h).

The product path has no CPU fallback: importing this module without the built library, or
calling into it without a B200-class device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmehhua.so")

MAX_LEVELS = 8
MAX_DETS = 256
MAX_NMS_PRE = 4096
ABI_VERSION = 6

E_ARG, E_WORKSPACE, E_CUDA, E_NODEVICE = -1, -2, -3, -4
ST_PAIR_OVERFLOW, ST_SELECT_SLOWPATH, ST_BAD_ALPHA, ST_CAPTURE_FALLBACK = 1, 2, 4, 8
MODE_NMS, MODE_ALL = 0, 1


class MehhuaError(RuntimeError):
    pass


class Level(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("deltas", C.c_void_p), ("lam", C.c_void_p),
                ("anchors", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("A", C.c_int32),
                ("reserved", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("head", C.c_int32), ("c_out", C.c_int32), ("num_levels", C.c_int32),
                ("nms_pre", C.c_int32), ("score_thr", C.c_float), ("nms_iou", C.c_float),
                ("max_per_img", C.c_int32), ("fg_thr", C.c_float), ("obj_thr", C.c_float),
                ("cluster_iou", C.c_float), ("lambda_scale", C.c_float), ("lambda_eps", C.c_float),
                ("use_lambda", C.c_int32), ("n_samples", C.c_int32), ("agg_object", C.c_int32),
                ("agg_scale", C.c_int32), ("agg_class", C.c_int32), ("cls_w", C.c_int32),
                ("means", C.c_float * 4), ("stds", C.c_float * 4), ("wh_ratio_clip", C.c_float),
                ("rescale", C.c_int32), ("pair_cap", C.c_int32), ("mode", C.c_int32),
                ("activation", C.c_int32), ("seed", C.c_uint64)]


BUFFER_FIELDS = ["score_rows", "lam_rows", "boxes", "topk_idx", "row_max", "row_argmax", "level_fg",
                 "dets", "det_labels", "det_flat", "n_det", "n_obj", "pair_row", "pair_obj",
                 "pair_cls", "pair_off", "lam_mean", "pair_unc", "image_scores", "level_maxconf", "group_unc", "pair_avg"]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


LevelArray = Level * MAX_LEVELS

# every symbol include/mehhua.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_CFG, _LV, _BUF = C.POINTER(Config), C.POINTER(Level), C.POINTER(Buffers)
SYMBOLS = {
    "mehhua_abi_version": (C.c_int, []),
    "mehhua_last_cuda_error": (C.c_char_p, []),
    "mehhua_launch_count": (C.c_uint64, []),
    "mehhua_rows_per_image": (C.c_int64, [_CFG, _LV]),
    "mehhua_workspace_bytes": (C.c_size_t, [_CFG, _LV, C.c_int32]),
    "mehhua_workspace_init": (C.c_int, [_P, C.c_size_t, _P]),
    "mehhua_read_status": (C.c_int, [_CFG, _LV, C.c_int32, _P, _P, C.POINTER(C.c_uint32)]),
    "mehhua_k1_alpha_topk": (C.c_int, [_CFG, _LV, C.c_int32, _P, _P, _BUF, _P, C.c_size_t, _P]),
    "mehhua_nms_objects": (C.c_int, [_CFG, _LV, C.c_int32, _BUF, _P, C.c_size_t, _P]),
    "mehhua_iou_pairs": (C.c_int, [_CFG, _LV, C.c_int32, _BUF, _P, C.c_size_t, _P]),
    "mehhua_k2_dirichlet_epi": (C.c_int, [_CFG, _LV, C.c_int32, _P, _P, _P, _BUF, _P, C.c_size_t, _P]),
    "mehhua_k3_hua": (C.c_int, [_CFG, _LV, C.c_int32, _BUF, _P, C.c_size_t, _P]),
    "mehhua_score_batch": (C.c_int, [_CFG, _LV, C.c_int32, _P, _P, _P, _BUF, _P, C.c_size_t, _P]),
    "mehhua_all_fg_rows": (C.c_int, [_CFG, _LV, C.c_int32, _BUF, _P, C.c_size_t, _P]),
    "mehhua_score_batch_all": (C.c_int, [_CFG, _LV, C.c_int32, _P, _BUF, _P, C.c_size_t, _P]),
    "mehhua_stage_timing_begin": (C.c_int, [C.c_int32]),
    "mehhua_stage_timing_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "mehhua_pool_topk_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "mehhua_pool_topk_workspace_bytes_k": (C.c_size_t, [C.c_int64, C.c_int32]),
    "mehhua_k4_pool_topk": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "mehhua_mi_workspace_bytes": (C.c_size_t, [_P, C.c_int32, C.c_int32]),
    "mehhua_mi_score_batch": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "mehhua_host_pin": (C.c_int, [_P, C.c_size_t]),
    "mehhua_host_unpin": (C.c_int, [_P]),
    "mehhua_host_is_pinned": (C.c_int, [_P]),
    "mehhua_debug_capture_counts": (C.c_int, [_CFG, _LV, C.c_int32, _P, _P, C.POINTER(C.c_int32)]),
    "mehhua_debug_philox": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mehhua_host_ctx_create": (C.c_int, [_CFG, _LV, C.c_int32, C.POINTER(_P)]),
    "mehhua_host_ctx_destroy": (None, [_P]),
    "mehhua_score_batch_host": (C.c_int, [_P, _LV, C.c_int32, _P, _P, _P, _P, C.POINTER(C.c_uint32)]),
}

_lib = None
_variants = {}
# A/B builds of the same sources (`make -C aod_meh_hua_b200/csrc variants`; experiments and tests only)
BF16_STAGE_LIB_PATH = os.path.join(_HERE, "libmehhua_bf16stage.so")    # K2 stages its draws as bfloat16
PHILOX10_LIB_PATH = os.path.join(_HERE, "libmehhua_philox10.so")        # K2's sampler runs Philox4x32-10


def _open(path: str) -> C.CDLL:
    if not os.path.isfile(path):
        raise MehhuaError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C aod_meh_hua_b200/csrc all` - there is no CPU fallback for the scoring path")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mehhua_abi_version() != ABI_VERSION:
        raise MehhuaError(f"{os.path.basename(path)} ABI {lib.mehhua_abi_version()} != binding ABI {ABI_VERSION}")
    return lib


def load() -> C.CDLL:
    """Load libmehhua.so (once).  Raises MehhuaError when it has not been built."""
    global _lib
    if _lib is None:
        _lib = _open(os.environ.get("MEHHUA_LIB", LIB_PATH))     # MEHHUA_LIB: an A/B build, for experiments
    return _lib


def load_variant(path: str) -> C.CDLL:
    """Load another build of the library (same ABI), e.g. PHILOX10_LIB_PATH."""
    if path not in _variants:
        _variants[path] = _open(path)
    return _variants[path]


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.mehhua_last_cuda_error().decode(errors="replace")
    name = {E_ARG: "MEHHUA_E_ARG", E_WORKSPACE: "MEHHUA_E_WORKSPACE", E_CUDA: "MEHHUA_E_CUDA",
            E_NODEVICE: "MEHHUA_E_NODEVICE"}.get(rc, str(rc))
    raise MehhuaError(f"{what} failed with {name}: {msg}")
