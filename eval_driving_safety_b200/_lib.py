"""ctypes loader for libb2attack.so (the C ABI declared in include/b2attack.h).

There is NO fallback: if the library is missing the import of any op fails
loudly, and every op refuses non-CUDA tensors.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2attack.so")

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)
_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/b2attack.h one to one
SIGNATURES = {
    "b2_version": (c_int, []),
    "b2_last_error": (ctypes.c_char_p, []),
    "b2_set_flag": (c_int, [ctypes.c_char_p, c_int]),
    "b2_pgd_update": (c_int, [_PP, _PP, _PP, _PP, c_int, c_int, c_int, c_i64, c_f32, c_f32, c_int,
                              _FP, _FP, _FP, _FP, c_vp]),
    "b2_pgd_update_l2_workspace_bytes": (c_i64, [c_int]),
    "b2_pgd_update_l2": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_f32, c_f32, c_int,
                                 _FP, _FP, _FP, _FP, c_vp, c_vp]),
    "b2_patch_apply": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, _IP, c_int, c_vp]),
    "b2_patch_update": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_f32, c_f32, _FP, _FP, c_vp, c_vp]),
    "b2_patch_apply_dev": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "b2_patch_update_dev": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_f32, c_f32, _FP, _FP, c_vp,
                                    c_vp]),
    "b2_patch_axpy": (c_int, [c_vp, c_vp, c_int, c_int, _FP, _FP, c_vp]),
    "b2_cost_volume_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "b2_cost_volume_bwd_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "b2_cost_volume_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "b2_grid_sample3d_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_i64,
                                     c_int, c_int, c_int, c_vp]),
    "b2_grid_sample2d_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_i64,
                                     c_int, c_int, c_int, c_vp]),
    "b2_grid_plan_count": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_i64, c_int, c_vp]),
    "b2_grid_plan_fill": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_i64,
                                  c_int, c_vp]),
    "b2_grid_plan_sort": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "b2_grid_sample_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp]),
    "b2_lift_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_i64, c_int,
                            c_vp]),
    "b2_conv3d": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_int, c_vp]),
    "b2_conv3d_fusion_caps": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "b2_conv3d_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_vp]),
    "b2_conv2d": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_int, c_int, c_vp]),
    "b2_conv2d_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2_conv2d_stat_rows": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "b2_conv2d_first_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "b2_conv2d_first_dgrad": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "b2_channel_concat": (c_int, [c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    "b2_channel_split": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_vp]),
    "b2_add_prefix": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "b2_conv3d_c1_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "b2_conv3d_c1_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "b2_conv3d_c1_dgrad": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "b2_groupnorm_workspace_bytes": (c_i64, [c_int, c_int]),
    "b2_groupnorm_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int,
                                 c_f32, c_int, c_vp, c_vp]),
    "b2_groupnorm_fwd_ext": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int,
                                     c_f32, c_int, c_vp, c_int, c_vp, c_vp]),
    "b2_groupnorm_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64,
                                 c_int, c_int, c_vp, c_vp]),
    "b2_groupnorm_bwd_ext": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64,
                                     c_int, c_int, c_vp, c_int, c_vp, c_vp]),
    "b2_bev_pool_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "b2_bev_pool_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "b2_depth_head_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_f32,
                                  c_vp]),
    "b2_depth_head_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "b2_depth_head_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32,
                                  c_f32, c_vp, c_vp]),
    "b2_roi_align_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "b2_roi_align_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_int, c_vp]),
    "b2_roi_align_pyramid_fwd": (c_int, [_PP, _IP, _IP, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp]),
    "b2_roi_align_pyramid_bwd": (c_int, [c_vp, c_vp, _PP, _IP, _IP, c_int, c_int, c_int, c_f32, c_int, c_vp]),
}

_lib = None


def load():
    """dlopen libb2attack.so and bind every symbol of the header; raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb2attack.so is not built (%s). Run `python -m eval_driving_safety_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def set_flag(name, value):
    """Kernel-variant switch (include/b2attack.h b2_set_flag); value None = back to the default."""
    check(load().b2_set_flag(name.encode(), -1 if value is None else int(value)), "set_flag(%s)" % name)


def last_error():
    return load().b2_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("libb2attack %s failed (code %d): %s" % (what, rc, last_error()))


def f32_array(values):
    if values is None:
        return None
    arr = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return arr


def ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
