"""ctypes binding of libsta_b200.so (C ABI: include/sta_b200.h).

PyTorch is used only as the owner of device memory and streams: every call passes raw
device pointers (`tensor.data_ptr()`) and the current CUDA stream handle.  There is no
CPU fallback: if the shared library is missing this module raises at import of `lib()`.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STA_B200_LIB") or os.path.join(_HERE, "csrc", "libsta_b200.so")  # env override: A/B timing of builds

EPI_BF16, EPI_GELU, EPI_F32, EPI_ROPE, EPI_PIXSHUF, EPI_HEAD = range(6)
PRECISION_BF16, PRECISION_X3 = 0, 1


class StaGemmDesc(Structure):
    _fields_ = [
        ("conv3x3", c_int), ("epi", c_int),
        ("A", c_void_p), ("lda", c_int64),
        ("W", c_void_p), ("ldw", c_int64),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("nimg", c_int), ("H", c_int), ("Wd", c_int), ("Cin", c_int),
        ("bias", c_void_p),
        ("out", c_void_p), ("ldo", c_int64),
        ("out2", c_void_p), ("resid", c_void_p), ("resid2", c_void_p),
        ("relu_main", c_int), ("rowmap_n", c_int),
        ("pos", c_void_p), ("rope_cols", c_int),
        ("ps_k", c_int), ("ps_cout", c_int), ("ps_h", c_int), ("ps_w", c_int),
        ("head_w", c_void_p), ("head_b", c_void_p), ("pts3d", c_void_p), ("conf", c_void_p),
        ("splitk_ws", c_void_p), ("splitk_ws_bytes", c_int64),
        ("split_precision", c_int),
    ]


_lib = None


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsta_b200.so is missing (%s). Build it with `python -m vista_slam_b200.build`; "
            "the STA path has no CPU/PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i, i64 = c_void_p, c_int, c_int64
    L.sta_last_error.restype = c_char_p
    L.sta_last_error.argtypes = []
    L.sta_version.restype = i
    L.sta_device_synchronize.restype = i
    L.sta_create.argtypes = [POINTER(vp)]
    L.sta_create_ex.argtypes = [POINTER(vp), i]
    L.sta_destroy.argtypes = [vp]
    L.sta_destroy.restype = None
    L.sta_load_tensor.argtypes = [vp, c_char_p, vp, POINTER(c_int64), i, i]
    L.sta_missing_tensors.argtypes = [vp]
    L.sta_weight_arena.argtypes = [vp, POINTER(vp), POINTER(c_int64)]
    L.sta_mark_all_loaded.argtypes = [vp]
    L.sta_encode.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    L.sta_decode.argtypes = [vp, vp, vp, vp, vp, i, i, POINTER(vp), POINTER(vp), vp]
    L.sta_head_pose.argtypes = [vp, vp, i, vp, vp, vp]
    L.sta_head_pts.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp, vp, vp]
    L.sta_graph_replays.argtypes = [vp]
    L.sta_graph_replays.restype = i64
    L.sta_regress_pairs.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sta_regress_pairs_begin.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp]
    L.sta_regress_pairs_finish.argtypes = [vp, POINTER(c_int), i, vp, vp, vp, vp, vp, vp, vp]
    L.sta_forward_pairs.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp]
    L.sta_forward_pairs_host.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp]
    L.sta_launch_count.argtypes = [vp]
    L.sta_launch_count.restype = i64
    L.sta_device_bytes.argtypes = [vp]
    L.sta_device_bytes.restype = i64
    L.sta_profile.argtypes = [vp, i]
    L.sta_profile_read.argtypes = [vp, POINTER(ctypes.c_double), POINTER(c_int64), POINTER(ctypes.c_double)]
    L.sta_op_gemm.argtypes = [POINTER(StaGemmDesc), vp]
    L.sta_op_attention.argtypes = [vp, i64, i, vp, i64, i, vp, i64, i, vp, i64, i, i, i, i, i, c_float, i, vp]
    L.sta_op_layernorm.argtypes = [vp, i, i, c_float, vp, vp, vp, vp, vp, vp, i, vp]
    L.sta_op_patch_im2col.argtypes = [vp, i, i, i, i, vp, vp]
    L.sta_op_upsample2x.argtypes = [vp, vp, i, i, i, i, vp]
    L.sta_op_im2col_3x3_s2.argtypes = [vp, vp, i, i, i, i, vp]
    L.sta_op_cast_f32_bf16.argtypes = [vp, vp, i64, i, i, vp]
    L.sta_op_rope2d.argtypes = [vp, vp, i, i, i, vp]
    L.sta_preprocess_shape.argtypes = [i, i, i, i, i, i, POINTER(c_int)]
    L.sta_preprocess_geometry.argtypes = [i, i, i, i, i, i, POINTER(c_int)]
    L.sta_preprocess_coeffs.argtypes = [i, i, POINTER(c_int), POINTER(c_int), POINTER(c_int), i64]
    L.sta_preprocess_rgb8.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, vp]
    L.sta_pointmap_scratch_bytes.argtypes = [i]
    L.sta_pointmap_scratch_bytes.restype = ctypes.c_size_t
    L.sta_pointmap_consumers.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, vp]
    L.sta_depth_scale.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    L.sta_pose_graph_scratch_bytes.argtypes = [i, i, i]
    L.sta_pose_graph_scratch_bytes.restype = ctypes.c_size_t
    L.sta_pose_graph_lm_step.argtypes = [vp, i, vp, vp, vp, i, vp, i, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp,
                                         vp, vp]
    _lib = L
    return L


def check(rc, what="sta call"):
    if rc != 0:
        msg = lib().sta_last_error()
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
