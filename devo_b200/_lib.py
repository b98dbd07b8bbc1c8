"""ctypes binding of libdevo_b200.so (the C ABI declared in include/devo_b200.h).

The product path is CUDA only: if the library is missing or an op is called with
non-CUDA tensors this module raises -- there is no CPU fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEVO_B200_LIB") or os.path.join(_HERE, "lib", "libdevo_b200.so")

F16, BF16, F32, F64 = 0, 1, 2, 3
_DTYPES = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32, torch.float64: F64}

_c = ctypes
_vp, _i, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); must list every symbol include/devo_b200.h declares
SIGNATURES = {
    "devo_abi_version": (_i, []),
    "devo_last_error": (_c.c_char_p, []),
    "devo_launch_count": (_c.c_uint64, []),
    "devo_copy_bytes": (_i, [_vp, _vp, _sz, _vp]),
    "devo_corr_forward": (_i, [_vp] * 6 + [_i] * 10 + [_vp]),
    "devo_corr_backward": (_i, [_vp] * 8 + [_i] * 10 + [_vp]),
    "devo_corr_backward_pm": (_i, [_vp] * 8 + [_i] * 6 + [_vp]),
    "devo_patchify_forward": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "devo_patchify_backward": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "devo_pyramid_pack": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "devo_pyramid_pack2": (_i, [_vp] * 3 + [_i] * 6 + [_vp]),
    "devo_gmap_pack": (_i, [_vp, _vp] + [_i] * 4 + [_vp]),
    "devo_patch_gather": (_i, [_vp] * 8 + [_i] * 8 + [_vp]),
    "devo_corr_lookup_fused": (_i, [_vp] * 6 + [_i] * 5 + [_vp]),
    "devo_corr_lookup_fused_ld": (_i, [_vp] * 6 + [_i] * 6 + [_vp]),
    "devo_pyramid_pack_split": (_i, [_vp] * 3 + [_i] * 5 + [_vp]),
    "devo_gmap_pack_split": (_i, [_vp] * 3 + [_i] * 3 + [_vp]),
    "devo_corr_lookup_fused_split": (_i, [_vp] * 8 + [_i] * 5 + [_vp]),
    "devo_graph_plan_workspace": (_sz, [_i]),
    "devo_graph_plan": (_i, [_vp, _vp, _i, _i64, _i64] + [_vp] * 8 + [_sz, _vp]),
    "devo_neighbors": (_i, [_vp] * 4 + [_i, _vp, _sz, _vp]),
    "devo_ba_workspace": (_sz, [_i, _i]),
    "devo_ba_forward": (_i, [_vp] * 9 + [_i] * 7 + [_vp, _sz, _vp, _vp]),
    "devo_ba_forward_planned": (_i, [_vp] * 9 + [_i] * 7 + [_vp] * 4 + [_vp, _sz, _vp, _vp]),
    "devo_ba_prepare": (_i, [_vp, _sz, _i, _i, _vp, _vp]),
    "devo_ba_forward_prepared": (_i, [_vp] * 9 + [_i] * 7 + [_vp] * 4 + [_vp, _sz, _vp, _vp, _vp]),
    "devo_ba_system_doubles": (_sz, [_i]),
    "devo_ba_sharded_accumulate": (_i, [_vp] * 9 + [_i] * 8 + [_vp, _vp, _sz, _vp, _vp]),
    "devo_ba_sharded_solve": (_i, [_vp, _vp] + [_i] * 5 + [_vp, _sz, _vp, _vp]),
    "devo_ba_sharded_solve_peer": (_i, [_vp, _vp, _i, _i, _c.c_uint64] + [_i] * 5 + [_vp, _sz, _vp, _vp]),
    "devo_reproject": (_i, [_vp] * 7 + [_i, _i, _vp]),
    "devo_transform_forward": (_i, [_vp] * 11 + [_i] * 4 + [_vp]),
    "devo_transform_backward": (_i, [_vp] * 12 + [_i] * 4 + [_vp]),
    "devo_glue_layernorm": (_i, [_i, _i] + [_vp] * 6 + [_c.c_float, _vp, _vp, _i, _i, _vp]),
    "devo_glue_gather_mask_cast": (_i, [_i, _vp, _vp, _vp, _i, _i, _vp]),
    "devo_glue_residual_add": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "devo_glue_gated_residual": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "devo_glue_relu_cast": (_i, [_i, _vp, _vp, _i64, _i, _vp]),
    "devo_glue_heads": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "devo_segment_softmax_sum": (_i, [_vp] * 5 + [_i, _vp, _i, _i, _i, _vp]),
    "devo_voxel_workspace": (_sz, [_i]),
    "devo_voxel_normalize": (_i, [_vp, _vp, _c.c_longlong, _i, _i, _vp, _vp, _sz, _vp]),
    "devo_events_to_voxel": (_i, [_vp, _vp, _vp, _vp, _c.c_longlong, _vp, _i, _i, _i, _vp]),
    "devo_gru_workspace": (_sz, [_i, _i]),
    "devo_gru_state_floats": (_sz, [_i]),
    "devo_gru_state_gather": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "devo_gru_update": (_i, [_vp, _vp, _i, _vp, _sz, _vp]),
}
for _n, _a in (("expm", 2), ("logm", 2), ("inv", 2), ("as_matrix", 2), ("projector", 2),
               ("expm_backward", 3), ("logm_backward", 3), ("inv_backward", 3),
               ("mul", 3), ("adj", 3), ("adjT", 3), ("act", 3), ("act4", 3), ("jinv", 3),
               ("mul_backward", 5), ("adj_backward", 5), ("adjT_backward", 5),
               ("act_backward", 5), ("act4_backward", 5)):
    SIGNATURES["devo_lie_" + _n] = (_i, [_i, _i] + [_vp] * _a + [_i64, _vp])


class PyramidStruct(ctypes.Structure):
    """devo_pyramid_t"""
    _fields_ = [("n_levels", _i), ("level", _vp * 4), ("H", _i * 4), ("W", _i * 4), ("scale", _c.c_float * 4)]


class GruWeightsStruct(ctypes.Structure):
    """devo_gru_weights_t"""
    _fields_ = [("W", _vp), ("W0", _vp), ("bias", _vp), ("ln_gamma", _vp), ("ln_beta", _vp), ("ln_eps", _c.c_float),
                ("head_W", _vp), ("head_b", _vp)]


class GruIoStruct(ctypes.Structure):
    """devo_gru_io_t"""
    _fields_ = [("E", _i), ("dim", _i), ("corr_ld", _i), ("corr16", _vp), ("state32", _vp), ("net16", _vp), ("imap16", _vp), ("kk", _vp),
                ("ix", _vp), ("jx", _vp),
                ("perm_kk", _vp), ("gstart_kk", _vp), ("ngroups_kk", _vp), ("gid_kk", _vp), ("max_groups_kk", _i),
                ("perm_ij", _vp), ("gstart_ij", _vp), ("ngroups_ij", _vp), ("gid_ij", _vp), ("max_groups_ij", _i),
                ("net16_out", _vp), ("delta", _vp), ("weight", _vp), ("coords", _vp), ("target32", _vp), ("weight32", _vp), ("tile_local", _i)]


_lib = None


def lib():
    """load (once) and return the ctypes handle; raises if the CUDA library is not built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "devo_b200: %s not found -- build it with `python -m devo_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        if h.devo_abi_version() != 1:
            raise RuntimeError("devo_b200: ABI version mismatch")
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().devo_last_error().decode("utf-8", "replace")
        raise RuntimeError("devo_b200 %s failed (code %d): %s" % (what, rc, msg))


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError("devo_b200: unsupported dtype %s" % t.dtype)


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("devo_b200: expected a CUDA tensor (the hot path has no CPU fallback), got %s" % t.device)


def require_contiguous(**named):
    for n, t in named.items():
        if t is not None and not t.is_contiguous():
            raise RuntimeError("devo_b200: %s must be contiguous" % n)


def require_dtype(t, dtype, name):
    if t.dtype != dtype:
        raise RuntimeError("devo_b200: %s must be %s, got %s" % (name, dtype, t.dtype))


def launch_count():
    return int(lib().devo_launch_count())


def copy_(dst, src):
    """dst.copy_(src) for two contiguous CUDA tensors of the same dtype and size, done by a kernel on the current stream
    (devo_copy_bytes): inside a captured step a memcpy node would run on a copy engine and queue behind host uploads"""
    require_cuda(dst, src)
    if dst.dtype != src.dtype or dst.numel() != src.numel() or not (dst.is_contiguous() and src.is_contiguous()):
        raise RuntimeError("devo_b200: copy_ needs two contiguous tensors of the same dtype and size")
    check(lib().devo_copy_bytes(dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size(), stream_ptr(dst.device)),
          "copy_bytes")
    return dst


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """grow-only per-(device, stream, tag) scratch buffer"""
    key = (device, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf
