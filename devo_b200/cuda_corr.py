"""`cuda_corr` -- drop-in for the reference's pybind module of the same name
(devo/altcorr/correlation.cpp:57-62): forward, backward, patchify_forward,
patchify_backward, same argument order and meaning, list-of-tensors results.

Host side only: validation, output allocation and the call into the C ABI
(include/devo_b200.h) on the current CUDA stream.
"""
import os

import torch

from . import _lib

_pack_cache = {}
_USE_CACHE = os.environ.get("DEVO_B200_CORR_CACHE", "1") != "0"
_FORCE_GENERIC = os.environ.get("DEVO_B200_CORR_IMPL", "") == "generic"


def _check_common(fmap1, fmap2, coords, ii, jj):
    _lib.require_cuda(fmap1, fmap2, coords, ii, jj)
    if fmap1.dim() != 5 or fmap2.dim() != 5 or coords.dim() != 5:
        raise RuntimeError("cuda_corr: fmap1 [B,Np,C,P,P], fmap2 [B,Nf,C,H,W], coords [B,E,2,P,P] expected")
    if fmap1.dtype != fmap2.dtype:
        raise RuntimeError("cuda_corr: fmap1/fmap2 dtype mismatch (%s vs %s)" % (fmap1.dtype, fmap2.dtype))
    if fmap1.shape[2] != fmap2.shape[2]:
        raise RuntimeError("cuda_corr: channel mismatch")
    _lib.require_dtype(coords, torch.float32, "coords")
    _lib.require_dtype(ii, torch.int64, "ii")
    _lib.require_dtype(jj, torch.int64, "jj")
    if ii.numel() != coords.shape[1] or jj.numel() != coords.shape[1]:
        raise RuntimeError("cuda_corr: ii/jj length must equal coords.shape[1]")


def pack_pixel_major(fmap, pool=1, out=None):
    """planar [N,C,H,W] -> pixel-major [N,H//pool,W//pool,C] (average pooled); f16/bf16.
    `out`: contiguous destination (e.g. a slot of the pyramid ring buffer) to pack into directly"""
    _lib.require_cuda(fmap)
    fmap = fmap.contiguous()
    N, C, H, W = fmap.shape
    if out is None:
        out = torch.empty(N, H // pool, W // pool, C, dtype=fmap.dtype, device=fmap.device)
    elif (not out.is_contiguous()) or out.dtype != fmap.dtype or out.numel() != N * (H // pool) * (W // pool) * C:
        raise RuntimeError("pack_pixel_major: out must be a contiguous [N,H/pool,W/pool,C] tensor of the input dtype")
    _lib.check(_lib.lib().devo_pyramid_pack(fmap.data_ptr(), out.data_ptr(), _lib.dtype_code(fmap),
                                            N, C, H, W, pool, _lib.stream_ptr(fmap.device)), "pyramid_pack")
    return out


def pack_pixel_major2(fmap, pool, out1, outp):
    """the two levels of a [1, pool] pyramid from one read of the frames (bit-identical to two pack_pixel_major calls);
    returns False -- nothing written -- when the shapes are outside what the fused kernel takes"""
    _lib.require_cuda(fmap, out1, outp)
    N, C, H, W = fmap.shape
    if (pool not in (2, 4, 8) or C % 8 or W % 8 or H % pool or W % pool or pool * 8 * (C + 8) * 2 > 48 * 1024
            or fmap.dtype not in (torch.float16, torch.bfloat16) or not fmap.is_contiguous()
            or not out1.is_contiguous() or not outp.is_contiguous() or out1.dtype != fmap.dtype or outp.dtype != fmap.dtype
            or out1.numel() != N * H * W * C or outp.numel() != N * (H // pool) * (W // pool) * C
            or (fmap.data_ptr() | out1.data_ptr() | outp.data_ptr()) & 15):
        return False
    _lib.check(_lib.lib().devo_pyramid_pack2(fmap.data_ptr(), out1.data_ptr(), outp.data_ptr(), _lib.dtype_code(fmap),
                                             N, C, H, W, pool, _lib.stream_ptr(fmap.device)), "pyramid_pack2")
    return True


def pack_gmap(gmap, out=None):
    """planar [Np,C,P,P] -> [Np,P*P,C]  (`out`: contiguous destination to pack into directly)"""
    _lib.require_cuda(gmap)
    gmap = gmap.contiguous()
    Np, C = gmap.shape[0], gmap.shape[1]
    PP = gmap.shape[2] * gmap.shape[3]
    if out is None:
        out = torch.empty(Np, PP, C, dtype=gmap.dtype, device=gmap.device)
    elif (not out.is_contiguous()) or out.dtype != gmap.dtype or out.numel() != Np * PP * C:
        raise RuntimeError("pack_gmap: out must be a contiguous [Np,P*P,C] tensor of the input dtype")
    _lib.check(_lib.lib().devo_gmap_pack(gmap.data_ptr(), out.data_ptr(), _lib.dtype_code(gmap), Np, C, PP,
                                         _lib.stream_ptr(gmap.device)), "gmap_pack")
    return out


def _cached(t, fn, tag=""):
    if not _USE_CACHE:
        return fn(t)
    # The entry keeps the SOURCE tensor alive: while it is cached its memory cannot be handed to another tensor, so an
    # equal (data_ptr, shape, dtype) can only be the same storage (views share the version counter).  Without that
    # reference a new tensor allocated at a recycled address with the same shape and version hit a stale entry.
    key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t.device, tag)
    hit = _pack_cache.get(key)
    if hit is not None and hit[0] == t._version:
        _pack_cache[key] = _pack_cache.pop(key)          # most recently used last
        return hit[1]
    out = fn(t)
    _pack_cache.pop(key, None)
    while len(_pack_cache) >= 8:
        _pack_cache.pop(next(iter(_pack_cache)))         # evict the least recently used entry
    _pack_cache[key] = (t._version, out, t)
    return out


def lookup_fused(gmap_pm, levels_pm, scales, coords, ii, jj, out=None):
    """fused multi-level lookup on pixel-major buffers.
    gmap_pm [Np,9,C]; levels_pm: list of [Nf,H_l,W_l,C]; coords [E,2,3,3] f32 (level-1 resolution).
    `out`: optional preallocated [E, >= 49*9*L] buffer (row padding beyond 49*9*L is left untouched).
    returns [E, 49*9*L] (dtype of the features), levels interleaved on the last axis exactly like
    torch.stack(corrs, -1).view(1, E, -1) in devo/devo.py:217."""
    L = len(levels_pm)
    E = coords.shape[0]
    Np, _, C = gmap_pm.shape
    if out is None:
        out = torch.empty(E, 49 * 9 * L, dtype=gmap_pm.dtype, device=gmap_pm.device)
    elif out.dim() != 2 or out.shape[0] != E or out.shape[1] < 49 * 9 * L or out.stride(1) != 1:
        raise RuntimeError("cuda_corr.lookup_fused: out must be [E, >= 441*L] with unit inner stride")
    pyr = _lib.PyramidStruct()
    pyr.n_levels = L
    for l, lv in enumerate(levels_pm):
        pyr.level[l] = lv.data_ptr()
        pyr.H[l] = lv.shape[1]
        pyr.W[l] = lv.shape[2]
        pyr.scale[l] = float(scales[l])
    import ctypes
    _lib.check(_lib.lib().devo_corr_lookup_fused_ld(gmap_pm.data_ptr(), ctypes.addressof(pyr), coords.data_ptr(),
                                                    ii.data_ptr(), jj.data_ptr(), out.data_ptr(), out.stride(0),
                                                    _lib.dtype_code(gmap_pm), Np, levels_pm[0].shape[0], C, E,
                                                    _lib.stream_ptr(gmap_pm.device)), "corr_lookup_fused")
    return out


def pack_pixel_major_split(fmap, pool=1):
    """float planar [N,C,H,W] -> two pixel-major half buffers (hi, lo) with a = hi + 2^-11 lo (pooling in float)"""
    _lib.require_cuda(fmap)
    _lib.require_dtype(fmap, torch.float32, "fmap")
    fmap = fmap.contiguous()
    N, C, H, W = fmap.shape
    hi = torch.empty(N, H // pool, W // pool, C, dtype=torch.float16, device=fmap.device)
    lo = torch.empty_like(hi)
    _lib.check(_lib.lib().devo_pyramid_pack_split(fmap.data_ptr(), hi.data_ptr(), lo.data_ptr(), N, C, H, W, pool,
                                                  _lib.stream_ptr(fmap.device)), "pyramid_pack_split")
    return hi, lo


def pack_gmap_split(gmap):
    """float planar [Np,C,P,P] -> (hi, lo) [Np,P*P,C] halves"""
    _lib.require_cuda(gmap)
    _lib.require_dtype(gmap, torch.float32, "gmap")
    gmap = gmap.contiguous()
    Np, C = gmap.shape[0], gmap.shape[1]
    PP = gmap.shape[2] * gmap.shape[3]
    hi = torch.empty(Np, PP, C, dtype=torch.float16, device=gmap.device)
    lo = torch.empty_like(hi)
    _lib.check(_lib.lib().devo_gmap_pack_split(gmap.data_ptr(), hi.data_ptr(), lo.data_ptr(), Np, C, PP,
                                               _lib.stream_ptr(gmap.device)), "gmap_pack_split")
    return hi, lo


def _pyramid_struct(levels_pm, scales):
    pyr = _lib.PyramidStruct()
    pyr.n_levels = len(levels_pm)
    for l, lv in enumerate(levels_pm):
        pyr.level[l] = lv.data_ptr()
        pyr.H[l] = lv.shape[1]
        pyr.W[l] = lv.shape[2]
        pyr.scale[l] = float(scales[l])
    return pyr


def lookup_fused_split(gmap_split, levels_split, scales, coords, ii, jj, out=None):
    """`lookup_fused` for float32 features given as (hi, lo) half pairs (pack_gmap_split / pack_pixel_major_split):
    three tensor-core passes with float accumulation; returns float32 [E, 49*9*L]"""
    import ctypes
    g_hi, g_lo = gmap_split
    L = len(levels_split)
    E = coords.shape[0]
    Np, _, C = g_hi.shape
    if out is None:
        out = torch.empty(E, 49 * 9 * L, dtype=torch.float32, device=g_hi.device)
    elif out.dim() != 2 or out.shape[0] != E or out.shape[1] < 49 * 9 * L or out.stride(1) != 1 or out.dtype != torch.float32:
        raise RuntimeError("cuda_corr.lookup_fused_split: out must be float32 [E, >= 441*L] with unit inner stride")
    p_hi = _pyramid_struct([lv[0] for lv in levels_split], scales)
    p_lo = _pyramid_struct([lv[1] for lv in levels_split], scales)
    _lib.check(_lib.lib().devo_corr_lookup_fused_split(g_hi.data_ptr(), g_lo.data_ptr(), ctypes.addressof(p_hi),
                                                       ctypes.addressof(p_lo), coords.data_ptr(), ii.data_ptr(), jj.data_ptr(),
                                                       out.data_ptr(), out.stride(0), Np, levels_split[0][0].shape[0], C, E,
                                                       _lib.stream_ptr(g_hi.device)), "corr_lookup_fused_split")
    return out


def _split_eligible(fmap1, fmap2, coords, radius):
    return (not _FORCE_GENERIC and fmap1.dtype == torch.float32 and fmap1.shape[0] == 1
            and fmap1.shape[2] in (64, 128) and fmap1.shape[3] == 3 and fmap1.shape[4] == 3 and radius == 3)


def _fast_eligible(fmap1, fmap2, coords, radius):
    return (not _FORCE_GENERIC and fmap1.dtype in (torch.float16, torch.bfloat16) and fmap1.shape[0] == 1
            and fmap1.shape[2] in (64, 128) and fmap1.shape[3] == 3 and fmap1.shape[4] == 3 and radius == 3)


def forward(fmap1, fmap2, coords, ii, jj, radius):
    """cuda_corr.forward -> [corr [B,E,2r+1,2r+1,P,P]]  (dims 2,3 = x-offset, y-offset)"""
    _check_common(fmap1, fmap2, coords, ii, jj)
    fmap1 = fmap1.contiguous()
    fmap2 = fmap2.contiguous()
    coords = coords.contiguous()
    ii = ii.contiguous()
    jj = jj.contiguous()
    B, Np, C, P, _ = fmap1.shape
    _, Nf, _, H, W = fmap2.shape
    E = coords.shape[1]
    D1 = 2 * radius + 1
    if E > 0 and _fast_eligible(fmap1, fmap2, coords, radius):
        lv = _cached(fmap2, lambda t: pack_pixel_major(t[0], 1))
        g = _cached(fmap1, lambda t: pack_gmap(t[0]))
        out = lookup_fused(g, [lv], [1.0], coords[0], ii, jj)
        return [out.view(1, E, D1, D1, P, P)]
    if E > 0 and _split_eligible(fmap1, fmap2, coords, radius):       # float32 features (training): split precision
        lv = _cached(fmap2, lambda t: pack_pixel_major_split(t[0], 1))
        g = _cached(fmap1, lambda t: pack_gmap_split(t[0]))
        out = lookup_fused_split(g, [lv], [1.0], coords[0], ii, jj)
        return [out.view(1, E, D1, D1, P, P)]
    out = torch.empty(B, E, D1, D1, P, P, dtype=fmap1.dtype, device=fmap1.device)
    _lib.check(_lib.lib().devo_corr_forward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), ii.data_ptr(),
                                            jj.data_ptr(), out.data_ptr(), _lib.dtype_code(fmap1), B, Np, Nf, C, H, W,
                                            E, P, radius, _lib.stream_ptr(fmap1.device)), "corr_forward")
    return [out]


def backward(fmap1, fmap2, coords, ii, jj, corr_grad, radius):
    """cuda_corr.backward -> [fmap1_grad, fmap2_grad]"""
    _check_common(fmap1, fmap2, coords, ii, jj)
    _lib.require_cuda(corr_grad)
    fmap1 = fmap1.contiguous()
    fmap2 = fmap2.contiguous()
    coords = coords.contiguous()
    grad = corr_grad.to(torch.float32).contiguous()
    B, Np, C, P, _ = fmap1.shape
    _, Nf, _, H, W = fmap2.shape
    E = coords.shape[1]
    if E > 0 and _split_eligible(fmap1, fmap2, coords, radius) and fmap2.dtype == torch.float32:
        # training shape: pixel-major volumes, vector reductions (csrc/corr_bwd_pm.cu); the layout changes are plumbing
        f2pm = _cached(fmap2, lambda t: t[0].permute(0, 2, 3, 1).contiguous(), tag="pm32")
        g1pm = torch.zeros(Np, P * P, C, dtype=torch.float32, device=fmap1.device)
        g2pm = torch.zeros(Nf, H, W, C, dtype=torch.float32, device=fmap1.device)
        _lib.check(_lib.lib().devo_corr_backward_pm(fmap1.data_ptr(), f2pm.data_ptr(), coords.data_ptr(),
                                                    ii.contiguous().data_ptr(), jj.contiguous().data_ptr(), grad.data_ptr(),
                                                    g1pm.data_ptr(), g2pm.data_ptr(), Np, Nf, C, H, W, E,
                                                    _lib.stream_ptr(fmap1.device)), "corr_backward_pm")
        return [g1pm.permute(0, 2, 1).reshape(1, Np, C, P, P).contiguous(), g2pm.permute(0, 3, 1, 2).contiguous()[None]]
    g1 = torch.empty_like(fmap1)
    g2 = torch.empty_like(fmap2)
    _lib.check(_lib.lib().devo_corr_backward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                             ii.contiguous().data_ptr(), jj.contiguous().data_ptr(), grad.data_ptr(),
                                             g1.data_ptr(), g2.data_ptr(), _lib.dtype_code(fmap1), B, Np, Nf, C, H, W,
                                             E, P, radius, _lib.stream_ptr(fmap1.device)), "corr_backward")
    return [g1, g2]


def patchify_forward(net, coords, radius):
    """cuda_corr.patchify_forward -> [patches [B,M,C,2r+2,2r+2]]"""
    _lib.require_cuda(net, coords)
    if net.dim() != 4 or coords.dim() != 3 or coords.shape[-1] != 2:
        raise RuntimeError("cuda_corr.patchify_forward: net [B,C,H,W], coords [B,M,2] expected")
    _lib.require_dtype(coords, torch.float32, "coords")
    net = net.contiguous()
    coords = coords.contiguous()
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    out = torch.empty(B, M, C, D, D, dtype=net.dtype, device=net.device)
    _lib.check(_lib.lib().devo_patchify_forward(net.data_ptr(), coords.data_ptr(), out.data_ptr(), _lib.dtype_code(net),
                                                B, C, H, W, M, radius, _lib.stream_ptr(net.device)), "patchify_forward")
    return [out]


def patchify_backward(net, coords, gradient, radius):
    """cuda_corr.patchify_backward -> [net_grad [B,C,H,W]]"""
    _lib.require_cuda(net, coords, gradient)
    _lib.require_dtype(coords, torch.float32, "coords")
    coords = coords.contiguous()
    gradient = gradient.to(net.dtype).contiguous()
    B, C, H, W = net.shape
    M = coords.shape[1]
    out = torch.empty(B, C, H, W, dtype=net.dtype, device=net.device)
    _lib.check(_lib.lib().devo_patchify_backward(gradient.data_ptr(), coords.data_ptr(), out.data_ptr(),
                                                 _lib.dtype_code(net), B, C, H, W, M, radius,
                                                 _lib.stream_ptr(net.device)), "patchify_backward")
    return [out]
