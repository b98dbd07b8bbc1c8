"""The tail of the reference's `Patchifier.forward` (devo/enet.py:123-200) -- everything after the two encoders -- on one
fused kernel (csrc/patch_gather.cu): patch centres -> context rows `imap`, patch features `gmap` (the reference's planar
layout and / or the pixel-major layout the fused lookup reads) and the (x, y, inverse depth) patches.

The encoders themselves (cuDNN convolutions, once per frame) are out of the section-8 scope: `PatchFrontend` takes their
outputs.  Patch selection: RANDOM (enet.py:146-149) or centres supplied by the caller (the reference's gradient / scorer
selectors produce integer centres too and can be passed in as `coords`)."""
import torch

from . import _lib


def gather_patches(fmap, imap, coords, disps=None, P=3, planar=True, pixel_major=False):
    """fmap [N,C,H,W], imap [N,D,H,W] (same float dtype; either may be None), coords [N,M,2] float (x, y) at feature
    resolution, disps [N,H,W] float or None (= ones).
    Returns (gmap [N*M,C,P,P] or None, gmap_pm [N*M,P*P,C] or None, imap [N*M,D] or None, patches [N*M,3,P,P] float32):
    exactly `altcorr.patchify(fmap, coords, P//2)`, `altcorr.patchify(imap, coords, 0)` and
    `altcorr.patchify(coords_grid_with_index(disps), coords, P//2)` of the reference (bilinear mode), rounded to the
    feature dtype."""
    ref = fmap if fmap is not None else imap
    _lib.require_cuda(ref, coords)
    _lib.require_dtype(coords, torch.float32, "coords")
    if coords.dim() != 3 or coords.shape[-1] != 2:
        raise RuntimeError("gather_patches: coords [N,M,2] expected")
    N, M = coords.shape[0], coords.shape[1]
    dev, dt = ref.device, ref.dtype
    H, W = ref.shape[-2], ref.shape[-1]
    C = D = 0
    if fmap is not None:
        fmap = fmap.contiguous()
        if fmap.dim() != 4 or fmap.shape[0] != N:
            raise RuntimeError("gather_patches: fmap [N,C,H,W] expected")
        C = fmap.shape[1]
    if imap is not None:
        imap = imap.contiguous()
        if imap.dim() != 4 or imap.shape[0] != N or imap.dtype != dt or tuple(imap.shape[-2:]) != (H, W):
            raise RuntimeError("gather_patches: imap [N,D,H,W] of the dtype and size of fmap expected")
        D = imap.shape[1]
    if disps is not None:
        _lib.require_dtype(disps, torch.float32, "disps")
        disps = disps.contiguous()
        if tuple(disps.shape) != (N, H, W):
            raise RuntimeError("gather_patches: disps [N,H,W] expected")
    coords = coords.contiguous()
    g = torch.empty(N * M, C, P, P, dtype=dt, device=dev) if (fmap is not None and planar) else None
    gpm = torch.empty(N * M, P * P, C, dtype=dt, device=dev) if (fmap is not None and pixel_major) else None
    if fmap is not None and g is None and gpm is None:
        raise RuntimeError("gather_patches: ask for at least one gmap layout")
    im = torch.empty(N * M, D, dtype=dt, device=dev) if imap is not None else None
    pt = torch.empty(N * M, 3, P, P, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().devo_patch_gather(_lib.ptr(fmap), _lib.ptr(imap), _lib.ptr(disps), coords.data_ptr(), _lib.ptr(g),
                                            _lib.ptr(gpm), _lib.ptr(im), pt.data_ptr(), _lib.dtype_code(ref), N, C, D, H, W, M, P,
                                            _lib.stream_ptr(dev)), "patch_gather")
    return g, gpm, im, pt


class PatchFrontend:
    """`Patchifier.forward` after the encoders, with the reference's return convention:
    (fmap, gmap [1,N*M,C,P,P], imap [1,N*M,D,1,1], patches [1,N*M,3,P,P], index [N*M])."""

    def __init__(self, patch_size=3, pixel_major=False):
        self.P = patch_size
        self.pixel_major = pixel_major

    def __call__(self, fmap, imap, patches_per_image=80, disps=None, coords=None, generator=None):
        """fmap [1,N,C,H,W], imap [1,N,D,H,W] (encoder outputs already divided by 4 like enet.py:125-126)"""
        b, n, c, h, w = fmap.shape
        if b != 1:
            raise RuntimeError("PatchFrontend: batch 1 (as everywhere in DEVO)")
        if coords is None:                                            # SelectionMethod.RANDOM (enet.py:146-149)
            x = torch.randint(1, w - 1, size=[n, patches_per_image], device=fmap.device, generator=generator)
            y = torch.randint(1, h - 1, size=[n, patches_per_image], device=fmap.device, generator=generator)
            coords = torch.stack([x, y], dim=-1).float()
        g, gpm, im, pt = gather_patches(fmap[0], imap[0], coords, None if disps is None else disps[0], self.P,
                                        planar=True, pixel_major=self.pixel_major)
        index = torch.arange(n, device=fmap.device).view(n, 1).repeat(1, coords.shape[1]).reshape(-1)
        out = (fmap, g[None], im.view(1, -1, im.shape[1], 1, 1), pt[None], index)
        return out + (gpm,) if self.pixel_major else out
