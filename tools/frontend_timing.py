"""Tail of the reference's Patchifier.forward (3 x altcorr.patchify + coordinate grid + gmap re-pack) against the fused patch
gather, one frame of 96 patches at 160x120:  python tools/frontend_timing.py"""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import altcorr, cuda_corr
from devo_b200.frontend import gather_patches
from bench import time_us
dev = "cuda"
fmap = torch.randn(1, 128, 120, 160, device=dev).half(); imap = torch.randn(1, 384, 120, 160, device=dev).half()
coords = torch.stack([torch.randint(1, 159, (1, 96), device=dev), torch.randint(1, 119, (1, 96), device=dev)], -1).float()
yy, xx = torch.meshgrid(torch.arange(120, device=dev).float(), torch.arange(160, device=dev).float(), indexing="ij")
def composed():
    im = altcorr.patchify(imap, coords, 0); gm = altcorr.patchify(fmap, coords, 1)
    grid = torch.stack([xx[None], yy[None], torch.ones(1, 120, 160, device=dev)], 1)
    pt = altcorr.patchify(grid, coords, 1)
    return cuda_corr.pack_gmap(gm[0].half())
s = torch.cuda.current_stream()
with torch.no_grad():
    print("reference-style tail (3 x altcorr.patchify + grid + gmap pack), eager: %.1f us" % time_us(composed, s, None, warm=10, n=100))
    print("devo_patch_gather, eager: %.1f us" % time_us(lambda: gather_patches(fmap, imap, coords, None, 3, planar=False, pixel_major=True), s, None, warm=10, n=100))
