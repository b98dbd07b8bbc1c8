"""Autograd wrappers of the correlation lookup and patch gather
(same call signatures as devo/altcorr/correlation.py:4-72)."""
import torch

from .. import cuda_corr


class CorrLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, ii, jj, radius, dropout):
        ctx.save_for_backward(fmap1, fmap2, coords, ii, jj)
        ctx.radius = radius
        ctx.dropout = dropout
        (out,) = cuda_corr.forward(fmap1, fmap2, coords, ii, jj, radius)
        return out

    @staticmethod
    def backward(ctx, grad):
        fmap1, fmap2, coords, ii, jj = ctx.saved_tensors
        if ctx.dropout < 1:
            # edge dropout in the backward pass only (reference: correlation.py:20-25)
            keep = torch.rand(len(ii), device=grad.device) < ctx.dropout
            coords, grad, ii, jj = coords[:, keep], grad[:, keep], ii[keep], jj[keep]
        g1, g2 = cuda_corr.backward(fmap1, fmap2, coords, ii, jj, grad, ctx.radius)
        return g1, g2, None, None, None, None, None


class PatchLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, coords, radius):
        ctx.radius = radius
        ctx.save_for_backward(net, coords)
        (patches,) = cuda_corr.patchify_forward(net, coords, radius)
        return patches

    @staticmethod
    def backward(ctx, grad):
        net, coords = ctx.saved_tensors
        (g,) = cuda_corr.patchify_backward(net, coords, grad, ctx.radius)
        return g, None, None


def patchify(net, coords, radius, mode="bilinear"):
    """gather (2r+2)^2 windows at floor(coords); 'bilinear' blends them to (2r+1)^2"""
    patches = PatchLayer.apply(net, coords, radius)
    if mode != "bilinear":
        return patches
    frac = (coords - coords.floor()).to(net.device)
    dx, dy = frac[:, :, None, None, None].unbind(dim=-1)
    d = 2 * radius + 1
    return ((1 - dy) * (1 - dx) * patches[..., :d, :d] + (1 - dy) * dx * patches[..., :d, 1:]
            + dy * (1 - dx) * patches[..., 1:, :d] + dy * dx * patches[..., 1:, 1:])


def corr(fmap1, fmap2, coords, ii, jj, radius=1, dropout=1):
    return CorrLayer.apply(fmap1, fmap2, coords, ii, jj, radius, dropout)
