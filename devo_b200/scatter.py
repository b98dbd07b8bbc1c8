"""Segment reductions the hot path needs from `torch_scatter` (not installed here):
scatter_sum(src, index, dim, dim_size) and scatter_softmax(src, index, dim)
(devo/ba.py:42-56, devo/blocks.py:42-43).  Differentiable torch compositions; the
inference engine uses the fused CUDA kernel cuda_ba.segment_softmax_sum instead."""
import torch


def _expand(src, index, dim):
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src)


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    idx = _expand(src, index, dim)
    if out is not None:
        return out.scatter_add_(dim, idx, src)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, src)


def scatter_softmax(src, index, dim=-1, dim_size=None):
    dim = dim % src.dim()
    idx = _expand(src, index, dim)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    mx = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(dim, idx, src.detach(), reduce="amax", include_self=True)
    ex = torch.exp(src - mx.gather(dim, idx))
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, ex)
    return ex / den.gather(dim, idx)
