"""Oracle (TEST INFRASTRUCTURE): functional stand-in for the two torch_scatter
calls the reference hot path makes (torch_scatter 2.0.9 is not installed):

    scatter_sum(src, index, dim, dim_size)   devo/ba.py:42,46,51,56; devo/blocks.py:43
    scatter_softmax(src, index, dim)          devo/blocks.py:42

Semantics follow the published torch_scatter API: `index` is a 1-D (or
broadcastable) tensor of group ids along `dim`.
"""
import torch


def _expand_index(src, index, dim):
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src)


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    idx = _expand_index(src, index, dim)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add(dim, idx, src)


def scatter_softmax(src, index, dim=-1):
    dim = dim % src.dim()
    idx = _expand_index(src, index, dim)
    n = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = n
    mx = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(dim, idx, src, reduce="amax", include_self=True)
    ex = torch.exp(src - mx.gather(dim, idx))
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, ex)
    return ex / den.gather(dim, idx)
