"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's `cuda_corr`
extension -- sparse patch correlation lookup and patch gather.

Follows devo/altcorr/correlation_kernel.cu:
    corr_forward_kernel            :82-136   (window dot products, OOB => 0)
    corr_cuda_forward  (host part) :193-233  (bilinear blend :221-230, permute :232)
    corr_backward                   :139-190, 236-286
    patchify_forward_kernel        :16-47, 288-307
    patchify_backward_kernel       :49-80, 309-333
and devo/altcorr/correlation.py:51-72 (python `patchify(mode='bilinear')`).

The reference has no test or fixture for these ops; they are pinned on the GPU
box against the reference's compiled extension (oracle/_ref/cuda_corr_ref.so).
Arithmetic is done in `compute_dtype` (default fp64): for half inputs the
reference accumulates in half (kernel :121-131), which is *less* accurate than
this oracle, so half-precision parity is judged as "error vs this fp64 oracle
no larger than the reference's".
"""
import torch


def _window_index(coords, radius, H, W):
    """coords [B,E,2,P,P] -> (i1, j1, valid) each [B,E,P,P,D,D] (a = row offset, b = col offset)"""
    D = 2 * radius + 2
    x = coords[:, :, 0]
    y = coords[:, :, 1]
    fx = torch.floor(x).to(torch.int64)
    fy = torch.floor(y).to(torch.int64)
    off = torch.arange(D, dtype=torch.int64) - radius
    i1 = fy[..., None, None] + off.view(D, 1)
    j1 = fx[..., None, None] + off.view(1, D)
    i1, j1 = torch.broadcast_tensors(i1, j1)
    valid = (i1 >= 0) & (i1 < H) & (j1 >= 0) & (j1 < W)
    return i1, j1, valid


def corr_volume(fmap1, fmap2, coords, ii, jj, radius, compute_dtype=torch.float64, chunk=128):
    """V[b,e,a,b',i0,j0]  (correlation_kernel.cu:82-136)"""
    B, E, _, P, _ = coords.shape
    C = fmap1.shape[2]
    H, W = fmap2.shape[3], fmap2.shape[4]
    D = 2 * radius + 2
    f1 = fmap1.to(compute_dtype)
    f2 = fmap2.to(compute_dtype).permute(0, 1, 3, 4, 2)          # [B,Nf,H,W,C]
    i1, j1, valid = _window_index(coords, radius, H, W)
    outs = []
    for s in range(0, E, chunk):
        sl = slice(s, min(E, s + chunk))
        n = sl.stop - sl.start
        ic = i1[:, sl].clamp(0, H - 1)
        jc = j1[:, sl].clamp(0, W - 1)
        jidx = jj[sl].view(1, n, 1, 1, 1, 1).expand(B, n, P, P, D, D)
        bidx = torch.arange(B).view(B, 1, 1, 1, 1, 1).expand_as(jidx)
        win = f2[bidx, jidx, ic, jc]                               # [B,n,P,P,D,D,C]
        win = win * valid[:, sl].unsqueeze(-1).to(compute_dtype)
        p = f1[:, ii[sl]].permute(0, 1, 3, 4, 2)                   # [B,n,P,P,C]
        v = (win * p[:, :, :, :, None, None, :]).sum(-1)           # [B,n,P,P,D,D]
        outs.append(v.permute(0, 1, 4, 5, 2, 3))                   # [B,n,D(a),D(b'),P,P]
    return torch.cat(outs, dim=1) if outs else torch.zeros(B, 0, D, D, P, P, dtype=compute_dtype)


def _blend(V, coords, D):
    """host part of corr_cuda_forward :221-230"""
    x = coords[:, :, 0, None, None].to(V.dtype)
    y = coords[:, :, 1, None, None].to(V.dtype)
    dx = x - torch.floor(x)
    dy = y - torch.floor(y)
    out = (1 - dx) * (1 - dy) * V[:, :, 0:D - 1, 0:D - 1]
    out = out + dx * (1 - dy) * V[:, :, 0:D - 1, 1:D]
    out = out + (1 - dx) * dy * V[:, :, 1:D, 0:D - 1]
    out = out + dx * dy * V[:, :, 1:D, 1:D]
    return out


def corr_forward(fmap1, fmap2, coords, ii, jj, radius, compute_dtype=torch.float64):
    """== cuda_corr.forward(...)[0]; returns [B,E,2r+1(x-off),2r+1(y-off),P,P] in compute_dtype"""
    D = 2 * radius + 2
    V = corr_volume(fmap1, fmap2, coords, ii, jj, radius, compute_dtype)
    return _blend(V, coords, D).permute(0, 1, 3, 2, 4, 5)


def corr_backward(fmap1, fmap2, coords, ii, jj, grad, radius, compute_dtype=torch.float64):
    """== cuda_corr.backward(...) -> (fmap1_grad, fmap2_grad); the transposed blend and the
    scatter of :139-190,252-269 are exactly the adjoint of corr_forward, obtained by autograd."""
    f1 = fmap1.detach().to(compute_dtype).requires_grad_(True)
    f2 = fmap2.detach().to(compute_dtype).requires_grad_(True)
    out = corr_forward(f1, f2, coords, ii, jj, radius, compute_dtype)
    g1, g2 = torch.autograd.grad(out, [f1, f2], grad.to(compute_dtype), allow_unused=True)
    if g1 is None:
        g1 = torch.zeros_like(f1)
    if g2 is None:
        g2 = torch.zeros_like(f2)
    return g1, g2


def patchify_forward(net, coords, radius):
    """== cuda_corr.patchify_forward(...)[0]  (:16-47); exact copy => same dtype, bit-exact"""
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    x = coords[:, :, 0]
    y = coords[:, :, 1]
    fx = torch.floor(x).to(torch.int64)
    fy = torch.floor(y).to(torch.int64)
    off = torch.arange(D, dtype=torch.int64) - radius
    i = fy[:, :, None, None] + off.view(D, 1)            # [B,M,D,1]
    j = fx[:, :, None, None] + off.view(1, D)
    i, j = torch.broadcast_tensors(i, j)
    valid = (i >= 0) & (i < H) & (j >= 0) & (j < W)
    ic, jc = i.clamp(0, H - 1), j.clamp(0, W - 1)
    bidx = torch.arange(B).view(B, 1, 1, 1).expand_as(ic)
    g = net.permute(0, 2, 3, 1)[bidx, ic, jc]            # [B,M,D,D,C]
    g = torch.where(valid.unsqueeze(-1), g, torch.zeros((), dtype=net.dtype))
    return g.permute(0, 1, 4, 2, 3).contiguous()        # [B,M,C,D(a: row),D(b: col)]


def patchify_backward(net, coords, gradient, radius):
    """== cuda_corr.patchify_backward(...)[0]  (:49-80): scatter-add"""
    B, C, H, W = net.shape
    D = 2 * radius + 2
    x = coords[:, :, 0]
    y = coords[:, :, 1]
    fx = torch.floor(x).to(torch.int64)
    fy = torch.floor(y).to(torch.int64)
    off = torch.arange(D, dtype=torch.int64) - radius
    i = fy[:, :, None, None] + off.view(D, 1)
    j = fx[:, :, None, None] + off.view(1, D)
    i, j = torch.broadcast_tensors(i, j)
    valid = (i >= 0) & (i < H) & (j >= 0) & (j < W)
    lin = (i.clamp(0, H - 1) * W + j.clamp(0, W - 1))    # [B,M,D,D]
    out = torch.zeros(B, C, H * W, dtype=gradient.dtype)
    g = gradient * valid.unsqueeze(2).to(gradient.dtype)  # [B,M,C,D,D]
    lin = lin.unsqueeze(2).expand_as(g)
    out.scatter_add_(2, lin.permute(0, 2, 1, 3, 4).reshape(B, C, -1), g.permute(0, 2, 1, 3, 4).reshape(B, C, -1))
    return out.view(B, C, H, W)


def patchify(net, coords, radius, mode="bilinear"):
    """devo/altcorr/correlation.py:51-68"""
    patches = patchify_forward(net, coords, radius)
    if mode == "bilinear":
        offset = coords - coords.floor()
        dx, dy = offset[:, :, None, None, None].unbind(dim=-1)
        d = 2 * radius + 1
        x00 = (1 - dy) * (1 - dx) * patches[..., :d, :d]
        x01 = (1 - dy) * dx * patches[..., :d, 1:]
        x10 = dy * (1 - dx) * patches[..., 1:, :d]
        x11 = dy * dx * patches[..., 1:, 1:]
        return x00 + x01 + x10 + x11
    return patches
