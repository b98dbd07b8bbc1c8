"""Differentiable bundle adjustment of the training graph: `BA(...)` has the signature and
semantics of devo/ba.py:86-182 (one damped Gauss-Newton step on the poses after the first
`fixedp` and on the inverse depth of every patch in unique(kk); functional: returns new
`(poses, patches)`; autograd flows through everything).

Formulation (the same one the CUDA fastba kernels use, see csrc/ba.cu): every residual row
is the sparse vector  g = [Ji at block i', Jj at block j']  (projective_ops.transform already returns
Ji = -Ad(Gij)^T Jj, sign included)  so that

    B = G^T W G,   v = G^T W r,   E_k = sum_{rows of patch k} w Jz g,   C_k, u_k likewise
    S = B - E^T Q E,  y = v - E^T Q u,  Q = 1/(C + lambda)

which turns the reference's nine per-edge 6x6 scatter_sums into two dense GEMMs and three
segment sums.  Constants as in the reference: validity Z>0.2 & |r|<250 & bounds (:98-106),
damping (ep + 1e-4 S) on the diagonal (:71), depth clamp [1e-3,10] (:176), a failed
Cholesky gives a zero update (:16-20)."""
import torch
import torch.nn.functional as F

from . import projective_ops as pops
from .scatter import scatter_sum


class CholeskySolver(torch.autograd.Function):
    """x = H^-1 b via Cholesky; zero (and zero gradient) when H is not positive definite (devo/ba.py:12-37).
    No host synchronisation: the factorisation's `info` stays on the device and selects the result."""

    @staticmethod
    def forward(ctx, H, b):
        U, info = torch.linalg.cholesky_ex(H)
        ok = (info == 0).view(-1, 1, 1)
        xs = torch.where(ok, torch.cholesky_solve(b, torch.where(ok, U, torch.eye(U.shape[-1], dtype=U.dtype, device=U.device))),
                         torch.zeros_like(b))
        ctx.save_for_backward(U, xs, ok)
        return xs

    @staticmethod
    def backward(ctx, grad_x):
        U, xs, ok = ctx.saved_tensors
        Us = torch.where(ok, U, torch.eye(U.shape[-1], dtype=U.dtype, device=U.device))
        dz = torch.where(ok, torch.cholesky_solve(grad_x, Us), torch.zeros_like(grad_x))
        return -torch.matmul(xs, dz.transpose(-1, -2)), dz


def _block_rows(J, blk, nfree, sign):
    """place the [1,E,2,6] Jacobian rows at pose block `blk` of a [1,E,2,nfree,6] row matrix"""
    ok = (blk >= 0) & (blk < nfree)
    sel = F.one_hot(blk.clamp(0, max(nfree - 1, 0)), max(nfree, 1)).to(J.dtype) * ok.to(J.dtype)[:, None]
    return sign * sel[None, :, None, :nfree, None] * J[:, :, :, None, :]


def BA(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, PRINT=False,
       fixedp=1, structure_only=False):
    """One damped Gauss-Newton step, differentiable, WITHOUT host synchronisation on CUDA:
      * the geometry (reprojection + centre-pixel Jacobians) is one fused forward / one fused backward launch
        (projective_ops._TransformFn) instead of ~60 autograd nodes;
      * the reference sizes the system by `max(ii.max().item(), jj.max().item()) + 1` and by `torch.unique(kk)` (two host
        round trips): here every pose of `poses` after the first `fixedp` and every patch of `patches` is an unknown --
        frames / patches without edges contribute empty rows, the damping (ep on the diagonal, lambda in C) makes their update
        exactly zero, and the blocks of the others are unchanged, so the result is the reference's;
      * a failed factorisation selects the zero update on the device (CholeskySolver)."""
    sync_free = poses.data.is_cuda and not (isinstance(lmbda, torch.Tensor) and lmbda.numel() > 1)
    n_frames = poses.data.shape[1] if sync_free else max(ii.max().item(), jj.max().item()) + 1
    coords, ok, (Ji, Jj, Jz) = pops.transform(poses, patches, intrinsics, ii, jj, kk, jacobian=True)
    c = coords.shape[3] // 2
    centre = coords[..., c, c, :]
    res = targets - centre
    ok = ok * (res.norm(dim=-1) < 250).float()
    ok = ok * ((centre[..., 0] > bounds[0]) & (centre[..., 1] > bounds[1]) &
               (centre[..., 0] < bounds[2]) & (centre[..., 1] < bounds[3])).float()
    if PRINT:
        print((res * ok[..., None]).norm(dim=-1).mean().item())

    E = ii.numel()
    nfree = n_frames - fixedp
    n6 = 6 * nfree
    r = (ok[..., None] * res).reshape(1, 2 * E, 1)
    w = (ok[..., None] * weights).reshape(1, 2 * E, 1)
    z = Jz.reshape(1, 2 * E, 1)

    if sync_free:
        kx, kq, m = None, kk, patches.shape[1]              # segment id = the patch index itself
    else:
        kx, kq = torch.unique(kk, return_inverse=True, sorted=True)
        m = kx.numel()
    rows_k = kq.repeat_interleave(2)
    C = scatter_sum(w * z * z, rows_k, dim=1, dim_size=m)
    u = scatter_sum(w * r * z, rows_k, dim=1, dim_size=m)
    if isinstance(lmbda, torch.Tensor):
        lmbda = lmbda.reshape(1, -1, 1) if lmbda.numel() == m else lmbda.reshape(())
    Q = 1.0 / (C + lmbda)

    dX = None
    if structure_only or nfree == 0:
        dZ = Q * u
    else:
        G = (_block_rows(Ji, ii - fixedp, nfree, 1.0) + _block_rows(Jj, jj - fixedp, nfree, 1.0)).reshape(1, 2 * E, n6)
        wG = w * G
        B = torch.matmul(G.transpose(1, 2), wG)
        v = torch.matmul(G.transpose(1, 2), w * r)
        Ek = scatter_sum(z * wG, rows_k, dim=1, dim_size=m)                    # [1,m,n6]
        S = B - torch.matmul(Ek.transpose(1, 2), Q * Ek)
        y = v - torch.matmul(Ek.transpose(1, 2), Q * u)
        S = S + (ep + 1e-4 * S) * torch.eye(n6, device=S.device, dtype=S.dtype)
        dX = CholeskySolver.apply(S, y)
        dZ = Q * (u - torch.matmul(Ek, dX))
        dX = dX.view(1, nfree, 6)

    x, y_, disps = patches.unbind(dim=2)
    if sync_free:
        step = dZ.view(1, m, 1, 1).expand(-1, -1, disps.shape[2], disps.shape[3])
    else:
        step = scatter_sum(dZ.view(1, m, 1, 1).expand(-1, -1, disps.shape[2], disps.shape[3]), kx.to(dZ.device),
                           dim=1, dim_size=disps.shape[1])
    patches = torch.stack([x, y_, (disps + step).clamp(min=1e-3, max=10.0)], dim=2)
    if dX is not None:
        full = torch.zeros(1, poses.shape[1], 6, device=dX.device, dtype=dX.dtype)
        full = torch.cat([full[:, :fixedp], dX, full[:, fixedp + nfree:]], dim=1)
        poses = poses.retr(full)
    return poses, patches
