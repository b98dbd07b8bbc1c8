"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's
training-time bundle adjustment devo/ba.py:86-182 (one Gauss-Newton step,
functional), with its helpers CholeskySolver :12-37, safe_scatter_add :40-46,
retractions :49-56, block_matmul/solve :58-76.  This is also the *CPU path*
BASELINE.json's config 1 times (`cpu_baseline`, `--impl reference`).

Differences to fastba that are kept: validity Z>0.2 & |r|<250 &
bounds; projection uses 1/clamp(Z,0.1); damping (ep + 1e-4*S) on the diagonal;
first `fixedp` poses fixed; depth clamp [1e-3,10]; Cholesky failure => zero
update; retraction through lietorch (renormalising).

Pinned against the reference's own devo/ba.py imported from /root/reference
(tests/golden/make_golden.py -> tests/golden/ba_*.pt).
"""
import torch

from . import lie, pops


def _scatter_mat(A, ii, jj, n, m):
    """safe_scatter_add_mat :40-42 ; A [1,E,p,q] -> [1,n*m,p,q]"""
    v = (ii >= 0) & (jj >= 0) & (ii < n) & (jj < m)
    out = torch.zeros((1, n * m) + tuple(A.shape[2:]), dtype=A.dtype)
    if v.any():
        out.index_add_(1, ii[v] * m + jj[v], A[:, v])
    return out


def _scatter_vec(b, ii, n):
    """safe_scatter_add_vec :44-46"""
    v = (ii >= 0) & (ii < n)
    out = torch.zeros((1, n) + tuple(b.shape[2:]), dtype=b.dtype)
    if v.any():
        out.index_add_(1, ii[v], b[:, v])
    return out


def _flat(A):
    """block matrix [1,n,m,p,q] -> [1,n*p,m*q]  (block_matmul :58-64)"""
    b, n, m, p, q = A.shape
    return A.permute(0, 1, 3, 2, 4).reshape(b, n * p, m * q)


def ba_step(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds,
            ep=100.0, fixedp=1, structure_only=False):
    """poses: raw SE3 data [1,N,7]; returns (poses_new [1,N,7], patches_new)"""
    n = int(max(ii.max().item(), jj.max().item())) + 1
    coords, v, (Ji, Jj, Jz) = pops.transform(poses, patches, intrinsics, ii, jj, kk, jacobian=True)
    p = coords.shape[3]
    r = targets - coords[..., p // 2, p // 2, :]
    v = v * (r.norm(dim=-1) < 250).to(v.dtype)
    cx_, cy_ = coords[..., p // 2, p // 2, 0], coords[..., p // 2, p // 2, 1]
    inb = (cx_ > bounds[0]) & (cy_ > bounds[1]) & (cx_ < bounds[2]) & (cy_ < bounds[3])
    v = v * inb.to(v.dtype)
    r = (v[..., None] * r).unsqueeze(-1)
    w = (v[..., None] * weights).unsqueeze(-1)
    wJiT = (w * Ji).transpose(2, 3)
    wJjT = (w * Jj).transpose(2, 3)
    wJzT = (w * Jz).transpose(2, 3)
    mm = torch.matmul
    Bii, Bij, Bji, Bjj = mm(wJiT, Ji), mm(wJiT, Jj), mm(wJjT, Ji), mm(wJjT, Jj)
    Eik, Ejk = mm(wJiT, Jz), mm(wJjT, Jz)
    vi, vj = mm(wJiT, r), mm(wJjT, r)
    n = n - fixedp
    ii_ = ii - fixedp
    jj_ = jj - fixedp
    kx, kq = torch.unique(kk, return_inverse=True, sorted=True)
    m = len(kx)
    B = (_scatter_mat(Bii, ii_, ii_, n, n) + _scatter_mat(Bij, ii_, jj_, n, n)
         + _scatter_mat(Bji, jj_, ii_, n, n) + _scatter_mat(Bjj, jj_, jj_, n, n)).view(1, n, n, 6, 6)
    E = (_scatter_mat(Eik, ii_, kq, n, m) + _scatter_mat(Ejk, jj_, kq, n, m)).view(1, n, m, 6, 1)
    C = _scatter_vec(mm(wJzT, Jz), kq, m)
    vv = (_scatter_vec(vi, ii_, n) + _scatter_vec(vj, jj_, n)).view(1, n, 1, 6, 1)
    ww = _scatter_vec(mm(wJzT, r), kq, m)
    if isinstance(lmbda, torch.Tensor):
        lmbda = lmbda.reshape(*C.shape)
    Q = 1.0 / (C + lmbda)
    EQ = E * Q[:, None]
    dX = None
    if structure_only or n == 0:
        dZ = (Q * ww).view(1, -1, 1, 1)
    else:
        Ef, EQf = _flat(E), _flat(EQ)                      # [1,6n,m]
        S = _flat(B) - mm(EQf, Ef.transpose(1, 2))
        y = _flat(vv) - mm(EQf, ww.view(1, m, 1))
        S = S + (ep + 1e-4 * S) * torch.eye(6 * n, dtype=S.dtype)
        U, info = torch.linalg.cholesky_ex(S)
        if torch.any(info):
            dXf = torch.zeros_like(y)
        else:
            dXf = torch.cholesky_solve(y, U)
        dZ = Q * (ww - mm(Ef.transpose(1, 2), dXf).view(1, m, 1, 1))
        dX = dXf.view(1, -1, 6)
        dZ = dZ.view(1, -1, 1, 1)
    x, y_, disps = patches.unbind(dim=2)
    upd = torch.zeros_like(disps)
    upd.index_add_(1, kx, dZ.expand(-1, -1, disps.shape[2], disps.shape[3]).contiguous())
    disps = (disps + upd).clamp(min=1e-3, max=10.0)
    patches = torch.stack([x, y_, disps], dim=2)
    if dX is not None:
        full = torch.zeros(1, poses.shape[1], 6, dtype=poses.dtype)
        full[:, fixedp:fixedp + n] = dX
        dG = lie.expm(3, full.view(-1, 6))
        poses = lie.mul(3, dG, poses.view(-1, 7)).view(1, -1, 7)
    return poses, patches
