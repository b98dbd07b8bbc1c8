"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's in-place
Gauss-Newton bundle adjustment `cuda_ba.forward` and of `cuda_ba.reproject`.

Follows devo/fastba/ba_cuda.cu:
    SE3 helpers actSO3/actSE3/adjSE3/relSE3/expSO3/expSE3/retrSE3   :18-156
    reprojection_residuals_and_hessian                               :214-365
    Schur complement / Cholesky / back-substitution (host, ATen)     :461-537
    pose_retr_kernel :160-188, patch_retr_kernel :191-211
    reproject :368-418, 543-575
Quirks kept: un-normalised quaternions (:56-67,138-156); small-angle thresholds
theta^2<1e-8 / theta>1e-4 (:79,122); only intrinsics[0] is used (:232-238);
x1 = fx*X/Z unguarded while d = 1/Z only for Z>=0.2 (:265-269); poses < t0 are
fixed (:282-283); depth reset d>20 -> 1, floor 1e-4 (:196-209).

The reference has no test for this op; pinned on the GPU box against
oracle/_ref/cuda_ba_ref.so.  Arithmetic in `dtype` (default fp64).
Returns new tensors (the reference mutates in place) plus a status flag.
"""
import numpy as np
import torch


def _act_so3(q, X):
    qv = q[:, :3]
    uv = 2.0 * torch.linalg.cross(qv, X)
    return X + q[:, 3:4] * uv + torch.linalg.cross(qv, uv)


def _rel_se3(ti, qi, tj, qj):
    """relSE3 :56-67  (qij = qj * conj(qi), no renormalisation)"""
    qij = torch.stack([
        -qj[:, 3] * qi[:, 0] + qj[:, 0] * qi[:, 3] - qj[:, 1] * qi[:, 2] + qj[:, 2] * qi[:, 1],
        -qj[:, 3] * qi[:, 1] + qj[:, 1] * qi[:, 3] - qj[:, 2] * qi[:, 0] + qj[:, 0] * qi[:, 2],
        -qj[:, 3] * qi[:, 2] + qj[:, 2] * qi[:, 3] - qj[:, 0] * qi[:, 1] + qj[:, 1] * qi[:, 0],
        qj[:, 3] * qi[:, 3] + qj[:, 0] * qi[:, 0] + qj[:, 1] * qi[:, 1] + qj[:, 2] * qi[:, 2]], dim=-1)
    tij = tj - _act_so3(qij, ti)
    return tij, qij


def _adj_se3(t, q, X):
    """adjSE3 :39-54 : Y = Ad(G)^T X"""
    qinv = torch.cat([-q[:, :3], q[:, 3:]], dim=-1)
    Y0 = _act_so3(qinv, X[:, :3])
    Y1 = _act_so3(qinv, X[:, 3:])
    u = torch.stack([
        t[:, 2] * X[:, 1] - t[:, 1] * X[:, 2],
        t[:, 0] * X[:, 2] - t[:, 2] * X[:, 0],
        t[:, 1] * X[:, 0] - t[:, 0] * X[:, 1]], dim=-1)
    return torch.cat([Y0, Y1 + _act_so3(qinv, u)], dim=-1)


def _exp_se3(xi):
    """expSO3 :71-92, expSE3 :107-135"""
    tau, phi = xi[:, :3], xi[:, 3:]
    th2 = (phi * phi).sum(-1)
    th4 = th2 * th2
    th = torch.sqrt(th2)
    small = th2 < 1e-8
    ts = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4, torch.sin(0.5 * ts) / ts)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4, torch.cos(0.5 * ts))
    q = torch.cat([imag[:, None] * phi, real[:, None]], dim=-1)
    big = th > 1e-4
    tb = torch.where(big, th, torch.ones_like(th))
    t2b = torch.where(big, th2, torch.ones_like(th2))
    a = (1 - torch.cos(tb)) / t2b
    b = (tb - torch.sin(tb)) / (tb * t2b)
    c1 = torch.linalg.cross(phi, tau)
    c2 = torch.linalg.cross(phi, c1)
    t = tau + torch.where(big[:, None], a[:, None] * c1 + b[:, None] * c2, torch.zeros_like(tau))
    return t, q


def _retr_se3(xi, t, q):
    """retrSE3 :138-156"""
    dt, dq = _exp_se3(xi)
    q1 = torch.stack([
        dq[:, 3] * q[:, 0] + dq[:, 0] * q[:, 3] + dq[:, 1] * q[:, 2] - dq[:, 2] * q[:, 1],
        dq[:, 3] * q[:, 1] + dq[:, 1] * q[:, 3] + dq[:, 2] * q[:, 0] - dq[:, 0] * q[:, 2],
        dq[:, 3] * q[:, 2] + dq[:, 2] * q[:, 3] + dq[:, 0] * q[:, 1] - dq[:, 1] * q[:, 0],
        dq[:, 3] * q[:, 3] - dq[:, 0] * q[:, 0] - dq[:, 1] * q[:, 1] - dq[:, 2] * q[:, 2]], dim=-1)
    t1 = _act_so3(dq, t) + dt
    return t1, q1


def edge_terms(poses, patches, intrinsics, target, weight, ii, jj, kk):
    """per-edge quantities of reprojection_residuals_and_hessian :240-332.
    returns dict of [E,...] tensors: r[E,2], w[E,2], Ji[E,2,6], Jj[E,2,6], Jz[E,2]"""
    fx, fy, cx, cy = [intrinsics[0, k] for k in range(4)]
    Pp = patches.shape[-1]
    c = Pp // 2 if Pp != 3 else 1                     # reference hard-codes [1][1]
    ti, qi = poses[ii, :3], poses[ii, 3:]
    tj, qj = poses[jj, :3], poses[jj, 3:]
    px, py, pd = patches[kk, 0, 1, 1], patches[kk, 1, 1, 1], patches[kk, 2, 1, 1]
    Xi = torch.stack([(px - cx) / fx, (py - cy) / fy, torch.ones_like(px)], dim=-1)
    tij, qij = _rel_se3(ti, qi, tj, qj)
    Xj = _act_so3(qij, Xi) + pd[:, None] * tij
    X, Y, Z, W = Xj[:, 0], Xj[:, 1], Xj[:, 2], pd
    d = torch.where(Z >= 0.2, 1.0 / Z, torch.zeros_like(Z))
    d2 = d * d
    x1 = fx * (X / Z) + cx
    y1 = fy * (Y / Z) + cy
    rx = target[:, 0] - x1
    ry = target[:, 1] - y1
    inb = (torch.sqrt(rx * rx + ry * ry) < 128) & (Z > 0.2) & (x1 > -64) & (y1 > -64) \
        & (x1 < 2 * cx + 64) & (y1 < 2 * cy + 64)
    mask = inb.to(Z.dtype)
    o = torch.zeros_like(Z)
    Jjx = torch.stack([fx * W * d, o, fx * -X * W * d2, fx * -X * Y * d2, fx * (1 + X * X * d2), fx * -Y * d], dim=-1)
    Jjy = torch.stack([o, fy * W * d, fy * -Y * W * d2, fy * (-1 - Y * Y * d2), fy * (X * Y * d2), fy * X * d], dim=-1)
    Jix = _adj_se3(tij, qij, Jjx)
    Jiy = _adj_se3(tij, qij, Jjy)
    Jzx = fx * (tij[:, 0] * d - tij[:, 2] * (X * d2))
    Jzy = fy * (tij[:, 1] * d - tij[:, 2] * (Y * d2))
    return dict(r=torch.stack([rx, ry], -1), w=mask[:, None] * weight,
                Ji=torch.stack([Jix, Jiy], 1), Jj=torch.stack([Jjx, Jjy], 1),
                Jz=torch.stack([Jzx, Jzy], -1), coords=torch.stack([x1, y1], -1), Z=Z)


def ba(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations,
       dtype=torch.float64, return_system=False):
    """== cuda_ba.forward; returns (poses_new [N,7], patches_new [Np,3,P,P], status)
    status = 0, or (iteration+1) at which the Cholesky factorisation failed
    (the reference raises there; updates of earlier iterations persist)."""
    poses = poses.reshape(-1, 7).to(dtype).clone()
    Pp = patches.shape[-1]
    patches = patches.reshape(-1, 3, Pp, Pp).to(dtype).clone()
    intr = intrinsics.reshape(-1, 4).to(dtype)
    target = target.reshape(-1, 2).to(dtype)
    weight = weight.reshape(-1, 2).to(dtype)
    lm = lmbda.reshape(-1).to(dtype)
    kx, ku = torch.unique(kk, sorted=True, return_inverse=True)
    N = t1 - t0
    M = kx.shape[0]
    status = 0
    system = None
    for itr in range(iterations):
        T = edge_terms(poses, patches, intr, target, weight, ii, jj, kk)
        ix = ii - t0
        jx = jj - t0
        B = torch.zeros(max(N, 0) * 6, max(N, 0) * 6, dtype=dtype)
        Em = torch.zeros(max(N, 0) * 6, M, dtype=dtype)
        C = torch.zeros(M, dtype=dtype)
        v = torch.zeros(max(N, 0) * 6, dtype=dtype)
        u = torch.zeros(M, dtype=dtype)
        E_ = ii.shape[0]
        # dense per-row sparse vector g (length 6N): -Ji at block i', +Jj at block j'
        G = torch.zeros(E_, 2, max(N, 0) * 6, dtype=dtype)
        ar6 = torch.arange(6)
        for blk, J, sgn in ((ix, T["Ji"], -1.0), (jx, T["Jj"], 1.0)):
            ok = (blk >= 0) & (blk < N) if N > 0 else torch.zeros_like(blk, dtype=torch.bool)
            e_ok = torch.nonzero(ok).squeeze(-1)
            if e_ok.numel():
                cols = (blk[e_ok] * 6)[:, None] + ar6[None, :]
                for rho in range(2):
                    G[e_ok[:, None], rho, cols] += sgn * J[e_ok, rho]
        w, r, Jz = T["w"], T["r"], T["Jz"]
        if N > 0:
            Gf = G.reshape(-1, 6 * N)
            wf = w.reshape(-1)
            B = (Gf * wf[:, None]).t() @ Gf
            v = (Gf * (wf * r.reshape(-1))[:, None]).sum(0)
            Em.index_add_(1, ku.repeat_interleave(2), (Gf * (wf * Jz.reshape(-1))[:, None]).t())
        C.index_add_(0, ku, (w * Jz * Jz).sum(-1))
        u.index_add_(0, ku, (w * r * Jz).sum(-1))
        Q = 1.0 / (C + lm)
        if N <= 0:
            dZ = Q * u
            dX = None
        else:
            EQ = Em * Q[None, :]
            S = B - EQ @ Em.t()
            y = v - EQ @ u
            S = S + torch.eye(6 * N, dtype=dtype) * (1e-4 * S + 1.0)
            if return_system and itr == 0:
                system = dict(B=B.clone(), E=Em.clone(), C=C.clone(), v=v.clone(), u=u.clone(), S=S.clone(), y=y.clone())
            U, info = torch.linalg.cholesky_ex(S)
            if int(info) != 0 or not torch.isfinite(S).all():
                status = itr + 1
                break
            dX = torch.cholesky_solve(y[:, None], U)[:, 0]
            dZ = Q * (u - Em.t() @ dX)
            tn, qn = _retr_se3(dX.view(N, 6), poses[t0:t1, :3], poses[t0:t1, 3:])
            poses[t0:t1, :3] = tn
            poses[t0:t1, 3:] = qn
        dcur = patches[kx, 2, 0, 0] + dZ
        dcur = torch.where(dcur > 20, torch.ones_like(dcur), dcur)
        dcur = torch.maximum(dcur, torch.full_like(dcur, 1e-4))
        patches[kx, 2] = dcur[:, None, None].expand(-1, Pp, Pp)
    if return_system:
        return poses, patches, status, system
    return poses, patches, status


def reproject(poses, patches, intrinsics, ii, jj, kk, dtype=torch.float64):
    """== cuda_ba.reproject -> [1,E,2,P,P]  (:368-418; no clamp on Z)"""
    poses = poses.reshape(-1, 7).to(dtype)
    Pp = patches.shape[-1]
    patches = patches.reshape(-1, 3, Pp, Pp).to(dtype)
    intr = intrinsics.reshape(-1, 4).to(dtype)
    fx, fy, cx, cy = [intr[0, k] for k in range(4)]
    tij, qij = _rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    E = ii.shape[0]
    px = patches[kk, 0].reshape(E, -1)
    py = patches[kk, 1].reshape(E, -1)
    pd = patches[kk, 2].reshape(E, -1)
    Xi = torch.stack([(px - cx) / fx, (py - cy) / fy, torch.ones_like(px)], dim=-1)       # [E,PP,3]
    q9 = qij[:, None, :].expand(E, Pp * Pp, 4).reshape(-1, 4)
    Xj = _act_so3(q9, Xi.reshape(-1, 3)).view(E, Pp * Pp, 3) + pd[..., None] * tij[:, None, :]
    x1 = fx * (Xj[..., 0] / Xj[..., 2]) + cx
    y1 = fy * (Xj[..., 1] / Xj[..., 2]) + cy
    return torch.stack([x1, y1], dim=1).view(1, E, 2, Pp, Pp)
