"""Oracle (TEST INFRASTRUCTURE): CPU restatement of devo/projective_ops.py
(iproj :19-30, proj :32-50, transform :53-105, point_cloud :107-109,
flow_mag :111-121) on raw SE3 data tensors, built on oracle.lie.

Pinned against the reference's own Python module (imported from
/root/reference on top of oracle.lie) by tests/golden/make_golden.py.
Forward values only; shapes as in the reference: poses [1,N,7],
patches [1,Np,3,P,P], intrinsics [1,N,4], ii/jj/kk int64 [E].
"""
import torch

from . import lie

MIN_DEPTH = 0.2


def iproj(patches, intrinsics):
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    i = torch.ones_like(d)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, i, d], dim=-1)


def proj(X, intrinsics, depth=False):
    X_, Y_, Z_, W_ = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z_.clamp(min=0.1)
    x = fx * (d * X_) + cx
    y = fy * (d * Y_) + cy
    if depth:
        return torch.stack([x, y, d], dim=-1)
    return torch.stack([x, y], dim=-1)


def relative_pose(poses, ii, jj, tonly=False):
    """Gij = poses[jj] * poses[ii].inv()   (projective_ops.py:61)"""
    Gi = poses[0, ii]
    Gj = poses[0, jj]
    Gij = lie.mul(3, Gj, lie.inv(3, Gi))
    if tonly:
        Gij = Gij.clone()
        Gij[:, 3:] = torch.as_tensor([0, 0, 0, 1], dtype=Gij.dtype)
    return Gij


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    E = ii.shape[0]
    X0 = iproj(patches[:, kk], intrinsics[:, ii])              # [1,E,P,P,4]
    P = X0.shape[2]
    Gij = relative_pose(poses, ii, jj, tonly)                  # [E,7]
    G9 = Gij[:, None, None, :].expand(E, P, P, 7).reshape(-1, 7)
    X1 = lie.act4(3, G9, X0.reshape(-1, 4)).view(1, E, P, P, 4)
    x1 = proj(X1, intrinsics[:, jj], depth)
    p = P
    if jacobian:
        X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
        o = torch.zeros_like(H)
        fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
        d = torch.zeros_like(Z)
        m = Z.abs() > 0.2
        d[m] = 1.0 / Z[m]
        Ja = torch.stack([
            H, o, o, o, Z, -Y,
            o, H, o, -Z, o, X,
            o, o, H, Y, -X, o,
            o, o, o, o, o, o], dim=-1).view(1, E, 4, 6)
        Jp = torch.stack([
            fx * d, o, -fx * X * d * d, o,
            o, fy * d, -fy * Y * d * d, o], dim=-1).view(1, E, 2, 4)
        Jj = torch.matmul(Jp, Ja)                                # [1,E,2,6]
        G2 = Gij[:, None, :].expand(E, 2, 7).reshape(-1, 7)
        Ji = -lie.adjT(3, G2, Jj.reshape(-1, 6)).view(1, E, 2, 6)
        M = lie.as_matrix(3, Gij).view(1, E, 4, 4)
        Jz = torch.matmul(Jp, M[..., :, 3:])
        return x1, (Z > 0.2).to(x1.dtype), (Ji, Jj, Jz)
    if valid:
        return x1, (X1[..., p // 2, p // 2, 2] > 0.2).to(x1.dtype)
    return x1


def point_cloud(poses, patches, intrinsics, ix):
    X0 = iproj(patches, intrinsics[:, ix])
    n, P = X0.shape[1], X0.shape[2]
    Gi = lie.inv(3, poses[0, ix])
    G9 = Gi[:, None, None, :].expand(n, P, P, 7).reshape(-1, 7)
    return lie.act4(3, G9, X0.reshape(-1, 4)).view(1, n, P, P, 4)


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    c0 = transform(poses, patches, intrinsics, ii, ii, kk)
    c1 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=False)
    c2 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    f1 = (c1 - c0).norm(dim=-1)
    f2 = (c2 - c0).norm(dim=-1)
    return beta * f1 + (1 - beta) * f2
