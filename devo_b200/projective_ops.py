"""Projective geometry of the update operator -- same functions and argument meaning as
devo/projective_ops.py:19-121 (iproj, proj, transform, point_cloud, flow_mag).

`transform` takes the fused CUDA kernel (one launch: Gj*Gi^-1, act4, projection,
optional centre-pixel Jacobians) whenever no autograd graph is needed; otherwise it is
composed from the differentiable lietorch ops exactly like the reference."""
import torch

from . import _lib
from .lietorch import SE3

MIN_DEPTH = 0.2
FUSED_AUTOGRAD = True      # transform under autograd: one fused forward + one fused backward launch (False: composed path)


def extract_intrinsics(intrinsics):
    return intrinsics[..., None, None, :].unbind(dim=-1)


def coords_grid(ht, wd, **kwargs):
    y, x = torch.meshgrid(torch.arange(ht).to(**kwargs).float(), torch.arange(wd).to(**kwargs).float(), indexing="ij")
    return torch.stack([x, y], dim=-1)


def iproj(patches, intrinsics):
    """pixel (x, y, inverse depth) -> homogeneous ray (X, Y, 1, d)"""
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    X, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=0.1)
    x = fx * (d * X) + cx
    y = fy * (d * Y) + cy
    if depth:
        return torch.stack([x, y, d], dim=-1)
    return torch.stack([x, y], dim=-1)


def _fused_ok(poses, patches, intrinsics, depth, index_tensors=(), allow_grad=False):
    if depth or not isinstance(poses, SE3):
        return False
    ts = (poses.data, patches, intrinsics)
    if any((not t.is_cuda) or t.dtype != torch.float32 for t in ts):
        return False
    # the kernel dereferences ii/jj/kk as device int64 arrays; the reference API also accepts CPU / int32 index tensors
    # (`patches[:, kk]` works with them), which must take the composed path
    if any((not torch.is_tensor(t)) or (not t.is_cuda) or t.dtype != torch.int64 or t.device != patches.device for t in index_tensors):
        return False
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        if not allow_grad or intrinsics.requires_grad:       # the fused backward has no intrinsics gradient
            return False
    return poses.data.dim() == 3 and poses.data.shape[0] == 1 and patches.dim() == 5 and patches.shape[0] == 1


def transform_fused(poses_data, patches, intrinsics, ii, jj, kk, jacobian=False, valid=False, tonly=False, layout=0):
    """raw entry: poses_data [1,N,7]; returns coords ([1,E,P,P,2] if layout==0 else [1,E,2,P,P]) [, valid, (Ji,Jj,Jz)]"""
    poses_data = poses_data.contiguous()
    patches = patches.contiguous()
    intrinsics = intrinsics.contiguous()
    ii, jj, kk = ii.contiguous(), jj.contiguous(), kk.contiguous()
    E = ii.numel()
    P = patches.shape[-1]
    dev = patches.device
    shape = (1, E, P, P, 2) if layout == 0 else (1, E, 2, P, P)
    coords = torch.empty(shape, dtype=torch.float32, device=dev)
    v = torch.empty(1, E, dtype=torch.float32, device=dev) if (jacobian or valid) else None
    Ji = torch.empty(1, E, 2, 6, dtype=torch.float32, device=dev) if jacobian else None
    Jj = torch.empty(1, E, 2, 6, dtype=torch.float32, device=dev) if jacobian else None
    Jz = torch.empty(1, E, 2, 1, dtype=torch.float32, device=dev) if jacobian else None
    _lib.check(_lib.lib().devo_transform_forward(
        poses_data.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
        coords.data_ptr(), _lib.ptr(v), _lib.ptr(Ji), _lib.ptr(Jj), _lib.ptr(Jz), E, P, layout, int(bool(tonly)),
        _lib.stream_ptr(dev)), "transform_forward")
    if jacobian:
        return coords, v, (Ji, Jj, Jz)
    if valid:
        return coords, v
    return coords


class _TransformFn(torch.autograd.Function):
    """projective_ops.transform (+ centre-pixel Jacobians) as ONE fused forward and ONE fused backward launch
    (csrc/ba.cu::transform_kernel, csrc/transform_grad.cu).  The pose gradient follows lietorch's convention (gradient
    w.r.t. a left tangent perturbation, slots 0..5 of the 7), so it composes with the group_ops Functions."""

    @staticmethod
    def forward(ctx, poses_data, patches, intrinsics, ii, jj, kk, jacobian, tonly):
        poses_data, patches, intrinsics = poses_data.contiguous(), patches.contiguous(), intrinsics.contiguous()
        ii, jj, kk = ii.contiguous(), jj.contiguous(), kk.contiguous()
        coords, v, (Ji, Jj, Jz) = transform_fused(poses_data, patches, intrinsics, ii, jj, kk, jacobian=True, tonly=tonly)
        ctx.save_for_backward(poses_data, patches, intrinsics, ii, jj, kk)
        ctx.tonly, ctx.jacobian = bool(tonly), bool(jacobian)
        ctx.mark_non_differentiable(v)
        return coords, v, Ji, Jj, Jz

    @staticmethod
    def backward(ctx, g_coords, _gv, g_Ji, g_Jj, g_Jz):
        poses_data, patches, intrinsics, ii, jj, kk = ctx.saved_tensors
        gp = torch.zeros_like(poses_data)
        gx = torch.zeros_like(patches)

        def c(t):
            return None if t is None else t.contiguous().float()
        g_coords, g_Ji, g_Jj, g_Jz = c(g_coords), c(g_Ji), c(g_Jj), c(g_Jz)
        _lib.check(_lib.lib().devo_transform_backward(
            poses_data.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
            _lib.ptr(g_coords), _lib.ptr(g_Ji), _lib.ptr(g_Jj), _lib.ptr(g_Jz), gp.data_ptr(), gx.data_ptr(),
            ii.numel(), patches.shape[-1], 0, int(ctx.tonly), _lib.stream_ptr(patches.device)), "transform_backward")
        return gp, gx, None, None, None, None, None, None


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    """reproject patch kk from frame ii into frame jj -> [b,E,P,P,2] (+valid / +Jacobians)"""
    if _fused_ok(poses, patches, intrinsics, depth, (ii, jj, kk)):
        return transform_fused(poses.data, patches, intrinsics, ii, jj, kk, jacobian=jacobian, valid=valid, tonly=tonly)
    # (translation-only + autograd keeps the composed path: the reference overwrites the rotation slots of Gij IN PLACE,
    #  which cuts the tangent-space gradient in a way only that graph reproduces; flow_mag, its one caller, runs without grad)
    if FUSED_AUTOGRAD and not tonly and _fused_ok(poses, patches, intrinsics, depth, (ii, jj, kk), allow_grad=True):
        coords, v, Ji, Jj, Jz = _TransformFn.apply(poses.data, patches, intrinsics, ii, jj, kk, jacobian, tonly)
        if jacobian:
            return coords, v, (Ji, Jj, Jz)
        if valid:
            return coords, v
        return coords

    X0 = iproj(patches[:, kk], intrinsics[:, ii])
    Gij = poses[:, jj] * poses[:, ii].inv()
    if tonly:
        Gij[..., 3:] = torch.as_tensor([0, 0, 0, 1], device=Gij.device)
    X1 = Gij[:, :, None, None] * X0
    p = X1.shape[2]
    x1 = proj(X1, intrinsics[:, jj], depth)

    if jacobian:
        # centre-pixel Jacobians, written out in closed form (reference: :73-100 builds Ja (4x6) and
        # Jp (2x4) and multiplies them):  Jj = d pi / d xi_j,  Ji = -Ad(Gij)^T Jj,  Jz = d pi / d depth
        X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
        fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
        zero = torch.zeros_like(Z)
        d = torch.where(Z.abs() > 0.2, 1.0 / torch.where(Z.abs() > 0.2, Z, torch.ones_like(Z)), zero)
        ax, az = fx * d, -fx * X * d * d
        by, bz = fy * d, -fy * Y * d * d
        Jj = torch.stack([
            torch.stack([ax * H, zero, az * H, az * Y, ax * Z - az * X, -ax * Y], dim=-1),
            torch.stack([zero, by * H, bz * H, bz * Y - by * Z, -bz * X, by * X], dim=-1)], dim=-2)
        Ji = -Gij[:, :, None].adjT(Jj)
        t = Gij.translation()[..., :3]   # through Act4 so the tangent-space gradient convention holds
        Jz = torch.stack([ax * t[..., 0] + az * t[..., 2], by * t[..., 1] + bz * t[..., 2]], dim=-1).unsqueeze(-1)
        return x1, (Z > 0.2).float(), (Ji, Jj, Jz)
    if valid:
        return x1, (X1[..., p // 2, p // 2, 2] > 0.2).float()
    return x1


def point_cloud(poses, patches, intrinsics, ix):
    return poses[:, ix, None, None].inv() * iproj(patches, intrinsics[:, ix])


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    c0 = transform(poses, patches, intrinsics, ii, ii, kk)
    c1 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=False)
    c2 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    return beta * (c1 - c0).norm(dim=-1) + (1 - beta) * (c2 - c0).norm(dim=-1)
