"""Seeded synthetic inputs of the BASELINE.json workload shape "S8" (8 frames x 96 patches/frame,
640x480 voxel frames => 160x120 feature maps, C=128, fully connected graph => 6144 edges), in plain
torch on the CPU (fp64, then cast).  No datasets / checkpoints are available offline; weights are
seeded random-init.  Self-contained: does not use the test oracle."""
import torch


def _quat_from_rotvec(phi):
    th = phi.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return torch.cat([torch.sin(0.5 * th) / th * phi, torch.cos(0.5 * th)], dim=-1)


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def _qrot(q, p):
    qv, w = q[..., :3], q[..., 3:]
    uv = 2 * torch.linalg.cross(qv.expand_as(p), p)
    return p + w * uv + torch.linalg.cross(qv.expand_as(p), uv)


def fully_connected_graph(n_frames, patches_per_frame):
    Np = n_frames * patches_per_frame
    kk = torch.arange(Np).repeat_interleave(n_frames)
    jj = torch.arange(n_frames).repeat(Np)
    return kk // patches_per_frame, jj, kk


def make_workload(n_frames=8, patches_per_frame=96, H4=120, W4=160, C=128, dim=384, seed=1234, motion=0.03,
                  noise=0.5, feat_dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64
    Np = n_frames * patches_per_frame
    fx = W4 / 2.0
    intr = torch.tensor([fx, fx, W4 / 2.0, H4 / 2.0], dtype=dt).repeat(n_frames, 1)
    # chained small camera motions (world -> camera)
    t = torch.zeros(n_frames, 3, dtype=dt)
    q = torch.zeros(n_frames, 4, dtype=dt)
    q[:, 3] = 1
    for f in range(1, n_frames):
        dq = _quat_from_rotvec(torch.randn(3, generator=g, dtype=dt) * motion)
        dtv = torch.randn(3, generator=g, dtype=dt) * motion
        q[f] = _qmul(dq, q[f - 1])
        t[f] = _qrot(dq, t[f - 1]) + dtv
    q = q / q.norm(dim=-1, keepdim=True)
    poses_gt = torch.cat([t, q], -1)
    cx = torch.randint(8, W4 - 8, (Np,), generator=g).to(dt)
    cy = torch.randint(8, H4 - 8, (Np,), generator=g).to(dt)
    off = torch.arange(-1, 2).to(dt)
    px = (cx[:, None, None] + off[None, None, :]).expand(Np, 3, 3)
    py = (cy[:, None, None] + off[None, :, None]).expand(Np, 3, 3)
    d = (torch.rand(Np, generator=g, dtype=dt) * 0.8 + 0.2)[:, None, None].expand(Np, 3, 3)
    patches_gt = torch.stack([px, py, d], 1).contiguous()                 # [Np,3,3,3]
    ii, jj, kk = fully_connected_graph(n_frames, patches_per_frame)
    # GT reprojection of the patch centres: Gij = Gj * Gi^-1
    qi, ti, qj, tj = q[ii], t[ii], q[jj], t[jj]
    qi_inv = torch.cat([-qi[:, :3], qi[:, 3:]], -1)
    qij = _qmul(qj, qi_inv)
    tij = tj - _qrot(qij, ti)
    X0 = torch.stack([(cx[kk] - intr[0, 2]) / fx, (cy[kk] - intr[0, 3]) / fx, torch.ones(kk.numel(), dtype=dt)], -1)
    X1 = _qrot(qij, X0) + d[kk, 0, 0][:, None] * tij
    uv = torch.stack([fx * X1[:, 0] / X1[:, 2] + intr[0, 2], fx * X1[:, 1] / X1[:, 2] + intr[0, 3]], -1)
    E = ii.numel()
    targets = uv + noise * torch.randn(E, 2, generator=g, dtype=dt)
    weights = torch.rand(E, 2, generator=g, dtype=dt) * 0.9 + 0.1
    # starting state of the iteration: GT perturbed (poses) and random depth
    pert_q = _quat_from_rotvec(torch.randn(n_frames, 3, generator=g, dtype=dt) * 0.01)
    pert_q[0] = torch.tensor([0, 0, 0, 1.0], dtype=dt)
    pert_t = torch.randn(n_frames, 3, generator=g, dtype=dt) * 0.01
    pert_t[0] = 0
    q0 = _qmul(pert_q, q)
    t0 = _qrot(pert_q, t) + pert_t
    poses0 = torch.cat([t0, q0 / q0.norm(dim=-1, keepdim=True)], -1)
    patches0 = patches_gt.clone()
    patches0[:, 2] = (0.5 * d + 0.5 * torch.rand(Np, generator=g, dtype=dt)[:, None, None])
    fmap = (torch.randn(n_frames, C, H4, W4, generator=g) / 4).to(feat_dtype)
    gmap = (torch.randn(Np, C, 3, 3, generator=g) / 4).to(feat_dtype)
    imap = (torch.randn(Np, dim, generator=g) / 4).to(feat_dtype)
    net = (torch.randn(E, dim, generator=g) * 0.1).to(feat_dtype)
    return dict(n_frames=n_frames, patches_per_frame=patches_per_frame, H4=H4, W4=W4, C=C, dim=dim, E=E,
                intrinsics=intr.float(), poses_gt=poses_gt.float(), poses0=poses0.float(), patches0=patches0.float(),
                targets=targets.float(), weights=weights.float(), ii=ii, jj=jj, kk=kk, fmap=fmap, gmap=gmap,
                imap=imap, net=net)


def make_update_module(seed=1234, dim=384):
    from .update import Update
    torch.manual_seed(seed)
    return Update(3, dim)


def corr_algorithmic_bytes(n_frames, patches_per_frame, E, C, H4, W4, levels, elem_bytes, P=3):
    """every referenced input element read once, every output written once"""
    Np = n_frames * patches_per_frame
    pyr = sum(n_frames * C * (H4 // s) * (W4 // s) for s in levels)
    return elem_bytes * (Np * C * P * P + pyr) + 4 * (E * 2 * P * P) + 8 * (2 * E) + elem_bytes * (E * len(levels) * 49 * P * P)


def gru_flops(E, n_groups_kk, n_groups_ij, dim=384, corr_k=896):
    """dense-layer FLOPs of one Update.forward (enet.py:80-99): per edge corr[0] (K = corr_k, zero-padded) + 16 layers of
    dim x dim (corr[2], corr[5], c1 x2, c2 x2, agg_kk.g/f, agg_ij.g/f, 2 x (gate, res[0], res[2])); per group agg.h"""
    return 2 * E * dim * (corr_k + 16 * dim) + 2 * (n_groups_kk + n_groups_ij) * dim * dim
