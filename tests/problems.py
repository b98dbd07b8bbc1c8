"""Seeded synthetic problems shared by the tests, bench.py and smoke() (SURVEY 8d recipes).
Everything is generated on the CPU in fp64 and cast, so every implementation sees the same
problem."""
import torch

from oracle import lie as olie
from oracle import pops as opops


def fully_connected_graph(n_frames, patches_per_frame):
    """edges exactly as devo/enet.py:300-301 builds them for initialisation"""
    Np = n_frames * patches_per_frame
    kk = torch.arange(Np).repeat_interleave(n_frames)
    jj = torch.arange(n_frames).repeat(Np)
    ii = kk // patches_per_frame
    return ii, jj, kk


def ba_problem(n_frames=8, patches_per_frame=96, seed=1234, H4=120, W4=160, motion=0.05, noise=0.5,
               init="identity"):
    """config 1/3 of BASELINE.json: GT poses = chained small motions, random patches with
    inverse depth U(0.2,1), targets = GT reprojection of the centres + N(0, noise^2) px,
    weights U(0.1,1); initial poses identity (or GT perturbed), initial depth U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64
    Np = n_frames * patches_per_frame
    fx = W4 / 2.0
    intr = torch.tensor([fx, fx, W4 / 2.0, H4 / 2.0], dtype=dt).view(1, 1, 4).repeat(1, n_frames, 1)
    xi = torch.randn(n_frames, 6, generator=g, dtype=dt) * motion
    xi[0] = 0
    T = [olie.expm(3, xi[0:1])]
    for t in range(1, n_frames):
        T.append(olie.mul(3, olie.expm(3, xi[t:t + 1]), T[-1]))
    poses_gt = torch.cat(T, 0).view(1, n_frames, 7)
    cx = torch.randint(8, W4 - 8, (Np,), generator=g).to(dt)
    cy = torch.randint(8, H4 - 8, (Np,), generator=g).to(dt)
    off = torch.arange(-1, 2).to(dt)
    px = (cx[:, None, None] + off[None, None, :]).expand(Np, 3, 3)
    py = (cy[:, None, None] + off[None, :, None]).expand(Np, 3, 3)
    d = (torch.rand(Np, generator=g, dtype=dt) * 0.8 + 0.2)[:, None, None].expand(Np, 3, 3)
    patches_gt = torch.stack([px, py, d], 1).view(1, Np, 3, 3, 3).contiguous()
    ii, jj, kk = fully_connected_graph(n_frames, patches_per_frame)
    coords_gt = opops.transform(poses_gt, patches_gt, intr, ii, jj, kk)
    E = ii.numel()
    targets = coords_gt[..., 1, 1, :] + noise * torch.randn(1, E, 2, generator=g, dtype=dt)
    weights = torch.rand(1, E, 2, generator=g, dtype=dt) * 0.9 + 0.1
    if init == "identity":
        poses0 = olie.expm(3, torch.zeros(n_frames, 6, dtype=dt)).view(1, n_frames, 7)
    else:
        pert = torch.randn(n_frames, 6, generator=g, dtype=dt) * 0.01
        pert[0] = 0
        poses0 = olie.mul(3, olie.expm(3, pert), poses_gt[0]).view(1, n_frames, 7)
    patches0 = patches_gt.clone()
    patches0[:, :, 2] = torch.rand(Np, generator=g, dtype=dt)[None, :, None, None].expand(1, Np, 3, 3)
    return dict(intrinsics=intr, poses_gt=poses_gt, patches_gt=patches_gt, poses0=poses0, patches0=patches0,
                targets=targets, weights=weights, ii=ii, jj=jj, kk=kk, coords_gt=coords_gt,
                bounds=[-64, -64, W4 + 64, H4 + 64])


def corr_problem(n_frames=8, patches_per_frame=96, C=128, H4=120, W4=160, seed=1234, levels=(1, 4),
                 oob_stress=False, dtype=torch.float16):
    """config 2: gmap ~ N(0,1)/4, pyramid = avg_pool2d of fmap ~ N(0,1)/4, coords = reprojection of grid
    patches under small random motion (or uniform random incl. out-of-bounds)."""
    g = torch.Generator().manual_seed(seed)
    Np = n_frames * patches_per_frame
    gmap = (torch.randn(1, Np, C, 3, 3, generator=g) / 4).to(dtype)
    fmap = (torch.randn(1, n_frames, C, H4, W4, generator=g) / 4).to(dtype)
    pyramid = []
    for s in levels:
        p = torch.nn.functional.avg_pool2d(fmap[0].float(), s, s).to(dtype)
        pyramid.append(p.view(1, n_frames, C, H4 // s, W4 // s))
    P = ba_problem(n_frames, patches_per_frame, seed=seed, H4=H4, W4=W4)
    ii, jj, kk = P["ii"], P["jj"], P["kk"]
    E = ii.numel()
    if oob_stress:
        x = torch.rand(1, E, 1, 3, 3, generator=g) * (W4 + 16) - 8
        y = torch.rand(1, E, 1, 3, 3, generator=g) * (H4 + 16) - 8
        coords = torch.cat([x, y], 2).float()
    else:
        coords = P["coords_gt"].permute(0, 1, 4, 2, 3).contiguous().float()     # [1,E,2,3,3]
    return dict(gmap=gmap, fmap=fmap, pyramid=pyramid, coords=coords, ii=ii, jj=jj, kk=kk, levels=levels)
