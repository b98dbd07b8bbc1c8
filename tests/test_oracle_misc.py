"""CPU: corr / patchify / neighbors / fastba / scatter oracles -- literal small-case loops,
edge cases and algebraic properties."""
import numpy as np
import pytest
import torch

from oracle import corr as ocorr
from oracle import fastba as ofba
from oracle import neighbors as onb
from oracle import scatter as osc
from problems import ba_problem


def _corr_loops(f1, f2, coords, ii, jj, R):
    """literal restatement of correlation_kernel.cu:82-136 + host blend :221-232 (tiny cases)"""
    B, E, _, P, _ = coords.shape
    C, H, W = f1.shape[2], f2.shape[3], f2.shape[4]
    D = 2 * R + 2
    V = torch.zeros(B, E, D, D, P, P, dtype=torch.float64)
    for b in range(B):
        for e in range(E):
            for i0 in range(P):
                for j0 in range(P):
                    x, y = coords[b, e, 0, i0, j0].item(), coords[b, e, 1, i0, j0].item()
                    for a in range(D):
                        for bb in range(D):
                            i1 = int(np.floor(y)) + a - R
                            j1 = int(np.floor(x)) + bb - R
                            if 0 <= i1 < H and 0 <= j1 < W:
                                V[b, e, a, bb, i0, j0] = (f1[b, ii[e], :, i0, j0].double() * f2[b, jj[e], :, i1, j1].double()).sum()
    x = coords[:, :, 0, None, None].double()
    y = coords[:, :, 1, None, None].double()
    dx, dy = x - x.floor(), y - y.floor()
    out = (1 - dx) * (1 - dy) * V[:, :, :D - 1, :D - 1] + dx * (1 - dy) * V[:, :, :D - 1, 1:] \
        + (1 - dx) * dy * V[:, :, 1:, :D - 1] + dx * dy * V[:, :, 1:, 1:]
    return out.permute(0, 1, 3, 2, 4, 5)


@pytest.mark.parametrize("R,P", [(1, 3), (3, 3), (2, 1)])
def test_corr_forward_matches_loops(R, P):
    torch.manual_seed(R * 10 + P)
    B, Np, Nf, C, H, W, E = 1, 4, 3, 8, 9, 11, 5
    f1 = torch.randn(B, Np, C, P, P)
    f2 = torch.randn(B, Nf, C, H, W)
    coords = torch.cat([torch.rand(B, E, 1, P, P) * (W + 6) - 3, torch.rand(B, E, 1, P, P) * (H + 6) - 3], 2)
    ii = torch.randint(0, Np, (E,))
    jj = torch.randint(0, Nf, (E,))
    got = ocorr.corr_forward(f1, f2, coords, ii, jj, R)
    ref = _corr_loops(f1, f2, coords, ii, jj, R)
    assert got.shape == (B, E, 2 * R + 1, 2 * R + 1, P, P)
    assert torch.allclose(got, ref, atol=1e-12)


def test_corr_backward_is_adjoint():
    """<corr(f1,f2), g> is bilinear: check the oracle's backward against finite differences"""
    torch.manual_seed(3)
    f1 = torch.randn(1, 3, 8, 3, 3, dtype=torch.float64)
    f2 = torch.randn(1, 2, 8, 7, 9, dtype=torch.float64)
    coords = torch.cat([torch.rand(1, 4, 1, 3, 3) * 9, torch.rand(1, 4, 1, 3, 3) * 7], 2)
    ii = torch.tensor([0, 2, 1, 0])
    jj = torch.tensor([1, 0, 1, 1])
    g = torch.randn(1, 4, 3, 3, 3, 3, dtype=torch.float64)
    g1, g2 = ocorr.corr_backward(f1, f2, coords, ii, jj, g, 1)
    d1 = torch.randn_like(f1)
    lhs = (ocorr.corr_forward(f1 + 1e-6 * d1, f2, coords, ii, jj, 1) - ocorr.corr_forward(f1 - 1e-6 * d1, f2, coords, ii, jj, 1)) / 2e-6
    assert torch.allclose((lhs * g).sum(), (g1 * d1).sum(), atol=1e-6)
    d2 = torch.randn_like(f2)
    lhs = (ocorr.corr_forward(f1, f2 + 1e-6 * d2, coords, ii, jj, 1) - ocorr.corr_forward(f1, f2 - 1e-6 * d2, coords, ii, jj, 1)) / 2e-6
    assert torch.allclose((lhs * g).sum(), (g2 * d2).sum(), atol=1e-6)


def test_patchify_copy_and_oob():
    torch.manual_seed(5)
    net = torch.randn(2, 5, 6, 7)
    coords = torch.tensor([[[0.0, 0.0], [6.0, 5.0], [3.4, 2.6], [-5.0, 1.0]]]).repeat(2, 1, 1)
    for R in (0, 1, 2):
        p = ocorr.patchify_forward(net, coords, R)
        D = 2 * R + 2
        assert p.shape == (2, 4, 5, D, D)
        for m in range(4):
            x0, y0 = int(np.floor(coords[0, m, 0])), int(np.floor(coords[0, m, 1]))
            for a in range(D):
                for b in range(D):
                    i, j = y0 + a - R, x0 + b - R
                    exp = net[1, :, i, j] if (0 <= i < 6 and 0 <= j < 7) else torch.zeros(5)
                    assert torch.equal(p[1, m, :, a, b], exp)
    # integer coords + bilinear == exact copy of the top-left (2R+1)^2 window (correlation.py:56-66)
    ci = coords.floor()
    assert torch.equal(ocorr.patchify(net, ci, 1), ocorr.patchify_forward(net, ci, 1)[..., :3, :3])
    # backward is the adjoint of forward
    g = torch.randn(2, 4, 5, 4, 4)
    gb = ocorr.patchify_backward(net, coords, g, 1)
    assert torch.allclose((ocorr.patchify_forward(net, coords, 1) * g).sum(), (gb * net).sum(), atol=1e-4)


def test_neighbors_matches_literal_loops_and_edge_cases():
    rng = np.random.RandomState(0)
    for E, nk, nj in [(1, 1, 1), (7, 2, 3), (64, 5, 4), (500, 37, 9)]:
        kk = torch.from_numpy(rng.randint(0, nk, E))
        jj = torch.from_numpy(rng.randint(0, nj, E))    # duplicates => stability matters
        a = onb.neighbors(kk, jj)
        b = onb.neighbors_loops(kk, jj)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ix, jx = onb.neighbors(torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long))
    assert ix.numel() == 0 and jx.numel() == 0
    # S8 structure: 8 consecutive edges per patch, already ordered by frame
    kk = torch.arange(6).repeat_interleave(4)
    jj = torch.arange(4).repeat(6)
    ix, jx = onb.neighbors(kk, jj)
    e = torch.arange(24)
    assert torch.equal(ix, torch.where(e % 4 == 0, -1, e - 1)) and torch.equal(jx, torch.where(e % 4 == 3, -1, e + 1))


def test_fastba_oracle_properties():
    P = ba_problem(n_frames=4, patches_per_frame=16, seed=3, init="perturbed", noise=0.2)
    lm = torch.tensor([1e-4])
    a = (P["intrinsics"], P["targets"], P["weights"], lm, P["ii"], P["jj"], P["kk"])

    def cost(poses, patches):
        T = ofba.edge_terms(poses.reshape(-1, 7), patches.reshape(-1, 3, 3, 3), P["intrinsics"].reshape(-1, 4),
                            P["targets"].reshape(-1, 2), P["weights"].reshape(-1, 2), P["ii"], P["jj"], P["kk"])
        return (T["w"] * T["r"] ** 2).sum()

    c0 = cost(P["poses0"], P["patches0"])
    p1, x1, st = ofba.ba(P["poses0"], P["patches0"][0], *a, 1, 4, 4)
    assert st == 0 and cost(p1, x1) < 0.5 * c0
    # fixed poses do not move; pose 0 (< t0) is fixed
    assert torch.equal(p1[0], P["poses0"][0, 0].double())
    p2, x2, _ = ofba.ba(P["poses0"], P["patches0"][0], *a, 2, 4, 1)
    assert torch.equal(p2[:2], P["poses0"][0, :2].double())
    # structure only (t1-t0 == 0): poses untouched, depths move
    p3, x3, _ = ofba.ba(P["poses0"], P["patches0"][0], *a, 4, 4, 1)
    assert torch.equal(p3, P["poses0"][0].double()) and not torch.equal(x3, P["patches0"][0].double())
    # all P*P depth entries of a patch are equal after the update, x/y untouched
    assert torch.equal(x3[:, 2], x3[:, 2, :1, :1].expand(-1, 3, 3)) and torch.equal(x3[:, :2], P["patches0"][0, :, :2].double())
    # zero iterations is the identity
    p4, x4, _ = ofba.ba(P["poses0"], P["patches0"][0], *a, 1, 4, 0)
    assert torch.equal(p4, P["poses0"][0].double())
    # reproject == edge_terms coords at the centre pixel
    rp = ofba.reproject(P["poses0"], P["patches0"][0], P["intrinsics"], P["ii"], P["jj"], P["kk"])
    T = ofba.edge_terms(P["poses0"].reshape(-1, 7), P["patches0"].reshape(-1, 3, 3, 3), P["intrinsics"].reshape(-1, 4),
                        P["targets"].reshape(-1, 2), P["weights"].reshape(-1, 2), P["ii"], P["jj"], P["kk"])
    assert torch.allclose(rp[0, :, :, 1, 1], T["coords"], atol=1e-12)


def test_scatter_ops():
    torch.manual_seed(1)
    x = torch.randn(1, 10, 4)
    idx = torch.tensor([0, 2, 2, 1, 0, 0, 3, 3, 3, 1])
    s = osc.scatter_sum(x, idx, dim=1, dim_size=5)
    for g in range(5):
        assert torch.allclose(s[0, g], x[0, idx == g].sum(0), atol=1e-6)
    w = osc.scatter_softmax(x, idx, dim=1)
    for g in range(4):
        assert torch.allclose(w[0, idx == g], torch.softmax(x[0, idx == g], dim=0), atol=1e-6)


def test_c_corr_oracle_matches_torch_oracle():
    """oracle/corr_c.c (the C + OpenMP restatement bench.py's CPU arm times) against oracle/corr.py (pinned on the GPU box
    against the reference's compiled extension): reprojected and out-of-bounds coordinates, both pyramid levels"""
    from oracle import corr as ocorr, corr_c
    from problems import corr_problem
    corr_c.build()
    for oob in (False, True):
        P = corr_problem(n_frames=3, patches_per_frame=16, seed=3, dtype=torch.float32, H4=48, W4=64, oob_stress=oob)
        for l, s in enumerate((1, 4)):
            a = ocorr.corr_forward(P["gmap"], P["pyramid"][l], P["coords"] / s, P["kk"], P["jj"], 3)
            b = corr_c.corr_forward(P["gmap"], P["pyramid"][l], P["coords"] / s, P["kk"], P["jj"], 3)
            assert b.shape == a.shape
            assert (a - b.double()).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1.0)


def test_voxel_oracle_matches_reference_golden():
    """oracle/voxel.py against tests/golden/voxel.pt, generated by the reference's own utils/voxel_utils.py::std / rescale and
    utils/event_utils.py::to_voxel_grid (tests/golden/make_golden_voxel.py)"""
    import os
    from oracle import voxel as ovox
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "voxel.pt"), weights_only=False)
    assert torch.equal(ovox.std(g["vox"]), g["std_seq"])
    assert torch.equal(ovox.std(g["vox"], sequence=False), g["std_frame"])
    assert torch.equal(ovox.rescale(g["vox"]), g["rescale"])
    e = g["events"]
    assert torch.equal(ovox.to_voxel_grid(e["xs"], e["ys"], e["ts"], e["ps"].copy(), H=24, W=32, nb_of_time_bins=5), g["grid"])
    # an empty group leaves std untouched (voxel_utils.py:18)
    z = g["vox"].clone()
    z[1] = 0
    assert torch.equal(ovox.std(z), z)
