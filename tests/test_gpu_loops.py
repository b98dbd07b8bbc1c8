"""GPU: the two loop-shaped configurations of BASELINE.json at test scale.

  * config 4 "full DEVO update loop": several consecutive update iterations through the fused engine (recurrent hidden
    state, in-place BA) against the op-by-op composition through the reference-shaped API (devo/devo.py:308-338).
  * config 5 "training step": one iteration of the training loop body (devo/enet.py:341-372: corr with autograd ->
    Update -> two differentiable ba.BA steps -> reprojection loss) -- gradients reach the feature maps and the update
    operator's parameters, and the fp32 gradients agree with the same graph evaluated in fp64."""
import pytest
import torch

from problems import ba_problem, corr_problem

pytestmark = pytest.mark.gpu


def test_update_loop_three_iterations_equals_composition():
    from devo_b200 import altcorr, fastba, projective_ops as pops, lietorch as lt
    from test_gpu_engine import _build
    op, up, P, C, imap = _build(seed=11)
    net = op.get_net().clone()
    poses, patches = op.poses.clone(), op.patches.clone()
    ii, jj, kk = op.ii, op.jj, op.kk
    lm = torch.as_tensor([1e-4], device="cuda")
    with torch.no_grad():
        for _ in range(3):
            coords = pops.transform(lt.SE3(poses), patches, op.intrinsics, ii, jj, kk).permute(0, 1, 4, 2, 3).contiguous()
            with torch.autocast("cuda", dtype=torch.float16):
                c1 = altcorr.corr(C["gmap"].cuda(), C["pyramid"][0].cuda(), coords / 1, kk, jj, 3)
                c2 = altcorr.corr(C["gmap"].cuda(), C["pyramid"][1].cuda(), coords / 4, kk, jj, 3)
                corr = torch.stack([c1, c2], -1).view(1, len(kk), -1)
                net, (delta, weight, _) = up(net, imap[None][:, kk], corr, None, ii, jj, kk)
            target = coords[..., 1, 1] + delta.float()
            fastba.BA(poses, patches, op.intrinsics, target, weight.float(), lm, ii, jj, kk, 1, op.Nf, 2)
            op.step()
            assert int(op.status.item()) == 0
    # three recurrent iterations: half-precision noise of the GRU feeds back through BA, so the bound is looser than
    # for a single step (tests/test_gpu_engine.py)
    assert torch.allclose(op.poses, poses, atol=2e-3), (op.poses - poses).abs().max().item()
    assert torch.allclose(op.patches, patches, atol=2e-2), (op.patches - patches).abs().max().item()
    assert (op.get_net() - net.float()).abs().mean().item() < 2e-2


def _training_graph(dtype, seed=3):
    """loss of one training-loop iteration, with leaves (fmap, gmap, update parameters) that require grad"""
    from devo_b200 import altcorr, ba as dba, lietorch as lt, projective_ops as pops
    from devo_b200.update import Update
    nf, m, H, W = 4, 12, 60, 80
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=seed, H4=H, W4=W, init="perturbed")
    C = corr_problem(n_frames=nf, patches_per_frame=m, H4=H, W4=W, seed=seed, dtype=torch.float32)
    torch.manual_seed(seed)
    up = Update(3).cuda().to(dtype)
    ii, jj, kk = P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda()
    E = ii.numel()
    gmap = C["gmap"].cuda().to(dtype).requires_grad_(True)
    fmap = C["fmap"].cuda().to(dtype).requires_grad_(True)
    pyr = [fmap, torch.nn.functional.avg_pool2d(fmap[0], 4, 4)[None]]
    imap = (torch.randn(1, nf * m, 384, device="cuda") / 4).to(dtype)
    net = torch.zeros(1, E, 384, device="cuda", dtype=dtype)
    poses = lt.SE3(P["poses0"].cuda().to(dtype))
    patches = P["patches0"].cuda().to(dtype)
    intr = P["intrinsics"].cuda().to(dtype)
    coords = pops.transform(poses, patches, intr, ii, jj, kk).permute(0, 1, 4, 2, 3).contiguous()
    c1 = altcorr.corr(gmap, pyr[0], (coords / 1).float(), kk, jj, 3)
    c2 = altcorr.corr(gmap, pyr[1], (coords / 4).float(), kk, jj, 3)
    corr = torch.stack([c1, c2], -1).view(1, E, -1)
    net, (delta, weight, _) = up(net, imap[:, kk], corr, None, ii, jj, kk)
    target = coords[..., 1, 1].detach() + delta
    bounds = [-64, -64, W + 64, H + 64]
    for _ in range(2):
        poses, patches = dba.BA(poses, patches, intr, target, weight, 1e-4, ii, jj, kk, bounds, ep=10.0, fixedp=1)
    gt = pops.transform(lt.SE3(P["poses_gt"].cuda().to(dtype)), P["patches_gt"].cuda().to(dtype), intr, ii, jj, kk)
    est = pops.transform(poses, patches, intr, ii, jj, kk)
    loss = (est - gt).norm(dim=-1).mean()
    return loss, gmap, fmap, up


def test_training_step_gradients_fp32_match_fp64():
    loss64, g64, f64, up64 = _training_graph(torch.float64)
    loss64.backward()
    loss32, g32, f32, up32 = _training_graph(torch.float32)
    loss32.backward()
    assert torch.isfinite(loss32) and abs(loss32.item() - loss64.item()) <= 1e-3 * max(1.0, abs(loss64.item()))
    pairs = [("gmap", g32.grad, g64.grad), ("fmap", f32.grad, f64.grad)]
    pairs += [(n, p32.grad, p64.grad) for (n, p32), (_, p64) in zip(up32.named_parameters(), up64.named_parameters())
              if p64.grad is not None]
    assert len(pairs) > 20
    for name, a, b in pairs:
        assert a is not None and torch.isfinite(a).all(), name
        scale = b.abs().max().item()
        if scale > 0:
            assert (a.double() - b).abs().max().item() <= 2e-2 * scale + 1e-7, (name, (a.double() - b).abs().max().item(), scale)
    assert g64.grad.abs().max().item() > 0 and f64.grad.abs().max().item() > 0       # the loss really reaches the features
