"""GPU: the fused UpdateOperator iteration equals the op-by-op composition through the
reference-shaped API (devo.py:308-338), and CUDA-graph replay equals eager execution."""
import pytest
import torch

from problems import ba_problem, corr_problem

pytestmark = pytest.mark.gpu


def _build(nf=4, m=24, H=60, W=80, seed=5):
    from devo_b200.engine import UpdateOperator
    from devo_b200.update import Update
    torch.manual_seed(seed)
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=seed, H4=H, W4=W, init="perturbed")
    C = corr_problem(n_frames=nf, patches_per_frame=m, H4=H, W4=W, seed=seed)
    up = Update(3).cuda().eval()
    E = P["ii"].numel()
    op = UpdateOperator(up, nf, m, E, H, W, t0=1)
    op.poses.copy_(P["poses0"].float())
    op.patches.copy_(P["patches0"].float())
    op.intrinsics.copy_(P["intrinsics"].float())
    op.set_graph(P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda())
    imap = (torch.randn(nf * m, 384) / 4).half().cuda()
    for f in range(nf):
        op.ingest_frame(f, C["fmap"][0, f].cuda(), C["gmap"][0, f * m:(f + 1) * m].cuda(), imap[f * m:(f + 1) * m])
    op.set_net(0.1 * torch.randn(1, E, 384, device="cuda"))
    return op, up, P, C, imap


def test_step_equals_op_by_op_composition():
    from devo_b200 import altcorr, fastba, projective_ops as pops, lietorch as lt
    op, up, P, C, imap = _build()
    net0 = op.get_net().clone()
    poses, patches = op.poses.clone(), op.patches.clone()
    ii, jj, kk = op.ii, op.jj, op.kk
    # ---- reference-shaped composition (devo.py:210-223,308-338)
    with torch.no_grad():
        coords = pops.transform(lt.SE3(poses), patches, op.intrinsics, ii, jj, kk).permute(0, 1, 4, 2, 3).contiguous()
        with torch.autocast("cuda", dtype=torch.float16):
            c1 = altcorr.corr(C["gmap"].cuda(), C["pyramid"][0].cuda(), coords / 1, kk, jj, 3)
            c2 = altcorr.corr(C["gmap"].cuda(), C["pyramid"][1].cuda(), coords / 4, kk, jj, 3)
            corr = torch.stack([c1, c2], -1).view(1, len(kk), -1)
            ctx = imap[None][:, kk]
            net, (delta, weight, _) = up(net0, ctx, corr, None, ii, jj, kk)
        target = coords[..., 1, 1] + delta.float()
        fastba.BA(poses, patches, op.intrinsics, target, weight.float(), torch.as_tensor([1e-4], device="cuda"), ii, jj, kk, 1, op.Nf, 2)
    op.step()
    assert int(op.status.item()) == 0
    assert torch.allclose(op.coords, coords, atol=1e-4)
    assert torch.allclose(op.get_net(), net.float(), atol=2e-2, rtol=2e-2)
    assert torch.allclose(op.poses, poses, atol=2e-4) and torch.allclose(op.patches, patches, atol=2e-3)


def test_graph_replay_equals_eager():
    op, up, P, C, imap = _build(seed=8)
    op.snapshot_geometry()
    net0 = op.get_net().clone()
    op.step(reset_geometry=True)
    ref = (op.poses.clone(), op.patches.clone(), op.get_net().clone())
    op.set_net(net0)
    op.capture(reset_geometry=True, warmup=2)
    op.set_net(net0)
    op.replay()
    torch.cuda.synchronize()
    assert torch.equal(op.poses, ref[0]) and torch.equal(op.patches, ref[1]) and torch.equal(op.get_net(), ref[2])
    assert int(op.status_sticky.item()) == 0


def test_overlapped_ingest_equals_serial_ingest():
    """ingest_frame(overlap=True) packs the new frame on a side stream that the iteration joins before the lookup"""
    res = []
    for overlap in (False, True):
        op, up, P, C, imap = _build(seed=9)
        m = op.M
        new = (torch.randn_like(C["fmap"][0, 1]) / 4).cuda()
        op.ingest_frame(1, new, C["gmap"][0, m:2 * m].cuda(), imap[m:2 * m], overlap=overlap)
        op.step()
        torch.cuda.synchronize()
        res.append((op.poses.clone(), op.patches.clone(), op.get_net().clone(), op.levels_pm[0].clone()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_state_arena_refresh_path_equals_set_graph_path():
    """bench.py's end-to-end step refreshes the operator's whole state arena (poses, patches, intrinsics, edge list) from an
    uploaded buffer and must then re-derive what set_graph derives (`refresh_pair_key`): same iteration, bit for bit, as
    installing the same state through set_graph.  (A frame-pair key computed any other way than the engine's breaks the
    radix width the engine promises to the plan.)"""
    op, up, P, C, imap = _build()
    net0 = op.get_net().clone()
    arena0 = op.state_arena.clone()
    op.snapshot_geometry()
    with torch.no_grad():
        op.step(reset_geometry=True)
    a_poses, a_patches, a_net = op.poses.clone(), op.patches.clone(), op.get_net().clone()
    # ---- the same state delivered as one buffer; the edge list arrives with it
    op.set_net(net0)
    op.ii.zero_(); op.jj.zero_(); op.kk.zero_(); op.pair_key.zero_()
    with torch.no_grad():
        op.state_arena.copy_(arena0)
        op.refresh_pair_key(same_graph=True)       # (keeps what set_graph verified about the list: the same program runs)
        op._iteration(reset_geometry=False)
    torch.cuda.synchronize()
    assert int(op.status.item()) == 0
    assert torch.equal(op.poses, a_poses) and torch.equal(op.patches, a_patches) and torch.equal(op.get_net(), a_net)
    nf = op.Nf
    assert int(op.plan_ij.ngroups.item()) == torch.unique(op.ii * nf + op.jj).numel()


def test_kernel_copy_and_ba_prepare_leave_no_copy_engine_work():
    """`_lib.copy_` (devo_copy_bytes) copies any byte count -- 16-byte vectors when pointers and size allow, bytes otherwise --
    and `cuda_ba.prepare` clears the status word and the ticket area of the workspace; both are kernels (counted by the
    library's launch counter), so a captured step holds no memset / memcpy nodes that would queue behind host uploads"""
    from devo_b200 import _lib, cuda_ba
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.randint(0, 255, (1 << 20,), dtype=torch.uint8, device="cuda", generator=g)
    for off_s, off_d, n in ((0, 0, 1 << 20), (0, 0, 4096 + 16), (1, 0, 1000), (0, 3, 77), (16, 32, 48), (5, 5, 1)):
        dst = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        l0 = _lib.launch_count()
        _lib.copy_(dst[off_d:off_d + n], src[off_s:off_s + n])
        assert _lib.launch_count() == l0 + 1
        assert torch.equal(dst[off_d:off_d + n], src[off_s:off_s + n])
        assert int(dst[:off_d].sum()) == 0 and int(dst[off_d + n:].sum()) == 0
    _lib.copy_(dst[:0], src[:0])                                     # nothing to do
    with pytest.raises(RuntimeError):
        _lib.copy_(dst[:8], src[:9])
    with pytest.raises(RuntimeError):
        _lib.copy_(dst[:8].view(torch.float32), src[:8])
    E, nfree = 6144, 7
    ws = torch.full((int(_lib.lib().devo_ba_workspace(E, nfree)),), 255, dtype=torch.uint8, device="cuda")
    status = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    before = ws.clone()
    l0 = _lib.launch_count()
    cuda_ba.prepare(E, nfree, status, ws)
    assert _lib.launch_count() == l0 + 1
    assert int(status) == 0
    changed = (ws != before).nonzero().flatten()
    assert changed.numel() == 128 and int(changed[-1] - changed[0]) == 127 and int(ws[changed].sum()) == 0
