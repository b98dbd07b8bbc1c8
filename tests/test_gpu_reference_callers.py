"""GPU: the REFERENCE's own Python callers, unmodified (staged verbatim copy, tests/ref_callers.py), running on this
library through `devo_b200.install_shims()` -- SURVEY 8a-K / 8b: "replace the .so, keep the Python".

  * devo/altcorr/correlation.py, devo/fastba/ba.py  : same wrapper, our backend vs the reference's compiled extension
  * devo/lietorch/groups.py + group_ops.py          : SE3/Sim3 class API on our lietorch_backends vs the golden fixture
  * devo/blocks.py::SoftAgg (torch_scatter shim)    : vs the fused segment kernel
  * devo/enet.py::Update.forward under autocast     : vs devo_gru_update (the fused tcgen05 update operator), S8 + S22
  * devo/devo.py::DEVO.update (the unit of work)    : vs UpdateOperator.step() on the same state
  * devo/devo.py::DEVO.__call__ x 15 (config 4)     : runs unchanged on synthetic voxel frames
"""
import os

import pytest
import torch

import ref_callers
from problems import ba_problem, corr_problem, fully_connected_graph

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_callers.available(), reason="oracle/_ref/devo_py not staged")]


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ---------------------------------------------------------------------------------------------- altcorr / fastba
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_reference_altcorr_wrapper_on_this_library(dtype):
    ns = ref_callers.load()
    Pm = corr_problem(n_frames=4, patches_per_frame=48, seed=3, dtype=dtype)
    args = [Pm["gmap"].cuda(), None, None, Pm["kk"].cuda(), Pm["jj"].cuda(), 3]
    outs = {}
    for kind in ("ref_ext", "ours"):
        ref_callers.use_backend(kind)
        res = []
        for lvl, s in enumerate((1, 4)):
            args[1], args[2] = Pm["pyramid"][lvl].cuda(), (Pm["coords"] / s).cuda()
            res.append(ns.altcorr.corr(*args))
        outs[kind] = torch.stack(res, -1).view(1, len(Pm["kk"]), -1)          # devo.py:217
    ref_callers.use_backend("ours")
    assert outs["ours"].shape == outs["ref_ext"].shape and outs["ours"].dtype == outs["ref_ext"].dtype
    tol = 1e-4 if dtype == torch.float32 else 6e-3      # fp16: the reference accumulates in half
    assert _rel(outs["ours"], outs["ref_ext"]) <= tol, _rel(outs["ours"], outs["ref_ext"])


def test_reference_altcorr_backward_and_patchify_on_this_library():
    ns = ref_callers.load()
    Pm = corr_problem(n_frames=3, patches_per_frame=16, seed=5, dtype=torch.float32, H4=48, W4=64)
    grads = {}
    xy = torch.stack([torch.randint(1, 62, (3, 20)), torch.randint(1, 46, (3, 20))], -1).float().cuda()
    for kind in ("ref_ext", "ours"):
        ref_callers.use_backend(kind)
        g = Pm["gmap"].cuda().requires_grad_(True)
        f = Pm["pyramid"][0].cuda().requires_grad_(True)
        out = ns.altcorr.corr(g, f, Pm["coords"].cuda(), Pm["kk"].cuda(), Pm["jj"].cuda(), 3, 1)
        # (the reference returns a permuted view: randn_like would follow its strides, so draw the noise by shape)
        noise = torch.randn(out.shape, generator=torch.Generator().manual_seed(0)).cuda()
        (out * noise).sum().backward()
        net = Pm["pyramid"][0][0].cuda()
        grads[kind] = (g.grad.clone(), f.grad.clone(), ns.altcorr.patchify(net, xy, 1), ns.altcorr.patchify(net, xy + 0.25, 1))
    ref_callers.use_backend("ours")
    assert _rel(grads["ours"][0], grads["ref_ext"][0]) <= 1e-4
    assert _rel(grads["ours"][1], grads["ref_ext"][1]) <= 1e-4
    assert torch.equal(grads["ours"][2], grads["ref_ext"][2])
    assert _rel(grads["ours"][3], grads["ref_ext"][3]) <= 1e-6


def test_reference_fastba_wrapper_on_this_library():
    """devo/fastba/ba.py: BA / neighbors / reproject through the reference's wrapper, our backend vs its own extension"""
    from oracle import fastba as ofba
    ns = ref_callers.load()
    P = ba_problem(n_frames=8, patches_per_frame=96, seed=11, init="perturbed")
    res = {}
    for kind in ("ref_ext", "ours"):
        ref_callers.use_backend(kind)
        poses = ns.lietorch.SE3(P["poses0"].float().cuda())
        patches, intr = P["patches0"].float().cuda(), P["intrinsics"].float().cuda()
        ii, jj, kk = P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda()
        ix, jx = ns.fastba.neighbors(kk, jj)
        co = ns.fastba.reproject(poses.data, patches, intr, ii, jj, kk)
        ns.fastba.BA(poses, patches, intr, P["targets"].float().cuda(), P["weights"].float().cuda(),
                     torch.as_tensor([1e-4], device="cuda"), ii, jj, kk, 1, 8, 2)
        res[kind] = (ix, jx, co, poses.data.clone(), patches.clone())
    ref_callers.use_backend("ours")
    d = lambda t: t.double()
    po, xo, st = ofba.ba(d(P["poses0"][0]), d(P["patches0"][0]), d(P["intrinsics"][0]), d(P["targets"][0]), d(P["weights"][0]),
                         torch.tensor([1e-4]).float().double(), P["ii"], P["jj"], P["kk"], 1, 8, 2)
    assert st == 0
    assert torch.equal(res["ours"][0], res["ref_ext"][0]) and torch.equal(res["ours"][1], res["ref_ext"][1])
    assert (res["ours"][2] - res["ref_ext"][2]).abs().max().item() <= 1e-3
    e_ref = (res["ref_ext"][3][0].double().cpu() - po).abs().max().item()
    e_ours = (res["ours"][3][0].double().cpu() - po).abs().max().item()
    assert e_ours <= 1e-5 and e_ours <= 1.5 * e_ref + 1e-6, (e_ours, e_ref)
    assert (res["ours"][4][0, :, 2].double().cpu() - xo[:, 2]).abs().max().item() <= 1e-5


# ---------------------------------------------------------------------------------------------- lietorch class API
def test_reference_groups_py_on_this_backend_matches_golden():
    """the reference's groups.py / group_ops.py (autograd Functions included) over OUR lietorch_backends reproduce the
    fixture its own Python produced over the CPU oracle (tests/golden/lie_autograd.pt)"""
    ns = ref_callers.load()
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "lie_autograd.pt"))
    SE3 = ns.lietorch.SE3
    assert SE3.group_id == 3 and ns.lietorch.Sim3.group_id == 4
    torch.manual_seed(0)
    xi = (0.3 * torch.randn(2, 5, 6, dtype=torch.float64, device="cuda")).requires_grad_(True)
    X = SE3.exp(xi)
    Y = SE3.exp(0.2 * torch.randn(2, 5, 6, dtype=torch.float64, device="cuda"))
    p = torch.randn(2, 5, 4, dtype=torch.float64, device="cuda")
    Z = X * Y.inv()
    q = Z.act(p)
    assert torch.allclose((Z * Z.inv()).log(), torch.zeros_like(xi), atol=1e-9)
    assert torch.allclose(Z.matrix() @ p[..., None], q[..., None], atol=1e-9)
    q.square().sum().backward()
    assert torch.isfinite(xi.grad).all() and xi.grad.abs().max() > 0
    # the retraction used by BA (groups.py:153-156) and AdjT (projective_ops.py:96)
    a = torch.randn(2, 5, 6, dtype=torch.float64, device="cuda")
    assert torch.allclose(Z.retr(a).data, (SE3.exp(a) * Z).data, atol=1e-12)
    assert Z.adjT(a).shape == a.shape
    # same values as this package's own class API (independent implementation of the host glue)
    from devo_b200 import lietorch as ours
    X2 = ours.SE3.exp(xi.detach())
    assert torch.equal(X2.data, X.data.detach())
    assert isinstance(g, dict)     # fixture presence (its contents are checked in tests/test_gpu_lie.py)


# ---------------------------------------------------------------------------------------------- SoftAgg / Update
def test_reference_softagg_on_scatter_shim_matches_segment_kernel():
    from devo_b200 import cuda_ba
    ns = ref_callers.load()
    torch.manual_seed(3)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(5, 17)]
    keep = torch.rand(ii.numel(), device="cuda") > 0.15
    ii, jj, kk = ii[keep], jj[keep], kk[keep]
    agg = ns.blocks.SoftAgg(384).cuda().eval()
    x = torch.randn(1, ii.numel(), 384, device="cuda")
    with torch.no_grad():
        for key, ngrp in ((kk, 5 * 17), (ii * 12345 + jj, 25)):
            ref = agg(x, key)
            plan = cuda_ba.GraphPlan(key, torch.zeros_like(key), -1, 1, want_neighbors=False)
            y = cuda_ba.segment_softmax_sum(agg.g(x), agg.f(x), plan, ngrp)
            got = agg.h(y)[:, plan.gid.long()]
            assert _rel(got, ref) <= 2e-5, _rel(got, ref)


def _update_case(nf, m, seed, ragged):
    from devo_b200 import cuda_ba
    torch.manual_seed(seed)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(nf, m)]
    if ragged:
        keep = torch.rand(ii.numel(), device="cuda") > 0.2
        ii, jj, kk = ii[keep], jj[keep], kk[keep]
    E, Np = ii.numel(), nf * m
    imap = (0.25 * torch.randn(1, Np, 384, device="cuda")).half()
    corr = torch.zeros(E, 896, device="cuda", dtype=torch.half)
    corr[:, :882] = (2.0 * torch.randn(E, 882, device="cuda")).half()
    plan_kk = cuda_ba.GraphPlan(kk, jj, Np, nf)
    plan_ij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(ii), -1, 1, want_neighbors=False)
    return ii, jj, kk, E, Np, imap, corr, plan_kk, plan_ij


@pytest.mark.parametrize("nf,m,seed,ragged,state", [(8, 96, 0, False, "f32"), (8, 96, 1, False, "f16"), (22, 96, 7, True, "f32"),
                                                    (4, 24, 2, True, "f32")])
def test_reference_update_forward_vs_fused_update_operator(nf, m, seed, ragged, state):
    """devo/enet.py::Update.forward under autocast (exactly how devo.py:312-316 calls it) vs devo_gru_update.
    state "f16": the half zero-state of the very first update; "f32": the float32 state every later update sees
    (GatedResidual returns float32 under autocast, devo.py:232-233 concatenates half zeros onto it)."""
    from devo_b200.update import PackedUpdateWeights, Update
    ns = ref_callers.load()
    ii, jj, kk, E, Np, imap, corr, plan_kk, plan_ij = _update_case(nf, m, seed, ragged)
    ref_up = ns.enet.Update(3).cuda().eval()
    with torch.no_grad():
        for p in ref_up.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    ours = Update(3).cuda().eval()
    ours.load_state_dict(ref_up.state_dict())        # same keys: reference checkpoints load unchanged
    net = 0.5 * torch.randn(1, E, 384, device="cuda") if state == "f32" else torch.zeros(1, E, 384, device="cuda", dtype=torch.half)
    with torch.no_grad(), torch.autocast("cuda", enabled=True):
        ref_net, (ref_d, ref_w, _) = ref_up(net, imap[:, kk], corr[:, :882].reshape(1, E, 882), None, ii, jj, kk)
    assert ref_net.dtype == torch.float32
    with torch.no_grad():
        out_net, (d, w, _) = ours.forward_mma(net, imap, kk, corr, plan_kk, plan_ij, Np, nf * nf, PackedUpdateWeights(ours, torch.float16, 896))
    torch.cuda.synchronize()
    scale = max(ref_net.abs().max().item(), 1.0)
    assert (out_net.float() - ref_net).abs().max().item() <= 3e-2 * scale
    assert (out_net.float() - ref_net).abs().mean().item() <= 1.5e-3 * scale
    assert (d.float() - ref_d.float()).abs().max().item() <= 2e-2
    assert (w.float() - ref_w.float()).abs().max().item() <= 1e-2


# ---------------------------------------------------------------------------------------------- DEVO.update / __call__
def _make_devo(ns, seed=0):
    torch.manual_seed(seed)
    cfg = ref_callers.default_cfg()
    net = ns.enet.eVONet(patch_selector=cfg.PATCH_SELECTOR.lower())
    return ns.devo.DEVO(cfg, net, evs=True, ht=480, wd=640), cfg


def test_reference_devo_update_vs_update_operator():
    """The unit of work of the benchmark: the body of DEVO.update (devo.py:308-338), run by the reference's own code on
    our backends, against UpdateOperator.step() on the same state (8 frames x 96 patches, fully connected)."""
    from devo_b200.engine import UpdateOperator
    from devo_b200.update import Update
    ns = ref_callers.use_backend("ours")
    slam, cfg = _make_devo(ns, 1)
    nf, M = 8, cfg.PATCHES_PER_FRAME
    P = ba_problem(n_frames=nf, patches_per_frame=M, seed=21, init="perturbed")
    C = corr_problem(n_frames=nf, patches_per_frame=M, seed=21)
    torch.manual_seed(5)
    imap = (torch.randn(nf * M, 384) / 4).half().cuda()
    slam.n, slam.m, slam.is_initialized = nf, nf * M, False
    slam.poses_[:nf] = P["poses0"][0].float().cuda()
    slam.patches_[:nf] = P["patches0"][0].float().cuda().view(nf, M, 3, 3, 3)
    slam.intrinsics_[:nf] = P["intrinsics"][0].float().cuda()
    slam.index_[:nf] = torch.arange(nf, device="cuda")[:, None]
    slam.imap_[:nf] = imap.view(nf, M, 384)
    slam.gmap_[:nf] = C["gmap"][0].cuda().view(nf, M, 128, 3, 3)
    slam.fmap1_[0, :nf] = C["pyramid"][0][0].cuda()
    slam.fmap2_[0, :nf] = C["pyramid"][1][0].cuda()
    slam.ii, slam.jj, slam.kk = P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda()
    E = slam.ii.numel()
    net0 = 0.1 * torch.randn(1, E, 384, device="cuda")
    slam.net = net0.clone()
    # ---- ours
    ours = Update(3).cuda().eval()
    ours.load_state_dict(slam.network.update.state_dict())
    op = UpdateOperator(ours, nf, M, E, 120, 160, t0=1)
    op.poses.copy_(slam.poses_[:nf][None])
    op.patches.copy_(slam.patches_[:nf].view(1, nf * M, 3, 3, 3))
    op.intrinsics.copy_(slam.intrinsics_[:nf][None])
    op.set_graph(slam.ii, slam.jj, slam.kk)
    for f in range(nf):
        op.ingest_frame(f, C["fmap"][0, f].cuda(), C["gmap"][0, f * M:(f + 1) * M].cuda(), imap[f * M:(f + 1) * M])
    op.set_net(net0)
    with torch.no_grad():
        slam.update()
        op.step()
    torch.cuda.synchronize()
    assert int(op.status.item()) == 0
    assert (op.get_net().float() - slam.net.float()).abs().max().item() <= 3e-2 * max(slam.net.abs().max().item(), 1.0)
    assert (op.poses[0] - slam.poses_[:nf]).abs().max().item() <= 5e-4
    assert (op.patches[0].view(nf, M, 3, 3, 3)[:, :, 2] - slam.patches_[:nf, :, 2]).abs().max().item() <= 5e-3


def test_reference_devo_call_runs_unchanged_config4():
    """BASELINE.json config 4: DEVO.__call__ x 15 on synthetic voxel frames [5,480,640] through the reference's own
    Python (patchify -> motion probe -> 12 initial updates -> update + keyframe), on this library's backends."""
    ns = ref_callers.use_backend("ours")
    slam, cfg = _make_devo(ns, 2)
    g = torch.Generator(device="cuda").manual_seed(7)
    intr = torch.tensor([320.0, 320.0, 320.0, 240.0], device="cuda")
    n_updates = [0]
    orig = slam.update

    def counted():
        n_updates[0] += 1
        return orig()
    slam.update = counted
    slam.motion_probe = lambda: 10.0          # random-init weights: force initialisation (devo.py:531-535 skips frames else)
    with torch.no_grad():
        for t in range(15):
            vox = (torch.rand(5, 480, 640, device="cuda", generator=g) < 0.1).float() * torch.randn(5, 480, 640, device="cuda", generator=g)
            slam(float(t), vox, intr)
    torch.cuda.synchronize()
    assert slam.is_initialized and n_updates[0] == 12 + 7
    assert torch.isfinite(slam.poses_[:slam.n]).all() and torch.isfinite(slam.patches_[:slam.n]).all()
    assert slam.net.dtype == torch.float32 and slam.net.shape[1] == slam.ii.numel()
    poses, tstamps = slam.terminate()
    assert poses.shape == (15, 7)


def test_patch_graph_vo_matches_reference_devo_loop():
    """BASELINE.json config 4 through THIS package's frame loop (devo_b200.vo.PatchGraphVO: pixel-major ring-buffer ingest,
    device-side graph analysis and hidden-state bookkeeping, fused update operator) against the reference's own DEVO class
    on the same frames, the same network and the same per-frame front-end outputs: identical edge lists and keyframe
    decisions, poses / depths within the half-precision noise that 19 recurrent updates accumulate."""
    from devo_b200.update import Update
    from devo_b200.vo import PatchGraphVO, frontend_from_reference_network
    ns = ref_callers.use_backend("ours")
    slam, cfg = _make_devo(ns, 3)
    g = torch.Generator(device="cuda").manual_seed(11)
    intr = torch.tensor([320.0, 320.0, 320.0, 240.0], device="cuda")
    frames = [(torch.rand(5, 480, 640, device="cuda", generator=g) < 0.1).float() * torch.randn(5, 480, 640, device="cuda", generator=g)
              for _ in range(15)]
    # one front-end pass per frame, shared by both loops
    fe = frontend_from_reference_network(slam.network, cfg)
    with torch.no_grad():
        cached = [fe(f) for f in frames]
    it = iter(cached)

    def ref_patchify(image, **kw):
        c = next(it)
        return (c["fmap"][None, None].clone(), c["gmap"][None].clone(), c["imap"][None, :, :, None, None].clone(),
                c["patches"][None].clone(), None, c["clr"])
    import types
    net = slam.network
    slam.network = types.SimpleNamespace(patchify=ref_patchify, update=net.update)     # DEVO only calls these two
    slam.motion_probe = lambda: 10.0
    torch.manual_seed(5)
    with torch.no_grad():
        for t, f in enumerate(frames):
            slam(float(t), f, intr)
    # ---- ours
    ours_up = Update(3).cuda().eval()
    ours_up.load_state_dict(net.update.state_dict())
    it2 = iter(cached)
    vo = PatchGraphVO(cfg, ours_up, lambda image: next(it2))
    vo.motion_probe = lambda: 10.0
    torch.manual_seed(5)
    for t, f in enumerate(frames):
        vo(float(t), f, intr)
    torch.cuda.synchronize()
    assert vo.n_updates == 12 + 7 and vo.is_initialized
    assert vo.n == slam.n and vo.m == slam.m
    assert torch.equal(vo.ii, slam.ii) and torch.equal(vo.jj, slam.jj) and torch.equal(vo.kk, slam.kk)
    assert torch.isfinite(vo.poses_[:vo.n]).all()
    dp = (vo.poses_[:vo.n] - slam.poses_[:slam.n]).abs().max().item()
    dd = (vo.patches_[:vo.n, :, 2] - slam.patches_[:slam.n, :, 2]).abs().max().item()
    net_ref = slam.net.float()
    dn = (vo.state.get() - net_ref).abs().mean().item() / max(net_ref.abs().mean().item(), 1e-6)
    print("PatchGraphVO vs reference DEVO after 15 frames: pose %.2e depth %.2e hidden-state rel %.2e" % (dp, dd, dn))
    assert dp <= 5e-2 and dn <= 0.1
    p1, _ = vo.terminate()
    p2, _ = slam.terminate()
    assert p1.shape == p2.shape == (15, 7)


def test_patch_graph_vo_motion_probe_matches_reference():
    """devo.py:241-256 through our loop: same median flow as the reference's motion_probe on the same state"""
    from devo_b200.update import Update
    from devo_b200.vo import PatchGraphVO, frontend_from_reference_network
    ns = ref_callers.use_backend("ours")
    slam, cfg = _make_devo(ns, 4)
    g = torch.Generator(device="cuda").manual_seed(3)
    intr = torch.tensor([320.0, 320.0, 320.0, 240.0], device="cuda")
    frames = [(torch.rand(5, 480, 640, device="cuda", generator=g) < 0.1).float() * torch.randn(5, 480, 640, device="cuda", generator=g)
              for _ in range(2)]
    fe = frontend_from_reference_network(slam.network, cfg)
    with torch.no_grad():
        cached = [fe(f) for f in frames]
    probes = {}
    it = iter(cached)

    def ref_patchify(image, **kw):
        c = next(it)
        return (c["fmap"][None, None].clone(), c["gmap"][None].clone(), c["imap"][None, :, :, None, None].clone(),
                c["patches"][None].clone(), None, c["clr"])
    import types
    net = slam.network
    slam.network = types.SimpleNamespace(patchify=ref_patchify, update=net.update)
    orig = slam.motion_probe
    slam.motion_probe = lambda: probes.setdefault("ref", orig())
    torch.manual_seed(1)
    with torch.no_grad():
        for t, f in enumerate(frames):
            slam(float(t), f, intr)
    ours_up = Update(3).cuda().eval()
    ours_up.load_state_dict(net.update.state_dict())
    it2 = iter(cached)
    vo = PatchGraphVO(cfg, ours_up, lambda image: next(it2))
    orig2 = vo.motion_probe
    vo.motion_probe = lambda: probes.setdefault("ours", orig2())
    torch.manual_seed(1)
    for t, f in enumerate(frames):
        vo(float(t), f, intr)
    a, b = float(probes["ref"]), float(probes["ours"])
    assert abs(a - b) <= 2e-2 * max(abs(a), 1.0), (a, b)
