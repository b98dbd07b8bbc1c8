"""GPU parity: fastba.BA / neighbors / reproject / graph plan / projective transform /
differentiable ba.BA (all through the C ABI) vs the CPU oracle.
Tolerances: index outputs bit-exact; BA pose / inverse-depth updates <= 1e-5 absolute on
well-conditioned synthetic graphs (BASELINE.json north_star); transform <= 1e-4 px."""
import numpy as np
import pytest
import torch

from oracle import ba as oba
from oracle import fastba as ofba
from oracle import neighbors as onb
from oracle import pops as opops
from problems import ba_problem, fully_connected_graph

pytestmark = pytest.mark.gpu


def _run_ba(P, t0, t1, iters, poses=None, patches=None):
    from devo_b200 import fastba
    poses = (P["poses0"] if poses is None else poses).float().cuda().contiguous()
    patches = (P["patches0"] if patches is None else patches).float().cuda().contiguous()
    fastba.BA(poses, patches, P["intrinsics"].float().cuda(), P["targets"].float().cuda(), P["weights"].float().cuda(),
              torch.tensor([1e-4], device="cuda"), P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda(), t0, t1, iters)
    return poses, patches


def _oracle_ba(P, t0, t1, iters):
    # same fp32-rounded inputs as the GPU sees, arithmetic in fp64
    f = lambda t: t.float().double()
    return ofba.ba(f(P["poses0"]), f(P["patches0"][0]), f(P["intrinsics"]), f(P["targets"]), f(P["weights"]),
                   torch.tensor([1e-4]).float().double(), P["ii"], P["jj"], P["kk"], t0, t1, iters)


@pytest.mark.parametrize("nf,m,t0,iters", [(2, 32, 1, 2), (4, 24, 1, 2), (8, 96, 1, 2), (8, 96, 1, 10), (6, 16, 3, 3),
                                            (13, 10, 1, 2), (21, 6, 1, 2), (26, 5, 1, 2)])   # 12 / 20 / 25 free poses: every entries-per-thread variant
def test_ba_vs_oracle(nf, m, t0, iters):
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=1234 + nf, init="perturbed", noise=0.3)
    poses, patches = _run_ba(P, t0, nf, iters)
    po, xo, st = _oracle_ba(P, t0, nf, iters)
    assert st == 0
    dp = (poses[0].double().cpu() - po).abs().max().item()
    dd = (patches[0, :, 2].double().cpu() - xo[:, 2]).abs().max().item()
    assert dp <= 1e-5, dp
    assert dd <= 1e-5, dd
    # poses outside [t0,t1) and the x/y channels are untouched, depth is constant over a patch
    assert torch.equal(poses[0, :t0].cpu(), P["poses0"][0, :t0].float())
    assert torch.equal(patches[0, :, :2].cpu(), P["patches0"][0, :, :2].float())
    assert torch.equal(patches[0, :, 2], patches[0, :, 2, :1, :1].expand(-1, 3, 3))


@pytest.mark.parametrize("seed", [0, 1])
def test_ba_arbitrary_edge_lists_vs_oracle(seed):
    """The C ABI takes ANY (ii, jj, kk): edges of one patch leaving from different frames (no per-patch pre-sum: the
    block-pair lists carry the diagonal terms), self edges i == j, duplicated edges, patches with 1 or 40 edges, a ragged
    random subset of a patch graph -- all against the fp64 oracle.  (The DEVO graph, where every edge of a patch leaves
    from the patch's frame, is the pre-summed fast case of the other tests.)"""
    nf, m = 7, 24
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=77 + seed, init="perturbed", noise=0.3)
    g = torch.Generator().manual_seed(seed)
    E0 = P["ii"].numel()
    keep = torch.rand(E0, generator=g) < 0.6                             # ragged subset
    keep[:3] = True
    sel = torch.nonzero(keep).flatten()
    ii, jj, kk = P["ii"][sel].clone(), P["jj"][sel].clone(), P["kk"][sel].clone()
    tg, wt = P["targets"][:, sel].clone(), P["weights"][:, sel].clone()
    # a third of the edges get a random source frame (the oracle and the kernel read poses[ii], whatever it is)
    rnd = torch.rand(ii.numel(), generator=g) < 0.33
    ii[rnd] = torch.randint(0, nf, (int(rnd.sum()),), generator=g)
    # some self edges and some exact duplicates
    selfe = torch.randperm(ii.numel(), generator=g)[:10]
    jj[selfe] = ii[selfe]
    dup = torch.randperm(ii.numel(), generator=g)[:25]
    ii, jj, kk = torch.cat([ii, ii[dup]]), torch.cat([jj, jj[dup]]), torch.cat([kk, kk[dup]])
    tg, wt = torch.cat([tg, tg[:, dup]], 1), torch.cat([wt, wt[:, dup]], 1)
    # one patch with 40 edges
    big = torch.randint(0, nf, (40,), generator=g)
    ii = torch.cat([ii, torch.full((40,), 2)]); jj = torch.cat([jj, big]); kk = torch.cat([kk, torch.full((40,), 2 * m + 5)])
    tg = torch.cat([tg, tg[:, :40]], 1); wt = torch.cat([wt, wt[:, :40] * 0.1], 1)
    # targets consistent with the edited graph: reproject with the ground truth, add noise
    cg = opops.transform(P["poses_gt"], P["patches_gt"], P["intrinsics"], ii, jj, kk)
    tg = cg[..., 1, 1, :] + 0.3 * torch.randn(1, ii.numel(), 2, generator=g, dtype=torch.float64)
    Q = dict(P, ii=ii, jj=jj, kk=kk, targets=tg, weights=wt)
    for t0, iters in ((1, 2), (2, 3)):
        poses, patches = _run_ba(Q, t0, nf, iters)
        po, xo, st = _oracle_ba(Q, t0, nf, iters)
        assert st == 0
        dp = (poses[0].double().cpu() - po).abs().max().item()
        dd = (patches[0, :, 2].double().cpu() - xo[:, 2]).abs().max().item()
        assert dp <= 1e-5, dp
        assert dd <= 1e-5, dd


def test_ba_from_identity_converges_like_oracle():
    """config 3 of BASELINE.json (8 keyframes x 96 patches, 10 GN iterations from identity poses)"""
    P = ba_problem(n_frames=8, patches_per_frame=96, seed=1234)
    poses, patches = _run_ba(P, 1, 8, 10)
    po, xo, st = _oracle_ba(P, 1, 8, 10)
    assert st == 0
    # after 10 iterations from a poor start, fp32 Jacobians vs fp64 differ a little more
    assert (poses[0].double().cpu() - po).abs().max().item() <= 2e-4
    c = opops.transform(poses.double().cpu(), patches.double().cpu(), P["intrinsics"], P["ii"], P["jj"], P["kk"])
    assert (c[..., 1, 1, :] - P["targets"]).norm(dim=-1).median() < 1.0


def test_ba_structure_only_zero_iterations_and_buffer_views():
    from devo_b200 import fastba
    P = ba_problem(n_frames=4, patches_per_frame=16, seed=3, init="perturbed")
    poses, patches = _run_ba(P, 4, 4, 1)            # t1 - t0 == 0 => structure only (ba_cuda.cu:494-506)
    po, xo, _ = _oracle_ba(P, 4, 4, 1)
    assert torch.equal(poses.cpu(), P["poses0"].float())
    assert (patches[0, :, 2].double().cpu() - xo[:, 2]).abs().max().item() <= 1e-5
    poses, patches = _run_ba(P, 1, 4, 0)
    assert torch.equal(poses.cpu(), P["poses0"].float()) and torch.equal(patches.cpu(), P["patches0"].float())
    # DEVO passes views of big ring buffers (devo.py:151-177): poses [1,4096,7], patches [1,4096*M,3,3,3]
    big_p = torch.zeros(1, 64, 7, device="cuda")
    big_p[..., 6] = 1
    big_x = torch.zeros(1, 64 * 16, 3, 3, 3, device="cuda")
    big_p[0, :4] = P["poses0"][0].float().cuda()
    big_x[0, :64] = P["patches0"][0].float().cuda()
    intr = torch.zeros(1, 64, 4, device="cuda")
    intr[0, :] = P["intrinsics"][0, 0].float().cuda()
    fastba.BA(big_p, big_x, intr, P["targets"].float().cuda(), P["weights"].float().cuda(),
              torch.tensor([1e-4], device="cuda"), P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda(), 1, 4, 2)
    p2, x2 = _run_ba(P, 1, 4, 2)
    assert torch.equal(big_p[0, :4], p2[0]) and torch.equal(big_x[0, :64], x2[0])      # deterministic, layout independent
    assert torch.equal(big_p[0, 4:, :6], torch.zeros(60, 6, device="cuda"))


def test_ba_masked_edges_and_failure_status():
    from devo_b200 import cuda_ba
    P = ba_problem(n_frames=3, patches_per_frame=8, seed=9, init="perturbed")
    # targets far away (> 128 px) mask every edge: nothing may move except the depth regulariser (dZ = 0)
    far = dict(P)
    far["targets"] = P["targets"] + 1000.0
    poses, patches = _run_ba(far, 1, 3, 2)
    assert torch.allclose(poses.cpu(), P["poses0"].float(), atol=1e-6)
    # a NaN weight makes the Schur system non-finite => status = iteration+1 (no update applied), and
    # STRICT raises like torch::linalg::cholesky does in the reference (caught by devo.py:336-340)
    badw = P["weights"].clone()
    badw[0, 3, 0] = float("nan")
    p0, x0 = P["poses0"].float().cuda(), P["patches0"].float().cuda()
    st = cuda_ba.forward_async(p0, x0, P["intrinsics"].float().cuda(), P["targets"].float().cuda(), badw.float().cuda(),
                               torch.tensor([1e-4], device="cuda"), P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda(), 1, 3, 2)
    assert int(st.item()) == 1
    assert torch.equal(p0.cpu(), P["poses0"].float()) and torch.equal(x0.cpu(), P["patches0"].float())
    bad = dict(P)
    bad["weights"] = badw
    with pytest.raises(RuntimeError):
        _run_ba(bad, 1, 3, 2)
    # a NaN *pose* only masks its edges here (mask comparisons are false), where the reference turns
    # 0*NaN into a failed factorisation: documented deviation
    badp = P["poses0"].clone()
    badp[0, 2, 0] = float("nan")
    poses, patches = _run_ba(P, 1, 3, 1, poses=badp)
    assert torch.isfinite(poses[0, :2]).all()


def test_ba_prepared_form_equals_planned_form_and_folds_the_status():
    """devo_ba_prepare + devo_ba_forward_prepared (the memsets of a call issued ahead, the status OR-ed into a sticky word
    by the last launch): same poses / depths bit for bit as devo_ba_forward_planned; a failed call leaves its status in the
    sticky word, a good one leaves it alone; the workspace can be reused call after call."""
    from devo_b200 import _lib, cuda_ba
    P = ba_problem(n_frames=5, patches_per_frame=16, seed=19, init="perturbed", noise=0.3)
    ii, jj, kk = P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda()
    plan = cuda_ba.GraphPlan(kk, jj, 5 * 16, 5)
    args = (P["intrinsics"].float().cuda(), P["targets"].float().cuda(), P["weights"].float().cuda(), torch.tensor([1e-4], device="cuda"),
            ii, jj, kk, 1, 5, 2)
    p0, x0 = P["poses0"].float().cuda(), P["patches0"].float().cuda()
    st0 = cuda_ba.forward_async(p0, x0, *args, plan=plan)
    ws = torch.empty(_lib.lib().devo_ba_workspace(ii.numel(), 4), dtype=torch.uint8, device="cuda").fill_(0xAB)     # garbage on purpose
    status = torch.full((1,), 77, dtype=torch.int32, device="cuda")
    sticky = torch.zeros(1, dtype=torch.int32, device="cuda")
    for rep in range(3):
        p1, x1 = P["poses0"].float().cuda(), P["patches0"].float().cuda()
        cuda_ba.prepare(ii.numel(), 4, status, ws)
        cuda_ba.forward_async(p1, x1, *args, status=status, plan=plan, workspace=ws, prepared=True, status_or=sticky)
        assert int(status.item()) == 0 == int(st0.item()) and int(sticky.item()) == 0
        assert torch.equal(p1, p0) and torch.equal(x1, x0)
    badw = P["weights"].float().clone()
    badw[0, 3, 0] = float("nan")
    bargs = (args[0], args[1], badw.cuda()) + args[3:]
    p2, x2 = P["poses0"].float().cuda(), P["patches0"].float().cuda()
    cuda_ba.prepare(ii.numel(), 4, status, ws)
    cuda_ba.forward_async(p2, x2, *bargs, status=status, plan=plan, workspace=ws, prepared=True, status_or=sticky)
    assert int(status.item()) == 1 and int(sticky.item()) == 1
    assert torch.equal(p2.cpu(), P["poses0"].float()) and torch.equal(x2.cpu(), P["patches0"].float())
    # ... and the next good call on the same workspace works again, the sticky word keeps the failure
    p3, x3 = P["poses0"].float().cuda(), P["patches0"].float().cuda()
    cuda_ba.prepare(ii.numel(), 4, status, ws)
    cuda_ba.forward_async(p3, x3, *args, status=status, plan=plan, workspace=ws, prepared=True, status_or=sticky)
    assert int(status.item()) == 0 and int(sticky.item()) == 1 and torch.equal(p3, p0) and torch.equal(x3, x0)


@pytest.mark.parametrize("E,nk,nj", [(1, 1, 1), (7, 2, 3), (300, 20, 6), (5000, 300, 8), (12000, 500, 12), (40000, 2000, 22)])
def test_neighbors_bit_exact(E, nk, nj):
    from devo_b200 import fastba
    rng = np.random.RandomState(E)
    kk = torch.from_numpy(rng.randint(0, nk, E))
    jj = torch.from_numpy(rng.randint(0, nj, E))
    ix, jx = fastba.neighbors(kk.cuda(), jj.cuda())
    rx, ry = onb.neighbors(kk, jj)
    assert ix.dtype == torch.int64 and ix.is_cuda
    assert torch.equal(ix.cpu(), rx) and torch.equal(jx.cpu(), ry)


def test_graph_plan_key_bounds_are_only_a_hint():
    """GraphPlan(ka, kb, max_ka, max_kb): correct bounds fix the radix width on the host; WRONG (too small) bounds must not
    change the result -- the kernel notices and measures the keys itself"""
    from devo_b200 import cuda_ba
    g = torch.Generator().manual_seed(0)
    ka = torch.randint(0, 700, (5000,), generator=g).cuda()
    kb = torch.randint(0, 9, (5000,), generator=g).cuda()
    ref = cuda_ba.GraphPlan(ka, kb, -1, -1)
    for bounds in ((700, 9), (1024, 16), (8, 2), (1, 1)):
        p = cuda_ba.GraphPlan(ka, kb, *bounds)
        assert int(p.ngroups) == int(ref.ngroups)
        n = int(ref.ngroups)
        assert torch.equal(p.perm, ref.perm) and torch.equal(p.gid, ref.gid)
        assert torch.equal(p.gstart[:n + 1], ref.gstart[:n + 1]) and torch.equal(p.gkey[:n], ref.gkey[:n])
        assert torch.equal(p.ix, ref.ix) and torch.equal(p.jx, ref.jx)


def test_graph_plan_equals_torch_unique():
    from devo_b200 import cuda_ba
    for E, nk in [(50, 7), (6144, 768), (20000, 1500)]:
        rng = np.random.RandomState(E)
        ka = torch.from_numpy(rng.randint(0, nk, E) * 3 + 5)
        kb = torch.from_numpy(rng.randint(0, 9, E))
        pl = cuda_ba.GraphPlan(ka.cuda(), kb.cuda())
        uq, inv = torch.unique(ka, sorted=True, return_inverse=True)
        G = int(pl.ngroups.item())
        assert G == uq.numel()
        assert torch.equal(pl.gkey[:G].cpu(), uq) and torch.equal(pl.gid.cpu().long(), inv)
        order = np.lexsort((np.arange(E), kb.numpy(), ka.numpy()))
        assert torch.equal(pl.perm.cpu().long(), torch.from_numpy(order))
        counts = torch.bincount(inv, minlength=G)
        assert torch.equal(pl.gstart[:G + 1].cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    ii, jj, kk = fully_connected_graph(8, 96)
    ix, jx = cuda_ba.neighbors(kk.cuda(), jj.cuda())
    e = torch.arange(6144)
    assert torch.equal(ix.cpu(), torch.where(e % 8 == 0, -1, e - 1)) and torch.equal(jx.cpu(), torch.where(e % 8 == 7, -1, e + 1))


def test_reproject_and_transform_vs_oracle():
    from devo_b200 import fastba, projective_ops as pops, lietorch as lt
    P = ba_problem(n_frames=5, patches_per_frame=20, seed=21, init="perturbed")
    f32 = lambda t: t.float().cuda()
    ii, jj, kk = P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda()
    rp = fastba.reproject(f32(P["poses0"]), f32(P["patches0"]), f32(P["intrinsics"]), ii, jj, kk)
    ro = ofba.reproject(P["poses0"].float(), P["patches0"][0].float(), P["intrinsics"].float(), P["ii"], P["jj"], P["kk"])
    assert rp.shape == (1, ii.numel(), 2, 3, 3)
    assert (rp.double().cpu() - ro).abs().max().item() <= 1e-3 * 1.0      # pixels, fp32 vs fp64 on O(100) px values
    G = lt.SE3(f32(P["poses0"]))
    a = (f32(P["patches0"]), f32(P["intrinsics"]), ii, jj, kk)
    od = lambda t: t.float().double()
    oa = (od(P["poses0"]), od(P["patches0"]), od(P["intrinsics"]), P["ii"], P["jj"], P["kk"])
    c, v, (Ji, Jj, Jz) = pops.transform(G, *a, jacobian=True)          # fused kernel (no grad)
    co, vo, (Jio, Jjo, Jzo) = opops.transform(*oa, jacobian=True)
    assert c.shape == co.shape and (c.double().cpu() - co).abs().max().item() <= 1e-3
    assert torch.equal(v.cpu().double(), vo)
    for x, y in ((Ji, Jio), (Jj, Jjo), (Jz, Jzo)):
        assert x.shape == y.shape
        assert ((x.double().cpu() - y).abs().max() / y.abs().max()).item() <= 1e-5
    ct = pops.transform(G, *a, tonly=True)
    assert (ct.double().cpu() - opops.transform(*oa, tonly=True)).abs().max().item() <= 1e-3
    # composed (autograd) path gives the same numbers as the fused kernel
    pr = a[0].clone().requires_grad_(True)
    c2, v2, (Ji2, Jj2, Jz2) = pops.transform(G, pr, *a[1:], jacobian=True)
    assert (c2 - c).abs().max().item() <= 2e-3 and torch.equal(v2, v)
    assert ((Ji2 - Ji).abs().max() / Ji.abs().max()).item() <= 1e-4 and ((Jz2 - Jz).abs().max() / Jz.abs().max()).item() <= 1e-4
    fm = pops.flow_mag(G, *a, beta=0.5)
    assert (fm.double().cpu() - opops.flow_mag(*oa, beta=0.5)).abs().max().item() <= 2e-3


def test_differentiable_ba_vs_oracle_and_golden():
    """devo_b200.ba.BA (training path, fp64 on CUDA) == reference ba.py golden vectors (config 1)"""
    import os
    from devo_b200 import ba as pba, lietorch as lt
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ba_config1.pt"))
    P = g["problem"]
    Gs, X = lt.SE3(P["poses0"].cuda()), P["patches0"].cuda()
    cu = lambda t: t.cuda()
    for it, (pg, dg) in enumerate(g["traj"][:4]):
        Gs, X = pba.BA(Gs, X, cu(P["intrinsics"]), cu(P["targets"]), cu(P["weights"]), 1e-4, cu(P["ii"]), cu(P["jj"]),
                       cu(P["kk"]), P["bounds"], ep=10.0, fixedp=1)
        assert torch.allclose(Gs.data.cpu(), pg, atol=1e-8), (it, (Gs.data.cpu() - pg).abs().max())
        assert torch.allclose(X[:, :, 2, 1, 1].cpu(), dg, atol=1e-8)
    g2 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ba_4x12.pt"))
    P = g2["problem"]
    a = (cu(P["intrinsics"]), cu(P["targets"]), cu(P["weights"]), 1e-4, cu(P["ii"]), cu(P["jj"]), cu(P["kk"]))
    G1, X1 = pba.BA(lt.SE3(cu(P["poses0"])), cu(P["patches0"]), *a, [20, 20, 140, 100], ep=10.0, fixedp=1)
    assert torch.allclose(G1.data.cpu(), g2["poses_a"], atol=1e-8) and torch.allclose(X1[:, :, 2, 1, 1].cpu(), g2["depth_a"], atol=1e-8)
    G2, X2 = pba.BA(lt.SE3(cu(P["poses0"])), cu(P["patches0"]), *a, P["bounds"], ep=100.0, fixedp=2, structure_only=True)
    assert torch.allclose(G2.data.cpu(), g2["poses_b"], atol=1e-10) and torch.allclose(X2[:, :, 2, 1, 1].cpu(), g2["depth_b"], atol=1e-8)
    # gradients flow to weights/targets/patches (training uses them)
    w = cu(P["weights"]).clone().requires_grad_(True)
    G3, X3 = pba.BA(lt.SE3(cu(P["poses0"])), cu(P["patches0"]), a[0], a[1], w, *a[3:], P["bounds"], ep=10.0, fixedp=1)
    (G3.data.sum() + X3.sum()).backward()
    assert torch.isfinite(w.grad).all() and w.grad.abs().sum() > 0


def test_segment_softmax_sum_and_update_planned_equals_reference_semantics():
    from devo_b200 import cuda_ba
    from devo_b200.update import Update
    torch.manual_seed(0)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(4, 12)]
    E = ii.numel()
    perm = torch.randperm(E, device="cuda")          # unsorted edge order
    ii, jj, kk = ii[perm], jj[perm], kk[perm]
    up = Update(3).cuda().eval()
    net = torch.randn(1, E, 384, device="cuda")
    inp = torch.randn(1, E, 384, device="cuda")
    corr = torch.randn(1, E, 882, device="cuda")
    with torch.no_grad():
        n1, (d1, w1, _) = up(net, inp, corr, None, ii, jj, kk)
        pk = cuda_ba.GraphPlan(kk, jj, 48, 4)
        pij = cuda_ba.GraphPlan(ii * 12345 + jj, kk, -1, -1, want_neighbors=False)
        n2, (d2, w2, _) = up.forward_planned(net, inp, corr, pk, pij, 48, 16)
    assert torch.allclose(n1, n2, atol=2e-4, rtol=1e-4) and torch.allclose(d1, d2, atol=2e-4) and torch.allclose(w1, w2, atol=2e-4)


def test_frozen_cast_equals_autocast():
    """the engine's cached-weight GRU path reproduces torch.autocast numerics"""
    from devo_b200 import cuda_ba
    from devo_b200.update import Update, FrozenCast
    torch.manual_seed(1)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(4, 12)]
    E = ii.numel()
    up = Update(3).cuda().eval()
    net = torch.randn(1, E, 384, device="cuda").half()
    inp = torch.randn(1, E, 384, device="cuda").half()
    corr = torch.randn(1, E, 882, device="cuda").half()
    pk = cuda_ba.GraphPlan(kk, jj, 48, 4)
    pij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(kk), -1, 1, want_neighbors=False)
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.float16):
            n1, (d1, w1, _) = up.forward_planned(net, inp, corr, pk, pij, 48, 16)
        n2, (d2, w2, _) = up.forward_planned(net, inp, corr, pk, pij, 48, 16, FrozenCast(torch.float16))
    assert n1.dtype == n2.dtype and d1.dtype == d2.dtype and w1.dtype == w2.dtype
    assert torch.allclose(n1, n2, atol=1e-3, rtol=1e-3) and torch.allclose(d1.float(), d2.float(), atol=1e-3)
    assert torch.allclose(w1.float(), w2.float(), atol=1e-3)


def test_ba_with_shared_plan_is_bitwise_identical():
    from devo_b200 import cuda_ba
    P = ba_problem(n_frames=5, patches_per_frame=20, seed=4, init="perturbed")
    a = lambda: (P["poses0"].float().cuda().contiguous(), P["patches0"].float().cuda().contiguous())
    rest = (P["intrinsics"].float().cuda(), P["targets"].float().cuda(), P["weights"].float().cuda(),
            torch.tensor([1e-4], device="cuda"), P["ii"].cuda(), P["jj"].cuda(), P["kk"].cuda(), 1, 5, 2)
    p1, x1 = a()
    cuda_ba.forward_async(p1, x1, *rest)
    p2, x2 = a()
    plan = cuda_ba.GraphPlan(P["kk"].cuda(), P["jj"].cuda(), 100, 5)
    st = cuda_ba.forward_async(p2, x2, *rest, plan=plan)
    assert int(st.item()) == 0 and torch.equal(p1, p2) and torch.equal(x1, x2)


def test_fused_gru_equals_planned_path():
    """forward_fused (hand-written glue kernels, ReLU in GEMM epilogues) == forward_planned (ATen ops)"""
    from devo_b200 import cuda_ba
    from devo_b200.update import Update, FrozenCast
    torch.manual_seed(2)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(5, 16)]
    E = ii.numel()
    perm = torch.randperm(E, device="cuda")
    ii, jj, kk = ii[perm], jj[perm], kk[perm]
    up = Update(3).cuda().eval()
    net = (0.3 * torch.randn(1, E, 384, device="cuda")).half()
    inp = (0.3 * torch.randn(1, E, 384, device="cuda")).half()
    corr = torch.randn(1, E, 882, device="cuda").half()
    pk = cuda_ba.GraphPlan(kk, jj, 80, 5)
    pij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(kk), -1, 1, want_neighbors=False)
    fc = FrozenCast(torch.float16)
    out16 = torch.empty_like(net)
    with torch.no_grad():
        n1, (d1, w1, _) = up.forward_planned(net, inp, corr, pk, pij, 80, 25, fc)
        n2, (d2, w2, _) = up.forward_fused(net, inp, corr, pk, pij, 80, 25, fc, net_out=out16)
    assert n1.dtype == n2.dtype == torch.float32 and d2.dtype == torch.float16 and w2.dtype == torch.float16
    assert torch.allclose(n1, n2, atol=2e-2, rtol=2e-2), (n1 - n2).abs().max()
    assert (n1 - n2).abs().mean() < 2e-3
    assert torch.allclose(d1.float(), d2.float(), atol=1e-2) and torch.allclose(w1.float(), w2.float(), atol=5e-3)
    assert torch.equal(out16, n2.half())


def test_s22_steady_state_shape():
    """DEVO steady state ("S22"): 22 source frames, ~45k edges, BA window of 10 poses -- exercises the
    multi-kernel graph-plan path (E > 16384), multi-batch accumulate CTAs and a 60x60 solve."""
    from devo_b200 import cuda_ba, fastba
    nf, m = 22, 96
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=22, init="perturbed", noise=0.3)
    # keep only edges within a 12-frame lifetime window, like PATCH_LIFETIME does
    keep = (P["ii"] - P["jj"]).abs() <= 12
    for k in ("ii", "jj", "kk"):
        P[k] = P[k][keep]
    P["targets"] = P["targets"][:, keep]
    P["weights"] = P["weights"][:, keep]
    E = P["ii"].numel()
    assert E > 16384
    ix, jx = fastba.neighbors(P["kk"].cuda(), P["jj"].cuda())
    rx, ry = onb.neighbors(P["kk"], P["jj"])
    assert torch.equal(ix.cpu(), rx) and torch.equal(jx.cpu(), ry)
    t0 = nf - 10
    poses, patches = _run_ba(P, t0, nf, 2)
    po, xo, st = _oracle_ba(P, t0, nf, 2)
    assert st == 0
    assert (poses[0].double().cpu() - po).abs().max().item() <= 1e-5
    assert (patches[0, :, 2].double().cpu() - xo[:, 2]).abs().max().item() <= 1e-5
    assert torch.equal(poses[0, :t0].cpu(), P["poses0"][0, :t0].float())


# ---- edge-sharded BA: one frame graph split over ranks by owning patch ------------------------------
def _sharded_inputs(P, rank, world, dev):
    from devo_b200 import dist as d
    Np = P["patches0"].shape[1]
    sel = d.shard_edges_by_patch(P["kk"], Np, rank, world)
    f = lambda t: t.float().to(dev).contiguous()
    return dict(poses=f(P["poses0"]), patches=f(P["patches0"]), intr=f(P["intrinsics"]),
                target=f(P["targets"][:, sel]), weight=f(P["weights"][:, sel]),
                ii=P["ii"][sel].to(dev), jj=P["jj"][sel].to(dev), kk=P["kk"][sel].to(dev), sel=sel)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("nf,m,t0,iters", [(8, 96, 1, 2), (5, 7, 2, 3)])
def test_sharded_ba_simulated_ranks_match_single_gpu(world, nf, m, t0, iters):
    """`world` ranks simulated on one GPU: each owns the edges of its patch block, the all-reduce is a plain sum of the
    per-rank systems.  Result must equal the single-GPU fastba.BA (same arithmetic, fp64 partial sums regrouped) and
    the poses must stay bitwise replicated across ranks."""
    from devo_b200 import cuda_ba, dist as d
    P = ba_problem(n_frames=nf, patches_per_frame=m, seed=77 + nf, init="perturbed", noise=0.3)
    ref_poses, ref_patches = _run_ba(P, t0, nf, iters)
    lm = torch.tensor([1e-4], device="cuda")
    ranks = []
    for r in range(world):
        I = _sharded_inputs(P, r, world, "cuda")
        I["ba"] = cuda_ba.ShardedBA(I["poses"], I["patches"], I["intr"], I["target"], I["weight"], lm, I["ii"], I["jj"], I["kk"], t0, nf)
        ranks.append(I)
    for itr in range(iters):
        total = sum(I["ba"].accumulate(itr).clone() for I in ranks)
        for I in ranks:
            I["ba"].system.copy_(total)
            I["ba"].solve(itr)
    for I in ranks:
        I["ba"].finish(iters)
        assert int(I["ba"].status.item()) == 0
    Np = nf * m
    for r, I in enumerate(ranks):
        assert torch.equal(I["poses"], ranks[0]["poses"])                       # replicated, bitwise
        assert (I["poses"] - ref_poses).abs().max().item() <= 1e-6
        lo, hi = d.patch_range(Np, r, world)
        assert (I["patches"][0, lo:hi] - ref_patches[0, lo:hi]).abs().max().item() <= 1e-6 if hi > lo else True
        other = torch.ones(Np, dtype=torch.bool)
        other[lo:hi] = False
        assert torch.equal(I["patches"][0, other].cpu(), P["patches0"][0, other].float())   # foreign patches untouched


def test_sharded_ba_failure_is_seen_by_every_rank():
    """a rank whose accumulate fails (here: NaN weights -> non-PD system after the sum) stops all ranks identically"""
    from devo_b200 import cuda_ba
    P = ba_problem(n_frames=4, patches_per_frame=8, seed=5, init="perturbed")
    lm = torch.tensor([1e-4], device="cuda")
    ranks = []
    for r in range(2):
        I = _sharded_inputs(P, r, 2, "cuda")
        if r == 1:
            I["weight"].fill_(float("nan"))
        I["ba"] = cuda_ba.ShardedBA(I["poses"], I["patches"], I["intr"], I["target"], I["weight"], lm, I["ii"], I["jj"], I["kk"], 1, 4)
        ranks.append(I)
    total = sum(I["ba"].accumulate(0).clone() for I in ranks)
    for I in ranks:
        I["ba"].system.copy_(total)
        I["ba"].solve(0)
        assert int(I["ba"].status.item()) != 0
        assert torch.equal(I["poses"].cpu(), P["poses0"].float())               # nothing applied


def _nccl_worker(rank, world, port, q, peer=False):
    import os
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from devo_b200 import cuda_ba, dist as d
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    P = ba_problem(n_frames=8, patches_per_frame=96, seed=85, init="perturbed", noise=0.3)
    I = _sharded_inputs(P, rank, world, torch.device("cuda", rank))
    st = cuda_ba.forward_sharded(I["poses"], I["patches"], I["intr"], I["target"], I["weight"], torch.tensor([1e-4], device=I["poses"].device),
                                 I["ii"], I["jj"], I["kk"], 1, 8, 2, peer=peer)
    d.gather_patch_depths(I["patches"], 8 * 96)
    torch.cuda.synchronize()
    q.put((rank, int(st.item()), I["poses"].cpu(), I["patches"].cpu()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ba_two_gpus_nccl():
    """the real thing: 2 processes, 2 GPUs, NCCL all-reduce of [S|y] per iteration (skipped on a 1-GPU box)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    P = ba_problem(n_frames=8, patches_per_frame=96, seed=85, init="perturbed", noise=0.3)
    ref_poses, ref_patches = _run_ba(P, 1, 8, 2)
    assert out[0][1] == 0 and out[1][1] == 0
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][3], out[1][3])     # replicas agree bitwise
    assert (out[0][2] - ref_poses.cpu()).abs().max().item() <= 1e-6
    assert (out[0][3] - ref_patches.cpu()).abs().max().item() <= 1e-6


def test_sharded_ba_two_gpus_peer_memory():
    """the all-reduce fused into the solve kernel over NVLink peer memory (torch symmetric memory; no NCCL call on the
    data path): same result as the NCCL form and as a single GPU, replicas bitwise identical"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 250)
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    P = ba_problem(n_frames=8, patches_per_frame=96, seed=85, init="perturbed", noise=0.3)
    ref_poses, ref_patches = _run_ba(P, 1, 8, 2)
    assert out[0][1] == 0 and out[1][1] == 0
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][3], out[1][3])     # replicas agree bitwise
    assert (out[0][2] - ref_poses.cpu()).abs().max().item() <= 1e-6
    assert (out[0][3] - ref_patches.cpu()).abs().max().item() <= 1e-6
