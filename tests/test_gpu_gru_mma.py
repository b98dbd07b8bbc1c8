"""GPU: the fused tcgen05 update operator (csrc/gru_mma.cu, devo_gru_update) against (a) the cuBLAS + glue-kernel path
(forward_fused, identical rounding points) and (b) the reference-shaped module forward under torch.autocast
(devo/enet.py:80-99).  Tolerances are half-precision noise: the two paths only differ in the summation order
inside the dot products (fp32 accumulation in both)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(nf, m, seed, dt=torch.float16):
    from devo_b200 import cuda_ba
    from devo_b200.update import Update
    from problems import fully_connected_graph
    torch.manual_seed(seed)
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(nf, m)]
    if seed % 2:      # ragged graph: drop a random fifth of the edges (keyframe removal leaves such graphs)
        keep = torch.rand(ii.numel(), device="cuda") > 0.2
        ii, jj, kk = ii[keep], jj[keep], kk[keep]
    E, Np = ii.numel(), nf * m
    up = Update(3).cuda().eval()
    with torch.no_grad():
        for p in up.parameters():      # non-trivial LayerNorm affine + biases
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    net = (0.5 * torch.randn(1, E, 384, device="cuda")).to(dt)
    imap = (0.25 * torch.randn(1, Np, 384, device="cuda")).to(dt)
    corr = torch.zeros(E, 896, device="cuda", dtype=dt)
    corr[:, :882] = (2.0 * torch.randn(E, 882, device="cuda")).to(dt)
    plan_kk = cuda_ba.GraphPlan(kk, jj, Np, nf)
    plan_ij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(ii), -1, 1, want_neighbors=False)
    return up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, nf * nf


@pytest.mark.parametrize("nf,m,seed", [(8, 96, 0), (4, 24, 1), (3, 5, 2), (5, 31, 3), (22, 96, 7)])   # last: S22, ragged, several waves of clusters
def test_gru_mma_matches_cublas_path(nf, m, seed):
    from devo_b200.update import FrozenCast, PackedUpdateWeights
    up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, pairs = _problem(nf, m, seed)
    fc = FrozenCast(torch.float16)
    with torch.no_grad():
        ref_net, (ref_d, ref_w, _) = up.forward_fused(net, imap[:, kk], corr.view(1, -1, 896), plan_kk, plan_ij, Np, pairs, fc)
        packed = PackedUpdateWeights(up, torch.float16, 896)
        out_net, (d, w, _) = up.forward_mma(net, imap, kk, corr, plan_kk, plan_ij, Np, pairs, packed)
    torch.cuda.synchronize()
    assert torch.isfinite(out_net).all() and torch.isfinite(d).all() and torch.isfinite(w).all()
    scale = ref_net.abs().max().item()
    assert (out_net.float() - ref_net).abs().max().item() <= 2e-2 * max(scale, 1.0)
    assert (out_net.float() - ref_net).abs().mean().item() <= 1e-3 * max(scale, 1.0)
    assert (d.float() - ref_d.float()).abs().max().item() <= 2e-2
    assert (w.float() - ref_w.float()).abs().max().item() <= 1e-2


def test_gru_mma_matches_module_forward_under_autocast():
    from devo_b200.update import PackedUpdateWeights
    up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, pairs = _problem(4, 24, 4)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref_net, (ref_d, ref_w, _) = up(net, imap[:, kk], corr[:, :882].reshape(1, -1, 882), None, ii, jj, kk)
    with torch.no_grad():
        out_net, (d, w, _) = up.forward_mma(net, imap, kk, corr, plan_kk, plan_ij, Np, pairs, PackedUpdateWeights(up, torch.float16, 896))
    assert (out_net.float() - ref_net.float()).abs().max().item() <= 3e-2 * max(ref_net.abs().max().item(), 1.0)
    assert (d.float() - ref_d.float()).abs().max().item() <= 2e-2
    assert (w.float() - ref_w.float()).abs().max().item() <= 1e-2


def test_gru_mma_bf16_and_persistent_state():
    """bf16 weights / activations; the hidden state lives in a GruState (tile-layout float32) that the kernel updates in
    place, and the optional half copy of the new state (net_out) is written by the same launch"""
    from devo_b200.update import FrozenCast, GruState, PackedUpdateWeights
    up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, pairs = _problem(4, 24, 6, dt=torch.bfloat16)
    fc = FrozenCast(torch.bfloat16)
    with torch.no_grad():
        ref_net, (ref_d, ref_w, _) = up.forward_fused(net, imap[:, kk], corr.view(1, -1, 896), plan_kk, plan_ij, Np, pairs, fc)
        state = GruState(net.shape[1], "cuda")
        copy16 = torch.empty_like(net)
        ret, (d, w, _) = up.forward_mma(net, imap, kk, corr, plan_kk, plan_ij, Np, pairs,
                                        PackedUpdateWeights(up, torch.bfloat16, 896), net_out=copy16, state=state)
    assert ret is state
    assert (state.get() - ref_net).abs().max().item() <= 1e-1 * max(ref_net.abs().max().item(), 1.0)
    assert torch.equal(copy16.float(), state.get().to(torch.bfloat16).float())
    assert (d.float() - ref_d.float()).abs().max().item() <= 1e-1


def test_gru_state_pack_unpack_and_gather():
    """GruState: row-major <-> tile layout round trip is exact; gather implements net[:, ~m] and zero rows for new edges"""
    from devo_b200.update import GruState
    for E in (1, 63, 64, 129, 6144, 777):
        x = torch.randn(1, E, 384, device="cuda")
        st = GruState(E, "cuda").set(x)
        assert torch.equal(st.get(), x)
        idx = torch.randint(-1, E, (E + 5,), device="cuda")
        got = st.gather(idx).get()[0]
        want = torch.where((idx >= 0)[:, None], x[0][idx.clamp_min(0)], torch.zeros(1, device="cuda"))
        assert torch.equal(got, want)


def test_gru_mma_is_deterministic():
    """pair hand-offs (A-operand double buffer, TMEM sets, staged TMA stores) must not race: repeated runs are bit-identical"""
    from devo_b200.update import PackedUpdateWeights
    up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, pairs = _problem(8, 96, 0)
    packed = PackedUpdateWeights(up, torch.float16, 896)
    outs = []
    with torch.no_grad():
        for _ in range(6):
            n, (d, w, _) = up.forward_mma(net, imap, kk, corr, plan_kk, plan_ij, Np, pairs, packed)
            outs.append((n.clone(), d.clone(), w.clone()))
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])


@pytest.mark.parametrize("nf,m,state_dtype", [(8, 96, torch.float32), (8, 96, torch.float16), (4, 40, torch.float32), (16, 12, torch.float32)])
def test_gru_mma_tile_local_program_matches_the_classic_form(nf, m, state_dtype):
    """devo_gru_io_t.tile_local: when every neighbour link stays inside its 64-edge tile (the patch-major all-pairs graph
    with 8, 4 or 16 edges per patch) the first four programs and the patch-wise segment reduction run as one launch: the
    epilogues permute rows inside the CTA's own A tile instead of gathering them across kernel boundaries, and the
    SoftAgg over a patch's edges walks the patch's chain in shared memory.  The neighbour exchange is the same arithmetic
    (bit-identical on its own); the in-tile softmax adds the rows of a patch in chain order where the segment kernel adds
    two interleaved halves, so the group rows may differ in the last bit of a half: compared at half-precision noise."""
    from devo_b200.update import PackedUpdateWeights, tile_local_graph
    up, net, imap, corr, ii, jj, kk, plan_kk, plan_ij, Np, pairs = _problem(nf, m, 0)
    assert tile_local_graph(plan_kk)
    packed = PackedUpdateWeights(up, torch.float16, 896)
    coords = torch.randn(1, ii.numel(), 2, 3, 3, device="cuda")
    outs = []
    with torch.no_grad():
        for tl in (False, True):
            n0 = net.to(state_dtype)
            res = []
            for it in range(2):                      # second update: the float32 state path even when the first was half
                n_out, (d, w, (tg, wt)) = up.forward_mma(n0, imap, kk, corr, plan_kk, plan_ij, Np, pairs, packed, coords=coords, tile_local=tl)
                res += [n_out.clone(), d.clone(), w.clone(), tg.clone(), wt.clone()]
                n0 = n_out
            outs.append(res)
    torch.cuda.synchronize()
    names = ["net", "delta", "weight", "target", "weight32"] * 2
    for nm, a, b in zip(names, *outs):
        assert torch.isfinite(b).all()
        err = (a.float() - b.float()).abs().max().item()
        scale = max(a.float().abs().max().item(), 1.0)
        print("tile-local vs classic %-8s max |diff| %.3e (scale %.2f)" % (nm, err, scale))
        assert err <= 4e-3 * scale, (nm, err)
        assert (a.float() - b.float()).abs().mean().item() <= 2e-4 * scale


def test_tile_local_graph_detects_straddling_patches():
    from devo_b200 import cuda_ba
    from devo_b200.update import tile_local_graph
    from problems import fully_connected_graph
    for nf, ok in ((8, True), (4, True), (16, True), (5, False), (12, False), (22, False)):
        ii, jj, kk = [t.cuda() for t in fully_connected_graph(nf, 16)]
        assert tile_local_graph(cuda_ba.GraphPlan(kk, jj, nf * 16, nf)) == ok
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(8, 16)]
    perm = torch.randperm(ii.numel(), device="cuda")                       # the DEVO loop's append order is not patch-major
    assert not tile_local_graph(cuda_ba.GraphPlan(kk[perm], jj[perm], 8 * 16, 8))
