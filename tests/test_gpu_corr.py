"""GPU parity: correlation lookup / patch gather (through the C ABI) vs the CPU oracle.
Tolerances (BASELINE.json north_star): correlation <= 1e-4 relative for fp32 inputs;
patchify / index scatter bit-exact; half inputs: error vs the fp64 oracle is bounded by
half output rounding (our accumulation is fp32, the reference's is half)."""
import pytest
import torch

from oracle import corr as ocorr
from problems import corr_problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _rand_problem(B, Np, Nf, C, H, W, E, P, seed, dtype, spread=4.0):
    g = torch.Generator().manual_seed(seed)
    f1 = (torch.randn(B, Np, C, P, P, generator=g) / 4).to(dtype)
    f2 = (torch.randn(B, Nf, C, H, W, generator=g) / 4).to(dtype)
    cx = torch.rand(B, E, 1, 1, 1, generator=g) * (W + 8) - 4
    cy = torch.rand(B, E, 1, 1, 1, generator=g) * (H + 8) - 4
    coords = torch.cat([cx + torch.rand(B, E, 1, P, P, generator=g) * spread, cy + torch.rand(B, E, 1, P, P, generator=g) * spread], 2)
    ii = torch.randint(0, Np, (E,), generator=g)
    jj = torch.randint(0, Nf, (E,), generator=g)
    return f1, f2, coords.float(), ii, jj


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-12), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("R,P,C", [(3, 3, 128), (1, 3, 24), (2, 2, 8)])
def test_generic_forward_vs_oracle(dtype, tol, R, P, C):
    from devo_b200 import _lib
    f1, f2, coords, ii, jj = _rand_problem(2, 20, 3, C, 30, 40, 150, P, 7, dtype)
    ref = ocorr.corr_forward(f1, f2, coords, ii, jj, R)
    D1 = 2 * R + 1
    out = torch.empty(2, 150, D1, D1, P, P, dtype=dtype, device="cuda")
    a = [t.cuda().contiguous() for t in (f1, f2, coords, ii, jj)]
    _lib.check(_lib.lib().devo_corr_forward(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(),
                                            a[4].data_ptr(), out.data_ptr(), _lib.dtype_code(a[0]), 2, 20, 3, C, 30, 40,
                                            150, P, R, _lib.stream_ptr()), "corr_forward")
    assert _rel(out, ref) <= tol


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("C", [128, 64])
def test_fast_path_vs_oracle_and_generic(dtype, C):
    """cuda_corr.forward takes the TMA + tcgen05 path for (half, C in {64,128}, P=3, r=3, B=1)"""
    from devo_b200 import cuda_corr
    f1, f2, coords, ii, jj = _rand_problem(1, 40, 5, C, 30, 40, 700, 3, 11, dtype, spread=2.5)
    coords[0, :50] += 1000.0           # windows entirely out of bounds -> zeros
    coords[0, 50:80, :, 2, 2] += 7.0   # one patch pixel far away -> per-output direct path
    ref = ocorr.corr_forward(f1, f2, coords, ii, jj, 3)
    a = [t.cuda() for t in (f1, f2, coords, ii, jj)]
    assert cuda_corr._fast_eligible(a[0], a[1], a[2], 3)
    (out,) = cuda_corr.forward(*a, 3)
    assert out.shape == ref.shape and out.dtype == dtype
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert _rel(out, ref) <= tol, _rel(out, ref)
    assert out[0, :50].abs().max().item() == 0.0
    # generic path on the same inputs agrees to output rounding
    from devo_b200 import _lib
    gen = torch.empty_like(out)
    _lib.check(_lib.lib().devo_corr_forward(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(),
                                            a[4].data_ptr(), gen.data_ptr(), _lib.dtype_code(a[0]), 1, 40, 5, C, 30, 40,
                                            700, 3, 3, _lib.stream_ptr()), "corr_forward")
    assert _rel(out, gen) <= tol


@pytest.mark.parametrize("C", [128, 64])
def test_split_precision_fast_path_float32(C):
    """float32 features (what training feeds altcorr) take the tensor-core kernel in split precision: each value as two
    halves (22 significant bits), three passes, float accumulation.  Tolerance 1e-5 of the largest output against the fp64
    oracle and against the float32 SIMT kernel -- the bar VERDICT r1 set was 1e-4."""
    from devo_b200 import _lib, cuda_corr
    f1, f2, coords, ii, jj = _rand_problem(1, 40, 5, C, 30, 40, 700, 3, 13, torch.float32, spread=2.5)
    f2[0, :, :8] *= 1e-3               # small-magnitude channels: the low halves must not underflow to nothing
    coords[0, :50] += 1000.0           # windows entirely out of bounds -> zeros
    coords[0, 50:80, :, 2, 2] += 7.0   # one patch pixel far away -> per-output direct path
    ref = ocorr.corr_forward(f1.double(), f2.double(), coords, ii, jj, 3)
    a = [t.cuda() for t in (f1, f2, coords, ii, jj)]
    assert cuda_corr._split_eligible(a[0], a[1], a[2], 3)
    (out,) = cuda_corr.forward(*a, 3)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel(out, ref) <= 1e-5, _rel(out, ref)
    assert out[0, :50].abs().max().item() == 0.0
    gen = torch.empty_like(out)
    _lib.check(_lib.lib().devo_corr_forward(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(),
                                            a[4].data_ptr(), gen.data_ptr(), _lib.dtype_code(a[0]), 1, 40, 5, C, 30, 40,
                                            700, 3, 3, _lib.stream_ptr()), "corr_forward")
    assert _rel(out, gen) <= 1e-5, _rel(out, gen)
    # the hi / lo buffers reproduce the float features to 2^-22 relative
    hi, lo = cuda_corr.pack_pixel_major_split(a[1][0], 1)
    back = (hi.float() + lo.float() / 2048.0).permute(0, 3, 1, 2)
    assert ((back - a[1][0]).abs() <= a[1][0].abs() * 2.0 ** -21 + 1e-7).all()


def test_split_precision_multilevel_matches_per_level_float32():
    """two pooled levels through lookup_fused_split == the float32 kernel per level, stacked like devo.py:217"""
    import torch.nn.functional as F
    from devo_b200 import cuda_corr
    f1, f2, coords, ii, jj = _rand_problem(1, 40, 5, 128, 32, 40, 300, 3, 17, torch.float32, spread=2.5)
    a = [t.cuda() for t in (f1, f2, coords, ii, jj)]
    g = cuda_corr.pack_gmap_split(a[0][0])
    lv = [cuda_corr.pack_pixel_major_split(a[1][0], s) for s in (1, 4)]
    out = cuda_corr.lookup_fused_split(g, lv, (1, 4), a[2][0], a[3], a[4])
    per = []
    for s in (1, 4):
        fm = a[1] if s == 1 else F.avg_pool2d(a[1][0], s, s)[None]
        gen = torch.empty(1, 300, 7, 7, 3, 3, dtype=torch.float32, device="cuda")
        from devo_b200 import _lib
        cs = (a[2] / s).contiguous()
        _lib.check(_lib.lib().devo_corr_forward(a[0].data_ptr(), fm.contiguous().data_ptr(), cs.data_ptr(), a[3].data_ptr(),
                                                a[4].data_ptr(), gen.data_ptr(), _lib.dtype_code(a[0]), 1, 40, 5, 128,
                                                fm.shape[3], fm.shape[4], 300, 3, 3, _lib.stream_ptr()), "corr_forward")
        per.append(gen)
    ref = torch.stack(per, -1).view(300, -1)
    assert _rel(out, ref) <= 1e-5, _rel(out, ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("pool,H,W,C", [(4, 120, 160, 128), (2, 30, 40, 64), (8, 32, 72, 128)])
def test_pyramid_pack2_is_bit_identical_to_two_packs(dtype, pool, H, W, C):
    """both levels of a [1, pool] pyramid from one read == pack(1) and pack(pool) == permute / F.avg_pool2d"""
    import torch.nn.functional as F
    from devo_b200 import cuda_corr
    g = torch.Generator().manual_seed(3)
    f = torch.randn(3, C, H, W, generator=g).to(dtype).cuda()
    o1 = torch.empty(3, H, W, C, dtype=dtype, device="cuda")
    op = torch.empty(3, H // pool, W // pool, C, dtype=dtype, device="cuda")
    assert cuda_corr.pack_pixel_major2(f, pool, o1, op)
    assert torch.equal(o1, cuda_corr.pack_pixel_major(f, 1)) and torch.equal(o1, f.permute(0, 2, 3, 1))
    assert torch.equal(op, cuda_corr.pack_pixel_major(f, pool))
    assert torch.equal(op, F.avg_pool2d(f, pool, pool).permute(0, 2, 3, 1))
    assert not cuda_corr.pack_pixel_major2(f[:, :, :, :W - 4].contiguous(), pool, o1, op)       # W % 8 != 0: declined, caller falls back


def test_fused_multilevel_layout_matches_stack():
    """lookup_fused over levels [1,4] == torch.stack([corr(l) for l], -1).view(1,E,-1) (devo.py:210-217)"""
    from devo_b200 import cuda_corr
    Pm = corr_problem(n_frames=3, patches_per_frame=20, C=128, H4=60, W4=80, seed=3)
    gm = cuda_corr.pack_gmap(Pm["gmap"][0].cuda())
    lv = [cuda_corr.pack_pixel_major(Pm["fmap"][0].cuda(), s) for s in (1, 4)]
    # the packer's pooled level equals avg_pool2d of the planar map (bit exact in half)
    for l, s in enumerate((1, 4)):
        assert torch.equal(lv[l].permute(0, 3, 1, 2).cpu(), Pm["pyramid"][l][0])
    coords = Pm["coords"].cuda()
    ii, jj = Pm["kk"].cuda(), Pm["jj"].cuda()
    out = cuda_corr.lookup_fused(gm, lv, (1, 4), coords[0], ii, jj)
    refs = [ocorr.corr_forward(Pm["gmap"], Pm["pyramid"][l], Pm["coords"] / s, Pm["kk"], Pm["jj"], 3) for l, s in enumerate((1, 4))]
    ref = torch.stack(refs, -1).reshape(1, Pm["kk"].numel(), -1)
    assert out.shape == (Pm["kk"].numel(), 882)
    assert _rel(out, ref[0]) <= 2e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 1e-11)])
def test_backward_vs_oracle(dtype, tol):
    from devo_b200 import cuda_corr
    f1, f2, coords, ii, jj = _rand_problem(1, 12, 3, 16, 20, 24, 60, 3, 5, dtype)
    g = torch.randn(1, 60, 7, 7, 3, 3)
    r1, r2 = ocorr.corr_backward(f1, f2, coords, ii, jj, g, 3)
    g1, g2 = cuda_corr.backward(f1.cuda(), f2.cuda(), coords.cuda(), ii.cuda(), jj.cuda(), g.cuda(), 3)
    assert _rel(g1, r1) <= tol and _rel(g2, r2) <= tol


@pytest.mark.parametrize("C,spread", [(64, 4.0), (128, 4.0), (128, 1.0), (64, 40.0)])
def test_backward_pixel_major_vs_oracle_and_generic(C, spread):
    """float32 at the training shape takes csrc/corr_bwd_pm.cu (pixel-major volumes, vector reductions); spread 40 puts the
    patch pixels so far apart that every pixel walks its own window (bounding box > 20 x 20)"""
    from devo_b200 import cuda_corr
    f1, f2, coords, ii, jj = _rand_problem(1, 12, 3, C, 20, 24, 60, 3, 11, torch.float32, spread=spread)
    g = torch.randn(1, 60, 7, 7, 3, 3)
    r1, r2 = ocorr.corr_backward(f1, f2, coords, ii, jj, g, 3)
    a = (f1.cuda(), f2.cuda(), coords.cuda(), ii.cuda(), jj.cuda(), g.cuda())
    assert cuda_corr._split_eligible(a[0], a[1], a[2], 3)
    g1, g2 = cuda_corr.backward(*a, 3)
    assert g1.shape == f1.shape and g2.shape == f2.shape and g1.is_contiguous() and g2.is_contiguous()
    assert _rel(g1, r1) <= 2e-5 and _rel(g2, r2) <= 2e-5
    cuda_corr._FORCE_GENERIC = True
    try:
        h1, h2 = cuda_corr.backward(*a, 3)
    finally:
        cuda_corr._FORCE_GENERIC = False
    assert _rel(g1, h1) <= 2e-5 and _rel(g2, h2) <= 2e-5
    # no edges, and an edge list that only reaches out-of-bounds windows
    e1, e2 = cuda_corr.backward(a[0], a[1], a[2][:, :0], a[3][:0], a[4][:0], a[5][:, :0], 3)
    assert float(e1.abs().max()) == 0 and float(e2.abs().max()) == 0
    z1, z2 = cuda_corr.backward(a[0], a[1], a[2] + 1000.0, a[3], a[4], a[5], 3)
    assert float(z1.abs().max()) == 0 and float(z2.abs().max()) == 0


def test_autograd_wrappers():
    from devo_b200 import altcorr
    f1, f2, coords, ii, jj = _rand_problem(1, 6, 2, 8, 12, 14, 20, 3, 9, torch.float32)
    a = f1.cuda().requires_grad_(True)
    b = f2.cuda().requires_grad_(True)
    out = altcorr.corr(a, b, coords.cuda(), ii.cuda(), jj.cuda(), 3, 1)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    r1, r2 = ocorr.corr_backward(f1, f2, coords, ii, jj, w.cpu(), 3)
    assert _rel(a.grad, r1) <= 2e-5 and _rel(b.grad, r2) <= 2e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("R", [0, 1, 3])
def test_patchify_bit_exact(dtype, R):
    from devo_b200 import cuda_corr, altcorr
    g = torch.Generator().manual_seed(R)
    net = torch.randn(2, 7, 13, 17, generator=g).to(dtype)
    coords = torch.cat([torch.rand(2, 30, 1, generator=g) * 21 - 2, torch.rand(2, 30, 1, generator=g) * 17 - 2], -1)
    coords[:, :10] = coords[:, :10].floor()
    ref = ocorr.patchify_forward(net, coords, R)
    (out,) = cuda_corr.patchify_forward(net.cuda(), coords.cuda(), R)
    assert torch.equal(out.cpu(), ref)                     # pure copy => bit exact
    if dtype in (torch.float32, torch.float64):
        gr = torch.randn_like(ref)
        (gb,) = cuda_corr.patchify_backward(net.cuda(), coords.cuda(), gr.cuda(), R)
        assert _rel(gb, ocorr.patchify_backward(net, coords, gr, R)) <= 1e-5
        pb = altcorr.patchify(net.cuda(), coords.cuda(), R)
        assert _rel(pb, ocorr.patchify(net, coords, R)) <= 1e-5


def test_empty_and_errors():
    from devo_b200 import cuda_corr
    f1 = torch.zeros(1, 4, 8, 3, 3, device="cuda")
    f2 = torch.zeros(1, 2, 8, 5, 5, device="cuda")
    e = torch.zeros(0, dtype=torch.long, device="cuda")
    (o,) = cuda_corr.forward(f1, f2, torch.zeros(1, 0, 2, 3, 3, device="cuda"), e, e, 3)
    assert o.shape == (1, 0, 7, 7, 3, 3)
    with pytest.raises(RuntimeError):
        cuda_corr.forward(f1.cpu(), f2.cpu(), torch.zeros(1, 0, 2, 3, 3), e.cpu(), e.cpu(), 3)   # no CPU fallback
    with pytest.raises(RuntimeError):
        cuda_corr.forward(f1, f2.half(), torch.zeros(1, 0, 2, 3, 3, device="cuda"), e, e, 3)


def test_full_size_properties_s8():
    """BASELINE.json config 2 at full size (E=6144, 160x120, C=128): size-independent properties --
    linearity in fmap1, zero for out-of-bounds windows, level interleave, determinism."""
    from devo_b200 import cuda_corr
    Pm = corr_problem(seed=1234)
    gm = cuda_corr.pack_gmap(Pm["gmap"][0].cuda())
    lv = [cuda_corr.pack_pixel_major(Pm["fmap"][0].cuda(), s) for s in (1, 4)]
    coords, ii, jj = Pm["coords"].cuda()[0], Pm["kk"].cuda(), Pm["jj"].cuda()
    out = cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj)
    out2 = cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj)
    assert torch.equal(out, out2) and torch.isfinite(out).all()
    outn = cuda_corr.lookup_fused(-gm, lv, (1, 4), coords, ii, jj)
    assert torch.equal(outn, -out)                                     # exact: sign flip commutes with rounding
    l0 = cuda_corr.lookup_fused(gm, lv[:1], (1,), coords, ii, jj)
    assert torch.equal(out.view(-1, 441, 2)[..., 0], l0)
    far = cuda_corr.lookup_fused(gm, lv, (1, 4), coords + 5000.0, ii, jj)
    assert far.abs().max().item() == 0.0
    # spot-check 64 random edges against the fp64 oracle
    sel = torch.randperm(ii.numel(), generator=torch.Generator().manual_seed(0))[:64]
    refs = [ocorr.corr_forward(Pm["gmap"], Pm["pyramid"][l], Pm["coords"][:, sel] / s, Pm["kk"][sel], Pm["jj"][sel], 3)
            for l, s in enumerate((1, 4))]
    ref = torch.stack(refs, -1).reshape(64, -1)
    assert _rel(out[sel.cuda()], ref) <= 2e-3


def test_backward_full_size_properties_s8_float32():
    """the pixel-major backward at full size (E = 6144, 160x120 and 40x30, C = 128, float32), through size-independent
    properties: the lookup is bilinear in (fmap1, fmap2), so <corr(f1, f2), g> = <f1, d/df1> = <f2, d/df2> (adjoint
    identity, forward in float64-accumulated reference on a sample; here against this library's float32 forward);
    linearity in the incoming gradient; exact zero for windows that lie outside the image"""
    from devo_b200 import cuda_corr
    Pm = corr_problem(seed=1234, dtype=torch.float32)
    f1 = Pm["gmap"].cuda()
    ii, jj = Pm["kk"].cuda(), Pm["jj"].cuda()
    E = ii.numel()
    gen = torch.Generator(device="cuda").manual_seed(5)
    for lvl, s in enumerate((1, 4)):
        f2 = Pm["pyramid"][lvl].cuda()
        coords = (Pm["coords"] / s).cuda()
        assert cuda_corr._split_eligible(f1, f2, coords, 3)
        g = torch.randn(1, E, 7, 7, 3, 3, device="cuda", generator=gen)
        h = torch.randn(1, E, 7, 7, 3, 3, device="cuda", generator=gen)
        (out,) = cuda_corr.forward(f1, f2, coords, ii, jj, 3)
        g1, g2 = cuda_corr.backward(f1, f2, coords, ii, jj, g, 3)
        assert torch.isfinite(g1).all() and torch.isfinite(g2).all()
        lhs = (out.double() * g.double()).sum().item()
        a1 = (f1.double() * g1.double()).sum().item()
        a2 = (f2.double() * g2.double()).sum().item()
        scale = (out.double() * g.double()).abs().sum().item()
        assert abs(lhs - a1) <= 1e-5 * scale and abs(lhs - a2) <= 1e-5 * scale, (lhs, a1, a2, scale)
        h1, h2 = cuda_corr.backward(f1, f2, coords, ii, jj, h, 3)
        c1, c2 = cuda_corr.backward(f1, f2, coords, ii, jj, 2.0 * g - 0.5 * h, 3)
        assert _rel(c1, 2.0 * g1 - 0.5 * h1) <= 2e-5 and _rel(c2, 2.0 * g2 - 0.5 * h2) <= 2e-5
        z1, z2 = cuda_corr.backward(f1, f2, coords + 5000.0, ii, jj, g, 3)
        assert z1.abs().max().item() == 0.0 and z2.abs().max().item() == 0.0


def test_pack_cache_is_not_fooled_by_recycled_addresses():
    """the drop-in cuda_corr.forward caches the pixel-major copy of its inputs; a NEW tensor that the caching allocator
    places at the address of a freed one (same shape, same version counter) must not hit the old entry"""
    from devo_b200 import cuda_corr
    P = corr_problem(n_frames=2, patches_per_frame=4, H4=24, W4=32, seed=5)
    ii, jj = P["kk"].cuda(), P["jj"].cuda()
    coords = P["coords"].cuda()
    outs, ptrs = [], []
    for rep in range(4):
        g = torch.Generator(device="cuda").manual_seed(rep)
        gmap = (torch.randn(P["gmap"].shape, device="cuda", generator=g) / 4).half()
        fmap = (torch.randn(P["pyramid"][0].shape, device="cuda", generator=g) / 4).half()
        ptrs.append(fmap.data_ptr())
        (got,) = cuda_corr.forward(gmap, fmap, coords, ii, jj, 3)
        ref = ocorr.corr_forward(gmap.cpu(), fmap.cpu(), coords.cpu(), P["kk"], P["jj"], 3)
        assert (got.cpu().double() - ref.double()).abs().max().item() <= 2e-2, rep
        del gmap, fmap, got
