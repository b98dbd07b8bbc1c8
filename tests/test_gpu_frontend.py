"""Fused patch gather (csrc/patch_gather.cu, devo_b200/frontend.py) == the tail of the reference's Patchifier.forward
(devo/enet.py:179-191): three bilinear-mode altcorr.patchify calls + coords_grid_with_index.  Bit-exact for the integer
patch centres every selector of the reference produces; float rounding only for fractional centres."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(dtype, N=3, C=128, D=384, H=30, W=40, M=24, seed=5, integer=True, border=False):
    g = torch.Generator().manual_seed(seed)
    fmap = (torch.randn(N, C, H, W, generator=g) / 4).to(dtype).cuda()
    imap = (torch.randn(N, D, H, W, generator=g) / 4).to(dtype).cuda()
    disps = (torch.rand(N, H, W, generator=g) + 0.1).cuda()
    if border:                                   # windows that stick out of the image on every side
        x = torch.randint(-2, W + 2, (N, M), generator=g)
        y = torch.randint(-2, H + 2, (N, M), generator=g)
    else:
        x = torch.randint(1, W - 1, (N, M), generator=g)
        y = torch.randint(1, H - 1, (N, M), generator=g)
    coords = torch.stack([x, y], -1).float()
    if not integer:
        coords = coords + torch.rand(N, M, 2, generator=g)
    return fmap, imap, disps, coords.cuda()


def _composed(fmap, imap, disps, coords, P=3):
    """what enet.py:179-191 does, on this library's altcorr (pinned to the reference extension elsewhere)"""
    from devo_b200 import altcorr
    N, _, H, W = fmap.shape
    im = altcorr.patchify(imap, coords, 0)
    gm = altcorr.patchify(fmap, coords, P // 2)
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda").float(), torch.arange(W, device="cuda").float(), indexing="ij")
    grid = torch.stack([xx[None].expand(N, -1, -1), yy[None].expand(N, -1, -1), disps], 1)      # coords_grid_with_index
    pt = altcorr.patchify(grid, coords, P // 2)
    return gm, im, pt


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("border", [False, True])
def test_patch_gather_integer_centres_bit_exact(dtype, border):
    from devo_b200 import cuda_corr
    from devo_b200.frontend import gather_patches
    fmap, imap, disps, coords = _inputs(dtype, border=border)
    gm, im, pt = _composed(fmap, imap, disps, coords)
    g, gpm, i2, p2 = gather_patches(fmap, imap, coords, disps, 3, planar=True, pixel_major=True)
    N, M = coords.shape[:2]
    assert torch.equal(g, gm.to(dtype).view(N * M, -1, 3, 3))
    assert torch.equal(i2, im.to(dtype).view(N * M, -1))
    assert torch.equal(p2, pt.view(N * M, 3, 3, 3))
    if dtype != torch.float32:                                       # the pixel-major copy == the engine's gmap_pack of it
        assert torch.equal(gpm, cuda_corr.pack_gmap(g))
    else:
        assert torch.equal(gpm, g.permute(0, 2, 3, 1).reshape(N * M, 9, -1))


def test_patch_gather_fractional_centres():
    from devo_b200.frontend import gather_patches
    fmap, imap, disps, coords = _inputs(torch.float32, integer=False, border=True)
    gm, im, pt = _composed(fmap, imap, disps, coords)
    g, _, i2, p2 = gather_patches(fmap, imap, coords, disps, 3)
    N, M = coords.shape[:2]
    assert (g - gm.view(N * M, -1, 3, 3)).abs().max().item() <= 1e-6
    assert (i2 - im.view(N * M, -1)).abs().max().item() <= 1e-6
    assert (p2 - pt.view(N * M, 3, 3, 3)).abs().max().item() <= 1e-4          # x / y up to 40: float rounding of the blend


def test_patch_frontend_matches_reference_patchifier_tail():
    """the reference's own altcorr.patchify (unchanged, on the reference's compiled extension) fed with the same
    encoder outputs and centres: the tail of Patchifier.forward, line by line"""
    import ref_callers
    if not ref_callers.available():
        pytest.skip("reference callers not staged (oracle/_ref/devo_py)")
    from devo_b200.frontend import PatchFrontend
    fmap, imap, disps, coords = _inputs(torch.float16, N=2, M=16)
    try:
        ns = ref_callers.use_backend("ref_ext")
        r_imap = ns.altcorr.patchify(imap, coords, 0).view(1, -1, 384, 1, 1)
        r_gmap = ns.altcorr.patchify(fmap, coords, 1).view(1, -1, 128, 3, 3)
        yy, xx = torch.meshgrid(torch.arange(30, device="cuda").float(), torch.arange(40, device="cuda").float(), indexing="ij")
        grid = torch.stack([xx[None].expand(2, -1, -1), yy[None].expand(2, -1, -1), disps], 1)
        r_patches = ns.altcorr.patchify(grid, coords, 1).view(1, -1, 3, 3, 3)
    finally:
        ref_callers.use_backend("ours")
    fe = PatchFrontend(3, pixel_major=True)
    f2, g2, i2, p2, index, gpm = fe(fmap[None], imap[None], 16, disps[None], coords=coords)
    assert torch.equal(g2, r_gmap.half()) and torch.equal(i2, r_imap.half()) and torch.equal(p2, r_patches)
    assert torch.equal(index, torch.arange(2, device="cuda").repeat_interleave(16))
    # RANDOM selection: centres strictly inside the image, reproducible with a generator
    gen = torch.Generator(device="cuda").manual_seed(3)
    out = fe(fmap[None], imap[None], 20, generator=gen)
    px = out[3][0, :, 0, 1, 1]
    assert out[1].shape == (1, 40, 128, 3, 3) and px.min().item() >= 1 and px.max().item() <= 38
    assert torch.equal(out[3][0, :, 2], torch.ones_like(out[3][0, :, 2]))        # disps=None -> ones


def test_patch_graph_vo_on_the_fused_front_end():
    """PatchGraphVO fed by `frontend_from_encoders` (encoders -> ONE gather launch -> pixel-major patches straight into the
    ring) == PatchGraphVO fed with the reference-style front end outputs (planar gmap, re-packed by the loop) for the same
    encoders and patch centres: bit-identical trajectories."""
    import types
    from devo_b200 import altcorr
    from devo_b200.update import Update
    from devo_b200.vo import PatchGraphVO, VOConfig, frontend_from_encoders
    torch.manual_seed(1)
    cfg = VOConfig(PATCHES_PER_FRAME=32, BUFFER_SIZE=64, OPTIMIZATION_WINDOW=6, PATCH_LIFETIME=5, REMOVAL_WINDOW=8)
    up = Update(3).cuda().eval()

    class Enc(torch.nn.Module):                               # stand-in encoders: the scope starts after them
        def __init__(self, c):
            super().__init__()
            self.conv = torch.nn.Conv2d(5, c, 8, stride=4, padding=2)

        def forward(self, x):
            b, n, c, h, w = x.shape
            return self.conv(x.view(b * n, c, h, w)).view(b, n, -1, h // 4, w // 4)
    fnet, inet = Enc(128).cuda().eval(), Enc(384).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(2)
    frames = [torch.randn(5, 480, 640, device="cuda", generator=g) for _ in range(10)]
    centres = [torch.stack([torch.randint(1, 159, (1, 32), device="cuda", generator=g),
                            torch.randint(1, 119, (1, 32), device="cuda", generator=g)], -1).float() for _ in frames]
    intr = torch.tensor([320.0, 320.0, 320.0, 240.0], device="cuda")

    def run(front_end):
        vo = PatchGraphVO(cfg, up, front_end)
        vo.motion_probe = lambda: 10.0
        torch.manual_seed(7)
        with torch.no_grad():
            for t, f in enumerate(frames):
                vo(float(t), f, intr)
        torch.cuda.synchronize()
        return vo

    it = iter(centres)
    fused = run(frontend_from_encoders(fnet, inet, cfg, select=lambda image, M: next(it)))

    it2 = iter(centres)

    def reference_style(image):                               # enet.py:179-191 spelled out on altcorr.patchify
        with torch.autocast("cuda", enabled=bool(cfg.MIXED_PRECISION)):
            fmap = fnet(image[None, None]) / 4.0
            imap = inet(image[None, None]) / 4.0
        fmap, imap = fmap[0], imap[0].to(fmap.dtype)
        coords = next(it2)
        im = altcorr.patchify(imap, coords, 0).to(fmap.dtype)
        gm = altcorr.patchify(fmap, coords, 1).to(fmap.dtype)
        yy, xx = torch.meshgrid(torch.arange(120, device="cuda").float(), torch.arange(160, device="cuda").float(), indexing="ij")
        grid = torch.stack([xx[None], yy[None], torch.ones(1, 120, 160, device="cuda")], 1)
        pt = altcorr.patchify(grid, coords, 1)
        return dict(fmap=fmap[0], gmap=gm[0], imap=im[0, :, :, 0, 0], patches=pt[0], clr=None)
    ref = run(reference_style)
    assert fused.n == ref.n and fused.n_updates == ref.n_updates > 0
    assert torch.equal(fused.ii, ref.ii) and torch.equal(fused.kk, ref.kk)
    assert torch.equal(fused.poses_[:fused.n], ref.poses_[:ref.n])
    assert torch.equal(fused.patches_[:fused.n], ref.patches_[:ref.n])
