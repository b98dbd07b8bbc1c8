"""altcorr backward at the training shape (config 5: 15 frames x 96 patches, E = 21600, C = 128, float32), per level:
the pixel-major kernel alone, the whole cuda_corr.backward call (kernel + layout plumbing) and the generic kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devo_b200 import _lib, cuda_corr, projective_ops as pops, synthetic, lietorch as lt

dev = torch.device("cuda", 0)
wl = synthetic.make_workload(n_frames=15, patches_per_frame=96, seed=1234, feat_dtype=torch.float32)
ii, jj, kk = wl["ii"].to(dev), wl["jj"].to(dev), wl["kk"].to(dev)
E = ii.numel()
fmap, gmap = wl["fmap"].to(dev)[None].float(), wl["gmap"].to(dev)[None].float()
poses, patches, intr = lt.SE3(wl["poses0"].to(dev)[None]), wl["patches0"].to(dev)[None], wl["intrinsics"].to(dev)[None]
coords = pops.transform(poses, patches, intr, ii, jj, kk).permute(0, 1, 4, 2, 3).contiguous()
g = torch.randn(1, E, 7, 7, 3, 3, device=dev)


def t_us(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for s in (1, 4):
    f2 = torch.nn.functional.avg_pool2d(fmap[0], s, s)[None].contiguous() if s > 1 else fmap
    c = (coords / s).contiguous()
    _, Nf, C, H, W = f2.shape
    Np = gmap.shape[1]
    f2pm = f2[0].permute(0, 2, 3, 1).contiguous()
    g1pm = torch.zeros(Np, 9, C, device=dev)
    g2pm = torch.zeros(Nf, H, W, C, device=dev)

    def kernel():
        _lib.check(_lib.lib().devo_corr_backward_pm(gmap.data_ptr(), f2pm.data_ptr(), c.data_ptr(), kk.data_ptr(), jj.data_ptr(),
                                                    g.data_ptr(), g1pm.data_ptr(), g2pm.data_ptr(), Np, Nf, C, H, W, E,
                                                    _lib.stream_ptr(dev)), "bwd")
    tk = t_us(kernel)
    tc = t_us(lambda: cuda_corr.backward(gmap, f2, c, kk, jj, g, 3))
    cuda_corr._FORCE_GENERIC = True
    tg = t_us(lambda: cuda_corr.backward(gmap, f2, c, kk, jj, g, 3), n=3)
    cuda_corr._FORCE_GENERIC = False
    px = 121 if s == 1 else 81
    print("level 1/%d: kernel %.0f us (%.0f GB/s of box reads + reductions), cuda_corr.backward %.0f us, generic kernel path %.0f us"
          % (s, tk, E * px * C * 4 * 2 / tk / 1e3, tc, tg))
