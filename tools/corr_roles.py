"""Role-level cycle accounting of corr_fast_kernel (library built with -DDEVO_CORR_TIMING, see tools/build_variant.sh):
per CTA, cycles each role spent blocked in each of its waits.  usage: DEVO_B200_LIB=...ct.so python tools/corr_roles.py"""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import _lib, cuda_corr, synthetic
from bench import build_engine, load_state

dev = torch.device("cuda")
op, up, wl = build_engine(dev)
load_state(op, wl, dev)
op.step()
coords = op.coords[0].clone()
gm = cuda_corr.pack_gmap(wl["gmap"].to(dev))
lv = [cuda_corr.pack_pixel_major(wl["fmap"].to(dev), s) for s in (1, 4)]
ii, jj = wl["kk"].to(dev), wl["jj"].to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
h = ctypes.CDLL(_lib.LIB_PATH)
names = ["prod0: wait coords", "prod0: wait empty", "prod0: total",
         "mma: wait tempty", "mma: wait full", "mma: total",
         "epi0: wait rfull", "epi0: wait tfull", "epi0: wait tcgen05.ld", "epi0: bar.sync", "epi0: total"]
for mode in ("cold", "warm"):
    for _ in range(3):
        if mode == "cold":
            flush.zero_()
        cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (148 * 16))()
    h.devo_corr_debug_clocks(buf)
    t = torch.tensor(list(buf), dtype=torch.float64).view(148, 16)
    print("== %s L2: cycles per CTA (mean / max over 148 CTAs; 83 items per CTA)" % mode)
    for q, n in enumerate(names):
        print("   %-24s %9.0f %9.0f" % (n, t[:, q].mean().item(), t[:, q].max().item()))
