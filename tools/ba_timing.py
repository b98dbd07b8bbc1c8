"""fastba on the S8 workload: CUDA-event time of a graph replay (2 Gauss-Newton iterations, shared plan), and -- when the
library was built with DEVO_NVCC_EXTRA=-DDEVO_BA_TIMING -- the clock64 stamps of the solve stages.
    python tools/ba_timing.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import _lib, cuda_ba, synthetic

wl = synthetic.make_workload()
dev = torch.device("cuda")
poses0, patches0 = wl["poses0"][None].to(dev).contiguous(), wl["patches0"][None].to(dev).contiguous()
poses, patches = poses0.clone(), patches0.clone()
intr, tgt, wgt = wl["intrinsics"][None].to(dev), wl["targets"][None].to(dev), wl["weights"][None].to(dev)
lm = torch.tensor([1e-4], device=dev)
ii, jj, kk = wl["ii"].to(dev), wl["jj"].to(dev), wl["kk"].to(dev)
plan = cuda_ba.GraphPlan(kk, jj, wl["n_frames"] * wl["patches_per_frame"], wl["n_frames"])
status = torch.zeros(1, dtype=torch.int32, device=dev)
ws = torch.empty(_lib.lib().devo_ba_workspace(ii.numel(), 7), dtype=torch.uint8, device=dev)


def run(iters):
    poses.copy_(poses0)
    patches.copy_(patches0)
    cuda_ba.forward_async(poses, patches, intr, tgt, wgt, lm, ii, jj, kk, 1, 8, iters, status=status, plan=plan, workspace=ws)


for iters in (1, 2):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run(iters)
        run(iters)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run(iters)
    for _ in range(5):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print("fastba S8, %d iteration(s), graph replay incl. 2 state-reset copies: %.1f us  (status %d)" % (iters, a.elapsed_time(b) * 10, int(status.item())))
h = ctypes.CDLL(_lib.LIB_PATH)
if hasattr(h, "devo_ba_debug_clocks"):
    buf = (ctypes.c_longlong * 24)()
    h.devo_ba_debug_clocks(buf)
    c = list(buf)
    r = lambda i: (c[i] - c[0]) / 1e3
    print("last accumulate launch, CTA grid/2: after the PDL wait %.1f | depth update applied %.1f" % (r(16), r(17)))
    print("CTA grid/2 (us since kernel start): edge terms done %.1f | E_k done %.1f | partial accumulated %.1f | ticket taken %.1f" % (r(8), r(9), r(10), r(11)))
    print("CTA grid/2 accumulate phase: thread 0 lists done %.1f, E_k rows done %.1f | thread 433 (y^T y corner) lists done %.1f, E_k rows done %.1f | barrier passed %.1f" %
          (r(12), r(13), r(14), r(15), r(7)))
    print("solving CTA: solve start %.1f | reduced %.1f | owned %.1f | eliminated %.1f | back-substituted %.1f | retracted %.1f" %
          (r(1), r(2), r(3), r(4), r(5), r(6)))
