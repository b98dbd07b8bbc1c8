"""Stand-alone time of the two segment softmax+sum reductions of an S8 update (kk: 768 groups x 8 rows, ij: 64 groups x 96
rows, 384 channels, fp16), graph replay, L2 warm:  python tools/segment_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import _lib, cuda_ba, synthetic

wl = synthetic.make_workload()
dev = torch.device("cuda")
ii, jj, kk = wl["ii"].to(dev), wl["jj"].to(dev), wl["kk"].to(dev)
E = ii.numel()
Np, N = wl["n_frames"] * wl["patches_per_frame"], wl["n_frames"]
plan_kk = cuda_ba.GraphPlan(kk, jj, Np, N)
plan_ij = cuda_ba.GraphPlan(ii * 12345 + jj, jj, 12345 * N + N, N) if hasattr(cuda_ba, "GraphPlan") else None
g = torch.randn(E, 384, device=dev).half()
f = torch.randn(E, 384, device=dev).half()
L = _lib.lib()


def run(plan, groups, y):
    _lib.check(L.devo_segment_softmax_sum(g.data_ptr(), f.data_ptr(), plan.perm.data_ptr(), plan.gstart.data_ptr(),
                                          plan.ngroups.data_ptr(), groups, y.data_ptr(), 1, E, 384, _lib.stream_ptr(dev)), "seg")


for name, plan, groups in (("kk", plan_kk, Np), ("ij", plan_ij, 64)):
    y = torch.empty(groups, 384, device=dev, dtype=torch.half)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run(plan, groups, y)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            run(plan, groups, y)
    for _ in range(3):
        gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    print("segment softmax+sum %s: %d groups, %.2f us per launch (10 back-to-back launches per graph, PDL)" % (name, int(plan.ngroups.item()), a.elapsed_time(b) * 1000 / 200))
