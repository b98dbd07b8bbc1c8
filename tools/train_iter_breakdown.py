"""Where a training-loop iteration (BASELINE.json config 5 shape) spends its time: CUDA events around each stage, forward
and backward.  python tools/train_iter_breakdown.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from devo_b200 import altcorr, ba as dba, lietorch as lt, projective_ops as pops, synthetic
    dev = torch.device("cuda", 0)
    wl = synthetic.make_workload(n_frames=15, patches_per_frame=96, seed=1234, feat_dtype=torch.float32)
    up = synthetic.make_update_module(seed=1234).to(dev)
    ii, jj, kk = wl["ii"].to(dev), wl["jj"].to(dev), wl["kk"].to(dev)
    E = ii.numel()
    fmap0, gmap0 = wl["fmap"].to(dev)[None].float(), wl["gmap"].to(dev)[None].float()
    imap = wl["imap"].to(dev)[None].float()
    poses0, patches0, intr = wl["poses0"].to(dev)[None], wl["patches0"].to(dev)[None], wl["intrinsics"].to(dev)[None]
    poses_gt = lt.SE3(wl["poses_gt"].to(dev)[None])
    bounds = [-64, -64, wl["W4"] + 64, wl["H4"] + 64]
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    for rep in range(4):
        marks.clear()
        mark("start")
        fmap = fmap0.clone().requires_grad_(True)
        gmap = gmap0.clone().requires_grad_(True)
        pyr = [fmap, torch.nn.functional.avg_pool2d(fmap[0], 4, 4)[None]]
        poses, patches = lt.SE3(poses0.clone()), patches0.clone()
        net = torch.zeros(1, E, 384, device=dev)
        coords = pops.transform(poses, patches, intr, ii, jj, kk).permute(0, 1, 4, 2, 3).contiguous()
        mark("transform")
        c1 = altcorr.corr(gmap, pyr[0], coords / 1, kk, jj, 3)
        c2 = altcorr.corr(gmap, pyr[1], coords / 4, kk, jj, 3)
        corr = torch.stack([c1, c2], -1).view(1, E, -1)
        mark("corr fwd")
        net, (delta, weight, _) = up(net, imap[:, kk], corr, None, ii, jj, kk)
        mark("update fwd")
        target = coords[..., 1, 1].detach() + delta
        for _ in range(2):
            poses, patches = dba.BA(poses, patches, intr, target, weight, 1e-4, ii, jj, kk, bounds, ep=10.0, fixedp=1)
        mark("ba x2 fwd")
        gt = pops.transform(poses_gt, patches0, intr, ii, jj, kk)
        est = pops.transform(poses, patches, intr, ii, jj, kk)
        loss = (est - gt).norm(dim=-1).mean()
        mark("loss fwd")
        # backward in pieces: loss -> (delta, weight) ; -> corr ; -> features
        gd, gw = torch.autograd.grad(loss, [delta, weight], retain_graph=True)
        mark("bwd loss+ba")
        (gc,) = torch.autograd.grad([delta, weight], [corr], [gd, gw], retain_graph=True)
        mark("bwd update")
        torch.autograd.grad(corr, [gmap, fmap], gc)
        mark("bwd corr")
        torch.cuda.synchronize()
    t0 = marks[0][1]
    prev = t0
    for name, e in marks[1:]:
        print("%-14s %8.3f ms" % (name, prev.elapsed_time(e)))
        prev = e
    print("%-14s %8.3f ms" % ("total", t0.elapsed_time(marks[-1][1])))


if __name__ == "__main__":
    main()
