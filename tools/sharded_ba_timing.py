"""Edge-sharded fastba on N GPUs (S8 graph split by owning patch): per-call time of the NCCL form (one all-reduce of
[S|y] per Gauss-Newton iteration) and of the peer-memory form (reduction fused into the solve kernel).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_ba_timing.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import cuda_ba, dist as d, synthetic


def main():
    rank, world, local = d.init("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    wl = synthetic.make_workload()
    Np = wl["n_frames"] * wl["patches_per_frame"]
    sel = d.shard_edges_by_patch(wl["kk"], Np, rank, world)
    f = lambda t: t.to(dev).contiguous()
    poses0, patches0 = f(wl["poses0"][None]), f(wl["patches0"][None])
    intr, tgt, wgt = f(wl["intrinsics"][None]), f(wl["targets"][None][:, sel]), f(wl["weights"][None][:, sel])
    ii, jj, kk = f(wl["ii"][sel]), f(wl["jj"][sel]), f(wl["kk"][sel])
    lm = torch.tensor([1e-4], device=dev)
    iters = 2
    res = {}
    for mode in ("nccl", "peer"):
        poses, patches = poses0.clone(), patches0.clone()
        ba = cuda_ba.ShardedBA(poses, patches, intr, tgt, wgt, lm, ii, jj, kk, 1, wl["n_frames"])
        if mode == "peer":
            ba.enable_peer()

        def call():
            poses.copy_(poses0)
            patches.copy_(patches0)
            for itr in range(iters):
                if mode == "peer":
                    ba.accumulate_peer(itr)
                    ba.solve_peer(itr)
                else:
                    ba.accumulate(itr)
                    dist.all_reduce(ba.system)
                    ba.solve(itr)
            ba.finish(iters)

        for _ in range(5):
            call()
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            call()
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / 50 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # device time of the kernels involved (the eager call above is launch-bound on the host)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(20):
                call()
            torch.cuda.synchronize(dev)
        kern = {}
        for ev in prof.key_averages():
            n = ev.key
            for tag in ("ncclDevKernel", "ba_solve_peer_kernel", "ba_solve_kernel", "ba_accumulate_kernel"):
                if tag in n:
                    tot = getattr(ev, "device_time_total", None)
                    if tot is None:
                        tot = getattr(ev, "cuda_time_total", 0.0)
                    c, tt = kern.get(tag, (0, 0.0))
                    kern[tag] = (c + ev.count, tt + tot)
        res[mode] = (float(t.item()), int(ba.status.item()), poses.clone(), {k: v[1] / max(v[0], 1) for k, v in kern.items()})
        dist.barrier()
    if rank == 0:
        same = torch.equal(res["nccl"][2], res["peer"][2])
        print("sharded fastba, S8 graph over %d GPUs, %d GN iterations per call (eager launches, max over ranks):" % (world, iters))
        print("   NCCL all-reduce of [S|y] per iteration : %.1f us per call (status %d)" % (res["nccl"][0], res["nccl"][1]))
        print("   reduction fused into the solve (peer)  : %.1f us per call (status %d)" % (res["peer"][0], res["peer"][1]))
        print("   poses identical between the two forms  : %s" % same)
        for mode in ("nccl", "peer"):
            print("   %s: mean device time per kernel (us): %s" % (mode, ", ".join("%s %.1f" % kv for kv in sorted(res[mode][3].items()))))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
