"""Which of its inputs makes the update operator slower with a cold L2?  The operator's graph is replayed after an L2 flush
and after re-touching (reading) chosen inputs:  python tools/gru_cold_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_engine, load_state

dev = torch.device("cuda")
op, up, wl = build_engine(dev)
load_state(op, wl, dev)
op.step()
torch.cuda.synchronize()
from devo_b200.update import GruState
scratch = GruState(op.E, dev).set(op.get_net())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g), torch.no_grad():
    op.update.forward_mma(None, op.imap, op.kk, op.corr_buf, op.plan_kk, op.plan_ij, op.Np, op.Nf * op.Nf, op.packed,
                          workspace=op._gru_ws, state=scratch, tile_local=op.tile_local)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sets = {
    "nothing (cold)": [],
    "weights": [op.packed.W, op.packed.W0],
    "hidden state": [scratch.buf],
    "corr rows": [op.corr_buf],
    "weights + state": [op.packed.W, op.packed.W0, scratch.buf],
    "weights + state + corr + ctx + plans": [op.packed.W, op.packed.W0, scratch.buf, op.corr_buf, op.imap, op.plan_kk.perm, op.plan_kk.gid,
                                            op.plan_ij.perm, op.plan_ij.gid, op.plan_kk.ix, op.plan_kk.jx, op.kk],
    "the GRU workspace too": [op.packed.W, op.packed.W0, scratch.buf, op.corr_buf, op.imap, op._gru_ws],
}
sink = torch.zeros(1, device=dev)
for name, ts in sets.items():
    t = []
    for k in range(25):
        flush.zero_()
        for x in ts:
            sink += x.view(torch.uint8).view(-1)[::64].sum() if x.dtype != torch.uint8 else x[::64].sum()     # one byte per 64: every line read
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        if k >= 5:
            t.append(e0.elapsed_time(e1) * 1e3)
    t.sort()
    print("re-touched after the flush: %-40s %7.1f us" % (name, t[len(t) // 2]))
t = []
for k in range(25):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t.append(e0.elapsed_time(e1) * 1e3)
t.sort()
print("no flush at all (warm)                                               %7.1f us" % t[len(t) // 2])
