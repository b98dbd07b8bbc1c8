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

# ---- what kind of "cold" is it?  (a) flush, then idle before the replay; (b) a flush that only READS (L2 full of clean
# lines); (c) flush, then a dense read of everything the operator touches; (d) a small flush (16 MiB: weights and state evicted?)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
everything = sets["the GRU workspace too"] + [op.plan_kk.perm, op.plan_kk.gid, op.plan_ij.perm, op.plan_ij.gid, op.plan_kk.ix, op.plan_kk.jx, op.kk]


def run(name, pre):
    t = []
    for k in range(25):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if k >= 5:
            t.append(e0.elapsed_time(e1) * 1e3)
    t.sort()
    print("%-68s %7.1f us" % (name, t[len(t) // 2]))


def dense_touch():
    global sink
    for x in everything:
        sink += x.view(torch.uint8).view(-1).sum()


run("write flush, then 200 us idle", lambda: (flush.zero_(), torch.cuda._sleep(400000)))
run("read-only flush (sum of 256 MiB)", lambda: big.sum())
run("write flush, read flush", lambda: (flush.zero_(), big.sum()))
run("write flush, then a dense read of every buffer the operator touches", lambda: (flush.zero_(), dense_touch()))
run("dense read of every buffer only (no flush)", dense_touch)
run("200 us idle only (no flush)", lambda: torch.cuda._sleep(400000))

# ---- is it the TLB?  a pre-kernel that touches many pages but hardly any data, and one that touches nothing new
huge = torch.empty(8 << 30, dtype=torch.uint8, device=dev)
huge.zero_()
torch.cuda.synchronize()
pages = huge[::2 << 20]          # one byte per 2 MiB page: 4096 pages, 4 KB of data
few = huge[:64 << 20:2 << 20]    # 32 pages


def touch(v):
    global sink
    sink += v.sum()


run("tiny kernel that touches no new memory (sink += 1)", lambda: sink.add_(1))
run("one byte from each of 32 other pages", lambda: touch(few))
run("one byte from each of 4096 other pages (8 GiB span, 4 KB of data)", lambda: touch(pages))
run("4096 pages, then 200 us idle", lambda: (touch(pages), torch.cuda._sleep(400000)))
run("no pre-kernel (warm), again", lambda: None)

# ---- how much foreign data does it take?
for mb in (2, 8, 32, 64, 128):
    part = big[:mb << 20]
    run("memset of %d MiB" % mb, lambda part=part: part.zero_())
for mb in (8, 32, 128):
    part = big[:mb << 20]
    run("read (sum) of %d MiB" % mb, lambda part=part: touch(part))
run("dense read of the weights only (6 MB, no flush)", lambda: touch(op.packed.W.view(torch.uint8).view(-1)))
run("dense read of the corr rows only (10.8 MB, no flush)", lambda: touch(op.corr_buf.view(torch.uint8).view(-1)))
run("dense read of the hidden state only (9.4 MB, no flush)", lambda: touch(scratch.buf.view(torch.uint8).view(-1)))

# ---- is it the SMs' L1 / shared-memory state?  the same 8 MiB read with ld.ca (L1 allocating), ld.cg (L2 only), ld.cs,
# with and without a large dynamic shared-memory request, on all SMs or a few
import ctypes
from devo_b200 import _lib
L = ctypes.CDLL(_lib.LIB_PATH)
L.devo_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
sink_u = torch.zeros(4, dtype=torch.int32, device=dev)
part = big[:8 << 20]


def dbg(mode, smem, blocks, buf=part):
    L.devo_debug_read(buf.data_ptr(), buf.numel(), mode, smem, blocks, sink_u.data_ptr(), _lib.stream_ptr(dev))


for mode, nm in ((0, "ld.ca"), (1, "ld.cg"), (2, "ld.cs")):
    run("8 MiB read, %s, 1184 CTAs, 1 KB smem" % nm, lambda mode=mode: dbg(mode, 1024, 1184))
run("8 MiB read, ld.cg, 1184 CTAs, 100 KB smem", lambda: dbg(1, 100 << 10, 1184))
run("8 MiB read, ld.cg, 148 CTAs, 200 KB smem", lambda: dbg(1, 200 << 10, 148))
run("8 MiB read, ld.ca, 148 CTAs, 200 KB smem", lambda: dbg(0, 200 << 10, 148))
run("8 MiB read, ld.ca, 8 CTAs, 1 KB smem", lambda: dbg(0, 1024, 8))
run("64 KB read, ld.ca, 1184 CTAs, 1 KB smem", lambda: dbg(0, 1024, 1184, big[:64 << 10]))
run("hidden state read, ld.cg, 1184 CTAs", lambda: dbg(1, 1024, 1184, scratch.buf.view(torch.uint8).view(-1)))
run("hidden state read, ld.ca, 1184 CTAs", lambda: dbg(0, 1024, 1184, scratch.buf.view(torch.uint8).view(-1)))

# ---- or is it simply whether the GPU was idle before the replay?  back-to-back replays, no idle gap
def burst(name, pre, n=20):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tp = 0.0
        if pre is not None:                       # time of the pre-kernels alone, subtracted below
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(n):
                pre()
            p1.record(); torch.cuda.synchronize()
            tp = p0.elapsed_time(p1) * 1e3 / n
        e0.record()
        for _ in range(n):
            if pre is not None:
                pre()
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n - tp)
    ts.sort()
    print("%-68s %7.1f us per replay (pre-kernel alone %.1f us)" % (name, ts[len(ts) // 2], tp))


burst("20 replays back to back", None)
burst("20 x (64 MiB ld.cg read on all SMs + replay)", lambda: dbg(1, 1024, 1184, big[:64 << 20]))
burst("20 x (64 MiB memset + replay)", lambda: big[:64 << 20].zero_())
burst("20 x (256 MiB memset + replay)", lambda: flush.zero_())

# ---- a wide kernel WITHOUT memory traffic (spin), and the flush followed by an ld.cg re-read of everything the operator touches
def spin(cycles, blocks):
    L.devo_debug_read(part.data_ptr(), cycles, 3, 1024, blocks, sink_u.data_ptr(), _lib.stream_ptr(dev))


def retouch():
    for x in everything:
        dbg(1, 1024, 1184, x.view(torch.uint8).view(-1))


burst("20 x (10 us spin on 148 x 8 CTAs + replay)", lambda: spin(19000, 1184))
burst("20 x (50 us spin on 148 x 8 CTAs + replay)", lambda: spin(95000, 1184))
burst("20 x (50 us spin on 8 CTAs + replay)", lambda: spin(95000, 8))
burst("20 x (256 MiB memset + ld.cg re-read of all operator buffers + replay)", lambda: (flush.zero_(), retouch()))
burst("20 x (ld.cg re-read of all operator buffers + replay)", retouch)
burst("20 x (8 MiB ld.cg read + replay)", lambda: dbg(1, 1024, 1184))
burst("20 x (32 MiB ld.cg read + replay)", lambda: dbg(1, 1024, 1184, big[:32 << 20]))

# ---- which buffers?  flush, then ONE ld.cg re-read kernel per chosen buffer (the torch reductions used above disturb L2 themselves)
def flush_then(bufs):
    flush.zero_()
    for x in bufs:
        dbg(1, 1024, 1184, x.view(torch.uint8).view(-1))


burst("20 x (flush + replay)", lambda: flush_then([]))
burst("20 x (flush + weights + replay)", lambda: flush_then([op.packed.W, op.packed.W0]))
burst("20 x (flush + hidden state + replay)", lambda: flush_then([scratch.buf]))
burst("20 x (flush + corr rows + replay)", lambda: flush_then([op.corr_buf]))
burst("20 x (flush + workspace + replay)", lambda: flush_then([op._gru_ws]))
burst("20 x (flush + weights + state + corr + ctx + replay)", lambda: flush_then([op.packed.W, op.packed.W0, scratch.buf, op.corr_buf, op.imap]))
burst("20 x (flush + weights + state + corr + ctx + workspace + replay)", lambda: flush_then([op.packed.W, op.packed.W0, scratch.buf, op.corr_buf, op.imap, op._gru_ws]))
print("sizes (MB): W %.1f W0 %.1f state %.1f corr %.1f imap %.1f workspace %.1f" % tuple(
    x.numel() * x.element_size() / 1e6 for x in (op.packed.W, op.packed.W0, scratch.buf, op.corr_buf, op.imap, op._gru_ws)))

# ---- code vs data: a SECOND operator instance (own state / rows / workspace, the same packed weights and the same kernels)
# replayed after the flush warms code + weights but none of the first instance's data
op2, _, wl2 = build_engine(dev)
load_state(op2, wl2, dev)
op2.packed = op.packed
op2.step()
torch.cuda.synchronize()
scratch2 = GruState(op2.E, dev).set(op2.get_net())
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2), torch.no_grad():
    op2.update.forward_mma(None, op2.imap, op2.kk, op2.corr_buf, op2.plan_kk, op2.plan_ij, op2.Np, op2.Nf * op2.Nf, op2.packed,
                           workspace=op2._gru_ws, state=scratch2, tile_local=op2.tile_local)
mine = [scratch.buf, op.corr_buf, op.imap, op._gru_ws, op.plan_kk.perm, op.plan_kk.gid, op.plan_ij.perm, op.plan_ij.gid, op.plan_kk.ix, op.plan_kk.jx, op.kk]
burst("20 x (flush + replay)                                       [all cold]", lambda: flush.zero_())
burst("20 x (flush + other instance's replay + replay)    [code + weights warm]", lambda: (flush.zero_(), g2.replay()))
burst("20 x (flush + ld.cg re-read of own data + replay)           [data warm]", lambda: flush_then(mine + [op.packed.W, op.packed.W0]))
burst("20 x (flush + other instance + re-read of own data + replay) [all warm]", lambda: (flush.zero_(), g2.replay(), [dbg(1, 1024, 1184, x.view(torch.uint8).view(-1)) for x in mine]))
burst("20 x (other instance's replay + replay), no flush", lambda: g2.replay())
