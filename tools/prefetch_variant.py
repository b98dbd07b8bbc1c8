"""Does pulling the update operator's state / weights into L2 WHILE the lookup runs shorten the (flushed) step?
A/B of the captured bench step with and without a small read kernel (ld.cg, few CTAs) forked at the start of the step.
python tools/prefetch_variant.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from devo_b200 import _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
op, up, wl = bench.build_engine(dev)
fmap, gmap, imap = bench.load_state(op, wl, dev)
M, f = wl["patches_per_frame"], wl["n_frames"] - 1
L = ctypes.CDLL(_lib.LIB_PATH)
L.devo_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
sink = torch.zeros(4, dtype=torch.int32, device=dev)
pf_stream = torch.cuda.Stream(device=dev)


def prefetch(bufs, blocks):
    for x in bufs:
        v = x.view(torch.uint8).view(-1)
        L.devo_debug_read(v.data_ptr(), v.numel(), 1, 1024, blocks, sink.data_ptr(), _lib.stream_ptr(dev))


def make(bufs, blocks):
    def body():
        cur = torch.cuda.current_stream(dev)
        if bufs:
            pf_stream.wait_stream(cur)
            with torch.cuda.stream(pf_stream):
                prefetch(bufs, blocks)
        op.ingest_frame(f, fmap[f], gmap[f * M:(f + 1) * M], imap[f * M:(f + 1) * M], overlap=True)
        op._iteration(reset_geometry=True)
        if bufs:
            cur.wait_stream(pf_stream)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3):
            body()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        body()
    return g


state = op.state.buf
variants = {
    "baseline": ([], 0),
    "weights, 8 CTAs": ([op.packed.W, op.packed.W0], 8),
    "weights + state, 16 CTAs": ([op.packed.W, op.packed.W0, state], 16),
    "weights + state + ctx, 32 CTAs": ([op.packed.W, op.packed.W0, state, op.imap], 32),
    "weights + state + workspace, 32 CTAs": ([op.packed.W, op.packed.W0, state, op._gru_ws], 32),
}
graphs = {k: make(*v) for k, v in variants.items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
for rep in range(2):
    for name, g in graphs.items():
        n = 100
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for k in range(10):
            flush.zero_(); g.replay()
        for k in range(n):
            flush.zero_()
            ev[k][0].record(stream); g.replay(); ev[k][1].record(stream)
        torch.cuda.synchronize(dev)
        us = sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
        print("%-40s %.1f us per flushed step (%.0f it/s)" % (name, us, 1e6 / us))
