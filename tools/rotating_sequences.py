"""The bench step on K independent sequences processed round-robin (model weights shared, every other buffer per
sequence): with K x ~87 MB of state the inputs of a step are larger than L2 and cold by themselves -- no flush kernel, no
dirty flush lines, and the kernels' code / the model weights stay as resident as a busy GPU keeps them.
python tools/rotating_sequences.py [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
Ks = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
ops = []
for k in range(max(Ks)):
    op, up, wl = bench.build_engine(dev)
    fm = bench.load_state(op, wl, dev)
    if ops:
        op.packed = ops[0][0].packed                 # one model: the packed weights are shared by all sequences
    ops.append((op, wl, fm))
graphs = []
for op, wl, (fmap, gmap, imap) in ops:
    M, f = wl["patches_per_frame"], wl["n_frames"] - 1

    def body(op=op, fmap=fmap, gmap=gmap, imap=imap, M=M, f=f):
        op.ingest_frame(f, fmap[f], gmap[f * M:(f + 1) * M], imap[f * M:(f + 1) * M], overlap=True)
        op._iteration(reset_geometry=True)
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
    graphs.append(g)
stream = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K in Ks:
    for mode in ("rotate", "rotate + flush"):
        n = 120
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for k in range(20):
            graphs[k % K].replay()
        for k in range(n):
            if mode != "rotate":
                flush.zero_()
            ev[k][0].record(stream)
            graphs[k % K].replay()
            ev[k][1].record(stream)
        torch.cuda.synchronize(dev)
        us = sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
        print("K = %d sequences, %-15s %.1f us per step (%.0f it/s)  status %s" % (K, mode, us, 1e6 / us, [int(o[0].status_sticky.item()) for o in ops[:K]]))
