"""Where does a step's time go beyond its four main kernels?  CUDA-graph replays (L2 flushed before each) of the bench step
and of the step with pieces removed:  python tools/step_variants.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_engine, load_state, time_us

dev = torch.device("cuda")
op, up, wl = build_engine(dev)
fmap, gmap, imap = load_state(op, wl, dev)
M, Nf = wl["patches_per_frame"], wl["n_frames"]
nf = Nf - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ingest(overlap):
    op.ingest_frame(nf, fmap[nf], gmap[nf * M:(nf + 1) * M], imap[nf * M:(nf + 1) * M], overlap=overlap)


variants = {
    "bench step (ingest on a side stream + iteration, geometry reset)": lambda: (ingest(True), op._iteration(reset_geometry=True)),
    "iteration only (geometry reset)": lambda: op._iteration(reset_geometry=True),
    "iteration only, no geometry reset": lambda: op._iteration(reset_geometry=False),
    "ingest only (main stream)": lambda: ingest(False),
    "ingest in line, then iteration": lambda: (ingest(False), op._iteration(reset_geometry=True)),
}
s = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s), torch.no_grad():
    for name, fn in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        t = time_us(g.replay, s, flush, warm=5, n=50)
        print("%-70s %7.1f us" % (name, t))
