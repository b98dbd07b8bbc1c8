"""Which leg bounds the end-to-end pipeline of bench.py (`e2e`): the per-step H2D copy, or the step's graph replay?
Runs bench.run_e2e in probe mode: the same host loop with only the copies, only the graph replays, and both; then
variants (the result read back by a DMA node instead of the copy kernel, a quarter of the upload, the upload in 8
chunks, a step without the state refresh and result copies)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
op, up, wl = bench.build_engine(dev)
bench.load_state(op, wl, dev)
for kw in ({}, dict(d2h_kernel=False), dict(h2d_frac=0.25), dict(chunks=8), dict(light_body=True)):
    print(kw, json.dumps(bench.run_e2e(op, wl, dev, 200, probe=True, **kw)))
