"""The two graph plans of an S8 step (kk x jj neighbours / patch groups, frame-pair groups), L2 flushed before each call:
    python tools/plan_timing.py"""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_engine, load_state, time_us
dev = torch.device("cuda")
op, up, wl = build_engine(dev); load_state(op, wl, dev)
s = torch.cuda.current_stream(); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("plan_kk %.1f us  plan_ij %.1f us" % (time_us(op.plan_kk.update, s, flush), time_us(op.plan_ij.update, s, flush)))
