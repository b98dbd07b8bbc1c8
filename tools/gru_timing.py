"""Where the time goes inside the fused update operator (csrc/gru_mma.cu): CUDA-event time of forward_mma vs the
cuBLAS + glue path, and the %globaltimer stamps CTA 0 of each of the 6 gru_mma launches records.
    python tools/gru_timing.py"""
import ctypes
import os
import sys

import torch

SL = 80            # debug stamps per launch (kDbgSlots in csrc/gru_mma.cu)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    from devo_b200 import _lib, cuda_ba
    from devo_b200.update import FrozenCast, PackedUpdateWeights, Update
    from problems import fully_connected_graph
    torch.manual_seed(0)
    nf, m = int(os.environ.get('GRU_NF', 8)), int(os.environ.get('GRU_M', 96))
    ii, jj, kk = [t.cuda() for t in fully_connected_graph(nf, m)]
    E, Np = ii.numel(), nf * m
    up = Update(3).cuda().eval()
    net = 0.5 * torch.randn(1, E, 384, device="cuda")
    imap = (0.25 * torch.randn(1, Np, 384, device="cuda")).half()
    corr = torch.zeros(E, 896, device="cuda", dtype=torch.float16)
    corr[:, :882] = torch.randn(E, 882, device="cuda").half()
    plan_kk = cuda_ba.GraphPlan(kk, jj, Np, nf)
    plan_ij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(ii), -1, 1, want_neighbors=False)
    fc = FrozenCast(torch.float16)
    packed = PackedUpdateWeights(up, torch.float16, 896)
    TL = os.environ.get('GRU_TILE_LOCAL', '0') == '1'      # the merged first program (needs a tile-local graph: S8 is one)
    ctx = imap[:, kk].contiguous()

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    with torch.no_grad():
        from devo_b200.update import GruState
        st = GruState(E, "cuda").set(net)
        t_mma = timeit(lambda: up.forward_mma(None, imap, kk, corr, plan_kk, plan_ij, Np, nf * nf, packed, state=st, tile_local=TL))
        net16 = net.half()
        t_cub = timeit(lambda: up.forward_fused(net16, ctx, corr.view(1, E, 896), plan_kk, plan_ij, Np, nf * nf, fc))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            up.forward_mma(None, imap, kk, corr, plan_kk, plan_ij, Np, nf * nf, packed, state=st, tile_local=TL)
        t_graph = timeit(g.replay)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            up.forward_fused(net16, ctx, corr.view(1, E, 896), plan_kk, plan_ij, Np, nf * nf, fc)
        t_cub_graph = timeit(g2.replay)
        print("forward_mma %.1f us (graph replay %.1f us)   cuBLAS+glue path %.1f us eager, %.1f us graph replay" % (t_mma, t_graph, t_cub, t_cub_graph))
        L = _lib.lib()
        L.devo_gru_debug_timing.argtypes = [ctypes.c_void_p]
        L.devo_gru_debug_timing(None)
        for _ in range(int(os.environ.get("GRU_REPS", 32))):      # the stamps of the LAST update survive (16-launch ring): clocks are up by then
            up.forward_mma(None, imap, kk, corr, plan_kk, plan_ij, Np, nf * nf, packed, state=st, tile_local=TL)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (2 * 16 * SL))()
        L.devo_gru_debug_timing(buf)
    names = ["corr+norm", "c1", "c2+agg_kk g,f", "h_kk+agg_ij g,f", "h_ij+gru+heads"]
    nl = [3, 2, 4, 3, 7]
    if TL:                     # tile-local graph: one program up to the frame-pair g / f, then the GRU program
        names = ["corr+norm|c1|c2|agg_kk|h_kk|g,f_ij", "h_ij+gru+heads"]
        nl = [12, 7]
    reps = int(os.environ.get("GRU_REPS", 32))
    # absolute view (ns clock shared by all SMs): when CTA 0 of each launch started, finished its prologue, issued its first
    # MMA, finished its last epilogue and exited -- the distance between "last epilogue of launch k" and "first MMA of
    # launch k+1" is what a kernel boundary (+ the gather prologue, + a segment reduction for two of them) costs
    base = None
    rows = []
    for k, n in enumerate(nl):
        k16 = (len(names) * (reps - 1) + k) % 16
        s_ = [buf[SL * k16 + q] for q in range(SL)]
        if base is None:
            base = s_[0]
        s0 = s_[0] if s_[0] else base                     # fused launch: only program 0 has the kernel-start stamp
        rows.append((names[k], (s0 - base) / 1e3, (s_[2] - base) / 1e3, (s_[4] - base) / 1e3, (s_[9 + 6 * (n - 1)] - base) / 1e3, (s_[3] - base) / 1e3))
    print("absolute (us): launch            start   pro-done  first-mma  last-epi   exit   | boundary = next first-mma - last-epi")
    for i, r in enumerate(rows):
        nxt = "%.1f" % (rows[i + 1][3] - r[4]) if i + 1 < len(rows) else "-"
        print("               %-16s %7.1f %9.1f %10.1f %9.1f %7.1f  | %s" % (r + (nxt,)))
    for k, (nm, n) in enumerate(zip(names, nl)):
        k16 = (len(names) * (reps - 1) + k) % 16
        s = [buf[SL * k16 + q] for q in range(SL)]
        cyc = [buf[16 * SL + SL * k16 + q] for q in range(SL)]
        mhz = (cyc[3] - cyc[0]) / max(s[3] - s[0], 1) * 1e3
        t0 = s[0] if s[0] else base
        rel = lambda q: (s[q] - t0) / 1e3 if s[q] else float("nan")
        # per layer: MMA issue start - issue end | epilogue: N-tile 0 ready, N-tile 1 ready, chunk loop done, layer done
        line = "%-16s setup %.1f pro %.1f |" % (nm, rel(1), rel(2))
        for l in range(n):
            b = 4 + 6 * l
            line += " L%d mma %.1f-%.1f epi %.1f/%.1f loop %.1f end %.1f |" % (l, rel(b), rel(b + 1), rel(b + 2), rel(b + 3), rel(b + 4), rel(b + 5))
        line += " exit %.1f us  (SM clock %.0f MHz)" % (rel(3), mhz)
        print(line)



if __name__ == "__main__":
    main()
