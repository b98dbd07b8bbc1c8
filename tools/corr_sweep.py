"""time devo_corr_lookup_fused on the S8 workload for the loaded library variant (DEVO_B200_LIB)"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import cuda_corr, synthetic

def timeit(fn, flush, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ts[3:])
    return 1e3 * t[len(t) // 2]

def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for C in (128, 64):
        wl = synthetic.make_workload(C=C)
        gm = cuda_corr.pack_gmap(wl["gmap"].to(dev))
        lv = [cuda_corr.pack_pixel_major(wl["fmap"].to(dev), s) for s in (1, 4)]
        # coords: GT reprojection of the patch grids (approximate: centre + grid)
        from bench import build_engine, load_state
        if C == 128:
            op, up, wl2 = build_engine(dev)
            load_state(op, wl2, dev)
            op.step()
            coords = op.coords[0].clone()
        ii, jj = wl["kk"].to(dev), wl["jj"].to(dev)
        res["C%d_L2" % C] = timeit(lambda: cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj), flush)
        res["C%d_lvl1" % C] = timeit(lambda: cuda_corr.lookup_fused(gm, lv[:1], (1,), coords, ii, jj), flush)
        res["C%d_lvl4" % C] = timeit(lambda: cuda_corr.lookup_fused(gm, lv[1:], (4,), coords, ii, jj), flush)
        # L2-warm (no flush)
        for _ in range(3): cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): cuda_corr.lookup_fused(gm, lv, (1, 4), coords, ii, jj)
        b.record(); torch.cuda.synchronize()
        res["C%d_L2_warm" % C] = 1e3 * a.elapsed_time(b) / 20
    print(os.environ.get("DEVO_B200_LIB", "default"), json.dumps({k: round(v, 1) for k, v in res.items()}))

main()
