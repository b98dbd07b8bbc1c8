import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import _lib, cuda_ba, synthetic
wl = synthetic.make_workload()
dev = torch.device("cuda")
a = lambda: (wl["poses0"][None].to(dev).contiguous(), wl["patches0"][None].to(dev).contiguous())
rest = (wl["intrinsics"][None].to(dev), wl["targets"][None].to(dev), wl["weights"][None].to(dev), torch.tensor([1e-4], device=dev),
        wl["ii"].to(dev), wl["jj"].to(dev), wl["kk"].to(dev), 1, 8, 1)
for _ in range(3):
    p, x = a(); cuda_ba.forward_async(p, x, *rest)
torch.cuda.synchronize()
h = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 16)()
h.devo_ba_debug_clocks(buf)
c = list(buf)
print("cycles: accumulate(start->solve start) %d | reduce %d | ownership %d | eliminate %d | backsub %d | solve end %d" %
      (c[1]-c[0], c[2]-c[1], c[3]-c[2], c[4]-c[3], c[5]-c[4], c[6]-c[5]))
