"""Oracle (TEST INFRASTRUCTURE): ctypes binding of oracle/corr_c.c -- the plain-C / OpenMP restatement of the reference's
correlation lookup (correlation_kernel.cu:82-136,193-233).  `build()` compiles it with gcc into oracle/_build/ (git-ignored).
Only tests/ and bench.py's CPU-baseline legs use it; nothing in devo_b200/ does."""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libcorr_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "corr_c.c")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(src):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", src, "-o", SO, "-lm"], check=True)
    return SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        h = ctypes.CDLL(SO)
        h.corr_oracle_forward.restype = ctypes.c_int
        h.corr_oracle_forward.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 6
        _lib = h
    return _lib


def corr_forward(fmap1, fmap2, coords, ii, jj, radius):
    """== cuda_corr.forward(...)[0] for float32 CPU tensors, batch 1: [1,E,2r+1,2r+1,P,P]"""
    assert fmap1.shape[0] == 1 and fmap2.shape[0] == 1 and coords.shape[0] == 1
    f1 = fmap1[0].float().contiguous()
    f2 = fmap2[0].float().contiguous()
    co = coords[0].float().contiguous()
    ii = ii.to(torch.int64).contiguous()
    jj = jj.to(torch.int64).contiguous()
    E, P = co.shape[0], co.shape[-1]
    C, H, W = f2.shape[1], f2.shape[2], f2.shape[3]
    D1 = 2 * radius + 1
    out = torch.empty(1, E, D1, D1, P, P, dtype=torch.float32)
    rc = _load().corr_oracle_forward(f1.data_ptr(), f2.data_ptr(), co.data_ptr(), ii.data_ptr(), jj.data_ptr(),
                                     out.data_ptr(), E, C, H, W, P, radius)
    if rc != 0:
        raise RuntimeError("corr_oracle_forward: unsupported size")
    return out
