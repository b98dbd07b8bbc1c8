"""Oracle (TEST INFRASTRUCTURE): `cuda_ba.neighbors` (devo/fastba/ba.cpp:104-149).

For every edge e: among the edges with the same ii[e] (callers pass the patch id
kk), stable-sorted by jj (frame id), ix[e] = edge index of the predecessor and
jx[e] = edge index of the successor, -1 at the ends.  Integer => bit-exact.
Also the index plumbing the update operator needs from torch.unique
(devo/blocks.py:41, devo/fastba/ba_cuda.cu:435-437).
"""
import numpy as np
import torch


def neighbors(ii, jj):
    ii_n = ii.cpu().numpy().astype(np.int64)
    jj_n = jj.cpu().numpy().astype(np.int64)
    E = ii_n.shape[0]
    order = np.lexsort((np.arange(E), jj_n, ii_n))       # by ii, then jj, then edge index (stable)
    ix = np.full(E, -1, dtype=np.int64)
    jx = np.full(E, -1, dtype=np.int64)
    if E > 1:
        same = ii_n[order[1:]] == ii_n[order[:-1]]
        ix[order[1:][same]] = order[:-1][same]
        jx[order[:-1][same]] = order[1:][same]
    return torch.from_numpy(ix), torch.from_numpy(jx)


def neighbors_loops(ii, jj):
    """literal restatement with python loops (small cases only)"""
    ii_l, jj_l = ii.tolist(), jj.tolist()
    groups = {}
    for e, k in enumerate(ii_l):
        groups.setdefault(k, []).append(e)
    ix = [-1] * len(ii_l)
    jx = [-1] * len(ii_l)
    for k, idx in groups.items():
        idx = sorted(idx, key=lambda e: jj_l[e])        # python sort is stable
        for t, e in enumerate(idx):
            ix[e] = idx[t - 1] if t > 0 else -1
            jx[e] = idx[t + 1] if t < len(idx) - 1 else -1
    return torch.tensor(ix, dtype=torch.int64), torch.tensor(jx, dtype=torch.int64)


def unique_inverse(x):
    """torch.unique(x, sorted=True, return_inverse=True)"""
    return torch.unique(x.cpu(), sorted=True, return_inverse=True)
