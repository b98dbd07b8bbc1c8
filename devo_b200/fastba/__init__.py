"""`fastba` operator API (devo/fastba/ba.py:4-8): BA (in place), neighbors, reproject."""
from .. import cuda_ba

neighbors = cuda_ba.neighbors
reproject = cuda_ba.reproject


def BA(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations=2):
    """`iterations` Gauss-Newton steps in place on poses[t0:t1] and the patches in unique(kk)"""
    return cuda_ba.forward(poses.data, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations)
