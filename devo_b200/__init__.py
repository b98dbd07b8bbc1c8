"""devo_b200 -- B200 (sm_100a) implementation of the DEVO update-operator hot path:
sparse patch correlation (altcorr), SE3/Sim3 Lie-group ops (lietorch), projective ops and
Gauss-Newton bundle adjustment (fastba, ba), behind the reference's operator API.

    devo_b200.cuda_corr / cuda_ba / lietorch_backends   drop-ins for the pybind modules
    devo_b200.altcorr / fastba / lietorch / ba / projective_ops   the Python operator API
    devo_b200.engine.UpdateOperator                      fused, CUDA-graph-captured iteration

All compute goes through include/devo_b200.h (libdevo_b200.so, hand-written CUDA); there
is no CPU path in this package.
"""
import sys
import types

__version__ = "0.1.0"


def install_shims():
    """Register this package's backends under the module names the reference's Python
    wrappers import (`cuda_corr`, `cuda_ba`, `lietorch_backends`, `torch_scatter`), so
    devo/altcorr/correlation.py, devo/fastba/ba.py, devo/lietorch/group_ops.py, devo/ba.py
    and devo/blocks.py run unchanged on top of libdevo_b200.so."""
    from . import cuda_ba, cuda_corr, lietorch_backends, scatter
    sys.modules["cuda_corr"] = cuda_corr
    sys.modules["cuda_ba"] = cuda_ba
    sys.modules["lietorch_backends"] = lietorch_backends
    ts = types.ModuleType("torch_scatter")
    ts.scatter_sum = scatter.scatter_sum
    ts.scatter_softmax = scatter.scatter_softmax
    sys.modules["torch_scatter"] = ts
