"""Autograd glue between torch and `lietorch_backends` (mirrors the op set of
devo/lietorch/group_ops.py:6-100: Exp, Log, Inv, Mul, Adj, AdjT, Act3, Act4, Jinv,
ToMatrix, ToVec, FromVec).  Gradients w.r.t. group elements are K-vectors stored in
the first K of N slots, as produced by the backend."""
import torch

from .. import lietorch_backends as _be


def _make_op(name, fwd, bwd):
    class _Op(torch.autograd.Function):
        @staticmethod
        def forward(ctx, group_id, *inputs):
            ctx.group_id = group_id
            ctx.save_for_backward(*inputs)
            return fwd(group_id, *inputs)

        @staticmethod
        def backward(ctx, grad):
            if bwd is None:
                raise AssertionError("Backward operation not implemented for %s" % name)
            grads = bwd(ctx.group_id, grad.contiguous(), *ctx.saved_tensors)
            return (None,) + tuple(grads)

    _Op.__name__ = name
    _Op.__qualname__ = name
    return _Op


Exp = _make_op("Exp", _be.expm, _be.expm_backward)
Log = _make_op("Log", _be.logm, _be.logm_backward)
Inv = _make_op("Inv", _be.inv, _be.inv_backward)
Mul = _make_op("Mul", _be.mul, _be.mul_backward)
Adj = _make_op("Adj", _be.adj, _be.adj_backward)
AdjT = _make_op("AdjT", _be.adjT, _be.adjT_backward)
Act3 = _make_op("Act3", _be.act, _be.act_backward)
Act4 = _make_op("Act4", _be.act4, _be.act4_backward)
Jinv = _make_op("Jinv", _be.Jinv, None)
ToMatrix = _make_op("ToMatrix", _be.as_matrix, None)


class FromVec(torch.autograd.Function):
    """embedding vector -> group object (identity forward; projector^+ backward)"""

    @staticmethod
    def forward(ctx, group_id, x):
        ctx.group_id = group_id
        ctx.save_for_backward(x)
        return x

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        J = _be.projector(ctx.group_id, x)
        return None, torch.matmul(grad.unsqueeze(-2), torch.linalg.pinv(J)).squeeze(-2)


class ToVec(torch.autograd.Function):
    """group object -> embedding vector (identity forward; projector backward)"""

    @staticmethod
    def forward(ctx, group_id, x):
        ctx.group_id = group_id
        ctx.save_for_backward(x)
        return x

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        J = _be.projector(ctx.group_id, x)
        return None, torch.matmul(grad.unsqueeze(-2), J).squeeze(-2)
