"""`lietorch_backends` -- drop-in for the reference's pybind module
(devo/lietorch/src/lietorch.cpp:286-316): the 19 batched Lie-group ops.
Inputs are 2-D [batch, dim] tensors; CHECK_CONTIGUOUS as in the reference
(lietorch.cpp:7).  CUDA only: there is no CPU backend in this build.
"""
import torch

from . import _lib

_DIMS = {1: (3, 4), 2: (4, 5), 3: (6, 7), 4: (7, 8)}   # group id -> (K tangent dim, N embedding dim)


def _prep(*ts):
    for t in ts:
        if not t.is_contiguous():
            raise RuntimeError("lietorch_backends: input must be contiguous")
        if not t.is_cuda:
            raise RuntimeError("lietorch_backends: CUDA tensors only (the CPU backend is not part of this build)")
        if t.dtype not in (torch.float32, torch.float64):
            raise RuntimeError("lietorch_backends: float32/float64 only, got %s" % t.dtype)
    d = ts[0].dtype
    for t in ts[1:]:
        if t.dtype != d:
            raise RuntimeError("lietorch_backends: dtype mismatch")


def _call(name, gid, ref, ptrs, n):
    _dims(gid)
    fn = getattr(_lib.lib(), "devo_lie_" + name)
    _lib.check(fn(int(gid), _lib.dtype_code(ref), *ptrs, int(n), _lib.stream_ptr(ref.device)), "lie_" + name)


def _dims(gid):
    try:
        return _DIMS[gid]
    except KeyError:
        raise RuntimeError("lietorch_backends: unknown group id %r (SO3=1, RxSO3=2, SE3=3, Sim3=4)" % (gid,))


def _new(ref, *shape):
    return torch.empty(*shape, dtype=ref.dtype, device=ref.device)


def expm(gid, a):
    _prep(a)
    X = _new(a, a.shape[0], _dims(gid)[1])
    _call("expm", gid, a, [a.data_ptr(), X.data_ptr()], a.shape[0])
    return X


def expm_backward(gid, grad, a):
    _prep(grad, a)
    da = _new(a, *a.shape)
    _call("expm_backward", gid, a, [grad.data_ptr(), a.data_ptr(), da.data_ptr()], a.shape[0])
    return [da]


def logm(gid, X):
    _prep(X)
    a = _new(X, X.shape[0], _dims(gid)[0])
    _call("logm", gid, X, [X.data_ptr(), a.data_ptr()], X.shape[0])
    return a


def logm_backward(gid, grad, X):
    _prep(grad, X)
    dX = _new(X, *X.shape)
    _call("logm_backward", gid, X, [grad.data_ptr(), X.data_ptr(), dX.data_ptr()], X.shape[0])
    return [dX]


def inv(gid, X):
    _prep(X)
    Y = _new(X, *X.shape)
    _call("inv", gid, X, [X.data_ptr(), Y.data_ptr()], X.shape[0])
    return Y


def inv_backward(gid, grad, X):
    _prep(grad, X)
    dX = _new(X, *X.shape)
    _call("inv_backward", gid, X, [grad.data_ptr(), X.data_ptr(), dX.data_ptr()], X.shape[0])
    return [dX]


def mul(gid, X, Y):
    _prep(X, Y)
    Z = _new(X, *X.shape)
    _call("mul", gid, X, [X.data_ptr(), Y.data_ptr(), Z.data_ptr()], X.shape[0])
    return Z


def mul_backward(gid, grad, X, Y):
    _prep(grad, X, Y)
    dX, dY = _new(X, *X.shape), _new(Y, *Y.shape)
    _call("mul_backward", gid, X, [grad.data_ptr(), X.data_ptr(), Y.data_ptr(), dX.data_ptr(), dY.data_ptr()], X.shape[0])
    return [dX, dY]


def _binary(name, gid, X, a):
    _prep(X, a)
    b = _new(a, *a.shape)
    _call(name, gid, X, [X.data_ptr(), a.data_ptr(), b.data_ptr()], X.shape[0])
    return b


def _binary_backward(name, gid, grad, X, a):
    _prep(grad, X, a)
    dX, da = _new(X, *X.shape), _new(a, *a.shape)
    _call(name, gid, X, [grad.data_ptr(), X.data_ptr(), a.data_ptr(), dX.data_ptr(), da.data_ptr()], X.shape[0])
    return [dX, da]


def adj(gid, X, a):
    return _binary("adj", gid, X, a)


def adj_backward(gid, grad, X, a):
    return _binary_backward("adj_backward", gid, grad, X, a)


def adjT(gid, X, a):
    return _binary("adjT", gid, X, a)


def adjT_backward(gid, grad, X, a):
    return _binary_backward("adjT_backward", gid, grad, X, a)


def act(gid, X, p):
    return _binary("act", gid, X, p)


def act_backward(gid, grad, X, p):
    return _binary_backward("act_backward", gid, grad, X, p)


def act4(gid, X, p):
    return _binary("act4", gid, X, p)


def act4_backward(gid, grad, X, p):
    return _binary_backward("act4_backward", gid, grad, X, p)


def as_matrix(gid, X):
    _prep(X)
    T = _new(X, X.shape[0], 4, 4)
    _call("as_matrix", gid, X, [X.data_ptr(), T.data_ptr()], X.shape[0])
    return T


def projector(gid, X):
    _prep(X)
    N = _dims(gid)[1]
    P = _new(X, X.shape[0], N, N)
    _call("projector", gid, X, [X.data_ptr(), P.data_ptr()], X.shape[0])
    return P


def Jinv(gid, X, a):
    _prep(X, a)
    b = _new(a, *a.shape)
    _call("jinv", gid, X, [X.data_ptr(), a.data_ptr(), b.data_ptr()], X.shape[0])
    return b
