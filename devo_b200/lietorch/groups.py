"""Lie-group classes SO3 / RxSO3 / SE3 / Sim3 with the reference's public surface
(devo/lietorch/groups.py:51-322): Identity, Random, exp/log, inv, mul, retr, adj, adjT,
Jinv, act (3- or 4-vectors), matrix, translation, vec/InitFromVec, scale, indexing,
view, `*`, cat/stack and LieGroupParameter.  Data layout: SO3 [qx,qy,qz,qw];
RxSO3 [q,s]; SE3 [t,q]; Sim3 [t,q,s]; tangent [tau, phi(, sigma)]."""
import numpy as np
import torch

from .broadcasting import broadcast_inputs
from .group_ops import Exp, Log, Inv, Mul, Adj, AdjT, Jinv, Act3, Act4, ToVec, FromVec


def _as_shape(batch_shape):
    if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list, torch.Size)):
        return tuple(batch_shape[0])
    return tuple(batch_shape)


class LieGroupParameter(torch.Tensor):
    """a tangent-space parameter around a fixed group element (for optimisers)"""

    from torch._C import _disabled_torch_function_impl
    __torch_function__ = _disabled_torch_function_impl

    def __new__(cls, group, requires_grad=True):
        data = torch.zeros(group.tangent_shape, device=group.data.device, dtype=group.data.dtype, requires_grad=True)
        return torch.Tensor._make_subclass(cls, data, requires_grad)

    def __init__(self, group):
        self.group = group

    def retr(self):
        return self.group.retr(self)

    def log(self):
        return self.retr().log()

    def inv(self):
        return self.retr().inv()

    def adj(self, a):
        return self.retr().adj(a)

    def __mul__(self, other):
        if isinstance(other, LieGroupParameter):
            return self.retr() * other.retr()
        return self.retr() * other

    def add_(self, update, alpha):
        self.group = self.group.exp(alpha * update) * self.group

    def __getitem__(self, index):
        return self.retr().__getitem__(index)


class LieGroup:
    group_name = None
    group_id = None
    manifold_dim = None
    embedded_dim = None
    id_elem = None

    def __init__(self, data):
        self.data = data

    def __repr__(self):
        return "{}: size={}, device={}, dtype={}".format(self.group_name, self.shape, self.device, self.dtype)

    # ---- properties
    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def tangent_shape(self):
        return self.data.shape[:-1] + (self.manifold_dim,)

    # ---- construction
    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        shape = _as_shape(batch_shape)
        data = cls.id_elem.reshape(1, -1)
        if "device" in kwargs:
            data = data.to(kwargs["device"])
        if "dtype" in kwargs:
            data = data.type(kwargs["dtype"])
        data = data.repeat(int(np.prod(shape)), 1)
        return cls(data).view(shape)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def InitFromVec(cls, data):
        return cls(cls.apply_op(FromVec, data))

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        shape = _as_shape(batch_shape)
        xi = torch.randn(shape + (cls.manifold_dim,), **kwargs)
        return cls.exp(sigma * xi)

    @classmethod
    def apply_op(cls, op, x, y=None):
        inputs, out_shape = broadcast_inputs(x, y)
        data = op.apply(cls.group_id, *inputs)
        return data.view(out_shape + (-1,))

    @classmethod
    def exp(cls, x):
        return cls(cls.apply_op(Exp, x))

    # ---- group ops
    def vec(self):
        return self.apply_op(ToVec, self.data)

    def log(self):
        return self.apply_op(Log, self.data)

    def inv(self):
        return self.__class__(self.apply_op(Inv, self.data))

    def mul(self, other):
        return self.__class__(self.apply_op(Mul, self.data, other.data))

    def retr(self, a):
        dX = self.__class__.apply_op(Exp, a)
        return self.__class__(self.apply_op(Mul, dX, self.data))

    def adj(self, a):
        return self.apply_op(Adj, self.data, a)

    def adjT(self, a):
        return self.apply_op(AdjT, self.data, a)

    def Jinv(self, a):
        return self.apply_op(Jinv, self.data, a)

    def act(self, p):
        if p.shape[-1] == 3:
            return self.apply_op(Act3, self.data, p)
        if p.shape[-1] == 4:
            return self.apply_op(Act4, self.data, p)
        raise ValueError("act expects points with 3 or 4 components")

    def matrix(self):
        I = torch.eye(4, dtype=self.dtype, device=self.device)
        I = I.view([1] * (self.data.dim() - 1) + [4, 4])
        return self.__class__(self.data[..., None, :]).act(I).transpose(-1, -2)

    def translation(self):
        p = torch.as_tensor([0.0, 0.0, 0.0, 1.0], dtype=self.dtype, device=self.device)
        p = p.view([1] * (self.data.dim() - 1) + [4])
        return self.apply_op(Act4, self.data, p)

    def __mul__(self, other):
        if isinstance(other, LieGroup):
            return self.mul(other)
        if isinstance(other, torch.Tensor):
            return self.act(other)
        return NotImplemented

    # ---- tensor-like plumbing
    def detach(self):
        return self.__class__(self.data.detach())

    def view(self, dims):
        return self.__class__(self.data.view(tuple(dims) + (self.embedded_dim,)))

    def __getitem__(self, index):
        return self.__class__(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def to(self, *args, **kwargs):
        return self.__class__(self.data.to(*args, **kwargs))

    def cpu(self):
        return self.__class__(self.data.cpu())

    def cuda(self):
        return self.__class__(self.data.cuda())

    def float(self, device=None):
        return self.__class__(self.data.float())

    def double(self, device=None):
        return self.__class__(self.data.double())

    def unbind(self, dim=0):
        return [self.__class__(x) for x in self.data.unbind(dim=dim)]


class SO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SO3", 1, 3, 4
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data[..., 3:7]
        super().__init__(data)


class RxSO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "RxSO3", 2, 4, 5
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        if isinstance(data, Sim3):
            data = data.data[..., 3:8]
        super().__init__(data)


class SE3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SE3", 3, 6, 7
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SO3):
            data = torch.cat([torch.zeros_like(data.data[..., :3]), data.data], -1)
        super().__init__(data)

    def scale(self, s):
        t, q = self.data.split([3, 4], -1)
        return SE3(torch.cat([t * s.unsqueeze(-1), q], dim=-1))


class Sim3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "Sim3", 4, 7, 8
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SO3):
            q = data.data
            data = torch.cat([torch.zeros_like(q[..., :3]), q, torch.ones_like(q[..., :1])], -1)
        elif isinstance(data, SE3):
            data = torch.cat([data.data, torch.ones_like(data.data[..., :1])], -1)
        elif isinstance(data, Sim3):
            data = data.data
        super().__init__(data)


def cat(group_list, dim):
    return group_list[0].__class__(torch.cat([X.data for X in group_list], dim=dim))


def stack(group_list, dim):
    return group_list[0].__class__(torch.stack([X.data for X in group_list], dim=dim))
