"""SO3 / RxSO3 / SE3 / Sim3 group objects over the B200 lietorch backend.

The public surface is the one the reference's callers use (devo/lietorch/groups.py:51-322 -- constructors `Identity`,
`IdentityLike`, `Random`, `InitFromVec`, `exp`; element methods `log inv mul retr adj adjT Jinv act matrix translation vec
scale`; `*`; tensor-like plumbing; `cat`, `stack`, `LieGroupParameter`); the implementation is this package's own: one
generic element class driven by a small per-group description table, with every backend call funnelled through
`_dispatch`.  Storage: SO3 [qx qy qz qw]; RxSO3 [q s]; SE3 [t q]; Sim3 [t q s]; tangent vectors [tau phi (sigma)].
"""
import math

import torch

from . import group_ops as _ops
from .broadcasting import broadcast_inputs


def _flatten_shape_args(args):
    """`Identity(2, 3)`, `Identity((2, 3))` and `Identity(torch.Size([2, 3]))` all mean batch shape (2, 3)"""
    if len(args) == 1 and not isinstance(args[0], int):
        return tuple(int(d) for d in args[0])
    return tuple(int(d) for d in args)


def _dispatch(group_id, fn, first, second=None):
    """run an autograd Function of group_ops on batch-flattened operands and restore the batch shape"""
    flat, batch = broadcast_inputs(first, second)
    return fn.apply(group_id, *flat).view(batch + (-1,))


class LieGroup:
    """an array of group elements; `data[..., embedded_dim]` holds the parameters"""

    # filled in by the concrete groups below
    group_name = ""
    group_id = 0
    manifold_dim = 0      # K: tangent dimension
    embedded_dim = 0      # N: stored parameters per element
    id_elem = None

    def __init__(self, data):
        self.data = data

    # ------------------------------------------------------------------ introspection
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
        return tuple(self.shape) + (self.manifold_dim,)

    def __repr__(self):
        return "%s: size=%s, device=%s, dtype=%s" % (self.group_name, self.shape, self.device, self.dtype)

    def _new(self, data):
        return type(self)(data)

    def _call(self, fn, other=None):
        return _dispatch(self.group_id, fn, self.data, other)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def apply_op(cls, op, x, y=None):
        return _dispatch(cls.group_id, op, x, y)

    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        shape = _flatten_shape_args(batch_shape)
        one = cls.id_elem.to(device=kwargs.get("device"), dtype=kwargs.get("dtype", cls.id_elem.dtype))
        n = math.prod(shape) if shape else 1
        return cls(one.expand(n, cls.embedded_dim).clone().view(shape + (cls.embedded_dim,)))

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        shape = _flatten_shape_args(batch_shape)
        return cls.exp(sigma * torch.randn(shape + (cls.manifold_dim,), **kwargs))

    @classmethod
    def InitFromVec(cls, data):
        return cls(_dispatch(cls.group_id, _ops.FromVec, data))

    @classmethod
    def exp(cls, x):
        return cls(_dispatch(cls.group_id, _ops.Exp, x))

    # ------------------------------------------------------------------ unary maps
    def log(self):
        return self._call(_ops.Log)

    def vec(self):
        return self._call(_ops.ToVec)

    def inv(self):
        return self._new(self._call(_ops.Inv))

    # ------------------------------------------------------------------ binary maps
    def mul(self, other):
        return self._new(self._call(_ops.Mul, other.data))

    def retr(self, a):
        """left retraction  Exp(a) * X"""
        step = _dispatch(self.group_id, _ops.Exp, a)
        return self._new(_dispatch(self.group_id, _ops.Mul, step, self.data))

    def adj(self, a):
        return self._call(_ops.Adj, a)

    def adjT(self, a):
        return self._call(_ops.AdjT, a)

    def Jinv(self, a):
        return self._call(_ops.Jinv, a)

    def act(self, p):
        width = p.shape[-1]
        if width not in (3, 4):
            raise ValueError("act expects points with 3 or 4 components, got %d" % width)
        return self._call(_ops.Act3 if width == 3 else _ops.Act4, p)

    def __mul__(self, other):
        if isinstance(other, LieGroup):
            return self.mul(other)
        if torch.is_tensor(other):
            return self.act(other)
        return NotImplemented

    # ------------------------------------------------------------------ homogeneous forms
    def _batch_ones(self):
        return [1] * (self.data.dim() - 1)

    def matrix(self):
        """4x4 homogeneous matrices: the group acting on the columns of I_4"""
        eye = torch.eye(4, dtype=self.dtype, device=self.device).view(self._batch_ones() + [4, 4])
        cols = self._new(self.data.unsqueeze(-2)).act(eye)
        return cols.transpose(-1, -2)

    def translation(self):
        origin = torch.zeros(self._batch_ones() + [4], dtype=self.dtype, device=self.device)
        origin[..., 3] = 1.0
        return self._call(_ops.Act4, origin)

    # ------------------------------------------------------------------ tensor-like plumbing
    def __getitem__(self, index):
        return self._new(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def view(self, dims):
        return self._new(self.data.view(tuple(dims) + (self.embedded_dim,)))

    def detach(self):
        return self._new(self.data.detach())

    def unbind(self, dim=0):
        return [self._new(part) for part in self.data.unbind(dim=dim)]

    def to(self, *args, **kwargs):
        return self._new(self.data.to(*args, **kwargs))

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self._new(self.data.cuda())

    def float(self, device=None):
        return self._new(self.data.to(torch.float32))

    def double(self, device=None):
        return self._new(self.data.to(torch.float64))


def _pad_parts(ref, translation=False, scale=False):
    """zero translation / unit scale columns shaped like the batch of `ref`"""
    out = []
    if translation:
        out.append(torch.zeros_like(ref[..., :3]))
    if scale:
        out.append(torch.ones_like(ref[..., :1]))
    return out


class SO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SO3", 1, 3, 4
    id_elem = torch.tensor([0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SE3):                      # rotation part of a rigid motion
            data = data.data[..., 3:7]
        super().__init__(data)


class RxSO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "RxSO3", 2, 4, 5
    id_elem = torch.tensor([0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        if isinstance(data, Sim3):                     # rotation + scale part of a similarity
            data = data.data[..., 3:8]
        super().__init__(data)


class SE3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SE3", 3, 6, 7
    id_elem = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SO3):
            q = data.data
            data = torch.cat(_pad_parts(q, translation=True) + [q], dim=-1)
        super().__init__(data)

    def scale(self, s):
        """multiply the translations by per-element factors `s`"""
        return SE3(torch.cat([self.data[..., :3] * s.unsqueeze(-1), self.data[..., 3:]], dim=-1))


class Sim3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "Sim3", 4, 7, 8
    id_elem = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SO3):
            q = data.data
            data = torch.cat(_pad_parts(q, translation=True) + [q] + _pad_parts(q, scale=True), dim=-1)
        elif isinstance(data, SE3):
            data = torch.cat([data.data] + _pad_parts(data.data, scale=True), dim=-1)
        elif isinstance(data, Sim3):
            data = data.data
        super().__init__(data)


class LieGroupParameter(torch.Tensor):
    """A trainable tangent-space offset around a fixed group element: `retr()` is the current estimate Exp(self) * group.
    The tensor itself is the (initially zero) offset, so optimisers see an ordinary leaf."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    def __new__(cls, group, requires_grad=True):
        zero = torch.zeros(group.tangent_shape, dtype=group.data.dtype, device=group.data.device, requires_grad=True)
        return torch.Tensor._make_subclass(cls, zero, requires_grad)

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
        rhs = other.retr() if isinstance(other, LieGroupParameter) else other
        return self.retr() * rhs

    def __getitem__(self, index):
        return self.retr()[index]

    def add_(self, update, alpha):
        """fold `alpha * update` into the anchor element (the offset itself stays zero)"""
        self.group = self.group.exp(alpha * update) * self.group


def cat(group_list, dim):
    return type(group_list[0])(torch.cat([g.data for g in group_list], dim=dim))


def stack(group_list, dim):
    return type(group_list[0])(torch.stack([g.data for g in group_list], dim=dim))
