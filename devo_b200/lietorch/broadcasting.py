"""Batch-dimension broadcasting for binary group ops (role of devo/lietorch/broadcasting.py)."""


def check_broadcastable(x, y):
    assert x.dim() == y.dim()
    for n, m in zip(x.shape[:-1], y.shape[:-1]):
        assert n == m or n == 1 or m == 1


def broadcast_inputs(x, y=None):
    """flatten the batch dims of x (and y, after expanding size-1 batch dims against each
    other) to 2-D contiguous tensors; returns (inputs, batch_shape)"""
    if y is None:
        return (x.reshape(-1, x.shape[-1]).contiguous(),), tuple(x.shape[:-1])
    check_broadcastable(x, y)
    out_shape = tuple(max(n, m) for n, m in zip(x.shape[:-1], y.shape[:-1]))
    xe = x.expand(out_shape + (x.shape[-1],))
    ye = y.expand(out_shape + (y.shape[-1],))
    return (xe.reshape(-1, x.shape[-1]).contiguous(), ye.reshape(-1, y.shape[-1]).contiguous()), out_shape
