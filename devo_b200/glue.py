"""Python face of the fused GRU glue kernels (csrc/gru_glue.cu, include/devo_b200.h "GRU glue")."""
import torch

from . import _lib


def _dt(t):
    return _lib.dtype_code(t)


def layernorm_add3(a, b, c, ln):
    """float32 LN(T(T(a+b)+c)) for half tensors a,b,c [rows,dim]  (norm(net + inp + corr), enet.py:82-83)"""
    rows, dim = a.shape
    out = torch.empty(rows, dim, dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().devo_glue_layernorm(0, _dt(a), a.data_ptr(), b.data_ptr(), _lib.ptr(c), 0, ln.weight.data_ptr(),
                                              ln.bias.data_ptr(), float(ln.eps), out.data_ptr(), 0, rows, dim,
                                              _lib.stream_ptr(a.device)), "glue_layernorm")
    return out


def layernorm_f32(x32, ln, half_dtype):
    """LN(x32) -> (float32 result, same values rounded to the autocast dtype)"""
    rows, dim = x32.shape
    out = torch.empty_like(x32)
    out16 = torch.empty(rows, dim, dtype=half_dtype, device=x32.device)
    _lib.check(_lib.lib().devo_glue_layernorm(1, _dt(out16), 0, 0, 0, x32.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(),
                                              float(ln.eps), out.data_ptr(), out16.data_ptr(), rows, dim,
                                              _lib.stream_ptr(x32.device)), "glue_layernorm")
    return out, out16


def layernorm_relu_half(x16, ln):
    """T(relu(LN(float(x16))))  (the LayerNorm -> ReLU -> Linear input of the corr MLP, enet.py:59-66)"""
    rows, dim = x16.shape
    out16 = torch.empty_like(x16)
    _lib.check(_lib.lib().devo_glue_layernorm(2, _dt(x16), x16.data_ptr(), 0, 0, 0, ln.weight.data_ptr(), ln.bias.data_ptr(),
                                              float(ln.eps), 0, out16.data_ptr(), rows, dim, _lib.stream_ptr(x16.device)),
               "glue_layernorm")
    return out16


def gather_mask_cast(x32, idx, half_dtype):
    rows, dim = x32.shape
    out = torch.empty(rows, dim, dtype=half_dtype, device=x32.device)
    _lib.check(_lib.lib().devo_glue_gather_mask_cast(_dt(out), x32.data_ptr(), idx.data_ptr(), out.data_ptr(), rows, dim,
                                                     _lib.stream_ptr(x32.device)), "glue_gather_mask_cast")
    return out


def residual_add_(net32, y16, gid=None, want_half=False):
    """net32 += y16[gid] (in place); optionally returns the sum rounded to y16's dtype"""
    rows, dim = net32.shape
    out16 = torch.empty(rows, dim, dtype=y16.dtype, device=net32.device) if want_half else None
    _lib.check(_lib.lib().devo_glue_residual_add(_dt(y16), net32.data_ptr(), y16.data_ptr(), _lib.ptr(gid), _lib.ptr(out16),
                                                 rows, dim, _lib.stream_ptr(net32.device)), "glue_residual_add")
    return out16


def gated_residual(x32, gate_pre, res):
    out = torch.empty_like(x32)
    _lib.check(_lib.lib().devo_glue_gated_residual(_dt(res), x32.data_ptr(), gate_pre.data_ptr(), res.data_ptr(), out.data_ptr(),
                                                   x32.numel(), _lib.stream_ptr(x32.device)), "glue_gated_residual")
    return out


def relu_cast(x32, half_dtype, relu=True, out=None):
    if out is None:
        out = torch.empty(x32.shape, dtype=half_dtype, device=x32.device)
    _lib.check(_lib.lib().devo_glue_relu_cast(_dt(out), x32.data_ptr(), out.data_ptr(), x32.numel(), int(relu),
                                              _lib.stream_ptr(x32.device)), "glue_relu_cast")
    return out


def heads(x32, W16, b16):
    """(delta [rows,2], weight [rows,2]) of the update operator's two heads from the float32 hidden state"""
    rows, dim = x32.shape
    delta = torch.empty(rows, 2, dtype=W16.dtype, device=x32.device)
    weight = torch.empty(rows, 2, dtype=W16.dtype, device=x32.device)
    _lib.check(_lib.lib().devo_glue_heads(_dt(W16), x32.data_ptr(), W16.data_ptr(), b16.data_ptr(), delta.data_ptr(),
                                          weight.data_ptr(), rows, dim, _lib.stream_ptr(x32.device)), "glue_heads")
    return delta, weight
