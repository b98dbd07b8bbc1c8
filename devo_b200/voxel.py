"""Event-voxel front of the frame path: event stream -> voxel grid and voxel-grid normalisation, same
functions and argument meaning as utils/event_utils.py::to_voxel_grid and utils/voxel_utils.py::std / rescale (training)
and the normalisation block of DEVO.__call__ (devo/devo.py:419-452, inference), on CUDA kernels (csrc/voxel.cu)."""
import torch

from . import _lib


def _normalize(x, groups, mode):
    _lib.require_cuda(x)
    xf = x.contiguous().float()
    out = torch.empty_like(xf)
    n = xf.numel() // groups
    L = _lib.lib()
    ws = _lib.workspace(L.devo_voxel_workspace(groups), xf.device, "voxel")
    _lib.check(L.devo_voxel_normalize(xf.data_ptr(), out.data_ptr(), n, groups, mode, 0, ws.data_ptr(), ws.numel(),
                                      _lib.stream_ptr(xf.device)), "voxel_normalize")
    return out


def std(voxs, sequence=True):
    """utils/voxel_utils.py:6-27: standardise the non-zero entries (per batch element over the whole sequence, or per frame);
    unchanged if any group has no event"""
    b, n, c, h, w = voxs.shape
    return _normalize(voxs, b if sequence else b * n, 0).view(b, n, c, h, w)


def rescale(voxs, sequence=True):
    """utils/voxel_utils.py:30-52: positives / max, negatives / -min (the reference takes the extrema over the whole tensor)"""
    b, n, c, h, w = voxs.shape
    return _normalize(voxs, 1, 1).view(b, n, c, h, w)


def normalize_frame(image, norm="std"):
    """the normalisation DEVO.__call__ applies to one voxel frame [1,1,bins,H,W] (devo/devo.py:419-452).  Returns the
    normalised frame, or None where the reference skips the frame ('rescale' with an empty polarity)."""
    norm = norm.lower()
    if norm == "none":
        return image
    b, n, v, h, w = image.shape
    if norm in ("standard", "std"):
        return _normalize(image, b * n, 0).view(b, n, v, h, w)
    if norm in ("rescale", "norm"):
        if not bool((image > 0).any()) or not bool((image < 0).any()):
            return None
        return _normalize(image, 1, 1).view(b, n, v, h, w)
    raise NotImplementedError(norm)


def to_voxel_grid(xs, ys, ts, ps, H=480, W=640, nb_of_time_bins=5, device="cuda"):
    """utils/event_utils.py:180-231: trilinear accumulation of an event stream (x, y, t sorted ascending, polarity in
    {0 or -1, 1}) into [bins, H, W]; arrays may be numpy or torch, host or device"""
    dev = torch.device(device)
    xs = torch.as_tensor(xs).to(dev, torch.float32).contiguous()
    ys = torch.as_tensor(ys).to(dev, torch.float32).contiguous()
    ts = torch.as_tensor(ts).to(dev, torch.float64).contiguous()
    ps = torch.as_tensor(ps).to(dev, torch.float32).contiguous()
    grid = torch.zeros(nb_of_time_bins, H, W, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().devo_events_to_voxel(xs.data_ptr(), ys.data_ptr(), ts.data_ptr(), ps.data_ptr(), xs.numel(),
                                               grid.data_ptr(), nb_of_time_bins, H, W, _lib.stream_ptr(dev)), "events_to_voxel")
    return grid
