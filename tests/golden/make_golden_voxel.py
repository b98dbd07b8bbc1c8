"""Golden vectors for the event-voxel front from the REFERENCE's own functions (utils/voxel_utils.py::std / rescale,
utils/event_utils.py::to_voxel_grid), imported from /root/reference in this container (h5py, absent here, is stubbed: the
functions used do not touch it).     python tests/golden/make_golden_voxel.py   -> tests/golden/voxel.pt"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DEVO_REFERENCE", "/root/reference")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    h5 = types.ModuleType("h5py")
    h5.File = object
    sys.modules.setdefault("h5py", h5)
    vu = _load(os.path.join(REF, "utils", "voxel_utils.py"), "ref_voxel_utils")
    eu = _load(os.path.join(REF, "utils", "event_utils.py"), "ref_event_utils")
    g = torch.Generator().manual_seed(3)
    vox = (torch.rand(2, 3, 5, 24, 32, generator=g) < 0.15).float() * torch.randn(2, 3, 5, 24, 32, generator=g)
    rng = np.random.RandomState(5)
    n = 4000
    xs, ys = rng.randint(0, 32, n).astype(np.float64), rng.randint(0, 24, n).astype(np.float64)
    ts = np.sort(rng.uniform(10.0, 10.05, n))
    ps = rng.randint(0, 2, n)
    out = dict(vox=vox, std_seq=vu.std(vox.clone()), std_frame=vu.std(vox.clone(), sequence=False), rescale=vu.rescale(vox.clone()),
               events=dict(xs=xs, ys=ys, ts=ts, ps=ps), grid=eu.to_voxel_grid(xs, ys, ts, ps.copy(), H=24, W=32, nb_of_time_bins=5))
    torch.save(out, os.path.join(HERE, "voxel.pt"))
    print({k: (tuple(v.shape) if hasattr(v, "shape") else type(v).__name__) for k, v in out.items()})


if __name__ == "__main__":
    main()
