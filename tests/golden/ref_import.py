"""Import the REFERENCE's own Python modules (devo/lietorch, devo/projective_ops.py,
devo/ba.py) from /root/reference in THIS container, on top of shims for the
native/third-party modules that are absent here:

    lietorch_backends -> oracle.lie              (Eigen is absent)
    torch_scatter     -> oracle.scatter          (torch_scatter 2.0.9 absent)
    cuda_ba, cuda_corr-> empty stubs             (CUDA-only; not exercised on CPU)

Used only by tests/golden/make_golden.py and by CPU tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

REF = os.environ.get("DEVO_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "devo"))


def load():
    """returns (ref_lietorch, ref_pops, ref_ba) modules of the reference"""
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from oracle import lie as olie
    from oracle import scatter as oscatter

    sys.modules["lietorch_backends"] = olie
    ts = types.ModuleType("torch_scatter")
    ts.scatter_sum = oscatter.scatter_sum
    ts.scatter_softmax = oscatter.scatter_softmax
    sys.modules["torch_scatter"] = ts
    for name, attrs in (("cuda_ba", ["forward", "neighbors", "reproject"]),
                        ("cuda_corr", ["forward", "backward", "patchify_forward", "patchify_backward"])):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, None)
        sys.modules[name] = m
    # import the reference package under an alias so it cannot shadow ours
    spec = importlib.util.spec_from_file_location(
        "refdevo", os.path.join(REF, "devo", "__init__.py"),
        submodule_search_locations=[os.path.join(REF, "devo")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["refdevo"] = pkg
    spec.loader.exec_module(pkg)
    lt = importlib.import_module("refdevo.lietorch")
    pops = importlib.import_module("refdevo.projective_ops")
    ba = importlib.import_module("refdevo.ba")
    return lt, pops, ba
