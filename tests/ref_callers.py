"""Run the REFERENCE's own Python callers unchanged on top of this library (GPU tests only).

`oracle/build_ref.py::stage_python()` stages a verbatim, git-ignored copy of the reference's
`devo/` and `utils/` Python files under oracle/_ref/devo_py/ (it travels to the GPU box with the
snapshot; /root/reference does not exist there).  `load()` puts that directory on sys.path, installs

  * devo_b200.install_shims():  cuda_corr, cuda_ba, lietorch_backends, torch_scatter  -> this library
  * inert stubs for the plotting / evaluation imports that are absent in this image
    (matplotlib, evo, yacs) -- none of them is on the hot path

and imports `devo.devo`, `devo.enet`, ... exactly as the reference's scripts do.  `use_backend()`
re-points the already imported wrappers at the reference's own compiled extensions
(oracle/_ref/cuda_corr_ref.so, cuda_ba_ref.so) or back at this library, so both can be compared
through the same unmodified Python.  lietorch_backends cannot be built from the reference here
(Eigen absent), so SE3 always runs on this library's backend.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref", "devo_py")


def available():
    return os.path.isfile(os.path.join(STAGED, "devo", "devo.py"))


class CfgNode(dict):
    """minimal stand-in for yacs.config.CfgNode (attribute access + merge_from_file)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self.update(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            self[k] = v

    def clone(self):
        return CfgNode(self)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, k):
            return _Anything()

    def _any(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = _stub("matplotlib", use=lambda *a, **k: None)
            mpl.pyplot = _stub("matplotlib.pyplot", __getattr__=_any)
            mpl.patches = _stub("matplotlib.patches", __getattr__=_any)
    if "evo" not in sys.modules:
        try:
            import evo  # noqa: F401
        except ImportError:
            evo = _stub("evo")
            evo.core = _stub("evo.core")
            evo.core.trajectory = _stub("evo.core.trajectory", PoseTrajectory3D=_Anything)
            evo.tools = _stub("evo.tools", plot=_Anything())
            evo.tools.plot = _stub("evo.tools.plot", __getattr__=_any)
    if "yacs" not in sys.modules:
        try:
            import yacs  # noqa: F401
        except ImportError:
            y = _stub("yacs")
            y.config = _stub("yacs.config", CfgNode=CfgNode)


_loaded = None


def load():
    """-> namespace with the reference modules: devo (devo.devo), enet, blocks, altcorr, fastba, lietorch, pops, ba, cfg"""
    global _loaded
    if _loaded is not None:
        return _loaded
    import devo_b200
    devo_b200.install_shims()
    _install_stubs()
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    ns = types.SimpleNamespace()
    ns.pkg = importlib.import_module("devo")
    assert os.path.abspath(ns.pkg.__path__[0]).startswith(STAGED), "a different `devo` package shadows the staged reference"
    ns.lietorch = importlib.import_module("devo.lietorch")
    ns.altcorr = importlib.import_module("devo.altcorr")
    ns.fastba = importlib.import_module("devo.fastba")
    ns.pops = importlib.import_module("devo.projective_ops")
    ns.ba = importlib.import_module("devo.ba")
    ns.blocks = importlib.import_module("devo.blocks")
    ns.enet = importlib.import_module("devo.enet")
    ns.devo = importlib.import_module("devo.devo")
    ns.config = importlib.import_module("devo.config")
    _loaded = ns
    return ns


def default_cfg():
    """the reference's config/default.yaml merged over devo/config.py (what evals/*.py pass to DEVO)"""
    ns = load()
    cfg = ns.config.cfg.clone()
    cfg.merge_from_file(os.path.join(STAGED, "config", "default.yaml"))
    return cfg


def use_backend(kind):
    """re-point the reference wrappers: 'ours' -> devo_b200 modules, 'ref_ext' -> the reference's compiled extensions"""
    ns = load()
    corr_mod = importlib.import_module("devo.altcorr.correlation")
    ba_mod = importlib.import_module("devo.fastba.ba")
    if kind == "ours":
        from devo_b200 import cuda_ba, cuda_corr
    elif kind == "ref_ext":
        from oracle.build_ref import load_ref
        cuda_corr, cuda_ba = load_ref("cuda_corr_ref"), load_ref("cuda_ba_ref")
        if cuda_corr is None or cuda_ba is None:
            raise RuntimeError("oracle/_ref extensions not built")
    else:
        raise ValueError(kind)
    corr_mod.cuda_corr = cuda_corr
    ba_mod.cuda_ba = cuda_ba
    for m in (ba_mod, ns.fastba):
        m.neighbors = cuda_ba.neighbors
        m.reproject = cuda_ba.reproject
    return ns
