"""Build the *unmodified-algorithm* reference CUDA extensions into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing under devo_b200/
may import this.

The reference ships two compilable torch extensions on the hot path
(reference: setup.py:13-25):
    cuda_corr = devo/altcorr/correlation.cpp + correlation_kernel.cu
    cuda_ba   = devo/fastba/ba.cpp + ba_cuda.cu
They are compiled from where they lie under /root/reference.  Two mechanical
torch-2.x API patches are applied to a scratch copy in a temp dir (never to the
repo, never to /root/reference):
    correlation_kernel.cu : `X.type()`  -> `X.scalar_type()` inside AT_DISPATCH  (4 sites)
    ba_cuda.cu            : `torch::linalg::cholesky(S)` -> `at::linalg_cholesky(S)` (1 site)
The pybind module names are changed to cuda_corr_ref / cuda_ba_ref (by the
TORCH_EXTENSION_NAME define only) so they can be imported next to ours.

lietorch_backends cannot be built: it needs Eigen 3.4.0 (reference
README.md:71-73, setup.py:27-29), which is not on this machine.

Outputs: oracle/_ref/cuda_corr_ref*.so, oracle/_ref/cuda_ba_ref*.so
(git-ignored, travel to the GPU box with gpurun).

`stage_python()` additionally stages an UNMODIFIED copy of the reference's Python callers
(devo/*.py, devo/{altcorr,fastba,lietorch}/*.py, utils/*.py, config/default.yaml) under
oracle/_ref/devo_py/ -- git-ignored like the rest of oracle/_ref, so the sources never enter
the history, but they travel to the GPU box, where /root/reference does not exist.  The GPU
tests (tests/test_gpu_reference_callers.py) import them on top of devo_b200.install_shims()
to prove that the reference's own callers run unchanged on this library.
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DEVO_REFERENCE", "/root/reference")


def _patched_copy(tmp):
    src_corr = os.path.join(REF, "devo", "altcorr")
    src_ba = os.path.join(REF, "devo", "fastba")
    files = {}
    for d, names in ((src_corr, ["correlation.cpp", "correlation_kernel.cu"]),
                     (src_ba, ["ba.cpp", "ba_cuda.cu"])):
        for n in names:
            with open(os.path.join(d, n)) as f:
                files[n] = f.read()
    k = files["correlation_kernel.cu"]
    for who in ("fmap1", "net"):
        k = k.replace("AT_DISPATCH_FLOATING_TYPES_AND_HALF(%s.type()," % who,
                      "AT_DISPATCH_FLOATING_TYPES_AND_HALF(%s.scalar_type()," % who)
    files["correlation_kernel.cu"] = k
    files["ba_cuda.cu"] = files["ba_cuda.cu"].replace(
        "torch::linalg::cholesky(S)", "at::linalg_cholesky(S)")
    for n, txt in files.items():
        with open(os.path.join(tmp, n), "w") as f:
            f.write(txt)


def build(verbose=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    stage_python()
    have = [f for f in os.listdir(OUT) if f.endswith(".so")]
    if any(f.startswith("cuda_corr_ref") for f in have) and any(f.startswith("cuda_ba_ref") for f in have):
        return True
    from torch.utils.cpp_extension import load
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    tmp = tempfile.mkdtemp(prefix="devo_ref_build_")
    try:
        _patched_copy(tmp)
        for name, srcs in (("cuda_corr_ref", ["correlation.cpp", "correlation_kernel.cu"]),
                           ("cuda_ba_ref", ["ba.cpp", "ba_cuda.cu"])):
            bdir = os.path.join(tmp, "build_" + name)
            os.makedirs(bdir, exist_ok=True)
            load(name=name, sources=[os.path.join(tmp, s) for s in srcs],
                 extra_cflags=["-O3"], extra_cuda_cflags=["-O3", "-lineinfo",
                 "-gencode=arch=compute_100a,code=sm_100a"],
                 build_directory=bdir, verbose=verbose, is_python_module=False)
            for f in os.listdir(bdir):
                if f.endswith(".so"):
                    shutil.copy(os.path.join(bdir, f), os.path.join(OUT, f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


PY_OUT = os.path.join(OUT, "devo_py")


def stage_python():
    """copy the reference's Python callers (verbatim) to oracle/_ref/devo_py/; returns the directory or None"""
    if not os.path.isdir(os.path.join(REF, "devo")):
        return PY_OUT if os.path.isdir(os.path.join(PY_OUT, "devo")) else None
    for sub in ("devo", "devo/altcorr", "devo/fastba", "devo/lietorch", "utils", "config"):
        src = os.path.join(REF, sub)
        dst = os.path.join(PY_OUT, sub)
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            if f.endswith((".py", ".yaml")) and os.path.isfile(os.path.join(src, f)):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return PY_OUT


def load_ref(name):
    """import oracle/_ref/<name>.so as a python module (GPU tests only)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("reference extensions built:" if ok else "reference not present", os.listdir(OUT) if ok else "")
