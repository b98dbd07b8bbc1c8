"""CPU: the C-ABI shared library loads and exports every symbol include/devo_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from devo_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from devo_b200 import build
        build.build()
    hdr = open(os.path.join(ROOT, "include", "devo_b200.h")).read()
    declared = set(re.findall(r"\b(devo_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert len(declared) >= 37
    h = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(h, sym), sym
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    assert L.devo_abi_version() == 1
    assert L.devo_graph_plan_workspace(6144) > 0 and L.devo_ba_workspace(6144, 7) > L.devo_graph_plan_workspace(6144)
    assert L.devo_launch_count() == 0 or L.devo_launch_count() > 0


def test_sm100a_cubin_with_blackwell_instructions():
    """the built library carries an sm_100a cubin whose SASS shows TMA and tcgen05 (B200_PROFILING.md)"""
    import shutil
    import subprocess
    from devo_b200 import _lib
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        return
    out = subprocess.run([cu, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic
    # the fused update operator: tcgen05 MMA, TMEM load AND store, TMA, cluster barrier, tcgen05.commit
    i = out.find("gru_mma_kernel")
    assert i >= 0
    j = out.find("Function :", i + 10)
    gru = out[i:j if j > 0 else len(out)]
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UCGABAR_ARV", "UTCBAR"):
        assert mnemonic in gru, mnemonic


def test_integration_doc_maps_every_symbol():
    """INTEGRATION.md names every exported entry point (lie functions through the `devo_lie_*` row)"""
    hdr = open(os.path.join(ROOT, "include", "devo_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(devo_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert "devo_lie_*" in doc
    missing = sorted(s for s in declared if s not in doc and not s.startswith("devo_lie_"))
    assert not missing, missing


def test_hot_kernels_do_not_spill():
    """static resources of the per-iteration kernels (cuobjdump --dump-resource-usage): no stack frame (= no register
    spills) in the tensor-core kernels the bench runs, and the fused update operator fits 448 threads per SM"""
    import shutil
    import subprocess
    from devo_b200 import _lib
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        return
    txt = subprocess.run([cu, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    res = {}
    fn = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and fn:
            res[fn] = (int(m.group(1)), int(m.group(2)))
            fn = None

    def find(*parts):
        hits = [v for k, v in res.items() if all(p in k for p in parts)]
        assert hits, parts
        return hits

    for reg, stack in find("gru_mma_kernel", "6__half"):
        assert stack == 0 and reg <= 128, (reg, stack)       # 448 threads x 128 registers <= 64 K per SM
    for reg, stack in find("corr_fast_kernel"):
        assert stack == 0, (reg, stack)
    for reg, stack in find("segment_softmax_sum_kernel") + find("transform_kernel"):
        assert stack == 0, (reg, stack)
