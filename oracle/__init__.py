"""CPU oracle for the DEVO update-operator hot path.

*** TEST INFRASTRUCTURE -- NOT PRODUCT CODE ***

Everything in this package is a CPU restatement (torch-CPU / numpy, fp64
capable) of the reference's algorithm for the hot path named in
BASELINE.json.  It exists to *check* the CUDA product in devo_b200/ and to
serve as the `cpu_baseline` / `--impl reference` arm of bench.py.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  Nothing under `devo_b200/` imports, links or executes anything
from here; the product fails loudly when its CUDA library is missing.

Parity pinning status:
  * lie.py              pinned by the reference's own test-suite properties
                        (devo/lietorch/run_tests.py: exp/log, inverse, adjoint
                        commutation, act-vs-matrix, finite-difference grads),
                        re-stated in tests/test_oracle_lie.py.  The reference
                        has no golden vectors for it and its C++ backend needs
                        Eigen (absent) => formulas follow include/*.h line by
                        line.
  * pops.py / ba.py     pinned against the reference's *own Python code*
                        (devo/projective_ops.py, devo/ba.py, devo/lietorch/*.py)
                        imported from /root/reference on top of oracle.lie as
                        the `lietorch_backends` module; fixtures in
                        tests/golden/ (tests/golden/make_golden.py).
  * corr.py / fastba.py / neighbors.py
                        the reference has no tests or fixtures for these
                        ("parity unpinned" by the reference's tests); they are
                        pinned on the GPU box against the reference's own CUDA
                        extensions compiled from /root/reference into
                        oracle/_ref (oracle/build_ref.py) -- see
                        tests/test_parity_vs_reference_ext.py.
"""
