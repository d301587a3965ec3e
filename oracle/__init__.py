"""CPU oracle for the AdapterSIS hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-CPU restatement (fp32 / fp64, no custom kernels)
of the reference's mathematics for the hot path named in BASELINE.json:
multi-scale deformable attention, the injector / extractor adapter blocks, the
DINOv2 transformer block stack, and the interleaved encoder of train.py.

Rules (enforced by tests/test_oracle_golden.py::test_oracle_is_test_infrastructure_only):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``adaptersis_b200/`` imports it -- the product path is the
    CUDA library and fails loudly when that library is missing.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against fixtures generated *by importing the reference's
own modules* in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/*.pt``); ``tests/test_oracle_golden.py`` checks every function
here against those fixtures.

Every function is written functionally over a ``state_dict`` carrying the
reference's parameter names, which doubles as the state_dict-compat proof.
The functions are dtype- and device-generic (fp32 or fp64; the tensors decide), so the ``-m gpu``
tests can also run the oracle ON the GPU in fp64 at BASELINE.json's full sizes (ViT-L/14, 588^2).
"""
from . import msda, layers, vit, adapter, encoder  # noqa: F401
