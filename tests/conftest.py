"""Test configuration.

* ``-m "not gpu"``: oracle vs golden vectors / reference known answers, host
  logic, ABI surface.  Runs anywhere (no GPU, no /root/reference needed).
* ``-m gpu``: parity tests proper — the CUDA path, called through the C ABI,
  against the oracle on identical seeded inputs.  Never reads /root/reference.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    from oracle.binding import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libref.so); built here when the
    reference sources are mounted, otherwise used prebuilt, otherwise skipped."""
    from oracle import binding
    if not binding.ref_available():
        try:
            binding.ensure_ref()
        except Exception as e:  # pragma: no cover
            pytest.skip(f"oracle/_ref not buildable here: {e}")
    if not binding.ref_available():
        pytest.skip("oracle/_ref/libref.so not present (reference sources not mounted)")
    return binding.Ref(n_threads=1)


@pytest.fixture(scope="session")
def golden():
    return np.load(REPO / "tests" / "golden" / "qgemv_golden.npz")


@pytest.fixture(scope="session")
def gpu_ops():
    """The product's ops module, initialised on cuda:0.  Fails (does not skip)
    if the CUDA extension cannot be used: there is no fallback to hide behind."""
    from llm_inference_b200 import ops
    ops.init_ops(1, device=0)
    return ops
