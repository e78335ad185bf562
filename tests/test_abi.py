"""C-ABI surface: the library builds for sm_100a, loads without a GPU, exports
every symbol include/llmi_cuda.h declares, and refuses to compute without a
device (no CPU fallback)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from llm_inference_b200 import _build, _lib


def test_library_builds_and_exports_every_declared_symbol():
    lib = _lib.load(build=True)
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in llmi_cuda.h but not exported"
    # and the Python signature table covers the header exactly
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.llmi_abi_version() == 1


def test_sass_is_sm_100a_with_bulk_copy_and_dp4a():
    _lib.load(build=True)
    out = subprocess.run(["cuobjdump", "-sass", str(_build.LIB)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
    assert "UBLKCP" in out.stdout, "activation staging must be a bulk async (TMA) copy"
    assert "IDP.4A" in out.stdout, "block dots must be dp4a"
    assert "LDG.E.NA.128" in out.stdout or "LDG.E.128" in out.stdout
    # the prefill kernel is on the 5th-generation tensor cores: int8 UMMA, tensor-memory loads, TMA tensor copy
    for mnemonic in ("UTCIMMA", "LDTM", "UTCBAR", "UTMALDG"):
        assert mnemonic in out.stdout, f"{mnemonic} missing: the tcgen05 prefill kernel did not compile for sm_100a"


def test_row_bytes_matches_reference_block_sizes():
    lib = _lib.load()
    # ops.h:11-31,89-92 struct sizes: 18/34/22 per 32, 144/210 per 256
    assert lib.llmi_row_bytes(2, 1152) == 36 * 18
    assert lib.llmi_row_bytes(8, 1152) == 36 * 34
    assert lib.llmi_row_bytes(6, 1152) == 36 * 22
    assert lib.llmi_row_bytes(12, 2560) == 10 * 144
    assert lib.llmi_row_bytes(14, 2560) == 10 * 210
    assert lib.llmi_row_bytes(1, 7) == 14 and lib.llmi_row_bytes(30, 7) == 14
    assert lib.llmi_row_bytes(12, 1152) == 0  # K % 256 != 0 (SURVEY hard part 8)
    assert lib.llmi_row_bytes(3, 1152) == 0   # Q4_1 unsupported


def test_calls_before_init_fail_with_state_error():
    lib = _lib.load()
    h = C.c_void_p()
    w = np.zeros(18, np.uint8)
    rc = lib.llmi_weight_upload(w.ctypes.data, 2, 32, 1, 0, 1, C.byref(h))
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; init state is shared with other tests")
    assert rc == 5 and b"llmi_init" in lib.llmi_last_error()


def test_init_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    rc = lib.llmi_init(0)
    assert rc == 4
    assert b"no CPU fallback" in lib.llmi_last_error()
    from llm_inference_b200 import ops
    with pytest.raises(RuntimeError):
        ops.init_ops(1)
    with pytest.raises(RuntimeError):
        ops.quantize_row_q8_0(np.zeros(32, np.float32))
