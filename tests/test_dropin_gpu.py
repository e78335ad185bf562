"""The reference's OWN code on the CUDA path, through the unchanged ops.h
boundary: (1) its three gtest files, compiled unmodified and linked against
host/ops_cuda.cpp instead of the matmul half of ops.cpp (oracle/Makefile
`dropin`), must pass on the B200 — including ModelTest.ForwardPass' golden
logits (model_test.cpp:409-460); (2) its Model::forward run on both builds over
a synthetic Gemma-3 GGUF must agree (same call sites model.cpp:557,754,784,803,
875,877,909,1000).  Needs the prebuilt oracle/_ref (it travels to the GPU box)."""
import subprocess

import numpy as np
import pytest

from llm_inference_b200 import synth
from oracle import binding

pytestmark = pytest.mark.gpu
REF = binding.HERE / "_ref"


def _need(name):
    p = REF / name
    if not p.exists():
        pytest.skip(f"{p} was not built (reference sources not mounted at build time)")
    return p


@pytest.mark.parametrize("which", ["ops", "gguf", "model"])
def test_reference_gtests_pass_on_the_cuda_dropin(which):
    exe = _need(f"dropin_{which}_test")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout
    if which == "model":
        assert "[       OK ] ModelTest.ForwardPass" in r.stdout


def test_reference_model_forward_cpu_vs_cuda_dropin():
    _need("libref.so")
    _need("libdropin.so")
    dims = synth.GemmaDims("tiny", 2, 256, 512, 4, 2, 64, 320)
    img = synth.build_gemma3_gguf(dims, synth.Q4_0, synth.F16, seed=11)
    cpu = binding.Ref("libref.so", n_threads=2).model(img)
    gpu = binding.Ref("libdropin.so", n_threads=1).model(img)
    toks = [5, 17, 200, 3]
    a, b = cpu.forward(toks, 0), gpu.forward(toks, 0)
    assert np.isfinite(a).all() and a.shape == (320,)
    assert np.abs(a - b).max() <= 2e-4 * np.abs(a).max()
    assert int(a.argmax()) == int(b.argmax())
    pos = len(toks)
    for _ in range(8):  # greedy decode with the KV cache, token-identical
        t = int(a.argmax())
        a, b = cpu.forward([t], pos), gpu.forward([t], pos)
        assert int(a.argmax()) == int(b.argmax())
        assert np.abs(a - b).max() <= 5e-4 * np.abs(a).max()
        pos += 1
