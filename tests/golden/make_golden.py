"""Generates tests/golden/qgemv_golden.npz from the COMPILED REFERENCE
(oracle/_ref/libref.so = /root/reference/{ops,gguf,model}.cpp built in place by
oracle/Makefile).  Run in the dev container only:

    python tests/golden/make_golden.py

The committed .npz pins the port oracle (and through it the CUDA path) to the
reference's actual outputs on machines where /root/reference does not exist.
Inputs are seeded numpy draws (llm_inference_b200.synth.random_blocks), stored
in the file next to the reference outputs so no regeneration is needed to test.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))

from llm_inference_b200 import synth  # noqa: E402
from oracle.binding import Ref, ensure_ref  # noqa: E402

CASES = [  # (ggml_type, K, N)
    (synth.Q4_0, 1152, 24), (synth.Q4_0, 32, 3), (synth.Q8_0, 1152, 16), (synth.Q5_0, 1152, 16),
    (synth.Q4_K, 2560, 16), (synth.Q4_K, 256, 2), (synth.Q6_K, 2560, 16), (synth.Q6_K, 256, 2),
    (synth.BF16, 1152, 16), (synth.F16, 1152, 16), (synth.F16, 1150, 5), (synth.F16, 7, 3),
]


def main() -> None:
    assert ensure_ref(), "reference sources not available"
    R = Ref(n_threads=1)
    out = {}
    rng = np.random.default_rng(20261018)
    # scalar conversions
    codes = np.arange(65536, dtype=np.uint16)
    table = R.f16_to_f32(codes)
    out["f16_table_sha256"] = np.frombuffer(hashlib.sha256(table.tobytes()).digest(), np.uint8)
    out["f16_table_sample_codes"] = np.array([0, 1, 773, 0x03ff, 0x0400, 0x3c00, 0x7bff, 0x7c00, 0x7e00, 0x8001,
                                              0xfbff, 0xfc00], np.uint16)
    out["f16_table_sample_vals"] = R.f16_to_f32(out["f16_table_sample_codes"])
    f = np.concatenate([
        (rng.standard_normal(4000) * 10.0 ** rng.uniform(-9, 6, 4000)).astype(np.float32),
        np.array([0, -0.0, 65504, 65519.996, 65520, 1e10, -1e10, np.inf, -np.inf, 2.0 ** -25, 2.0 ** -24,
                  2.0 ** -14, 5.96e-8, 2.98e-8, 2.9802322e-08, 2.9802326e-08, 6.1e-5, 6.097e-5], np.float32)])
    out["f32_to_f16_in"] = f
    out["f32_to_f16_out"] = R.f32_to_f16(f)
    # quantizers (incl. an all-zero block, a tie in |max| with opposite signs, tiny and huge values)
    x = rng.standard_normal(1024).astype(np.float32)
    x[32:64] = 0.0
    x[256:512] = 0.0
    x[512] = 3.5
    x[513] = -3.5
    x[600] = -3.5
    x[768:800] *= 1e-30
    x[800:832] *= 1e20
    out["q8_in"] = x
    out["q8_0_out"] = R.quantize_row_q8_0(x)
    out["q8_k_out"] = R.quantize_row_q8_k(x)
    # mat-vecs
    for i, (t, k, n) in enumerate(CASES):
        if t in (synth.F16,) and k % 32:
            w = rng.standard_normal((n, k)).astype(np.float16).view(np.uint8).ravel()
        else:
            w = synth.random_blocks(t, n, k, seed=100 + i)
        xv = rng.standard_normal(k).astype(np.float32)
        o = R.mat_vec_mul(t, w, xv, n, k)
        out[f"mv{i}_meta"] = np.array([t, k, n], np.int64)
        out[f"mv{i}_w"] = w
        out[f"mv{i}_x"] = xv
        out[f"mv{i}_o"] = o
    # row dequantizers
    for t in (synth.Q8_0, synth.Q5_0, synth.Q4_K, synth.Q6_K):
        row = synth.random_blocks(t, 1, 512, seed=900 + t)
        out[f"deq{t}_row"] = row
        out[f"deq{t}_out"] = R.dequantize_row(t, row, 512)
    path = Path(__file__).with_name("qgemv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes,", len(CASES), "mat-vec cases")


if __name__ == "__main__":
    main()
