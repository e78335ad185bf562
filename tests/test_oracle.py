"""Pins the port oracle (oracle/qgemv_oracle.c) to the reference:
(1) the reference's own known-answer tests for the path, restated with the same
    inputs and expected values; (2) golden vectors produced by the compiled
    reference (tests/golden/make_golden.py); (3) the compiled reference itself
    on fresh random inputs when oracle/_ref is available."""
import hashlib

import numpy as np
import pytest

from llm_inference_b200 import synth


# ---- (1) reference known-answer tests -------------------------------------

def test_ka_fp16_matvec(port):
    # ops_test.cpp:73-93 — 2x4 F16 weights {1..8} . 0.5 -> {5, 13}
    w = np.array([0x3c00, 0x4000, 0x4200, 0x4400, 0x4500, 0x4600, 0x4700, 0x4800], np.uint16)
    o = port.mat_vec_mul(synth.F16, w, np.full(4, 0.5, np.float32), 2, 4)
    assert abs(o[0] - 5.0) < 1e-3 and abs(o[1] - 13.0) < 1e-3


def test_ka_q4_k(port):
    # ops_test.cpp:138-171 — d=1, dmin=0, every 6-bit scale 1, nibbles 2, x=1 -> 512
    b = np.zeros(144, np.uint8)
    b[0:2] = np.frombuffer(np.float16(1.0).tobytes(), np.uint8)
    b[4:8] = 1
    b[12:16] = 1
    b[16:] = 0x22
    o = port.mat_vec_mul(synth.Q4_K, b, np.ones(256, np.float32), 1, 256)
    assert abs(o[0] - 512.0) < 1e-3


def test_ka_q6_k(port):
    # ops_test.cpp:173-202 — ql=0x11, qh=0xAA (q = 33-32 = 1), scales 1, d=1, x=1 -> 256
    b = np.zeros(210, np.uint8)
    b[0:128] = 0x11
    b[128:192] = 0xAA
    b[192:208] = 1
    b[208:210] = np.frombuffer(np.float16(1.0).tobytes(), np.uint8)
    o = port.mat_vec_mul(synth.Q6_K, b, np.ones(256, np.float32), 1, 256)
    assert abs(o[0] - 256.0) < 1e-3


def test_ka_q8_0(port):
    # ops_test.cpp:203-227 — d=1, qs=2, x=1 -> 64
    b = np.zeros(34, np.uint8)
    b[0:2] = np.frombuffer(np.float16(1.0).tobytes(), np.uint8)
    b[2:] = 2
    o = port.mat_vec_mul(synth.Q8_0, b, np.ones(32, np.float32), 1, 32)
    assert abs(o[0] - 64.0) < 1e-2


def test_ka_q5_0(port):
    # ops_test.cpp:229-257 — d=1, qs=0x11, qh=~0 (q = 17-16 = 1), x=1 -> 32
    b = np.zeros(22, np.uint8)
    b[0:2] = np.frombuffer(np.float16(1.0).tobytes(), np.uint8)
    b[2:6] = 0xFF
    b[6:] = 0x11
    o = port.mat_vec_mul(synth.Q5_0, b, np.ones(32, np.float32), 1, 32)
    assert abs(o[0] - 32.0) < 1e-3


def test_ka_q4_0_nibble_order(port):
    # gguf_test.cpp:150-280 — 4x32 Q4_0, byte patterns F0,E1,D2,C3, f16 scale bit
    # patterns 3800,3666,3333,3000, x = 1..32; expected = float dequant dot with
    # low nibble <-> element j, high nibble <-> element j+16; tolerance 15.0
    pats, scales = [0xF0, 0xE1, 0xD2, 0xC3], [0x3800, 0x3666, 0x3333, 0x3000]
    w = np.zeros((4, 18), np.uint8)
    for r in range(4):
        w[r, 0:2] = np.frombuffer(np.uint16(scales[r]).tobytes(), np.uint8)
        w[r, 2:] = pats[r]
    x = np.arange(1, 33, dtype=np.float32)
    o = port.mat_vec_mul(synth.Q4_0, w, x, 4, 32)
    for r in range(4):
        d = float(np.array([scales[r]], np.uint16).view(np.float16)[0])
        lo, hi = (pats[r] & 0xF) - 8, (pats[r] >> 4) - 8
        exp = d * (lo * x[:16].sum() + hi * x[16:].sum())
        assert abs(o[r] - exp) < 15.0
        assert abs(o[r] - exp) < 0.02 * abs(exp) + 0.5  # and in fact much closer


def test_ka_f16_table(port):
    # gguf_test.cpp:63-83 — table values incl. subnormal 773 -> 4.6e-5
    v = port.f16_to_f32(np.array([0x3c00, 0xc000, 773, 0], np.uint16))
    assert v[0] == 1.0 and v[1] == -2.0 and v[3] == 0.0
    assert abs(v[2] - 773 * 2.0 ** -24) < 1e-12 and abs(v[2] - 4.6e-5) < 1e-6


def test_dispatcher_rejects_f16_and_f32(port):
    # ops.cpp:933-956: F16 / F32 are not accepted by mat_vec_mul
    o = np.zeros(1, np.float32)
    import ctypes as C
    rc = port.L.orc_mat_vec_mul(1, o.ctypes.data_as(C.POINTER(C.c_float)), None, None, 1, 32)
    assert rc == 1
    assert port.L.orc_mat_vec_mul(0, o.ctypes.data_as(C.POINTER(C.c_float)), None, None, 1, 32) == 1


# ---- (2) golden vectors from the compiled reference -------------------------

def test_golden_f16_conversions(port, golden):
    table = port.f16_to_f32(np.arange(65536, dtype=np.uint16))
    assert hashlib.sha256(table.tobytes()).digest() == golden["f16_table_sha256"].tobytes()
    got = port.f16_to_f32(golden["f16_table_sample_codes"])
    assert np.array_equal(got.view(np.uint32), golden["f16_table_sample_vals"].view(np.uint32))
    assert np.array_equal(port.f32_to_f16(golden["f32_to_f16_in"]), golden["f32_to_f16_out"])


def test_golden_quantizers_bit_exact(port, golden):
    x = golden["q8_in"]
    assert np.array_equal(port.quantize_row_q8_0(x), golden["q8_0_out"])
    assert np.array_equal(port.quantize_row_q8_k(x), golden["q8_k_out"])


def _golden_cases(golden):
    i = 0
    while f"mv{i}_meta" in golden:
        t, k, n = (int(v) for v in golden[f"mv{i}_meta"])
        yield t, k, n, golden[f"mv{i}_w"], golden[f"mv{i}_x"], golden[f"mv{i}_o"]
        i += 1


def test_golden_matvec_bit_exact(port, golden):
    # the port reproduces the compiled reference's fp32 results bit for bit
    n_cases = 0
    for t, k, n, w, x, o_ref in _golden_cases(golden):
        o = port.mat_vec_mul(t, w, x, n, k)
        assert np.array_equal(o.view(np.uint32), o_ref.view(np.uint32)), (synth.TYPE_NAMES[t], k, n)
        n_cases += 1
    assert n_cases >= 12


def test_golden_row_dequantizers(port, golden):
    for t in (synth.Q8_0, synth.Q5_0, synth.Q4_K, synth.Q6_K):
        out = port.dequantize_row(t, golden[f"deq{t}_row"], 512)
        assert np.array_equal(out.view(np.uint32), golden[f"deq{t}_out"].view(np.uint32)), synth.TYPE_NAMES[t]


# ---- (3) the compiled reference itself --------------------------------------

SHAPES = [(synth.Q4_0, 1152, 96), (synth.Q4_0, 6912, 40), (synth.Q8_0, 3840, 48), (synth.Q5_0, 2560, 32),
          (synth.Q4_K, 2560, 64), (synth.Q4_K, 10240, 16), (synth.Q6_K, 2560, 64), (synth.Q6_K, 10240, 16),
          (synth.BF16, 2560, 32), (synth.F16, 5376, 48), (synth.F16, 1155, 9)]


@pytest.mark.parametrize("t,k,n", SHAPES, ids=[f"{synth.TYPE_NAMES[t]}-{k}x{n}" for t, k, n in SHAPES])
def test_port_matches_compiled_reference(port, ref, t, k, n):
    rng = np.random.default_rng(k * 31 + n + t)
    if t == synth.F16 and k % 32:
        w = rng.standard_normal((n, k)).astype(np.float16).view(np.uint8).ravel()
    else:
        w = synth.random_blocks(t, n, k, seed=k + n + t)
    x = rng.standard_normal(k).astype(np.float32)
    a, b = port.mat_vec_mul(t, w, x, n, k), ref.mat_vec_mul(t, w, x, n, k)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_port_quantizers_match_compiled_reference(port, ref):
    rng = np.random.default_rng(5)
    for scale in (1.0, 1e-6, 300.0):
        x = (rng.standard_normal(21504) * scale).astype(np.float32)
        assert np.array_equal(port.quantize_row_q8_0(x), ref.quantize_row_q8_0(x))
        assert np.array_equal(port.quantize_row_q8_k(x), ref.quantize_row_q8_k(x))


def test_port_conversions_match_compiled_reference(port, ref):
    codes = np.arange(65536, dtype=np.uint16)
    assert np.array_equal(port.f16_to_f32(codes).view(np.uint32), ref.f16_to_f32(codes).view(np.uint32))
    rng = np.random.default_rng(9)
    f = (rng.standard_normal(50000) * 10.0 ** rng.uniform(-10, 6, 50000)).astype(np.float32)
    assert np.array_equal(port.f32_to_f16(f), ref.f32_to_f16(f))


def test_block_dots_are_consistent(port):
    # the integer dots the GPU must reproduce: recompute o from them in double
    k, n = 1152, 8
    w = synth.random_blocks(synth.Q4_0, n, k, seed=3)
    x = np.random.default_rng(3).standard_normal(k).astype(np.float32)
    o, dots = port.mat_vec_mul(synth.Q4_0, w, x, n, k, want_dots=True)
    xq = port.quantize_row_q8_0(x).reshape(-1, 34)
    dx = xq[:, :2].copy().view(np.float16).astype(np.float64).ravel()
    dw = w.reshape(n, k // 32, 18)[:, :, :2].copy().view(np.float16).astype(np.float64).reshape(n, -1)
    exact = (dots.reshape(n, -1) * dw * dx).sum(axis=1)
    assert np.abs(o - exact).max() <= 1e-5 * np.abs(exact).max()


@pytest.mark.parametrize("t,k,n", [(synth.Q4_0, 1152, 64), (synth.Q4_0, 1184, 33), (synth.Q8_0, 2592, 17),
                                   (synth.Q4_0, 96, 8), (synth.Q8_0, 32, 8), (synth.Q4_0, 5376, 24),
                                   (synth.Q4_K, 1280, 19), (synth.Q6_K, 2560, 19), (synth.Q5_0, 1184, 19),
                                   (synth.F16, 1160, 19), (synth.BF16, 1152, 19)])
def test_canonical_order_restatement_is_within_the_bound_of_the_reference_order(port, t, k, n):
    """The device's summation order restated on the CPU (orc_gemv_*_canonical: chunks of 16 blocks, four chains,
    (s0+s1)+(s2+s3), chunks left to right — what the GPU kernels are checked against BIT FOR BIT in
    tests/test_gemv_gpu.py) uses the reference's per-block terms, so it differs from the reference-order port only
    by fp32 summation order: <= 1e-5 of max|o| (north_star's bar), and exactly equal when a row is one block."""
    w = synth.random_blocks(t, n, k, seed=k + n)
    x = np.random.default_rng(k).standard_normal(k).astype(np.float32)
    ref_order = port.mat_vec_mul(t, w, x, n, k)
    canonical = port.mat_vec_mul_canonical(t, w, x, n, k)
    scale = float(np.abs(ref_order).max())
    assert float(np.abs(canonical - ref_order).max()) <= 1e-5 * scale
    if k == 32 and t == synth.Q8_0:  # one block: (dot*dw)*dx, a single rounding sequence either way
        assert np.array_equal(canonical.view(np.uint32), ref_order.view(np.uint32))
