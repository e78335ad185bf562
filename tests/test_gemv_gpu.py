"""Parity of the CUDA path (through the C ABI) with the oracle.

Bars (north_star / SURVEY §7 hard part 1):
  * Q8_0 / Q8_K activation quantization: bit-exact;
  * per-block integer dots: bit-exact;
  * fp32 outputs: |o_gpu - o_ref| <= 1e-5 * max_r |o_ref|  (summation order is
    the only difference; per-element relative error is meaningless under
    cancellation).
The checker is the port oracle (bit-exact with the compiled reference, see
test_oracle.py) and, when oracle/_ref travelled to this box, the compiled
reference itself."""
import numpy as np
import pytest

from llm_inference_b200 import synth
from llm_inference_b200.synth import BF16, F16, Q4_0, Q4_K, Q5_0, Q6_K, Q8_0

pytestmark = pytest.mark.gpu

TOL = 1e-5  # relative to max |o_ref| over the rows


def _check(o, o_ref, what=""):
    scale = float(np.abs(o_ref).max())
    err = float(np.abs(o - o_ref).max())
    assert err <= TOL * scale, f"{what}: max err {err:.3e} > {TOL}*{scale:.3e}"
    return err / scale if scale else 0.0


def _weights(t, n, k, seed):
    return synth.random_blocks(t, n, k, seed=seed)


def _run(ops, t, w, x, n, k, shape=(0, 0), row_range=None, ring=(0, 0, 0, 0)):
    ops.set_gemv_shape(*shape)
    ops.set_gemv_ring(*ring)
    rb, re = row_range or (0, n)
    dw = ops.DeviceWeight(w, t, k, n, rb, re)
    dx = ops.DeviceVector(k, x)
    do = ops.DeviceVector(n, np.full(n, np.nan, np.float32))
    act = ops.Activation(k)
    ops.mat_vec_mul_dev(dw, dx, act, do)
    ops.device_sync()
    o = do.get()
    extra = {}
    if t in (Q4_0, Q8_0, Q4_K, Q6_K):
        extra["dots"] = ops.block_dots(dw, act)
        extra["xq"] = act.export_q8_0(k) if t in (Q4_0, Q8_0) else act.export_q8_k(k)
    for h in (dw, dx, do, act):
        h.close()
    ops.set_gemv_shape(0, 0)
    ops.set_gemv_ring(0, 0, 0, 0)
    return o, extra


# ---------------------------------------------------------------- quantizers

@pytest.mark.parametrize("k", [32, 1152, 2560, 21504])
def test_quantize_q8_0_bit_exact(gpu_ops, port, k):
    rng = np.random.default_rng(k)
    for scale in (1.0, 1e-4, 300.0):
        x = (rng.standard_normal(k) * scale).astype(np.float32)
        if k >= 64:
            x[32:64] = 0.0  # an all-zero block (d == 0 -> id == 0)
        got = gpu_ops.quantize_row_q8_0(x)
        assert np.array_equal(got, port.quantize_row_q8_0(x))


@pytest.mark.parametrize("k", [256, 2560, 10240])
def test_quantize_q8_k_bit_exact(gpu_ops, port, k):
    rng = np.random.default_rng(k + 1)
    for scale in (1.0, 1e-4, 300.0):
        x = (rng.standard_normal(k) * scale).astype(np.float32)
        x[0:256] = 0.0          # all-zero super-block
        if k > 256:
            x[256 + 7] = 9.0    # tie in |max|: first occurrence (positive) must win
            x[256 + 99] = -9.0
        got = gpu_ops.quantize_row_q8_k(x)
        assert np.array_equal(got, port.quantize_row_q8_k(x))


def test_quantizers_match_golden(gpu_ops, golden):
    x = golden["q8_in"]
    assert np.array_equal(gpu_ops.quantize_row_q8_0(x), golden["q8_0_out"])
    assert np.array_equal(gpu_ops.quantize_row_q8_k(x), golden["q8_k_out"])


# ------------------------------------------------------------------- mat-vec

SHAPES = [
    # BASELINE.json config 1: gemma-3-1b FFN gate shape
    (Q4_0, 1152, 6912),
    # other gemma-3 shapes (1b/4b/27b): short-wide, K-split paths, odd slab counts
    (Q4_0, 1152, 256), (Q4_0, 6912, 1152), (Q4_0, 5376, 4096), (Q4_0, 21504, 1348),
    (Q4_0, 32, 5), (Q4_0, 96, 17),
    (Q8_0, 3840, 2048), (Q8_0, 15360, 520), (Q8_0, 32, 9),
    (Q5_0, 2560, 1024), (Q5_0, 21504, 40), (Q5_0, 64, 3),
    (Q4_K, 2560, 2048), (Q4_K, 10240, 2560), (Q4_K, 256, 11),
    (Q6_K, 2560, 1024), (Q6_K, 10240, 2560), (Q6_K, 256, 11),
    (BF16, 2560, 512), (BF16, 1150, 13),
    (F16, 1152, 4096), (F16, 5376, 1000), (F16, 1150, 13), (F16, 7, 3),
]


@pytest.mark.parametrize("t,k,n", SHAPES, ids=[f"{synth.TYPE_NAMES[t]}-{k}x{n}" for t, k, n in SHAPES])
def test_matvec_parity(gpu_ops, port, t, k, n):
    rng = np.random.default_rng(1000 * t + k + n)
    if t in (F16, BF16) and k % 8:
        w = (rng.standard_normal((n, k)) * 0.05).astype(np.float16).view(np.uint8).ravel() if t == F16 else \
            ((rng.standard_normal((n, k)) * 0.05).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16) \
            .view(np.uint8).ravel()
    else:
        w = _weights(t, n, k, seed=t + k + n)
    x = rng.standard_normal(k).astype(np.float32)
    o, extra = _run(gpu_ops, t, w, x, n, k)
    if t in (Q4_0, Q8_0, Q4_K, Q6_K):
        o_ref, dots_ref = port.mat_vec_mul(t, w, x, n, k, want_dots=True)
        xq_ref = port.quantize_row_q8_0(x) if t in (Q4_0, Q8_0) else port.quantize_row_q8_k(x)
        assert np.array_equal(extra["xq"], xq_ref), "activation quants must be bit-exact"
        assert np.array_equal(extra["dots"], dots_ref), "integer block dots must be bit-exact"
    else:
        o_ref = port.mat_vec_mul(t, w, x, n, k)
    assert not np.isnan(o).any()
    _check(o, o_ref, f"{synth.TYPE_NAMES[t]} {k}->{n}")


@pytest.mark.parametrize("t,k,n", [(Q4_0, 1152, 1030), (Q8_0, 3840, 264), (Q4_K, 2560, 136), (Q6_K, 2560, 136),
                                   (F16, 1152, 520), (Q5_0, 1152, 72), (BF16, 1152, 72)],
                         ids=lambda v: str(v))
def test_result_is_independent_of_the_grid(gpu_ops, port, t, k, n):
    # canonical chunked summation: every CTA shape gives the same bits
    w = _weights(t, n, k, seed=77 + t)
    x = np.random.default_rng(t).standard_normal(k).astype(np.float32)
    base, _ = _run(gpu_ops, t, w, x, n, k, shape=(4, 1))
    _check(base, port.mat_vec_mul(t, w, x, n, k), "shape=(4,1)")
    for shape in ((4, 4), (8, 1), (8, 3), (16, 1), (16, 7), (0, 0)):
        o, _ = _run(gpu_ops, t, w, x, n, k, shape=shape)
        assert np.array_equal(o.view(np.uint32), base.view(np.uint32)), f"shape={shape}"


@pytest.mark.parametrize("t,k,n", [(Q4_0, 1152, 6912), (Q4_0, 6912, 1160), (Q4_0, 21504, 520), (Q8_0, 3840, 1544),
                                   (Q4_K, 2560, 2056), (Q6_K, 2560, 1032), (F16, 1152, 3080), (Q5_0, 1152, 1032),
                                   (BF16, 1184, 520), (Q4_0, 1184, 203), (Q4_0, 32, 8), (Q8_0, 512, 40)],
                         ids=lambda v: str(v))
def test_persistent_ring_kernel_is_bitwise_the_slab_kernel(gpu_ops, port, t, k, n):
    """gemv_ring_kernel (persistent CTAs, per-warp bulk-copy rings, item-granular CTA ranges with flagged hand-over of
    the chunk partials of split slabs) against gemv_slab_kernel: same items, same canonical order => same bits, for
    every grid size and ring depth, including slabs split over two and over three CTAs (K = 21504: 42 chunks)."""
    w = _weights(t, n, k, seed=11 + t + n)
    x = np.random.default_rng(t + k).standard_normal(k).astype(np.float32)
    base, _ = _run(gpu_ops, t, w, x, n, k, ring=(1, 0, 0, 0))
    _check(base, port.mat_vec_mul(t, w, x, n, k), "slab kernel")
    for ring in ((2, 0, 0, 0), (2, 1, 2, 8), (2, 3, 2, 8), (2, 4, 4, 8), (2, 2, 3, 16), (2, 1, 2, 16)):
        for _ in range(2):  # twice: the hand-over words must be back to zero after a launch
            o, _e = _run(gpu_ops, t, w, x, n, k, ring=ring)
            assert np.array_equal(o.view(np.uint32), base.view(np.uint32)), f"ring={ring}"
    # a row shard (ragged slab count) through the ring kernel
    rb, re = 8 * ((n // 8) // 3), n
    o, _e = _run(gpu_ops, t, w, x, n, k, row_range=(rb, re), ring=(2, 2, 3, 8))
    assert np.array_equal(o[rb:re].view(np.uint32), base[rb:re].view(np.uint32))


def test_persistent_ring_kernel_batched_launch(gpu_ops):
    """q/k/v as one persistent grid: CTA ranges cross the matrix boundaries."""
    ops = gpu_ops
    k = 2560
    x = np.random.default_rng(9).standard_normal(k).astype(np.float32)
    dx, act = ops.DeviceVector(k, x), ops.Activation(k)
    for t in (Q4_0, Q4_K, Q8_0):
        shapes = [2048, 1024, 1032]
        dws = [ops.DeviceWeight(_weights(t, n, k, seed=n), t, k, n) for n in shapes]
        outs = {m: [ops.DeviceVector(n, np.full(n, np.nan, np.float32)) for n in shapes] for m in (1, 2)}
        act.prepare(dws[0], dx)
        for m in (1, 2):
            ops.set_gemv_ring(m, 0, 0, 0)
            ops.gemv_batch(dws, act, outs[m])
        ops.set_gemv_ring(0, 0, 0, 0)
        ops.device_sync()
        for a, b in zip(outs[1], outs[2]):
            assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32))
        for h in dws + outs[1] + outs[2]:
            h.close()


def test_matvec_matches_golden_from_compiled_reference(gpu_ops, golden):
    i = 0
    while f"mv{i}_meta" in golden:
        t, k, n = (int(v) for v in golden[f"mv{i}_meta"])
        o, _ = _run(gpu_ops, t, golden[f"mv{i}_w"], golden[f"mv{i}_x"], n, k)
        _check(o, golden[f"mv{i}_o"], f"golden case {i}")
        i += 1
    assert i >= 12


def test_matvec_against_compiled_reference_if_present(gpu_ops):
    from oracle import binding
    if not binding.ref_available():
        pytest.skip("oracle/_ref did not travel to this box")
    R = binding.Ref(n_threads=4)
    rng = np.random.default_rng(42)
    for t, k, n in [(Q4_0, 1152, 6912), (Q8_0, 3840, 512), (Q4_K, 2560, 512), (Q6_K, 2560, 512), (F16, 1152, 2048)]:
        w = _weights(t, n, k, seed=5 + t)
        x = rng.standard_normal(k).astype(np.float32)
        o, _ = _run(gpu_ops, t, w, x, n, k)
        _check(o, R.mat_vec_mul(t, w, x, n, k), f"vs compiled reference {synth.TYPE_NAMES[t]}")


def test_row_sharded_handles_reproduce_the_full_result_bitwise(gpu_ops):
    # the multi-GPU image of ops.cpp:439-448: a row is computed start-to-finish
    # by one device, so any row partition gives bit-identical rows
    t, k, n = Q4_0, 5376, 4096
    w = _weights(t, n, k, seed=9)
    x = np.random.default_rng(9).standard_normal(k).astype(np.float32)
    full, _ = _run(gpu_ops, t, w, x, n, k)
    for parts in (2, 4, 8):
        got = np.zeros(n, np.float32)
        for r in range(parts):
            rb, re = r * n // parts, (r + 1) * n // parts
            o, _ = _run(gpu_ops, t, w, x, n, k, row_range=(rb, re))
            got[rb:re] = o[rb:re]
        assert np.array_equal(got.view(np.uint32), full.view(np.uint32)), f"{parts}-way shard"


def test_full_size_properties_27b_gate(gpu_ops, port):
    # BASELINE full size (27b gate 5376->21504, 65 MB): size-independent checks
    t, k, n = Q4_0, 5376, 21504
    w = _weights(t, n, k, seed=27)
    rng = np.random.default_rng(27)
    x = rng.standard_normal(k).astype(np.float32)
    o, extra = _run(gpu_ops, t, w, x, n, k)
    # (a) a random sample of rows against the oracle
    rows = rng.choice(n, 256, replace=False)
    rb = k // 32 * 18
    w_rows = np.concatenate([w[r * rb:(r + 1) * rb] for r in rows])
    o_ref = port.mat_vec_mul(t, w_rows, x, len(rows), k)
    scale = float(np.abs(o_ref).max())
    assert np.abs(o[rows] - o_ref).max() <= TOL * scale
    # (b) checksum of checksums: sum of all integer block dots per block column
    #     equals the dot of the column-summed nibbles with the quantized x
    dots = extra["dots"].reshape(n, k // 32).astype(np.int64)
    blocks = w.reshape(n, k // 32, 18)[:, :, 2:]
    lo = (blocks & 0x0F).astype(np.int64) - 8
    hi = (blocks >> 4).astype(np.int64) - 8
    colsum = np.concatenate([lo.sum(axis=0), hi.sum(axis=0)], axis=1)  # [nb, 32]
    xq = extra["xq"].reshape(-1, 34)[:, 2:].view(np.int8).astype(np.int64)
    assert np.array_equal(dots.sum(axis=0), (colsum * xq).sum(axis=1))
    # (c) homogeneity: scaling x by a power of two scales o exactly (quants unchanged)
    o2, _ = _run(gpu_ops, t, w, (x * 4.0).astype(np.float32), n, k)
    assert np.array_equal(o2, o * 4.0)


def test_zero_and_empty_inputs(gpu_ops, port):
    # x == 0 -> every output exactly 0 (d == 0 blocks)
    for t, k in ((Q4_0, 1152), (Q4_K, 2560), (Q6_K, 2560), (F16, 1152)):
        n = 24
        w = _weights(t, n, k, seed=1)
        o, _ = _run(gpu_ops, t, w, np.zeros(k, np.float32), n, k)
        assert np.array_equal(o, np.zeros(n, np.float32))
    # empty row range: nothing is written
    w = _weights(Q4_0, 16, 64, seed=2)
    dw = gpu_ops.DeviceWeight(w, Q4_0, 64, 16, 5, 5)
    dx = gpu_ops.DeviceVector(64, np.ones(64, np.float32))
    do = gpu_ops.DeviceVector(16, np.full(16, 7.0, np.float32))
    act = gpu_ops.Activation(64)
    gpu_ops.mat_vec_mul_dev(dw, dx, act, do)
    gpu_ops.device_sync()
    assert np.array_equal(do.get(), np.full(16, 7.0, np.float32))


# ------------------------------------------- ops.h mirror: errors & dispatch

def _gguf_with(t, k, n, seed=3):
    g = synth.GGUFBuilder()
    w = _weights(t, n, k, seed) if t != 0 else np.zeros(n * k, np.float32)
    g.add_tensor("w", t, (k, n), w)
    from llm_inference_b200.gguf import GGUFFile
    f = GGUFFile(g.build())
    return f, f.tensor("w"), w


def test_ops_h_dispatch_and_error_strings(gpu_ops, port):
    ops = gpu_ops
    rng = np.random.default_rng(8)
    for t, k, n in [(Q4_0, 1152, 64), (Q4_K, 256, 8), (Q6_K, 256, 8), (Q8_0, 64, 8), (Q5_0, 64, 8), (BF16, 64, 8)]:
        f, ti, w = _gguf_with(t, k, n)
        x = rng.standard_normal(k).astype(np.float32)
        o = []
        res = ops.mat_vec_mul(o, ti, f, x)
        assert len(o) == n  # resized like std::vector (ops.cpp:200)
        _check(res, port.mat_vec_mul(t, w, x, n, k))
        name = {Q4_0: "q4_0", Q4_K: "q4_k", Q6_K: "q6_k", Q8_0: "q8_0", Q5_0: "q5_0", BF16: "bf16"}[t]
        with pytest.raises(RuntimeError, match=f"mat_vec_mul_{name}: input vector size mismatch"):
            ops.mat_vec_mul(o, ti, f, x[:-1])
        ops.registry_clear()
    # F16 / F32 tensors are rejected by the dispatcher (ops.cpp:952-955)
    for t in (F16, 0):
        f, ti, _ = _gguf_with(t, 64, 4)
        with pytest.raises(RuntimeError, match=f"mat_vec_mul: unsupported tensor type {t}"):
            ops.mat_vec_mul([], ti, f, np.zeros(64, np.float32))
    # mat_vec_mul_fp16 size errors (ops.cpp:458-463) and the ops_test known answer
    w = np.array([0x3c00, 0x4000, 0x4200, 0x4400, 0x4500, 0x4600, 0x4700, 0x4800], np.uint16)
    o = ops.mat_vec_mul_fp16([], w, np.full(4, 0.5, np.float32), 2, 4)
    assert abs(o[0] - 5.0) < 1e-3 and abs(o[1] - 13.0) < 1e-3
    with pytest.raises(RuntimeError, match="mat_vec_mul_fp16: input vector size mismatch"):
        ops.mat_vec_mul_fp16([], w, np.zeros(3, np.float32), 2, 4)
    with pytest.raises(RuntimeError, match="mat_vec_mul_fp16: weight matrix size mismatch"):
        ops.mat_vec_mul_fp16([], w, np.zeros(4, np.float32), 3, 4)
    ops.registry_clear()


def test_reference_known_answers_on_gpu(gpu_ops):
    ops = gpu_ops
    # ops_test.cpp:138-257 constant-fill single blocks
    b = np.zeros(144, np.uint8); b[0:2] = np.frombuffer(np.float16(1).tobytes(), np.uint8)
    b[4:8] = 1; b[12:16] = 1; b[16:] = 0x22
    cases = [(Q4_K, b, 256, 512.0, 1e-3)]
    b = np.zeros(210, np.uint8); b[0:128] = 0x11; b[128:192] = 0xAA; b[192:208] = 1
    b[208:210] = np.frombuffer(np.float16(1).tobytes(), np.uint8)
    cases.append((Q6_K, b, 256, 256.0, 1e-3))
    b = np.zeros(34, np.uint8); b[0:2] = np.frombuffer(np.float16(1).tobytes(), np.uint8); b[2:] = 2
    cases.append((Q8_0, b, 32, 64.0, 1e-2))
    b = np.zeros(22, np.uint8); b[0:2] = np.frombuffer(np.float16(1).tobytes(), np.uint8); b[2:6] = 0xFF; b[6:] = 0x11
    cases.append((Q5_0, b, 32, 32.0, 1e-3))
    for t, blk, k, expect, tol in cases:
        o, _ = _run(ops, t, blk, np.ones(k, np.float32), 1, k)
        assert abs(o[0] - expect) < tol, (synth.TYPE_NAMES[t], o[0])


def test_batched_launch_equals_separate_launches(gpu_ops, port):
    # q/k/v (and gate/up) of a layer go out as one grid: same bits as three calls
    ops = gpu_ops
    k = 1152
    x = np.random.default_rng(5).standard_normal(k).astype(np.float32)
    dx, act = ops.DeviceVector(k, x), ops.Activation(k)
    for t in (Q4_0, Q8_0, F16):
        shapes = [1024, 256, 264]
        ws = [_weights(t, n, k, seed=n) for n in shapes]
        dws = [ops.DeviceWeight(w, t, k, n) for w, n in zip(ws, shapes)]
        outs_b = [ops.DeviceVector(n) for n in shapes]
        outs_s = [ops.DeviceVector(n) for n in shapes]
        act.prepare(dws[0], dx)
        ops.gemv_batch(dws, act, outs_b)
        for w, o in zip(dws, outs_s):
            ops.gemv(w, act, o)
        ops.device_sync()
        for w, n, ob, os_ in zip(ws, shapes, outs_b, outs_s):
            assert np.array_equal(ob.get().view(np.uint32), os_.get().view(np.uint32))
            _check(ob.get(), port.mat_vec_mul(t, w, x, n, k), "batched")
    w4, w8 = ops.DeviceWeight(_weights(Q4_0, 8, k, 1), Q4_0, k, 8), ops.DeviceWeight(_weights(Q8_0, 8, k, 1), Q8_0, k, 8)
    act.prepare(w4, dx)
    with pytest.raises(RuntimeError, match="share one format"):
        ops.gemv_batch([w4, w8], act, [ops.DeviceVector(8), ops.DeviceVector(8)])


@pytest.mark.parametrize("t", [synth.Q4_0, synth.Q8_0, synth.Q4_K, synth.Q6_K, synth.Q5_0, synth.BF16, synth.F16])
def test_token_batched_matvec_is_bitwise_n_single_calls(gpu_ops, t):
    """llmi_gemm_tokens (prefill) == n_tokens x llmi_mat_vec_mul_dev, bit for bit: ragged N, partial last
    K-chunk, token counts on both sides of the kernel switches (token loop < 16 <= token-per-lane dp4a < 128 <=
    tcgen05 int8 for Q4_0 / Q8_0) and of the tile / lane-group sizes, and a row-shard handle."""
    ops = gpu_ops
    kq = t in (synth.Q4_K, synth.Q6_K)
    for k, n in ((512, 40), (1280 if kq else 1184, 203), (2560 if kq else 2592, 77)):
        w_host = synth.random_blocks(t, n, k, seed=k + n)
        w = ops.DeviceWeight(w_host, t, k, n)
        from llm_inference_b200.shard import shard_blocks
        shard = ops.DeviceWeight(shard_blocks(w_host, synth.row_bytes(t, k), (8, 32)), t, k, n, 8, 32,
                                 blocks_are_shard=True)
        act = ops.Activation(k)
        for m in (1, 5, 16, 37, 70) + ((131,) if t in (synth.Q4_0, synth.Q8_0) else ()):
            x = np.random.default_rng(m).standard_normal((m, k)).astype(np.float32)
            x[0, :32] = 0.0  # an all-zero block
            xs, out = ops.DeviceVector(m * k, x), ops.DeviceVector(m * n, np.full(m * n, np.nan, np.float32))
            ops.gemm_tokens(w, xs, m, out)
            got = out.get().reshape(m, n)
            one, o1 = ops.DeviceVector(k), ops.DeviceVector(n)
            for i in range(m):
                one.set(x[i])
                ops.mat_vec_mul_dev(w, one, act, o1)
                assert np.array_equal(got[i].view(np.uint32), o1.get().view(np.uint32)), (t, k, n, m, i)
            out.set(np.full(m * n, np.nan, np.float32))
            ops.gemm_tokens(shard, xs, m, out)
            part = out.get().reshape(m, n)
            assert np.array_equal(part[:, 8:32].view(np.uint32), got[:, 8:32].view(np.uint32))
            assert np.isnan(part[:, :8]).all() and np.isnan(part[:, 32:]).all()
            for v in (xs, out, one, o1):
                v.close()
        act.close()
        w.close()
        shard.close()


@pytest.mark.parametrize("t", [Q4_0, Q8_0, Q4_K, Q6_K, Q5_0, F16, BF16])
def test_every_kernel_is_bitwise_the_canonical_order_oracle(gpu_ops, port, t):
    """The CPU restatement of the device's summation order (oracle/qgemv_oracle.c, orc_gemv_*_canonical: the
    reference's per-block terms; K-chunks, four sub-lane chains per chunk, (s0+s1)+(s2+s3), chunks left to right)
    against every GPU kernel of the path — one-token, token loop (5 tokens), token-per-lane dp4a (37), tcgen05 int8
    (131; Q4_0 / Q8_0) — BIT FOR BIT, on ragged shapes with a partial last chunk."""
    ops = gpu_ops
    kq = t in (Q4_K, Q6_K)
    for k, n in ((1280 if kq else 1184, 203), (2560 if kq else 2592, 77), (512, 40)):
        w_host = _weights(t, n, k, seed=3 * k + n)
        w = ops.DeviceWeight(w_host, t, k, n)
        act = ops.Activation(k)
        for m in (1, 5, 37) + ((131,) if t in (Q4_0, Q8_0) else ()):
            x = np.random.default_rng(m + k).standard_normal((m, k)).astype(np.float32)
            want = np.stack([port.mat_vec_mul_canonical(t, w_host, x[i], n, k) for i in range(m)])
            if m == 1:
                dx, do = ops.DeviceVector(k, x[0]), ops.DeviceVector(n)
                ops.mat_vec_mul_dev(w, dx, act, do)
                got = do.get().reshape(1, n)
            else:
                xs, out = ops.DeviceVector(m * k, x), ops.DeviceVector(m * n)
                ops.gemm_tokens(w, xs, m, out)
                got = out.get().reshape(m, n)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (t, k, n, m)


@pytest.mark.parametrize("t", [Q4_0, Q8_0])
def test_tensor_core_prefill_direct_store_is_bitwise_single_calls(gpu_ops, t):
    """A token batch large enough that every CTA of the tcgen05 int8 kernel walks ALL K-chunks of its tile
    (grid.z == 1: rows/128 x tokens/32 >= 2 CTAs per SM): the chunk partials are then added in registers and each
    (token, row) is stored once — no partials buffer, no reduce launch (umma_prefill.cuh, DIRECT).  Sampled tokens,
    bit for bit against the one-token mat-vec; ragged N, ragged token tile and a K that ends inside a chunk."""
    ops = gpu_ops
    k, n, m = 1184 if t == Q4_0 else 1056, 4099, 331
    w_host = synth.random_blocks(t, n, k, seed=91)
    w = ops.DeviceWeight(w_host, t, k, n)
    act = ops.Activation(k)
    x = np.random.default_rng(5).standard_normal((m, k)).astype(np.float32)
    xs, out = ops.DeviceVector(m * k, x), ops.DeviceVector(m * n, np.full(m * n, np.nan, np.float32))
    ops.gemm_tokens(w, xs, m, out)
    got = out.get().reshape(m, n)
    assert not np.isnan(got).any()
    one, o1 = ops.DeviceVector(k), ops.DeviceVector(n)
    for i in (0, 31, 32, 170, 319, 320, 330):
        one.set(x[i])
        ops.mat_vec_mul_dev(w, one, act, o1)
        assert np.array_equal(got[i].view(np.uint32), o1.get().view(np.uint32)), (t, i)
    for v in (xs, out, one, o1, act, w):
        v.close()


@pytest.mark.parametrize("t", [Q4_0, Q8_0, Q4_K, Q6_K, Q5_0, F16, BF16])
def test_fast_prefill_gemm_is_within_its_stated_tolerance(gpu_ops, t):
    """llmi_set_prefill_mode(1): the dequantize-to-bf16 tcgen05 GEMM (gemm_bf16.cuh).  Not the parity path — the bar it
    states is |o - o_exact| <= 2e-2 * max|o_exact| per call against the exact token-batched mat-vec (bf16 rounds both
    operands to 8 mantissa bits; measured ~2e-3); ragged row and token tiles, one and several token tiles."""
    ops = gpu_ops
    kq = t in (Q4_K, Q6_K)
    try:
        for k, n, m in ((1280 if kq else 1152, 203, 131), (2560, 520, 300), (512, 136, 64)):
            w = ops.DeviceWeight(_weights(t, n, k, seed=5 * k + n), t, k, n)
            x = np.random.default_rng(m + k).standard_normal((m, k)).astype(np.float32)
            xs, out = ops.DeviceVector(m * k, x), ops.DeviceVector(m * n, np.full(m * n, np.nan, np.float32))
            ops.set_prefill_mode(False)
            ops.gemm_tokens(w, xs, m, out)
            ops.device_sync()
            exact = out.get().reshape(m, n).copy()
            out2 = ops.DeviceVector(m * n, np.full(m * n, np.nan, np.float32))
            ops.set_prefill_mode(True)
            ops.gemm_tokens(w, xs, m, out2)
            ops.device_sync()
            fast = out2.get().reshape(m, n)
            assert not np.isnan(fast).any()
            scale = float(np.abs(exact).max())
            err = float(np.abs(fast - exact).max())
            assert err <= 2e-2 * scale, (synth.TYPE_NAMES[t], k, n, m, err / scale)
            assert err > 0 or t in (F16, BF16)  # it IS another arithmetic: a bit-equal result would mean the exact path ran
            for h in (w, xs, out, out2):
                h.close()
    finally:
        ops.set_prefill_mode(False)
