"""TEST INFRASTRUCTURE — ctypes bindings for the parity oracles.

Two checkers live here, neither is ever on the product path:

* ``Port``  — ``oracle/liboracle.so``, our C restatement (``qgemv_oracle.c``).
* ``Ref``   — ``oracle/_ref/libref.so``, the reference's own ops/gguf/model
  sources compiled in place (``oracle/Makefile``); present whenever it was
  built in the dev container (it travels to the GPU box with the snapshot).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("LLMI_REFERENCE", "/root/reference"))

Q8_0_BYTES = 34
Q8_K_BYTES = 292

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u16p = C.POINTER(C.c_uint16)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def build(target: str = "oracle", quiet: bool = True) -> None:
    """Run ``make <target>`` in oracle/ (gcc/g++ only)."""
    subprocess.run(
        ["make", "-C", str(HERE), target],
        check=True,
        stdout=subprocess.DEVNULL if quiet else None,
        stderr=subprocess.STDOUT if quiet else None,
    )


def row_bytes(ggml_type: int, k: int) -> int:
    per = {2: (32, 18), 8: (32, 34), 6: (32, 22), 12: (256, 144), 14: (256, 210),
           1: (1, 2), 30: (1, 2), 0: (1, 4)}[ggml_type]
    assert k % per[0] == 0, f"K={k} not a multiple of {per[0]}"
    return k // per[0] * per[1]


class Port:
    """The C restatement (kind == "port")."""

    def __init__(self) -> None:
        so = HERE / "liboracle.so"
        src = HERE / "qgemv_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            build("oracle")
        L = C.CDLL(str(so))
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [C.c_float]
        L.orc_bf16_to_f32.restype = C.c_float
        L.orc_bf16_to_f32.argtypes = [C.c_uint16]
        L.orc_nearest_int.restype = C.c_int
        L.orc_nearest_int.argtypes = [C.c_float]
        L.orc_quantize_row_q8_0.argtypes = [_f32p, _u8p, C.c_size_t]
        L.orc_quantize_row_q8_k.argtypes = [_f32p, _u8p, C.c_size_t]
        for n in ("q4_0", "q8_0", "q4_k", "q6_k"):
            getattr(L, f"orc_gemv_{n}").argtypes = [_f32p, _u8p, _f32p, C.c_size_t, C.c_size_t, _i32p]
        L.orc_gemv_q5_0.argtypes = [_f32p, _u8p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_gemv_q4_0_canonical.argtypes = [_f32p, _u8p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_gemv_q8_0_canonical.argtypes = [_f32p, _u8p, _f32p, C.c_size_t, C.c_size_t]
        for n in ("q4_k", "q6_k", "q5_0"):
            getattr(L, f"orc_gemv_{n}_canonical").argtypes = [_f32p, _u8p, _f32p, C.c_size_t, C.c_size_t]
        for n in ("f16", "bf16"):
            getattr(L, f"orc_gemv_{n}_canonical").argtypes = [_f32p, _u16p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_gemv_bf16.argtypes = [_f32p, _u16p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_gemv_f16.argtypes = [_f32p, _u16p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_mat_vec_mul.restype = C.c_int
        L.orc_mat_vec_mul.argtypes = [C.c_uint32, _f32p, _u8p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_dequantize_row.restype = C.c_int
        L.orc_dequantize_row.argtypes = [C.c_uint32, _u8p, C.c_size_t, _f32p]
        L.orc_row_exact.restype = C.c_double
        L.orc_row_exact.argtypes = [C.c_uint32, _u8p, _f32p, C.c_size_t, C.POINTER(C.c_double)]
        self.L = L

    kind = "port"

    def f16_to_f32(self, h: np.ndarray) -> np.ndarray:
        h = np.ascontiguousarray(h, np.uint16)
        return np.array([self.L.orc_f16_to_f32(int(v)) for v in h.ravel()], np.float32).reshape(h.shape)

    def f32_to_f16(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, np.float32)
        return np.array([self.L.orc_f32_to_f16(float(v)) for v in f.ravel()], np.uint16).reshape(f.shape)

    def quantize_row_q8_0(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros(x.size // 32 * Q8_0_BYTES, np.uint8)
        self.L.orc_quantize_row_q8_0(_p(x, _f32p), _p(y, _u8p), x.size)
        return y

    def quantize_row_q8_k(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros(x.size // 256 * Q8_K_BYTES, np.uint8)
        self.L.orc_quantize_row_q8_k(_p(x, _f32p), _p(y, _u8p), x.size)
        return y

    def mat_vec_mul(self, ggml_type: int, w: np.ndarray, x: np.ndarray, n_rows: int, n_cols: int,
                    want_dots: bool = False):
        """Returns o (and the per-block integer dots when asked)."""
        w = np.ascontiguousarray(w).view(np.uint8).ravel()
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros(n_rows, np.float32)
        names = {2: "q4_0", 8: "q8_0", 12: "q4_k", 14: "q6_k"}
        if ggml_type in names:
            per = {2: n_cols // 32, 8: n_cols // 32, 12: n_cols // 32, 14: n_cols // 128}[ggml_type]
            dots = np.zeros(n_rows * per, np.int32) if want_dots else None
            getattr(self.L, f"orc_gemv_{names[ggml_type]}")(
                _p(o, _f32p), _p(w, _u8p), _p(x, _f32p), n_rows, n_cols,
                _p(dots, _i32p) if want_dots else None)
            return (o, dots) if want_dots else o
        assert not want_dots
        if ggml_type == 6:
            self.L.orc_gemv_q5_0(_p(o, _f32p), _p(w, _u8p), _p(x, _f32p), n_rows, n_cols)
        elif ggml_type == 30:
            self.L.orc_gemv_bf16(_p(o, _f32p), _p(w.view(np.uint16), _u16p), _p(x, _f32p), n_rows, n_cols)
        elif ggml_type == 1:
            self.L.orc_gemv_f16(_p(o, _f32p), _p(w.view(np.uint16), _u16p), _p(x, _f32p), n_rows, n_cols)
        else:
            raise RuntimeError(f"mat_vec_mul: unsupported tensor type {ggml_type}")
        return o

    def mat_vec_mul_canonical(self, ggml_type: int, w: np.ndarray, x: np.ndarray, n_rows: int, n_cols: int) -> np.ndarray:
        """Mat-vec with the reference's per-block terms summed in the DEVICE's canonical order: what every GPU
        kernel of the path must reproduce bit for bit (all seven formats)."""
        w = np.ascontiguousarray(w).view(np.uint8).ravel()
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros(n_rows, np.float32)
        name = {2: "q4_0", 8: "q8_0", 12: "q4_k", 14: "q6_k", 6: "q5_0", 1: "f16", 30: "bf16"}[ggml_type]
        fn = getattr(self.L, f"orc_gemv_{name}_canonical")
        if ggml_type in (1, 30):
            fn(_p(o, _f32p), _p(w.view(np.uint16), _u16p), _p(x, _f32p), n_rows, n_cols)
        else:
            fn(_p(o, _f32p), _p(w, _u8p), _p(x, _f32p), n_rows, n_cols)
        return o

    def dequantize_row(self, ggml_type: int, row: np.ndarray, n_cols: int) -> np.ndarray:
        row = np.ascontiguousarray(row).view(np.uint8).ravel()
        out = np.zeros(n_cols, np.float32)
        rc = self.L.orc_dequantize_row(ggml_type, _p(row, _u8p), n_cols, _p(out, _f32p))
        if rc:
            raise RuntimeError(f"dequantize_row: unsupported type {ggml_type}")
        return out

    def row_exact(self, ggml_type: int, w_row: np.ndarray, x: np.ndarray, n_cols: int):
        w_row = np.ascontiguousarray(w_row).view(np.uint8).ravel()
        x = np.ascontiguousarray(x, np.float32)
        sa = C.c_double(0)
        v = self.L.orc_row_exact(ggml_type, _p(w_row, _u8p), _p(x, _f32p), n_cols, C.byref(sa))
        return v, sa.value


def ref_available() -> bool:
    return (HERE / "_ref" / "libref.so").exists()


def ensure_ref() -> bool:
    """Build oracle/_ref when the reference sources are present (dev container)."""
    if REF_SRC.joinpath("ops.cpp").exists():
        build("ref")
    return ref_available()


class Ref:
    """The reference's own compiled CPU code (kind == "reference")."""

    kind = "reference"

    def __init__(self, lib: str = "libref.so", n_threads: int = 1) -> None:
        so = HERE / "_ref" / lib
        if not so.exists():
            raise FileNotFoundError(f"{so} not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(str(so))
        L.ref_last_error.restype = C.c_char_p
        L.ref_init_ops.argtypes = [C.c_int]
        L.ref_f16_to_f32_n.argtypes = [_u16p, _f32p, C.c_uint64]
        L.ref_f32_to_f16_n.argtypes = [_f32p, _u16p, C.c_uint64]
        L.ref_bf16_to_f32.restype = C.c_float
        L.ref_bf16_to_f32.argtypes = [C.c_uint16]
        L.ref_quantize_row_q8_0.argtypes = [_f32p, C.c_uint64, _u8p]
        L.ref_quantize_row_q8_k.argtypes = [_f32p, C.c_uint64, _u8p]
        L.ref_tensor_create.restype = C.c_void_p
        L.ref_tensor_create.argtypes = [C.c_uint32, _u8p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.ref_tensor_free.argtypes = [C.c_void_p]
        L.ref_tensor_mat_vec_mul.argtypes = [C.c_void_p, _f32p, C.c_uint64, _f32p, C.c_uint64]
        L.ref_tensor_mat_vec_mul_loop.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_int]
        L.ref_mat_vec_mul_fp16.argtypes = [_u16p, C.c_uint64, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, _f32p]
        L.ref_f16_create.restype = C.c_void_p
        L.ref_f16_create.argtypes = [_u16p, C.c_uint64, C.c_uint64]
        L.ref_f16_free.argtypes = [C.c_void_p]
        L.ref_f16_mat_vec_mul_loop.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
        L.ref_dequantize_row.argtypes = [C.c_uint32, _u8p, C.c_uint64, _f32p]
        L.ref_rms_norm.argtypes = [_f32p, C.c_uint64, C.c_double, _f32p]
        L.ref_softmax.argtypes = [_f32p, C.c_uint64]
        L.ref_rope.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int]
        L.ref_model_create.restype = C.c_void_p
        L.ref_model_create.argtypes = [_u8p, C.c_uint64, C.c_int]
        L.ref_model_free.argtypes = [C.c_void_p]
        L.ref_model_forward.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _f32p, C.c_uint64]
        L.ref_model_vocab.restype = C.c_uint64
        L.ref_model_vocab.argtypes = [C.c_void_p]
        self.L = L
        self.n_threads = n_threads
        L.ref_init_ops(n_threads)

    def _err(self) -> str:
        return self.L.ref_last_error().decode()

    def f16_to_f32(self, h: np.ndarray) -> np.ndarray:
        h = np.ascontiguousarray(h, np.uint16)
        f = np.zeros(h.shape, np.float32)
        self.L.ref_f16_to_f32_n(_p(h, _u16p), _p(f, _f32p), h.size)
        return f

    def f32_to_f16(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, np.float32)
        h = np.zeros(f.shape, np.uint16)
        self.L.ref_f32_to_f16_n(_p(f, _f32p), _p(h, _u16p), f.size)
        return h

    def quantize_row_q8_0(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros(x.size // 32 * Q8_0_BYTES, np.uint8)
        if self.L.ref_quantize_row_q8_0(_p(x, _f32p), x.size, _p(y, _u8p)):
            raise RuntimeError(self._err())
        return y

    def quantize_row_q8_k(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros(x.size // 256 * Q8_K_BYTES, np.uint8)
        if self.L.ref_quantize_row_q8_k(_p(x, _f32p), x.size, _p(y, _u8p)):
            raise RuntimeError(self._err())
        return y

    def tensor(self, ggml_type: int, w: np.ndarray, n_rows: int, n_cols: int) -> "RefTensor":
        return RefTensor(self, ggml_type, w, n_rows, n_cols)

    def mat_vec_mul(self, ggml_type: int, w: np.ndarray, x: np.ndarray, n_rows: int, n_cols: int) -> np.ndarray:
        if ggml_type == 1:
            return self.mat_vec_mul_fp16(w, x, n_rows, n_cols)
        t = self.tensor(ggml_type, w, n_rows, n_cols)
        try:
            return t.mat_vec_mul(x)
        finally:
            t.close()

    def mat_vec_mul_fp16(self, w: np.ndarray, x: np.ndarray, n_rows: int, n_cols: int) -> np.ndarray:
        w = np.ascontiguousarray(w).view(np.uint16).ravel()
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros(n_rows, np.float32)
        if self.L.ref_mat_vec_mul_fp16(_p(w, _u16p), w.size, _p(x, _f32p), x.size, n_rows, n_cols, _p(o, _f32p)):
            raise RuntimeError(self._err())
        return o

    def dequantize_row(self, ggml_type: int, row: np.ndarray, n_cols: int) -> np.ndarray:
        row = np.ascontiguousarray(row).view(np.uint8).ravel()
        out = np.zeros(n_cols, np.float32)
        if self.L.ref_dequantize_row(ggml_type, _p(row, _u8p), n_cols, _p(out, _f32p)):
            raise RuntimeError(self._err())
        return out

    def rms_norm(self, x: np.ndarray, eps: float) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros_like(x)
        self.L.ref_rms_norm(_p(x, _f32p), x.size, eps, _p(o, _f32p))
        return o

    def rope(self, t: np.ndarray, n_rot: int, base: float, scale: float, pos: int) -> np.ndarray:
        t = np.ascontiguousarray(t, np.float32).copy()
        nt, nh, hd = t.shape
        self.L.ref_rope(_p(t, _f32p), nt, nh, hd, n_rot, base, scale, pos)
        return t

    def model(self, gguf: np.ndarray) -> "RefModel":
        return RefModel(self, gguf)


class RefTensor:
    def __init__(self, ref: Ref, ggml_type: int, w: np.ndarray, n_rows: int, n_cols: int) -> None:
        self.ref = ref
        w = np.ascontiguousarray(w).view(np.uint8).ravel()
        self.n_rows, self.n_cols = n_rows, n_cols
        self.h = ref.L.ref_tensor_create(ggml_type, _p(w, _u8p), w.size, n_cols, n_rows)
        if not self.h:
            raise RuntimeError(ref._err())

    def mat_vec_mul(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros(self.n_rows, np.float32)
        if self.ref.L.ref_tensor_mat_vec_mul(self.h, _p(x, _f32p), x.size, _p(o, _f32p), o.size):
            raise RuntimeError(self.ref._err())
        return o

    def loop(self, x: np.ndarray, iters: int) -> None:
        x = np.ascontiguousarray(x, np.float32)
        if self.ref.L.ref_tensor_mat_vec_mul_loop(self.h, _p(x, _f32p), x.size, iters):
            raise RuntimeError(self.ref._err())

    def close(self) -> None:
        if self.h:
            self.ref.L.ref_tensor_free(self.h)
            self.h = None


class RefModel:
    """Model(GGUFFile&) + forward(tokens, pos) of the reference (model.h:72-91)."""

    def __init__(self, ref: Ref, gguf: np.ndarray) -> None:
        self.ref = ref
        self._img = np.ascontiguousarray(gguf, np.uint8)  # borrowed by GGUFFile: keep alive
        self.h = ref.L.ref_model_create(_p(self._img, _u8p), self._img.size, 0)
        if not self.h:
            raise RuntimeError(ref._err())
        self.vocab = int(ref.L.ref_model_vocab(self.h))

    def forward(self, tokens, pos: int) -> np.ndarray:
        tk = np.ascontiguousarray(tokens, np.int32)
        logits = np.zeros(self.vocab, np.float32)
        if self.ref.L.ref_model_forward(self.h, _p(tk, _i32p), tk.size, pos, _p(logits, _f32p), logits.size):
            raise RuntimeError(self.ref._err())
        return logits

    def close(self) -> None:
        if self.h:
            self.ref.L.ref_model_free(self.h)
            self.h = None
