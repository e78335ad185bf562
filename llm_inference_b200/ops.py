"""Python mirror of the reference's operator interface for the quantized mat-vec
path (ops.h:38-70, 88-105), on top of the C ABI (include/llmi_cuda.h).

Same names, argument meaning and error behaviour as the reference:

* ``init_ops(n_threads)``                                     ops.h:38
* ``mat_vec_mul(o, w_tensor, gguf_file, x)`` + typed variants  ops.h:53-70
* ``mat_vec_mul_fp16(o, w, x, n_rows, n_cols)``                ops.h:41
* ``quantize_row_q8_0(x, y, size)`` / ``quantize_row_q8_k``    ops.h:94,104

``o``/``y`` are Python lists or numpy arrays that are resized/overwritten like
the reference's ``std::vector&`` out-parameters; every function also returns
the result as a numpy array.  Errors are ``RuntimeError`` with the reference's
message strings (ops.cpp:197,459,462,620,714,793,846,901,953).

This is the *host-vector tier*: each call copies x to the device and o back and
synchronises, exactly like one reference call.  The timed path is the device
tier (``DeviceWeight`` / ``DeviceVector`` / ``Activation`` below and
``llm_inference_b200.model``).
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib
from .gguf import GGUFFile, TensorInfo
from .synth import BF16, F16, Q4_0, Q4_K, Q5_0, Q6_K, Q8_0, row_bytes

_NAMES = {Q4_0: "q4_0", Q4_K: "q4_k", Q6_K: "q6_k", Q8_0: "q8_0", Q5_0: "q5_0", BF16: "bf16"}
_initialised = False


def init_ops(n_threads: int = 1, device: int = 0) -> None:
    """ops.h:38.  The thread count is meaningless on the GPU (row partitioning
    became grid partitioning); kept for signature compatibility."""
    global _initialised
    del n_threads
    _lib.check(_lib.load().llmi_init(device))
    _initialised = True


def _need_init() -> C.CDLL:
    if not _initialised:
        # the reference dereferences a null pool here (ops.cpp:18,447); we raise
        raise RuntimeError("init_ops() must be called before any mat_vec_mul")
    return _lib.load()


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _assign(out, values: np.ndarray):
    if out is None:
        return
    if isinstance(out, list):
        out[:] = values.tolist()
    elif isinstance(out, np.ndarray):
        out.resize(values.shape, refcheck=False)
        out[...] = values


# ------------------------------------------------------------------ device tier

class DeviceWeight:
    """One uploaded + repacked matrix (llmi_weight_upload)."""

    def __init__(self, blocks: np.ndarray, ggml_type: int, n_cols: int, n_rows: int,
                 row_begin: int = 0, row_end: int | None = None, handle=None, blocks_are_shard: bool = False) -> None:
        L = _need_init()
        self.type, self.n_cols, self.n_rows = ggml_type, n_cols, n_rows
        self.row_begin = row_begin
        self.row_end = n_rows if row_end is None else row_end
        self.owned = handle is None
        if handle is None:
            raw = np.ascontiguousarray(blocks).view(np.uint8).ravel()
            base = _ptr(raw)
            if blocks_are_shard:  # `blocks` holds only rows [row_begin,row_end): address of the virtual row 0
                base -= self.row_begin * row_bytes(ggml_type, n_cols)
            h = C.c_void_p()
            _lib.check(L.llmi_weight_upload(base, ggml_type, n_cols, n_rows, self.row_begin, self.row_end,
                                            C.byref(h)))
            handle = h
        self.h = handle

    def device_bytes(self) -> int:
        return int(_lib.load().llmi_weight_device_bytes(self.h))

    def close(self) -> None:
        if self.h and self.owned:
            _lib.load().llmi_weight_free(self.h)
        self.h = None


class DeviceVector:
    """fp32 vector in device memory."""

    def __init__(self, n: int, host: np.ndarray | None = None) -> None:
        L = _need_init()
        self.n = n
        p = C.c_void_p()
        _lib.check(L.llmi_dev_alloc(4 * n, C.byref(p)))
        self.p = p
        if host is not None:
            self.set(host)

    def set(self, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host, np.float32)
        assert host.size == self.n
        _lib.check(_lib.load().llmi_h2d(self.p, _ptr(host), 4 * self.n))

    def get(self) -> np.ndarray:
        out = np.empty(self.n, np.float32)
        _lib.check(_lib.load().llmi_d2h(_ptr(out), self.p, 4 * self.n))
        return out

    def close(self) -> None:
        if self.p:
            _lib.load().llmi_dev_free(self.p)
            self.p = None


class TorchVector:
    """fp32 torch CUDA tensor viewed as a device vector (for torch.distributed collectives)."""

    def __init__(self, tensor) -> None:
        assert tensor.is_cuda and tensor.dtype.is_floating_point and tensor.element_size() == 4
        self.t = tensor
        self.n = tensor.numel()
        self.p = C.c_void_p(tensor.data_ptr())

    def close(self) -> None:
        self.t = None


class Activation:
    """Quantized / staged activation vector (llmi_act_*)."""

    def __init__(self, max_cols: int) -> None:
        L = _need_init()
        h = C.c_void_p()
        _lib.check(L.llmi_act_create(max_cols, C.byref(h)))
        self.h = h

    def prepare(self, w: DeviceWeight, x: DeviceVector, stream=None) -> None:
        _lib.check(_lib.load().llmi_act_prepare(w.h, x.p, self.h, stream))

    def quantize_q8_0(self, x: DeviceVector, stream=None) -> None:
        _lib.check(_lib.load().llmi_quantize_q8_0(x.p, x.n, self.h, stream))

    def quantize_q8_k(self, x: DeviceVector, stream=None) -> None:
        _lib.check(_lib.load().llmi_quantize_q8_k(x.p, x.n, self.h, stream))

    def export_q8_0(self, n: int) -> np.ndarray:
        out = np.zeros(n // 32 * 34, np.uint8)
        _lib.check(_lib.load().llmi_act_export_q8_0(self.h, _ptr(out)))
        return out

    def export_q8_k(self, n: int) -> np.ndarray:
        out = np.zeros(n // 256 * 292, np.uint8)
        _lib.check(_lib.load().llmi_act_export_q8_k(self.h, _ptr(out)))
        return out

    def close(self) -> None:
        if self.h:
            _lib.load().llmi_act_free(self.h)
            self.h = None


def gemv(w: DeviceWeight, act: Activation, out: DeviceVector, stream=None) -> None:
    _lib.check(_lib.load().llmi_gemv(w.h, act.h, out.p, stream))


def gemv_batch(ws, act: Activation, outs, stream=None) -> None:
    """Up to 3 same-format matrices consuming one activation, as one grid (llmi_gemv_batch)."""
    n = len(ws)
    wa = (C.c_void_p * n)(*[w.h for w in ws])
    oa = (C.c_void_p * n)(*[o.p for o in outs])
    _lib.check(_lib.load().llmi_gemv_batch(wa, oa, n, act.h, stream))


def mat_vec_mul_dev(w: DeviceWeight, x: DeviceVector, act: Activation, out: DeviceVector, stream=None) -> None:
    _lib.check(_lib.load().llmi_mat_vec_mul_dev(w.h, x.p, act.h, out.p, stream))


def gemm_tokens(w: DeviceWeight, xs: DeviceVector, n_tokens: int, out: DeviceVector, stream=None) -> None:
    """Token-batched mat-vec (prefill): xs = [n_tokens][n_cols], out = [n_tokens][n_rows] (llmi_gemm_tokens)."""
    _lib.check(_lib.load().llmi_gemm_tokens(w.h, xs.p, n_tokens, out.p, stream))


def block_dots(w: DeviceWeight, act: Activation) -> np.ndarray:
    per = {Q4_0: w.n_cols // 32, Q8_0: w.n_cols // 32, Q4_K: w.n_cols // 32, Q6_K: w.n_cols // 128}[w.type]
    out = np.zeros((w.row_end - w.row_begin) * per, np.int32)
    _lib.check(_lib.load().llmi_debug_block_dots(w.h, act.h, _ptr(out)))
    return out


def set_gemv_shape(warps: int = 0, slabs_per_cta: int = 0) -> None:
    """Pin the mat-vec CTA shape (benches); 0 = heuristic.  Results never change."""
    _lib.check(_lib.load().llmi_set_gemv_shape(warps, slabs_per_cta))


def set_gemv_ring(mode: int = 0, ctas_per_sm: int = 0, depth: int = 0, warps: int = 0) -> None:
    """Select the persistent ring mat-vec kernel: 0 heuristic, 1 never, 2 wherever it fits.  Same bits."""
    _lib.check(_lib.load().llmi_set_gemv_ring(mode, ctas_per_sm, depth, warps))


def set_prefill_mode(fast: bool) -> None:
    """False (default): exact token-batched mat-vec.  True: dequantize-to-bf16 tcgen05 GEMM (tolerance, not bit-exact)."""
    _lib.check(_lib.load().llmi_set_prefill_mode(1 if fast else 0))


def device_sync() -> None:
    _lib.check(_lib.load().llmi_device_sync())


# ------------------------------------------------- ops.h mirror (host vectors)

def _registry_weight(w_tensor: TensorInfo, gguf_file: GGUFFile) -> DeviceWeight:
    L = _need_init()
    n_cols, n_rows = int(w_tensor.shape[0]), int(w_tensor.shape[1])
    data = gguf_file.get_tensor_data(w_tensor)  # borrowed from the GGUF image, like the mmap pointer
    h = C.c_void_p()
    _lib.check(L.llmi_registry_get(_ptr(data), w_tensor.tensor_type, n_cols, n_rows, C.byref(h)))
    return DeviceWeight(None, w_tensor.tensor_type, n_cols, n_rows, handle=h)


def _typed(name: str, ggml_type: int):
    def fn(o, w_tensor: TensorInfo, gguf_file: GGUFFile, x) -> np.ndarray:
        n_cols, n_rows = int(w_tensor.shape[0]), int(w_tensor.shape[1])
        xv = np.ascontiguousarray(x, np.float32).ravel()
        if xv.size != n_cols:
            raise RuntimeError(f"mat_vec_mul_{name}: input vector size mismatch")
        if w_tensor.tensor_type != ggml_type:
            raise RuntimeError(f"mat_vec_mul_{name}: tensor is not {name}")
        w = _registry_weight(w_tensor, gguf_file)
        out = np.empty(n_rows, np.float32)
        _lib.check(_lib.load().llmi_host_mat_vec_mul(w.h, _ptr(xv), xv.size, _ptr(out), out.size))
        _assign(o, out)
        return out

    fn.__name__ = f"mat_vec_mul_{name}"
    fn.__doc__ = f"ops.h mat_vec_mul_{name}(o, w_tensor, gguf_file, x) on the B200."
    return fn


mat_vec_mul_q4_0 = _typed("q4_0", Q4_0)   # ops.h:56, ops.cpp:188-451
mat_vec_mul_q4_k = _typed("q4_k", Q4_K)   # ops.h:60, ops.cpp:614-706
mat_vec_mul_q6_k = _typed("q6_k", Q6_K)   # ops.h:63, ops.cpp:708-785
mat_vec_mul_q8_0 = _typed("q8_0", Q8_0)   # ops.h:65, ops.cpp:787-838
mat_vec_mul_q5_0 = _typed("q5_0", Q5_0)   # ops.h:67, ops.cpp:840-893
mat_vec_mul_bf16 = _typed("bf16", BF16)   # ops.h:69, ops.cpp:895-931


def mat_vec_mul(o, w_tensor: TensorInfo, gguf_file: GGUFFile, x) -> np.ndarray:
    """ops.h:53 dispatcher (ops.cpp:933-956): a K mismatch is only logged here
    (the typed function then throws); F16/F32 tensors are not accepted."""
    xv = np.asarray(x)
    if xv.size != w_tensor.shape[0]:
        print(f"mat_vec_mul size mismatch: tensor: {w_tensor.name} w_tensor.shape[0]={w_tensor.shape[0]} "
              f"x.size()={xv.size}", file=sys.stderr)
    fn = {Q4_0: mat_vec_mul_q4_0, Q4_K: mat_vec_mul_q4_k, Q6_K: mat_vec_mul_q6_k, Q8_0: mat_vec_mul_q8_0,
          Q5_0: mat_vec_mul_q5_0, BF16: mat_vec_mul_bf16}.get(w_tensor.tensor_type)
    if fn is None:
        raise RuntimeError(f"mat_vec_mul: unsupported tensor type {w_tensor.tensor_type}")
    return fn(o, w_tensor, gguf_file, x)


def mat_vec_mul_fp16(o, w, x, n_rows: int, n_cols: int) -> np.ndarray:
    """ops.h:41 (ops.cpp:455-612): w is the row-major f16 bit pattern matrix."""
    L = _need_init()
    xv = np.ascontiguousarray(x, np.float32).ravel()
    wv = w if isinstance(w, np.ndarray) and w.dtype == np.uint16 and w.flags.c_contiguous else \
        np.ascontiguousarray(w, np.uint16)
    if xv.size != n_cols:
        raise RuntimeError("mat_vec_mul_fp16: input vector size mismatch")
    if wv.size != n_rows * n_cols:
        raise RuntimeError("mat_vec_mul_fp16: weight matrix size mismatch")
    h = C.c_void_p()
    _lib.check(L.llmi_registry_get(_ptr(wv), F16, n_cols, n_rows, C.byref(h)))
    out = np.empty(n_rows, np.float32)
    _lib.check(L.llmi_host_mat_vec_mul(h, _ptr(xv), xv.size, _ptr(out), out.size))
    _assign(o, out)
    return out


def quantize_row_q8_0(x, y=None, size: int | None = None) -> np.ndarray:
    """ops.h:94.  Returns size/32 BlockQ8_0 records (34 B each) as uint8."""
    L = _need_init()
    xv = np.ascontiguousarray(x, np.float32).ravel()
    n = xv.size if size is None else size
    out = np.zeros(n // 32 * 34, np.uint8)
    _lib.check(L.llmi_host_quantize_row_q8_0(_ptr(xv), n, _ptr(out)))
    _assign(y, out)
    return out


def quantize_row_q8_k(x, y=None, size: int | None = None) -> np.ndarray:
    """ops.h:104.  Returns size/256 block_q8_K records (292 B each) as uint8."""
    L = _need_init()
    xv = np.ascontiguousarray(x, np.float32).ravel()
    n = xv.size if size is None else size
    out = np.zeros(n // 256 * 292, np.uint8)
    _lib.check(L.llmi_host_quantize_row_q8_k(_ptr(xv), n, _ptr(out)))
    _assign(y, out)
    return out


def registry_clear() -> None:
    _lib.check(_lib.load().llmi_registry_clear())


__all__ = [
    "init_ops", "mat_vec_mul", "mat_vec_mul_q4_0", "mat_vec_mul_q4_k", "mat_vec_mul_q6_k", "mat_vec_mul_q8_0",
    "mat_vec_mul_q5_0", "mat_vec_mul_bf16", "mat_vec_mul_fp16", "quantize_row_q8_0", "quantize_row_q8_k",
    "DeviceWeight", "DeviceVector", "TorchVector", "Activation", "gemv", "gemv_batch", "mat_vec_mul_dev", "block_dots", "set_gemv_shape", "set_gemv_ring", "set_prefill_mode",
    "device_sync", "registry_clear", "row_bytes",
]
