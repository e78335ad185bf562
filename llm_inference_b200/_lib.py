"""ctypes loader for libllmi_cuda.so (C ABI: include/llmi_cuda.h).

There is no CPU fallback anywhere in this package: if the library is missing it
is built with nvcc; if it cannot be loaded, or no B200 is present when a compute
entry point is called, the call raises."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from . import _build

_u8p = C.c_void_p
_lib = None

# every symbol include/llmi_cuda.h declares: (restype, argtypes)
_vp, _u32, _u64, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
SIGNATURES = {
    "llmi_init": (_int, [_int]),
    "llmi_shutdown": (_int, []),
    "llmi_last_error": (C.c_char_p, []),
    "llmi_abi_version": (_int, []),
    "llmi_sm_count": (_int, []),
    "llmi_weight_upload": (_int, [_vp, _u32, _u64, _u64, _u64, _u64, C.POINTER(_vp)]),
    "llmi_weight_free": (_int, [_vp]),
    "llmi_weight_dims": (_int, [_vp, C.POINTER(_u32), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64),
                                 C.POINTER(_u64)]),
    "llmi_weight_device_bytes": (_u64, [_vp]),
    "llmi_row_bytes": (_u64, [_u32, _u64]),
    "llmi_registry_get": (_int, [_vp, _u32, _u64, _u64, C.POINTER(_vp)]),
    "llmi_registry_clear": (_int, []),
    "llmi_act_create": (_int, [_u64, C.POINTER(_vp)]),
    "llmi_act_free": (_int, [_vp]),
    "llmi_quantize_q8_0": (_int, [_vp, _u64, _vp, _vp]),
    "llmi_quantize_q8_k": (_int, [_vp, _u64, _vp, _vp]),
    "llmi_round_f16": (_int, [_vp, _u64, _vp, _vp]),
    "llmi_stage_f32": (_int, [_vp, _u64, _vp, _vp]),
    "llmi_act_prepare": (_int, [_vp, _vp, _vp, _vp]),
    "llmi_act_export_q8_0": (_int, [_vp, _vp]),
    "llmi_act_export_q8_k": (_int, [_vp, _vp]),
    "llmi_gemv": (_int, [_vp, _vp, _vp, _vp]),
    "llmi_gemv_batch": (_int, [_vp, _vp, _int, _vp, _vp]),
    "llmi_mat_vec_mul_dev": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "llmi_set_gemv_shape": (_int, [_int, _int]),
    "llmi_set_gemv_ring": (_int, [_int, _int, _int, _int]),
    "llmi_set_prefill_mode": (_int, [_int]),
    "llmi_debug_block_dots": (_int, [_vp, _vp, _vp]),
    "llmi_host_mat_vec_mul": (_int, [_vp, _vp, _u64, _vp, _u64]),
    "llmi_host_quantize_row_q8_0": (_int, [_vp, _u64, _vp]),
    "llmi_host_quantize_row_q8_k": (_int, [_vp, _u64, _vp]),
    "llmi_model_load": (_int, [_vp, _u64, _u32, C.POINTER(_vp)]),
    "llmi_model_load_shard": (_int, [_vp, _u64, _u32, _int, _int, C.POINTER(_vp)]),
    "llmi_shard_range": (_int, [_u64, _int, _int, C.POINTER(_u64), C.POINTER(_u64)]),
    "llmi_model_comm_handle": (_int, [_vp, _vp]),
    "llmi_model_comm_connect": (_int, [_vp, _vp]),
    "llmi_model_comm_error": (_int, [_vp]),
    "llmi_model_comm_reset": (_int, [_vp]),
    "llmi_model_comm_disconnect": (_int, [_vp]),
    "llmi_model_free": (_int, [_vp]),
    "llmi_model_info": (_int, [_vp, C.POINTER(_u32), C.POINTER(_u64)]),
    "llmi_model_forward": (_int, [_vp, _vp, _int, _int, _vp]),
    "llmi_model_decode_greedy": (_int, [_vp, C.c_int32, _int, _int, _vp, C.POINTER(C.c_float)]),
    "llmi_model_last_logits": (_int, [_vp, _vp]),
    "llmi_gemm_tokens": (_int, [_vp, _vp, _u32, _vp, _vp]),
    "llmi_model_launches_per_step": (_int, [_vp]),
    "llmi_model_decode_path": (_int, [_vp]),
    "llmi_model_last_forward_stats": (_int, [_vp, _vp, _vp]),
    "llmi_dev_alloc": (_int, [_u64, C.POINTER(_vp)]),
    "llmi_dev_free": (_int, [_vp]),
    "llmi_h2d": (_int, [_vp, _vp, _u64]),
    "llmi_d2h": (_int, [_vp, _vp, _u64]),
    "llmi_device_sync": (_int, []),
}


def declared_symbols() -> list[str]:
    """Function names declared in include/llmi_cuda.h (parsed from the header)."""
    text = (_build.REPO / "include" / "llmi_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(llmi_[a-z0-9_]+)\s*\(", text)))


def load(build: bool = True) -> C.CDLL:
    """Load (building first if needed) libllmi_cuda.so and type its symbols."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build_cuda()
    if not _build.LIB.exists():
        raise RuntimeError(f"{_build.LIB} is missing: the CUDA extension must be built (python -m "
                           "llm_inference_b200._build); there is no CPU fallback")
    lib = C.CDLL(str(_build.LIB))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class LlmiError(RuntimeError):
    def __init__(self, code: int, msg: str) -> None:
        super().__init__(msg)
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise LlmiError(rc, load().llmi_last_error().decode())
