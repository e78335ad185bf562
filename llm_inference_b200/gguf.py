"""Minimal GGUF v3 reader mirroring the reference's GGUFFile surface
(gguf.h:81-121, gguf.cpp:258-304, 354-356): header, metadata KV, tensor infos,
and ``get_tensor_data`` as a zero-copy view.  Host-side plumbing only."""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

from .synth import row_bytes

GGUF_MAGIC = 0x46554747
_SCALARS = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}


@dataclass
class TensorInfo:
    """gguf.h:81-87.  shape[0] = n_cols (K, contiguous), shape[1] = n_rows (N)."""
    name: str
    shape: list
    total_elements: int
    tensor_type: int
    tensor_offset: int


class GGUFFile:
    def __init__(self, data) -> None:
        if isinstance(data, (str, bytes)) and not isinstance(data, bytes):
            data = np.memmap(data, dtype=np.uint8, mode="r")
        self.data = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data
        self._pos = 0
        magic, self.version, n_tensors, n_kv = self._unpack("<IIQQ")
        if magic != GGUF_MAGIC:
            raise RuntimeError("Invalid GGUF magic number")  # gguf.cpp:274-276
        self.metadata = {}
        for _ in range(n_kv):
            k = self._str()
            (t,) = self._unpack("<I")
            self.metadata[k] = self._value(t)
        self.tensor_infos = []
        for _ in range(n_tensors):
            name = self._str()
            (nd,) = self._unpack("<I")
            shape = list(self._unpack("<" + "Q" * nd))
            t, off = self._unpack("<IQ")
            self.tensor_infos.append(TensorInfo(name, shape, int(np.prod(shape)), t, off))
        self.data_section_start = (self._pos + 31) // 32 * 32  # gguf.cpp:301-303
        self._by_name = {t.name: t for t in self.tensor_infos}

    def _unpack(self, fmt: str):
        n = struct.calcsize(fmt)
        if self._pos + n > self.data.size:
            raise RuntimeError("Read beyond end of file")
        v = struct.unpack_from(fmt, self.data, self._pos)
        self._pos += n
        return v

    def _str(self) -> str:
        (n,) = self._unpack("<Q")
        if self._pos + n > self.data.size:
            raise RuntimeError("String length exceeds file size")
        s = bytes(self.data[self._pos:self._pos + n]).decode("utf-8", "replace")
        self._pos += n
        return s

    def _value(self, t: int):
        if t in _SCALARS:
            return self._unpack(_SCALARS[t])[0]
        if t == 8:
            return self._str()
        if t == 9:
            et, n = self._unpack("<IQ")
            return [self._value(et) for _ in range(n)]
        raise RuntimeError("Unsupported GGUF value type")

    def tensor(self, name: str) -> TensorInfo:
        return self._by_name[name]

    def get_tensor_data(self, t: TensorInfo) -> np.ndarray:
        """View of the tensor's bytes (gguf.cpp:354-356)."""
        start = self.data_section_start + t.tensor_offset
        if len(t.shape) == 2:
            n = t.shape[1] * row_bytes(t.tensor_type, t.shape[0])
        else:
            n = row_bytes(t.tensor_type, t.total_elements)
        return self.data[start:start + n]
