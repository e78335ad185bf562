"""Python mirror of the reference's Model (model.h:72-117) on the device-resident
forward (include/llmi_cuda.h, llmi_model_*): ``Model(gguf_image)``,
``forward(tokens, pos) -> logits`` of the last token, plus the greedy
generation loop of main.cpp run entirely on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, ops


class Model:
    """``world > 1``: this process holds rank ``rank``'s row shard of every matrix; call
    :meth:`connect` (after every rank has constructed its Model) before the first forward / decode.  All ranks
    must then make the same calls; logits and tokens are bit-identical to the single-GPU model."""

    def __init__(self, gguf_image, max_positions: int = 4096, device: int = 0, world: int = 1, rank: int = 0) -> None:
        ops.init_ops(1, device)
        L = _lib.load()
        img = np.ascontiguousarray(np.frombuffer(gguf_image, np.uint8) if isinstance(gguf_image, (bytes, bytearray))
                                   else gguf_image, np.uint8)
        h = C.c_void_p()
        _lib.check(L.llmi_model_load_shard(img.ctypes.data, img.size, max_positions, world, rank, C.byref(h)))
        self.h = h
        self.world, self.rank = world, rank
        dims = (C.c_uint32 * 8)()
        wb = C.c_uint64()
        _lib.check(L.llmi_model_info(h, dims, C.byref(wb)))
        (self.n_layer, self.n_embd, self.n_ff, self.n_head, self.n_head_kv, self.head_dim, self.vocab,
         self.max_positions) = (int(v) for v in dims)
        self.weight_bytes = int(wb.value)

    def comm_handle(self) -> bytes:
        """This rank's exchange buffer as a CUDA IPC handle (64 bytes)."""
        buf = (C.c_uint8 * 64)()
        _lib.check(_lib.load().llmi_model_comm_handle(self.h, buf))
        return bytes(buf)

    def connect(self, group=None) -> None:
        """Maps the peers' exchange buffers.  The 64-byte handles are all-gathered through ``torch.distributed``
        (any backend — this is the only collective a sharded model ever calls); afterwards the mat-vec kernels
        write their rows straight into the peers' buffers."""
        if self.world == 1:
            return
        import torch.distributed as dist

        handles = [None] * self.world
        dist.all_gather_object(handles, self.comm_handle(), group=group)
        blob = b"".join(handles)
        _lib.check(_lib.load().llmi_model_comm_connect(self.h, blob))
        dist.barrier(group=group)

    @property
    def comm_error(self) -> bool:
        return bool(_lib.load().llmi_model_comm_error(self.h))

    def forward(self, tokens, pos: int) -> np.ndarray:
        """Model::forward(tokens, pos): logits of the last token (model.cpp:706-1048)."""
        tk = np.ascontiguousarray(tokens, np.int32)
        logits = np.empty(self.vocab, np.float32)
        _lib.check(_lib.load().llmi_model_forward(self.h, tk.ctypes.data, tk.size, pos, logits.ctypes.data))
        return logits

    def decode_greedy(self, first_token: int, pos: int, n_steps: int):
        """(generated token ids, device milliseconds) — main.cpp:172-221 on the device."""
        out = np.zeros(n_steps, np.int32)
        ms = C.c_float()
        _lib.check(_lib.load().llmi_model_decode_greedy(self.h, int(first_token), pos, n_steps, out.ctypes.data,
                                                        C.byref(ms)))
        return out, float(ms.value)

    def last_logits(self) -> np.ndarray:
        logits = np.empty(self.vocab, np.float32)
        _lib.check(_lib.load().llmi_model_last_logits(self.h, logits.ctypes.data))
        return logits

    def last_forward_stats(self):
        """(device milliseconds, kernel launches) of the last forward() call."""
        ms, n = C.c_float(), C.c_int()
        _lib.check(_lib.load().llmi_model_last_forward_stats(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def launches_per_step(self) -> int:
        return int(_lib.load().llmi_model_launches_per_step(self.h))

    @property
    def persistent(self) -> bool:
        """True when this model decodes with the persistent kernel (LLMI_DECODE=mega at load time)."""
        return bool(_lib.load().llmi_model_decode_path(self.h))

    def disconnect(self, group=None) -> None:
        """Tear-down of a sharded model, first half: unmap the peers' buffers, then wait for every rank to have done
        so (a rank must not free a buffer a peer still has mapped).  :meth:`close` afterwards."""
        if self.world == 1 or not self.h:
            return
        import torch.distributed as dist

        _lib.check(_lib.load().llmi_model_comm_disconnect(self.h))
        dist.barrier(group=group)

    def close(self) -> None:
        if self.h:
            _lib.load().llmi_model_free(self.h)
            self.h = None
