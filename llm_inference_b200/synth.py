"""Synthetic inputs for the quantized mat-vec path (no network: every weight is
generated).  Random quant blocks with *sane* f16 scales (never random bits, so
no Inf/NaN/subnormal scales), ggml-style quantizers for random-init matrices,
and a GGUF v3 image writer that produces files the reference parser accepts
(layout: gguf.cpp:281-303; metadata keys the loader needs: model.cpp:73-167;
tensor names: model.cpp:174-233).

Block formats (ops.h:11-31, 89-102; sizes verified against the reference
structs): Q4_0 18 B, Q8_0 34 B, Q5_0 22 B per 32; Q4_K 144 B, Q6_K 210 B per 256.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q6_K, BF16 = 0, 1, 2, 6, 8, 12, 14, 30

TYPE_NAMES = {F32: "F32", F16: "F16", Q4_0: "Q4_0", Q5_0: "Q5_0", Q8_0: "Q8_0",
              Q4_K: "Q4_K", Q6_K: "Q6_K", BF16: "BF16"}
_BLOCK = {Q4_0: (32, 18), Q8_0: (32, 34), Q5_0: (32, 22), Q4_K: (256, 144),
          Q6_K: (256, 210), F16: (1, 2), BF16: (1, 2), F32: (1, 4)}


def block_geometry(ggml_type: int) -> tuple[int, int]:
    """(elements per block, bytes per block) of a weight format."""
    return _BLOCK[ggml_type]


def row_bytes(ggml_type: int, n_cols: int) -> int:
    blk, nbytes = _BLOCK[ggml_type]
    if n_cols % blk:
        raise ValueError(f"{TYPE_NAMES[ggml_type]}: K={n_cols} is not a multiple of {blk}")
    return n_cols // blk * nbytes


def algorithmic_bytes(ggml_type: int, n_rows: int, n_cols: int) -> int:
    """reference-format weight bytes + fp32 x in + fp32 o out."""
    return n_rows * row_bytes(ggml_type, n_cols) + 4 * n_cols + 4 * n_rows


def _f16_bits(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float16).view(np.uint16)


def random_blocks(ggml_type: int, n_rows: int, n_cols: int, seed: int = 1) -> np.ndarray:
    """Raw weight bytes [n_rows * row_bytes] of uniformly random quant blocks."""
    rng = np.random.default_rng(seed)
    blk, nbytes = _BLOCK[ggml_type]
    nb = n_rows * (n_cols // blk)
    if n_cols % blk:
        raise ValueError("K not a multiple of the block size")
    if ggml_type in (F16, BF16):
        w = (rng.standard_normal((n_rows, n_cols), dtype=np.float32) * 0.05)
        if ggml_type == F16:
            return _f16_bits(w).view(np.uint8).ravel()
        bits = w.view(np.uint32)
        bits = (bits + 0x7FFF + ((bits >> 16) & 1)) >> 16  # RNE to bf16
        return bits.astype(np.uint16).view(np.uint8).ravel()
    out = np.zeros((nb, nbytes), np.uint8)
    if ggml_type == Q4_0:
        out[:, 0:2] = _f16_bits(rng.uniform(0.005, 0.025, nb)).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = rng.integers(0, 256, (nb, 16), dtype=np.uint8)
    elif ggml_type == Q8_0:
        out[:, 0:2] = _f16_bits(rng.uniform(0.0005, 0.003, nb)).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = rng.integers(-127, 128, (nb, 32), dtype=np.int8).view(np.uint8)
    elif ggml_type == Q5_0:
        out[:, 0:2] = _f16_bits(rng.uniform(0.003, 0.012, nb)).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = rng.integers(0, 256, (nb, 20), dtype=np.uint8)
    elif ggml_type == Q4_K:
        out[:, 0:2] = _f16_bits(rng.uniform(0.0002, 0.0008, nb)).view(np.uint8).reshape(nb, 2)
        out[:, 2:4] = _f16_bits(rng.uniform(0.0002, 0.0008, nb)).view(np.uint8).reshape(nb, 2)
        out[:, 4:] = rng.integers(0, 256, (nb, 140), dtype=np.uint8)
    elif ggml_type == Q6_K:
        out[:, 0:192] = rng.integers(0, 256, (nb, 192), dtype=np.uint8)
        out[:, 192:208] = rng.integers(-64, 64, (nb, 16), dtype=np.int8).view(np.uint8)
        out[:, 208:210] = _f16_bits(rng.uniform(0.00005, 0.0002, nb)).view(np.uint8).reshape(nb, 2)
    else:
        raise ValueError(f"unsupported type {ggml_type}")
    return out.ravel()


def quantize_q4_0(w: np.ndarray) -> np.ndarray:
    """ggml-style Q4_0 of a float matrix [N, K] -> raw bytes.  d = max/-8 where
    max is the signed value of largest magnitude; q = min(15, int(x/d + 8.5));
    low nibble <-> element j, high nibble <-> element j+16."""
    w = np.ascontiguousarray(w, np.float32)
    n, k = w.shape
    b = w.reshape(n * k // 32, 32)
    idx = np.argmax(np.abs(b), axis=1)
    mx = b[np.arange(b.shape[0]), idx]
    d = mx / -8.0
    idv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0).astype(np.float32)
    q = np.minimum(15, (b * idv[:, None] + 8.5).astype(np.int32)).astype(np.uint8)
    out = np.zeros((b.shape[0], 18), np.uint8)
    out[:, 0:2] = _f16_bits(d).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q[:, :16] | (q[:, 16:] << 4)
    return out.ravel()


def quantize_q8_0(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, np.float32)
    n, k = w.shape
    b = w.reshape(n * k // 32, 32)
    d = (np.abs(b).max(axis=1) / 127.0).astype(np.float32)
    idv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0).astype(np.float32)
    q = np.rint(b * idv[:, None]).astype(np.int8)
    out = np.zeros((b.shape[0], 34), np.uint8)
    out[:, 0:2] = _f16_bits(d).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out.ravel()


def random_init_weight(ggml_type: int, n_rows: int, n_cols: int, seed: int, std: float | None = None) -> np.ndarray:
    """Random-init matrix N(0, 1/sqrt(K)) in the requested storage format.
    Q4_0/Q8_0/F16/BF16 are real quantizations of the float matrix; the k-quants
    and Q5_0 use random blocks whose dequantized magnitude is of the same order.
    ``std`` overrides the 1/sqrt(K) target."""
    rng = np.random.default_rng(seed)
    target = np.float32(std if std is not None else 1.0 / np.sqrt(n_cols))
    if ggml_type in (Q4_0, Q8_0, F16, BF16, F32):
        w = rng.standard_normal((n_rows, n_cols), dtype=np.float32) * target
        if ggml_type == Q4_0:
            return quantize_q4_0(w)
        if ggml_type == Q8_0:
            return quantize_q8_0(w)
        if ggml_type == F16:
            return _f16_bits(w).view(np.uint8).ravel()
        if ggml_type == F32:
            return w.view(np.uint8).ravel()
        bits = w.view(np.uint32)
        return ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16).astype(np.uint16).view(np.uint8).ravel()
    raw = random_blocks(ggml_type, n_rows, n_cols, seed).reshape(-1, _BLOCK[ggml_type][1]).copy()
    # rescale the f16 super-scales so dequantized weights are ~N(0, 1/sqrt(K))
    s = target
    nb = raw.shape[0]
    if ggml_type == Q4_K:   # w = d*sc*q - dmin*m, sc,m<=63, q<=15
        raw[:, 0:2] = _f16_bits(rng.uniform(0.5, 1.5, nb) * s / 120.0).view(np.uint8).reshape(nb, 2)
        raw[:, 2:4] = _f16_bits(rng.uniform(0.5, 1.5, nb) * s / 16.0).view(np.uint8).reshape(nb, 2)
    elif ggml_type == Q6_K:  # w = d*sc*(q-32)
        raw[:, 208:210] = _f16_bits(rng.uniform(0.5, 1.5, nb) * s / 600.0).view(np.uint8).reshape(nb, 2)
    elif ggml_type == Q5_0:  # w = d*(q-16)
        raw[:, 0:2] = _f16_bits(rng.uniform(0.5, 1.5, nb) * s / 9.0).view(np.uint8).reshape(nb, 2)
    return raw.ravel()


# --------------------------------------------------------------------- GGUF

_GGUF_MAGIC = 0x46554747
_T_U32, _T_F32, _T_BOOL, _T_STR, _T_ARR = 4, 6, 7, 8, 9


def _s(b: bytearray, s: str) -> None:
    e = s.encode()
    b += struct.pack("<Q", len(e)) + e


@dataclass
class GGUFBuilder:
    """Builds a GGUF v3 image in memory."""

    kv: list = field(default_factory=list)
    tensors: list = field(default_factory=list)

    def add_str(self, k: str, v: str) -> None:
        self.kv.append((k, _T_STR, v))

    def add_u32(self, k: str, v: int) -> None:
        self.kv.append((k, _T_U32, v))

    def add_f32(self, k: str, v: float) -> None:
        self.kv.append((k, _T_F32, v))

    def add_str_array(self, k: str, v: list[str]) -> None:
        self.kv.append((k, _T_ARR, (_T_STR, v)))

    def add_bool_array(self, k: str, v: list[bool]) -> None:
        self.kv.append((k, _T_ARR, (_T_BOOL, v)))

    def add_tensor(self, name: str, ggml_type: int, shape: tuple[int, ...], data: np.ndarray) -> None:
        """shape is GGUF order: shape[0] = K (contiguous), shape[1] = N."""
        self.tensors.append((name, ggml_type, tuple(int(s) for s in shape),
                             np.ascontiguousarray(data).view(np.uint8).ravel()))

    def build(self, align_tensors: int = 32) -> np.ndarray:
        b = bytearray()
        b += struct.pack("<IIQQ", _GGUF_MAGIC, 3, len(self.tensors), len(self.kv))
        for k, t, v in self.kv:
            _s(b, k)
            b += struct.pack("<I", t)
            if t == _T_STR:
                _s(b, v)
            elif t == _T_U32:
                b += struct.pack("<I", v)
            elif t == _T_F32:
                b += struct.pack("<f", v)
            elif t == _T_ARR:
                et, items = v
                b += struct.pack("<IQ", et, len(items))
                for it in items:
                    if et == _T_STR:
                        _s(b, it)
                    elif et == _T_BOOL:
                        b += struct.pack("<B", 1 if it else 0)
        offs, off = [], 0
        for _, _, _, data in self.tensors:
            off = (off + align_tensors - 1) // align_tensors * align_tensors
            offs.append(off)
            off += data.size
        for (name, t, shape, _), o in zip(self.tensors, offs):
            _s(b, name)
            b += struct.pack("<I", len(shape))
            for d in shape:
                b += struct.pack("<Q", d)
            b += struct.pack("<IQ", t, o)
        start = (len(b) + 31) // 32 * 32  # the reference hard-codes 32 (gguf.cpp:301-303)
        img = np.zeros(start + off, np.uint8)
        img[: len(b)] = np.frombuffer(bytes(b), np.uint8)
        for (_, _, _, data), o in zip(self.tensors, offs):
            img[start + o: start + o + data.size] = data
        return img


@dataclass
class GemmaDims:
    name: str
    n_layer: int
    n_embd: int
    n_ff: int
    n_head: int
    n_head_kv: int
    head_dim: int
    vocab: int


GEMMA3 = {
    "gemma-3-1b": GemmaDims("gemma-3-1b", 26, 1152, 6912, 4, 1, 256, 262144),
    "gemma-3-4b": GemmaDims("gemma-3-4b", 34, 2560, 10240, 8, 4, 256, 262208),
    "gemma-3-12b": GemmaDims("gemma-3-12b", 48, 3840, 15360, 16, 8, 256, 262208),
    "gemma-3-27b": GemmaDims("gemma-3-27b", 62, 5376, 21504, 32, 16, 128, 262208),
}


def q4_k_m_layer_types(n_layer: int) -> list[dict]:
    """llama.cpp's Q4_K_M recipe: attn_v and ffn_down get Q6_K on the first and
    last eighth of the layers and on every third layer in between; the rest Q4_K."""
    out = []
    for i in range(n_layer):
        more = i < n_layer // 8 or i >= 7 * n_layer // 8 or (i - n_layer // 8) % 3 == 2
        out.append({"attn_q": Q4_K, "attn_k": Q4_K, "attn_output": Q4_K, "ffn_gate": Q4_K,
                    "ffn_up": Q4_K, "attn_v": Q6_K if more else Q4_K,
                    "ffn_down": Q6_K if more else Q4_K})
    return out


def build_gemma3_gguf(dims: GemmaDims, weight_type: int | str = Q4_0, embd_type: int = F16,
                      seed: int = 1234, n_layer: int | None = None, vocab: int | None = None,
                      distinct_layers: bool = True, embd_std: float = 1.0) -> np.ndarray:
    """Random-init Gemma-3 GGUF image.  ``weight_type`` is a ggml
    type id or "q4_k_m".  ``n_layer``/``vocab`` shrink the model for tests.
    With ``distinct_layers=False`` layers >= 1 reuse layer 0's quantized bytes
    (generation time only; every layer still has its own tensor data).
    ``embd_std`` < 1 makes the (tied) embeddings small next to the layer outputs,
    so greedy decoding depends on the whole network instead of echoing its input."""
    L = n_layer if n_layer is not None else dims.n_layer
    V = vocab if vocab is not None else dims.vocab
    E, F, H, HK, D = dims.n_embd, dims.n_ff, dims.n_head, dims.n_head_kv, dims.head_dim
    g = GGUFBuilder()
    a = "gemma3"
    g.add_str("general.architecture", a)
    g.add_u32(f"{a}.block_count", L)
    g.add_u32(f"{a}.embedding_length", E)
    g.add_u32(f"{a}.feed_forward_length", F)
    g.add_u32(f"{a}.attention.head_count", H)
    g.add_u32(f"{a}.attention.head_count_kv", HK)
    g.add_f32(f"{a}.attention.layer_norm_rms_epsilon", 1e-6)
    g.add_f32(f"{a}.rope.freq_base", 1e6)
    g.add_u32(f"{a}.attention.key_length", D)
    g.add_u32(f"{a}.attention.value_length", D)
    # Model::load_vocabulary needs the key (model.cpp:1054); forward() never
    # touches the strings, so a token table of a few entries is enough.
    g.add_str_array("tokenizer.ggml.tokens", ["<pad>", "<eos>", "<bos>", "<unk>"])
    g.add_u32("tokenizer.ggml.bos_token_id", 2)

    rng = np.random.default_rng(seed)
    if embd_type == F16:
        emb = _f16_bits(rng.standard_normal((V, E), dtype=np.float32) * np.float32(embd_std)).view(np.uint8).ravel()
    else:
        emb = random_init_weight(embd_type, V, E, seed + 7, std=embd_std)  # embeddings ~N(0, embd_std)
    g.add_tensor("token_embd.weight", embd_type, (E, V), emb)
    g.add_tensor("output_norm.weight", F32, (E,),
                 (1 + 0.1 * rng.standard_normal(E)).astype(np.float32))
    per_layer = q4_k_m_layer_types(L) if weight_type == "q4_k_m" else None
    shapes = {"attn_q": (E, H * D), "attn_k": (E, HK * D), "attn_v": (E, HK * D),
              "attn_output": (H * D, E), "ffn_gate": (E, F), "ffn_up": (E, F), "ffn_down": (F, E)}
    cache: dict = {}
    for i in range(L):
        for nm, n in (("attn_norm", E), ("ffn_norm", E), ("post_attention_norm", E),
                      ("post_ffw_norm", E), ("attn_q_norm", D), ("attn_k_norm", D)):
            g.add_tensor(f"blk.{i}.{nm}.weight", F32, (n,),
                         (1 + 0.1 * rng.standard_normal(n)).astype(np.float32))
        for nm, (k, n) in shapes.items():
            t = per_layer[i][nm] if per_layer else weight_type
            key = (nm, t)
            if distinct_layers or key not in cache:
                cache[key] = random_init_weight(t, n, k, seed + 1000 * i + 17 * list(shapes).index(nm))
            g.add_tensor(f"blk.{i}.{nm}.weight", t, (k, n), cache[key])
    return g.build()
