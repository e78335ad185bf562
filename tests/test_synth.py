"""Host-side logic: the synthetic GGUF writer produces images both our reader
and the reference's own parser/model accept, with the shapes SURVEY §8(d) lists."""
import numpy as np
import pytest

from llm_inference_b200 import synth
from llm_inference_b200.gguf import GGUFFile


def test_block_sizes_and_algorithmic_bytes():
    assert synth.row_bytes(synth.Q4_0, 1152) == 648 and synth.row_bytes(synth.Q6_K, 2560) == 2100
    # config 1 = 4,478,976 + 4,608 + 27,648
    assert synth.algorithmic_bytes(synth.Q4_0, 6912, 1152) == 4_511_232
    with pytest.raises(ValueError):
        synth.row_bytes(synth.Q4_K, 1152)


def test_quantizers_round_trip():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((8, 64)).astype(np.float32)
    q = synth.quantize_q4_0(w).reshape(-1, 18)
    d = q[:, :2].copy().view(np.float16).astype(np.float32)
    lo = (q[:, 2:] & 0xF).astype(np.float32) - 8
    hi = (q[:, 2:] >> 4).astype(np.float32) - 8
    deq = np.concatenate([lo, hi], axis=1) * d
    assert np.abs(deq.reshape(8, 64) - w).max() <= np.abs(w).max() / 7
    q8 = synth.quantize_q8_0(w).reshape(-1, 34)
    d8 = q8[:, :2].copy().view(np.float16).astype(np.float32)
    deq8 = q8[:, 2:].view(np.int8).astype(np.float32) * d8
    assert np.abs(deq8.reshape(8, 64) - w).max() <= np.abs(w).max() / 100


def test_gguf_writer_reader_round_trip():
    dims = synth.GemmaDims("tiny", 2, 256, 512, 4, 2, 64, 96)
    img = synth.build_gemma3_gguf(dims, "q4_k_m", synth.Q6_K, seed=3)
    f = GGUFFile(img)
    assert f.metadata["general.architecture"] == "gemma3"
    assert f.metadata["gemma3.block_count"] == 2 and f.metadata["gemma3.attention.key_length"] == 64
    t = f.tensor("blk.1.ffn_down.weight")
    assert t.shape == [512, 256] and t.tensor_type in (synth.Q4_K, synth.Q6_K)
    assert f.get_tensor_data(t).size == 256 * synth.row_bytes(t.tensor_type, 512)
    e = f.tensor("token_embd.weight")
    assert e.tensor_type == synth.Q6_K and e.shape == [256, 96]
    names = {ti.name for ti in f.tensor_infos}
    for nm in ("attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up", "ffn_down", "attn_norm",
               "ffn_norm", "post_attention_norm", "post_ffw_norm", "attn_q_norm", "attn_k_norm"):
        assert f"blk.0.{nm}.weight" in names
    assert all((f.data_section_start + ti.tensor_offset) % 32 == 0 for ti in f.tensor_infos)


def test_reference_model_accepts_the_synthetic_gguf(ref):
    dims = synth.GemmaDims("tiny", 2, 256, 512, 4, 2, 64, 96)
    for wt, et in ((synth.Q4_0, synth.F16), ("q4_k_m", synth.Q6_K), (synth.Q8_0, synth.Q8_0)):
        m = ref.model(synth.build_gemma3_gguf(dims, wt, et, seed=5))
        logits = m.forward([1, 2, 3], 0)
        assert logits.shape == (96,) and np.isfinite(logits).all() and np.abs(logits).max() > 1e-3
        nxt = m.forward([int(logits.argmax())], 3)
        assert np.isfinite(nxt).all()
        m.close()


def test_gemma3_shapes_match_survey():
    d = synth.GEMMA3["gemma-3-27b"]
    assert (d.n_layer, d.n_embd, d.n_ff, d.n_head, d.n_head_kv, d.head_dim, d.vocab) == (62, 5376, 21504, 32, 16, 128, 262208)
    d = synth.GEMMA3["gemma-3-1b"]
    assert (d.n_layer, d.n_embd, d.n_ff, d.n_head * d.head_dim, d.vocab) == (26, 1152, 6912, 1024, 262144)
