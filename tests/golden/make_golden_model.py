"""Generates tests/golden/model_golden.npz from the COMPILED REFERENCE
(oracle/_ref/libref.so): a tiny random-init Gemma-3 GGUF image (stored in the
file, so the test needs neither numpy RNG stability nor /root/reference), a
prompt, the reference's logits after the prefill and along a greedy decode.
Run in the dev container only:  python tests/golden/make_golden_model.py"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import synth  # noqa: E402
from oracle.binding import Ref, ensure_ref  # noqa: E402


def main() -> None:
    assert ensure_ref()
    R = Ref(n_threads=1)
    out = {}
    cases = {
        "q4_0": (synth.GemmaDims("tiny-q4_0", 2, 128, 256, 2, 1, 64, 64), synth.Q4_0, synth.F16),
        "q4_k_m": (synth.GemmaDims("tiny-q4_k_m", 1, 256, 512, 4, 2, 64, 48), "q4_k_m", synth.Q6_K),
    }
    for name, (dims, wt, et) in cases.items():
        img = synth.build_gemma3_gguf(dims, wt, et, seed=2026, embd_std=0.004)
        m = R.model(img)
        prompt = np.array([3, 11, 7, 30, 2, 19], np.int32)
        logits = [m.forward(prompt, 0)]
        toks, pos = [], len(prompt)
        for _ in range(12):
            t = int(logits[-1].argmax())
            toks.append(t)
            logits.append(m.forward([t], pos))
            pos += 1
        m.close()
        out[f"{name}_image"] = img
        out[f"{name}_prompt"] = prompt
        out[f"{name}_tokens"] = np.array(toks, np.int32)
        out[f"{name}_logits"] = np.stack(logits)
        srt = np.sort(np.stack(logits), axis=1)
        print(name, img.size, "bytes; tokens", toks, "min top1-top2 margin", float((srt[:, -1] - srt[:, -2]).min()))
    path = Path(__file__).with_name("model_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
