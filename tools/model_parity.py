"""Dev tool: per-step logits error of the device-resident forward vs the compiled
reference (oracle/_ref) on synthetic models.  Usage: python tools/model_parity.py"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402
from oracle import binding  # noqa: E402

R = binding.Ref(n_threads=8)
cases = [
    ("tiny q4_0 std.004", synth.GemmaDims("t", 2, 128, 256, 2, 1, 64, 64), synth.Q4_0, synth.F16, 0.004),
    ("tiny q4_0 std1", synth.GemmaDims("t", 2, 128, 256, 2, 1, 64, 64), synth.Q4_0, synth.F16, 1.0),
    ("small q4_0 std.004", synth.GemmaDims("s", 3, 512, 1024, 4, 2, 128, 512), synth.Q4_0, synth.F16, 0.004),
    ("small q4_0 std1", synth.GemmaDims("s", 3, 512, 1024, 4, 2, 128, 512), synth.Q4_0, synth.F16, 1.0),
    ("small q4_k_m std.004", synth.GemmaDims("s", 3, 512, 1024, 4, 2, 128, 512), "q4_k_m", synth.Q6_K, 0.004),
    ("small q8_0 std.004", synth.GemmaDims("s", 3, 512, 1024, 4, 2, 128, 512), synth.Q8_0, synth.Q8_0, 0.004),
]
for name, dims, wt, et, std in cases:
    img = synth.build_gemma3_gguf(dims, wt, et, seed=7, embd_std=std)
    ref, m = R.model(img), Model(img, max_positions=128)
    prompt = np.arange(3, 11, dtype=np.int32)
    errs, same = [], 0
    for n in range(1, len(prompt) + 1):  # prefill lengths: isolates the position at which errors appear
        ref2, m2 = R.model(img), Model(img, max_positions=128)
        a, b = ref2.forward(prompt[:n], 0), m2.forward(prompt[:n], 0)
        errs.append(float(np.abs(a - b).max() / np.abs(a).max()))
        ref2.close(); m2.close()
    a, b = ref.forward(prompt, 0), m.forward(prompt, 0)
    pos, derr = len(prompt), []
    for step in range(24):
        ta, tb = int(a.argmax()), int(b.argmax())
        same += ta == tb
        a, b = ref.forward([ta], pos), m.forward([ta], pos)
        derr.append(float(np.abs(a - b).max() / np.abs(a).max()))
        pos += 1
    print(f"{name}: prefill errs {[f'{e:.1e}' for e in errs]}\n   decode errs {[f'{e:.1e}' for e in derr]} tokens same {same}/24")
    ref.close(); m.close()
