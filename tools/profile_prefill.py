"""Dev tool: one prompt through llmi_model_forward (batched prefill), for an ncu launch list.
    python tools/profile_prefill.py [workload] [prompt_len] [layers]"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-1b-q4_0"
p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
layers = int(sys.argv[3]) if len(sys.argv) > 3 else None
dims_name, wt, et = bench.WORKLOADS[wl]
img = synth.build_gemma3_gguf(synth.GEMMA3[dims_name], wt, et, seed=1234, distinct_layers=False, n_layer=layers)
m = Model(img, max_positions=max(256, p + 8))
prompt = (np.arange(p, dtype=np.int32) * 7919 + 13) % m.vocab
for _ in range(2):
    m.forward(prompt, 0)
    print("prefill ms, launches:", m.last_forward_stats())
