"""Dev tool: a handful of decode steps of a workload with a tiny prompt, so that
`ncu --metrics gpu__time_duration.sum` stays cheap.  Usage:
    python tools/profile_step.py [workload] [prompt_len] [steps] [layers]"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-1b-q4_0"
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
layers = int(sys.argv[4]) if len(sys.argv) > 4 else None
dims_name, wt, et = bench.WORKLOADS[wl]
img = synth.build_gemma3_gguf(synth.GEMMA3[dims_name], wt, et, seed=1234, distinct_layers=False, n_layer=layers)
m = Model(img, max_positions=256)
prompt = (np.arange(p, dtype=np.int32) * 7919 + 13) % m.vocab
lg = m.forward(prompt, 0)
toks, ms = m.decode_greedy(int(lg.argmax()), p, steps)
print("decode ms/step", ms / steps, "launches/step", m.launches_per_step, toks[:4])
