"""Dev tool: the throughput prefill's GEMM token tile (LLMI_FAST_TNF = 256 | 128 | unset: chosen by grid fill) — whole-prompt
time and whether the logits are bit-identical across the settings (a row-sharded model may pick another tile than the
single-GPU model it must equal).
    python tools/fast_tnf_ab.py [workload] [prompt_len] [layers]"""
import json
import os
import sys
import zlib
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-27b-q4_0"
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
layers = int(sys.argv[3]) if len(sys.argv) > 3 else None
dims_name, wt, et = bench.WORKLOADS[wl]
img = synth.build_gemma3_gguf(synth.GEMMA3[dims_name], wt, et, seed=1234, distinct_layers=False, n_layer=layers)
os.environ["LLMI_PREFILL"] = "fast"
out = {"workload": wl, "prompt": n_p}
for tnf in ("256", "", "128"):
    if tnf:
        os.environ["LLMI_FAST_TNF"] = tnf
    else:
        os.environ.pop("LLMI_FAST_TNF", None)
    m = Model(img, max_positions=n_p + 48)
    prompt = ((np.arange(n_p, dtype=np.int64) * 7919 + 13) % m.vocab).astype(np.int32)
    lg = m.forward(prompt, 0)
    best = None
    for _ in range(3):
        lg = m.forward(prompt, 0)
        ms, _ = m.last_forward_stats()
        best = ms if best is None else min(best, ms)
    out[tnf or "auto"] = {"ms": round(best, 2), "logits_crc": zlib.crc32(lg.tobytes())}
    m.close()
out["bitwise_equal"] = len({v["logits_crc"] for k, v in out.items() if isinstance(v, dict)}) == 1
print(json.dumps(out))
