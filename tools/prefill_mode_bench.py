"""Dev tool: whole-prompt prefill (llmi_model_forward) in the exact and in the throughput mode (LLMI_PREFILL=fast), and
how far the greedy continuation of the fast-mode prompt agrees with the exact one.
    python tools/prefill_mode_bench.py [workload] [prompt_len] [layers]"""
import json
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-1b-q4_0"
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 512
layers = int(sys.argv[3]) if len(sys.argv) > 3 else None
dims_name, wt, et = bench.WORKLOADS[wl]
img = synth.build_gemma3_gguf(synth.GEMMA3[dims_name], wt, et, seed=1234, distinct_layers=False, n_layer=layers)
res = {}
for mode in ("exact", "fast"):
    if mode == "fast":
        os.environ["LLMI_PREFILL"] = "fast"
    else:
        os.environ.pop("LLMI_PREFILL", None)
    m = Model(img, max_positions=n_p + 48)
    prompt = ((np.arange(n_p, dtype=np.int64) * 7919 + 13) % m.vocab).astype(np.int32)
    lg = m.forward(prompt, 0)
    best = None
    for _ in range(2):
        lg = m.forward(prompt, 0)
        ms, launches = m.last_forward_stats()
        best = ms if best is None else min(best, ms)
    first = int(lg.argmax())
    toks, _ = m.decode_greedy(first, n_p, 32)
    res[mode] = {"ms": best, "tok_s": n_p / best * 1e3, "launches": launches, "first": first, "tokens": [int(t) for t in toks],
                 "logits": lg.copy()}
    m.close()
os.environ.pop("LLMI_PREFILL", None)
from llm_inference_b200 import ops  # noqa: E402
ops.set_prefill_mode(False)
le, lf = res["exact"].pop("logits"), res["fast"].pop("logits")
agree = 0
for a, b in zip([res["exact"]["first"]] + res["exact"]["tokens"], [res["fast"]["first"]] + res["fast"]["tokens"]):
    if a != b:
        break
    agree += 1
print(json.dumps({"workload": wl, "prompt": n_p, "exact_ms": round(res["exact"]["ms"], 2), "fast_ms": round(res["fast"]["ms"], 2),
                  "exact_tok_s": round(res["exact"]["tok_s"]), "fast_tok_s": round(res["fast"]["tok_s"]),
                  "speedup": round(res["exact"]["ms"] / res["fast"]["ms"], 2),
                  "logits_max_err_rel_to_max": float(np.abs(le - lf).max() / np.abs(le).max()),
                  "greedy_tokens_agree_leading": agree, "of": 33,
                  "exact_tokens_head": res["exact"]["tokens"][:8], "fast_tokens_head": res["fast"]["tokens"][:8]}))
