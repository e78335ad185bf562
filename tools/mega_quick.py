"""Dev check of the persistent decode kernel: tokens vs the per-launch path + ms/token, one workload.
    python tools/mega_quick.py [workload] [steps]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from llm_inference_b200 import ops  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-1b-q4_0"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ops.init_ops(1, device=0)
t0 = time.time()
img = bench.build_image(wl)
print(f"image {len(img)/1e9:.2f} GB in {time.time()-t0:.1f}s", flush=True)
res = {}
modes = os.environ.get("QUICK_MODES", "legacy,mega" if os.environ.get("QUICK_LEGACY", "1") == "1" else "mega").split(",")
for mode in modes:
    if mode == "legacy":
        os.environ["LLMI_DECODE"] = "legacy"
    else:
        os.environ["LLMI_DECODE"] = "mega"
    t0 = time.time()
    m = Model(img, max_positions=64 + K + 16)
    t_load = time.time() - t0
    prompt = (np.arange(16, dtype=np.int32) * 7919 + 13) % m.vocab
    first = int(m.forward(prompt, 0).argmax())
    m.decode_greedy(first, 16, 4)
    best = None
    for _ in range(3):
        toks, ms = m.decode_greedy(first, 16, K)
        best = ms if best is None else min(best, ms)
    res[mode] = toks
    print(f"{wl} {mode}: load {t_load:.1f}s  {best/K:.4f} ms/token  {K/best*1e3:.1f} tok/s  launches/step "
          f"{m.launches_per_step}  tokens {list(toks[:6])}", flush=True)
    m.close()
if len(res) == 2:
    print("tokens equal:", bool(np.array_equal(res["legacy"], res["mega"])))
