"""Dev tool: launch one GEMV case a few times (no graph) so ncu can capture it.
Usage: python tools/ncu_case.py Q4_0 5376 21504 [WxS] [launches]"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import ops, synth  # noqa: E402

name, k, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
shape = tuple(int(v) for v in sys.argv[4].split("x")) if len(sys.argv) > 4 else (0, 0)
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 8
t = {v: kk for kk, v in synth.TYPE_NAMES.items()}[name]
ops.init_ops(1, 0)
copies = 4
raw = synth.random_blocks(t, n, k, seed=1)
ws = [ops.DeviceWeight(raw, t, k, n) for _ in range(copies)]
x = ops.DeviceVector(k, np.random.default_rng(0).standard_normal(k).astype(np.float32))
o = ops.DeviceVector(n)
act = ops.Activation(k)
ops.set_gemv_shape(*shape)
act.prepare(ws[0], x)
for i in range(launches):
    ops.gemv(ws[i % copies], act, o)
ops.device_sync()
print("done", float(np.abs(o.get()).max()))
