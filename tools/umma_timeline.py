"""Dev tool: per-role clock64 timeline of CTA (0,0,0) of gemm_umma_kernel (built here with -DLLMI_UMMA_TIMING).
Rows: producer stage issue, MMA sees stage full, MMA issued block, epilogue warp 4: dots of block landed / next
block's TMEM load issued (= its tfull seen) / block folded.    python tools/umma_timeline.py [q8|q4] [tokens]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import _build  # noqa: E402

_build.build_cuda(force=True, extra=["-DLLMI_UMMA_TIMING"])
import torch  # noqa: E402
from llm_inference_b200 import _lib, ops, synth  # noqa: E402

ops.init_ops(1, 0)
L = _lib.load()
q8 = (sys.argv[1] if len(sys.argv) > 1 else "q8") == "q8"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t, k, n = (synth.Q8_0, 3840, 15360) if q8 else (synth.Q4_0, 5376, 21504)
w = ops.DeviceWeight(synth.random_blocks(t, n, k, seed=1), t, k, n)
x = ops.DeviceVector(m * k, np.random.default_rng(0).standard_normal(m * k).astype(np.float32))
o = ops.DeviceVector(m * n)
for _ in range(3):
    ops.gemm_tokens(w, x, m, o)
torch.cuda.synchronize()
st = np.zeros((10, 160), np.int64)
L.llmi_debug_umma_stamps.argtypes = [C.c_void_p]
L.llmi_debug_umma_stamps(st.ctypes.data)
t0 = st[0, 0]
names = ["producer issues stage", "MMA sees stage full", "MMA issued block", "epi: dots landed", "epi: next tfull seen",
         "epi: block folded", "epi: block done (+stores)", "epi: stage top", "epi: scales requested", "epi: dfull seen"]
for r, nm in enumerate(names):
    row = st[r] - t0
    cnt = 12 if r in (0, 1, 7, 8, 9) else 40
    print(f"{nm:24s}", " ".join(str(int(v)) for v in row[:cnt]))
for r in (2, 3, 5):
    d = np.diff(st[r, 8:120])
    print(f"{names[r]:24s} median delta per block: {np.median(d):.0f} cycles")
