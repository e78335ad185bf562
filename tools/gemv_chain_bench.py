"""Dev tool: where do the ~2.3 us of a small dependent mat-vec go?  Builds the library with
-DLLMI_GEMV_TIMING (on the box it runs on), replays a CUDA graph of dependent mat-vecs of one shape
(PDL on, like the decode step) and prints the %globaltimer stamps of CTA 0 / thread 0 per launch:
entry -> pdl_wait returned -> activation landed -> all items folded -> outputs stored -> next entry.
    python tools/gemv_chain_bench.py [K N] ..."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import _build  # noqa: E402

_build.build_cuda(force=True, extra=["-DLLMI_GEMV_TIMING"])
import torch  # noqa: E402
from llm_inference_b200 import _lib, ops, synth  # noqa: E402

ops.init_ops(1)
L = _lib.load()
shapes = [(1152, 6912), (1024, 1152), (6912, 1152), (1152, 256), (5376, 21504)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for k, n in shapes:
    copies = 24
    raw = synth.random_blocks(synth.Q4_0, n, k, seed=1)
    ws = [ops.DeviceWeight(raw, synth.Q4_0, k, n) for _ in range(copies)]
    x = ops.DeviceVector(k, np.random.default_rng(0).standard_normal(k).astype(np.float32))
    o = ops.DeviceVector(n)
    act = ops.Activation(k)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        act.prepare(ws[0], x, s.cuda_stream)
        for w in ws:
            ops.gemv(w, act, o, s.cuda_stream)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for w in ws:
                ops.gemv(w, act, o, s.cuda_stream)
        for _ in range(3):
            g.replay()
        s.synchronize()
        n_l = C.c_uint()
        L.llmi_debug_gemv_stamps(None, C.byref(n_l), 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s)
        s.synchronize()
    st = np.zeros((256, 6), np.uint64)
    L.llmi_debug_gemv_stamps(st.ctypes.data, C.byref(n_l), 1)
    st = st[:copies].astype(np.int64)
    d = np.stack([st[:, 1] - st[:, 0], st[:, 2] - st[:, 1], st[:, 3] - st[:, 2], st[:, 4] - st[:, 3]], 1)[4:]
    gap = (st[1:, 0] - st[:-1, 4])[4:]
    per = (st[-1, 4] - st[4, 0]) / (copies - 5)
    print(f"Q4_0 {k}->{n}: {e0.elapsed_time(e1) * 1e3 / copies:.2f} us/launch by events, {per / 1e3:.2f} us entry-to-entry; "
          f"median ns: entry->pdl_wait {np.median(d[:, 0]):.0f}, ->activation landed {np.median(d[:, 1]):.0f}, "
          f"->items folded {np.median(d[:, 2]):.0f}, ->stored {np.median(d[:, 3]):.0f}, "
          f"end->next entry {np.median(gap):.0f}")
    for w in ws:
        w.close()
