"""Phase timeline of the persistent decode kernel (dev tool).  Builds the library with -DLLMI_MEGA_TIMING into a
scratch copy, runs a few decode steps and prints, for CTA 0 and the last CTA, per program entry of the LAST step:
wait+prologue time, mat-vec time (us) and the totals per entry kind.
    python tools/mega_timeline.py [workload] [steps]"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import _build  # noqa: E402

_build.build_cuda(force=True, extra=["-DLLMI_MEGA_TIMING"])
import bench  # noqa: E402
from llm_inference_b200 import _lib, ops  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-1b-q4_0"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
os.environ["LLMI_DECODE"] = "mega"
ops.init_ops(1, device=0)
img = bench.build_image(wl)
m = Model(img, max_positions=64 + K + 16)
prompt = (np.arange(16, dtype=np.int32) * 7919 + 13) % m.vocab
first = int(m.forward(prompt, 0).argmax())
m.decode_greedy(first, 16, 4)
toks, ms = m.decode_greedy(first, 16, K)
print(f"{wl}: {ms/K:.4f} ms/token")
L = _lib.load()
buf = np.zeros((2, 1024, 16), np.uint64)
tu = {"gemma-3-1b-q4_0": "q4", "gemma-3-27b-q4_0": "q4", "gemma-3-12b-q8_0": "q8", "gemma-3-4b-q4_k_m": "kq"}[wl]
fn = getattr(L, f"llmi_debug_mega_stamps_{tu}")
fn.argtypes = [C.c_void_p]
fn(buf.ctypes.data)
n_layer = m.n_layer
for c in range(2):
    st = buf[c].astype(np.int64)
    n = int((st[:, 0] > 0).sum())
    t0 = st[0, 0]
    kinds = {}
    print(f"--- CTA {'0' if c == 0 else 'last'}: {n} entries, step total {(st[n-1,2]-t0)/1e3:.1f} us")
    for pc in range(n):
        a, b, e = st[pc, 0], st[pc, 1], st[pc, 2]
        if b == 0:  # attention entry
            name, pro, mv = "attn", 0.0, (e - a) / 1e3
        else:
            name, pro, mv = f"gemv{pc % 5 if n_layer else 0}", (b - a) / 1e3, (e - b) / 1e3
        k = kinds.setdefault(name, [0.0, 0.0, 0])
        k[0] += pro; k[1] += mv; k[2] += 1
        if pc < 12 or pc >= n - 3:
            print(f"  pc {pc:4d} {name:6s} start {(a-t0)/1e3:9.2f} us  wait+prologue {pro:7.2f}  work {mv:7.2f}")
    for name, (pro, mv, cnt) in sorted(kinds.items()):
        print(f"  {name}: n={cnt} wait+prologue {pro:.1f} us (avg {pro/cnt:.2f})  work {mv:.1f} us (avg {mv/cnt:.2f})")
m_cyc = np.zeros((1024, 32), np.int64)
fc = getattr(L, f"llmi_debug_mega_cycles_{tu}")
fc.argtypes = [C.c_void_p]
fc(m_cyc.ctypes.data)
# cycle stamps of CTA 0 / thread 0: 0 entry setup done, 1 prefetch issued, 2 hint seen, 3 norm: static loads issued,
# 4 flagged words in, 5 first block sum, 6 first scale, 7 h updated, 8 second scale, 9 xn in smem, 10 activation
# quantized, 12 prologue barrier passed, 13 first item: addresses ready, 14 items done, 15 barrier, 16 epilogue, 17 barrier
order = [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 12, 13, 21, 22, 23, 24, 14, 15, 16, 17]
for k in (0, 3):
    rows = [m_cyc[pc] for pc in range(5, 120) if pc % 5 == k]
    a = np.array(rows, dtype=np.float64)
    d = [(a[:, order[i + 1]] - a[:, order[i]]).mean() for i in range(len(order) - 1)]
    print(f"gemv{k} cycles between stamps " + " ".join(f"{order[i]}-{order[i+1]}:{v:.0f}" for i, v in enumerate(d)))
m.close()
