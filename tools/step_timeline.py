"""Dev tool: where does a decode step's time go, launch by launch?  Builds the library with -DLLMI_TIMELINE (on the
box it runs on: every kernel's CTA 0 / thread 0 stamps %globaltimer at entry, when griddepcontrol.wait returns and at
its last statement), runs a few greedy decode steps of a workload and prints, for one step in the middle, every
launch of one layer and the per-kind totals of the whole step.  'stage' of launch k = wait(k+1) - wait(k): the time
the step's critical path spends in it, hand-over to the next kernel included.
    python tools/step_timeline.py [workload] [prompt_len] [steps] [layers]
Environment knobs of the library apply (LLMI_GEMV_RING=..., LLMI_NO_PDL=1, ...)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import _build  # noqa: E402

_build.build_cuda(force=True, extra=["-DLLMI_TIMELINE"])
import bench  # noqa: E402
from llm_inference_b200 import _lib, synth  # noqa: E402
from llm_inference_b200.model import Model  # noqa: E402

KIND = {1: "gemv_slab", 2: "gemv_ring", 3: "norm_act", 4: "attention", 5: "geglu", 6: "embed", 7: "finish"}
wl = sys.argv[1] if len(sys.argv) > 1 else "gemma-3-27b-q4_0"
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
layers = int(sys.argv[4]) if len(sys.argv) > 4 else None
dims_name, wt, et = bench.WORKLOADS[wl]
img = synth.build_gemma3_gguf(synth.GEMMA3[dims_name], wt, et, seed=1234, distinct_layers=False, n_layer=layers)
m = Model(img, max_positions=256)
prompt = (np.arange(p, dtype=np.int32) * 7919 + 13) % m.vocab
lg = m.forward(prompt, 0)
tok = int(lg.argmax())
m.decode_greedy(tok, p, 2)  # warm-up + graph capture
L = C.CDLL(str(_build.LIB))
rec = np.dtype([("t", np.uint64, 10), ("kind", np.uint32), ("ctas", np.uint32)])


def dump(reset):
    out = []
    for fn in ("llmi_debug_timeline_gemv", "llmi_debug_timeline_glue"):
        buf = np.zeros(8192, rec)
        n = C.c_uint()
        getattr(L, fn)(buf.ctypes.data_as(C.c_void_p), C.byref(n), reset)
        out.append(buf[: min(n.value, 8192)])
    return np.concatenate(out)


dump(1)
toks, ms = m.decode_greedy(tok, p + 2, steps)
r = dump(1)
r = r[np.argsort(r["t"][:, 0])]
per = len(r) // steps
print(f"{wl}: {ms / steps:.4f} ms/step by events, {len(r)} records = {per} launches/step")
a = r[per * (steps // 2): per * (steps // 2 + 1)]
t = a["t"].astype(np.int64)
wait = np.where(t[:, 1] > 0, t[:, 1], t[:, 0])
stage = np.append(wait[1:] - wait[:-1], t[-1, 2] - wait[-1])
print(f"step span {(t[-1, 2] - t[0, 0]) / 1e3:.1f} us")
print("per kind: launches, sum of stage us, mean stage us, mean (entry->wait) us, mean (wait->CTA0 end) us")
for k in sorted(set(a["kind"])):
    sel = a["kind"] == k
    print(f"  {KIND.get(int(k), k):10s} {sel.sum():4d} {stage[sel].sum() / 1e3:9.1f} {stage[sel].mean() / 1e3:8.2f} "
          f"{(wait[sel] - t[sel, 0]).mean() / 1e3:8.2f} {(t[sel, 2] - wait[sel]).mean() / 1e3:8.2f}")
# one layer from the middle of the step (8 launches)
mid = per // 2
mid -= (mid - 2) % 8 if per > 20 else 0
print("one layer (kind, ctas, entry rel. us, wait rel. us, cta0-end rel. us, stage us):")
base = wait[mid]
for i in range(mid, min(mid + 9, per)):
    print(f"  {KIND.get(int(a['kind'][i]), a['kind'][i]):10s} {a['ctas'][i]:5d} {(t[i, 0] - base) / 1e3:8.2f} {(wait[i] - base) / 1e3:8.2f} "
          f"{(t[i, 2] - base) / 1e3:8.2f} {stage[i] / 1e3:8.2f}")

# kernel-specific phase stamps (t[3..]): medians relative to the wait stamp, per kind (records that carry the stamp)
for k in sorted(set(a["kind"])):
    sel = a["kind"] == k
    parts = []
    for c in range(3, 10):
        ok = sel & (t[:, c] > 0)
        if ok.any():
            parts.append(f"t{c}={np.median(t[ok, c] - wait[ok]) / 1e3:.2f}")
    if parts:
        print(f"  {KIND.get(int(k), k)} phases (median us after wait):", ", ".join(parts),
              f"end={np.median(t[sel, 2] - wait[sel]) / 1e3:.2f}")
