"""Dev tool: time the decode GEMV kernels at Gemma-3 shapes on one B200.

Each (format, K, N) is timed inside a CUDA graph that launches the GEMV over a
rotation of distinct weight copies totalling >= 2x L2 (so small matrices are
read from HBM, not L2), CUDA events on the capturing stream.  Prints one JSON
line per case: algorithmic GB/s (SURVEY §8d byte count) and fraction of the
measured HBM peak.  Usage: python tools/gemv_sweep.py [--shapes 0x0,4x4,8x1] [--quick]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import ops, synth  # noqa: E402
from llm_inference_b200.synth import F16, Q4_0, Q4_K, Q5_0, Q6_K, Q8_0, BF16  # noqa: E402

L2_BYTES = 126 << 20


def peak_gbs() -> float:
    p = REPO / "MEASURED_PEAKS.json"
    return json.loads(p.read_text())["hbm_gbs"] if p.exists() else 6650.0


def time_case(t, k, n, shape, iters=20, with_quant=False, ring=(1, 0, 0, 0)):
    wbytes = n * synth.row_bytes(t, k)
    copies = max(2, min(64, -(-2 * L2_BYTES // wbytes)))
    raw = synth.random_blocks(t, n, k, seed=1)
    ws = [ops.DeviceWeight(raw, t, k, n) for _ in range(copies)]
    x = ops.DeviceVector(k, np.random.default_rng(0).standard_normal(k).astype(np.float32))
    o = ops.DeviceVector(n)
    act = ops.Activation(k)
    ops.set_gemv_shape(*shape)
    ops.set_gemv_ring(*ring)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        act.prepare(ws[0], x, s.cuda_stream)
        for w in ws:  # warm-up
            ops.gemv(w, act, o, s.cuda_stream)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for w in ws:
                if with_quant:
                    act.prepare(w, x, s.cuda_stream)
                ops.gemv(w, act, o, s.cuda_stream)
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    for h in ws + [x, o, act]:
        h.close()
    ops.set_gemv_shape(0, 0)
    ops.set_gemv_ring(0, 0, 0, 0)
    return us, copies


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="0x0", help="comma list of WxS, 0x0 = heuristic")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--with-quant", action="store_true")
    ap.add_argument("--rings", default="", help="comma list of CPSxDEPTHxWARPS for the persistent ring kernel (e.g. 2x2x16,3x2x8)")
    ap.add_argument("--fmt", default="", help="only this format (e.g. Q4_0)")
    a = ap.parse_args()
    ops.init_ops(1, 0)
    peak = peak_gbs()
    cases = [
        (Q4_0, 1152, 6912), (Q4_0, 6912, 1152), (Q4_0, 1152, 1024), (Q4_0, 1152, 256),
        (Q4_0, 2560, 10240), (Q4_0, 10240, 2560), (Q4_0, 2560, 2048),
        (Q4_0, 5376, 21504), (Q4_0, 21504, 5376), (Q4_0, 5376, 4096), (Q4_0, 4096, 5376), (Q4_0, 5376, 2048),
        (F16, 1152, 262144), (F16, 5376, 262208),
        (Q8_0, 3840, 15360), (Q8_0, 15360, 3840), (Q8_0, 3840, 262208),
        (Q4_K, 2560, 10240), (Q4_K, 10240, 2560), (Q6_K, 10240, 2560), (Q6_K, 2560, 262208),
        (Q5_0, 2560, 10240), (BF16, 2560, 10240),
    ]
    if a.quick:
        cases = [(Q4_0, 1152, 6912), (Q4_0, 2560, 10240), (Q4_0, 10240, 2560), (Q4_0, 5376, 21504), (Q4_0, 21504, 5376),
                 (Q4_0, 5376, 4096), (F16, 1152, 262144), (Q8_0, 3840, 15360), (Q4_K, 2560, 10240), (Q6_K, 10240, 2560)]
    variants = [("slab " + sh, tuple(int(v) for v in sh.split("x")), (1, 0, 0, 0)) for sh in a.shapes.split(",") if sh]
    variants += [("ring " + rg, (0, 0), (2,) + tuple(int(v) for v in rg.split("x"))) for rg in a.rings.split(",") if rg]
    for t, k, n in cases:
        if a.fmt and synth.TYPE_NAMES[t] != a.fmt:
            continue
        for name, shape, ring in variants:
            us, copies = time_case(t, k, n, shape, with_quant=a.with_quant, ring=ring)
            b = synth.algorithmic_bytes(t, n, k)
            gbs = b / us * 1e-3
            print(json.dumps({"fmt": synth.TYPE_NAMES[t], "K": k, "N": n, "shape": name, "us": round(us, 3),
                              "alg_MB": round(b / 1e6, 3), "GBps": round(gbs, 1),
                              "frac_of_measured_peak": round(gbs / peak, 4), "copies": copies}), flush=True)


if __name__ == "__main__":
    main()
