"""Dev tool: time llmi_gemm_tokens (token-batched mat-vec = the prefill GEMM) at Gemma-3 shapes.
    [LLMI_NO_UMMA=1] [FAST=1] python tools/prefill_gemm_bench.py [tokens ...]
FAST=1: the dequantize-to-bf16 tcgen05 GEMM (llmi_set_prefill_mode(1)); also prints TFLOP/s and the share of the measured
bf16 peak (MEASURED_PEAKS.json), and the error against the exact path.
Prints per case: ms, TMAC/s (int8 multiply-accumulates of the block dots) and the share of the dense int8 tensor peak
(4.5 PFLOP/s = 2250 TMAC/s nominal)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import ops, synth  # noqa: E402

ops.init_ops(1, 0)
FAST = os.environ.get("FAST") == "1"
_pk = REPO / "MEASURED_PEAKS.json"
BF16_PEAK = json.loads(_pk.read_text()).get("bf16_tflops_sustained", 1405.8) if _pk.exists() else 1405.8
toks = [int(v) for v in sys.argv[1:]] or [64, 256]
cases = [(synth.Q4_0, 1152, 6912), (synth.Q4_0, 6912, 1152), (synth.Q4_0, 5376, 21504), (synth.Q4_0, 21504, 5376),
         (synth.Q8_0, 3840, 15360), (synth.Q4_K, 2560, 10240), (synth.Q6_K, 10240, 2560)]
if os.environ.get("CASE"):  # one case only (ncu captures)
    cases = [cases[int(os.environ["CASE"])]]
for t, k, n in cases:
    w = ops.DeviceWeight(synth.random_blocks(t, n, k, seed=1), t, k, n)
    for m in toks:
        x = ops.DeviceVector(m * k, np.random.default_rng(0).standard_normal(m * k).astype(np.float32))
        o = ops.DeviceVector(m * n)
        err = None
        if FAST:
            ops.set_prefill_mode(False)
            ops.gemm_tokens(w, x, m, o)
            torch.cuda.synchronize()
            exact = o.get().copy()
            ops.set_prefill_mode(True)
            ops.gemm_tokens(w, x, m, o)
            torch.cuda.synchronize()
            err = float(np.abs(o.get() - exact).max() / np.abs(exact).max())
        for _ in range(2):
            ops.gemm_tokens(w, x, m, o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ops.gemm_tokens(w, x, m, o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tmacs = n * k * m / ms / 1e9
        print(json.dumps({"fmt": synth.TYPE_NAMES[t], "K": k, "N": n, "tokens": m, "ms": round(ms, 4),
                          "TMAC_s": round(tmacs, 1), "frac_int8_peak": round(tmacs / 2250, 4),
                          "umma": os.environ.get("LLMI_NO_UMMA") != "1", "fast": FAST,
                          "TFLOP_s": round(2 * tmacs, 1), "frac_bf16_peak": round(2 * tmacs / BF16_PEAK, 4) if FAST else None,
                          "max_err_rel_to_max": err}), flush=True)
        x.close()
        o.close()
    w.close()
