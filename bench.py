#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 quantized mat-vec path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload gemma-3-1b-q4_0|gemma-3-4b-q4_k_m|gemma-3-12b-q8_0|gemma-3-27b-q4_0]

Metric (BASELINE.json): decode tok/s + Q4_0 GEMV HBM GB/s (% of B200 peak) at
Gemma-3 shapes.  A *step* is one decode token's pass over the hot path named by
north_star: every mat_vec_mul of Model::forward for one token (7 per layer —
q, k, v, attn_output, gate, up, down: model.cpp:754,784,803,557,875,877,909 —
plus the logits mat-vec, model.cpp:1000/1027), each with the activation
quantization the reference performs inside the call (ops.cpp:209-210 etc.).
The default workload is BASELINE.json configs[1] (gemma-3-1b Q4_0, F16 tied
logits); weights are synthetic random blocks of exactly those shapes/formats.

  value     tok/s with weights AND activations resident in HBM, the whole step
            replayed as one CUDA graph, CUDA events on the launching stream.
            Consecutive steps read 1 GB of distinct weights (> 126 MB L2), so no
            L2 flush is needed between timed iterations ("inputs larger than L2").
  e2e       tok/s through the reference-facing host-vector API (ops.h drop-in
            semantics: host x in, host o out, synchronous, per call) — H2D/D2H
            copies inside the timed region.
  roofline  for the kernel class with the largest share of the step, timed live
            with CUDA events on its own launches: algorithmic bytes (SURVEY §8d:
            reference-format weight bytes + 4K + 4N per call) / time, against
            MEASURED_PEAKS.json hbm_gbs ("of measured"; fallback 6650 "of fallback").
  cpu_baseline  the reference's own CPU code (oracle/_ref, kind "reference") or
            the port oracle, all host threads, bounded sample of the same step.

N > 1: one process per GPU (torchrun); every rank decodes its own replica of the
workload (independent sequences, no data-path collective): scaling "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from llm_inference_b200 import synth  # noqa: E402
from llm_inference_b200.synth import F16, Q4_0, Q4_K, Q6_K, Q8_0  # noqa: E402

WORKLOADS = {
    # name: (dims, layer weight type or "q4_k_m", logits/embedding type)
    "gemma-3-1b-q4_0": ("gemma-3-1b", Q4_0, F16),
    "gemma-3-4b-q4_k_m": ("gemma-3-4b", "q4_k_m", Q6_K),
    "gemma-3-12b-q8_0": ("gemma-3-12b", Q8_0, Q8_0),
    "gemma-3-27b-q4_0": ("gemma-3-27b", Q4_0, F16),
}


def step_calls(workload: str, n_layer: int | None = None):
    """The mat-vec calls of one decode token: list of (name, type, K, N, x_key).
    x_key groups calls that consume the same activation vector (the reference
    re-quantizes per call; so do we in the timed step, to stay call-for-call)."""
    dims_name, wt, et = WORKLOADS[workload]
    d = synth.GEMMA3[dims_name]
    L = d.n_layer if n_layer is None else n_layer
    E, F, H, HK, D = d.n_embd, d.n_ff, d.n_head, d.n_head_kv, d.head_dim
    per_layer = synth.q4_k_m_layer_types(L) if wt == "q4_k_m" else None
    calls = []
    for i in range(L):
        ty = (lambda nm: per_layer[i][nm]) if per_layer else (lambda nm: wt)
        calls += [(f"blk.{i}.attn_q", ty("attn_q"), E, H * D, "xE"), (f"blk.{i}.attn_k", ty("attn_k"), E, HK * D, "xE"),
                  (f"blk.{i}.attn_v", ty("attn_v"), E, HK * D, "xE"),
                  (f"blk.{i}.attn_output", ty("attn_output"), H * D, E, "xHD"),
                  (f"blk.{i}.ffn_gate", ty("ffn_gate"), E, F, "xE"), (f"blk.{i}.ffn_up", ty("ffn_up"), E, F, "xE"),
                  (f"blk.{i}.ffn_down", ty("ffn_down"), F, E, "xF")]
    calls.append(("logits", et, E, d.vocab, "xE"))
    return calls


def algorithmic_bytes(calls) -> int:
    return sum(synth.algorithmic_bytes(t, n, k) for _, t, k, n, _ in calls)


def peak_gbs():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def gen_weights(calls, seed=1):
    """Raw reference-layout blocks per call.  Random blocks (sane f16 scales);
    identical shapes share one generated buffer per (type, K, N) and layer
    parity, but every call still gets its own device copy."""
    cache, out = {}, []
    for idx, (name, t, k, n, _) in enumerate(calls):
        key = (t, k, n, idx % 2)
        if key not in cache:
            cache[key] = synth.random_blocks(t, n, k, seed=seed + len(cache))
        out.append(cache[key])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ reference arm

def cpu_step_rate(calls, weights, budget_s: float, threads: int):
    """tok/s of the reference's CPU implementation on a bounded sample of the step."""
    from oracle import binding
    kind = "reference" if binding.ref_available() else "port"
    rng = np.random.default_rng(0)
    xs = {}
    for _, _, k, _, key in calls:
        xs.setdefault((key, k), rng.standard_normal(k).astype(np.float32))
    if kind == "reference":
        R = binding.Ref(n_threads=threads)
    else:
        R = binding.Port()
        threads = 1
    # bounded sample: the first layers + the logits call, scaled to a whole step by call bytes
    total_b = algorithmic_bytes(calls)
    sample, sample_b = [], 0
    n_layers_sample = 0
    for i, c in enumerate(calls):
        if c[0] == "logits" or i < 7 * 2:
            sample.append(i)
            sample_b += synth.algorithmic_bytes(c[1], c[3], c[2])
    n_layers_sample = min(2, (len(calls) - 1) // 7)
    handles = []
    for i in sample:
        _, t, k, n, _ = calls[i]
        if kind == "reference" and t != F16:
            handles.append(R.tensor(t, weights[i], n, k))
        else:
            handles.append(None)
    f16_handles = {}
    import ctypes as C
    fp = C.POINTER(C.c_float)
    for i in sample:  # one-off host copies of F16 matrices (the reference copies token_embd once too, model.cpp:46-55)
        _, t, k, n, _ = calls[i]
        if kind == "reference" and t == F16:
            w16 = np.ascontiguousarray(weights[i]).view(np.uint16)
            f16_handles[i] = R.L.ref_f16_create(w16.ctypes.data_as(C.POINTER(C.c_uint16)), n, k)

    def run(i, h):
        _, t, k, n, key = calls[i]
        x = xs[(key, k)]
        if h is not None:
            h.loop(x, 1)
        elif kind == "reference":
            R.L.ref_f16_mat_vec_mul_loop(f16_handles[i], x.ctypes.data_as(fp), None, 1)
        else:
            R.mat_vec_mul(t, weights[i], x, n, k)

    t_layers = t_logits = 0.0
    reps = 0
    for i, h in zip(sample, handles):  # warm call
        run(i, h)
    while t_layers + t_logits < budget_s and reps < 1000:
        for i, h in zip(sample, handles):
            t0 = time.perf_counter()
            run(i, h)
            dt = time.perf_counter() - t0
            if calls[i][0] == "logits":
                t_logits += dt
            else:
                t_layers += dt
        reps += 1
    for h in handles:
        if h is not None:
            h.close()
    for h in f16_handles.values():
        R.L.ref_f16_free(h)
    n_layers = (len(calls) - 1) // 7
    # layers are identical in shape: the layer part scales linearly, the logits call is measured whole
    est_step = (t_layers / reps) * n_layers / max(1, n_layers_sample) + t_logits / reps
    return {"value": 1.0 / est_step, "unit": "tok/s", "cores": threads, "kind": kind,
            "sample": f"{n_layers_sample} of {n_layers} layers (7 mat-vecs each, scaled linearly) + the whole logits "
                      f"mat-vec, {reps} reps, {t_layers + t_logits:.1f} s of CPU time, init_ops({threads})",
            "GBps": total_b / est_step / 1e9, "ms_per_step": est_step * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    calls = step_calls(args.workload)
    weights = gen_weights(calls)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    for _ in range(max(0, args.warmup - 1)):
        cpu_step_rate(calls, weights, 0.5, threads)
    vals = [cpu_step_rate(calls, weights, max(1.0, 20.0 / max(1, args.steps)), threads) for _ in range(args.steps)]
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1])
    cb["value"] = v
    line = {"impl": "reference", "metric": "decode tok/s (hot path: all mat-vecs of one token)", "value": v,
            "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8xint4->int32 block dots, fp32 accumulate (AVX2 on host cores)", "data": "synthetic",
            "config": {"workload": args.workload, "calls_per_step": len(calls),
                       "algorithmic_MB_per_step": algorithmic_bytes(calls) / 1e6},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm

def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from llm_inference_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ops.init_ops(1, device=local)  # fails loudly if the CUDA extension / a B200 is missing

    calls = step_calls(args.workload)
    weights = gen_weights(calls)
    rng = np.random.default_rng(rank)
    dws = [ops.DeviceWeight(w, t, k, n) for w, (_, t, k, n, _) in zip(weights, calls)]
    xs, outs, acts = {}, {}, {}
    for _, _, k, n, key in calls:
        if (key, k) not in xs:
            xs[(key, k)] = ops.DeviceVector(k, rng.standard_normal(k).astype(np.float32))
            acts[k] = ops.Activation(k)
        outs.setdefault(n, ops.DeviceVector(n))

    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    def enqueue(subset=None):
        n_launch = 0
        for i, (_, t, k, n, key) in enumerate(calls):
            if subset is not None and not subset(i):
                continue
            ops.mat_vec_mul_dev(dws[i], xs[(key, k)], acts[k], outs[n], sp)  # quantize + GEMV, like one reference call
            n_launch += 2
        return n_launch

    def make_graph(subset=None):
        with torch.cuda.stream(stream):
            enqueue(subset)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                n = enqueue(subset)
        return g, n

    def time_graph(g, iters, warm):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                g.replay()
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                g.replay()
            e1.record(stream)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters  # ms

    g_step, launches_per_step = make_graph()
    with ClockSampler(local) as clocks:
        ms = time_graph(g_step, args.steps, max(3, args.warmup))
        # keep the sampler alive long enough to see the clocks under load
        if ms * args.steps < 400:
            time_graph(g_step, int(400 / ms) + 1, 0)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * 1e3 / ms

    # per kernel class (live, CUDA events on the launches of that class only)
    classes = {}
    for i, (_, t, k, n, _) in enumerate(calls):
        classes.setdefault(synth.TYPE_NAMES[t], []).append(i)
    kern = []
    for name, idxs in classes.items():
        sel = set(idxs)
        g, nl = make_graph(lambda i, sel=sel: i in sel)
        # classes smaller than L2 are rotated with the full step between timings by replaying g_step first
        reps = max(5, args.steps // 4)
        tot = 0.0
        for _ in range(reps):
            g_step.replay()
            tot += time_graph(g, 1, 0)
        cms = tot / reps
        # quantize launches are part of the class' calls; split them off by timing quantize-free GEMVs
        b = sum(synth.algorithmic_bytes(calls[i][1], calls[i][3], calls[i][2]) for i in idxs)
        kern.append({"kernel": f"gemv_slab_kernel<{name}> (+ its activation prep)", "calls": len(idxs),
                     "ms_per_step": cms, "share_of_step": cms / ms, "algorithmic_MB": b / 1e6,
                     "GBps": b / cms / 1e6, "us_per_call": cms * 1e3 / len(idxs)})
    kern.sort(key=lambda r: -r["ms_per_step"])
    peak, peak_src = peak_gbs()
    dom = kern[0]
    roof = {"bound": "hbm", "achieved": dom["GBps"], "peak": peak, "unit": "GB/s", "frac": dom["GBps"] / peak,
            "traffic": None, "kernel": dom["kernel"], "peak_source": f"of {peak_src}",
            "frac_of_nominal_8TBps": dom["GBps"] / 8000.0,
            "note": "achieved = algorithmic bytes of this class' calls / CUDA-event time of exactly those launches "
                    "(includes the activation-prep kernel of each call and inter-kernel gaps)"}

    # end-to-end through the reference-facing host-vector API
    e2e = None
    if rank == 0 or world > 1:
        hx = {(key, k): xs[(key, k)].get() for (key, k) in xs}
        ho = {n: np.empty(n, np.float32) for n in outs}
        from llm_inference_b200 import _lib
        L = _lib.load()

        def host_step():
            for i, (_, t, k, n, key) in enumerate(calls):
                x = hx[(key, k)]
                _lib.check(L.llmi_host_mat_vec_mul(dws[i].h, x.ctypes.data, k, ho[n].ctypes.data, n))

        for _ in range(2):
            host_step()
        n_e2e = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host_step()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world / dt, "unit": "tok/s",
               "h2d_bytes_per_step": int(sum(4 * c[2] for c in calls)),
               "d2h_bytes_per_step": int(sum(4 * c[3] for c in calls)),
               "api": "llmi_host_mat_vec_mul per call (ops.h mat_vec_mul drop-in semantics: host x in, host o out, "
                      "synchronous)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_step_rate(calls, weights, args.cpu_seconds, os.cpu_count() or 1)

    if rank == 0:
        line = {"metric": "decode tok/s (hot path: all mat-vecs of one token)", "value": value, "unit": "tok/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8xint4->int32 block dots (dp4a), fp16 scales, fp32 accumulate", "data": "synthetic",
                "config": {"workload": args.workload, "calls_per_step": len(calls),
                           "algorithmic_MB_per_step": algorithmic_bytes(calls) / 1e6,
                           "parallelism": "single GPU" if world == 1 else f"{world} independent replicas",
                           "l2": "each step streams %.0f MB of distinct weights (> 126 MB L2): inputs larger than L2"
                                 % (algorithmic_bytes(calls) / 1e6),
                           "timing": "one CUDA graph per step, CUDA events on the launching stream, max over ranks"},
                "step_GBps": algorithmic_bytes(calls) / ms / 1e6, "step_frac_of_peak": algorithmic_bytes(calls) / ms / 1e6 / peak,
                "roofline": roof, "kernels": kern, "e2e": e2e, "cpu_baseline": cpu,
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gemma-3-1b-q4_0", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
