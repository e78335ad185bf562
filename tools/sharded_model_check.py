"""Row-sharded device-resident model (llmi_model_load_shard) against the single-GPU model: the logits of a prompt
and of every decode step and the greedy tokens must be BIT-IDENTICAL, because a row is computed start to finish
on one device in the canonical order and the exchange only transports it.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/sharded_model_check.py [--same-device] [--model small|1b] [--steps 32]
--same-device puts every rank on cuda:0 (two processes time-slicing one GPU: slow, but it exercises the whole
protocol on a 1-GPU box).  torch.distributed (gloo) only carries the 64-byte IPC handles and the barriers."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--model", default="small")
    ap.add_argument("--weights", default="q4_0")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=8)
    ap.add_argument("--prefill", default="exact", choices=["exact", "fast"],
                    help="fast: the bf16 tensor-core prompt mode (both the sharded and the single-GPU model)")
    a = ap.parse_args()
    import torch.distributed as dist

    from llm_inference_b200 import ops, synth
    from llm_inference_b200.model import Model

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = 0 if a.same_device else int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    if a.model == "small":
        dims = synth.GemmaDims("small", 3, 512, 1024, 4, 2, 128, 520)  # vocab 520: 65 slabs, ragged over the ranks
    elif a.model == "wide":  # 8 KV heads: every rank of a world of 8 owns one (head-sharded attention of a prompt batch)
        dims = synth.GemmaDims("wide", 2, 512, 1024, 16, 8, 64, 520)
    else:
        dims = synth.GEMMA3[a.model if a.model in synth.GEMMA3 else "gemma-3-" + a.model]
    wt = {"q4_0": synth.Q4_0, "q8_0": synth.Q8_0, "q4_k_m": "q4_k_m"}[a.weights]
    et = {"q4_0": synth.F16, "q8_0": synth.Q8_0, "q4_k_m": synth.Q6_K}[a.weights]
    img = synth.build_gemma3_gguf(dims, wt, et, seed=7, embd_std=0.004)
    t_max = a.prompt + a.steps + 8
    ops.init_ops(1, local)
    ops.set_prefill_mode(a.prefill == "fast")  # read when a model is loaded (it sizes the token batches)
    m = Model(img, max_positions=t_max, device=local, world=world, rank=rank)
    m.connect()
    prompt = (np.arange(a.prompt, dtype=np.int32) * 37 + 11) % dims.vocab
    t0 = time.time()
    lg = m.forward(prompt, 0)
    prompt_ms, prompt_launches = m.last_forward_stats()
    first = int(lg.argmax())
    toks, ms = m.decode_greedy(first, a.prompt, a.steps)
    lg_after = m.forward([int(toks[-1])], a.prompt + a.steps)  # full logits through the exchange once more
    wall = time.time() - t0
    err = m.comm_error
    ok, detail = True, {}
    if rank == 0:
        ref = Model(img, max_positions=t_max, device=local)
        rl = ref.forward(prompt, 0)
        ref_prompt_ms = ref.last_forward_stats()[0]
        rtoks, rms = ref.decode_greedy(int(rl.argmax()), a.prompt, a.steps)
        rl_after = ref.forward([int(rtoks[-1])], a.prompt + a.steps)
        ok = (not err and np.array_equal(rl.view(np.uint32), lg.view(np.uint32)) and np.array_equal(rtoks, toks)
              and np.array_equal(rl_after.view(np.uint32), lg_after.view(np.uint32)))
        detail = {"prompt_logits_bitwise": bool(np.array_equal(rl.view(np.uint32), lg.view(np.uint32))),
                  "tokens_equal": bool(np.array_equal(rtoks, toks)),
                  "final_logits_bitwise": bool(np.array_equal(rl_after.view(np.uint32), lg_after.view(np.uint32))),
                  "prefill": a.prefill, "prompt_tokens": a.prompt, "prompt_ms_sharded": prompt_ms,
                  "prompt_launches_sharded": prompt_launches, "prompt_ms_single": ref_prompt_ms,
                  "ms_per_token_sharded": ms / a.steps, "ms_per_token_single": rms / a.steps,
                  "launches_per_step": m.launches_per_step, "tokens_head": [int(t) for t in toks[:8]]}
        ref.close()
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok and not err))
    dist.barrier()
    m.disconnect()
    m.close()
    if rank == 0:
        print(json.dumps({"check": "sharded model == single-GPU model (bitwise)", "world": world, "model": a.model,
                          "weights": a.weights, "same_device": a.same_device, "ok": all(flags), "comm_error": err,
                          "wall_s": round(wall, 2), **detail}), flush=True)
    dist.destroy_process_group()
    return 0 if all(flags) else 1


if __name__ == "__main__":
    sys.exit(main())
