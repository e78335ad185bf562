"""Row-sharded device-resident model (llmi_model_load_shard) on ONE GPU: the ranks are separate
processes time-slicing cuda:0, wired through CUDA IPC exactly as on an NVLink box, so the whole exchange protocol
(flagged 64-bit stores from the mat-vec epilogues into every rank's buffer, consumers spinning on the tag, argmax
keys swapped between ranks, ragged row ranges) runs here.  Bar: prompt logits, every greedy token and the final
logits BIT-IDENTICAL to the single-GPU model (tools/sharded_model_check.py does the comparison on rank 0)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world,weights,port", [(2, "q4_0", 29541), (3, "q4_k_m", 29542), (2, "q8_0", 29543)])
def test_sharded_model_is_bitwise_the_single_gpu_model(gpu_ops, world, weights, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "sharded_model_check.py"),
           "--same-device", "--steps", "6", "--prompt", "4", "--weights", weights]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=300)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and not res["comm_error"], res
    assert res["prompt_logits_bitwise"] and res["tokens_equal"] and res["final_logits_bitwise"], res


@pytest.mark.parametrize("world,model,weights,prompt,prefill,batch,port", [
    (2, "small", "q4_0", 40, "exact", None, 29544),   # token-lane kernel, one batch; each rank attends with its own heads
    (3, "small", "q8_0", 37, "exact", "16", 29545),  # three batches, the last one ragged; ragged row ranges; attention replicated
    (2, "small", "q4_0", 100, "fast", None, 29546),   # bf16 tensor-core mode: GEMM epilogues store into the peer, q stays local
    (3, "small", "q8_0", 150, "fast", "64", 29547),
    (8, "wide", "q4_0", 100, "fast", None, 29548),    # a world of 8, one KV head per rank
    (8, "wide", "q4_0", 70, "exact", "32", 29549),
])
def test_sharded_prompt_batches_are_bitwise_the_single_gpu_batches(gpu_ops, world, model, weights, prompt, prefill, batch,
                                                                   port):
    """Prompts of a sharded model go through the token-batched kernels (run_batch): every rank computes its rows of
    every token and bx_exchange_kernel all-gathers the [token][row] tiles over peer memory.  Logits after the prompt,
    the greedy tokens that continue from its KV cache and the final logits must equal the single-GPU model's in the
    same mode bit for bit."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "sharded_model_check.py"),
           "--same-device", "--steps", "4", "--prompt", str(prompt), "--weights", weights, "--prefill", prefill,
           "--model", model]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.pop("LLMI_PREFILL", None)
    env["LLMI_SEQ_NORM_MIN_TOKENS"] = "1"  # norm stages on a token slice per rank even for these short prompts
    if batch:
        env["LLMI_PREFILL_BATCH"] = batch
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=300)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and not res["comm_error"], res
    assert res["prompt_logits_bitwise"] and res["tokens_equal"] and res["final_logits_bitwise"], res
    # batched, not token by token: a 3-layer model costs ~20 launches per token in run_step
    assert res["prompt_launches_sharded"] < 12 * prompt, res
