"""Multi-GPU parity check (run under torchrun, one rank per GPU):
row-sharded mat-vecs + NCCL all-gather must equal the unsharded single-GPU
result bit for bit, for every format.  Prints one line per rank-0 case."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from llm_inference_b200 import ops, shard, synth  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ops.init_ops(1, device=local)
stream = torch.cuda.current_stream()
ok_all = True
for t, k, n in [(synth.Q4_0, 5376, 4096), (synth.Q4_0, 1152, 1030), (synth.Q8_0, 3840, 2048), (synth.Q4_K, 2560, 1024),
                (synth.Q6_K, 2560, 136), (synth.F16, 1152, 4096), (synth.Q5_0, 256, 72), (synth.BF16, 256, 72)]:
    w = synth.random_blocks(t, n, k, seed=t + k)
    x = np.random.default_rng(k).standard_normal(k).astype(np.float32)
    ranges = shard.row_ranges(n, world)
    b, e = ranges[rank]
    rb = synth.row_bytes(t, k)
    dw = ops.DeviceWeight(shard.shard_blocks(w, rb, (b, e)), t, k, n, b, e, blocks_are_shard=True)
    full_w = ops.DeviceWeight(w, t, k, n)
    dx = ops.DeviceVector(k, x)
    act = ops.Activation(k)
    out_t = torch.full((n,), float("nan"), device="cuda")
    ref_t = torch.zeros(n, device="cuda")
    sp = stream.cuda_stream
    ops.mat_vec_mul_dev(dw, dx, act, ops.TorchVector(out_t), sp)
    shard.allgather_rows(out_t, ranges, rank)
    ops.mat_vec_mul_dev(full_w, dx, act, ops.TorchVector(ref_t), sp)
    torch.cuda.synchronize()
    ok = bool(torch.equal(out_t.view(torch.int32), ref_t.view(torch.int32)))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_all &= bool(flag.item())
    if rank == 0:
        print(f"{synth.TYPE_NAMES[t]} {k}->{n} x{world}: sharded+all-gather bit-identical = {bool(flag.item())}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
