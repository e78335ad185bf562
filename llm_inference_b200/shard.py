"""Output-row sharding of the mat-vec path across the GPUs of one node
 — the multi-GPU image of the reference's thread partition
(ops.cpp:439-448: rows split into contiguous chunks, one worker each).

Every rank holds rows ``[begin, end)`` of every weight matrix (slab-aligned, so
no 8-row slab straddles two ranks) and the full, replicated activation vector;
a row is computed start-to-finish on one device in the canonical summation
order, so the gathered result is bit-identical to the single-GPU result.  The
one exchange step per mat-vec is an all-gather of the fp32 output slices, done
in place on the full-length output vector.  The collective goes through
``torch.distributed`` (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

SLAB = 8  # rows per slab of the device layout (llmi_internal.h LLMI_SLAB)


def row_ranges(n_rows: int, world: int, align: int = SLAB) -> list[tuple[int, int]]:
    """Contiguous, ``align``-aligned row ranges, as even as possible; the last
    ranks may be empty when the matrix has fewer slabs than ranks."""
    n_units = -(-n_rows // align)
    out, start = [], 0
    for r in range(world):
        cnt = n_units // world + (1 if r < n_units % world else 0)
        b, e = min(n_rows, start * align), min(n_rows, (start + cnt) * align)
        out.append((b, e))
        start += cnt
    assert out[0][0] == 0 and out[-1][1] == n_rows
    return out


def equal_ranges(n_rows: int, world: int, align: int = SLAB) -> list[tuple[int, int]] | None:
    """Equal-sized aligned ranges (what an in-place all-gather needs), or None."""
    if n_rows % (world * align):
        return None
    per = n_rows // world
    return [(r * per, (r + 1) * per) for r in range(world)]


def token_slices(n_tok: int, world: int) -> list[tuple[int, int]]:
    """Contiguous token slices of a prompt batch, ceil(n_tok / world) each (the last ranks' may be short or empty):
    where the norm stages of a sharded batch run (model.cu run_batch)."""
    per = -(-n_tok // world)
    return [(min(n_tok, r * per), min(n_tok, (r + 1) * per)) for r in range(world)]


def kv_head_ranges(n_head_kv: int, world: int) -> list[tuple[int, int]] | None:
    """KV heads per rank for the attention of a sharded prompt batch, or None when the heads do not divide (then every
    rank attends with every head)."""
    if n_head_kv % world:
        return None
    per = n_head_kv // world
    return [(r * per, (r + 1) * per) for r in range(world)]


def allgather_columns(batch, ranges, rank: int, group=None):
    """In-place all-gather of a [tokens][rows] batch whose columns ``ranges[rank]`` this rank has filled: the host-side
    statement of what bx_exchange_kernel / the GEMM epilogue's peer stores do on the device (every rank ends up with every
    rank's column block; no arithmetic on the way)."""
    import torch.distributed as dist

    for src, (b, e) in enumerate(ranges):
        if e > b:
            blk = batch[:, b:e].contiguous()
            dist.broadcast(blk, src=src, group=group)
            if src != rank:
                batch[:, b:e] = blk
    return batch


def allgather_rows(full, ranges, rank: int, group=None):
    """In-place all-gather of a full-length vector whose ``ranges[rank]`` slice
    this rank has filled.  ``full`` is a torch tensor (cuda with NCCL, cpu with
    gloo).  Equal slices use all_gather_into_tensor; ragged ones a list gather."""
    import torch
    import torch.distributed as dist

    sizes = {e - b for b, e in ranges}
    b, e = ranges[rank]
    if len(sizes) == 1 and dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(full, full[b:e], group=group)
        return full
    parts = [torch.empty(re - rb, dtype=full.dtype, device=full.device) for rb, re in ranges]
    dist.all_gather(parts, full[b:e].contiguous(), group=group) if len(sizes) == 1 else \
        _ragged_gather(parts, full[b:e].contiguous(), ranges, rank, group)
    for (rb, re), p in zip(ranges, parts):
        full[rb:re] = p
    return full


def _ragged_gather(parts, mine, ranges, rank, group):
    import torch.distributed as dist

    for src, (rb, re) in enumerate(ranges):
        if re > rb:
            buf = mine if src == rank else parts[src]
            dist.broadcast(buf, src=src, group=group)
            if src == rank:
                parts[src].copy_(mine)


def shard_blocks(blocks: np.ndarray, row_bytes: int, rng: tuple[int, int]) -> np.ndarray:
    """The raw reference-layout bytes of rows [begin, end) (host side view)."""
    b, e = rng
    return np.ascontiguousarray(blocks).view(np.uint8).ravel()[b * row_bytes:e * row_bytes]
