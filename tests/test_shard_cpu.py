"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups.  Each rank
computes its row shard with the port oracle standing in for the device kernel
(rows are independent, ops.cpp:439-448), the slices are all-gathered in place,
and the result must equal the unsharded mat-vec bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_inference_b200 import shard, synth


def test_row_ranges_are_slab_aligned_and_cover():
    for n, w in [(6912, 2), (21504, 8), (1152, 8), (1030, 4), (5, 2), (262208, 8), (8, 4)]:
        rs = shard.row_ranges(n, w)
        assert len(rs) == w and rs[0][0] == 0 and rs[-1][1] == n
        for (b0, e0), (b1, e1) in zip(rs, rs[1:]):
            assert e0 == b1 and b0 <= e0
        assert all(b % 8 == 0 for b, e in rs if e > b)
        sizes = [e - b for b, e in rs]
        assert max(sizes) - min(sizes) < 16 or n < 8 * w  # one slab of imbalance + a ragged last slab
    assert shard.equal_ranges(21504, 8) == [(i * 2688, (i + 1) * 2688) for i in range(8)]
    assert shard.equal_ranges(1030, 4) is None


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, cases, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Port
    P = Port()
    ok = True
    for t, k, n in cases:
        w = synth.random_blocks(t, n, k, seed=t + k + n)          # every rank holds the same file
        x = np.random.default_rng(k).standard_normal(k).astype(np.float32)
        ranges = shard.row_ranges(n, world)
        rb = synth.row_bytes(t, k)
        b, e = ranges[rank]
        full = torch.full((n,), float("nan"))
        if e > b:
            mine = P.mat_vec_mul(t, shard.shard_blocks(w, rb, (b, e)), x, e - b, k)
            full[b:e] = torch.from_numpy(mine)
        shard.allgather_rows(full, ranges, rank)
        ref = P.mat_vec_mul(t, w, x, n, k)
        ok &= bool(np.array_equal(full.numpy().view(np.uint32), ref.view(np.uint32)))
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        out_q.put(int(res.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matvec_allgather_equals_unsharded(world):
    cases = [(synth.Q4_0, 1152, 6912), (synth.Q4_0, 1152, 1030), (synth.Q8_0, 256, 40), (synth.Q4_K, 512, 136),
             (synth.F16, 128, 520), (synth.Q6_K, 256, 5)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


def _batch_worker(rank: int, world: int, port: int, cases, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Port
    P = Port()
    ok = True
    for t, k, n, n_tok in cases:
        w = synth.random_blocks(t, n, k, seed=t + k + n)
        xs = np.random.default_rng(k + n_tok).standard_normal((n_tok, k)).astype(np.float32)
        ranges = shard.row_ranges(n, world)
        rb = synth.row_bytes(t, k)
        b, e = ranges[rank]
        batch = torch.full((n_tok, n), float("nan"))
        if e > b:  # this rank's rows of EVERY token: the column block [b, e) of the [token][row] batch
            mine = shard.shard_blocks(w, rb, (b, e))
            for m in range(n_tok):
                batch[m, b:e] = torch.from_numpy(P.mat_vec_mul(t, mine, xs[m], e - b, k))
        shard.allgather_columns(batch, ranges, rank)
        ref = np.stack([P.mat_vec_mul(t, w, xs[m], n, k) for m in range(n_tok)])
        ok &= bool(np.array_equal(batch.numpy().view(np.uint32), ref.view(np.uint32)))
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        out_q.put(int(res.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_token_batch_allgather_equals_unsharded(world):
    """The prompt path of a sharded model on CPU: every rank computes its rows of every token, the column
    blocks are all-gathered (gloo), and the [token][row] batch equals the unsharded one bit for bit."""
    cases = [(synth.Q4_0, 256, 136, 5), (synth.Q8_0, 128, 40, 3), (synth.Q4_K, 512, 72, 4), (synth.F16, 64, 20, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


def test_token_slices_and_head_ranges():
    """The partitions run_batch uses for the norm stages (token slices) and the attention (KV heads per rank)."""
    for n, w in [(2048, 8), (2048, 3), (100, 8), (37, 3), (8, 8), (5, 8)]:
        sl = shard.token_slices(n, w)
        assert len(sl) == w and sl[0][0] == 0 and max(e for _, e in sl) == n
        assert all(b0 <= e0 and e0 == b1 for (b0, e0), (b1, _) in zip(sl, sl[1:]) if e0 < n)
        assert sum(e - b for b, e in sl) == n
    assert shard.token_slices(2048, 8)[3] == (768, 1024)
    assert shard.token_slices(100, 8)[7] == (91, 100)
    assert shard.kv_head_ranges(16, 8) == [(2 * r, 2 * r + 2) for r in range(8)]
    assert shard.kv_head_ranges(16, 3) is None and shard.kv_head_ranges(1, 2) is None
    # the q rows a rank computes are exactly its heads' when the KV heads divide (gemma-3-27b: H 32, HK 16, D 128)
    for w in (2, 4, 8):
        rows = shard.row_ranges(32 * 128, w)
        heads = shard.kv_head_ranges(16, w)
        assert [(b * 2 * 128, e * 2 * 128) for b, e in heads] == rows


def test_c_abi_shard_range_is_the_python_partition():
    """llmi_shard_range (what llmi_model_load_shard applies to every matrix) == shard.row_ranges: contiguous,
    slab-aligned, covering, as even as possible — for ragged row counts and more ranks than slabs."""
    import ctypes as C

    from llm_inference_b200 import _lib
    L = _lib.load()
    for n in (0, 1, 7, 8, 9, 40, 203, 256, 520, 1152, 6912, 262144, 262208):
        for world in (1, 2, 3, 4, 5, 8):
            want = shard.row_ranges(n, world) if n else [(0, 0)] * world
            for rank in range(world):
                b, e = C.c_uint64(), C.c_uint64()
                assert L.llmi_shard_range(n, world, rank, C.byref(b), C.byref(e)) == 0
                assert (b.value, e.value) == want[rank], (n, world, rank)
    b, e = C.c_uint64(), C.c_uint64()
    assert L.llmi_shard_range(8, 2, 2, C.byref(b), C.byref(e)) != 0  # rank out of range
