"""N > 1 host path on CPU: two gloo ranks each own half of the KV heads, exchange only the absmax
(allreduce MAX of fp32 bit patterns) and must reproduce the unsharded codes and scales bit for bit.
The per-shard arithmetic is done by the oracle here (there is no GPU); the plumbing under test is
kvq_b200.distributed (shard ranges, the bit-pattern allreduce)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kvq_b200.distributed import allreduce_absmax_, shard_heads
        from oracle import kvq_oracle as o

        g = torch.Generator().manual_seed(1234)
        x = torch.randn(2, 8, 16, 32, generator=g).to(torch.bfloat16)
        x[0, 5, 3, 7] = 9.5  # the global max lives in rank 1's shard (heads 4..7) when world == 2
        mine = shard_heads(x, rank, world).contiguous()

        # per tensor: local absmax -> bits -> allreduce(MAX) -> quantize shard with the global max
        bits = torch.tensor([o.absmax(mine)], dtype=torch.float32).view(torch.int32)
        allreduce_absmax_(bits, True)
        gmax = float(bits.view(torch.float32)[0])
        q8, s8 = o.quantize_int8_per_tensor(mine, absmax_override=gmax)
        p4, s4, _ = o.quantize_int4_per_tensor_packed(mine, absmax_override=gmax)

        # per token: T absmax words
        tb = torch.stack([torch.tensor(o.absmax(mine[:, :, t].contiguous())) for t in range(x.size(2))]).view(torch.int32)
        allreduce_absmax_(tb, True)

        full_q8, full_s8 = o.quantize_int8_per_tensor(x)
        full_p4, full_s4, _ = o.quantize_int4_per_tensor_packed(x)
        _, _, full_tok = o.quantize_per_token(x, 8)
        want_tb = torch.stack([torch.tensor(o.absmax(x[:, :, t].contiguous())) for t in range(x.size(2))]).view(torch.int32)
        ok = (torch.equal(q8, shard_heads(full_q8, rank, world)) and torch.equal(p4, shard_heads(full_p4, rank, world))
              and s8.view(torch.int16) == full_s8.view(torch.int16) and s4.view(torch.int16) == full_s4.view(torch.int16)
              and torch.equal(tb, want_tb) and gmax == 9.5 and full_tok.numel() == x.size(2))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_head_sharded_scale_exchange_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_allreduce_is_noop_without_group():
    from kvq_b200.distributed import allreduce_absmax_
    b = torch.tensor([5], dtype=torch.int32)
    assert allreduce_absmax_(b, None) is b


def _worker_halves(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # bench.py's N > 1 step: the absmax words of the batch are reduced in two halves with async allreduce(MAX) on views of one
        # buffer, each waited for just before its quantize launch -- must equal one allreduce over the whole buffer
        n, half = 64, 32
        g = torch.Generator().manual_seed(100 + rank)
        bits = torch.rand(n, generator=g).view(torch.int32).clone()      # non-negative fp32 bit patterns: int order == float order
        whole = bits.clone()
        works = [dist.all_reduce(bits[lo: lo + cnt], op=dist.ReduceOp.MAX, async_op=True) for lo, cnt in ((0, half), (half, n - half))]
        for w in works:
            w.wait()
        dist.all_reduce(whole, op=dist.ReduceOp.MAX)
        gathered = [torch.empty_like(whole) for _ in range(world)]
        dist.all_gather(gathered, whole)
        ok = torch.equal(bits, whole) and all(torch.equal(t, whole) for t in gathered)
        ok = ok and torch.equal(whole.view(torch.float32), torch.stack([torch.rand(n, generator=torch.Generator().manual_seed(100 + r))
                                                                        for r in range(world)]).max(0).values)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_half_async_allreduce_equals_one_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_halves, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def _worker_exchange_fallback(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kvq_b200.distributed import AbsmaxExchange
        # no CUDA device here: symmetric memory cannot be set up, so the exchange must say why and fall back to the allreduce
        ex = AbsmaxExchange(torch.device("cpu"), None)
        bits = torch.tensor([1.0 + rank, 7.0 - 3 * rank, 0.0], dtype=torch.float32).view(torch.int32).clone()
        out = ex.exchange_(bits)
        want = torch.tensor([float(world), 7.0, 0.0], dtype=torch.float32).view(torch.int32)
        q.put((rank, bool(not ex.available and ex.why and out is bits and torch.equal(bits, want))))
    finally:
        dist.destroy_process_group()


def test_absmax_exchange_falls_back_to_allreduce_without_symmetric_memory():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_exchange_fallback, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_absmax_exchange_single_rank_is_identity():
    from kvq_b200.distributed import AbsmaxExchange
    ex = AbsmaxExchange(torch.device("cpu"), None)
    b = torch.tensor([3], dtype=torch.int32)
    assert not ex.available and ex.why == "single rank" and ex.exchange_(b) is b


def test_shared_exchange_is_none_without_peers():
    """One process, no process group: the stores' exchange is unavailable ("single rank") and cached; callers take the other path."""
    from kvq_b200.distributed import shared_exchange
    assert shared_exchange(torch.device("cpu"), None) is None
    assert shared_exchange(torch.device("cpu"), None) is None
