"""Head-sharded KV cache across GPUs (one process per GPU, torch.distributed over NCCL / NVLink).

The cache shards by KV head (dim 1 of [B,H,T,D]): eviction, pooling and attention are independent per
head, so the ONLY cross-GPU coupling on the path is the quantization scale, which the reference takes
over all heads (src/quantization/ops.py:27,:48 on a tensor that spans every head, :330,:342).  The
shards therefore exchange one thing: an allreduce(MAX) of the absmax between phase 1 and phase 2.

The absmax travels as the bit pattern of a non-negative fp32 in an int32 tensor: integer max equals
float max, MAX is exact and order-independent, so scales and codes are bit-identical to the 1-GPU run.
Payload: 4 bytes (per tensor) or 4*T bytes (per token) -- latency-bound, NVLink bandwidth is irrelevant.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch


def head_shard_range(n_heads: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Heads [lo, hi) owned by `rank`.  Heads must divide evenly (Llama-3 KV heads: 8 -> 8/4/2/1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    if n_heads % world_size != 0:
        raise ValueError(f"{n_heads} KV heads do not shard evenly over {world_size} ranks")
    per = n_heads // world_size
    return rank * per, (rank + 1) * per


def shard_heads(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """The [B, H/world, T, D] view of x owned by `rank` (no copy)."""
    lo, hi = head_shard_range(x.size(1), rank, world_size)
    return x[:, lo:hi]


def allreduce_absmax_(bits: torch.Tensor, group=None) -> torch.Tensor:
    """In-place allreduce(MAX) of absmax bit patterns over `group`; no-op when group is None or has one
    rank.  `group` may be a ProcessGroup or True for the default group."""
    if group is None:
        return bits
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("a head-sharded cache needs an initialised torch.distributed process group")
    pg = None if group is True else group
    if dist.get_world_size(pg) == 1:
        return bits
    if bits.dtype != torch.int32:
        raise RuntimeError("absmax bits must be int32 (bit patterns of non-negative fp32)")
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=pg)
    return bits


def world_info() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    import os

    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))
