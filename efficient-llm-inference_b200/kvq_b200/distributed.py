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


class AbsmaxExchange:
    """The absmax all-to-all max through peer-mapped memory (csrc/xchg.cu) instead of an NCCL allreduce.

    Every rank of `group` allocates one small symmetric buffer (torch.distributed._symmetric_memory: CUDA VMM handles mapped into
    every process of the node) and exchanges its pointer table once; afterwards `exchange_(bits)` is ONE kernel launch per rank and
    step -- no host-side collective call, no NCCL kernel.  Same result as `allreduce_absmax_` bit for bit.  When symmetric
    memory cannot be set up (single rank, a device without peer mapping) `available` is False and `exchange_` falls back to
    the NCCL allreduce."""

    def __init__(self, device: torch.device, group=None):
        import torch.distributed as dist
        from . import _lib as L

        self.device, self.group, self.seq, self.available, self.why = device, group, 0, False, ""
        self.fused_seq = 0  # step counter of the publish / gather pair (its own region of the buffer)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world == 1:
            self.why = "single rank"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(L.lib().kvq_absmax_exchange_buffer_bytes())
            self._buf = symm_mem.empty(nbytes // 4, dtype=torch.int32, device=device)
            self._buf.zero_()
            self._hdl = symm_mem.rendezvous(self._buf, group if group is not None else dist.group.WORLD)
            self._peers = int(self._hdl.buffer_ptrs_dev)
            self._local = self._buf.data_ptr()
            torch.cuda.synchronize(device)
            dist.barrier(group)  # every buffer zeroed before anyone's first exchange
            self.available = True
        except Exception as e:  # pragma: no cover - depends on the node
            self.why = f"{type(e).__name__}: {e}"

    def exchange_(self, bits: torch.Tensor) -> torch.Tensor:
        """In-place max over ranks of `bits` (int32 bit patterns of non-negative fp32, <= 1024 words), on torch's current stream."""
        if not self.available:
            return allreduce_absmax_(bits, True if self.group is None else self.group) if self.world > 1 else bits
        from . import _lib as L
        if bits.dtype != torch.int32 or not bits.is_contiguous():
            raise RuntimeError("absmax bits must be a contiguous int32 tensor")
        self.seq += 1
        L.check(L.lib().kvq_absmax_exchange(bits.data_ptr(), bits.numel(), self._peers, self.rank, self.world, self.seq,
                                            L.stream_ptr(bits.device)), "kvq_absmax_exchange")
        L.count_launch()
        return bits

    def next_fused_seq(self) -> int:
        """Sequence number for one kvq_absmax_batched_publish / kvq_quant_batched_gather pair (same on every rank)."""
        self.fused_seq += 1
        return self.fused_seq

    @property
    def peer_table_ptr(self) -> int:
        """Device address of the table of all ranks' buffer pointers."""
        return self._peers

    @property
    def local_buffer_ptr(self) -> int:
        return self._local


_SHARED = {}


def shared_exchange(device: torch.device, group=None) -> Optional["AbsmaxExchange"]:
    """One AbsmaxExchange per (device, group) for every layer store of a head-sharded cache: one buffer, one step counter (all
    ranks issue the same calls in the same order).  Created on first use -- a collective: every rank of `group` must get here.
    None when symmetric memory is unavailable (the callers then use the NCCL allreduce)."""
    key = (torch.device(device).index, None if group is None or group is True else id(group))
    ex = _SHARED.get(key)
    if ex is None:
        ex = _SHARED[key] = AbsmaxExchange(torch.device(device), None if group is True else group)
    return ex if ex.available else None
