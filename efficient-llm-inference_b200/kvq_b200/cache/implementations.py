"""KV-cache eviction, paging and chunk-summary strategies on B200: the reference's
`src/cache/implementations.py` API over hand-written sm_100a kernels (csrc/evict.cu, csrc/pool.cu).

Reference (/root/reference/src/cache/implementations.py):
    PagedKVCache :10-121, trim_kv_sliding_window :124-140, trim_kv_prefix_window :143-154,
    trim_kv_strided :157-190, trim_kv_block_old :193-245, trim_kv_budget_old :248-292,
    chunk_summarize_kv :295-346.

Every trim_* computes its kept-token set as a closed form (kvq_evict_plan_make, host arithmetic) and
one gather kernel per tensor evaluates source indices in registers: no arange / linspace / index
tensor / parts list, the cache is read once for the kept rows only.  As in the reference, a cache that
is already short enough is returned unchanged (same tensor objects) and the sliding window returns
views.
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch

from .. import _lib as L


# ----------------------------------------------------------------------------- eviction plans


def evict_plan(strategy: int, T: int, window: int = 0, prefix: int = 0, stride: int = 1, block: int = 1, keep: int = 1,
               budget: int = 0) -> L.EvictPlan:
    """Host arithmetic only (no CUDA): the closed-form description of the kept set for length T."""
    plan = L.EvictPlan()
    L.check(L.lib().kvq_evict_plan_make(strategy, int(T), int(window), int(prefix), int(stride), int(block), int(keep),
                                        int(budget), ctypes.byref(plan)), "kvq_evict_plan_make")
    return plan


def kept_token_indices(plan: L.EvictPlan) -> List[int]:
    """The kept source tokens, in output order, computed on the host (for inspection and tests)."""
    buf = (ctypes.c_int64 * max(int(plan.n_total), 1))()
    L.check(L.lib().kvq_evict_indices_host(ctypes.byref(plan), ctypes.cast(buf, ctypes.c_void_p)), "kvq_evict_indices_host")
    return list(buf[: plan.n_total])


def _old_table(plan: L.EvictPlan, device) -> Optional[torch.Tensor]:
    """Only when float32 linspace produced duplicates (plan.needs_table): the older region's tokens, built on the device
    (linspace -> trunc -> unique_consecutive as one kernel; the indices never exist on the host)."""
    if not plan.needs_table:
        return None
    table = torch.empty(int(plan.n_old), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        L.check(L.lib().kvq_evict_old_table_device(ctypes.byref(plan), table.data_ptr(), L.stream_ptr(table.device)),
                "kvq_evict_old_table_device")
        L.count_launch()
    return table


def gather_tokens(x: torch.Tensor, plan: L.EvictPlan, table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x[:, :, kept, :] as a new contiguous tensor; x is [B,H,T,D] of any 1/2/4-byte dtype (so the same
    kernel compacts fp caches, INT8 codes and packed INT4 bytes)."""
    L.require_cuda(x)
    if x.dim() != 4:
        raise RuntimeError("expected a [B, H, T, D] tensor")
    B, H, T, D = x.shape
    out = torch.empty((B, H, int(plan.n_total), D), dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out
    if table is None:
        table = _old_table(plan, x.device)
    with L.device_guard(x):
        sh, st = L.shape_stride(x)
        L.check(L.lib().kvq_evict_gather(x.data_ptr(), x.element_size(), sh, st, ctypes.byref(plan),
                                         table.data_ptr() if table is not None else None, out.data_ptr(), None,
                                         L.stream_ptr(x.device)), "kvq_evict_gather")
        L.count_launch()
    return out


def gather_tokens_batched(xs: List[torch.Tensor], plan: L.EvictPlan, table: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """gather_tokens for n same-layout tensors (all layers x {K, V}) in ONE launch; returns views of one [n,B,H,kept,D] buffer."""
    x0 = xs[0]
    B, H, T, D = x0.shape
    out = torch.empty((len(xs), B, H, int(plan.n_total), D), dtype=x0.dtype, device=x0.device)
    outs = list(out.unbind(0))
    if out.numel() == 0:
        return outs
    if table is None:
        table = _old_table(plan, x0.device)
    with L.device_guard(x0):
        sh, st = L.shape_stride(x0)
        tabs = L.ptr_tables(xs, outs)  # alive until the launch is enqueued
        L.check(L.lib().kvq_evict_gather_batched(tabs[0].data_ptr(), tabs[1].data_ptr(), len(xs), x0.element_size(),
                                                 sh, st, ctypes.byref(plan), table.data_ptr() if table is not None else None, None,
                                                 L.stream_ptr(x0.device)), "kvq_evict_gather_batched")
        L.count_launch()
    return outs


def _batchable(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.stride(3) == 1 and (x.size(3) * x.element_size()) % 16 == 0 and x.size(3) * x.element_size() <= 4096 and \
        all((s * x.element_size()) % 16 == 0 for s in x.stride()[:3])


def _trim(past_key_values, strategy: int, **params) -> tuple:
    """All layers x {K, V} that share a layout go through ONE gather launch (the reference loops per layer and per tensor,
    implementations.py:176-190, :218-245, :268-292); anything else one launch per tensor."""
    pkv = list(past_key_values)
    flat = [t for kv in pkv for t in kv]
    if len(flat) > 2 and flat[0].is_cuda and _batchable(flat[0]) and L.same_layout(flat):
        plan = evict_plan(strategy, flat[0].size(2), **params)
        if plan.identity:
            return tuple((k, v) for k, v in pkv)
        outs = gather_tokens_batched(flat, plan)
        return tuple((outs[2 * i], outs[2 * i + 1]) for i in range(len(pkv)))
    trimmed = []
    plans = {}
    for k, v in pkv:
        T = k.size(2)
        ent = plans.get(T)
        if ent is None:
            plan = evict_plan(strategy, T, **params)
            ent = plans[T] = (plan, None if plan.identity else _old_table(plan, k.device))
        plan, table = ent
        if plan.identity:
            trimmed.append((k, v))
            continue
        trimmed.append((gather_tokens(k, plan, table), gather_tokens(v, plan, table)))
    return tuple(trimmed)


def trim_kv_sliding_window(past_key_values: tuple, window_size: int) -> tuple:
    """Keep the last `window_size` tokens.  Like the reference this is a view, not a copy (:136-138)."""
    trimmed = []
    for k, v in past_key_values:
        if k.size(2) > window_size:
            k = k[:, :, -window_size:, :]
            v = v[:, :, -window_size:, :]
        trimmed.append((k, v))
    return tuple(trimmed)


def trim_kv_prefix_window(past_key_values, prefix_len: int, window_size: int):
    """Keep the first prefix_len and the last window_size tokens."""
    return _trim(past_key_values, L.PREFIX_WINDOW, window=window_size, prefix=prefix_len)


def trim_kv_strided(past_key_values, window_size: int, stride: int, prefix_len: int = 0):
    """Keep the prefix, the dense tail, and every `stride`-th token of the older region."""
    assert stride >= 1
    return _trim(past_key_values, L.STRIDED, window=window_size, prefix=prefix_len, stride=stride)


def trim_kv_block_old(past_key_values, window_size: int, block_size: int = 64, keep_per_block: int = 8,
                      prefix_len: int = 0):
    """Keep the prefix, the dense tail, and the last keep_per_block tokens of every block_size block of
    the older region (blocks anchored at prefix_len; the last, partial block keeps what it has)."""
    assert block_size >= 1
    assert 1 <= keep_per_block <= block_size
    return _trim(past_key_values, L.BLOCK, window=window_size, prefix=prefix_len, block=block_size, keep=keep_per_block)


def trim_kv_budget_old(past_key_values, window_size: int, old_budget: int = 64, prefix_len: int = 0):
    """Keep the prefix, the dense tail, and old_budget tokens sampled uniformly (float32 linspace,
    truncated, consecutive duplicates removed) from the older region."""
    assert old_budget >= 0
    return _trim(past_key_values, L.BUDGET, window=window_size, prefix=prefix_len, budget=old_budget)


# ----------------------------------------------------------------------------- chunk summary


def _chunk_mean(x: torch.Tensor, chunk_size: int, keep_last: int) -> torch.Tensor:
    B, H, T, D = x.shape
    keep_eff = min(keep_last, T)
    old_len = T - keep_eff
    if old_len <= 0:
        return x
    n_chunks = (old_len + chunk_size - 1) // chunk_size
    out = torch.empty((B, H, n_chunks + keep_eff, D), dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out
    with L.device_guard(x):
        sh, st = L.shape_stride(x)
        L.check(L.lib().kvq_chunk_mean(x.data_ptr(), L.dtype_code(x.dtype), sh, st, int(chunk_size), int(keep_last),
                                       out.data_ptr(), L.stream_ptr(x.device)), "kvq_chunk_mean")
        L.count_launch()
    return out


def chunk_summarize_kv(past_key_values: tuple, chunk_size: int, keep_last: int) -> tuple:
    """Replace all but the last keep_last tokens by mean-pooled chunk summaries; one fused launch per
    tensor (segmented mean + copy of the recent tokens).  The zero-padded last chunk is divided by
    chunk_size, like the reference."""
    pkv = list(past_key_values)
    for k, v in pkv:
        L.require_cuda(k, v)
    flat = [t for kv in pkv for t in kv]
    if len(flat) > 2 and _batchable(flat[0]) and L.same_layout(flat):  # one launch for all layers x {K, V}
        x0 = flat[0]
        B, H, T, D = x0.shape
        keep_eff = min(keep_last, T)
        old_len = T - keep_eff
        if old_len <= 0:
            return tuple((k, v) for k, v in pkv)
        n_chunks = (old_len + chunk_size - 1) // chunk_size
        out = torch.empty((len(flat), B, H, n_chunks + keep_eff, D), dtype=x0.dtype, device=x0.device)
        outs = list(out.unbind(0))
        if out.numel():
            with L.device_guard(x0):
                sh, st = L.shape_stride(x0)
                tabs = L.ptr_tables(flat, outs)  # alive until the launch is enqueued
                L.check(L.lib().kvq_chunk_mean_batched(tabs[0].data_ptr(), tabs[1].data_ptr(), len(flat),
                                                       L.dtype_code(x0.dtype), sh, st, int(chunk_size), int(keep_last),
                                                       L.stream_ptr(x0.device)), "kvq_chunk_mean_batched")
                L.count_launch()
        return tuple((outs[2 * i], outs[2 * i + 1]) for i in range(len(pkv)))
    summarized = []
    for k, v in pkv:
        if k.size(2) - min(keep_last, k.size(2)) <= 0:
            summarized.append((k, v))
            continue
        summarized.append((_chunk_mean(k, chunk_size, keep_last), _chunk_mean(v, chunk_size, keep_last)))
    return tuple(summarized)


# ----------------------------------------------------------------------------- paged store


class PagedKVCache:
    """Paged KV cache of one layer: K and V live in a physical pool [n_blocks, B, H, block_size, D];
    a device-resident block table maps logical block -> physical block (the reference keeps a Python
    list whose index is the logical block, :36-54).  Physical blocks are handed out from the top of the
    pool downwards so the table is never the identity: the gather kernel really goes through it."""

    def __init__(self, block_size: int = 64, device: str = "cuda", dtype: torch.dtype = torch.float16):
        self.block_size = int(block_size)
        self.device = device
        self.dtype = dtype
        self.t_filled = 0
        self._B: Optional[int] = None
        self._H: Optional[int] = None
        self._D: Optional[int] = None
        self._k_pool: Optional[torch.Tensor] = None
        self._v_pool: Optional[torch.Tensor] = None
        self._table: Optional[torch.Tensor] = None  # int32 [pool capacity], logical -> physical
        self._table_host: List[int] = []
        self._n_blocks = 0

    def num_blocks(self) -> int:
        return self._n_blocks

    @property
    def k_blocks(self) -> List[torch.Tensor]:
        return [self._k_pool[p] for p in self._table_host]

    @property
    def v_blocks(self) -> List[torch.Tensor]:
        return [self._v_pool[p] for p in self._table_host]

    @property
    def block_table(self) -> torch.Tensor:
        return self._table[: self._n_blocks]

    def _alloc_block(self, B: int, H: int, D: int) -> None:
        cap = 0 if self._k_pool is None else self._k_pool.size(0)
        if self._n_blocks == cap:
            new_cap = max(4, cap * 2)
            shape = (new_cap, B, H, self.block_size, D)
            k = torch.empty(shape, device=self.device, dtype=self.dtype)
            v = torch.empty(shape, device=self.device, dtype=self.dtype)
            table = torch.zeros(new_cap, dtype=torch.int32, device=self.device)
            if cap:
                # existing physical ids keep their slot counted from the top of the pool
                k[new_cap - cap:].copy_(self._k_pool)
                v[new_cap - cap:].copy_(self._v_pool)
                self._table_host = [p + (new_cap - cap) for p in self._table_host]
                table[: self._n_blocks] = torch.tensor(self._table_host, dtype=torch.int32)
            self._k_pool, self._v_pool, self._table = k, v, table
            cap = new_cap
        phys = cap - 1 - self._n_blocks
        self._table_host.append(phys)
        self._table[self._n_blocks] = phys
        self._n_blocks += 1

    @torch.no_grad()
    def append(self, k_1tok: torch.Tensor, v_1tok: torch.Tensor) -> None:
        """Append one token's K and V, each [B,H,1,D]."""
        assert k_1tok.dim() == 4 and v_1tok.dim() == 4
        B, H, one, D = k_1tok.shape
        assert one == 1
        L.require_cuda(k_1tok, v_1tok)
        if self._B is None:
            self._B, self._H, self._D = B, H, D
        if self.num_blocks() == 0 or (self.t_filled % self.block_size) == 0:
            self._alloc_block(B, H, D)
        with L.device_guard(k_1tok):
            stream = L.stream_ptr(k_1tok.device)
            for x, pool in ((k_1tok, self._k_pool), (v_1tok, self._v_pool)):
                if x.dtype != self.dtype:
                    x = x.to(self.dtype)  # the reference's slice assignment casts (:78-79)
                sh, st = L.shape_stride(x)
                L.check(L.lib().kvq_paged_append(x.data_ptr(), x.element_size(), sh, st, pool.data_ptr(),
                                                 self._table.data_ptr(), self.block_size, self.t_filled, stream),
                        "kvq_paged_append")
                L.count_launch()
        self.t_filled += 1

    @torch.no_grad()
    def append_tokens(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Append n tokens [B,H,n,D] at once: same result as n append() calls (the reference's prompt loop,
        benchmarker.py:528-530), one launch per tensor."""
        assert k.dim() == 4 and v.dim() == 4
        B, H, n, D = k.shape
        if n == 0:
            return
        L.require_cuda(k, v)
        if self._B is None:
            self._B, self._H, self._D = B, H, D
        while self.num_blocks() * self.block_size < self.t_filled + n:
            self._alloc_block(B, H, D)
        with L.device_guard(k):
            stream = L.stream_ptr(k.device)
            for x, pool in ((k, self._k_pool), (v, self._v_pool)):
                if x.dtype != self.dtype:
                    x = x.to(self.dtype)
                sh, st = L.shape_stride(x)
                L.check(L.lib().kvq_paged_append(x.data_ptr(), x.element_size(), sh, st, pool.data_ptr(),
                                                 self._table.data_ptr(), self.block_size, self.t_filled, stream),
                        "kvq_paged_append")
                L.count_launch()
        self.t_filled += n

    @torch.no_grad()
    def get_kv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Contiguous (K, V), each [B,H,T,D], stitched through the block table in one launch each."""
        if self.t_filled == 0:
            raise ValueError("Empty cache")
        assert self._B is not None and self._H is not None and self._D is not None
        B, H, D, T = self._B, self._H, self._D, self.t_filled
        outs = []
        with L.device_guard(self._k_pool):
            stream = L.stream_ptr(self._k_pool.device)
            for pool in (self._k_pool, self._v_pool):
                out = torch.empty((B, H, T, D), device=self.device, dtype=self.dtype)
                L.check(L.lib().kvq_paged_gather(pool.data_ptr(), pool.element_size(), L.I64x4(B, H, T, D),
                                                 self._table.data_ptr(), self.block_size, out.data_ptr(), stream),
                        "kvq_paged_gather")
                L.count_launch()
                outs.append(out)
        return outs[0], outs[1]

    def allocated_bytes(self) -> int:
        """Bytes of the blocks in use, K and V, including the unused tail of the last block (:108-114)."""
        if self.num_blocks() == 0:
            return 0
        per_block = self._B * self._H * self.block_size * self._D * self.dtype.itemsize
        return 2 * self.num_blocks() * per_block

    def used_bytes(self) -> int:
        if self.t_filled == 0:
            return 0
        assert self._B is not None and self._H is not None and self._D is not None
        return self.t_filled * self._B * self._H * self._D * self.dtype.itemsize * 2
