"""Paged quantized KV store + attention straight off the block table ("next" row 3).

The reference's PagedKVCache (/root/reference/src/cache/implementations.py:18-114) keeps fp blocks in a Python list and
re-concatenates them before every use (`get_kv`, :83-106): paging is simulated, every decode step still copies the whole
cache.  Here the blocks hold INT8 / packed-INT4 codes with one scale per token (the QuantizedLayerKV format,
src/quantization/ops.py:137-269), the block table lives on the device, and the decode-attention kernel reads K/V through
it (`kvq_decode_attn_paged`): nothing is stitched.  Dropping the oldest tokens (sliding window) is a block-table edit.

    pool   codes  [n_phys, B, H, block_size, W]   W = D (INT8) or D/2 (INT4)
           scales [n_phys, block_size]            input dtype (what the reference returns) and fp32 (what attention reads)
    table  int32  [n_logical]                     logical block i -> physical block
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import _lib as L
from ..attention import decode_attention_paged


class _PagedStream:
    """One K or V stream of a layer."""

    def __init__(self, bits: int, block_size: int):
        self.bits, self.bs = bits, block_size
        self.codes: Optional[torch.Tensor] = None
        self.scales: Optional[torch.Tensor] = None
        self.scales_f32: Optional[torch.Tensor] = None
        self.table: Optional[torch.Tensor] = None   # device int32, capacity = pool capacity
        self.table_host: List[int] = []             # logical -> physical, live blocks only
        self.free: List[int] = []
        self.T = 0                                  # tokens stored (logical 0 .. T-1); block 0 is always full-from-0
        self.D = 0
        self._stride3 = None
        self._stride3_of = None
        self._bits_scratch: Optional[torch.Tensor] = None

    def _grow(self, x: torch.Tensor, n_blocks: int) -> None:
        B, H, _, D = x.shape
        W = D if self.bits == 8 else (D + 1) // 2
        cap = 0 if self.codes is None else self.codes.size(0)
        if self.codes is not None and len(self.free) >= n_blocks:
            return
        need = len(self.table_host) + n_blocks
        new_cap = max(4, 2 * cap, 1 << (need - 1).bit_length())
        codes = torch.empty((new_cap, B, H, self.bs, W), dtype=torch.int8 if self.bits == 8 else torch.uint8, device=x.device)
        scales = torch.empty((new_cap, self.bs), dtype=x.dtype, device=x.device)
        s32 = torch.empty((new_cap, self.bs), dtype=torch.float32, device=x.device)
        table = torch.zeros(new_cap, dtype=torch.int32, device=x.device)
        if cap:  # physical ids are stable across growth
            codes[:cap].copy_(self.codes)
            scales[:cap].copy_(self.scales)
            s32[:cap].copy_(self.scales_f32)
            table[:cap].copy_(self.table)
        # hand new blocks out from the top of the pool downwards: the table is never the identity
        self.free.extend(range(cap, new_cap))
        self.codes, self.scales, self.scales_f32, self.table, self.D = codes, scales, s32, table, D

    def _sync_table(self) -> None:
        n = len(self.table_host)
        if n:
            self.table[:n].copy_(torch.tensor(self.table_host, dtype=torch.int32), non_blocking=True)

    def append_per_token(self, x: torch.Tensor, eps: float = 1e-8, group=None) -> None:
        """n tokens, one scale each, quantized straight into their blocks: one launch per block touched.  `group`: the store
        holds a head shard; the per-token absmax is max-combined over the group -- inside the quantize kernel through peer-mapped
        memory (kvq_quant_per_token_fused_exchange), or absmax -> NCCL allreduce(MAX) -> quantize where that is unavailable --
        so codes and scales equal the unsharded store's, bit for bit."""
        B, H, n, D = x.shape
        if self.codes is not None and (tuple(self.codes.shape[1:3]) != (B, H) or self.D != D or self.scales.dtype != x.dtype):
            raise RuntimeError("appended tensor does not match the cache")
        first_new = (self.T + self.bs - 1) // self.bs
        last = (self.T + n + self.bs - 1) // self.bs
        self._grow(x, last - first_new)
        if last > first_new:
            for _ in range(last - first_new):
                self.table_host.append(self.free.pop())
            self._sync_table()
        lib, dt, stream = L.lib(), L.dtype_code(x.dtype), L.stream_ptr(x.device)
        codes = self.codes
        if self._stride3_of is not codes:  # {sB,sH,sT} of one block, bytes == elements for the 1-byte codes
            self._stride3, self._stride3_of = L.I64x3(codes.stride(1), codes.stride(2), codes.stride(3)), codes
        bits = self._bits_scratch
        if bits is None or bits.numel() < min(n, self.bs) or bits.device != x.device:
            bits = self._bits_scratch = torch.empty(max(self.bs, 64), dtype=torch.int32, device=x.device)
        es = self.scales.element_size()
        ex = None
        if group is not None and float(eps) == 1e-8:
            from ..distributed import shared_exchange
            ex = shared_exchange(x.device, group)   # None: no peer-mapped memory on this node -> the allreduce form
        done = 0
        while done < n:  # destination pointers by arithmetic on the pools: the decode loop comes through here every step
            t = self.T + done
            blk, w = divmod(t, self.bs)
            m = min(self.bs - w, n - done)
            phys = self.table_host[blk]
            src = x if m == n else x[:, :, done: done + m]
            sh, st = L.shape_stride(src)
            slot = phys * self.bs + w
            dst = codes.data_ptr() + phys * codes.stride(0) + w * codes.stride(3)
            if group is None:
                L.check(lib.kvq_quant_per_token_fused(src.data_ptr(), dt, sh, st, self.bits, float(eps), dst, self._stride3,
                                                      self.scales.data_ptr() + slot * es, self.scales_f32.data_ptr() + 4 * slot,
                                                      bits.data_ptr(), stream), "kvq_quant_per_token_fused")
                L.count_launch()
                done += m
                continue
            if ex is not None:  # head shard, one launch: the max over ranks inside the quantize kernel (peer-mapped memory)
                rc = lib.kvq_quant_per_token_fused_exchange(src.data_ptr(), dt, sh, st, self.bits, float(eps), dst, self._stride3,
                                                            self.scales.data_ptr() + slot * es, self.scales_f32.data_ptr() + 4 * slot,
                                                            None, ex.peer_table_ptr, ex.rank, ex.world, 0, stream)   # 0: device-resident step counter
                if rc != L.KVQ_ERR_UNSUPPORTED:   # (unsupported layout: the same on every rank -> all take the form below)
                    L.check(rc, "kvq_quant_per_token_fused_exchange")
                    L.count_launch()
                    done += m
                    continue
            # head shard, fallback: absmax of the shard -> allreduce(MAX) -> quantize with the global per-token scale
            from ..distributed import allreduce_absmax_
            gb = torch.empty(m, dtype=torch.int32, device=x.device)
            L.check(lib.kvq_absmax(src.data_ptr(), dt, sh, st, L.PER_TOKEN, gb.data_ptr(), 0, stream), "kvq_absmax")
            allreduce_absmax_(gb, group)
            fn = lib.kvq_quant_int8 if self.bits == 8 else lib.kvq_quant_int4_pack
            L.check(fn(src.data_ptr(), dt, sh, st, L.PER_TOKEN | L.ABSMAX_COVERS_INPUT, gb.data_ptr(), float(eps), dst, self._stride3,
                       self.scales.data_ptr() + slot * es, self.scales_f32.data_ptr() + 4 * slot, stream), "kvq_quant (per token)")
            L.count_launch(2)
            done += m
        self.T += n

    def drop_front_blocks(self, n_blocks: int) -> None:
        if n_blocks <= 0:
            return
        self.free.extend(self.table_host[:n_blocks])
        self.table_host = self.table_host[n_blocks:]
        self.T -= n_blocks * self.bs
        self._sync_table()

    def gather(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Contiguous codes [B,H,T,W] and input-dtype scales [T] (two gather launches)."""
        B, H, W = self.codes.size(1), self.codes.size(2), self.codes.size(4)
        out = torch.empty((B, H, self.T, W), dtype=self.codes.dtype, device=self.codes.device)
        sc = torch.empty(self.T, dtype=self.scales.dtype, device=self.codes.device)
        stream = L.stream_ptr(out.device)
        L.check(L.lib().kvq_paged_gather(self.codes.data_ptr(), 1, L.I64x4(B, H, self.T, W), self.table.data_ptr(), self.bs,
                                         out.data_ptr(), stream), "kvq_paged_gather")
        L.check(L.lib().kvq_paged_gather(self.scales.data_ptr(), self.scales.element_size(), L.I64x4(1, 1, self.T, 1),
                                         self.table.data_ptr(), self.bs, sc.data_ptr(), stream), "kvq_paged_gather")
        L.count_launch(2)
        return out, sc

    def dequantize(self, out_dtype: torch.dtype) -> torch.Tensor:
        codes, sc = self.gather()
        B, H = codes.shape[:2]
        out = torch.empty((B, H, self.T, self.D), dtype=out_dtype, device=codes.device)
        fn = L.lib().kvq_dequant_int8 if self.bits == 8 else L.lib().kvq_dequant_int4_unpack
        L.check(fn(codes.data_ptr(), L.I64x4(B, H, self.T, self.D), L.stride3(codes), sc.data_ptr(), L.dtype_code(sc.dtype),
                   L.PER_TOKEN, out.data_ptr(), L.dtype_code(out_dtype), L.stream_ptr(out.device)), "kvq_dequant (per token)")
        L.count_launch()
        return out


class PagedQuantizedLayerKV:
    """QuantizedLayerKV (modes "int8", "int4", "mixed") on a paged pool; same per-token codes and scales, bit for bit."""

    def __init__(self, mode: str = "int8", block_size: int = 64, device: str = "cuda", compute_dtype: torch.dtype = torch.float16,
                 group=None):
        assert mode in ["int8", "int4", "mixed"], f"Invalid mode: {mode}"
        self.group = group  # torch.distributed group of a head-sharded store (None: single GPU)
        if block_size < 16 or block_size & (block_size - 1):
            raise ValueError("block_size must be a power of two >= 16")
        self.mode, self.block_size, self.device, self.compute_dtype = mode, int(block_size), device, compute_dtype
        self.k_bits = 4 if mode == "int4" else 8
        self.v_bits = 8 if mode == "int8" else 4
        self._k = _PagedStream(self.k_bits, self.block_size)
        self._v = _PagedStream(self.v_bits, self.block_size)
        self._ws: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return self._k.T

    def num_blocks(self) -> int:
        return len(self._k.table_host)

    @property
    def block_table(self) -> torch.Tensor:
        return self._k.table[: self.num_blocks()]

    @torch.no_grad()
    def append_tokens(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Append n tokens [B,H,n,D], each with its own scale (QuantizedKVCache.init_from_prompt_past / append_from_past
        semantics, ops.py:323-342)."""
        L.require_cuda(k, v)
        if k.dim() != 4 or k.shape != v.shape:
            raise RuntimeError("k and v must be [B,H,n,D] of the same shape")
        if k.size(3) % 2:
            raise RuntimeError("head_dim must be even")
        if k.size(2) == 0:
            return
        with L.device_guard(k):
            self._k.append_per_token(k, group=self.group)
            self._v.append_per_token(v, group=self.group)

    append = append_tokens

    @torch.no_grad()
    def get_kv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Dequantized contiguous (K, V) [B,H,T,D] in compute_dtype (the reference's access pattern; attention does not need it)."""
        if len(self) == 0:
            raise ValueError("Empty cache")
        with L.device_guard(self._k.codes):
            return self._k.dequantize(self.compute_dtype), self._v.dequantize(self.compute_dtype)

    @torch.no_grad()
    def quantized_contiguous(self):
        """(k_codes, k_scales, v_codes, v_scales) stitched: for tests / export."""
        with L.device_guard(self._k.codes):
            kc, ks = self._k.gather()
            vc, vs = self._v.gather()
        return kc, ks, vc, vs

    def length_after_sliding_window(self, window_size: int) -> int:
        """len(self) after evict_sliding_window(window_size), without touching the store."""
        return len(self) - max(0, (len(self) - int(window_size)) // self.block_size) * self.block_size

    def evict_sliding_window(self, window_size: int) -> int:
        """Keep at least the last `window_size` tokens by dropping whole leading blocks: a block-table edit, no data moves
        (block-granular version of trim_kv_to_window, implementations.py:124-141).  Returns the new length."""
        drop = max(0, (len(self) - int(window_size)) // self.block_size)
        self._k.drop_front_blocks(drop)
        self._v.drop_front_blocks(drop)
        return len(self)

    @torch.no_grad()
    def attend(self, q: torch.Tensor, sm_scale: Optional[float] = None, k_cur: Optional[torch.Tensor] = None,
               v_cur: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
               seq_len: Optional[int] = None) -> torch.Tensor:
        """Decode attention of q [B,Hq,1,D] over the stored tokens (or the first seq_len of them), read through the block table."""
        if len(self) == 0:
            raise ValueError("Empty cache")
        n = len(self) if seq_len is None else int(seq_len)
        if not 1 <= n <= len(self):
            raise ValueError(f"seq_len must be in [1, {len(self)}]")
        k, v = self._k, self._v
        B, H, bs = k.codes.size(1), k.codes.size(2), self.block_size
        if q.dim() != 4 or q.size(2) != 1 or q.size(0) != B or q.size(3) != k.D or not q.is_cuda:
            raise RuntimeError("q must be a CUDA tensor [B, Hq, 1, D] matching the cache")
        Hq, D = q.size(1), q.size(3)
        if Hq % H or Hq // H not in (1, 2, 4, 8) or (k_cur is None) != (v_cur is None):
            if self._ws is None:
                self._ws = torch.empty(8 << 20, dtype=torch.float32, device=q.device)
            return decode_attention_paged(q, k.codes, k.scales_f32, self.k_bits, v.codes, v.scales_f32, self.v_bits, k.table, n,
                                          sm_scale=sm_scale, out_dtype=out_dtype, workspace=self._ws, k_cur=k_cur, v_cur=v_cur)
        # per-step path of KvqFusedCache: the same C call without the wrapper's tensor bookkeeping
        import ctypes
        if k_cur is not None:
            if tuple(k_cur.shape) != (B, H, 1, D) or tuple(v_cur.shape) != (B, H, 1, D):
                raise RuntimeError("k_cur / v_cur must be [B, H, 1, D]")
            k_cur, v_cur = k_cur.to(q.dtype).contiguous(), v_cur.to(q.dtype).contiguous()
        qc = q.contiguous()
        out_dtype = out_dtype or q.dtype
        out = torch.empty((B, Hq, 1, D), dtype=out_dtype, device=q.device)
        with L.device_guard(q):
            lib = L.lib()
            ns = ctypes.c_int(1)
            nbytes = lib.kvq_decode_attn_workspace_bytes(B, Hq, H, n, D, ctypes.byref(ns))
            ws = self._ws
            if ws is None or ws.numel() * 4 < nbytes or ws.device != q.device:
                ws = self._ws = torch.empty(max(nbytes, 1 << 20) // 4, dtype=torch.float32, device=q.device)
            L.check(lib.kvq_decode_attn_paged(qc.data_ptr(), L.dtype_code(q.dtype), k.codes.data_ptr(), self.k_bits, None,
                                              k.scales_f32.data_ptr(), v.codes.data_ptr(), self.v_bits, None, v.scales_f32.data_ptr(),
                                              k.table.data_ptr(), bs, B, Hq, H, n, D,
                                              float(D) ** -0.5 if sm_scale is None else float(sm_scale),
                                              k_cur.data_ptr() if k_cur is not None else None,
                                              v_cur.data_ptr() if v_cur is not None else None, out.data_ptr(),
                                              L.dtype_code(out_dtype), ws.data_ptr(), ns.value, L.stream_ptr(q.device)),
                    "kvq_decode_attn_paged")
            L.count_launch(2)
        return out

    def allocated_bytes(self) -> int:
        """Bytes of the blocks in use (codes + fp16-sized scales as the reference counts them, ops.py:357-363)."""
        n = self.num_blocks()
        if n == 0:
            return 0
        per_tok = lambda s: s.codes.size(1) * s.codes.size(2) * s.codes.size(4)  # noqa: E731
        return n * self.block_size * (per_tok(self._k) + per_tok(self._v) + 2 * 2)
