"""Fused dequantize + decode attention over a quantized KV cache (csrc/attn.cu).

Not a function of the reference: there, every decode step dequantizes the whole cache, concatenates it,
wraps it in a DynamicCache and lets transformers compute softmax(q k^T / sqrt(D)) v
(/root/reference/src/benchmarking/benchmarker.py:470-477, src/quantization/ops.py:213-269).  This op
replaces that round trip for the cached tokens: codes are read once and dequantized in registers.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib as L


def run_grouped(q: torch.Tensor, H: int, supported, call):
    """The kernels are built for Hq/H in `supported` (powers of two <= 8).  Any other grouping -- Qwen2-7B's 7 query heads
    per KV head, 3, 5, 6, or more than 8 -- runs as chunks of at most 8 query heads per KV head, each padded with zero
    queries to the next supported size (a zero query gives a uniform softmax that is simply dropped).
    `call(q_sub)` evaluates q_sub [B, H*g, 1, D] with g in `supported` and returns [B, H*g, 1, D]."""
    B, Hq, _, D = q.shape
    if Hq % H:
        raise RuntimeError("Hq must be a multiple of H")
    G = Hq // H
    if G in supported:
        return call(q)
    q5 = q.contiguous().view(B, H, G, 1, D)
    outs = []
    for c0 in range(0, G, 8):
        g = min(8, G - c0)
        gp = min(x for x in supported if x >= g)
        qs = q.new_zeros((B, H, gp, 1, D))
        qs[:, :, :g] = q5[:, :, c0: c0 + g]
        outs.append(call(qs.view(B, H * gp, 1, D)).view(B, H, gp, 1, D)[:, :, :g])
    return torch.cat(outs, dim=2).reshape(B, Hq, 1, D)


@torch.no_grad()
def decode_attention_quantized(q: torch.Tensor, k_codes: torch.Tensor, k_scales: torch.Tensor, k_bits: int,
                               v_codes: torch.Tensor, v_scales: torch.Tensor, v_bits: int,
                               sm_scale: Optional[float] = None, out_dtype: Optional[torch.dtype] = None,
                               workspace: Optional[torch.Tensor] = None, k_cur: Optional[torch.Tensor] = None,
                               v_cur: Optional[torch.Tensor] = None, variant: str = "auto") -> torch.Tensor:
    """out[b,hq] = softmax_t(q[b,hq] . K[b,hq//G,t] * sm_scale) @ V[b,hq//G]  with K/V = codes * scale[t].

    variant: "auto" (default: the library's choice per Hq/H), "cuda_core" (split-K GEMV on CUDA cores) or "imma"
    (tensor-map TMA + INT8 tensor-core MMA); see enum kvq_attn_variant in include/kvq.h.

    q [B,Hq,1,D] (fp32/fp16/bf16); k_codes int8 [B,H,T,D] (k_bits 8) or uint8 [B,H,T,D/2] (k_bits 4);
    k_scales fp32 [T]; v likewise.  D in {64,128}; kernels exist for Hq/H in {1,2,4,8}, any other grouping is run in padded
    chunks (run_grouped).  Returns [B,Hq,1,D].
    k_cur / v_cur (optional, [B,H,1,D]): the decode step's own token, attended to exactly (unquantized) on top
    of the T cached tokens, as the reference's loop does.
    """
    L.require_cuda(q, k_codes, v_codes, k_scales, v_scales)
    if q.dim() != 4 or q.size(2) != 1:
        raise RuntimeError("q must be [B, Hq, 1, D] (one decode token)")
    B, Hq, _, D = q.shape
    H, T = k_codes.size(1), k_codes.size(2)
    if H and Hq % H == 0 and Hq // H not in (1, 2, 4, 8):
        return run_grouped(q, H, (1, 2, 4, 8), lambda qs: decode_attention_quantized(
            qs, k_codes, k_scales, k_bits, v_codes, v_scales, v_bits, sm_scale=sm_scale if sm_scale is not None else float(D) ** -0.5,
            out_dtype=out_dtype, workspace=workspace, k_cur=k_cur, v_cur=v_cur, variant=variant))
    if k_codes.size(0) != B or v_codes.size(0) != B or v_codes.size(1) != H or v_codes.size(2) != T:
        raise RuntimeError("K / V code tensors do not match q")
    if k_scales.dtype != torch.float32 or v_scales.dtype != torch.float32:
        raise RuntimeError("attention reads fp32 scales (QuantizedLayerKV.quantized())")
    if k_scales.numel() != T or v_scales.numel() != T:
        raise RuntimeError("need one scale per cached token")
    if sm_scale is None:
        sm_scale = float(D) ** -0.5
    out_dtype = out_dtype or q.dtype
    qc = q.contiguous()
    if (k_cur is None) != (v_cur is None):
        raise RuntimeError("k_cur and v_cur must be given together")
    if k_cur is not None:
        L.require_cuda(k_cur, v_cur)
        if tuple(k_cur.shape) != (B, H, 1, D) or tuple(v_cur.shape) != (B, H, 1, D):
            raise RuntimeError("k_cur / v_cur must be [B, H, 1, D]")
        k_cur, v_cur = k_cur.to(q.dtype).contiguous(), v_cur.to(q.dtype).contiguous()
    k_scales, v_scales = k_scales.contiguous(), v_scales.contiguous()
    out = torch.empty((B, Hq, 1, D), dtype=out_dtype, device=q.device)
    with L.device_guard(q):
        ns = ctypes.c_int(1)
        nbytes = L.lib().kvq_decode_attn_workspace_bytes(B, Hq, H, T, D, ctypes.byref(ns))
        if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
            workspace = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=q.device)
        L.check(L.lib().kvq_decode_attn_variant(qc.data_ptr(), L.dtype_code(q.dtype), k_codes.data_ptr(), int(k_bits),
                                                L.stride3(k_codes), k_scales.data_ptr(), v_codes.data_ptr(), int(v_bits),
                                                L.stride3(v_codes), v_scales.data_ptr(), B, Hq, H, T, D, float(sm_scale),
                                                k_cur.data_ptr() if k_cur is not None else None,
                                                v_cur.data_ptr() if v_cur is not None else None,
                                                out.data_ptr(), L.dtype_code(out_dtype), workspace.data_ptr(), ns.value,
                                                L.ATTN_VARIANTS[variant], L.stream_ptr(q.device)), "kvq_decode_attn")
        L.count_launch(2)
    return out


@torch.no_grad()
def decode_attention_paged(q: torch.Tensor, k_pool: torch.Tensor, k_scale_pool: torch.Tensor, k_bits: int,
                           v_pool: torch.Tensor, v_scale_pool: torch.Tensor, v_bits: int, block_table: torch.Tensor,
                           seq_len: int, sm_scale: Optional[float] = None, out_dtype: Optional[torch.dtype] = None,
                           workspace: Optional[torch.Tensor] = None, k_cur: Optional[torch.Tensor] = None,
                           v_cur: Optional[torch.Tensor] = None, variant: str = "auto") -> torch.Tensor:
    """decode_attention_quantized over a PAGED store: the kernel walks the block table, nothing is stitched.

    k_pool int8 [n_phys,B,H,block_size,D] (k_bits 8) or uint8 [..., D/2] (k_bits 4), contiguous; k_scale_pool fp32
    [n_phys, block_size]; v likewise; block_table int32 [>= ceil(seq_len / block_size)] (device): logical -> physical
    block; seq_len = tokens stored.  Kernels for Hq/H in {1,2,4,8}, other groupings in padded chunks; block_size a power of two >= 16.
    (The reference's PagedKVCache re-concatenates its blocks before every use: implementations.py:83-106.)"""
    L.require_cuda(q, k_pool, v_pool, k_scale_pool, v_scale_pool, block_table)
    if q.dim() != 4 or q.size(2) != 1:
        raise RuntimeError("q must be [B, Hq, 1, D] (one decode token)")
    B, Hq, _, D = q.shape
    if k_pool.dim() != 5 or v_pool.dim() != 5 or not k_pool.is_contiguous() or not v_pool.is_contiguous():
        raise RuntimeError("pools must be contiguous [n_phys, B, H, block_size, row]")
    H, bs = k_pool.size(2), k_pool.size(3)
    if k_pool.size(1) != B or v_pool.size(1) != B or v_pool.size(2) != H or v_pool.size(3) != bs:
        raise RuntimeError("K / V pools do not match q")
    if H and Hq % H == 0 and Hq // H not in (1, 2, 4, 8):
        return run_grouped(q, H, (1, 2, 4, 8), lambda qs: decode_attention_paged(
            qs, k_pool, k_scale_pool, k_bits, v_pool, v_scale_pool, v_bits, block_table, seq_len,
            sm_scale=sm_scale if sm_scale is not None else float(D) ** -0.5, out_dtype=out_dtype, workspace=workspace,
            k_cur=k_cur, v_cur=v_cur, variant=variant))
    if k_scale_pool.dtype != torch.float32 or v_scale_pool.dtype != torch.float32:
        raise RuntimeError("attention reads fp32 scales")
    if tuple(k_scale_pool.shape) != (k_pool.size(0), bs) or tuple(v_scale_pool.shape) != (v_pool.size(0), bs):
        raise RuntimeError("scale pools must be [n_phys, block_size]")
    if block_table.dtype != torch.int32 or block_table.numel() * bs < seq_len or seq_len < 1:
        raise RuntimeError("block_table must be int32 and cover seq_len tokens")
    if sm_scale is None:
        sm_scale = float(D) ** -0.5
    out_dtype = out_dtype or q.dtype
    qc = q.contiguous()
    if (k_cur is None) != (v_cur is None):
        raise RuntimeError("k_cur and v_cur must be given together")
    if k_cur is not None:
        L.require_cuda(k_cur, v_cur)
        if tuple(k_cur.shape) != (B, H, 1, D) or tuple(v_cur.shape) != (B, H, 1, D):
            raise RuntimeError("k_cur / v_cur must be [B, H, 1, D]")
        k_cur, v_cur = k_cur.to(q.dtype).contiguous(), v_cur.to(q.dtype).contiguous()
    k_scale_pool, v_scale_pool, block_table = k_scale_pool.contiguous(), v_scale_pool.contiguous(), block_table.contiguous()
    out = torch.empty((B, Hq, 1, D), dtype=out_dtype, device=q.device)
    with L.device_guard(q):
        ns = ctypes.c_int(1)
        nbytes = L.lib().kvq_decode_attn_workspace_bytes(B, Hq, H, seq_len, D, ctypes.byref(ns))
        if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
            workspace = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=q.device)
        L.check(L.lib().kvq_decode_attn_paged_variant(qc.data_ptr(), L.dtype_code(q.dtype), k_pool.data_ptr(), int(k_bits), None,
                                                      k_scale_pool.data_ptr(), v_pool.data_ptr(), int(v_bits), None,
                                                      v_scale_pool.data_ptr(), block_table.data_ptr(), int(bs), B, Hq, H, int(seq_len), D,
                                                      float(sm_scale), k_cur.data_ptr() if k_cur is not None else None,
                                                      v_cur.data_ptr() if v_cur is not None else None, out.data_ptr(),
                                                      L.dtype_code(out_dtype), workspace.data_ptr(), ns.value, L.ATTN_VARIANTS[variant],
                                                      L.stream_ptr(q.device)), "kvq_decode_attn_paged")
        L.count_launch(2)
    return out
