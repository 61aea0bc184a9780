"""ctypes front end of the CPU oracle (oracle/kvq_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of kvq_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(efficient-llm-inference_b200/kvq_b200) never does.

The functions take and return torch CPU tensors and carry the reference's names and return arities
(src/quantization/ops.py, src/cache/implementations.py in /root/reference) so that parity tests read
like calls into the reference.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkvq_oracle.so")

F32, F16, BF16 = 0, 1, 2
SLIDING, PREFIX_WINDOW, STRIDED, BLOCK, BUDGET = 0, 1, 2, 3, 4

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "kvq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CC": "gcc"})
    return _SO


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        c_i64, c_f, c_vp, c_int = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_int
        L.kvqo_num_threads.restype = c_int
        L.kvqo_set_num_threads.argtypes = [c_int]
        L.kvqo_absmax.restype = c_f
        L.kvqo_absmax.argtypes = [c_vp, c_int, c_i64]
        L.kvqo_scale_from_absmax.restype = c_f
        L.kvqo_scale_from_absmax.argtypes = [c_f, c_f, c_f]
        L.kvqo_quant_int8_given_absmax.restype = c_f
        L.kvqo_quant_int8_given_absmax.argtypes = [c_vp, c_int, c_i64, c_f, c_f, c_vp, c_vp]
        L.kvqo_quant_int8.restype = c_f
        L.kvqo_quant_int8.argtypes = [c_vp, c_int, c_i64, c_f, c_vp, c_vp]
        L.kvqo_quant_int4_given_absmax.restype = c_f
        L.kvqo_quant_int4_given_absmax.argtypes = [c_vp, c_int, c_i64, c_i64, c_f, c_f, c_vp, c_vp]
        L.kvqo_quant_int4.restype = c_f
        L.kvqo_quant_int4.argtypes = [c_vp, c_int, c_i64, c_i64, c_f, c_vp, c_vp]
        L.kvqo_dequant_int8.restype = None
        L.kvqo_dequant_int8.argtypes = [c_vp, c_i64, c_vp, c_int, c_vp, c_int]
        L.kvqo_dequant_int4.restype = None
        L.kvqo_dequant_int4.argtypes = [c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_int]
        L.kvqo_quant_per_token.restype = None
        L.kvqo_quant_per_token.argtypes = [c_vp, c_int, c_i64, c_i64, c_i64, c_int, c_f, c_vp, c_vp, c_vp]
        L.kvqo_dequant_per_token.restype = None
        L.kvqo_dequant_per_token.argtypes = [c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_int]
        L.kvqo_evict_indices.restype = c_i64
        L.kvqo_evict_indices.argtypes = [c_int] + [c_i64] * 7 + [c_vp]
        L.kvqo_gather_rows.restype = None
        L.kvqo_gather_rows.argtypes = [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]
        L.kvqo_chunk_mean.restype = c_i64
        L.kvqo_chunk_mean.argtypes = [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]
        L.kvqo_decode_attn.restype = None
        L.kvqo_decode_attn.argtypes = [c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp] + [c_i64] * 5 + [c_f, c_vp]
        L.kvqo_check_division_free.restype = c_i64
        L.kvqo_check_division_free.argtypes = [c_int, ctypes.c_uint16, c_f, c_f]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().kvqo_num_threads())


def set_num_threads(n: int) -> None:
    lib().kvqo_set_num_threads(int(n))


def _c(x: torch.Tensor) -> torch.Tensor:
    assert x.device.type == "cpu", "the oracle is a CPU checker"
    return x.contiguous()


def _p(x: torch.Tensor) -> int:
    return x.data_ptr()


# ----------------------------------------------------------------------------- area 1


def absmax(x: torch.Tensor) -> float:
    x = _c(x)
    return float(lib().kvqo_absmax(_p(x), _DT[x.dtype], x.numel()))


def quantize_int8_per_tensor(x: torch.Tensor, eps: float = 1e-8, absmax_override: Optional[float] = None
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """ops.py:10-30 -> (int8 codes shaped like x, 0-dim scale in x.dtype)."""
    x = _c(x)
    q = torch.empty(x.shape, dtype=torch.int8)
    scale = torch.empty((), dtype=x.dtype)
    if absmax_override is None:
        lib().kvqo_quant_int8(_p(x), _DT[x.dtype], x.numel(), eps, _p(q), _p(scale))
    else:
        lib().kvqo_quant_int8_given_absmax(_p(x), _DT[x.dtype], x.numel(), absmax_override, eps, _p(q), _p(scale))
    return q, scale


def quantize_int4_per_tensor_packed(x: torch.Tensor, eps: float = 1e-8, absmax_override: Optional[float] = None
                                    ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """ops.py:33-65 -> (uint8 packed [..., ceil(D/2)], 0-dim scale in x.dtype, orig_last)."""
    x = _c(x)
    d = x.shape[-1]
    rows = x.numel() // d if d else 0
    packed = torch.empty((*x.shape[:-1], (d + 1) // 2), dtype=torch.uint8)
    scale = torch.empty((), dtype=x.dtype)
    if absmax_override is None:
        lib().kvqo_quant_int4(_p(x), _DT[x.dtype], rows, d, eps, _p(packed), _p(scale))
    else:
        lib().kvqo_quant_int4_given_absmax(_p(x), _DT[x.dtype], rows, d, absmax_override, eps, _p(packed), _p(scale))
    return packed, scale, d


def dequantize_int8_per_tensor(q: torch.Tensor, scale: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """ops.py:68-90 (torch branch)."""
    q, scale = _c(q), _c(scale)
    out = torch.empty(q.shape, dtype=out_dtype)
    lib().kvqo_dequant_int8(_p(q), q.numel(), _p(scale), _DT[scale.dtype], _p(out), _DT[out_dtype])
    return out


def dequantize_int4_per_tensor_packed(packed: torch.Tensor, scale: torch.Tensor, orig_last_dim: int,
                                      out_dtype: torch.dtype) -> torch.Tensor:
    """ops.py:93-133 (torch branch)."""
    packed, scale = _c(packed), _c(scale)
    p = packed.shape[-1]
    rows = packed.numel() // p if p else 0
    out = torch.empty((*packed.shape[:-1], orig_last_dim), dtype=out_dtype)
    lib().kvqo_dequant_int4(_p(packed), rows, p, orig_last_dim, _p(scale), _DT[scale.dtype], _p(out), _DT[out_dtype])
    return out


def quantize_per_token(x: torch.Tensor, bits: int, eps: float = 1e-8):
    """One scale per token of x [B,H,T,D] (ops.py:333-342 calling :174-210).
    -> (codes [B,H,T,D] int8 | [B,H,T,ceil(D/2)] uint8, scales [T] in x.dtype, scales_f32 [T])."""
    x = _c(x)
    B, H, T, D = x.shape
    q = torch.empty((B, H, T, D), dtype=torch.int8) if bits == 8 else torch.empty((B, H, T, (D + 1) // 2), dtype=torch.uint8)
    scales = torch.empty((T,), dtype=x.dtype)
    s32 = torch.empty((T,), dtype=torch.float32)
    lib().kvqo_quant_per_token(_p(x), _DT[x.dtype], B * H, T, D, bits, eps, _p(q), _p(scales), _p(s32))
    return q, scales, s32


def dequantize_per_token(q: torch.Tensor, bits: int, scales: torch.Tensor, D: int, out_dtype: torch.dtype) -> torch.Tensor:
    """ops.py:213-269: every token dequantized with its own scale, concatenated along dim 2."""
    q, scales = _c(q), _c(scales)
    B, H, T = q.shape[:3]
    out = torch.empty((B, H, T, D), dtype=out_dtype)
    lib().kvqo_dequant_per_token(_p(q), bits, B * H, T, D, _p(scales), _DT[scales.dtype], _p(out), _DT[out_dtype])
    return out


# ----------------------------------------------------------------------------- area 2


def evict_indices(strategy: int, T: int, window: int = 0, prefix: int = 0, stride: int = 1, block: int = 1,
                  keep: int = 1, budget: int = 0) -> torch.Tensor:
    """Kept token indices of trim_kv_* (implementations.py:124-292) for a cache of length T."""
    buf = np.empty(T + max(prefix, 0) + 1, dtype=np.int64)
    n = lib().kvqo_evict_indices(strategy, T, window, prefix, stride, block, keep, budget, buf.ctypes.data)
    return torch.from_numpy(buf[:n].copy())


def gather_tokens(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x[:, :, idx, :] for x [B,H,T,D] (index_select + cat of implementations.py:174-189)."""
    x, idx = _c(x), _c(idx.to(torch.int64))
    B, H, T, D = x.shape
    out = torch.empty((B, H, idx.numel(), D), dtype=x.dtype)
    lib().kvqo_gather_rows(_p(x), B * H, T, D * x.element_size(), _p(idx), idx.numel(), _p(out))
    return out


def _trim(pkv, strategy, **kw):
    out = []
    for k, v in pkv:
        T = k.size(2)
        idx = evict_indices(strategy, T, **kw)
        if idx.numel() == T and strategy != PREFIX_WINDOW:
            out.append((k, v))
        elif idx.numel() == T and bool((idx == torch.arange(T)).all()):
            out.append((k, v))
        else:
            out.append((gather_tokens(k, idx), gather_tokens(v, idx)))
    return tuple(out)


def trim_kv_sliding_window(pkv, window_size: int):
    return _trim(pkv, SLIDING, window=window_size)


def trim_kv_prefix_window(pkv, prefix_len: int, window_size: int):
    return _trim(pkv, PREFIX_WINDOW, window=window_size, prefix=prefix_len)


def trim_kv_strided(pkv, window_size: int, stride: int, prefix_len: int = 0):
    assert stride >= 1
    return _trim(pkv, STRIDED, window=window_size, prefix=prefix_len, stride=stride)


def trim_kv_block_old(pkv, window_size: int, block_size: int = 64, keep_per_block: int = 8, prefix_len: int = 0):
    assert block_size >= 1
    assert 1 <= keep_per_block <= block_size
    return _trim(pkv, BLOCK, window=window_size, prefix=prefix_len, block=block_size, keep=keep_per_block)


def trim_kv_budget_old(pkv, window_size: int, old_budget: int = 64, prefix_len: int = 0):
    assert old_budget >= 0
    return _trim(pkv, BUDGET, window=window_size, prefix=prefix_len, budget=old_budget)


# ----------------------------------------------------------------------------- area 3


def chunk_mean(x: torch.Tensor, chunk_size: int, keep_last: int) -> torch.Tensor:
    x = _c(x)
    B, H, T, D = x.shape
    Tn = lib().kvqo_chunk_mean(_p(x), _DT[x.dtype], B * H, T, D, chunk_size, keep_last, None)
    if Tn == T and T - min(keep_last, T) <= 0:
        return x
    out = torch.empty((B, H, Tn, D), dtype=x.dtype)
    lib().kvqo_chunk_mean(_p(x), _DT[x.dtype], B * H, T, D, chunk_size, keep_last, _p(out))
    return out


def chunk_summarize_kv(pkv, chunk_size: int, keep_last: int):
    """implementations.py:295-346."""
    return tuple((chunk_mean(k, chunk_size, keep_last), chunk_mean(v, chunk_size, keep_last)) for k, v in pkv)


# ----------------------------------------------------------------------------- area 4


def decode_attention(q: torch.Tensor, kq: torch.Tensor, k_bits: int, k_scales: torch.Tensor, vq: torch.Tensor,
                     v_bits: int, v_scales: torch.Tensor, sm_scale: Optional[float] = None) -> torch.Tensor:
    """softmax(q k^T * sm_scale) v over dequant(codes, scale[t]); q [B,Hq,1,D] -> fp32 [B,Hq,1,D]."""
    B, Hq, one, D = q.shape
    assert one == 1
    H, T = kq.shape[1], kq.shape[2]
    if sm_scale is None:
        sm_scale = float(D) ** -0.5
    qf = _c(q.float()).view(B, Hq, D)
    kq, vq = _c(kq), _c(vq)
    ks, vs = _c(k_scales.float()), _c(v_scales.float())
    out = torch.empty((B, Hq, D), dtype=torch.float32)
    lib().kvqo_decode_attn(_p(qf), _p(kq), k_bits, _p(ks), _p(vq), v_bits, _p(vs), B, Hq, H, T, D, sm_scale, _p(out))
    return out.view(B, Hq, 1, D)


def check_division_free(dtype: torch.dtype, absmax_bits: int, qmax: float, eps: float = 1e-8) -> int:
    return int(lib().kvqo_check_division_free(_DT[dtype], absmax_bits, qmax, eps))
