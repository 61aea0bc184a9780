"""ctypes binding of libkvq_sm100.so (C ABI: include/kvq.h).

There is no CPU path and no fallback: if the shared library is missing, or a tensor is not on a CUDA
device, the call raises.  torch is used for device memory and streams only; every kernel launch goes
through the C ABI with raw pointers.
"""

from __future__ import annotations

import contextlib
import ctypes
import os
import threading
from typing import Optional, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KVQ_LIB", os.path.join(_HERE, "libkvq_sm100.so"))

F32, F16, BF16 = 0, 1, 2
PER_TENSOR, PER_TOKEN = 0, 1
ABSMAX_COVERS_INPUT = 0x100  # OR into granularity / bits when absmax is the phase-1 result of the same data
SLIDING, PREFIX_WINDOW, STRIDED, BLOCK, BUDGET = 0, 1, 2, 3, 4
ATTN_VARIANTS = {"auto": 0, "cuda_core": 1, "imma": 2}  # enum kvq_attn_variant

DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

KVQ_OK, KVQ_ERR_INVALID, KVQ_ERR_CUDA, KVQ_ERR_UNSUPPORTED, KVQ_ERR_NO_DEVICE = 0, -1, -2, -3, -4

c_i64, c_int, c_f, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p
I64x4 = c_i64 * 4
I64x3 = c_i64 * 3


class EvictPlan(ctypes.Structure):
    """struct kvq_evict_plan (include/kvq.h)."""

    _fields_ = [
        ("strategy", ctypes.c_int32), ("identity", ctypes.c_int32), ("T", c_i64), ("n_prefix", c_i64),
        ("n_old", c_i64), ("old_start", c_i64), ("tail_start", c_i64), ("n_tail", c_i64), ("n_total", c_i64),
        ("stride", c_i64), ("keep", c_i64), ("budget", c_i64), ("needs_table", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


# every symbol include/kvq.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "kvq_version": (c_int, []),
    "kvq_last_error": (ctypes.c_char_p, []),
    "kvq_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "kvq_absmax": (c_int, [c_vp, c_int, I64x4, I64x4, c_int, c_vp, c_int, c_vp]),
    "kvq_quant_int8": (c_int, [c_vp, c_int, I64x4, I64x4, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kvq_quant_int4_pack": (c_int, [c_vp, c_int, I64x4, I64x4, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kvq_quant_per_token_fused": (c_int, [c_vp, c_int, I64x4, I64x4, c_int, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "kvq_dequant_int8": (c_int, [c_vp, I64x4, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp]),
    "kvq_dequant_int4_unpack": (c_int, [c_vp, I64x4, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp]),
    "kvq_absmax_batched": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp]),
    "kvq_quant_batched": (c_int, [c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp]),
    "kvq_dequant_batched": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_int, c_vp]),
    "kvq_quant_per_token_fused_batched": (c_int, [c_vp, c_int, c_i64, I64x4, I64x4, c_int, c_f, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "kvq_quant_per_token_fused_exchange": (c_int, [c_vp, c_int, I64x4, I64x4, c_int, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int,
                                                   ctypes.c_uint32, c_vp]),
    "kvq_evict_gather_batched": (c_int, [c_vp, c_vp, c_i64, c_int, I64x4, I64x4, ctypes.POINTER(EvictPlan), c_vp, c_vp, c_vp]),
    "kvq_evict_old_table_device": (c_int, [ctypes.POINTER(EvictPlan), c_vp, c_vp]),
    "kvq_chunk_mean_batched": (c_int, [c_vp, c_vp, c_i64, c_int, I64x4, I64x4, c_i64, c_i64, c_vp]),
    "kvq_absmax_exchange_buffer_bytes": (c_i64, []),
    "kvq_absmax_exchange": (c_int, [c_vp, c_i64, c_vp, c_int, c_int, ctypes.c_uint32, c_vp]),
    "kvq_absmax_batched_publish": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_int, c_int, ctypes.c_uint32, c_vp]),
    "kvq_quant_batched_gather": (c_int, [c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.c_uint32, c_vp]),
    "kvq_evict_plan_make": (c_int, [c_int] + [c_i64] * 7 + [ctypes.POINTER(EvictPlan)]),
    "kvq_evict_indices_host": (c_int, [ctypes.POINTER(EvictPlan), c_vp]),
    "kvq_evict_indices_device": (c_int, [ctypes.POINTER(EvictPlan), c_vp, c_vp, c_vp]),
    "kvq_evict_gather": (c_int, [c_vp, c_int, I64x4, I64x4, ctypes.POINTER(EvictPlan), c_vp, c_vp, c_vp, c_vp]),
    "kvq_paged_append": (c_int, [c_vp, c_int, I64x4, I64x4, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "kvq_paged_gather": (c_int, [c_vp, c_int, I64x4, c_vp, c_i64, c_vp, c_vp]),
    "kvq_chunk_mean": (c_int, [c_vp, c_int, I64x4, I64x4, c_i64, c_i64, c_vp, c_vp]),
    "kvq_decode_attn_workspace_bytes": (c_i64, [c_i64] * 5 + [ctypes.POINTER(c_int)]),
    "kvq_decode_attn_ex": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp] + [c_i64] * 5
                           + [c_f, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
    "kvq_decode_attn": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp] + [c_i64] * 5
                        + [c_f, c_vp, c_int, c_vp, c_int, c_vp]),
    "kvq_decode_attn_default_variant": (c_int, [c_i64]),
    "kvq_decode_attn_variant": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp] + [c_i64] * 5
                                + [c_f, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp]),
    "kvq_decode_attn_paged_variant": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp] + [c_i64] * 6
                                      + [c_f, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp]),
    "kvq_decode_attn_paged": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp] + [c_i64] * 6
                              + [c_f, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
}

_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()

# launches made through the C ABI since import (bench.py reports it as gpu_launches)
launch_count = 0


def lib() -> ctypes.CDLL:
    """Load libkvq_sm100.so once (thread-safe; the reference's lazy singleton had no lock,
    src/cuda/extensions.py:138-147).  Raises if the library is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `make -C efficient-llm-inference_b200/csrc` "
                        "(or __graft_entry__.build()).  kvq_b200 has no CPU or torch fallback.")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def last_error() -> str:
    msg = lib().kvq_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    """Non-zero status -> RuntimeError (the reference raises RuntimeError through TORCH_CHECK)."""
    if rc != KVQ_OK:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")


def count_launch(n: int = 1) -> None:
    global launch_count
    launch_count += n


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "kvq_b200 runs on CUDA (sm_100a) tensors only; got a tensor on "
                f"'{t.device}'.  There is no CPU fallback by design.")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dt]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {dt}: expected float32, float16 or bfloat16") from None


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device: torch.device) -> int:
    """cudaStream_t of torch's current stream on `device` (the raw getter is ~30x cheaper than building a Stream object:
    the decode loop asks three times per layer per step)."""
    if _raw_stream is not None:
        idx = device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(device).cuda_stream


def as4d(x: torch.Tensor) -> torch.Tensor:
    """View any tensor as [B,H,T,D] without copying when possible (last dim stays last)."""
    if x.dim() == 4:
        return x
    if x.dim() == 0:
        return x.reshape(1, 1, 1, 1)
    if x.dim() < 4:
        return x.reshape((1,) * (4 - x.dim()) + tuple(x.shape))
    return x.reshape(-1, *x.shape[-3:])  # copies only if the leading dims cannot be merged


def shape_stride(x4: torch.Tensor) -> Tuple[ctypes.Array, ctypes.Array]:
    return I64x4(*x4.shape), I64x4(*x4.stride())


def stride3(t4: torch.Tensor) -> ctypes.Array:
    """{sB,sH,sT} of a code/output tensor whose last dim must be contiguous."""
    if t4.size(-1) > 1 and t4.stride(-1) != 1:
        raise RuntimeError("code / output tensors need a contiguous last dimension")
    return I64x3(*t4.stride()[:3])


def ptr_table(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Device int64 array of the data pointers of `tensors` (one small H2D copy): the argument of the *_batched entry points.
    KEEP THE RETURNED TENSOR ALIVE until the launch that reads it has been enqueued: `ptr_table(x).data_ptr()` alone frees
    the table at once and the next allocation reuses its memory."""
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64).to(tensors[0].device, non_blocking=True)


def ptr_tables(*lists: Sequence[torch.Tensor]) -> torch.Tensor:
    """Several pointer tables of equal length in ONE device tensor [k, n] (one H2D copy); row i is `lists[i]`'s table, at
    result[i].data_ptr().  Same lifetime rule as ptr_table."""
    dev = lists[0][0].device
    return torch.tensor([[t.data_ptr() for t in ts] for ts in lists], dtype=torch.int64).to(dev, non_blocking=True)


def same_layout(tensors: Sequence[torch.Tensor]) -> bool:
    """True when the tensors can share one batched launch: same shape / stride / dtype / device, 16-byte-aligned pointers."""
    t0 = tensors[0]
    sh, st, dt, dev = t0.shape, t0.stride(), t0.dtype, t0.device
    for t in tensors:
        if t.shape != sh or t.stride() != st or t.dtype != dt or t.device != dev or t.data_ptr() % 16:
            return False
    return True


_NO_GUARD = contextlib.nullcontext()


def device_guard(t: torch.Tensor):
    """Make t's device current for the launch; free when it already is (the usual case)."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)
