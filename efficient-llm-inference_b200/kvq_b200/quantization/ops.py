"""INT8 / packed-INT4 KV quantization on B200: the reference's `src/quantization/ops.py` API over
hand-written sm_100a kernels (csrc/quant.cu) reached through the C ABI (include/kvq.h).

Same names, argument meaning, return arities and error behaviour as the reference
(/root/reference/src/quantization/ops.py):
    quantize_int8_per_tensor          :10-30     -> (int8 codes, 0-dim scale in x.dtype)
    quantize_int4_per_tensor_packed   :33-65     -> (uint8 packed, 0-dim scale, orig_last)
    dequantize_int8_per_tensor        :68-90
    dequantize_int4_per_tensor_packed :93-133
    QuantizedLayerKV                  :136-290
    QuantizedKVCache                  :293-363

Differences that are not visible in results: scales stay on the device (no float(scale) sync, reference
:87,:117); the per-layer store is one preallocated [B,H,cap,D] code tensor plus a [cap] scale vector
instead of Python lists of per-token tensors, so get_kv() is ONE launch per tensor instead of T launches
and a torch.cat; init_from_prompt_past() is ONE launch per tensor instead of T.
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import _lib as L
from ..distributed import allreduce_absmax_


def _absmax_bits(x4: torch.Tensor, granularity: int, group=None) -> torch.Tensor:
    """Phase 1: max|x| as fp32 bit patterns (int32 tensor; non-negative so integer max == float max)."""
    n = x4.size(2) if granularity == L.PER_TOKEN else 1
    bits = torch.empty(n, dtype=torch.int32, device=x4.device)
    sh, st = L.shape_stride(x4)
    L.check(L.lib().kvq_absmax(x4.data_ptr(), L.dtype_code(x4.dtype), sh, st, granularity, bits.data_ptr(), 0,
                               L.stream_ptr(x4.device)), "kvq_absmax")
    L.count_launch()
    allreduce_absmax_(bits, group)  # head-sharded cache: the only collective on the path
    return bits


def _check_input(x: torch.Tensor) -> None:
    L.require_cuda(x)
    L.dtype_code(x.dtype)
    if x.numel() == 0:
        # x_fp32.abs().max() on an empty tensor raises in the reference (ops.py:27)
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")


@torch.no_grad()
def quantize_int8_per_tensor(x: torch.Tensor, eps: float = 1e-8, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Symmetric per-tensor INT8: q = clamp(round(x/scale), -127, 127), scale = max(max|x|/127, eps).

    Two launches: absmax reduction, then quantize.  `group`: optional torch.distributed group over
    which the tensor is head-sharded; the max is all-reduced so every shard uses the global scale.
    """
    _check_input(x)
    with L.device_guard(x):
        x4 = L.as4d(x)
        bits = _absmax_bits(x4, L.PER_TENSOR, group)
        q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
        scale = torch.empty((), dtype=x.dtype, device=x.device)
        sh, st = L.shape_stride(x4)
        L.check(L.lib().kvq_quant_int8(x4.data_ptr(), L.dtype_code(x.dtype), sh, st, L.PER_TENSOR | L.ABSMAX_COVERS_INPUT,
                                       bits.data_ptr(), float(eps), q.data_ptr(), None, scale.data_ptr(), None,
                                       L.stream_ptr(x.device)), "kvq_quant_int8")
        L.count_launch()
    return q, scale


@torch.no_grad()
def quantize_int4_per_tensor_packed(x: torch.Tensor, eps: float = 1e-8, group=None
                                    ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Symmetric per-tensor INT4 in [-8, 7], two codes per byte along the last dim (even index in the
    high nibble; an odd last dim is padded with one zero code)."""
    _check_input(x)
    if x.dim() == 0:
        raise IndexError("Dimension specified as -1 but tensor has no dimensions")  # q.size(-1), ops.py:53
    with L.device_guard(x):
        x4 = L.as4d(x)
        bits = _absmax_bits(x4, L.PER_TENSOR, group)
        orig_last = x.size(-1)
        packed = torch.empty((*x.shape[:-1], (orig_last + 1) // 2), dtype=torch.uint8, device=x.device)
        scale = torch.empty((), dtype=x.dtype, device=x.device)
        sh, st = L.shape_stride(x4)
        L.check(L.lib().kvq_quant_int4_pack(x4.data_ptr(), L.dtype_code(x.dtype), sh, st, L.PER_TENSOR | L.ABSMAX_COVERS_INPUT,
                                            bits.data_ptr(), float(eps), packed.data_ptr(), None, scale.data_ptr(), None,
                                            L.stream_ptr(x.device)), "kvq_quant_int4_pack")
        L.count_launch()
    return packed, scale, orig_last


def _scale_arg(scale, device) -> torch.Tensor:
    if not isinstance(scale, torch.Tensor):
        scale = torch.tensor(float(scale), dtype=torch.float32, device=device)
    L.require_cuda(scale)
    if scale.dtype not in L.DTYPE_CODE:
        scale = scale.float()
    if scale.numel() != 1:
        raise RuntimeError(f"per-tensor dequantize takes ONE scale, got a tensor of {scale.numel()} elements")
    return scale.reshape(1)


@torch.no_grad()
def dequantize_int8_per_tensor(q: torch.Tensor, scale: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """(q.float() * scale.float()).to(out_dtype) in one launch; any out_dtype in fp32/fp16/bf16
    (the reference's kernel handled fp16 only, ops.py:85)."""
    L.require_cuda(q)
    if q.dtype != torch.int8:
        raise RuntimeError("q must be int8")  # extensions.py:72
    scale = _scale_arg(scale, q.device)
    out = torch.empty(q.shape, dtype=out_dtype, device=q.device)
    if q.numel() == 0:
        return out
    with L.device_guard(q):
        q4 = L.as4d(q.contiguous())
        sh, _ = L.shape_stride(q4)
        L.check(L.lib().kvq_dequant_int8(q4.data_ptr(), sh, None, scale.data_ptr(), L.dtype_code(scale.dtype),
                                         L.PER_TENSOR, out.data_ptr(), L.dtype_code(out_dtype),
                                         L.stream_ptr(q.device)), "kvq_dequant_int8")
        L.count_launch()
    return out


@torch.no_grad()
def dequantize_int4_per_tensor_packed(packed: torch.Tensor, scale: torch.Tensor, orig_last_dim: int,
                                      out_dtype: torch.dtype) -> torch.Tensor:
    """Unpack (high nibble -> even index), subtract 8, keep the first orig_last_dim of each row,
    multiply by scale; one launch, each byte read once."""
    L.require_cuda(packed)
    if packed.dtype != torch.uint8:
        raise RuntimeError("packed must be uint8")  # extensions.py:90
    if packed.dim() < 1:
        raise RuntimeError("packed must have at least 1 dim")  # extensions.py:93
    orig_last_dim = int(orig_last_dim)
    P = packed.size(-1)
    if not (0 <= orig_last_dim <= 2 * P):
        raise RuntimeError(f"orig_last_dim {orig_last_dim} does not fit {P} packed bytes per row")
    scale = _scale_arg(scale, packed.device)
    out = torch.empty((*packed.shape[:-1], orig_last_dim), dtype=out_dtype, device=packed.device)
    if out.numel() == 0:
        return out
    with L.device_guard(packed):
        p4 = L.as4d(packed.contiguous())
        B, H, T = p4.shape[:3]
        sh = L.I64x4(B, H, T, orig_last_dim)
        s3 = L.I64x3(H * T * P, T * P, P)  # rows are P bytes apart even when orig_last < 2P
        L.check(L.lib().kvq_dequant_int4_unpack(p4.data_ptr(), sh, s3, scale.data_ptr(), L.dtype_code(scale.dtype),
                                                L.PER_TENSOR, out.data_ptr(), L.dtype_code(out_dtype),
                                                L.stream_ptr(packed.device)), "kvq_dequant_int4_unpack")
        L.count_launch()
    return out


class TensorTable:
    """Device-resident table of the data pointers of n same-shape contiguous CUDA tensors, for the batched
    (one launch per phase for a whole step) entry points.  Build it once and reuse it: creating it copies n
    pointers to the device."""

    def __init__(self, tensors):
        tensors = list(tensors)
        if not tensors:
            raise ValueError("empty tensor list")
        L.require_cuda(*tensors)
        t0 = tensors[0]
        for t in tensors:
            if t.shape != t0.shape or t.dtype != t0.dtype or t.device != t0.device or not t.is_contiguous():
                raise RuntimeError("batched entry points need same-shape, same-dtype, contiguous tensors on one device")
        self.tensors = tensors
        self.shape, self.dtype, self.device, self.numel = t0.shape, t0.dtype, t0.device, t0.numel()
        self.ptrs = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64).to(t0.device, non_blocking=True)

    def __len__(self):
        return len(self.tensors)


@torch.no_grad()
def quantize_per_tensor_batched(xs: "TensorTable", codes: "TensorTable", bits: int, scales: torch.Tensor,
                                eps: float = 1e-8, group=None, absmax_bits: Optional[torch.Tensor] = None,
                                exchange=None) -> torch.Tensor:
    """quantize_int8_per_tensor / quantize_int4_per_tensor_packed applied to every tensor of `xs` (each with its own
    per-tensor scale) in TWO launches total.  Codes go to the preallocated tensors of `codes`, scale.to(x.dtype) of tensor i
    to scales[i].  Returns the absmax bit patterns (over all ranks when head-sharded).

    Head-sharded (scale over all heads, reference ops.py:27,:48): pass `group` for one NCCL allreduce(MAX) between the two
    launches, or `exchange` (an available kvq_b200.distributed.AbsmaxExchange) for the form with NO collective launch: the
    absmax kernel publishes each tensor's word into every rank's peer-mapped buffer and the quantize kernel gathers them."""
    n = len(xs)
    if len(codes) != n or scales.numel() < n or scales.dtype != xs.dtype:
        raise RuntimeError("codes / scales do not match the input table")
    if bits == 4 and xs.shape[-1] % 2:
        raise RuntimeError("batched INT4 needs an even last dimension")
    if absmax_bits is None:
        absmax_bits = torch.empty(n, dtype=torch.int32, device=xs.device)
    if exchange is not None and getattr(exchange, "available", False) and xs.numel:
        with torch.cuda.device(xs.device):
            st, seq = L.stream_ptr(xs.device), exchange.next_fused_seq()
            L.check(L.lib().kvq_absmax_batched_publish(xs.ptrs.data_ptr(), L.dtype_code(xs.dtype), n, xs.numel, absmax_bits.data_ptr(),
                                                       exchange.peer_table_ptr, exchange.rank, exchange.world, seq, st),
                    "kvq_absmax_batched_publish")
            L.check(L.lib().kvq_quant_batched_gather(xs.ptrs.data_ptr(), L.dtype_code(xs.dtype), n, xs.numel, int(bits) | L.ABSMAX_COVERS_INPUT,
                                                     absmax_bits.data_ptr(), float(eps), codes.ptrs.data_ptr(), scales.data_ptr(), None,
                                                     exchange.local_buffer_ptr, exchange.world, seq, st), "kvq_quant_batched_gather")
            L.count_launch(2)
        return absmax_bits
    if exchange is not None and group is None:
        group = True if exchange.group is None else exchange.group  # exchange not available on this node: the allreduce
    with torch.cuda.device(xs.device):
        st = L.stream_ptr(xs.device)
        L.check(L.lib().kvq_absmax_batched(xs.ptrs.data_ptr(), L.dtype_code(xs.dtype), n, xs.numel, absmax_bits.data_ptr(), st),
                "kvq_absmax_batched")
        allreduce_absmax_(absmax_bits, group)
        L.check(L.lib().kvq_quant_batched(xs.ptrs.data_ptr(), L.dtype_code(xs.dtype), n, xs.numel, int(bits) | L.ABSMAX_COVERS_INPUT, absmax_bits.data_ptr(),
                                          float(eps), codes.ptrs.data_ptr(), scales.data_ptr(), None, st), "kvq_quant_batched")
        L.count_launch(2)
    return absmax_bits


@torch.no_grad()
def dequantize_per_tensor_batched(codes: "TensorTable", bits: int, scales: torch.Tensor, outs: "TensorTable") -> None:
    """dequantize_int8_per_tensor / dequantize_int4_per_tensor_packed for every tensor of `codes` in ONE launch."""
    n = len(codes)
    if len(outs) != n or scales.numel() < n:
        raise RuntimeError("outs / scales do not match the code table")
    with torch.cuda.device(codes.device):
        L.check(L.lib().kvq_dequant_batched(codes.ptrs.data_ptr(), int(bits), n, outs.numel, scales.data_ptr(),
                                            L.dtype_code(scales.dtype), outs.ptrs.data_ptr(), L.dtype_code(outs.dtype),
                                            L.stream_ptr(codes.device)), "kvq_dequant_batched")
        L.count_launch(1)


class _TokenStore:
    """Codes of one K or V stream: [B,H,cap,W] (W = D for INT8, ceil(D/2) for INT4) plus one scale per
    token in the input dtype (what the reference returns) and the same value widened to fp32 (what the
    fused attention kernel reads).  The live tokens are rows [start, start + T) of dim 2, contiguous; capacity
    doubles; a sliding window only advances `start` (no copy) until the tail of the buffer is reached."""

    def __init__(self, bits: int):
        self.bits = bits
        self.codes: Optional[torch.Tensor] = None
        self.scales: Optional[torch.Tensor] = None      # [cap] input dtype
        self.scales_f32: Optional[torch.Tensor] = None  # [cap] fp32
        self.start = 0
        self.T = 0
        self.D = 0
        self.n_scale_entries = 0  # how many scale tensors the reference would hold (one per append)
        self._bits_scratch: Optional[torch.Tensor] = None
        self._stride3 = None      # ctypes {sB,sH,sT} of `codes`, rebuilt when the buffer is reallocated
        self._stride3_of = None

    # ---- views of the live region
    def live_codes(self) -> torch.Tensor:
        return self.codes[:, :, self.start: self.start + self.T]

    def live_scales(self) -> torch.Tensor:
        return self.scales[self.start: self.start + self.T]

    def live_scales_f32(self) -> torch.Tensor:
        return self.scales_f32[self.start: self.start + self.T]

    def _reserve(self, x: torch.Tensor, n_new: int) -> None:
        B, H, _, D = x.shape
        W = D if self.bits == 8 else (D + 1) // 2
        need = self.T + n_new
        if self.codes is None:
            cap = max(256, 1 << (need - 1).bit_length())
            self.D = D
            self.codes = torch.empty((B, H, cap, W), dtype=torch.int8 if self.bits == 8 else torch.uint8, device=x.device)
            self.scales = torch.empty(cap, dtype=x.dtype, device=x.device)
            self.scales_f32 = torch.empty(cap, dtype=torch.float32, device=x.device)
            return
        if tuple(self.codes.shape[:2]) != (B, H) or self.D != D or self.scales.dtype != x.dtype:
            raise RuntimeError("appended tensor does not match the cache: "
                               f"expected [B={self.codes.size(0)},H={self.codes.size(1)},n,D={self.D}] {self.scales.dtype}")
        cap = self.codes.size(2)
        if self.start + need <= cap:
            return
        # out of room at the tail: move the live tokens to the front of a buffer that is large enough
        # (a window close to the capacity would otherwise re-copy the live store every cap - need steps: grow so that a copy
        # is amortised over at least `need` appends)
        if 2 * need <= cap:
            # start + need > cap and need <= cap / 2  =>  start > T: the live rows move to the front of the SAME buffers
            # (source and destination cannot overlap), no allocation
            self.codes[:, :, : self.T].copy_(self.live_codes())
            self.scales[: self.T].copy_(self.live_scales())
            self.scales_f32[: self.T].copy_(self.live_scales_f32())
            self.start = 0
            return
        new_cap = max(cap, 2 << (need - 1).bit_length())
        codes = torch.empty((B, H, new_cap, W), dtype=self.codes.dtype, device=x.device)
        codes[:, :, : self.T].copy_(self.live_codes())
        scales = torch.empty(new_cap, dtype=self.scales.dtype, device=x.device)
        scales[: self.T].copy_(self.live_scales())
        s32 = torch.empty(new_cap, dtype=torch.float32, device=x.device)
        s32[: self.T].copy_(self.live_scales_f32())
        self.codes, self.scales, self.scales_f32, self.start = codes, scales, s32, 0

    def append_per_token(self, x: torch.Tensor, eps: float = 1e-8, group=None) -> None:
        """n tokens, one scale each, in one launch (two + an allreduce when head-sharded)."""
        n = x.size(2)
        self._reserve(x, n)
        lo = self.start + self.T
        sh, st = L.shape_stride(x)
        lib, dt, stream = L.lib(), L.dtype_code(x.dtype), L.stream_ptr(x.device)
        if group is None:
            # the decode loop comes through here twice per layer per step: destination pointers by arithmetic on the
            # buffers (rows [lo, lo + n) of every plane), no tensor views
            codes = self.codes
            if self._stride3_of is not codes:
                self._stride3, self._stride3_of = L.stride3(codes), codes
            bits = self._bits_scratch
            if bits is None or bits.numel() < n or bits.device != x.device:  # absmax words: scratch, reused across steps
                bits = self._bits_scratch = torch.empty(max(n, 64), dtype=torch.int32, device=x.device)
            L.check(lib.kvq_quant_per_token_fused(x.data_ptr(), dt, sh, st, self.bits, float(eps),
                                                  codes.data_ptr() + lo * codes.stride(2), self._stride3,
                                                  self.scales.data_ptr() + lo * self.scales.element_size(),
                                                  self.scales_f32.data_ptr() + 4 * lo, bits.data_ptr(), stream),
                    "kvq_quant_per_token_fused")
            L.count_launch()
        else:
            dst = self.codes[:, :, lo: lo + n]
            sc = self.scales[lo: lo + n]
            s32 = self.scales_f32[lo: lo + n]
            bits = _absmax_bits(x, L.PER_TOKEN, group)
            fn = lib.kvq_quant_int8 if self.bits == 8 else lib.kvq_quant_int4_pack
            L.check(fn(x.data_ptr(), dt, sh, st, L.PER_TOKEN | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), float(eps), dst.data_ptr(),
                       L.stride3(dst), sc.data_ptr(), s32.data_ptr(), stream), "kvq_quant (per token)")
            L.count_launch()
        self.T += n
        self.n_scale_entries += n

    @staticmethod
    def append_per_token_batched(stores: "List[_TokenStore]", xs: List[torch.Tensor], eps: float = 1e-8) -> bool:
        """append_per_token for n stores of the same bit width in ONE launch (kvq_quant_per_token_fused_batched).  Returns
        False -- nothing done -- when the stores / inputs do not share a layout; the caller then loops."""
        if len(stores) < 2 or not L.same_layout(xs):
            return False
        x0 = xs[0]
        n = x0.size(2)
        if x0.stride(3) != 1 or (x0.size(3) * x0.element_size()) % 16 or any((s * x0.element_size()) % 16 for s in x0.stride()[:3]):
            return False
        for st_, x in zip(stores, xs):
            st_._reserve(x, n)
        s0 = stores[0]
        lo, cs, bits = s0.start + s0.T, s0.codes.stride(), s0.bits
        for st_ in stores:
            if st_.bits != bits or st_.start + st_.T != lo or st_.codes.stride() != cs or st_.scales.dtype != x0.dtype:
                return False
        sh, st4 = L.shape_stride(x0)
        tabs = L.ptr_tables(xs, [s_.codes for s_ in stores], [s_.scales for s_ in stores], [s_.scales_f32 for s_ in stores])
        L.check(L.lib().kvq_quant_per_token_fused_batched(
            tabs[0].data_ptr(), L.dtype_code(x0.dtype), len(xs), sh, st4, bits, float(eps), tabs[1].data_ptr(), L.stride3(s0.codes),
            lo, tabs[2].data_ptr(), tabs[3].data_ptr(), L.stream_ptr(x0.device)), "kvq_quant_per_token_fused_batched")
        L.count_launch()
        for st_ in stores:
            st_.T += n
            st_.n_scale_entries += n
        return True

    def append_per_token_exchange(self, x: torch.Tensor, exchange, eps: float = 1e-8) -> bool:
        """Head-sharded append_per_token in ONE launch, the max over ranks taken INSIDE the kernel through peer-mapped memory
        (kvq_quant_per_token_fused_exchange) -- instead of absmax kernel + NCCL allreduce + quantize kernel.  Returns False --
        nothing appended -- when the layout / size does not allow it; the caller then takes the allreduce path.  The step number
        lives on the device (seq = 0), so a decode step that contains this call can be captured into a CUDA graph and replayed.
        Every rank must make the same calls in the same order."""
        n = x.size(2)
        if exchange is None or float(eps) != 1e-8 or n > 1024 or x.stride(3) != 1 or (x.size(3) * x.element_size()) % 16:
            return False
        self._reserve(x, n)
        lo = self.start + self.T
        sh, st = L.shape_stride(x)
        codes = self.codes
        if self._stride3_of is not codes:
            self._stride3, self._stride3_of = L.stride3(codes), codes
        rc = L.lib().kvq_quant_per_token_fused_exchange(
            x.data_ptr(), L.dtype_code(x.dtype), sh, st, self.bits, float(eps), codes.data_ptr() + lo * codes.stride(2), self._stride3,
            self.scales.data_ptr() + lo * self.scales.element_size(), self.scales_f32.data_ptr() + 4 * lo, None,
            exchange.peer_table_ptr, exchange.rank, exchange.world, 0, L.stream_ptr(x.device))   # 0: device-resident step counter (capturable)
        if rc == L.KVQ_ERR_UNSUPPORTED:   # layout (every rank sees the same one)
            return False
        L.check(rc, "kvq_quant_per_token_fused_exchange")
        L.count_launch()
        self.T += n
        self.n_scale_entries += n
        return True

    def append_shared_scale(self, x: torch.Tensor, eps: float = 1e-8, group=None) -> None:
        """The reference's append() given n > 1 tokens: ONE scale over the whole [B,H,n,D] block
        (ops.py:181-210 calls the per-tensor helper on whatever it is handed)."""
        n = x.size(2)
        self._reserve(x, n)
        lo = self.start + self.T
        dst = self.codes[:, :, lo: lo + n]
        bits = _absmax_bits(x, L.PER_TENSOR, group)
        one = torch.empty(1, dtype=x.dtype, device=x.device)
        one32 = torch.empty(1, dtype=torch.float32, device=x.device)
        sh, st = L.shape_stride(x)
        fn = L.lib().kvq_quant_int8 if self.bits == 8 else L.lib().kvq_quant_int4_pack
        L.check(fn(x.data_ptr(), L.dtype_code(x.dtype), sh, st, L.PER_TENSOR | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), float(eps),
                   dst.data_ptr(), L.stride3(dst), one.data_ptr(), one32.data_ptr(), L.stream_ptr(x.device)), "kvq_quant (block)")
        L.count_launch()
        self.scales[lo: lo + n] = one
        self.scales_f32[lo: lo + n] = one32
        self.T += n
        self.n_scale_entries += 1

    def dequantize(self, out_dtype: torch.dtype) -> torch.Tensor:
        B, H = self.codes.shape[:2]
        out = torch.empty((B, H, self.T, self.D), dtype=out_dtype, device=self.codes.device)
        src = self.live_codes()
        scales = self.live_scales()
        sh = L.I64x4(B, H, self.T, self.D)
        fn = L.lib().kvq_dequant_int8 if self.bits == 8 else L.lib().kvq_dequant_int4_unpack
        L.check(fn(src.data_ptr(), sh, L.stride3(src), scales.data_ptr(), L.dtype_code(scales.dtype), L.PER_TOKEN,
                   out.data_ptr(), L.dtype_code(out_dtype), L.stream_ptr(out.device)), "kvq_dequant (per token)")
        L.count_launch()
        return out

    def keep_last(self, n: int) -> None:
        """Sliding window on the store: drop all but the last n tokens by moving `start` (no data movement)."""
        if n < self.T:
            self.start += self.T - n
            self.T = n
            self.n_scale_entries = n

    def evict(self, plan, table, kept_idx: torch.Tensor) -> None:
        """Compact the store to the tokens kept by an eviction plan: codes through the same index-computing gather
        kernel as fp caches (1-byte elements), scales by the device-computed index list.  With one scale per token this
        equals quantizing the trimmed fp cache, bit for bit."""
        from ..cache.implementations import gather_tokens
        n = int(plan.n_total)
        new_codes = gather_tokens(self.live_codes(), plan, table)
        cap = max(256, 1 << max(n - 1, 0).bit_length())
        codes = torch.empty((new_codes.size(0), new_codes.size(1), cap, new_codes.size(3)), dtype=self.codes.dtype,
                            device=self.codes.device)
        codes[:, :, :n].copy_(new_codes)
        scales = torch.empty(cap, dtype=self.scales.dtype, device=self.scales.device)
        s32 = torch.empty(cap, dtype=torch.float32, device=self.scales.device)
        scales[:n] = self.live_scales().index_select(0, kept_idx)
        s32[:n] = self.live_scales_f32().index_select(0, kept_idx)
        self.codes, self.scales, self.scales_f32 = codes, scales, s32
        self.start, self.T, self.n_scale_entries = 0, n, n

    def nbytes(self) -> int:
        if self.codes is None:
            return 0
        B, H, _, W = self.codes.shape
        return B * H * self.T * W * self.codes.element_size() + self.n_scale_entries * self.scales.element_size()

    def token_views(self) -> List[torch.Tensor]:
        return [] if self.codes is None else [self.codes[:, :, self.start + t: self.start + t + 1] for t in range(self.T)]

    def scale_views(self) -> List[torch.Tensor]:
        return [] if self.scales is None else [self.scales[self.start + t] for t in range(self.T)]


class QuantizedLayerKV:
    """Quantized KV cache of one transformer layer.

    mode "int8": K and V INT8; "int4": both packed INT4; "mixed": K INT8, V INT4.
    (reference ops.py:136-290)
    """

    def __init__(self, mode: str = "int8", device: str = "cuda", compute_dtype: torch.dtype = torch.float16,
                 group=None):
        assert mode in ["int8", "int4", "mixed"], f"Invalid mode: {mode}"
        self.mode = mode
        self.device = device
        self.compute_dtype = compute_dtype
        self.group = group  # torch.distributed group of a head-sharded cache (None: single GPU)
        self._k = _TokenStore(4 if mode == "int4" else 8)
        self._v = _TokenStore(8 if mode == "int8" else 4)
        self._attn_ws: Optional[torch.Tensor] = None  # split-K workspace of attend()

    # --- reference-visible storage, as views of the contiguous store ----------------------------
    @property
    def k_store(self) -> List[torch.Tensor]:
        return self._k.token_views()

    @property
    def v_store(self) -> List[torch.Tensor]:
        return self._v.token_views()

    @property
    def k_scales(self) -> List[torch.Tensor]:
        return self._k.scale_views()

    @property
    def v_scales(self) -> List[torch.Tensor]:
        return self._v.scale_views()

    @property
    def k_meta(self) -> List[Optional[int]]:
        return [self._k.D if self._k.bits == 4 else None] * self._k.T

    @property
    def v_meta(self) -> List[Optional[int]]:
        return [self._v.D if self._v.bits == 4 else None] * self._v.T

    def __len__(self) -> int:
        return self._k.T

    @staticmethod
    def _check_pair(k: torch.Tensor, v: torch.Tensor) -> None:
        L.require_cuda(k, v)
        if k.dim() != 4 or v.dim() != 4:
            raise RuntimeError("expected K and V of shape [B, H, n, D]")
        if k.numel() == 0 or v.numel() == 0:
            raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")

    @torch.no_grad()
    def append(self, k_1tok: torch.Tensor, v_1tok: torch.Tensor) -> None:
        """Append one token's K and V, each [B,H,1,D], quantized with its own scale."""
        self._check_pair(k_1tok, v_1tok)
        with L.device_guard(k_1tok):
            if k_1tok.size(2) == 1 and v_1tok.size(2) == 1:
                self._append_per_token_pair(k_1tok, v_1tok)
                return
            for store, x in ((self._k, k_1tok), (self._v, v_1tok)):
                if x.size(2) == 1:
                    store.append_per_token(x, group=self.group)
                else:
                    store.append_shared_scale(x, group=self.group)

    @torch.no_grad()
    def append_tokens(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Append n tokens [B,H,n,D], each with its own scale: identical to n append() calls of one
        token each (what init_from_prompt_past does, ops.py:339-342), in one launch per tensor."""
        self._check_pair(k, v)
        with L.device_guard(k):
            self._append_per_token_pair(k, v)

    def _append_per_token_pair(self, k: torch.Tensor, v: torch.Tensor) -> None:
        if self.group is not None:
            # head-sharded: the max over ranks inside the quantize kernel (one launch per tensor); the NCCL allreduce form when
            # peer-mapped memory is not available or the step is being captured
            from ..distributed import shared_exchange
            ex = shared_exchange(k.device, self.group)
            if ex is not None:
                for store, x in ((self._k, k), (self._v, v)):
                    if not store.append_per_token_exchange(x, ex):
                        store.append_per_token(x, group=self.group)
                return
        self._k.append_per_token(k, group=self.group)
        self._v.append_per_token(v, group=self.group)

    @torch.no_grad()
    def get_kv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Dequantized (K, V), each [B,H,T,D] in compute_dtype."""
        if self._k.T == 0:
            raise ValueError("Empty cache")
        with L.device_guard(self._k.codes):
            return self._k.dequantize(self.compute_dtype), self._v.dequantize(self.compute_dtype)

    _STRATEGIES = {"sliding_window": L.SLIDING, "prefix_window": L.PREFIX_WINDOW, "strided": L.STRIDED, "block": L.BLOCK,
                   "budget": L.BUDGET}

    @torch.no_grad()
    def evict(self, strategy: str, window_size: int = 256, prefix_len: int = 0, stride: int = 4, block_size: int = 64,
              keep_per_block: int = 8, old_budget: int = 64) -> int:
        """Apply one of the reference's eviction strategies (src/cache/implementations.py:124-292, same parameter names
        as the trim_kv_* functions) directly to the quantized store -- the reference only offers quantization and
        eviction as separate methods (benchmarker.py:673-686).  Returns the new length.  Composition is exact: with one
        scale per token, evicting codes equals quantizing the evicted fp cache."""
        import ctypes
        from ..cache.implementations import _old_table, evict_plan
        if strategy not in self._STRATEGIES:
            raise ValueError(f"unknown eviction strategy {strategy!r}; one of {sorted(self._STRATEGIES)}")
        T = self._k.T
        if T == 0:
            raise ValueError("Empty cache")
        plan = evict_plan(self._STRATEGIES[strategy], T, window=window_size, prefix=prefix_len, stride=stride, block=block_size,
                          keep=keep_per_block, budget=old_budget)
        if plan.identity:
            return T
        if strategy == "sliding_window":  # a view in the reference (:136-138); here: move the start of the live region
            self._k.keep_last(int(plan.n_total))
            self._v.keep_last(int(plan.n_total))
            return self._k.T
        dev = self._k.codes.device
        with torch.cuda.device(dev):
            table = _old_table(plan, dev)
            kept = torch.empty(int(plan.n_total), dtype=torch.int64, device=dev)
            L.check(L.lib().kvq_evict_indices_device(ctypes.byref(plan), table.data_ptr() if table is not None else None,
                                                     kept.data_ptr(), L.stream_ptr(dev)), "kvq_evict_indices_device")
            L.count_launch()
            self._k.evict(plan, table, kept)
            self._v.evict(plan, table, kept)
        return self._k.T

    def length_after_evict(self, strategy: str, window_size: int = 256, prefix_len: int = 0, stride: int = 4, block_size: int = 64,
                           keep_per_block: int = 8, old_budget: int = 64) -> int:
        """len(self) after evict(...) with the same arguments, without touching the store (host arithmetic)."""
        from ..cache.implementations import evict_plan
        if strategy not in self._STRATEGIES:
            raise ValueError(f"unknown eviction strategy {strategy!r}; one of {sorted(self._STRATEGIES)}")
        if self._k.T == 0:
            return 0
        return int(evict_plan(self._STRATEGIES[strategy], self._k.T, window=window_size, prefix=prefix_len, stride=stride,
                              block=block_size, keep=keep_per_block, budget=old_budget).n_total)

    def quantized(self):
        """(k_codes, k_scales_f32, k_bits, v_codes, v_scales_f32, v_bits) views of the live store, for
        the fused dequant decode attention."""
        if self._k.T == 0:
            raise ValueError("Empty cache")
        return (self._k.live_codes(), self._k.live_scales_f32(), self._k.bits,
                self._v.live_codes(), self._v.live_scales_f32(), self._v.bits)

    @torch.no_grad()
    def attend(self, q: torch.Tensor, seq_len: Optional[int] = None, k_cur: Optional[torch.Tensor] = None,
               v_cur: Optional[torch.Tensor] = None, sm_scale: Optional[float] = None,
               out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Fused dequant decode attention of q [B,Hq,1,D] over the first seq_len stored tokens (default: all), plus the step's
        own token k_cur / v_cur [B,H,1,D] exactly.  Same kernel call as decode_attention_quantized(*self.quantized(), ...)
        without building the views: the per-step path of KvqFusedCache."""
        import ctypes
        k, v = self._k, self._v
        T = k.T if seq_len is None else int(seq_len)
        if not 1 <= T <= k.T:
            raise ValueError("Empty cache" if k.T == 0 else f"seq_len must be in [1, {k.T}]")
        L.require_cuda(q)
        if q.dim() != 4 or q.size(2) != 1 or q.size(0) != k.codes.size(0) or q.size(3) != k.D:
            raise RuntimeError("q must be [B, Hq, 1, D] matching the cache")
        B, Hq, _, D = q.shape
        H = k.codes.size(1)
        if Hq % H == 0 and Hq // H not in (1, 2, 4, 8):
            from ..attention import run_grouped
            return run_grouped(q, H, (1, 2, 4, 8), lambda qs: self.attend(qs, T, k_cur, v_cur, float(D) ** -0.5 if sm_scale is None else sm_scale,
                                                                         out_dtype))
        if (k_cur is None) != (v_cur is None):
            raise RuntimeError("k_cur and v_cur must be given together")
        if k_cur is not None:
            if tuple(k_cur.shape) != (B, H, 1, D) or tuple(v_cur.shape) != (B, H, 1, D):
                raise RuntimeError("k_cur / v_cur must be [B, H, 1, D]")
            k_cur, v_cur = k_cur.to(q.dtype).contiguous(), v_cur.to(q.dtype).contiguous()
        qc = q.contiguous()
        out_dtype = out_dtype or q.dtype
        out = torch.empty((B, Hq, 1, D), dtype=out_dtype, device=q.device)
        for st in (k, v):
            if st._stride3_of is not st.codes:
                st._stride3, st._stride3_of = L.stride3(st.codes), st.codes
        with L.device_guard(q):
            lib = L.lib()
            ns = ctypes.c_int(1)
            nbytes = lib.kvq_decode_attn_workspace_bytes(B, Hq, H, T, D, ctypes.byref(ns))
            ws = self._attn_ws
            if ws is None or ws.numel() * 4 < nbytes or ws.device != q.device:
                ws = self._attn_ws = torch.empty(max(nbytes, 1 << 20) // 4, dtype=torch.float32, device=q.device)
            L.check(lib.kvq_decode_attn_ex(qc.data_ptr(), L.dtype_code(q.dtype),
                                           k.codes.data_ptr() + k.start * k.codes.stride(2), k.bits, k._stride3,
                                           k.scales_f32.data_ptr() + 4 * k.start,
                                           v.codes.data_ptr() + v.start * v.codes.stride(2), v.bits, v._stride3,
                                           v.scales_f32.data_ptr() + 4 * v.start, B, Hq, H, T, D,
                                           float(D) ** -0.5 if sm_scale is None else float(sm_scale),
                                           k_cur.data_ptr() if k_cur is not None else None,
                                           v_cur.data_ptr() if v_cur is not None else None, out.data_ptr(),
                                           L.dtype_code(out_dtype), ws.data_ptr(), ns.value, L.stream_ptr(q.device)),
                    "kvq_decode_attn")
            L.count_launch(2)
        return out

    def estimated_bytes(self) -> int:
        """Stored codes + scales, as the reference counts them (ops.py:271-290)."""
        return self._k.nbytes() + self._v.nbytes()


class QuantizedKVCache:
    """Quantized caches of all layers (reference ops.py:293-363)."""

    def __init__(self, n_layers: int, mode: str = "int8", device: str = "cuda",
                 compute_dtype: torch.dtype = torch.float16, group=None):
        self.layers = [QuantizedLayerKV(mode=mode, device=device, compute_dtype=compute_dtype, group=group)
                       for _ in range(n_layers)]

    def _append_all(self, ks: List[torch.Tensor], vs: List[torch.Tensor]) -> None:
        """Per-token append of every layer's K and V: one launch per bit width for ALL layers when they share a layout
        (the usual case), else one launch per tensor."""
        layers = self.layers[: len(ks)]
        if len(ks) > 1 and all(l.group is None for l in layers) and all(k.is_cuda and k.dim() == 4 and k.numel() for k in ks + vs):
            with L.device_guard(ks[0]):
                kst, vst = [l._k for l in layers], [l._v for l in layers]
                if kst[0].bits == vst[0].bits and ks[0].shape == vs[0].shape:
                    if _TokenStore.append_per_token_batched(kst + vst, ks + vs):
                        return
                else:
                    done_k = _TokenStore.append_per_token_batched(kst, ks)
                    done_v = _TokenStore.append_per_token_batched(vst, vs)
                    if done_k and done_v:
                        return
                    for l, k, v in zip(layers, ks, vs):
                        if not done_k:
                            l._check_pair(k, v)
                            l._k.append_per_token(k, group=None)
                        if not done_v:
                            l._check_pair(k, v)
                            l._v.append_per_token(v, group=None)
                    return
        for l, k, v in zip(layers, ks, vs):
            l.append_tokens(k, v)

    @torch.no_grad()
    def append_from_past(self, past_key_values: tuple) -> None:
        """Append the last token of every layer's (k, v)."""
        pkv = list(past_key_values)
        self._append_all([k[:, :, -1:, :] for k, _ in pkv], [v[:, :, -1:, :] for _, v in pkv])

    @torch.no_grad()
    def init_from_prompt_past(self, past_key_values: tuple) -> None:
        """Quantize the whole prompt, one scale per token (the reference loops over t, ops.py:339-342)."""
        pkv = [(k, v) for k, v in past_key_values]
        if all(k.size(2) > 0 for k, _ in pkv):
            self._append_all([k for k, _ in pkv], [v for _, v in pkv])
            return
        for layer_idx, (k, v) in enumerate(pkv):
            if k.size(2) > 0:
                self.layers[layer_idx].append_tokens(k, v)

    @torch.no_grad()
    def to_past_key_values(self) -> tuple:
        return tuple(layer.get_kv() for layer in self.layers)

    def estimated_bytes(self) -> int:
        return sum(layer.estimated_bytes() for layer in self.layers)
