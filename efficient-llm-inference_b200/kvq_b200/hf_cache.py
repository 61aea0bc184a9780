"""A `transformers` cache whose layers are quantized stores, plus the attention function that reads them
fused ("next" row 1).

The reference's quantized decode loop dequantizes the WHOLE cache every step, wraps it in a DynamicCache and
lets the model concatenate the new token and run attention on fp tensors
(/root/reference/src/benchmarking/benchmarker.py:465-486, src/quantization/ops.py:213-269, :345-355): O(T) bytes
written and re-read per step just to feed attention.  Here the model talks to the quantized store directly:

  * `KvqFusedCache.update()` (called by every attention layer) quantizes and appends the incoming tokens with
    one scale per token -- exactly what `QuantizedKVCache.init_from_prompt_past` / `append_from_past` do -- and
    hands the layer only the tokens it was given;
  * prefill (more than one query token) attends over those exact prompt tokens with the model's stock
    attention, like the reference's first forward;
  * a decode step (one query token) calls the fused kernel: INT8/INT4 codes of the cached tokens are
    dequantized in registers, the step's own token is attended to exactly (unquantized), as in the reference.

Use `fused_attention(model)` as a context manager: it registers the attention function under the name
"kvq_fused" and switches `model.config._attn_implementation` to it for the duration.
"""

from __future__ import annotations

import contextlib
import threading
from typing import Optional, Tuple

import torch
from transformers.cache_utils import Cache, CacheLayerMixin

from .attention import decode_attention_quantized  # noqa: F401  (re-exported for users of this module)
from .cache.paged_quantized import PagedQuantizedLayerKV
from .quantization.ops import QuantizedLayerKV

ATTN_NAME = "kvq_fused"

# per host thread: layer index -> (cache layer, the key tensor its update() returned).  The attention function of the same
# module call pops the entry and checks that the `key` it was handed IS that tensor, so two models (or two caches) in
# flight can never read each other's layers.
_tls = threading.local()


def _pending() -> dict:
    d = getattr(_tls, "pending", None)
    if d is None:
        d = _tls.pending = {}
    return d


def _mask_is_trivial(mask: Optional[torch.Tensor]) -> bool:
    """True when `mask` lets the one query token see every key (None, all-True bool, all-zero additive).  The same mask
    object serves every layer of a forward: the device sync is paid once per step, not once per layer."""
    if mask is None:
        return True
    key = (mask.data_ptr(), tuple(mask.shape), mask._version)
    if getattr(_tls, "mask_key", None) != key:
        _tls.mask_key = key
        _tls.mask_ok = bool(mask.all()) if mask.dtype == torch.bool else bool((mask == 0).all())
    return _tls.mask_ok


class KvqCacheLayer(CacheLayerMixin):
    """One layer of the cache: a QuantizedLayerKV store (codes + per-token scales on the device)."""

    is_sliding = False

    def __init__(self, layer_idx: int, mode: str, compute_dtype: torch.dtype, group=None, evict: Optional[dict] = None,
                 paged_block_size: Optional[int] = None):
        super().__init__()
        self.layer_idx = layer_idx
        self.mode = mode
        self.evict = evict  # e.g. {"strategy": "sliding_window", "window_size": 256}: applied to the store after every step
        self.paged_block_size = paged_block_size
        if paged_block_size is not None:
            if evict is not None and evict.get("strategy") != "sliding_window":
                raise ValueError("the paged store evicts by dropping leading blocks: sliding_window only")
        self.store = self._new_store(compute_dtype, group)
        self.cached_before_update = 0

    def _new_store(self, compute_dtype, group):
        if self.paged_block_size is not None:
            return PagedQuantizedLayerKV(mode=self.mode, block_size=self.paged_block_size, device="cuda", compute_dtype=compute_dtype,
                                         group=group)
        return QuantizedLayerKV(mode=self.mode, device="cuda", compute_dtype=compute_dtype, group=group)

    def lazy_initialization(self, key_states: torch.Tensor, value_states: torch.Tensor) -> None:
        self.dtype, self.device = key_states.dtype, key_states.device
        self.is_initialized = True

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, *args, **kwargs
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        if not self.is_initialized:
            self.lazy_initialization(key_states, value_states)
        if self.evict is not None and len(self.store) > 0:
            # trim what earlier steps accumulated (the reference's loops trim after every step, benchmarker.py:199-205):
            # doing it here, before the append, keeps `cached_before_update` tokens contiguous for this step's attention
            if self.paged_block_size is not None:
                self.store.evict_sliding_window(self.evict.get("window_size", 256))  # whole leading blocks: a table edit
            else:
                self.store.evict(**self.evict)
        self.cached_before_update = len(self.store)
        self.store.append_tokens(key_states, value_states)  # one scale per token, one launch per tensor
        _pending()[self.layer_idx] = (self, key_states)
        return key_states, value_states                     # only the new tokens: the rest stays quantized

    def get_seq_length(self) -> int:
        """Length the NEXT forward will see.  With `evict` set the trim is applied lazily at the next update(), but it is
        reported here already, so position ids and mask sizes equal those of the reference's loops, which trim right after
        every step (benchmarker.py:199-205)."""
        n = len(self.store)
        if self.evict is None or n == 0:
            return n
        if self.paged_block_size is not None:
            return self.store.length_after_sliding_window(self.evict.get("window_size", 256))
        return self.store.length_after_evict(**self.evict)

    def get_mask_sizes(self, query_length: int) -> Tuple[int, int]:
        return self.get_seq_length() + query_length, 0

    def get_max_cache_shape(self) -> int:
        return -1

    # the store is not a pair of fp tensors: operations that assume so are not supported
    def offload(self):
        raise NotImplementedError("KvqCacheLayer keeps its codes on the GPU")

    def prefetch(self):
        return None

    def reset(self) -> None:
        self.store = self._new_store(self.store.compute_dtype, getattr(self.store, "group", None))

    def reorder_cache(self, beam_idx):
        raise NotImplementedError("beam search is not supported on the quantized store")


class KvqFusedCache(Cache):
    """`past_key_values` object backed by quantized stores (modes "int8", "int4", "mixed")."""

    def __init__(self, n_layers: int, mode: str = "int8", compute_dtype: torch.dtype = torch.float16, group=None,
                 evict: Optional[dict] = None, paged_block_size: Optional[int] = None):
        """evict (optional): a QuantizedLayerKV.evict() keyword dict, e.g. {"strategy": "sliding_window",
        "window_size": 256} -- BASELINE config C2's "quant_mixed + sliding_window 256" as ONE cache object.
        paged_block_size (optional): keep the codes in a paged pool (PagedQuantizedLayerKV) read through its block table by
        the attention kernel; a sliding window then drops whole leading blocks."""
        assert mode in ["int8", "int4", "mixed"], f"Invalid mode: {mode}"
        super().__init__(layers=[KvqCacheLayer(i, mode, compute_dtype, group, evict, paged_block_size) for i in range(n_layers)])
        self.mode = mode

    def estimated_bytes(self) -> int:
        """Same accounting as QuantizedKVCache.estimated_bytes (reference ops.py:357-363); blocks in use for a paged store."""
        return sum(layer.store.allocated_bytes() if layer.paged_block_size is not None else layer.store.estimated_bytes()
                   for layer in self.layers)

    def to_past_key_values(self) -> tuple:
        """Dequantized ((k, v), ...) like QuantizedKVCache.to_past_key_values, for inspection."""
        return tuple(layer.store.get_kv() for layer in self.layers)


def _stock_attention(module, query, key, value, attention_mask, dropout, scaling, **kwargs):
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    return ALL_ATTENTION_FUNCTIONS["sdpa"](module, query, key, value, attention_mask, dropout=dropout, scaling=scaling, **kwargs)


def kvq_fused_attention(module, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                        attention_mask: Optional[torch.Tensor], dropout: float = 0.0, scaling: Optional[float] = None,
                        **kwargs):
    """Attention function registered under "kvq_fused" (transformers AttentionInterface signature).

    query [B,Hq,q_len,D]; key / value are what KvqCacheLayer.update returned: the new tokens only.
      * no cached tokens (prefill): stock attention over the tokens at hand, like the reference's first forward;
      * one query token, every cached token visible: the fused dequant kernel (cached tokens from codes, the step's token exact);
      * anything the fused kernel does not implement -- a mask that hides cached tokens (left padding), a sliding-window or
        soft-capped attention module, more than one query token on a non-empty cache -- takes the dequantize path:
        cached tokens dequantized (QuantizedLayerKV.get_kv), concatenated with the new ones, stock attention with the mask.
    """
    ent = _pending().pop(getattr(module, "layer_idx", None), None)
    if ent is None or ent[1] is not key:
        return _stock_attention(module, query, key, value, attention_mask, dropout, scaling, **kwargs)  # not our cache
    layer = ent[0]
    T = layer.cached_before_update
    if T == 0:
        return _stock_attention(module, query, key, value, attention_mask, dropout, scaling, **kwargs)
    special = (kwargs.get("sliding_window") is not None or kwargs.get("softcap") is not None
               or getattr(module, "sliding_window", None) is not None or getattr(module, "attn_logit_softcapping", None) is not None)
    if query.size(2) == 1 and key.size(2) == 1 and not special and _mask_is_trivial(attention_mask):
        if layer.paged_block_size is not None:
            out = layer.store.attend(query, sm_scale=scaling, k_cur=key, v_cur=value, out_dtype=query.dtype, seq_len=T)
        else:
            out = layer.store.attend(query, T, key, value, scaling, query.dtype)
        return out.transpose(1, 2), None  # [B, 1, Hq, D], like the stock attention functions
    kd, vd = layer.store.get_kv()
    k_all = torch.cat([kd[:, :, :T].to(key.dtype), key], dim=2)
    v_all = torch.cat([vd[:, :, :T].to(value.dtype), value], dim=2)
    return _stock_attention(module, query, k_all, v_all, attention_mask, dropout, scaling, **kwargs)


def register() -> None:
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    try:
        from transformers import AttentionInterface
        AttentionInterface.register(ATTN_NAME, kvq_fused_attention)
    except Exception:
        ALL_ATTENTION_FUNCTIONS[ATTN_NAME] = kvq_fused_attention
    try:  # the causal-mask builder is looked up under the same name: reuse the sdpa one
        from transformers import AttentionMaskInterface
        from transformers.masking_utils import sdpa_mask
        AttentionMaskInterface.register(ATTN_NAME, sdpa_mask)
    except Exception:
        pass


@contextlib.contextmanager
def fused_attention(model):
    """Route `model`'s attention through the quantized store for the duration of the block."""
    register()
    old = model.config._attn_implementation
    model.config._attn_implementation = ATTN_NAME
    try:
        yield
    finally:
        model.config._attn_implementation = old
        _pending().clear()
