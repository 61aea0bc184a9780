"""KVCacheBenchmarker: greedy decoding of a HuggingFace causal LM under twelve KV-cache policies, with
the cache work done by the B200 kernels of this package.

Same public surface as the reference (/root/reference/src/benchmarking/benchmarker.py:23-832): every
generate_* method, its arguments and return tuple, and benchmark_method with its result-dict keys
(:811-832, including the quirks: `block_size` is reported for paged_attention only, and
`est_kv_cache_mb_avg` is NaN for the trim policies).  The twelve near-identical loops of the reference
are one loop here (`_greedy`) parameterised by what happens to the cache between steps.

transformers >= 5 removed DynamicCache.from_legacy_cache / to_legacy_cache (the reference pins
4.57.3); `_as_cache` / `_as_tuple` cover both.
"""

from __future__ import annotations

import math
import time
from typing import Callable, Optional, Tuple

import torch
from transformers import DynamicCache

from ..cache import (
    PagedKVCache,
    chunk_summarize_kv,
    trim_kv_block_old,
    trim_kv_budget_old,
    trim_kv_prefix_window,
    trim_kv_sliding_window,
    trim_kv_strided,
)
from ..core.utils import get_cpu_mem_mb, get_gpu_peak_mb, mb, reset_gpu_peak
from ..quantization import QuantizedKVCache

LegacyKV = Tuple[Tuple[torch.Tensor, torch.Tensor], ...]


def _as_tuple(past) -> LegacyKV:
    """Model cache object -> ((k, v), ...) with tensors [B,H,T,D]."""
    if isinstance(past, tuple):
        return past
    if hasattr(past, "to_legacy_cache"):
        return past.to_legacy_cache()
    return tuple((layer.keys, layer.values) for layer in past.layers)


def _as_cache(kv: LegacyKV):
    """((k, v), ...) -> the cache object the model accepts."""
    if hasattr(DynamicCache, "from_legacy_cache"):
        return DynamicCache.from_legacy_cache(kv)
    return DynamicCache(ddp_cache_data=kv)


class KVCacheBenchmarker:
    """Benchmarks KV-cache strategies: none, full, sliding window, prefix+window, strided, block-sparse,
    budget-sparse, quantized (int8 / int4 / mixed), paged, chunk-summary."""

    VALID_METHODS = ("no_cache", "full_cache", "sliding_window", "quant_int8", "quant_int4", "quant_mixed",
                     "paged_attention", "chunked_cache", "prefix_window", "strided_cache", "block_cache",
                     "budget_cache")

    def __init__(self, model, tokenizer, device: str = "cuda"):
        self.model = model
        self.tokenizer = tokenizer
        self.device = device

    # ------------------------------------------------------------------ shared pieces

    def _encode(self, prompt: str, truncate: bool) -> torch.Tensor:
        if truncate:  # reference :77, :117, :223, :271, :321, :378
            enc = self.tokenizer(prompt, return_tensors="pt", truncation=True, max_length=1024)
        else:         # reference :172, :438, :510, :592
            enc = self.tokenizer(prompt, return_tensors="pt")
        return enc.input_ids.to(self.device)

    def _finish(self, input_ids: torch.Tensor, generated: torch.Tensor) -> Tuple[str, int]:
        n_new = generated.shape[-1] - input_ids.shape[-1]
        return self.tokenizer.decode(generated[0], skip_special_tokens=True), n_new

    @torch.no_grad()
    def _greedy(self, input_ids: torch.Tensor, max_new_tokens: int,
                after_forward: Callable[[LegacyKV], object], clamp_vocab: bool = False):
        """Prompt forward, then max_new_tokens single-token steps.  `after_forward(kv_tuple)` receives the
        model's cache after every forward and returns what to pass as past_key_values next (a cache
        object, or a zero-argument callable building it, evaluated only if another forward follows),
        which is where each policy trims / quantizes / pages / summarises."""
        out = self.model(input_ids=input_ids, use_cache=True)
        past = after_forward(_as_tuple(out.past_key_values))
        logits = out.logits[:, -1, :]
        generated = input_ids.clone()
        vocab = self.model.config.vocab_size
        for _ in range(max_new_tokens):
            next_token = torch.argmax(logits, dim=-1, keepdim=True)
            if clamp_vocab:
                next_token = torch.clamp(next_token, 0, vocab - 1)
            generated = torch.cat([generated, next_token], dim=-1)
            if callable(past):
                past = past()
            out = self.model(input_ids=next_token, use_cache=True, past_key_values=past)
            logits = out.logits[:, -1, :]
            past = after_forward(_as_tuple(out.past_key_values))
        return generated, past

    def _trimmed(self, prompt, max_new_tokens, truncate, trim):
        ids = self._encode(prompt, truncate)
        generated, _ = self._greedy(ids, max_new_tokens, lambda kv: _as_cache(trim(kv)))
        return self._finish(ids, generated)

    # ------------------------------------------------------------------ generation methods

    @torch.no_grad()
    def generate_no_cache(self, prompt: str, max_new_tokens: int = 32) -> Tuple[str, int]:
        """No KV cache: the whole sequence is re-fed every step; stops at EOS."""
        ids = self._encode(prompt, True)
        generated = ids.clone()
        vocab = self.model.config.vocab_size
        for _ in range(max_new_tokens):
            logits = self.model(input_ids=generated, use_cache=False).logits[:, -1, :]
            next_token = torch.clamp(torch.argmax(logits, dim=-1, keepdim=True), 0, vocab - 1)
            generated = torch.cat([generated, next_token], dim=-1)
            if next_token.item() == self.tokenizer.eos_token_id:
                break
        return self._finish(ids, generated)

    @torch.no_grad()
    def generate_with_cache(self, prompt: str, max_new_tokens: int = 32) -> Tuple[str, int]:
        """Standard full KV cache."""
        ids = self._encode(prompt, True)
        generated, _ = self._greedy(ids, max_new_tokens, _as_cache, clamp_vocab=True)
        return self._finish(ids, generated)

    @torch.no_grad()
    def generate_with_sliding_window(self, prompt: str, max_new_tokens: int = 32, window_size: int = 256
                                     ) -> Tuple[str, int]:
        """Keep only the last window_size positions after every step."""
        return self._trimmed(prompt, max_new_tokens, False, lambda kv: trim_kv_sliding_window(kv, window_size))

    @torch.no_grad()
    def generate_with_prefix_window(self, prompt: str, max_new_tokens: int = 32, window_size: int = 256,
                                    prefix_len: int = 32) -> Tuple[str, int]:
        """Keep prefix + tail window."""
        return self._trimmed(prompt, max_new_tokens, True,
                             lambda kv: trim_kv_prefix_window(kv, prefix_len=prefix_len, window_size=window_size))

    @torch.no_grad()
    def generate_with_strided_cache(self, prompt: str, max_new_tokens: int = 32, window_size: int = 256,
                                    stride: int = 4, prefix_len: int = 0) -> Tuple[str, int]:
        """Keep tail window + strided old tokens (+ optional prefix)."""
        return self._trimmed(prompt, max_new_tokens, True,
                             lambda kv: trim_kv_strided(kv, window_size=window_size, stride=stride, prefix_len=prefix_len))

    @torch.no_grad()
    def generate_with_block_cache(self, prompt: str, max_new_tokens: int = 32, window_size: int = 256,
                                  block_size: int = 64, keep_per_block: int = 8, prefix_len: int = 0) -> Tuple[str, int]:
        """Keep tail window + block-sparse older tokens."""
        return self._trimmed(prompt, max_new_tokens, True,
                             lambda kv: trim_kv_block_old(kv, window_size=window_size, block_size=block_size,
                                                          keep_per_block=keep_per_block, prefix_len=prefix_len))

    @torch.no_grad()
    def generate_with_budget_cache(self, prompt: str, max_new_tokens: int = 32, window_size: int = 256,
                                   old_budget: int = 64, prefix_len: int = 0) -> Tuple[str, int]:
        """Keep tail window + fixed-budget sampled older tokens (+ optional prefix)."""
        return self._trimmed(prompt, max_new_tokens, True,
                             lambda kv: trim_kv_budget_old(kv, window_size=window_size, old_budget=old_budget,
                                                           prefix_len=prefix_len))

    @torch.no_grad()
    def generate_with_quantized_kv(self, prompt: str, max_new_tokens: int = 32, mode: str = "int8",
                                   fused_attention: bool = False, evict: Optional[dict] = None,
                                   paged_block_size: Optional[int] = None) -> Tuple[str, int, float]:
        """Decode with the KV cache stored quantized (int8 / int4 / mixed).  Returns (text, n_new, cache MB).

        Prompt: one fused absmax+quantize launch per layer tensor (one scale per token).  Each step: one
        dequantize launch per layer tensor feeds the model, then the new token is appended.

        fused_attention=True (not in the reference): the model reads the quantized store directly through
        `KvqFusedCache` and the fused dequant decode-attention kernel instead of dequantizing the whole cache into
        a DynamicCache every step (kvq_b200/hf_cache.py); same cache contents, same per-token scales.
        evict (with fused_attention): a QuantizedLayerKV.evict() keyword dict, e.g. {"strategy": "sliding_window",
        "window_size": 256}: quantization composed with an eviction strategy in one cache (BASELINE config C2).
        paged_block_size (with fused_attention): keep the codes in a paged pool read through its block table by the
        attention kernel (quantization composed with paging; a sliding window then drops whole blocks)."""
        ids = self._encode(prompt, False)
        if fused_attention:
            return self._generate_fused(ids, max_new_tokens, mode, evict, paged_block_size)
        if evict is not None or paged_block_size is not None:
            raise ValueError("a composed quantize + evict / paged cache needs fused_attention=True (KvqFusedCache)")
        compute_dtype = torch.float16 if str(self.device).startswith("cuda") else torch.float32
        state = {}

        def after_forward(kv: LegacyKV):
            qcache = state.get("q")
            if qcache is None:
                qcache = state["q"] = QuantizedKVCache(n_layers=len(kv), mode=mode, device=self.device,
                                                       compute_dtype=compute_dtype)
                qcache.init_from_prompt_past(kv)
            else:
                qcache.append_from_past(kv)
            return lambda: _as_cache(qcache.to_past_key_values())

        generated, _ = self._greedy(ids, max_new_tokens, after_forward)
        text, n_new = self._finish(ids, generated)
        return text, n_new, mb(state["q"].estimated_bytes())

    @torch.no_grad()
    def _generate_fused(self, ids: torch.Tensor, max_new_tokens: int, mode: str, evict: Optional[dict] = None,
                        paged_block_size: Optional[int] = None) -> Tuple[str, int, float]:
        from ..hf_cache import KvqFusedCache, fused_attention
        cfg = self.model.config
        n_layers = getattr(cfg, "num_hidden_layers", None) or cfg.n_layer
        dtype = next(self.model.parameters()).dtype
        cache = KvqFusedCache(n_layers, mode=mode, compute_dtype=dtype, evict=evict, paged_block_size=paged_block_size)
        with fused_attention(self.model):
            out = self.model(input_ids=ids, use_cache=True, past_key_values=cache)
            logits = out.logits[:, -1, :]
            generated = ids.clone()
            for _ in range(max_new_tokens):
                next_token = torch.argmax(logits, dim=-1, keepdim=True)
                generated = torch.cat([generated, next_token], dim=-1)
                out = self.model(input_ids=next_token, use_cache=True, past_key_values=cache)
                logits = out.logits[:, -1, :]
        text, n_new = self._finish(ids, generated)
        return text, n_new, mb(cache.estimated_bytes())

    @torch.no_grad()
    def generate_with_paged_attention(self, prompt: str, max_new_tokens: int = 32, block_size: int = 64
                                      ) -> Tuple[str, int, float, float, int]:
        """KV stored in fixed-size blocks behind a block table, stitched for the model every step.
        Returns (text, n_new, alloc_mb, used_mb, num_blocks)."""
        ids = self._encode(prompt, False)
        dtype = torch.float16 if str(self.device).startswith("cuda") else torch.float32
        state = {}

        def after_forward(kv: LegacyKV):
            pages = state.get("p")
            if pages is None:
                pages = state["p"] = [PagedKVCache(block_size=block_size, device=self.device, dtype=dtype) for _ in kv]
                for cache, (k, v) in zip(pages, kv):
                    cache.append_tokens(k, v)  # the reference appends the prompt token by token (:528-530): same result
            else:
                for cache, (k, v) in zip(pages, kv):
                    cache.append(k[:, :, -1:, :], v[:, :, -1:, :])
            return lambda: _as_cache(tuple(c.get_kv() for c in pages))

        generated, _ = self._greedy(ids, max_new_tokens, after_forward)
        text, n_new = self._finish(ids, generated)
        pages = state["p"]
        alloc = sum(c.allocated_bytes() for c in pages)
        used = sum(c.used_bytes() for c in pages)
        return text, n_new, mb(alloc), mb(used), sum(c.num_blocks() for c in pages)

    @torch.no_grad()
    def generate_with_chunked_cache(self, prompt: str, max_new_tokens: int = 32, chunk_size: int = 64,
                                    keep_last: int = 256) -> Tuple[str, int, float]:
        """Older KV replaced by mean-pooled chunk summaries, recent tokens exact.  Returns (text, n_new, MB)."""
        ids = self._encode(prompt, False)
        generated, past = self._greedy(
            ids, max_new_tokens, lambda kv: _as_cache(chunk_summarize_kv(kv, chunk_size=chunk_size, keep_last=keep_last)))
        text, n_new = self._finish(ids, generated)
        est_bytes = sum(k.numel() * k.element_size() + v.numel() * v.element_size() for k, v in _as_tuple(past))
        return text, n_new, est_bytes / (1024 ** 2)

    # ------------------------------------------------------------------ benchmarking

    def benchmark_method(self, prompts: list, method: str = "no_cache", max_new_tokens: int = 32, window_size: int = 256,
                         block_size: int = 64, chunk_size: int = 64, keep_last: int = 256, mode: str = "int8",
                         prefix_len: int = 32, stride: int = 4, keep_per_block: int = 8, old_budget: int = 64,
                         fused_attention: bool = False, evict: Optional[dict] = None,
                         paged_block_size: Optional[int] = None) -> dict:
        """Run `method` over `prompts` (sequentially) and report time, tokens/s and memory.
        `fused_attention` (extension, default off) applies to the quant_* methods, see generate_with_quantized_kv."""
        assert method in self.VALID_METHODS, f"Invalid method: {method}"
        nan = float("nan")
        runners = {
            "no_cache": lambda p: (self.generate_no_cache(p, max_new_tokens)[1], 0.0),
            "full_cache": lambda p: (self.generate_with_cache(p, max_new_tokens)[1], nan),
            "sliding_window": lambda p: (self.generate_with_sliding_window(p, max_new_tokens, window_size=window_size)[1], nan),
            "quant_int8": lambda p: self.generate_with_quantized_kv(p, max_new_tokens, mode="int8", fused_attention=fused_attention, evict=evict, paged_block_size=paged_block_size)[1:3],
            "quant_int4": lambda p: self.generate_with_quantized_kv(p, max_new_tokens, mode="int4", fused_attention=fused_attention, evict=evict, paged_block_size=paged_block_size)[1:3],
            "quant_mixed": lambda p: self.generate_with_quantized_kv(p, max_new_tokens, mode="mixed", fused_attention=fused_attention, evict=evict, paged_block_size=paged_block_size)[1:3],
            "paged_attention": lambda p: self.generate_with_paged_attention(p, max_new_tokens, block_size=block_size)[1:3],
            "chunked_cache": lambda p: self.generate_with_chunked_cache(p, max_new_tokens, chunk_size=chunk_size,
                                                                        keep_last=keep_last)[1:3],
            "prefix_window": lambda p: (self.generate_with_prefix_window(p, max_new_tokens=max_new_tokens,
                                                                         window_size=window_size, prefix_len=prefix_len)[1], nan),
            "strided_cache": lambda p: (self.generate_with_strided_cache(p, max_new_tokens=max_new_tokens, window_size=window_size,
                                                                         stride=stride, prefix_len=prefix_len)[1], nan),
            "block_cache": lambda p: (self.generate_with_block_cache(p, max_new_tokens=max_new_tokens, window_size=window_size,
                                                                     block_size=block_size, keep_per_block=keep_per_block,
                                                                     prefix_len=prefix_len)[1], nan),
            "budget_cache": lambda p: (self.generate_with_budget_cache(p, max_new_tokens=max_new_tokens, window_size=window_size,
                                                                       old_budget=old_budget, prefix_len=prefix_len)[1], nan),
        }
        run = runners[method]
        on_gpu = str(self.device).startswith("cuda")

        reset_gpu_peak(self.device)
        start_cpu = get_cpu_mem_mb()
        if on_gpu:  # CUDA events on the current stream, like the reference (:693-697)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        else:
            t0 = time.time()

        total_new_tokens, est_mbs = 0, []
        for prompt in prompts:
            n_new, est = run(prompt)
            total_new_tokens += n_new
            est_mbs.append(est)

        if on_gpu:
            ev1.record()
            torch.cuda.synchronize()
            elapsed = ev0.elapsed_time(ev1) / 1000.0
        else:
            elapsed = time.time() - t0

        finite = [x for x in est_mbs if isinstance(x, float) and not math.isnan(x)]
        windowed = ("sliding_window", "prefix_window", "strided_cache", "block_cache", "budget_cache")
        return {
            "method": method,
            "elapsed_sec": elapsed,
            "total_new_tokens": total_new_tokens,
            "tokens_per_sec": total_new_tokens / elapsed if elapsed > 0 else float("inf"),
            "cpu_mem_used_mb": get_cpu_mem_mb() - start_cpu,
            "gpu_peak_mb": get_gpu_peak_mb(self.device),
            "window_size": window_size if method in windowed else None,
            "block_size": block_size if method == "paged_attention" else None,
            "chunk_size": chunk_size if method == "chunked_cache" else None,
            "est_kv_cache_mb_avg": sum(finite) / len(finite) if finite else nan,
            "prefix_len": prefix_len if method in windowed[1:] else None,
            "stride": stride if method == "strided_cache" else None,
            "keep_per_block": keep_per_block if method == "block_cache" else None,
            "old_budget": old_budget if method == "budget_cache" else None,
        }
