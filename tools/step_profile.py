"""Where does a decode step go?  GPT-2 small (random init), C2 shape by default: wall time per step of the HF DynamicCache
loop vs the fused quantized-store loop, and a cProfile of the latter (host-side overhead of the cache adapter)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
from transformers import DynamicCache, GPT2Config, GPT2LMHeadModel  # noqa: E402

from kvq_b200.hf_cache import KvqFusedCache, fused_attention  # noqa: E402

B, T0, NEW = int(os.environ.get("SP_B", "16")), int(os.environ.get("SP_T", "960")), 64
torch.manual_seed(0)
model = GPT2LMHeadModel(GPT2Config()).half().cuda().eval()
ids = torch.randint(0, 50257, (B, T0), device="cuda")


@torch.no_grad()
def loop(cache, ctx):
    with ctx:
        out = model(input_ids=ids, use_cache=True, past_key_values=cache)
        tok = out.logits[:, -1:].argmax(-1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(NEW):
            out = model(input_ids=tok, use_cache=True, past_key_values=cache)
            tok = out.logits[:, -1:].argmax(-1)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / NEW * 1e3


import contextlib  # noqa: E402

for rep in range(2):
    print("HF DynamicCache          ms/step", round(loop(DynamicCache(), contextlib.nullcontext()), 3))
    print("fused mixed              ms/step", round(loop(KvqFusedCache(12, "mixed"), fused_attention(model)), 3))
    print("fused mixed + window 256 ms/step", round(loop(KvqFusedCache(12, "mixed", evict={"strategy": "sliding_window", "window_size": 256}), fused_attention(model)), 3))
    print("fused mixed paged 64     ms/step", round(loop(KvqFusedCache(12, "mixed", paged_block_size=64), fused_attention(model)), 3))
pr = cProfile.Profile()
pr.enable()
loop(KvqFusedCache(12, "mixed"), fused_attention(model))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
