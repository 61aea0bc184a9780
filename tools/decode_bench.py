"""tok/s of KVCacheBenchmarker.benchmark_method on a random-init GPT-2 small (BASELINE C1 / C2 shapes).

    python tools/decode_bench.py [--device cuda|cpu] [--impl kvq|reference] [--batch 1] [--prompt-len 64] [--new 64] [--methods ...]

--impl reference imports the unmodified reference from /root/reference (build container only, CPU) with
the two-method DynamicCache shim of SURVEY 8c; --impl kvq runs this package on the GPU."""

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StubTokenizer:
    eos_token_id = 50256

    def __init__(self, batch, length):
        g = torch.Generator().manual_seed(0)
        self.ids = torch.randint(0, 50257, (batch, length), generator=g)

    def __call__(self, prompt, return_tensors="pt", truncation=False, max_length=None):
        ids = self.ids
        if truncation and max_length:
            ids = ids[:, :max_length]
        return type("Enc", (), {"input_ids": ids})()

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--impl", default="kvq")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt-len", type=int, default=64)
    ap.add_argument("--new", type=int, default=64)
    ap.add_argument("--window", type=int, default=256)
    ap.add_argument("--evict-window", type=int, default=0, help="with --fused: compose the quantized store with a sliding window")
    ap.add_argument("--paged-block", type=int, default=0, help="with --fused: keep the codes in a paged pool of this block size")
    ap.add_argument("--fused", action="store_true", help="quant_* methods through KvqFusedCache + fused decode attention")
    ap.add_argument("--methods", nargs="*", default=["full_cache", "quant_int8", "quant_int4", "quant_mixed", "sliding_window",
                                                     "strided_cache", "block_cache", "budget_cache", "chunked_cache", "paged_attention"])
    args = ap.parse_args()
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config()).eval()
    if args.device == "cuda":
        model = model.half().cuda()
    if args.impl == "reference":
        sys.dont_write_bytecode = True
        sys.path.insert(0, "/root/reference")
        from transformers import DynamicCache
        if not hasattr(DynamicCache, "from_legacy_cache"):
            DynamicCache.to_legacy_cache = lambda self: tuple((l.keys, l.values) for l in self.layers)
            DynamicCache.from_legacy_cache = classmethod(lambda cls, t: cls(ddp_cache_data=t))
        from src.benchmarking.benchmarker import KVCacheBenchmarker
    else:
        sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
        from kvq_b200.benchmarking import KVCacheBenchmarker
    bench = KVCacheBenchmarker(model, StubTokenizer(args.batch, args.prompt_len), device=args.device)
    for method in args.methods:
        kw = {"fused_attention": True} if (args.fused and args.impl == "kvq") else {}
        if kw and args.evict_window and method.startswith("quant_"):
            kw["evict"] = {"strategy": "sliding_window", "window_size": args.evict_window}
        if kw and args.paged_block and method.startswith("quant_"):
            kw["paged_block_size"] = args.paged_block
        if args.device == "cuda":
            bench.benchmark_method(["x"], method=method, max_new_tokens=4, window_size=args.window, **kw)  # warm
        t0 = time.time()
        r = bench.benchmark_method(["x"], method=method, max_new_tokens=args.new, window_size=args.window, **kw)
        r["fused_attention"] = bool(kw)
        r["evict"] = kw.get("evict")
        r["paged_block_size"] = kw.get("paged_block_size")
        r["wall_sec"] = time.time() - t0
        r["batch"], r["prompt_len"], r["impl"], r["device"] = args.batch, args.prompt_len, args.impl, args.device
        r["tokens_per_sec_x_batch"] = r["tokens_per_sec"] * args.batch
        print(json.dumps({k: (None if isinstance(v, float) and v != v else v) for k, v in r.items()}), flush=True)


if __name__ == "__main__":
    main()
