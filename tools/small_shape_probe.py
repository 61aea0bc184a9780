"""Per-call time of the store operations at GPT-2 C1 / C2 shapes (launch-bound regime): which call is slow?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200

def t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

for (B, H, T, D) in ((1, 12, 100, 64), (16, 12, 1000, 64)):
    k = torch.randn(B, H, T, D, device="cuda", dtype=torch.float16)
    k1 = torch.randn(B, H, 1, D, device="cuda", dtype=torch.float16)
    for mode in ("int8", "int4", "mixed"):
        s = kvq_b200.QuantizedLayerKV(mode, "cuda", torch.float16)
        s.append_tokens(k, k)
        print(f"[{B},{H},{T},{D}] {mode:5s} get_kv {t(s.get_kv):7.1f} us   append(1 tok, K+V) {t(lambda: (s.append_tokens(k1, k1), s._k.keep_last(T), s._v.keep_last(T))):7.1f} us")
