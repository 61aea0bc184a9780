"""One launch pair per (variant, G) at the C5 shape for ncu captures: python tools/attn_profile.py [variant] [G ...]
   ATTN_H=1 (KV heads: the per-rank shard of an 8-GPU run), ATTN_B, ATTN_T override the shape."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200
variant = sys.argv[1] if len(sys.argv) > 1 else "imma"
Gs = [int(x) for x in sys.argv[2:]] or [1, 8]
B, H, T, D = int(os.environ.get('ATTN_B', 8)), int(os.environ.get('ATTN_H', 8)), int(os.environ.get('ATTN_T', 65536)), 128
g = torch.Generator(device="cuda").manual_seed(1)
kc = torch.randint(-127, 128, (B, H, T, D), device="cuda", dtype=torch.int8, generator=g)
vc = torch.randint(0, 256, (B, H, T, D // 2), device="cuda", dtype=torch.uint8, generator=g)
ks = torch.rand(T, device="cuda", generator=g) * 0.01 + 0.01
vs = torch.rand(T, device="cuda", generator=g) * 0.01 + 0.01
ws = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for G in Gs:
    q = torch.randn(B, H * G, 1, D, device="cuda", dtype=torch.bfloat16, generator=g)
    for _ in range(3):
        flush.zero_()
        kvq_b200.decode_attention_quantized(q, kc, ks, 8, vc, vs, 4, workspace=ws, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0.record()
        kvq_b200.decode_attention_quantized(q, kc, ks, 8, vc, vs, 4, workspace=ws, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nb = B * H * T * (D + D // 2) + 8 * T
    print(f"{variant} G={G}: {min(ts) * 1e3:.1f} us best, {sum(ts) / len(ts) * 1e3:.1f} us mean -> {nb / min(ts) / 1e6:.0f} GB/s")
