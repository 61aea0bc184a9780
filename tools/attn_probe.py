"""One launch of the fused decode attention per mode on the C5 per-GPU shard (ncu target)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200
dev = torch.device("cuda")
B, H, T, D = 8, 1, 65536, 128
G = int(os.environ.get("G", "1"))
for mode, kb, vb in (("mixed", 8, 4), ("int8", 8, 8), ("int4", 4, 4)):
    kc = torch.randint(-127, 128, (B, H, T, D if kb == 8 else D // 2), device=dev, dtype=torch.int8)
    vc = torch.randint(0, 255, (B, H, T, D if vb == 8 else D // 2), device=dev, dtype=torch.uint8)
    if kb == 4: kc = kc.view(torch.uint8)
    if vb == 8: vc = vc.view(torch.int8)
    ks = torch.rand(T, device=dev) * 0.01 + 0.01
    vs = torch.rand(T, device=dev) * 0.01 + 0.01
    q = torch.randn(B, H * G, 1, D, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        out = kvq_b200.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
