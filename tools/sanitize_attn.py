"""Small decode-attention cases for compute-sanitizer (memcheck / racecheck / synccheck): every kernel variant once.
   compute-sanitizer --tool racecheck python tools/sanitize_attn.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200  # noqa: E402

torch.manual_seed(0)
for (B, Hq, H, T, D, bs) in ((1, 2, 2, 700, 64, 16), (2, 8, 2, 515, 128, 64), (1, 8, 1, 1300, 128, 128), (1, 4, 2, 300, 64, 32)):
    for mode in ("int8", "mixed", "int4"):
        k, v = torch.randn(B, H, T, D, device="cuda"), torch.randn(B, H, T, D, device="cuda")
        q = torch.randn(B, Hq, 1, D, device="cuda")
        flat = kvq_b200.QuantizedLayerKV(mode, "cuda", torch.float32)
        flat.append_tokens(k, v)
        kc, ks, kb, vc, vs, vb = flat.quantized()
        a = kvq_b200.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb, k_cur=k[:, :, :1], v_cur=v[:, :, :1])
        if Hq // H != 2:
            paged = kvq_b200.PagedQuantizedLayerKV(mode, block_size=bs, compute_dtype=torch.float32)
            paged.append_tokens(k, v)
            b = paged.attend(q, k_cur=k[:, :, :1], v_cur=v[:, :, :1])
            assert torch.equal(a, b)
torch.cuda.synchronize()
print("sanitize_attn ok")
