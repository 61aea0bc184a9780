"""Per-token fused quantizer at the C3 shape (n tensors in one launch) for timing / ncu: python tools/pertoken_profile.py [bits] [n]"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
from kvq_b200 import _lib as L
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, H, T, D = 16, int(sys.argv[3]) if len(sys.argv) > 3 else 8, 32768, 128
xs = [torch.randn(B, H, T, D, device="cuda", dtype=torch.bfloat16) for _ in range(n)]
W = D if bits == 8 else D // 2
qs = [torch.empty(B, H, T, W, device="cuda", dtype=torch.uint8) for _ in range(n)]
sc = [torch.empty(T, device="cuda", dtype=torch.bfloat16) for _ in range(n)]
s32 = [torch.empty(T, device="cuda", dtype=torch.float32) for _ in range(n)]
tabs = L.ptr_tables(xs, qs, sc, s32)
sh, st = L.shape_stride(xs[0])
lib, stream = L.lib(), L.stream_ptr(xs[0].device)
def run():
    L.check(lib.kvq_quant_per_token_fused_batched(tabs[0].data_ptr(), L.BF16, n, sh, st, bits, 1e-8, tabs[1].data_ptr(), None, 0,
                                                  tabs[2].data_ptr(), tabs[3].data_ptr(), stream), "q")
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nb = n * B * H * T * D * (2 + bits / 8)
print(f"per-token fused INT{bits}, {n} tensors: {ms:.3f} ms -> {nb / ms / 1e6:.0f} GB/s ({nb / ms / 1e6 / 8000:.3f} of 8 TB/s)")
