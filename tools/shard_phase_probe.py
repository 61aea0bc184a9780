"""Per-phase time of the C3 step on ONE GPU at the per-rank shard shapes of N = 1, 2, 4, 8 (same bytes per launch as a rank of
the N-GPU run sees; 64 tensors each): shows which phase loses efficiency when the tensors get small.
   python tools/shard_phase_probe.py [heads ...]      (launch-shape knobs are compile-time macros: -DKVQ_FLAT_MIN_CPC=... etc.)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
from kvq_b200 import _lib as L
from kvq_b200.quantization import TensorTable

lib = L.lib()
dev = torch.device("cuda")
B, T, D, NT = 16, 32768, 128, 64
for Hl in [int(a) for a in sys.argv[1:]] or [8, 4, 2, 1]:
    nt = NT if Hl <= 4 else 32  # 64 x 1 GiB inputs + outputs do not fit next to each other at Hl = 8
    xs = [torch.randn(B, Hl, T, D, device=dev, dtype=torch.bfloat16) for _ in range(nt)]
    pk = [torch.empty(B, Hl, T, D // 2, device=dev, dtype=torch.uint8) for _ in range(nt)]
    outs = [torch.empty(B, Hl, T, D, device=dev, dtype=torch.bfloat16) for _ in range(nt)]
    tx, tp, to = TensorTable(xs), TensorTable(pk), TensorTable(outs)
    bits = torch.zeros(nt, dtype=torch.int32, device=dev)
    scales = torch.empty(nt, dtype=torch.bfloat16, device=dev)
    n = xs[0].numel()
    st = L.stream_ptr(dev)
    phases = {
        "absmax": (2.0, lambda: lib.kvq_absmax_batched(tx.ptrs.data_ptr(), L.BF16, nt, n, bits.data_ptr(), st)),
        "quant ": (2.5, lambda: lib.kvq_quant_batched(tx.ptrs.data_ptr(), L.BF16, nt, n, 4 | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), 1e-8, tp.ptrs.data_ptr(), scales.data_ptr(), None, st)),
        "dequant": (2.5, lambda: lib.kvq_dequant_batched(tp.ptrs.data_ptr(), 4, nt, n, scales.data_ptr(), L.BF16, to.ptrs.data_ptr(), L.BF16, st)),
    }
    line = f"Hl={Hl} ({n * 2 >> 20} MB tensors x {nt})"
    for name, (bpe, fn) in phases.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        line += f"  {name} {ms:7.3f} ms {bpe * n * nt / ms / 1e6:7.0f} GB/s"
    print(line, flush=True)
    del xs, pk, outs, tx, tp, to
    torch.cuda.empty_cache()
