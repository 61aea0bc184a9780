"""Per-kernel HBM throughput of libkvq_sm100.so on the BASELINE shapes (CUDA events on the launching
stream, inputs rotated through buffers larger than L2).  Development tool; bench.py is the contract."""

import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200  # noqa: E402
from kvq_b200 import _lib as L  # noqa: E402
from kvq_b200.cache import evict_plan  # noqa: E402
from kvq_b200.cache.implementations import gather_tokens, _chunk_mean  # noqa: E402

PEAK = 6544.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
    iters = int(os.environ.get('MB_ITERS', iters))
    warm = int(os.environ.get('MB_WARM', warm))
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def timeit_graph(fn, n_variants, reps=10):
    """GPU time per call with the CPU out of the way: capture `reps` rounds over all variants in one CUDA
    graph (the C ABI launches on torch's current stream, so they are captured) and replay it."""
    for i in range(n_variants):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for i in range(n_variants):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (reps * n_variants))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report_graph(name, nbytes, fn, n_variants):
    med, best = timeit_graph(fn, n_variants)
    gbs = nbytes / med / 1e6
    print(f"{name:44s} {med:8.4f} ms  {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f}% of measured {PEAK:.0f}  {100 * gbs / 8000:5.1f}% of 8 TB/s  (best {nbytes / best / 1e6:.0f}) [graph]", flush=True)
    return gbs


def report(name, nbytes, fn, **kw):
    med, best = timeit(fn, **kw)
    gbs = nbytes / med / 1e6
    print(f"{name:44s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f}% of measured {PEAK:.0f}  {100 * gbs / 8000:5.1f}% of 8 TB/s  (best {nbytes / best / 1e6:.0f})", flush=True)
    return gbs


def main():
    dev = torch.device("cuda")
    lib = L.lib()
    st = L.stream_ptr(dev)
    which = sys.argv[1:] or ["copy", "quant", "token", "evict", "chunk", "attn"]
    B, H, T, D = 16, int(os.environ.get("MB_H", "8")), 32768, 128  # MB_H=1: the per-GPU head shard at N = 8 (128 MB tensors)
    N = B * H * T * D
    nbuf = 4
    if "copy" in which:
        a = [torch.empty(N, dtype=torch.bfloat16, device=dev).normal_() for _ in range(2)]
        b = torch.empty(N, dtype=torch.bfloat16, device=dev)
        report("torch copy_ bf16 1Gi elems (r+w)", 4 * N, lambda i: b.copy_(a[i % 2]))
        del a, b
    if "quant" in which or "token" in which:
        xs = [torch.randn(B, H, T, D, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    if "quant" in which:
        bits = torch.zeros(1, dtype=torch.int32, device=dev)
        sh, sd = L.shape_stride(xs[0])
        report("absmax per-tensor bf16 (N*2)", 2 * N, lambda i: lib.kvq_absmax(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 0, bits.data_ptr(), 0, st))
        q8 = torch.empty(B, H, T, D, dtype=torch.int8, device=dev)
        p4 = torch.empty(B, H, T, D // 2, dtype=torch.uint8, device=dev)
        sc = torch.empty(1, dtype=torch.bfloat16, device=dev)
        report("quant int8 phase2 bf16 (N*3)", 3 * N, lambda i: lib.kvq_quant_int8(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 0, bits.data_ptr(), 1e-8, q8.data_ptr(), None, sc.data_ptr(), None, st))
        report("quant int4+pack phase2 bf16 (N*2.5)", int(2.5 * N), lambda i: lib.kvq_quant_int4_pack(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 0, bits.data_ptr(), 1e-8, p4.data_ptr(), None, sc.data_ptr(), None, st))
        outs = [torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        q8s = [q8, q8.clone()]
        p4s = [p4, p4.clone(), p4.clone(), p4.clone()]
        report("dequant int8 -> bf16 (N*3)", 3 * N, lambda i: lib.kvq_dequant_int8(q8s[i % 2].data_ptr(), sh, None, sc.data_ptr(), L.BF16, 0, outs[i % 2].data_ptr(), L.BF16, st))
        report("dequant int4 unpack -> bf16 (N*2.5)", int(2.5 * N), lambda i: lib.kvq_dequant_int4_unpack(p4s[i % 4].data_ptr(), sh, None, sc.data_ptr(), L.BF16, 0, outs[i % 2].data_ptr(), L.BF16, st))
        report("api quantize_int4_per_tensor_packed (N*4.5)", int(4.5 * N), lambda i: kvq_b200.quantize_int4_per_tensor_packed(xs[i % nbuf]), iters=8)
        report("api dequantize_int4 (N*2.5)", int(2.5 * N), lambda i: kvq_b200.dequantize_int4_per_tensor_packed(p4s[i % 4], sc, D, torch.bfloat16), iters=8)
        # fp32 / fp16 inputs (half the tokens to keep memory equal)
        xf = [torch.randn(B, H, T // 2, D, device=dev, dtype=torch.float32) for _ in range(2)]
        shf, sdf = L.shape_stride(xf[0])
        report("quant int8 phase2 fp32 in (N/2*5)", 5 * N // 2, lambda i: lib.kvq_quant_int8(xf[i % 2].data_ptr(), L.F32, shf, sdf, 0, bits.data_ptr(), 1e-8, q8.data_ptr(), None, None, None, st))
        del xf, outs, q8s, p4s
    if "token" in which:
        q8 = torch.empty(B, H, T, D, dtype=torch.int8, device=dev)
        p4 = torch.empty(B, H, T, D // 2, dtype=torch.uint8, device=dev)
        scT = torch.empty(T, dtype=torch.bfloat16, device=dev)
        s32 = torch.empty(T, dtype=torch.float32, device=dev)
        bitsT = torch.empty(T, dtype=torch.int32, device=dev)
        sh, sd = L.shape_stride(xs[0])
        report("per-token fused absmax+quant int8 (N*3)", 3 * N, lambda i: lib.kvq_quant_per_token_fused(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 8, 1e-8, q8.data_ptr(), None, scT.data_ptr(), s32.data_ptr(), bitsT.data_ptr(), st))
        report("per-token fused absmax+quant int4 (N*2.5)", int(2.5 * N), lambda i: lib.kvq_quant_per_token_fused(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 4, 1e-8, p4.data_ptr(), None, scT.data_ptr(), s32.data_ptr(), bitsT.data_ptr(), st))
        report("per-token absmax only (N*2)", 2 * N, lambda i: lib.kvq_absmax(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 1, bitsT.data_ptr(), 0, st))
        report("per-token quant int8 phase2 (N*3)", 3 * N, lambda i: lib.kvq_quant_int8(xs[i % nbuf].data_ptr(), L.BF16, sh, sd, 1, bitsT.data_ptr(), 1e-8, q8.data_ptr(), None, scT.data_ptr(), s32.data_ptr(), st))
        outs = [torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        report("per-token dequant int8 -> bf16 (N*3)", 3 * N, lambda i: lib.kvq_dequant_int8(q8.data_ptr(), sh, None, scT.data_ptr(), L.BF16, 1, outs[i % 2].data_ptr(), L.BF16, st))
        report("per-token dequant int4 -> bf16 (N*2.5)", int(2.5 * N), lambda i: lib.kvq_dequant_int4_unpack(p4.data_ptr(), sh, None, scT.data_ptr(), L.BF16, 1, outs[i % 2].data_ptr(), L.BF16, st))
        # C2 shape: GPT-2 small fp16 B16 H12 T1024 D64, 12 layers worth of K tensors (25 MB each)
        c2 = [torch.randn(16, 12, 1024, 64, device=dev, dtype=torch.float16) for _ in range(24)]
        c2q = torch.empty(16, 12, 1024, 64, dtype=torch.int8, device=dev)
        sh2, sd2 = L.shape_stride(c2[0])
        n2 = c2[0].numel()
        report_graph("C2 per-token fused int8 fp16 [16,12,1024,64]", 3 * n2, lambda i: lib.kvq_quant_per_token_fused(c2[i % 24].data_ptr(), L.F16, sh2, sd2, 8, 1e-8, c2q.data_ptr(), None, scT.data_ptr(), s32.data_ptr(), bitsT.data_ptr(), L.stream_ptr(dev)), 24)
        del outs, c2
    if "quant" in which or "token" in which:
        del xs
    torch.cuda.empty_cache()
    if "evict" in which or "chunk" in which:
        # C4: Llama-3-8B shape, T=131072; B=16 to be larger than L2 per tensor (4 GiB), rotate 2 buffers
        Bc, Hc, Tc, Dc = 4, 8, 131072, 128
        ks = [torch.randn(Bc, Hc, Tc, Dc, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        R = Bc * Hc * Dc * 2
    if "evict" in which:
        for name, stg, kw in (("strided w256 s4", L.STRIDED, dict(window=256, stride=4)), ("block 64/8 w256", L.BLOCK, dict(window=256, block=64, keep=8)),
                              ("budget 64 w256", L.BUDGET, dict(window=256, budget=64)), ("prefix32+window256", L.PREFIX_WINDOW, dict(window=256, prefix=32)),
                              ("strided s2 (half)", L.STRIDED, dict(window=256, stride=2))):
            plan = evict_plan(stg, Tc, **kw)
            out = torch.empty(Bc, Hc, int(plan.n_total), Dc, dtype=torch.bfloat16, device=dev)
            shc, sdc = L.shape_stride(ks[0])
            report(f"evict {name} kept={plan.n_total} (2*kept*R)", 2 * int(plan.n_total) * R,
                   lambda i: lib.kvq_evict_gather(ks[i % 2].data_ptr(), 2, shc, sdc, ctypes.byref(plan), None, out.data_ptr(), None, st))
    if "chunk" in which:
        old = Tc - 256
        nb = old * R + (-(-old // 64)) * R + 2 * 256 * R
        report("chunk mean c64 keep256 T=131072", nb, lambda i: _chunk_mean(ks[i % 2], 64, 256))
    if "evict" in which or "chunk" in which:
        del ks
    torch.cuda.empty_cache()
    if "attn" in which:
        # C5 per-GPU shard at 8 GPUs: 1 KV head, B=8, T=65536, D=128; several layers rotate
        Ba, Ha, Ta, Da = 8, 1, 65536, 128
        for G in [int(x) for x in os.environ.get("MB_ATTN_G", "1,8").split(",")]:
            for mode, kb, vb in (("mixed", 8, 4), ("int8", 8, 8), ("int4", 4, 4)):
                nl = 6
                kcs = [torch.randint(-127, 128, (Ba, Ha, Ta, Da if kb == 8 else Da // 2), device=dev, dtype=torch.int8) for _ in range(nl)]
                vcs = [torch.randint(0, 255, (Ba, Ha, Ta, Da if vb == 8 else Da // 2), device=dev, dtype=torch.uint8) for _ in range(nl)]
                if kb == 4:
                    kcs = [k.view(torch.uint8) for k in kcs]
                if vb == 8:
                    vcs = [v.view(torch.int8) for v in vcs]
                ksc = torch.rand(Ta, device=dev) * 0.01 + 0.01
                vsc = torch.rand(Ta, device=dev) * 0.01 + 0.01
                q = torch.randn(Ba, Ha * G, 1, Da, device=dev, dtype=torch.bfloat16)
                nbytes = Ba * Ha * Ta * (Da * kb // 8 + Da * vb // 8) + 8 * Ta
                ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)
                g = report_graph(f"decode attn {mode} G={G} B8 H1 T65536 D128", nbytes,
                                 lambda i: kvq_b200.decode_attention_quantized(q, kcs[i % nl], ksc, kb, vcs[i % nl], vsc, vb, workspace=ws), nl)
                del kcs, vcs
    if "paged" in which:
        # same C5 shard through a shuffled block table (INT8-K / INT4-V): cost of the indirection per block size
        Ba, Ha, Ta, Da = 8, 1, 65536, 128
        for G in [int(x) for x in os.environ.get("MB_ATTN_G", "1,8").split(",")]:
            for bs in (16, 64, 128):
                nb = Ta // bs
                nl = 4
                kps = [torch.randint(-127, 128, (nb, Ba, Ha, bs, Da), device=dev, dtype=torch.int8) for _ in range(nl)]
                vps = [torch.randint(0, 255, (nb, Ba, Ha, bs, Da // 2), device=dev, dtype=torch.uint8) for _ in range(nl)]
                ksp = torch.rand(nb, bs, device=dev) * 0.01 + 0.01
                vsp = torch.rand(nb, bs, device=dev) * 0.01 + 0.01
                table = torch.randperm(nb, device=dev).to(torch.int32)
                q = torch.randn(Ba, Ha * G, 1, Da, device=dev, dtype=torch.bfloat16)
                nbytes = Ba * Ha * Ta * (Da + Da // 2) + 8 * Ta
                ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)
                report_graph(f"paged decode attn mixed G={G} block {bs} B8 H1 T65536 D128", nbytes,
                             lambda i: kvq_b200.decode_attention_paged(q, kps[i % nl], ksp, 8, vps[i % nl], vsp, 4, table, Ta, workspace=ws), nl)
                del kps, vps
    print("launches", L.launch_count)


if __name__ == "__main__":
    main()
