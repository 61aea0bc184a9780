"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU, torch) in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the outputs are committed as small fixtures and
this script is committed next to them.  Every array that is bf16 is stored as its uint16 bit
pattern (suffix `_bf16`), fp16 likewise (`_f16`).

Fixture files
  quant.npz      per-tensor INT8 / INT4 quantize + dequantize (ops.py:10-133), KATs of SURVEY section 4
  pertoken.npz   QuantizedLayerKV append/get_kv/estimated_bytes in int8 / int4 / mixed (ops.py:136-290)
  evict.npz      index sets of trim_kv_* (implementations.py:124-292) over a parameter sweep
  evict_big.npz  million-token budget cases with float32-linspace duplicates (delta encoded)
  chunk.npz      chunk_summarize_kv (implementations.py:295-346)
  paged.npz      PagedKVCache append/get_kv/bytes (implementations.py:10-121)
  attn.npz       fp32 eager attention over the reference's dequantized cache (benchmarker.py:470-477)
  benchmarker.npz  KVCacheBenchmarker.generate_with_quantized_kv / generate_with_cache on a seeded random-init GPT-2 small
                 (benchmarker.py:423-491): greedy tokens, n_new and est_mb for int8 / int4 / mixed (SURVEY section 4 KAT)
"""

import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from src.quantization import ops as rq  # noqa: E402
from src.cache import implementations as rc  # noqa: E402

rq.get_cuda_extension = lambda: None  # CPU container: force the torch branch explicitly

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.uint16).numpy().copy()
    return t.numpy().copy()


def gen(shape, dtype, seed, outlier=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g, dtype=torch.float32)
    if outlier:
        x.view(-1)[::97] *= 8
    return x.to(dtype)


def make_quant():
    out = {}
    cases = []
    # KATs
    kat8 = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 127.0])
    kat4 = torch.tensor([0.5, 1.5, 2.5, -2.5, -7.0, 7.0, 6.5])
    shapes = [("kat8", kat8), ("kat4", kat4)]
    for name, dt in DT.items():
        shapes.append((f"zeros_{name}", torch.zeros(4, dtype=dt)))
        shapes.append((f"rand_odd_{name}", gen((2, 3, 5, 7), dt, 11)))
        shapes.append((f"rand_d64_{name}", gen((2, 4, 9, 64), dt, 12)))
        shapes.append((f"outlier_d128_{name}", gen((1, 2, 33, 128), dt, 13, outlier=True)))
        shapes.append((f"tiny_{name}", (gen((3, 16), dt, 14) * 1e-9).to(dt)))  # eps clamp path
        shapes.append((f"one_{name}", gen((1,), dt, 15)))
    for name, x in shapes:
        cases.append(name)
        tag = {torch.float32: "f32", torch.float16: "f16", torch.bfloat16: "bf16"}[x.dtype]
        out[f"{name}/x_{tag}"] = bits(x)
        q8, s8 = rq.quantize_int8_per_tensor(x)
        out[f"{name}/q8"] = q8.numpy()
        out[f"{name}/s8_{tag}"] = bits(s8)
        p4, s4, last = rq.quantize_int4_per_tensor_packed(x)
        out[f"{name}/p4"] = p4.numpy()
        out[f"{name}/s4_{tag}"] = bits(s4)
        out[f"{name}/last"] = np.int64(last)
        for oname, odt in DT.items():
            out[f"{name}/dq8_{oname}"] = bits(rq.dequantize_int8_per_tensor(q8, s8, odt))
            out[f"{name}/dq4_{oname}"] = bits(rq.dequantize_int4_per_tensor_packed(p4, s4, last, odt))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "quant.npz"), **out)


def make_pertoken():
    out = {}
    cases = []
    for name, dt in DT.items():
        for (B, H, T, D) in [(2, 3, 6, 16), (1, 2, 5, 7), (2, 12, 4, 64)]:
            for mode in ("int8", "int4", "mixed"):
                case = f"{mode}_{name}_{B}x{H}x{T}x{D}"
                cases.append(case)
                k = gen((B, H, T, D), dt, 100 + T + D, outlier=True)
                v = gen((B, H, T, D), dt, 200 + T + D)
                cdt = dt  # compute dtype = storage dtype here; f16 input -> f16 out etc.
                layer = rq.QuantizedLayerKV(mode=mode, device="cpu", compute_dtype=cdt)
                for t in range(T):
                    layer.append(k[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
                kk, vv = layer.get_kv()
                out[f"{case}/k_{name}"] = bits(k)
                out[f"{case}/v_{name}"] = bits(v)
                out[f"{case}/kstore"] = torch.cat(layer.k_store, dim=2).numpy()
                out[f"{case}/vstore"] = torch.cat(layer.v_store, dim=2).numpy()
                out[f"{case}/kscales_{name}"] = bits(torch.stack(layer.k_scales))
                out[f"{case}/vscales_{name}"] = bits(torch.stack(layer.v_scales))
                out[f"{case}/kdq_{name}"] = bits(kk)
                out[f"{case}/vdq_{name}"] = bits(vv)
                out[f"{case}/bytes"] = np.int64(layer.estimated_bytes())
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "pertoken.npz"), **out)


def index_set(fn, T, **kw):
    x = torch.arange(T, dtype=torch.float32).view(1, 1, T, 1)
    (k, _), = fn(((x, x),), **kw)
    return k.reshape(-1).to(torch.int64).numpy()


def make_evict():
    out = {}
    rows = []  # (strategy, T, window, prefix, stride, block, keep, budget)
    Ts = [1, 2, 7, 40, 257, 300, 1000, 4097]
    for T in Ts:
        for w in (0, 1, 8, 256):
            rows.append((0, T, w, 0, 1, 1, 1, 0))
            for p in (0, 3, 32):
                rows.append((1, T, w, p, 1, 1, 1, 0))
                for s in (1, 2, 4, 7):
                    rows.append((2, T, w, p, s, 1, 1, 0))
                for (blk, keep) in ((1, 1), (10, 2), (64, 8), (64, 64), (5, 3)):
                    rows.append((3, T, w, p, 1, blk, keep, 0))
                for bud in (0, 1, 2, 5, 64, 333):
                    rows.append((4, T, w, p, 1, 1, 1, bud))
    # long-context budget/linspace stress (float32 linspace at large magnitudes, near-1 steps)
    for (T, w, p, bud) in [(131072, 256, 0, 64), (131072, 256, 0, 50000), (131072, 256, 4, 130000),
                           (65536, 0, 0, 65535), (32768, 256, 32, 32000), (100003, 17, 5, 99000),
                           (16384, 256, 0, 4096), (20000, 100, 7, 19892), (70001, 1, 0, 69999)]:
        rows.append((4, T, w, p, 1, 1, 1, bud))
    for (T, w, p, s) in [(131072, 256, 0, 4), (131072, 256, 16, 3)]:
        rows.append((2, T, w, p, s, 1, 1, 0))
    for (T, w, p, blk, keep) in [(131072, 256, 0, 64, 8), (131071, 255, 3, 100, 7)]:
        rows.append((3, T, w, p, 1, blk, keep, 0))
    flat, offs = [], [0]
    for (st, T, w, p, s, blk, keep, bud) in rows:
        if st == 0:
            idx = index_set(rc.trim_kv_sliding_window, T, window_size=w)
        elif st == 1:
            idx = index_set(rc.trim_kv_prefix_window, T, prefix_len=p, window_size=w)
        elif st == 2:
            idx = index_set(rc.trim_kv_strided, T, window_size=w, stride=s, prefix_len=p)
        elif st == 3:
            idx = index_set(rc.trim_kv_block_old, T, window_size=w, block_size=blk, keep_per_block=keep, prefix_len=p)
        else:
            idx = index_set(rc.trim_kv_budget_old, T, window_size=w, old_budget=bud, prefix_len=p)
        flat.append(idx)
        offs.append(offs[-1] + len(idx))
    out["params"] = np.array(rows, dtype=np.int64)
    out["offsets"] = np.array(offs, dtype=np.int64)
    out["indices"] = np.concatenate(flat).astype(np.int32)
    # gathered data for a few cases (values, not only indices)
    for name, dt in DT.items():
        k = gen((2, 3, 40, 8), dt, 300)
        v = gen((2, 3, 40, 8), dt, 301)
        out[f"data/k_{name}"] = bits(k)
        out[f"data/v_{name}"] = bits(v)
        (k2, v2), = rc.trim_kv_block_old(((k, v),), window_size=8, block_size=10, keep_per_block=2, prefix_len=3)
        out[f"data/block_k_{name}"] = bits(k2)
        out[f"data/block_v_{name}"] = bits(v2)
        (k3, v3), = rc.trim_kv_budget_old(((k, v),), window_size=8, old_budget=5, prefix_len=2)
        out[f"data/budget_k_{name}"] = bits(k3)
    np.savez_compressed(os.path.join(HERE, "evict.npz"), **out)


def make_evict_big():
    """Million-token budget cases where float32 linspace yields duplicates that unique_consecutive removes
    (the kernel's table mode).  Stored as first index + int8 deltas (the lists are ~1M long)."""
    out = {}
    rows = [(964858, 148, 1, 951071), (1367489, 194, 13, 1364206), (1000003, 256, 0, 999000)]
    for i, (T, w, p, bud) in enumerate(rows):
        idx = index_set(rc.trim_kv_budget_old, T, window_size=w, old_budget=bud, prefix_len=p)
        d = np.diff(idx)
        assert d.min() >= 0 and d.max() < 127
        out[f"{i}/first"] = np.int64(idx[0])
        out[f"{i}/deltas"] = d.astype(np.int8)
    out["params"] = np.array(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "evict_big.npz"), **out)


def make_chunk():
    out = {}
    cases = []
    ones = torch.ones(1, 1, 10, 2)
    (ko, _), = rc.chunk_summarize_kv(((ones, ones),), chunk_size=4, keep_last=4)
    out["kat/out_f32"] = bits(ko)
    for name, dt in DT.items():
        for (B, H, T, D, chunk, keep) in [(2, 3, 50, 8, 4, 6), (1, 2, 130, 64, 64, 2), (2, 2, 9, 16, 4, 20),
                                          (1, 1, 64, 128, 16, 0), (2, 4, 77, 32, 5, 13)]:
            case = f"{name}_{B}x{H}x{T}x{D}_c{chunk}_k{keep}"
            cases.append(case)
            k = gen((B, H, T, D), dt, 400 + T)
            (ko, _), = rc.chunk_summarize_kv(((k, k),), chunk_size=chunk, keep_last=keep)
            out[f"{case}/x_{name}"] = bits(k)
            out[f"{case}/out_{name}"] = bits(ko)
            out[f"{case}/params"] = np.array([chunk, keep], dtype=np.int64)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "chunk.npz"), **out)


def make_paged():
    out = {}
    for name, dt in DT.items():
        B, H, T, D, blk = 2, 3, 11, 8, 4
        k = gen((B, H, T, D), dt, 500)
        v = gen((B, H, T, D), dt, 501)
        pc = rc.PagedKVCache(block_size=blk, device="cpu", dtype=dt)
        for t in range(T):
            pc.append(k[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
        kk, vv = pc.get_kv()
        out[f"{name}/k"] = bits(k)
        out[f"{name}/v"] = bits(v)
        out[f"{name}/kk"] = bits(kk)
        out[f"{name}/vv"] = bits(vv)
        out[f"{name}/meta"] = np.array([blk, pc.num_blocks(), pc.allocated_bytes(), pc.used_bytes(), pc.t_filled], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "paged.npz"), **out)


def make_attn():
    """fp32 eager attention (transformers eager_attention_forward formula) over the reference's
    dequantized per-token cache.  Hq = G*H query heads read kv head hq // G (repeat_kv)."""
    out = {}
    cases = []
    for (B, Hq, H, T, D, mode) in [(2, 4, 4, 37, 16, "int8"), (2, 8, 2, 50, 64, "mixed"), (1, 6, 3, 21, 8, "int4"),
                                   (2, 12, 12, 33, 64, "mixed")]:
        case = f"{mode}_{B}x{Hq}x{H}x{T}x{D}"
        cases.append(case)
        k = gen((B, H, T, D), torch.float32, 600 + T, outlier=True)
        v = gen((B, H, T, D), torch.float32, 601 + T)
        q = gen((B, Hq, 1, D), torch.float32, 602 + T)
        layer = rq.QuantizedLayerKV(mode=mode, device="cpu", compute_dtype=torch.float32)
        for t in range(T):
            layer.append(k[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
        kd, vd = layer.get_kv()
        G = Hq // H
        kr = kd.repeat_interleave(G, dim=1)
        vr = vd.repeat_interleave(G, dim=1)
        w = torch.softmax(torch.matmul(q, kr.transpose(-1, -2)) * (D ** -0.5), dim=-1)
        o = torch.matmul(w, vr)
        out[f"{case}/q"] = bits(q)
        out[f"{case}/kstore"] = torch.cat(layer.k_store, dim=2).numpy()
        out[f"{case}/vstore"] = torch.cat(layer.v_store, dim=2).numpy()
        out[f"{case}/kscales"] = bits(torch.stack(layer.k_scales))
        out[f"{case}/vscales"] = bits(torch.stack(layer.v_scales))
        out[f"{case}/out"] = bits(o)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "attn.npz"), **out)


class StubTokenizer:
    """The only tokenizer members the benchmarker touches (benchmarker.py:76-78,95,99); decode() returns the ids."""

    eos_token_id = 50256

    def __init__(self, ids):
        self.ids = ids

    def __call__(self, prompt, return_tensors="pt", truncation=False, max_length=None):
        return type("Enc", (), {"input_ids": self.ids})()

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)


def make_benchmarker():
    """The reference's quantized decode loop itself (unmodified; transformers >= 5 needs the two-method DynamicCache shim of
    SURVEY 8c) on GPT2LMHeadModel(GPT2Config()) under torch.manual_seed(0), fp32, CPU: 16 prompt ids, 8 new tokens."""
    from transformers import DynamicCache, GPT2Config, GPT2LMHeadModel
    if not hasattr(DynamicCache, "from_legacy_cache"):
        DynamicCache.to_legacy_cache = lambda self: tuple((l.keys, l.values) for l in self.layers)
        DynamicCache.from_legacy_cache = classmethod(lambda cls, t: cls(ddp_cache_data=t))
    from src.benchmarking.benchmarker import KVCacheBenchmarker
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config()).eval()
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 50257, (1, 16), generator=g)
    bench = KVCacheBenchmarker(model, StubTokenizer(ids), device="cpu")
    out = {"prompt_ids": ids.numpy(), "new_tokens": np.array(8)}
    text, n_new = bench.generate_with_cache("x", max_new_tokens=8)
    out["full/tokens"] = np.array([int(t) for t in text.split()])
    for mode in ("int8", "int4", "mixed"):
        text, n_new, est_mb = bench.generate_with_quantized_kv("x", max_new_tokens=8, mode=mode)
        out[f"{mode}/tokens"] = np.array([int(t) for t in text.split()])
        out[f"{mode}/n_new"] = np.array(n_new)
        out[f"{mode}/est_mb"] = np.array(est_mb, dtype=np.float64)
        print("benchmarker", mode, n_new, repr(est_mb))
    np.savez_compressed(os.path.join(HERE, "benchmarker.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "benchmarker":
        make_benchmarker()
        sys.exit(0)
    make_quant()
    make_pertoken()
    make_evict()
    make_evict_big()
    make_chunk()
    make_paged()
    make_attn()
    make_benchmarker()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
