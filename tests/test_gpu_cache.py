"""Areas 2 and 3 parity on the GPU, through the C ABI: eviction index sets and gathered rows bit-exact
(golden + oracle), paged store bit-exact, chunk summaries within the stated tolerance (recent tokens
bit-exact), plus long-context (BASELINE C4, T = 131072) property checks."""

import ctypes

import numpy as np
import pytest
import torch

from _golden import DT, bits_equal, load, typed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kvq():
    import kvq_b200
    assert torch.cuda.is_available()
    return kvq_b200


def gen(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).to(dtype)


def device_indices(plan):
    from kvq_b200 import _lib as L
    from kvq_b200.cache.implementations import _old_table
    out = torch.empty(max(int(plan.n_total), 1), dtype=torch.int64, device="cuda")
    table = _old_table(plan, "cuda")
    L.check(L.lib().kvq_evict_indices_device(ctypes.byref(plan), table.data_ptr() if table is not None else None, out.data_ptr(),
                                             L.stream_ptr(out.device)), "kvq_evict_indices_device")
    return out[: plan.n_total].cpu().numpy()


def test_device_index_sets_match_golden(kvq):
    """The __device__ index function the gather kernel uses, evaluated on the GPU for every golden case
    (incl. float32 linspace at T = 131072 and the million-token duplicate cases in table mode)."""
    from kvq_b200.cache import evict_plan
    g = load("evict.npz")
    params, offs, flat = g["params"], g["offsets"], g["indices"]
    for i, (st, T, w, p, s, blk, keep, bud) in enumerate(params.tolist()):
        plan = evict_plan(st, T, window=w, prefix=p, stride=s, block=blk, keep=keep, budget=bud)
        assert np.array_equal(device_indices(plan), flat[offs[i]:offs[i + 1]].astype(np.int64)), (st, T, w, p, s, blk, keep, bud)
    gb = load("evict_big.npz")
    for i, (T, w, p, bud) in enumerate(gb["params"].tolist()):
        want = np.concatenate([[int(gb[f"{i}/first"])], int(gb[f"{i}/first"]) + np.cumsum(gb[f"{i}/deltas"].astype(np.int64))])
        assert np.array_equal(device_indices(evict_plan(4, T, window=w, prefix=p, budget=bud)), want)


def test_golden_gathered_values(kvq):
    g = load("evict.npz")
    for tag in DT:
        k, v = typed(g, f"data/k_{tag}", tag).cuda(), typed(g, f"data/v_{tag}", tag).cuda()
        (k2, v2), = kvq.trim_kv_block_old(((k, v),), window_size=8, block_size=10, keep_per_block=2, prefix_len=3)
        assert bits_equal(k2.cpu(), typed(g, f"data/block_k_{tag}", tag)) and bits_equal(v2.cpu(), typed(g, f"data/block_v_{tag}", tag))
        (k3, _), = kvq.trim_kv_budget_old(((k, v),), window_size=8, old_budget=5, prefix_len=2)
        assert bits_equal(k3.cpu(), typed(g, f"data/budget_k_{tag}", tag))


CASES = [dict(fn="trim_kv_prefix_window", kw=dict(prefix_len=4, window_size=8)),
         dict(fn="trim_kv_prefix_window", kw=dict(prefix_len=3, window_size=0)),
         dict(fn="trim_kv_strided", kw=dict(window_size=8, stride=4, prefix_len=2)),
         dict(fn="trim_kv_strided", kw=dict(window_size=256, stride=4)),
         dict(fn="trim_kv_block_old", kw=dict(window_size=8, block_size=10, keep_per_block=2, prefix_len=3)),
         dict(fn="trim_kv_block_old", kw=dict(window_size=256, block_size=64, keep_per_block=8)),
         dict(fn="trim_kv_budget_old", kw=dict(window_size=8, old_budget=5, prefix_len=2)),
         dict(fn="trim_kv_budget_old", kw=dict(window_size=256, old_budget=64)),
         dict(fn="trim_kv_budget_old", kw=dict(window_size=16, old_budget=0, prefix_len=1)),
         dict(fn="trim_kv_budget_old", kw=dict(window_size=16, old_budget=100000)),
         dict(fn="trim_kv_sliding_window", kw=dict(window_size=256))]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 3, 40, 8), (2, 4, 1500, 64), (1, 8, 2049, 128), (2, 2, 300, 7), (1, 2, 777, 96)])
def test_oracle_trims(kvq, oracle, dtype, shape):
    k, v = gen(shape, dtype, 11), gen(shape, dtype, 12)
    hf = k.cuda().transpose(1, 2).contiguous().transpose(1, 2)  # strided [B,H,T,D] view
    for case in CASES:
        (gk, gv), = getattr(kvq, case["fn"])(((hf, v.cuda()),), **case["kw"])
        (ok, ov), = getattr(oracle, case["fn"])(((k, v),), **case["kw"])
        assert bits_equal(gk.cpu().contiguous(), ok.contiguous()) and bits_equal(gv.cpu().contiguous(), ov.contiguous()), case


def test_gather_on_quantized_codes(kvq, oracle):
    """The same kernel compacts INT8 codes and packed INT4 bytes (1-byte elements): evict(quantize(x)) ==
    quantize(evict(x)) for per-token scales."""
    from kvq_b200 import _lib as L
    from kvq_b200.cache import evict_plan
    from kvq_b200.cache.implementations import gather_tokens
    x = gen((2, 4, 1000, 128), torch.float16, 5)
    layer = kvq.QuantizedLayerKV("mixed", "cuda", torch.float16)
    layer.append_tokens(x.cuda(), x.cuda())
    kc, _, _, vc, _, _ = layer.quantized()
    plan = evict_plan(L.STRIDED, 1000, window=256, stride=4)
    idx = oracle.evict_indices(2, 1000, window=256, stride=4)
    assert torch.equal(gather_tokens(kc, plan).cpu(), kc.cpu()[:, :, idx]) and torch.equal(gather_tokens(vc, plan).cpu(), vc.cpu()[:, :, idx])
    (xe, _), = kvq.trim_kv_strided(((x.cuda(), x.cuda()),), 256, 4)
    l2 = kvq.QuantizedLayerKV("mixed", "cuda", torch.float16)
    l2.append_tokens(xe, xe)
    assert torch.equal(l2.quantized()[0], gather_tokens(kc, plan))


def test_golden_paged(kvq):
    g = load("paged.npz")
    for tag, dt in DT.items():
        k, v = typed(g, f"{tag}/k", tag).cuda(), typed(g, f"{tag}/v", tag).cuda()
        blk, nblocks, alloc, used, filled = g[f"{tag}/meta"].tolist()
        pc = kvq.PagedKVCache(block_size=blk, device="cuda", dtype=dt)
        for t in range(k.shape[2]):
            pc.append(k[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
        kk, vv = pc.get_kv()
        assert bits_equal(kk.cpu(), typed(g, f"{tag}/kk", tag)) and bits_equal(vv.cpu(), typed(g, f"{tag}/vv", tag))
        assert (pc.num_blocks(), pc.allocated_bytes(), pc.used_bytes(), pc.t_filled) == (nblocks, alloc, used, filled)
        assert len(pc.k_blocks) == nblocks and pc.k_blocks[0].shape == (k.shape[0], k.shape[1], blk, k.shape[3])


@pytest.mark.parametrize("shape,blk", [((2, 12, 300, 64), 64), ((1, 2, 50, 7), 4), ((2, 8, 131, 128), 16)])
def test_paged_growth_and_table(kvq, shape, blk):
    k, v = gen(shape, torch.float16, 1).cuda(), gen(shape, torch.float16, 2).cuda()
    hf = k.transpose(1, 2).contiguous().transpose(1, 2)
    pc = kvq.PagedKVCache(block_size=blk, device="cuda", dtype=torch.float16)
    for t in range(shape[2]):
        pc.append(hf[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
    kk, vv = pc.get_kv()
    assert torch.equal(kk, k) and torch.equal(vv, v)
    table = pc.block_table.cpu().tolist()
    assert len(set(table)) == len(table) == pc.num_blocks() == -(-shape[2] // blk)
    assert table != sorted(table) or len(table) == 1  # blocks are not laid out in logical order


CHUNK_TOL = {torch.float32: dict(rtol=1e-6, atol=1e-6), torch.float16: dict(rtol=2 ** -10, atol=1e-6),
             torch.bfloat16: dict(rtol=2 ** -7, atol=1e-6)}


def test_golden_chunk(kvq):
    g = load("chunk.npz")
    got = kvq.chunk_summarize_kv(((torch.ones(1, 1, 10, 2, device="cuda"),) * 2,), 4, 4)[0][0]
    assert got[0, 0, :, 0].tolist() == [1.0, 0.5, 1, 1, 1, 1]
    for case in g["cases"]:
        tag = case.split("_")[0]
        x = typed(g, f"{case}/x_{tag}", tag).cuda()
        chunk, keep = g[f"{case}/params"].tolist()
        want = typed(g, f"{case}/out_{tag}", tag)
        (got, got_v), = kvq.chunk_summarize_kv(((x, x),), chunk_size=chunk, keep_last=keep)
        assert got.shape == want.shape and got.dtype == want.dtype and torch.equal(got, got_v)
        torch.testing.assert_close(got.cpu().float(), want.float(), **CHUNK_TOL[DT[tag]])
        keep_eff = min(keep, x.shape[2])
        if keep_eff and x.shape[2] - keep_eff > 0:
            assert bits_equal(got[:, :, -keep_eff:, :].cpu().contiguous(), want[:, :, -keep_eff:, :].contiguous())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,chunk,keep", [((2, 4, 1000, 64), 64, 256), ((1, 8, 2049, 128), 64, 256), ((2, 3, 77, 7), 5, 13),
                                              ((1, 2, 100, 96), 7, 0), ((2, 2, 50, 16), 64, 3)])
def test_oracle_chunk(kvq, oracle, dtype, shape, chunk, keep):
    x = gen(shape, dtype, 21)
    hf = x.cuda().transpose(1, 2).contiguous().transpose(1, 2)
    (got, _), = kvq.chunk_summarize_kv(((hf, hf),), chunk, keep)
    want = oracle.chunk_mean(x, chunk, keep)
    assert got.shape == want.shape
    torch.testing.assert_close(got.cpu().float(), want.float(), **CHUNK_TOL[dtype])


def test_long_context_properties(kvq, oracle):
    """BASELINE C4: Llama-3-8B shape, T = 131072, bf16 (B=1, H=8, D=128).  Full size, checked through
    properties: kept counts, rows equal torch indexing with the oracle's index set, composition
    (re-trimming a trimmed cache equals trimming twice on indices), chunk means against fp64 torch."""
    B, H, T, D = 1, 8, 131072, 128
    torch.manual_seed(1234)
    k = torch.randn(B, H, T, D, device="cuda", dtype=torch.bfloat16)
    for fn, st, kw, okw, n in (("trim_kv_strided", 2, dict(window_size=256, stride=4), dict(window=256, stride=4), 32960),
                               ("trim_kv_block_old", 3, dict(window_size=256, block_size=64, keep_per_block=8), dict(window=256, block=64, keep=8), 16608),
                               ("trim_kv_budget_old", 4, dict(window_size=256, old_budget=64), dict(window=256, budget=64), 320)):
        (gk, gv), = getattr(kvq, fn)(((k, k),), **kw)
        idx = oracle.evict_indices(st, T, **okw).cuda()
        assert gk.shape == (B, H, n, D) and torch.equal(gk, k[:, :, idx]) and torch.equal(gv, gk)
        (g2, _), = getattr(kvq, fn)(((gk, gk),), **kw)  # the benchmarker re-applies the trim every step
        idx2 = oracle.evict_indices(st, n, **okw).cuda()
        assert torch.equal(g2, gk[:, :, idx2])
    (ck, _), = kvq.chunk_summarize_kv(((k, k),), 64, 256)
    assert ck.shape == (B, H, 2044 + 256, D)
    want = k[:, :, : T - 256].double().view(B, H, 2044, 64, D).mean(3)
    torch.testing.assert_close(ck[:, :, :2044].double(), want, rtol=2 ** -7, atol=1e-6)
    assert torch.equal(ck[:, :, 2044:], k[:, :, T - 256:])


@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
def test_evict_on_quantized_store_equals_quantizing_the_evicted_cache(kvq, oracle, mode):
    """QuantizedLayerKV.evict(): every strategy applied to the code store == the reference's trim applied to the fp cache
    followed by per-token quantization (codes and scales bit-identical, so dequantized tensors too)."""
    x = gen((2, 4, 1100, 64), torch.float16, 31)
    y = gen((2, 4, 1100, 64), torch.float16, 32)
    for strat, kw, fn, fkw in (("sliding_window", dict(window_size=256), "trim_kv_sliding_window", dict(window_size=256)),
                               ("prefix_window", dict(window_size=128, prefix_len=16), "trim_kv_prefix_window", dict(window_size=128, prefix_len=16)),
                               ("strided", dict(window_size=256, stride=4, prefix_len=2), "trim_kv_strided", dict(window_size=256, stride=4, prefix_len=2)),
                               ("block", dict(window_size=256, block_size=64, keep_per_block=8), "trim_kv_block_old", dict(window_size=256, block_size=64, keep_per_block=8)),
                               ("budget", dict(window_size=256, old_budget=64), "trim_kv_budget_old", dict(window_size=256, old_budget=64)),
                               ("sliding_window", dict(window_size=5000), "trim_kv_sliding_window", dict(window_size=5000))):
        layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float16)
        layer.append_tokens(x.cuda(), y.cuda())
        n = layer.evict(strat, **kw)
        (tx, ty), = getattr(oracle, fn)(((x, y),), **fkw)
        assert n == tx.size(2) == len(layer)
        kb, vb = (4 if mode == "int4" else 8), (8 if mode == "int8" else 4)
        ok, oks, _ = oracle.quantize_per_token(tx.contiguous(), kb)
        ov, ovs, _ = oracle.quantize_per_token(ty.contiguous(), vb)
        kc, ks, _, vc, vs, _ = layer.quantized()
        assert torch.equal(kc.cpu(), ok) and torch.equal(vc.cpu(), ov), strat
        assert torch.equal(ks.cpu(), oks.float()) and torch.equal(vs.cpu(), ovs.float()), strat
        layer.append_tokens(x[:, :, :3].cuda(), y[:, :, :3].cuda())  # the store keeps working after compaction
        assert len(layer) == n + 3 and layer.estimated_bytes() > 0


def test_paged_bulk_append_equals_token_by_token(kvq):
    k, v = gen((2, 3, 150, 64), torch.float16, 8).cuda(), gen((2, 3, 150, 64), torch.float16, 9).cuda()
    hf = k.transpose(1, 2).contiguous().transpose(1, 2)
    a, b = kvq.PagedKVCache(16, "cuda", torch.float16), kvq.PagedKVCache(16, "cuda", torch.float16)
    a.append_tokens(hf[:, :, :70], v[:, :, :70])
    a.append(hf[:, :, 70:71], v[:, :, 70:71])
    a.append_tokens(hf[:, :, 71:], v[:, :, 71:])
    for t in range(150):
        b.append(hf[:, :, t:t + 1], v[:, :, t:t + 1])
    (ka, va), (kb, vb) = a.get_kv(), b.get_kv()
    assert torch.equal(ka, k) and torch.equal(va, v) and torch.equal(kb, k) and torch.equal(vb, v)
    assert (a.num_blocks(), a.allocated_bytes(), a.used_bytes(), a.t_filled) == (b.num_blocks(), b.allocated_bytes(), b.used_bytes(), b.t_filled)
