"""The layer-batched entry points (one launch for all layers x {K, V}): kvq_evict_gather_batched,
kvq_chunk_mean_batched, kvq_quant_per_token_fused_batched, the device-built budget table, and the vectorised
paged append -- each bit-identical to the per-tensor path and to the oracle (reference loops they replace:
src/cache/implementations.py:176-190, :218-245, :268-292, :316-346; src/quantization/ops.py:323-342)."""

import ctypes

import numpy as np
import pytest
import torch

from _golden import bits_equal, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kvq():
    import kvq_b200
    assert torch.cuda.is_available()
    return kvq_b200


def gen(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).to(dtype)


TRIMS = [("trim_kv_prefix_window", dict(prefix_len=4, window_size=8)),
         ("trim_kv_strided", dict(window_size=64, stride=4, prefix_len=2)),
         ("trim_kv_block_old", dict(window_size=100, block_size=64, keep_per_block=8)),
         ("trim_kv_block_old", dict(window_size=8, block_size=10, keep_per_block=2, prefix_len=3)),
         ("trim_kv_budget_old", dict(window_size=32, old_budget=64)),
         ("trim_kv_budget_old", dict(window_size=16, old_budget=0, prefix_len=1))]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 4, 1500, 64), (1, 8, 2049, 128), (3, 2, 333, 8)])
def test_layer_batched_trims_match_oracle(kvq, oracle, dtype, shape, monkeypatch):
    """A multi-layer tuple goes through ONE kvq_evict_gather_batched launch; results equal the oracle's per layer."""
    from kvq_b200 import _lib as L
    n_layers = 5
    cpu = [(gen(shape, dtype, 100 + i), gen(shape, dtype, 200 + i)) for i in range(n_layers)]
    dev = tuple((k.cuda(), v.cuda()) for k, v in cpu)
    for fn, kw in TRIMS:
        before = L.launch_count
        got = getattr(kvq, fn)(dev, **kw)
        used = L.launch_count - before
        want = getattr(oracle, fn)(tuple(cpu), **kw)
        assert used == 1, (fn, used)  # one launch for all 10 tensors (a needs_table budget plan would add one)
        for (gk, gv), (wk, wv) in zip(got, want):
            assert bits_equal(gk.cpu().contiguous(), wk.contiguous()) and bits_equal(gv.cpu().contiguous(), wv.contiguous()), (fn, kw)


def test_layer_batched_trim_falls_back_for_mixed_layouts(kvq, oracle):
    a, b = gen((1, 2, 300, 64), torch.float16, 1), gen((1, 2, 301, 64), torch.float16, 2)  # two lengths: per-tensor launches
    got = kvq.trim_kv_strided(((a.cuda(), a.cuda()), (b.cuda(), b.cuda())), window_size=32, stride=3)
    want = oracle.trim_kv_strided(((a, a), (b, b)), window_size=32, stride=3)
    for (gk, gv), (wk, wv) in zip(got, want):
        assert bits_equal(gk.cpu(), wk) and bits_equal(gv.cpu(), wv)
    odd = gen((1, 2, 300, 7), torch.float16, 3)  # 14-byte rows: generic kernel, per tensor
    got = kvq.trim_kv_block_old(((odd.cuda(), odd.cuda()),) * 3, window_size=16, block_size=8, keep_per_block=2)
    want = oracle.trim_kv_block_old(((odd, odd),) * 3, window_size=16, block_size=8, keep_per_block=2)
    assert all(bits_equal(g[0].cpu(), w[0]) for g, w in zip(got, want))


def test_budget_table_is_built_on_the_device(kvq):
    """needs_table plans (float32 linspace produced duplicates): the device-built table equals the host index list,
    for the golden million-token cases."""
    from kvq_b200 import _lib as L
    from kvq_b200.cache import evict_plan
    from kvq_b200.cache.implementations import _old_table, kept_token_indices
    gb = load("evict_big.npz")
    seen = 0
    for i, (T, w, p, bud) in enumerate(gb["params"].tolist()):
        plan = evict_plan(L.BUDGET, T, window=w, prefix=p, budget=bud)
        if not plan.needs_table:
            continue
        seen += 1
        host = kept_token_indices(plan)[plan.n_prefix: plan.n_prefix + plan.n_old]
        table = _old_table(plan, "cuda")
        assert table.dtype == torch.int32 and table.cpu().tolist() == host
    assert seen >= 1


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_layer_batched_chunk_summary(kvq, oracle, dtype):
    from kvq_b200 import _lib as L
    shape = (2, 4, 700, 64)
    cpu = [(gen(shape, dtype, 300 + i), gen(shape, dtype, 400 + i)) for i in range(4)]
    dev = tuple((k.cuda(), v.cuda()) for k, v in cpu)
    before = L.launch_count
    got = kvq.chunk_summarize_kv(dev, chunk_size=64, keep_last=100)
    assert L.launch_count - before == 1
    single = [kvq.chunk_summarize_kv((kv,), chunk_size=64, keep_last=100)[0] for kv in dev]
    want = oracle.chunk_summarize_kv(tuple(cpu), chunk_size=64, keep_last=100)
    tol = {torch.float32: 1e-6, torch.float16: 2 ** -10, torch.bfloat16: 2 ** -7}[dtype]
    for (gk, gv), (sk, sv), (wk, wv) in zip(got, single, want):
        assert torch.equal(gk, sk) and torch.equal(gv, sv)  # same kernel, same order of additions
        torch.testing.assert_close(gk.cpu().float(), wk.float(), rtol=tol, atol=1e-6)
        assert torch.equal(gk[:, :, -100:].cpu(), wk[:, :, -100:])
    # nothing older than keep_last: same objects back
    same = kvq.chunk_summarize_kv(dev, chunk_size=64, keep_last=700)
    assert same[0][0] is dev[0][0]


@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_layer_batched_per_token_append(kvq, oracle, mode, dtype):
    """QuantizedKVCache.init_from_prompt_past / append_from_past: one launch per bit width for ALL layers; codes and scales
    bit-identical to the oracle's per-token quantizer."""
    from kvq_b200 import _lib as L
    n_layers, B, H, T, D = 6, 2, 4, 37, 64
    kb, vb = (4 if mode == "int4" else 8), (8 if mode == "int8" else 4)
    past = [(gen((B, H, T, D), dtype, 10 + i), gen((B, H, T, D), dtype, 50 + i)) for i in range(n_layers)]
    qc = kvq.QuantizedKVCache(n_layers=n_layers, mode=mode, device="cuda", compute_dtype=dtype)
    before = L.launch_count
    qc.init_from_prompt_past(tuple((k.cuda(), v.cuda()) for k, v in past))
    assert L.launch_count - before == (1 if kb == vb else 2)
    nxt = [(torch.cat([k, gen((B, H, 1, D), dtype, 90 + i)], 2), torch.cat([v, gen((B, H, 1, D), dtype, 130 + i)], 2))
           for i, (k, v) in enumerate(past)]
    before = L.launch_count
    qc.append_from_past(tuple((k.cuda(), v.cuda()) for k, v in nxt))
    assert L.launch_count - before == (1 if kb == vb else 2)
    for i, (k, v) in enumerate(nxt):
        kc, ks, _, vc, vs, _ = qc.layers[i].quantized()
        ok, oks, _ = oracle.quantize_per_token(k, kb)
        ov, ovs, _ = oracle.quantize_per_token(v, vb)
        assert torch.equal(kc.cpu(), ok) and torch.equal(vc.cpu(), ov), i
        assert torch.equal(ks.cpu(), oks.float()) and torch.equal(vs.cpu(), ovs.float()), i
        assert torch.equal(qc.layers[i]._k.live_scales().cpu(), oks)


def test_nan_and_custom_eps(kvq, oracle):
    """torch's clamp(min=eps) propagates NaN: a NaN input gives a NaN scale (not eps).  A custom eps and a calibrated absmax
    take the exact-division kernels: codes equal the oracle's true division."""
    x = gen((2, 2, 16, 64), torch.float16, 3)
    x[0, 0, 3, 5] = float("nan")
    for fn in (kvq.quantize_int8_per_tensor, lambda t: kvq.quantize_int4_per_tensor_packed(t)[:2]):
        _, s = fn(x.cuda())
        assert torch.isnan(s).item()
    y = gen((2, 3, 33, 64), torch.bfloat16, 4) * 1e-3
    for eps in (1e-3, 3.3e-5):
        q, s = kvq.quantize_int8_per_tensor(y.cuda(), eps=eps)
        oq, os_ = oracle.quantize_int8_per_tensor(y, eps=eps)
        assert torch.equal(q.cpu(), oq) and float(s) == float(os_)
        p, s4, _ = kvq.quantize_int4_per_tensor_packed(y.cuda(), eps=eps)
        op, os4, _ = oracle.quantize_int4_per_tensor_packed(y, eps=eps)
        assert torch.equal(p.cpu(), op) and float(s4) == float(os4)


def test_store_capacity_growth_amortises(kvq):
    """A window just under the capacity must not re-copy the store every other step."""
    layer = kvq.QuantizedLayerKV("int8", "cuda", torch.float16)
    x = gen((1, 2, 1, 64), torch.float16, 1).cuda()
    copies, last = 0, None
    for t in range(1200):
        layer.append(x, x)
        layer.evict("sliding_window", window_size=250)
        if layer._k.codes is not last:
            copies += last is not None
            last = layer._k.codes
    assert len(layer) == 250 and copies <= 3, copies


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(16, 1, 2048, 128), (4, 2, 1500, 64), (2, 4, 1100, 128), (1, 1, 1025, 64), (3, 1, 4099, 32)])
def test_small_tokens_several_per_cta(kvq, oracle, dtype, shape):
    """Small tokens (a head shard at N = 8: 4 KB per token) take the several-tokens-per-CTA form of the fused per-token quantizer
    (32 / 64 / 128 threads per token): codes and scales bit-identical to the oracle, for contiguous and strided inputs."""
    x = gen(shape, dtype, shape[2])
    for mode, bits in (("int8", 8), ("int4", 4)):
        layer = kvq.QuantizedLayerKV(mode, "cuda", dtype)
        xd = x.cuda()
        hf = xd.transpose(1, 2).contiguous().transpose(1, 2)   # [B,T,H,D] memory viewed as [B,H,T,D]
        layer.append_tokens(hf, xd)
        kc, ks, _, vc, vs, _ = layer.quantized()
        oq, osc, _ = oracle.quantize_per_token(x, bits)
        assert torch.equal(kc.cpu(), oq) and torch.equal(vc.cpu(), oq)
        assert torch.equal(ks.cpu(), osc.float()) and torch.equal(vs.cpu(), osc.float())


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("world", [1, 2, 5])
def test_publish_gather_on_one_device_equals_the_plain_batched_pair(kvq, bits, world):
    """kvq_absmax_batched_publish / kvq_quant_batched_gather (the head-sharded pair with the exchange inside the kernels) with
    all `world` ranks played by ONE device: every rank publishes into every rank's buffer, then every rank gathers.  Codes and
    scales of every shard equal the single-GPU batched quantization of the full tensors, for several steps in a row (slot
    parity, arrival counters re-armed, the max moving between ranks).  Reference: scale over all heads, ops.py:27,:48."""
    from kvq_b200 import _lib as L
    from kvq_b200.quantization.ops import TensorTable, quantize_per_tensor_batched
    lib, dev = L.lib(), torch.device("cuda")
    nbytes = int(lib.kvq_absmax_exchange_buffer_bytes())
    bufs = [torch.zeros(nbytes // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64).to(dev)
    B, H, T, D, nt = 2, 2 * world, 2048, 64, 3       # 2048 * 64 * 2 heads * 2 B * 2 = 1 MiB per shard: many CTAs per tensor
    W = D if bits == 8 else D // 2
    cdt = torch.int8 if bits == 8 else torch.uint8
    stream = L.stream_ptr(dev)
    for step in range(1, 6):
        g = torch.Generator(device="cuda").manual_seed(100 * step + world)
        fulls = [(torch.randn(B, H, T, D, device=dev, generator=g) * (i + step)).to(torch.bfloat16) for i in range(nt)]
        fulls[step % nt][1, (3 * step) % H, 17, 5] = 500.0 + step
        fq = [torch.empty(B, H, T, W, device=dev, dtype=cdt) for _ in range(nt)]
        fs = torch.empty(nt, device=dev, dtype=torch.bfloat16)
        fb = quantize_per_tensor_batched(TensorTable(fulls), TensorTable(fq), bits, fs)
        Hl = H // world
        shards = [[f[:, r * Hl:(r + 1) * Hl].contiguous() for f in fulls] for r in range(world)]
        tabs = [TensorTable(s) for s in shards]
        local = [torch.empty(nt, dtype=torch.int32, device=dev) for _ in range(world)]
        for r in range(world):
            L.check(lib.kvq_absmax_batched_publish(tabs[r].ptrs.data_ptr(), L.BF16, nt, tabs[r].numel, local[r].data_ptr(), peers.data_ptr(),
                                                   r, world, step, stream), "publish")
        for r in range(world):
            sq = [torch.empty(B, Hl, T, W, device=dev, dtype=cdt) for _ in range(nt)]
            qt = TensorTable(sq)
            ss = torch.empty(nt, device=dev, dtype=torch.bfloat16)
            comb = torch.empty(nt, dtype=torch.int32, device=dev)
            L.check(lib.kvq_quant_batched_gather(tabs[r].ptrs.data_ptr(), L.BF16, nt, tabs[r].numel, bits | L.ABSMAX_COVERS_INPUT, comb.data_ptr(),
                                                 1e-8, qt.ptrs.data_ptr(), ss.data_ptr(), None, bufs[r].data_ptr(), world, step, stream), "gather")
            assert torch.equal(comb, fb), (step, r)
            assert torch.equal(ss.view(torch.int16), fs.view(torch.int16))
            for a, f in zip(sq, fq):
                assert torch.equal(a, f[:, r * Hl:(r + 1) * Hl])
            # the local words are the shard's own absmax
            want = torch.stack([s.abs().max().float() for s in shards[r]]).view(torch.int32)
            assert torch.equal(local[r], want)


def test_publish_gather_argument_checks(kvq):
    from kvq_b200 import _lib as L
    lib, dev = L.lib(), torch.device("cuda")
    buf = torch.zeros(int(lib.kvq_absmax_exchange_buffer_bytes()) // 4, dtype=torch.int32, device=dev)
    peers = torch.tensor([buf.data_ptr()], dtype=torch.int64).to(dev)
    x = torch.zeros(1, 1, 8, 64, device=dev, dtype=torch.bfloat16)
    xt = torch.tensor([x.data_ptr()], dtype=torch.int64).to(dev)
    bits = torch.zeros(1, dtype=torch.int32, device=dev)
    st = L.stream_ptr(dev)
    assert lib.kvq_absmax_batched_publish(xt.data_ptr(), L.BF16, 1, x.numel(), bits.data_ptr(), peers.data_ptr(), 0, 1, 0, st) == -1   # seq 0
    assert lib.kvq_absmax_batched_publish(xt.data_ptr(), L.BF16, 1, x.numel(), bits.data_ptr(), peers.data_ptr(), 1, 1, 1, st) == -1   # rank >= world
    assert lib.kvq_absmax_batched_publish(xt.data_ptr(), L.BF16, 2000, x.numel(), bits.data_ptr(), peers.data_ptr(), 0, 1, 1, st) == -1  # too many tensors
    assert lib.kvq_absmax_batched_publish(xt.data_ptr(), L.BF16, 1, x.numel(), bits.data_ptr(), None, 0, 1, 1, st) == -1
    q = torch.empty(1, 1, 8, 64, device=dev, dtype=torch.int8)
    qt = torch.tensor([q.data_ptr()], dtype=torch.int64).to(dev)
    sc = torch.empty(1, device=dev, dtype=torch.bfloat16)
    assert lib.kvq_quant_batched_gather(xt.data_ptr(), L.BF16, 1, x.numel(), 8, bits.data_ptr(), 1e-8, qt.data_ptr(), sc.data_ptr(), None, None, 1, 1, st) == -1
    assert lib.kvq_quant_batched_gather(xt.data_ptr(), L.BF16, 1, x.numel(), 8, bits.data_ptr(), 1e-8, qt.data_ptr(), sc.data_ptr(), None, buf.data_ptr(), 1, 0, st) == -1


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("shape", [(8, 2, 1, 128), (2, 3, 37, 64), (1, 1, 5, 8)])
def test_per_token_exchange_with_one_rank_equals_the_plain_fused_quantizer(kvq, bits, shape):
    """kvq_quant_per_token_fused_exchange with world = 1 (the rank's own buffer is the only peer): same codes, scales and absmax
    words as kvq_quant_per_token_fused, call after call (slot parity / step tags).  The 2-GPU form is in test_gpu_multi.py."""
    from kvq_b200 import _lib as L
    lib, dev = L.lib(), torch.device("cuda")
    buf = torch.zeros(int(lib.kvq_absmax_exchange_buffer_bytes()) // 4, dtype=torch.int32, device=dev)
    peers = torch.tensor([buf.data_ptr()], dtype=torch.int64).to(dev)
    B, H, T, D = shape
    W = D if bits == 8 else D // 2
    st = L.stream_ptr(dev)
    for step in range(1, 5):
        x = (gen(shape, torch.bfloat16, 50 + step) * step).to(dev)
        outs = []
        for form in ("plain", "exchange", "exchange_device_step"):
            q = torch.zeros(B, H, T, W, device=dev, dtype=torch.uint8)
            sc = torch.zeros(T, device=dev, dtype=torch.bfloat16)
            s32 = torch.zeros(T, device=dev, dtype=torch.float32)
            am = torch.zeros(T, device=dev, dtype=torch.int32)
            sh, sx = L.shape_stride(x)
            if form == "plain":
                L.check(lib.kvq_quant_per_token_fused(x.data_ptr(), L.BF16, sh, sx, bits, 1e-8, q.data_ptr(), L.stride3(q), sc.data_ptr(), s32.data_ptr(),
                                                      am.data_ptr(), st), "plain")
            else:   # step number from the host, or (0) from the counter in the buffer
                L.check(lib.kvq_quant_per_token_fused_exchange(x.data_ptr(), L.BF16, sh, sx, bits, 1e-8, q.data_ptr(), L.stride3(q), sc.data_ptr(),
                                                               s32.data_ptr(), am.data_ptr(), peers.data_ptr(), 0, 1, step if form == "exchange" else 0,
                                                               st), "exchange")
            outs.append((q, sc, s32, am))
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a, b)
        # the device counter: `step` completed calls, arrivals re-armed
        stepw = buf.view(torch.uint8)[-16:].view(torch.int32)
        assert stepw[0].item() == step and stepw[1].item() == 0
    # argument checks: a bad rank, a custom eps (exact-division path: not in the single-pass kernel)
    x = gen((1, 1, 4, 64), torch.bfloat16, 1).to(dev)
    q = torch.zeros(1, 1, 4, 64, device=dev, dtype=torch.uint8)
    sh, sx = L.shape_stride(x)
    assert lib.kvq_quant_per_token_fused_exchange(x.data_ptr(), L.BF16, sh, sx, 8, 1e-8, q.data_ptr(), L.stride3(q), None, None, None, peers.data_ptr(), 1, 1, 3, st) == -1
    assert lib.kvq_quant_per_token_fused_exchange(x.data_ptr(), L.BF16, sh, sx, 8, 1e-3, q.data_ptr(), L.stride3(q), None, None, None, peers.data_ptr(), 0, 1, 9, st) == -3
