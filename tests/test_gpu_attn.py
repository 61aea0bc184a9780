"""Area 4 on the GPU: fused dequant decode attention against the golden fp32 eager attention over the
reference's dequantized cache and against the oracle.  Tolerance (stated): fp32 accumulation with a
different summation order and exp2-domain softmax -> rtol 1e-4 / atol 1e-5 for fp32 outputs; for 16-bit
outputs one more rounding of the result: fp16 rtol = atol = 2e-3 (2^-11 relative + slack), bf16
rtol = atol = 8e-3 (2^-8 relative)."""

import pytest
import torch

from _golden import load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kvq():
    import kvq_b200
    assert torch.cuda.is_available()
    return kvq_b200


def test_golden_attention(kvq):
    g = load("attn.npz")
    ran = 0
    for case in g["cases"]:
        mode = case.split("_")[0]
        D = int(case.split("x")[-1])
        # head_dim 64 / 128 take the tiled kernels, the other golden cases (16, 8) the generic any-head_dim kernel
        kb, vb = (4 if mode == "int4" else 8), (8 if mode == "int8" else 4)
        q = torch.from_numpy(g[f"{case}/q"]).cuda()
        for variant in ("cuda_core", "imma"):
            out = kvq.decode_attention_quantized(q, torch.from_numpy(g[f"{case}/kstore"]).cuda(), torch.from_numpy(g[f"{case}/kscales"]).cuda(), kb,
                                                 torch.from_numpy(g[f"{case}/vstore"]).cuda(), torch.from_numpy(g[f"{case}/vscales"]).cuda(), vb,
                                                 variant=variant)
            torch.testing.assert_close(out.cpu(), torch.from_numpy(g[f"{case}/out"]), rtol=1e-4, atol=1e-5)
        ran += 1
    assert ran >= 4


@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
@pytest.mark.parametrize("B,Hq,H,T,D", [(2, 4, 4, 300, 64), (1, 8, 8, 1000, 128), (2, 8, 2, 513, 128), (1, 16, 2, 2000, 64),
                                         (16, 12, 12, 1024, 64), (1, 8, 1, 5000, 128), (2, 4, 2, 1, 128), (1, 2, 2, 129, 64),
                                         (2, 8, 2, 700, 64), (3, 16, 4, 257, 128), (1, 32, 4, 40000, 128),
                                         (1, 3, 3, 1, 64), (2, 2, 2, 17, 128), (1, 8, 1, 1, 128), (2, 8, 2, 2, 64), (1, 1, 1, 33000, 128),
                                         (2, 8, 4, 777, 128), (1, 24, 12, 1500, 64), (4, 4, 2, 130, 64), (1, 2, 1, 9000, 128)])
@pytest.mark.parametrize("variant", ["cuda_core", "imma"])
def test_oracle_attention(kvq, oracle, mode, B, Hq, H, T, D, variant):
    g = torch.Generator().manual_seed(B * 1000 + T)
    k = torch.randn(B, H, T, D, generator=g)
    k.view(-1)[::97] *= 8
    v = torch.randn(B, H, T, D, generator=g)
    q = torch.randn(B, Hq, 1, D, generator=g)
    layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
    layer.append_tokens(k.cuda(), v.cuda())
    kc, ks, kb, vc, vs, vb = layer.quantized()
    want = oracle.decode_attention(q, kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu())
    if Hq // H not in (1, 2, 4, 8) and variant == "imma":
        pytest.skip("padded groupings are covered by test_any_group_size")
    out = kvq.decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb, variant=variant)
    torch.testing.assert_close(out.cpu(), want, rtol=1e-4, atol=1e-5)
    for dt in (torch.float16, torch.bfloat16):
        o16 = kvq.decode_attention_quantized(q.cuda().to(dt), kc, ks, kb, vc, vs, vb, variant=variant)
        want16 = oracle.decode_attention(q.to(dt).float(), kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu())
        assert o16.dtype == dt
        tol = 2e-3 if dt == torch.float16 else 8e-3
        torch.testing.assert_close(o16.cpu().float(), want16, rtol=tol, atol=tol)


def test_matches_dequantize_then_attend(kvq):
    """End-to-end equivalence with what the reference loop does on the GPU: dequantize the whole cache to
    fp16 (get_kv) and run fp32 eager attention on it."""
    B, Hq, H, T, D = 2, 12, 12, 700, 64
    k, v = torch.randn(B, H, T, D, device="cuda", dtype=torch.float16), torch.randn(B, H, T, D, device="cuda", dtype=torch.float16)
    q = torch.randn(B, Hq, 1, D, device="cuda", dtype=torch.float16)
    layer = kvq.QuantizedLayerKV("mixed", "cuda", torch.float16)
    layer.append_tokens(k, v)
    kd, vd = layer.get_kv()
    w = torch.softmax(torch.matmul(q.float(), kd.float().transpose(-1, -2)) * D ** -0.5, dim=-1)
    want = torch.matmul(w, vd.float())
    kc, ks, kb, vc, vs, vb = layer.quantized()
    out = kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb, out_dtype=torch.float32)
    # get_kv rounds code*scale to fp16 before attention; the fused kernel keeps fp32: 2^-11 relative per element
    torch.testing.assert_close(out, want, rtol=2e-3, atol=2e-3)


def test_bad_arguments(kvq):
    q = torch.zeros(1, 2, 1, 600, device="cuda")
    kc = torch.zeros(1, 2, 4, 600, dtype=torch.int8, device="cuda")
    s = torch.ones(4, device="cuda")
    with pytest.raises(RuntimeError, match="head_dim"):  # 1 .. 512
        kvq.decode_attention_quantized(q, kc, s, 8, kc, s, 8)
    q = torch.zeros(1, 3, 1, 64, device="cuda")
    kc = torch.zeros(1, 2, 4, 64, dtype=torch.int8, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of H"):  # (3 query heads on ONE KV head would run, zero-padded to 4)
        kvq.decode_attention_quantized(q, kc, s, 8, kc, s, 8)


def test_exact_current_token_merge(kvq, oracle):
    """kvq_decode_attn_ex: cached tokens from codes + the step's own token unquantized (what the reference's decode step
    attends to, benchmarker.py:473-477).  Expected = fp32 attention over [dequant(cached) ++ exact current]."""
    for (B, Hq, H, T, D, mode) in [(2, 4, 4, 300, 64, "mixed"), (1, 8, 2, 257, 128, "int8"), (2, 12, 12, 130, 64, "int4"), (1, 2, 2, 1, 128, "mixed")]:
        g = torch.Generator().manual_seed(T)
        k, v = torch.randn(B, H, T, D, generator=g), torch.randn(B, H, T, D, generator=g)
        q = torch.randn(B, Hq, 1, D, generator=g)
        kcur, vcur = torch.randn(B, H, 1, D, generator=g) * 2, torch.randn(B, H, 1, D, generator=g)
        layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
        layer.append_tokens(k.cuda(), v.cuda())
        kc, ks, kb, vc, vs, vb = layer.quantized()
        kd = oracle.dequantize_per_token(kc.cpu(), kb, ks.cpu(), D, torch.float32)
        vd = oracle.dequantize_per_token(vc.cpu(), vb, vs.cpu(), D, torch.float32)
        G = Hq // H
        kk = torch.cat([kd, kcur], 2).repeat_interleave(G, 1).double()
        vv = torch.cat([vd, vcur], 2).repeat_interleave(G, 1).double()
        w = torch.softmax(torch.matmul(q.double(), kk.transpose(-1, -2)) * D ** -0.5, -1)
        want = torch.matmul(w, vv).float()
        got = kvq.decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb, k_cur=kcur.cuda(), v_cur=vcur.cuda())
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)


def test_attention_on_a_slid_store(kvq, oracle):
    """After a zero-copy sliding window the live region starts at an arbitrary token offset (scales no longer 16-byte
    aligned): the fused kernels must read exactly that region."""
    B, H, T, D = 2, 4, 777, 128
    g = torch.Generator().manual_seed(5)
    k, v, q = torch.randn(B, H, T, D, generator=g), torch.randn(B, H, T, D, generator=g), torch.randn(B, H, 1, D, generator=g)
    layer = kvq.QuantizedLayerKV("mixed", "cuda", torch.float32)
    layer.append_tokens(k.cuda(), v.cuda())
    for w in (701, 333, 130):
        assert layer.evict("sliding_window", window_size=w) == w
        kc, ks, kb, vc, vs, vb = layer.quantized()
        assert kc.shape[2] == w
        ok, oks, _ = oracle.quantize_per_token(k[:, :, T - w:].contiguous(), 8)
        ov, ovs, _ = oracle.quantize_per_token(v[:, :, T - w:].contiguous(), 4)
        assert torch.equal(kc.cpu(), ok) and torch.equal(vs.cpu(), ovs.float())
        want = oracle.decode_attention(q, ok, 8, oks.float(), ov, 4, ovs.float())
        got = kvq.decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb)
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)
    kd, _ = layer.get_kv()
    assert kd.shape == (B, H, 130, D)
    # appending past the end of the buffer compacts the live region and keeps its contents
    before = layer.quantized()[0].clone()
    layer.append_tokens(k[:, :, :400].cuda(), v[:, :, :400].cuda())
    assert len(layer) == 530 and torch.equal(layer.quantized()[0][:, :, :130], before)


@pytest.mark.parametrize("mode", ["int8", "mixed"])
def test_store_attend_is_the_same_kernel_call(kvq, mode):
    """QuantizedLayerKV.attend (pointer arithmetic on the buffers, the per-step path of KvqFusedCache) == decode_attention_quantized
    on the views, bit for bit: full length, a prefix (seq_len), after a sliding window moved the live region, with the step's own token."""
    g = torch.Generator().manual_seed(11)
    B, H, T, D = 2, 2, 700, 128
    k, v = torch.randn(B, H, T, D, generator=g).half().cuda(), torch.randn(B, H, T, D, generator=g).half().cuda()
    q = torch.randn(B, 4 * H, 1, D, generator=g).half().cuda()
    kc1, vc1 = torch.randn(B, H, 1, D, generator=g).half().cuda(), torch.randn(B, H, 1, D, generator=g).half().cuda()
    layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float16)
    layer.append_tokens(k, v)
    for trim in (None, 300):
        if trim:
            layer.evict("sliding_window", window_size=trim)
        kc, ks, kb, vc, vs, vb = layer.quantized()
        n = kc.size(2)
        assert torch.equal(layer.attend(q), kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb))
        assert torch.equal(layer.attend(q, n - 5, kc1, vc1, 0.11, torch.float32),
                           kvq.decode_attention_quantized(q, kc[:, :, : n - 5], ks[: n - 5], kb, vc[:, :, : n - 5], vs[: n - 5], vb,
                                                          sm_scale=0.11, out_dtype=torch.float32, k_cur=kc1, v_cur=vc1))
    with pytest.raises(ValueError):
        layer.attend(q, 10 ** 6)
    with pytest.raises(RuntimeError):
        layer.attend(q[:, :, :, :64])
    with pytest.raises(ValueError):
        kvq.QuantizedLayerKV(mode, "cuda", torch.float16).attend(q)


@pytest.mark.parametrize("variant", ["cuda_core", "imma"])
def test_full_size_against_fp32_and_across_variants(kvq, variant):
    """C5's real length (T = 65 536) on one KV head.  (1) Parity at this size: fp32 dequantize + softmax on the GPU (the oracle's
    formula, torch fp32; 4 accumulations of 64k terms -> compared at rtol 2e-4 of the largest output) for G = 8 and for the
    MHA shape (codes repeated).  (2) The kernel families check each other: GQA, MHA shape, strided rows (register-staged
    fallback / strided tensor map) within accumulation noise; the paged instantiation bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(5)
    B, G, T, D = 1, 8, 65536, 128
    kc = torch.randint(-127, 128, (B, 1, T, D), device="cuda", dtype=torch.int8, generator=g)
    vc = torch.randint(0, 256, (B, 1, T, D // 2), device="cuda", dtype=torch.uint8, generator=g)
    ks = torch.rand(T, device="cuda", generator=g) * 0.004 + 0.002
    vs = torch.rand(T, device="cuda", generator=g) * 0.01 + 0.01
    q = torch.randn(B, G, 1, D, device="cuda", generator=g)
    kd = kc[0, 0].double() * ks.double()[:, None]
    vd = torch.stack([(vc[0, 0] >> 4).double() - 8.0, (vc[0, 0] & 15).double() - 8.0], dim=-1).reshape(T, D) * vs.double()[:, None]
    p = torch.softmax(q[0, :, 0].double() @ kd.T * D ** -0.5, dim=-1)
    want = (p @ vd).float()[None, :, None, :]
    gqa = kvq.decode_attention_quantized(q, kc, ks, 8, vc, vs, 4, variant=variant)
    torch.testing.assert_close(gqa, want, rtol=2e-4, atol=2e-4 * float(want.abs().max()))
    mha = kvq.decode_attention_quantized(q, kc.expand(B, G, T, D).contiguous(), ks, 8, vc.expand(B, G, T, D // 2).contiguous(), vs, 4,
                                         variant=variant)
    torch.testing.assert_close(mha, want, rtol=2e-4, atol=2e-4 * float(want.abs().max()))
    torch.testing.assert_close(mha, gqa, rtol=2e-4, atol=2e-5)
    # strided rows (a [B,H,T,2W] buffer viewed at half width): register-staged fp32 kernel (cuda_core) / strided tensor map (imma)
    kpad = torch.zeros(B, 1, T, 2 * D, device="cuda", dtype=torch.int8)
    kpad[..., :D] = kc
    vpad = torch.zeros(B, 1, T, D, device="cuda", dtype=torch.uint8)
    vpad[..., : D // 2] = vc
    split = kvq.decode_attention_quantized(q, kpad[..., :D], ks, 8, vpad[..., : D // 2], vs, 4, variant=variant)
    torch.testing.assert_close(split, gqa, rtol=2e-4, atol=2e-5)
    # paged: same tokens in shuffled 128-token blocks
    bs = 128
    nb = T // bs
    perm = torch.randperm(nb, device="cuda", generator=g)
    table = torch.empty(nb, dtype=torch.int32, device="cuda")
    table[:] = perm.to(torch.int32)
    k_pool = torch.empty(nb, B, 1, bs, D, device="cuda", dtype=torch.int8)
    v_pool = torch.empty(nb, B, 1, bs, D // 2, device="cuda", dtype=torch.uint8)
    k_pool[perm] = kc.view(B, 1, nb, bs, D).permute(2, 0, 1, 3, 4)
    v_pool[perm] = vc.view(B, 1, nb, bs, D // 2).permute(2, 0, 1, 3, 4)
    ks_pool = torch.empty(nb, bs, device="cuda")
    vs_pool = torch.empty(nb, bs, device="cuda")
    ks_pool[perm] = ks.view(nb, bs)
    vs_pool[perm] = vs.view(nb, bs)
    paged = kvq.decode_attention_paged(q, k_pool, ks_pool, 8, v_pool, vs_pool, 4, table, T, variant=variant)
    assert torch.equal(paged, gqa)
    # weights sum to one: with V == const the output is that constant (exact up to the scale's rounding), for every variant
    vc1 = torch.full_like(vc, 0xBB)  # nibble 11 -> code +3
    one = kvq.decode_attention_quantized(q, kc, ks, 8, vc1, torch.full_like(vs, 0.5), 4, variant=variant)
    torch.testing.assert_close(one, torch.full_like(one, 1.5), rtol=1e-5, atol=0)


@pytest.mark.parametrize("variant", ["cuda_core", "imma"])
@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
@pytest.mark.parametrize("bs", [16, 64, 128, 256])
def test_paged_variants_equal_contiguous(kvq, variant, mode, bs):
    """Each kernel family over a paged store (every block-size class: several boxes per tile, one box, a piece of a large
    block) is bit-identical to the same family over the contiguous store, also for a length that ends inside a block."""
    g = torch.Generator().manual_seed(bs)
    B, H, G, T, D = 2, 2, 4, 1000, 128
    k, v = torch.randn(B, H, T, D, generator=g), torch.randn(B, H, T, D, generator=g)
    q = torch.randn(B, H * G, 1, D, generator=g).cuda()
    layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
    layer.append_tokens(k.cuda(), v.cuda())
    kc, ks, kb, vc, vs, vb = layer.quantized()
    paged = kvq.PagedQuantizedLayerKV(mode, block_size=bs, compute_dtype=torch.float32)
    paged.append_tokens(k.cuda(), v.cuda())
    pk, pv = paged._k, paged._v
    for n in (T, 777):
        want = kvq.decode_attention_quantized(q, kc[:, :, :n], ks[:n], kb, vc[:, :, :n], vs[:n], vb, variant=variant)
        got = kvq.decode_attention_paged(q, pk.codes, pk.scales_f32, kb, pv.codes, pv.scales_f32, vb, pk.table, n, variant=variant)
        assert torch.equal(got, want), (variant, mode, bs, n)


@pytest.mark.parametrize("G", [3, 5, 6, 7, 12, 16])
def test_any_group_size(kvq, oracle, G):
    """Groupings the kernels are not built for (Qwen2-7B: 7 query heads per KV head) run as zero-padded chunks of <= 8."""
    g = torch.Generator().manual_seed(G)
    B, H, T, D = 2, 2, 333, 64
    k, v = torch.randn(B, H, T, D, generator=g), torch.randn(B, H, T, D, generator=g)
    q = torch.randn(B, H * G, 1, D, generator=g)
    kcur, vcur = torch.randn(B, H, 1, D, generator=g), torch.randn(B, H, 1, D, generator=g)
    layer = kvq.QuantizedLayerKV("mixed", "cuda", torch.float32)
    layer.append_tokens(k.cuda(), v.cuda())
    kc, ks, kb, vc, vs, vb = layer.quantized()
    want = oracle.decode_attention(q, kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu())
    got = kvq.decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)
    assert torch.equal(layer.attend(q.cuda()), got)
    paged = kvq.PagedQuantizedLayerKV("mixed", block_size=32, compute_dtype=torch.float32)
    paged.append_tokens(k.cuda(), v.cuda())
    torch.testing.assert_close(paged.attend(q.cuda()).cpu(), want, rtol=1e-4, atol=1e-5)
    # with the step's own token: equals attention over T + 1 tokens whose last one is exact -> compare the two wrappers
    a = kvq.decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb, k_cur=kcur.cuda(), v_cur=vcur.cuda())
    b = layer.attend(q.cuda(), None, kcur.cuda(), vcur.cuda())
    assert torch.equal(a, b)


@pytest.mark.parametrize("variant", ["cuda_core", "imma"])
@pytest.mark.parametrize("G,mode", [(8, "int8"), (8, "mixed"), (1, "int8")])
def test_back_to_back_calls_on_a_warm_l2(kvq, variant, G, mode):
    """Regression for a cross-proxy race found on B200: the ring's stages are read through the generic proxy (ldmatrix /
    ld.shared) and refilled through the async proxy (TMA / bulk copy).  When the tiles come from a warm L2 -- the same cache
    attended to again and again, as in a decode loop -- the refill could overtake reads that had been issued but not performed
    (rare wrong V rows: l exact, accumulators off by 1e-2 .. 1e-1, two runs in three at G = 8).  Fixed by
    fence.proxy.async before the stage is released.  40 identical calls must give bit-identical, correct results."""
    g = torch.Generator(device="cuda").manual_seed(0)
    B, H, T, D = 1, 32, 3000, 128
    k = torch.randn(B, H, T, D, device="cuda", generator=g)
    v = torch.randn(B, H, T, D, device="cuda", generator=g)
    q = torch.randn(B, H * G, 1, D, device="cuda", generator=g)
    layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
    layer.append_tokens(k, v)
    kc, ks, kb, vc, vs, vb = layer.quantized()
    kd, vd = layer.get_kv()
    p = torch.softmax(torch.einsum("bhgd,bhtd->bhgt", q[:, :, 0].double().view(B, H, G, D), kd.double()) * D ** -0.5, dim=-1)
    want = torch.einsum("bhgt,bhtd->bhgd", p, vd.double()).reshape(B, H * G, 1, D).float()
    first = None
    for _ in range(40):
        out = kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb, variant=variant)
        if first is None:
            first = out
            torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-5)
        else:
            assert torch.equal(out, first)


@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
@pytest.mark.parametrize("B,Hq,H,T,D", [(2, 6, 3, 300, 96), (1, 4, 4, 700, 256), (2, 8, 2, 130, 80), (1, 5, 1, 65, 7), (1, 2, 2, 1, 33),
                                         (1, 24, 2, 200, 64)])
def test_any_head_dim(kvq, oracle, mode, B, Hq, H, T, D):
    """Models whose head_dim is not 64 / 128 (96: Phi-3-mini class, 256: Gemma, odd sizes: the reference accepts any model,
    benchmarker.py:473-477) and groupings beyond 8 run the generic kernel: same tolerance as the tiled ones, with and without
    the step's own token, through the function, the store and the HF cache path."""
    g = torch.Generator().manual_seed(D * 7 + T)
    k, v = torch.randn(B, H, T, D, generator=g), torch.randn(B, H, T, D, generator=g)
    q = torch.randn(B, Hq, 1, D, generator=g)
    layer = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
    layer.append_tokens(k.cuda(), v.cuda())
    kc, ks, kb, vc, vs, vb = layer.quantized()
    want = oracle.decode_attention(q, kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu())
    import kvq_b200._lib as L
    from kvq_b200.attention import decode_attention_quantized
    got = decode_attention_quantized(q.cuda(), kc, ks, kb, vc, vs, vb) if Hq // H <= 8 or D not in (64, 128) else None
    if got is not None:
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.attend(q.cuda()).cpu(), want, rtol=1e-4, atol=1e-5)
    o16 = layer.attend(q.cuda().half())
    assert o16.dtype == torch.float16
    torch.testing.assert_close(o16.cpu().float(), oracle.decode_attention(q.half().float(), kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu()),
                               rtol=2e-3, atol=2e-3)
