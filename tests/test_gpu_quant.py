"""Area 1 parity on the GPU, through the C ABI: codes / packed nibbles / scales bit-exact, dequantized
values bit-exact (fp32 multiply + one RN conversion), against (a) the golden fixtures produced by the
reference, (b) the CPU oracle on seeded inputs incl. strided views and the HF transposed layout, and
(c) size-independent properties at the full BASELINE C3 tensor size."""

import numpy as np
import pytest
import torch

from _golden import DT, TAG, bits_equal, load, tensor, typed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kvq():
    import kvq_b200
    assert torch.cuda.is_available()
    return kvq_b200


def gen(shape, dtype, seed, outlier=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g, dtype=torch.float32)
    if outlier:
        x.view(-1)[::97] *= 8
    return x.to(dtype)


def test_golden_per_tensor(kvq):
    g = load("quant.npz")
    for case in g["cases"]:
        x = tensor(g, f"{case}/x")
        tag = TAG[x.dtype]
        xc = x.cuda()
        q8, s8 = kvq.quantize_int8_per_tensor(xc)
        assert q8.dtype == torch.int8 and q8.shape == x.shape and s8.dim() == 0 and s8.dtype == x.dtype
        assert torch.equal(q8.cpu(), torch.from_numpy(g[f"{case}/q8"])), case
        assert bits_equal(s8.cpu(), typed(g, f"{case}/s8_{tag}", tag)), case
        p4, s4, last = kvq.quantize_int4_per_tensor_packed(xc)
        assert p4.dtype == torch.uint8 and isinstance(last, int)
        assert torch.equal(p4.cpu(), torch.from_numpy(g[f"{case}/p4"])), case
        assert bits_equal(s4.cpu(), typed(g, f"{case}/s4_{tag}", tag)), case
        assert last == int(g[f"{case}/last"])
        for otag, odt in DT.items():
            assert bits_equal(kvq.dequantize_int8_per_tensor(q8, s8, odt).cpu(), typed(g, f"{case}/dq8_{otag}", otag)), (case, otag)
            assert bits_equal(kvq.dequantize_int4_per_tensor_packed(p4, s4, last, odt).cpu(),
                              typed(g, f"{case}/dq4_{otag}", otag)), (case, otag)


SHAPES = [(2, 3, 5, 7), (1, 1, 1, 16), (2, 4, 9, 64), (3, 8, 130, 128), (1, 12, 300, 64), (2, 2, 17, 96), (1, 2, 3, 272),
          (4, 8, 512, 128)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_per_tensor(kvq, oracle, dtype, shape):
    x = gen(shape, dtype, 1234 + sum(shape), outlier=True)
    q8, s8 = kvq.quantize_int8_per_tensor(x.cuda())
    oq8, os8 = oracle.quantize_int8_per_tensor(x)
    assert torch.equal(q8.cpu(), oq8) and bits_equal(s8.cpu(), os8)
    p4, s4, last = kvq.quantize_int4_per_tensor_packed(x.cuda())
    op4, os4, olast = oracle.quantize_int4_per_tensor_packed(x)
    assert torch.equal(p4.cpu(), op4) and bits_equal(s4.cpu(), os4) and last == olast
    for odt in (torch.float32, torch.float16, torch.bfloat16):
        assert bits_equal(kvq.dequantize_int8_per_tensor(q8, s8, odt).cpu(), oracle.dequantize_int8_per_tensor(oq8, os8, odt))
        assert bits_equal(kvq.dequantize_int4_per_tensor_packed(p4, s4, last, odt).cpu(),
                          oracle.dequantize_int4_per_tensor_packed(op4, os4, olast, odt))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_views_and_transposed_layouts(kvq, oracle, dtype):
    base = gen((2, 40, 6, 64), dtype, 99, outlier=True)          # [B, T, H, D] as HF GPT-2 produces it
    hf = base.cuda().transpose(1, 2)                              # [B, H, T, D] view, strides (T*H*D, D, H*D, 1)
    ref = base.transpose(1, 2).contiguous()
    views = [(hf, ref), (hf[:, :, 7:8, :], ref[:, :, 7:8, :]), (hf[:, :, -16:, :], ref[:, :, -16:, :]),
             (hf[:, 1:3], ref[:, 1:3]), (hf[..., 16:48], ref[..., 16:48])]
    for xv, rv in views:
        assert not xv.is_contiguous()
        q8, s8 = kvq.quantize_int8_per_tensor(xv)
        oq8, os8 = oracle.quantize_int8_per_tensor(rv)
        assert torch.equal(q8.cpu(), oq8) and bits_equal(s8.cpu(), os8)
        p4, s4, _ = kvq.quantize_int4_per_tensor_packed(xv)
        op4, os4, _ = oracle.quantize_int4_per_tensor_packed(rv)
        assert torch.equal(p4.cpu(), op4) and bits_equal(s4.cpu(), os4)
    # innermost stride != 1 (generic path)
    xt = gen((2, 3, 16, 8), dtype, 5).cuda().transpose(2, 3)
    q8, s8 = kvq.quantize_int8_per_tensor(xt)
    oq8, os8 = oracle.quantize_int8_per_tensor(xt.cpu().contiguous())
    assert torch.equal(q8.cpu(), oq8) and bits_equal(s8.cpu(), os8)


def test_edge_inputs(kvq):
    for dt, want in ((torch.float32, 1e-8), (torch.float16, 0.0), (torch.bfloat16, 1.0012e-8)):
        q, s = kvq.quantize_int8_per_tensor(torch.zeros(4, dtype=dt, device="cuda"))
        assert q.tolist() == [0, 0, 0, 0] and float(s) == pytest.approx(want, rel=1e-3, abs=0)
    q, s = kvq.quantize_int8_per_tensor(torch.tensor(3.0, device="cuda"))  # 0-dim input
    assert q.shape == () and int(q) == 127
    with pytest.raises(RuntimeError):
        kvq.quantize_int8_per_tensor(torch.empty(0, device="cuda"))
    with pytest.raises(RuntimeError, match="int8"):
        kvq.dequantize_int8_per_tensor(torch.zeros(4, dtype=torch.uint8, device="cuda"), torch.tensor(1.0, device="cuda"), torch.float16)
    with pytest.raises(RuntimeError, match="uint8"):
        kvq.dequantize_int4_per_tensor_packed(torch.zeros(4, dtype=torch.int8, device="cuda"), torch.tensor(1.0, device="cuda"), 8, torch.float16)
    # python-float scale is accepted like a 0-dim tensor
    out = kvq.dequantize_int8_per_tensor(torch.tensor([1, -2], dtype=torch.int8, device="cuda"), 0.5, torch.float32)
    assert out.tolist() == [0.5, -1.0]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_dequant_packed_path_corner_cases(kvq, oracle, dtype):
    """The 16-bit dequant path multiplies code * scale in half precision (exact operands, one rounding).
    Sweep every code against scales that hit subnormal / tiny / huge products and mismatched scale dtypes."""
    codes8 = torch.arange(-127, 128, dtype=torch.int8).repeat(2)[:496].reshape(1, 1, 31, 16)
    nib = torch.arange(0, 256, dtype=torch.uint8).repeat(2)[:496].reshape(1, 1, 31, 16)
    for sval in (0.0, 6e-8, 1.2e-7, 3.1e-5, 0.007, 0.33, 1.0, 3.0, 511.0, 6.5e4 / 127):
        for sdt in (dtype, torch.float32):
            s = torch.tensor(sval, dtype=sdt)
            for odt in (dtype, torch.float32):
                got = kvq.dequantize_int8_per_tensor(codes8.cuda(), s.cuda(), odt)
                assert bits_equal(got.cpu(), oracle.dequantize_int8_per_tensor(codes8, s, odt)), (sval, sdt, odt)
                got = kvq.dequantize_int4_per_tensor_packed(nib.cuda(), s.cuda(), 32, odt)
                assert bits_equal(got.cpu(), oracle.dequantize_int4_per_tensor_packed(nib, s, 32, odt)), (sval, sdt, odt)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("bits", [8, 4])
def test_batched_multi_tensor_matches_single(kvq, oracle, dtype, bits):
    """One launch per phase over all tensors of a step == the per-tensor helpers applied one by one (bit-exact)."""
    from kvq_b200.quantization import TensorTable, dequantize_per_tensor_batched, quantize_per_tensor_batched
    xs = [gen((2, 4, 37, 64), dtype, 50 + i, outlier=(i % 2 == 0)) * (0.1 + i) for i in range(5)]
    xs[3] = torch.zeros_like(xs[3])  # eps path inside a batch
    xd = [x.to(dtype).cuda() for x in xs]
    W = 64 if bits == 8 else 32
    codes = [torch.empty((2, 4, 37, W), dtype=torch.int8 if bits == 8 else torch.uint8, device="cuda") for _ in xs]
    outs = [torch.empty((2, 4, 37, 64), dtype=dtype, device="cuda") for _ in xs]
    scales = torch.empty(len(xs), dtype=dtype, device="cuda")
    tx, tc, to = TensorTable(xd), TensorTable(codes), TensorTable(outs)
    quantize_per_tensor_batched(tx, tc, bits, scales)
    dequantize_per_tensor_batched(tc, bits, scales, to)
    for i, x in enumerate(xs):
        x = x.to(dtype)
        if bits == 8:
            oq, os_ = oracle.quantize_int8_per_tensor(x)
            want = oracle.dequantize_int8_per_tensor(oq, os_, dtype)
        else:
            oq, os_, last = oracle.quantize_int4_per_tensor_packed(x)
            want = oracle.dequantize_int4_per_tensor_packed(oq, os_, last, dtype)
        assert torch.equal(codes[i].cpu(), oq) and bits_equal(scales[i].cpu(), os_) and bits_equal(outs[i].cpu(), want), i


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_calibrated_absmax_clips_like_the_reference_formula(kvq, oracle, dtype):
    """C ABI called with an absmax SMALLER than max|x| (no KVQ_ABSMAX_COVERS_INPUT flag): clamp(round(x/scale), lo, hi)
    must clip exactly like the reference formula (ops.py:29, :50) evaluated with that scale."""
    from kvq_b200 import _lib as L
    x = gen((2, 4, 33, 64), dtype, 77, outlier=True)
    xc = x.cuda()
    half_max = float(x.float().abs().max()) * 0.37
    bits = torch.tensor([half_max], dtype=torch.float32).view(torch.int32).cuda()
    half_max = float(bits.view(torch.float32)[0])
    sh, st = L.shape_stride(xc)
    q8 = torch.empty(x.shape, dtype=torch.int8, device="cuda")
    L.check(L.lib().kvq_quant_int8(xc.data_ptr(), L.dtype_code(dtype), sh, st, L.PER_TENSOR, bits.data_ptr(), 1e-8, q8.data_ptr(), None,
                                   None, None, L.stream_ptr(xc.device)), "kvq_quant_int8")
    want8, _ = oracle.quantize_int8_per_tensor(x, absmax_override=half_max)
    assert torch.equal(q8.cpu(), want8) and int(want8.max()) == 127 and int(want8.min()) == -127
    p4 = torch.empty((2, 4, 33, 32), dtype=torch.uint8, device="cuda")
    L.check(L.lib().kvq_quant_int4_pack(xc.data_ptr(), L.dtype_code(dtype), sh, st, L.PER_TENSOR, bits.data_ptr(), 1e-8, p4.data_ptr(),
                                        None, None, None, L.stream_ptr(xc.device)), "kvq_quant_int4_pack")
    want4, _, _ = oracle.quantize_int4_per_tensor_packed(x, absmax_override=half_max)
    assert torch.equal(p4.cpu(), want4)
    # strided view -> rows kernel, same contract
    xv = xc.transpose(1, 2).contiguous().transpose(1, 2)
    shv, stv = L.shape_stride(xv)
    L.check(L.lib().kvq_quant_int8(xv.data_ptr(), L.dtype_code(dtype), shv, stv, L.PER_TENSOR, bits.data_ptr(), 1e-8, q8.data_ptr(), None,
                                   None, None, L.stream_ptr(xc.device)), "kvq_quant_int8")
    assert torch.equal(q8.cpu(), want8)


def test_kvq_ext_facade(kvq, oracle):
    ext = kvq.get_cuda_extension()
    q = torch.randint(-127, 128, (3, 5, 16), dtype=torch.int8)
    out = ext.dequant_int8_to_fp16(q.cuda(), 0.0123)
    assert bits_equal(out.cpu(), oracle.dequantize_int8_per_tensor(q, torch.tensor(0.0123), torch.float16))
    p = torch.randint(0, 256, (2, 4, 4), dtype=torch.uint8)
    out = ext.dequant_int4_packed_to_fp16(p.cuda(), 0.5, 7)
    want = oracle.dequantize_int4_per_tensor_packed(p, torch.tensor(0.5), 8, torch.float16)
    want[..., 7:] = 0
    assert out.shape == (2, 4, 8) and bits_equal(out.cpu(), want)
    with pytest.raises(RuntimeError, match="contiguous"):
        ext.dequant_int8_to_fp16(q.cuda().transpose(0, 1), 1.0)


def test_golden_per_token_store(kvq):
    g = load("pertoken.npz")
    for case in g["cases"]:
        mode, tag = case.split("_")[0], case.split("_")[1]
        k, v = tensor(g, f"{case}/k").cuda(), tensor(g, f"{case}/v").cuda()
        T = k.shape[2]
        for bulk in (False, True):
            layer = kvq.QuantizedLayerKV(mode=mode, device="cuda", compute_dtype=DT[tag])
            if bulk:
                layer.append_tokens(k, v)
            else:
                for t in range(T):
                    layer.append(k[:, :, t:t + 1, :], v[:, :, t:t + 1, :])
            assert len(layer.k_store) == T and len(layer.k_scales) == T
            assert torch.equal(torch.cat(layer.k_store, dim=2).cpu(), torch.from_numpy(g[f"{case}/kstore"])), case
            assert torch.equal(torch.cat(layer.v_store, dim=2).cpu(), torch.from_numpy(g[f"{case}/vstore"])), case
            assert bits_equal(torch.stack(layer.k_scales).cpu(), typed(g, f"{case}/kscales_{tag}", tag)), case
            assert bits_equal(torch.stack(layer.v_scales).cpu(), typed(g, f"{case}/vscales_{tag}", tag)), case
            kd, vd = layer.get_kv()
            assert bits_equal(kd.cpu(), typed(g, f"{case}/kdq_{tag}", tag)), case
            assert bits_equal(vd.cpu(), typed(g, f"{case}/vdq_{tag}", tag)), case
            assert layer.estimated_bytes() == int(g[f"{case}/bytes"]), case


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(16, 12, 70, 64), (4, 8, 300, 128), (1, 2, 9, 7), (2, 40, 33, 128), (64, 32, 3, 128)])
def test_oracle_per_token(kvq, oracle, dtype, shape):
    """C2-like (B16,H12,D64) and Llama-like shapes; (64,32,3,128) exceeds the register-resident token
    size and takes the two-loop variant; (1,2,9,7) the generic path."""
    k = gen(shape, dtype, 777, outlier=True)
    hf_layout = k.cuda().transpose(1, 2).contiguous().transpose(1, 2)  # [B,H,T,D] view of [B,T,H,D] memory
    for mode in ("int8", "int4", "mixed"):
        kb = 4 if mode == "int4" else 8
        vb = 8 if mode == "int8" else 4
        layer = kvq.QuantizedLayerKV(mode=mode, device="cuda", compute_dtype=dtype)
        layer.append_tokens(hf_layout, k.cuda())
        oq, os_, _ = oracle.quantize_per_token(k, kb)
        ov, ovs, _ = oracle.quantize_per_token(k, vb)
        kc, ks32, kbits, vc, vs32, vbits = layer.quantized()
        assert (kbits, vbits) == (kb, vb)
        assert torch.equal(kc.cpu(), oq) and torch.equal(vc.cpu(), ov)
        assert bits_equal(layer._k.scales[: shape[2]].cpu(), os_) and bits_equal(layer._v.scales[: shape[2]].cpu(), ovs)
        assert torch.equal(ks32.cpu(), os_.float()) and torch.equal(vs32.cpu(), ovs.float())
        kd, vd = layer.get_kv()
        assert bits_equal(kd.cpu(), oracle.dequantize_per_token(oq, kb, os_, shape[3], dtype))
        assert bits_equal(vd.cpu(), oracle.dequantize_per_token(ov, vb, ovs, shape[3], dtype))


def test_store_growth_and_block_append(kvq, oracle):
    """Capacity doubling keeps earlier tokens; append() with n > 1 tokens uses ONE scale for the block,
    like the reference's per-tensor helper would (ops.py:181-210)."""
    x = gen((2, 4, 700, 64), torch.float16, 3)
    layer = kvq.QuantizedLayerKV("int8", "cuda", torch.float16)
    layer.append_tokens(x[:, :, :200].cuda(), x[:, :, :200].cuda())
    layer.append(x[:, :, 200:500].cuda(), x[:, :, 200:500].cuda())      # block of 300: shared scale, grows past 256/512
    layer.append_tokens(x[:, :, 500:].cuda(), x[:, :, 500:].cuda())
    kd, _ = layer.get_kv()
    a, sa, _ = oracle.quantize_per_token(x[:, :, :200].contiguous(), 8)
    b, sb = oracle.quantize_int8_per_tensor(x[:, :, 200:500].contiguous())
    c, sc, _ = oracle.quantize_per_token(x[:, :, 500:].contiguous(), 8)
    want = torch.cat([oracle.dequantize_per_token(a, 8, sa, 64, torch.float16), oracle.dequantize_int8_per_tensor(b, sb, torch.float16),
                      oracle.dequantize_per_token(c, 8, sc, 64, torch.float16)], dim=2)
    assert bits_equal(kd.cpu(), want)
    assert layer.estimated_bytes() == 2 * (x.numel() + (200 + 1 + 200) * 2)


def test_quantized_cache_container(kvq, oracle):
    L_, B, H, T, D = 3, 2, 4, 20, 64
    past = tuple((gen((B, H, T, D), torch.float16, 10 + i).cuda(), gen((B, H, T, D), torch.float16, 20 + i).cuda()) for i in range(L_))
    qc = kvq.QuantizedKVCache(n_layers=L_, mode="mixed", device="cuda", compute_dtype=torch.float16)
    qc.init_from_prompt_past(past)
    step = tuple((torch.cat([k, gen((B, H, 1, D), torch.float16, 30 + i).cuda()], 2), torch.cat([v, gen((B, H, 1, D), torch.float16, 40 + i).cuda()], 2))
                 for i, (k, v) in enumerate(past))
    qc.append_from_past(step)
    out = qc.to_past_key_values()
    assert len(out) == L_ and out[0][0].shape == (B, H, T + 1, D)
    for i, (k, v) in enumerate(step):
        oq, os_, _ = oracle.quantize_per_token(k.cpu(), 8)
        ov, ovs, _ = oracle.quantize_per_token(v.cpu(), 4)
        assert bits_equal(out[i][0].cpu(), oracle.dequantize_per_token(oq, 8, os_, D, torch.float16))
        assert bits_equal(out[i][1].cpu(), oracle.dequantize_per_token(ov, 4, ovs, D, torch.float16))
    assert qc.estimated_bytes() == L_ * (B * H * (T + 1) * (D + D // 2) + 2 * (T + 1) * 2)


def test_full_size_properties(kvq):
    """BASELINE C3: one Llama-3-8B-shape layer tensor, B16 x H8 x T32768 x D128 bf16 = 1 GiB.
    Size-independent checks: (1) scale == max|x|/7 computed by torch on the device; (2) codes of a
    random 1M-element sample equal the exact formula evaluated in float64; (3) |x - dequant| <= scale/2
    everywhere inside the clamp range; (4) quantize(dequantize(q)) == q (idempotence); (5) int8 and int4
    codes are consistent with the same absmax; (6) a checksum of the packed bytes equals the checksum of
    the same tensor quantized in 8 head shards with the global absmax (linearity over heads)."""
    torch.manual_seed(1234)
    B, H, T, D = 16, 8, 32768, 128
    x = torch.randn(B, H, T, D, device="cuda", dtype=torch.bfloat16)
    absmax = x.abs().max().float()
    p4, s4, last = kvq.quantize_int4_per_tensor_packed(x)
    assert last == D and p4.shape == (B, H, T, D // 2)
    # scale on the CPU: CUDA torch turns tensor / python-scalar into a multiply by the reciprocal
    scale = torch.clamp(absmax.cpu() / 7.0, min=1e-8).cuda()
    assert float(s4) == float(scale.to(torch.bfloat16))
    idx = torch.randint(0, x.numel(), (1 << 20,), device="cuda")
    xs = x.view(-1)[idx].double()
    codes = torch.clamp(torch.round((xs.float() / scale)), -8, 7).to(torch.int64)
    byte = p4.view(-1)[idx // 2].to(torch.int64)
    nib = torch.where(idx % 2 == 0, byte >> 4, byte & 15) - 8
    assert torch.equal(nib, codes)
    dq = kvq.dequantize_int4_per_tensor_packed(p4, s4, last, torch.bfloat16)
    err = (dq.float() - x.float()).abs().max()
    assert float(err) <= float(s4.float()) * 0.5 * (1 + 2 ** -7) + float(absmax) * 2 ** -8
    p4b, s4b, _ = kvq.quantize_int4_per_tensor_packed(dq)
    # dq's absmax is 7*s4 (s4 rounded to bf16), so its scale differs slightly from the first pass; codes must not
    assert torch.equal(p4b, p4)
    del dq, p4b
    q8, s8 = kvq.quantize_int8_per_tensor(x)
    assert float(s8) == float(torch.clamp(absmax.cpu() / 127.0, min=1e-8).to(torch.bfloat16))
    assert int(q8.abs().max()) == 127
    del q8
    total = int(p4.view(-1).to(torch.int64).sum())
    from kvq_b200 import _lib as L
    from kvq_b200.quantization.ops import _absmax_bits
    bits = _absmax_bits(x, L.PER_TENSOR)
    shard_sum = 0
    for h in range(H):
        xs = x[:, h:h + 1]
        out = torch.empty((B, 1, T, D // 2), dtype=torch.uint8, device="cuda")
        sh, st = L.shape_stride(xs)
        L.check(L.lib().kvq_quant_int4_pack(xs.data_ptr(), L.BF16, sh, st, L.PER_TENSOR, bits.data_ptr(), 1e-8, out.data_ptr(),
                                            None, None, None, L.stream_ptr(x.device)), "shard")
        assert torch.equal(out, p4[:, h:h + 1])
        shard_sum += int(out.view(-1).to(torch.int64).sum())
    assert shard_sum == total
