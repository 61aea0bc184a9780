"""Paged quantized store + attention through the block table: same codes / scales as the contiguous
QuantizedLayerKV, attention bit-identical to the contiguous kernel on the same tokens and within tolerance of the oracle."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kvq():
    import kvq_b200
    return kvq_b200


def _kv(B, H, T, D, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    k = torch.randn(B, H, T, D, generator=g)
    k.view(-1)[::89] *= 6
    v = torch.randn(B, H, T, D, generator=g)
    return k.to(dtype).cuda(), v.to(dtype).cuda()


@pytest.mark.parametrize("mode", ["int8", "int4", "mixed"])
@pytest.mark.parametrize("block_size", [16, 64, 128, 256])
def test_store_matches_contiguous_store(kvq, mode, block_size):
    """Prompt in one go, then ragged appends that straddle block boundaries: codes, scales and get_kv bit-identical."""
    B, H, T, D = 2, 3, 300, 64
    k, v = _kv(B, H, T, D, 7, torch.float16)
    flat = kvq.QuantizedLayerKV(mode, "cuda", torch.float16)
    paged = kvq.PagedQuantizedLayerKV(mode, block_size=block_size, compute_dtype=torch.float16)
    cuts = [0, 131, 132, 133, 200, 260, 300]
    for a, b in zip(cuts[:-1], cuts[1:]):
        flat.append_tokens(k[:, :, a:b], v[:, :, a:b])
        paged.append_tokens(k[:, :, a:b], v[:, :, a:b])
    assert len(paged) == len(flat) == T and paged.num_blocks() == -(-T // block_size)
    kc, ks, _, vc, vs, _ = flat.quantized()
    pk, pks, pv, pvs = paged.quantized_contiguous()
    assert torch.equal(pk, kc) and torch.equal(pv, vc)
    assert torch.equal(pks.float(), ks) and torch.equal(pvs.float(), vs)
    a, b = flat.get_kv(), paged.get_kv()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    table = paged.block_table.cpu().tolist()
    assert len(set(table)) == len(table) and table != sorted(table)  # real indirection, not the identity


@pytest.mark.parametrize("mode", ["int8", "mixed", "int4"])
@pytest.mark.parametrize("B,Hq,H,T,D,bs", [(2, 4, 4, 300, 64, 16), (1, 8, 8, 1000, 128, 64), (2, 8, 2, 513, 128, 128),
                                             (1, 16, 2, 2000, 64, 32), (1, 8, 1, 5000, 128, 256), (2, 4, 1, 1, 128, 16),
                                             (1, 32, 4, 20000, 128, 64), (2, 8, 4, 777, 128, 32), (1, 24, 12, 1500, 64, 64)])
def test_paged_attention_equals_contiguous_and_oracle(kvq, oracle, mode, B, Hq, H, T, D, bs):
    k, v = _kv(B, H, T, D, B * 1000 + T)
    q = torch.randn(B, Hq, 1, D, generator=torch.Generator().manual_seed(T)).cuda()
    flat = kvq.QuantizedLayerKV(mode, "cuda", torch.float32)
    flat.append_tokens(k, v)
    paged = kvq.PagedQuantizedLayerKV(mode, block_size=bs, compute_dtype=torch.float32)
    paged.append_tokens(k, v)
    kc, ks, kb, vc, vs, vb = flat.quantized()
    want = kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb)
    got = paged.attend(q)
    assert torch.equal(got, want)  # same tiles, same order, same arithmetic
    ref = oracle.decode_attention(q.cpu(), kc.cpu(), kb, ks.cpu(), vc.cpu(), vb, vs.cpu())
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-5)
    # the step's own token, exact
    kcur, vcur = _kv(B, H, 1, D, 5)
    assert torch.equal(paged.attend(q, k_cur=kcur, v_cur=vcur),
                       kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb, k_cur=kcur, v_cur=vcur))
    for dt in (torch.float16, torch.bfloat16):
        assert torch.equal(paged.attend(q.to(dt)), kvq.decode_attention_quantized(q.to(dt), kc, ks, kb, vc, vs, vb))


def test_sliding_window_is_a_table_edit(kvq):
    B, H, T, D, bs = 1, 2, 1000, 128, 64
    k, v = _kv(B, H, T, D, 3, torch.float16)
    q = torch.randn(B, 8 * H, 1, D, generator=torch.Generator().manual_seed(1)).half().cuda()
    paged = kvq.PagedQuantizedLayerKV("mixed", block_size=bs)
    paged.append_tokens(k, v)
    pool_ptr = paged._k.codes.data_ptr()
    n = paged.evict_sliding_window(256)
    kept = T - ((T - 256) // bs) * bs  # whole leading blocks dropped: at least the window survives
    assert n == kept == len(paged) and 256 <= kept < 256 + bs and paged._k.codes.data_ptr() == pool_ptr
    flat = kvq.QuantizedLayerKV("mixed", "cuda", torch.float16)
    flat.append_tokens(k[:, :, T - kept:], v[:, :, T - kept:])
    kc, ks, kb, vc, vs, vb = flat.quantized()
    assert torch.equal(paged.attend(q), kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb))
    # freed blocks are reused by later appends; the result keeps matching
    k2, v2 = _kv(B, H, 200, D, 4, torch.float16)
    paged.append_tokens(k2, v2)
    flat.append_tokens(k2, v2)
    kc, ks, kb, vc, vs, vb = flat.quantized()
    assert torch.equal(paged.attend(q), kvq.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb))
    assert paged._k.codes.data_ptr() == pool_ptr


def test_bad_arguments(kvq):
    with pytest.raises(ValueError):
        kvq.PagedQuantizedLayerKV("int8", block_size=48)
    p = kvq.PagedQuantizedLayerKV("int8", block_size=16)
    with pytest.raises(ValueError):
        p.attend(torch.zeros(1, 2, 1, 64, device="cuda"))
    k, v = _kv(1, 2, 40, 64, 1, torch.float16)
    p.append_tokens(k, v)
    q2 = torch.randn(1, 4, 1, 64, device="cuda", dtype=torch.float16)  # two query heads per KV head
    flat = kvq.QuantizedLayerKV("int8", "cuda", torch.float16)
    flat.append_tokens(k, v)
    assert torch.equal(p.attend(q2), flat.attend(q2))
    with pytest.raises(RuntimeError):
        p.append_tokens(k[:, :1], v[:, :1])
