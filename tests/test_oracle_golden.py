"""Pins the CPU oracle (oracle/kvq_oracle.c) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer / byte / index results bit-exact; dequantized values
bit-exact; pooled summaries and attention within the stated tolerance."""

import numpy as np
import pytest
import torch

from _golden import DT, bits_equal, load, tensor, typed


def test_quant_dequant_per_tensor(oracle):
    g = load("quant.npz")
    for case in g["cases"]:
        x = tensor(g, f"{case}/x")
        tag = {torch.float32: "f32", torch.float16: "f16", torch.bfloat16: "bf16"}[x.dtype]
        q8, s8 = oracle.quantize_int8_per_tensor(x)
        assert torch.equal(q8, torch.from_numpy(g[f"{case}/q8"])), case
        assert bits_equal(s8, typed(g, f"{case}/s8_{tag}", tag)), case
        p4, s4, last = oracle.quantize_int4_per_tensor_packed(x)
        assert torch.equal(p4, torch.from_numpy(g[f"{case}/p4"])), case
        assert bits_equal(s4, typed(g, f"{case}/s4_{tag}", tag)), case
        assert last == int(g[f"{case}/last"])
        for otag, odt in DT.items():
            d8 = oracle.dequantize_int8_per_tensor(q8, s8, odt)
            assert bits_equal(d8, typed(g, f"{case}/dq8_{otag}", otag)), (case, otag)
            d4 = oracle.dequantize_int4_per_tensor_packed(p4, s4, last, odt)
            assert bits_equal(d4, typed(g, f"{case}/dq4_{otag}", otag)), (case, otag)


def test_known_answers(oracle):
    """SURVEY section 4 known-answer table (obtained by running the reference)."""
    q, s = oracle.quantize_int8_per_tensor(torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 127.0]))
    assert q.tolist() == [0, 2, 2, 0, -2, 127] and float(s) == 1.0
    p, s, last = oracle.quantize_int4_per_tensor_packed(torch.tensor([0.5, 1.5, 2.5, -2.5, -7.0, 7.0, 6.5]))
    assert bytes(p.tolist()) == bytes.fromhex("8aa61fe8") and last == 7 and tuple(p.shape) == (4,)
    assert oracle.dequantize_int4_per_tensor_packed(p, s, last, torch.float32).tolist() == [0, 2, 2, -2, -7, 7, 6]
    for dt, want in ((torch.float32, 1e-8), (torch.float16, 0.0), (torch.bfloat16, 1.0012e-8)):
        q, s = oracle.quantize_int8_per_tensor(torch.zeros(4, dtype=dt))
        assert q.tolist() == [0, 0, 0, 0]
        assert float(s) == pytest.approx(want, rel=1e-3, abs=0)


def test_per_token_cache(oracle):
    g = load("pertoken.npz")
    for case in g["cases"]:
        mode, tag = case.split("_")[0], case.split("_")[1]
        k, v = tensor(g, f"{case}/k"), tensor(g, f"{case}/v")
        D = k.shape[-1]
        kb = 4 if mode == "int4" else 8
        vb = 8 if mode == "int8" else 4
        kq, ks, _ = oracle.quantize_per_token(k, kb)
        vq, vs, _ = oracle.quantize_per_token(v, vb)
        assert torch.equal(kq, torch.from_numpy(g[f"{case}/kstore"])), case
        assert torch.equal(vq, torch.from_numpy(g[f"{case}/vstore"])), case
        assert bits_equal(ks, typed(g, f"{case}/kscales_{tag}", tag)), case
        assert bits_equal(vs, typed(g, f"{case}/vscales_{tag}", tag)), case
        kd = oracle.dequantize_per_token(kq, kb, ks, D, DT[tag])
        vd = oracle.dequantize_per_token(vq, vb, vs, D, DT[tag])
        assert bits_equal(kd, typed(g, f"{case}/kdq_{tag}", tag)), case
        assert bits_equal(vd, typed(g, f"{case}/vdq_{tag}", tag)), case
        est = kq.numel() + vq.numel() + (ks.numel() + vs.numel()) * ks.element_size()
        assert est == int(g[f"{case}/bytes"]), case


def test_evict_index_sets(oracle):
    g = load("evict.npz")
    params, offs, flat = g["params"], g["offsets"], g["indices"]
    for i, (st, T, w, p, s, blk, keep, bud) in enumerate(params.tolist()):
        want = flat[offs[i]:offs[i + 1]].astype(np.int64)
        got = oracle.evict_indices(st, T, window=w, prefix=p, stride=s, block=blk, keep=keep, budget=bud).numpy()
        assert np.array_equal(got, want), (st, T, w, p, s, blk, keep, bud)


def test_evict_known_answers(oracle):
    T = 40
    assert oracle.evict_indices(0, T, window=8).tolist() == list(range(32, 40))
    assert oracle.evict_indices(1, T, window=8, prefix=4).tolist() == [0, 1, 2, 3] + list(range(32, 40))
    assert oracle.evict_indices(2, T, window=8, prefix=2, stride=4).tolist() == [0, 1] + list(range(2, 32, 4)) + list(range(32, 40))
    assert oracle.evict_indices(3, T, window=8, prefix=3, block=10, keep=2).tolist() == [0, 1, 2, 11, 12, 21, 22, 30, 31] + list(range(32, 40))
    assert oracle.evict_indices(4, T, window=8, prefix=2, budget=5).tolist() == [0, 1, 2, 9, 16, 23, 31] + list(range(32, 40))


def test_evict_gathered_values(oracle):
    g = load("evict.npz")
    for tag in DT:
        k, v = typed(g, f"data/k_{tag}", tag), typed(g, f"data/v_{tag}", tag)
        (k2, v2), = oracle.trim_kv_block_old(((k, v),), window_size=8, block_size=10, keep_per_block=2, prefix_len=3)
        assert bits_equal(k2, typed(g, f"data/block_k_{tag}", tag))
        assert bits_equal(v2, typed(g, f"data/block_v_{tag}", tag))
        (k3, _), = oracle.trim_kv_budget_old(((k, v),), window_size=8, old_budget=5, prefix_len=2)
        assert bits_equal(k3, typed(g, f"data/budget_k_{tag}", tag))


# pooled summaries: fp32 accumulate, one rounding to dtype; summation order differs from torch's
CHUNK_TOL = {"f32": dict(rtol=1e-6, atol=1e-6), "f16": dict(rtol=2 ** -10, atol=1e-6), "bf16": dict(rtol=2 ** -7, atol=1e-6)}


def test_chunk_summaries(oracle):
    g = load("chunk.npz")
    ones = torch.ones(1, 1, 10, 2)
    got = oracle.chunk_mean(ones, 4, 4)
    assert torch.equal(got, torch.from_numpy(g["kat/out_f32"]))
    assert got[0, 0, :, 0].tolist() == [1.0, 0.5, 1, 1, 1, 1]
    for case in g["cases"]:
        tag = case.split("_")[0]
        x = typed(g, f"{case}/x_{tag}", tag)
        chunk, keep = g[f"{case}/params"].tolist()
        want = typed(g, f"{case}/out_{tag}", tag)
        got = oracle.chunk_mean(x, chunk, keep)
        assert got.shape == want.shape and got.dtype == want.dtype, case
        torch.testing.assert_close(got.float(), want.float(), **CHUNK_TOL[tag])
        keep_eff = min(keep, x.shape[2])
        if keep_eff and x.shape[2] - keep_eff > 0:  # recent tokens are copied exactly
            assert bits_equal(got[:, :, -keep_eff:, :].contiguous(), want[:, :, -keep_eff:, :].contiguous())


def test_attention(oracle):
    g = load("attn.npz")
    for case in g["cases"]:
        mode = case.split("_")[0]
        kb = 4 if mode == "int4" else 8
        vb = 8 if mode == "int8" else 4
        q = torch.from_numpy(g[f"{case}/q"])
        out = oracle.decode_attention(q, torch.from_numpy(g[f"{case}/kstore"]), kb, torch.from_numpy(g[f"{case}/kscales"]),
                                      torch.from_numpy(g[f"{case}/vstore"]), vb, torch.from_numpy(g[f"{case}/vscales"]))
        # fp32 eager attention in torch vs double-accumulated oracle: stated tolerance 1e-5
        torch.testing.assert_close(out, torch.from_numpy(g[f"{case}/out"]), rtol=1e-5, atol=1e-5)


def test_benchmarker_fixture_matches_the_survey_kat():
    """benchmarker.npz was produced by the unmodified reference loop (make_golden.py::make_benchmarker): its est_mb values are
    the SURVEY section 4 known answers, and equal the closed form  layers * tokens * (code bytes + scale bytes)."""
    g = load("benchmarker.npz")
    kat = {"int8": 0.424072265625, "int4": 0.213134765625, "mixed": 0.318603515625}
    L_, H, D, T = 12, 12, 64, 16 + int(g["new_tokens"])
    for mode, want in kat.items():
        kb, vb = (4 if mode == "int4" else 8), (8 if mode == "int8" else 4)
        nbytes = L_ * T * (H * D * kb // 8 + H * D * vb // 8 + 2 * 4)   # fp32 scales on the CPU path (ops.py:271-290)
        assert float(g[f"{mode}/est_mb"]) == want == nbytes / 1024 ** 2
        assert g[f"{mode}/tokens"].tolist()[:16] == g["prompt_ids"][0].tolist() and len(g[f"{mode}/tokens"]) == T
    assert g["int8/tokens"].tolist() == g["full/tokens"].tolist()   # SURVEY: the int8 text equals the full-cache text
