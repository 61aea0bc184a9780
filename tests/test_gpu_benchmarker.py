"""KVCacheBenchmarker on the GPU with a random-init GPT-2 (no weights / tokenizer files exist offline):
all twelve methods run through the B200 kernels, the result dict has the reference's keys and quirks
(benchmarker.py:811-832), and policies that cannot drop anything reproduce the full-cache tokens."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu


class StubTokenizer:
    """The only members the benchmarker touches (reference benchmarker.py:76-78,95,99)."""

    eos_token_id = 50256

    def __init__(self, batch=1):
        self.batch = batch

    def __call__(self, prompt, return_tensors="pt", truncation=False, max_length=None):
        ids = torch.tensor([[(ord(c) * 31 + i * 7) % 1000 for i, c in enumerate(prompt)]] * self.batch)
        if truncation and max_length:
            ids = ids[:, :max_length]
        return type("Enc", (), {"input_ids": ids})()

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)


@pytest.fixture(scope="module")
def bench():
    from transformers import GPT2Config, GPT2LMHeadModel
    from kvq_b200.benchmarking import KVCacheBenchmarker
    torch.manual_seed(0)
    cfg = GPT2Config(n_layer=3, n_head=4, n_embd=256, n_positions=512, vocab_size=1000)
    model = GPT2LMHeadModel(cfg).half().cuda().eval()
    return KVCacheBenchmarker(model, StubTokenizer(), device="cuda")


PROMPT = "the quick brown fox jumps over the lazy dog " * 2  # 88 tokens
KEYS = {"method", "elapsed_sec", "total_new_tokens", "tokens_per_sec", "cpu_mem_used_mb", "gpu_peak_mb", "window_size",
        "block_size", "chunk_size", "est_kv_cache_mb_avg", "prefix_len", "stride", "keep_per_block", "old_budget"}


def test_all_twelve_methods(bench):
    from kvq_b200.benchmarking import KVCacheBenchmarker
    for method in KVCacheBenchmarker.VALID_METHODS:
        r = bench.benchmark_method([PROMPT, PROMPT[:50]], method=method, max_new_tokens=6, window_size=32, block_size=16,
                                   chunk_size=8, keep_last=24, prefix_len=4, stride=3, keep_per_block=2, old_budget=10)
        assert set(r) == KEYS and r["method"] == method
        assert r["total_new_tokens"] == 12 and r["elapsed_sec"] > 0 and r["tokens_per_sec"] > 0
        assert (r["block_size"] == 16) == (method == "paged_attention")       # quirk: not reported for block_cache
        assert (r["chunk_size"] == 8) == (method == "chunked_cache")
        assert (r["window_size"] == 32) == (method in ("sliding_window", "prefix_window", "strided_cache", "block_cache", "budget_cache"))
        assert (r["stride"] == 3) == (method == "strided_cache")
        nan_methods = ("full_cache", "sliding_window", "prefix_window", "strided_cache", "block_cache", "budget_cache")
        assert math.isnan(r["est_kv_cache_mb_avg"]) == (method in nan_methods)


def test_quantized_cache_size_matches_reference_formula(bench):
    L_, H, D = 3, 4, 64
    for mode, per_tok in (("int8", 2 * D), ("int4", D), ("mixed", D + D // 2)):
        text, n_new, est_mb = bench.generate_with_quantized_kv(PROMPT, max_new_tokens=5, mode=mode)
        T = len(PROMPT) + 5  # every generated token's KV is appended (benchmarker.py:486)
        want = L_ * (H * T * per_tok + 2 * T * 2)  # codes + one fp16 scale per token for K and for V (ops.py:271-290)
        assert n_new == 5 and est_mb == want / 1024 ** 2
        assert text == bench.generate_with_quantized_kv(PROMPT, max_new_tokens=5, mode=mode)[0]  # deterministic


def test_lossless_policies_reproduce_full_cache(bench):
    full, n = bench.generate_with_cache(PROMPT, 8)
    assert n == 8
    assert bench.generate_with_sliding_window(PROMPT, 8, window_size=4096)[0] == full
    assert bench.generate_with_prefix_window(PROMPT, 8, window_size=4096, prefix_len=4)[0] == full
    assert bench.generate_with_strided_cache(PROMPT, 8, window_size=4096, stride=4)[0] == full
    assert bench.generate_with_block_cache(PROMPT, 8, window_size=4096)[0] == full
    assert bench.generate_with_budget_cache(PROMPT, 8, window_size=4096)[0] == full
    assert bench.generate_with_chunked_cache(PROMPT, 8, chunk_size=8, keep_last=4096)[0] == full
    text, n_new, alloc_mb, used_mb, nblocks = bench.generate_with_paged_attention(PROMPT, 8, block_size=16)
    assert text == full                              # paging is exact
    T = len(PROMPT) + 8
    assert nblocks == 3 * -(-T // 16) and used_mb == 3 * T * 4 * 64 * 2 * 2 / 1024 ** 2 and alloc_mb >= used_mb


def test_batch16_quant_mixed_with_window(bench):
    """BASELINE C2 shape in miniature: batch 16, mixed INT8-K / INT4-V cache, then a sliding window."""
    from kvq_b200.benchmarking import KVCacheBenchmarker
    b16 = KVCacheBenchmarker(bench.model, StubTokenizer(batch=16), device="cuda")
    r = b16.benchmark_method([PROMPT], method="quant_mixed", max_new_tokens=4)
    assert r["total_new_tokens"] == 4 and r["est_kv_cache_mb_avg"] > 0
    r = b16.benchmark_method([PROMPT], method="sliding_window", max_new_tokens=4, window_size=32)
    assert r["total_new_tokens"] == 4


def test_fused_attention_cache_matches_dequantize_path(bench):
    """KvqFusedCache + fused decode attention vs the reference-style loop (dequantize everything, DynamicCache): same
    quantized store (codes and scales bit-identical), next-token logits within fp16 noise, same cache size."""
    from kvq_b200.hf_cache import KvqFusedCache, fused_attention
    from kvq_b200.quantization import QuantizedKVCache
    from transformers import DynamicCache
    model = bench.model
    ids = StubTokenizer()(PROMPT).input_ids.cuda()
    for mode in ("int8", "mixed"):
        # reference-style: prefill, quantize the prompt, one step on the dequantized cache
        out = model(input_ids=ids, use_cache=True)
        kv = tuple((l.keys, l.values) for l in out.past_key_values.layers)
        qc = QuantizedKVCache(n_layers=len(kv), mode=mode, device="cuda", compute_dtype=torch.float16)
        qc.init_from_prompt_past(kv)
        nxt = torch.argmax(out.logits[:, -1, :], dim=-1, keepdim=True)
        ref = model(input_ids=nxt, use_cache=True, past_key_values=DynamicCache(ddp_cache_data=qc.to_past_key_values()))
        # fused: the model writes into / reads from the quantized store
        cache = KvqFusedCache(len(kv), mode=mode, compute_dtype=torch.float16)
        with fused_attention(model):
            out2 = model(input_ids=ids, use_cache=True, past_key_values=cache)
            assert torch.equal(torch.argmax(out2.logits[:, -1, :], -1, keepdim=True), nxt)
            for a, b in zip(cache.layers, qc.layers):  # identical stores after the prompt
                assert torch.equal(a.store.quantized()[0], b.quantized()[0]) and torch.equal(a.store.quantized()[3], b.quantized()[3])
                assert torch.equal(a.store.quantized()[1], b.quantized()[1])
            got = model(input_ids=nxt, use_cache=True, past_key_values=cache)
        assert cache.get_seq_length() == ids.size(1) + 1
        torch.testing.assert_close(got.logits.float(), ref.logits.float(), rtol=2e-2, atol=2e-2)
    t_fused = bench.generate_with_quantized_kv(PROMPT, 8, mode="int8", fused_attention=True)
    t_plain = bench.generate_with_quantized_kv(PROMPT, 8, mode="int8")
    assert t_fused[1] == t_plain[1] == 8 and t_fused[2] == t_plain[2]
    agree = sum(a == b for a, b in zip(t_fused[0].split(), t_plain[0].split())) / len(t_plain[0].split())
    assert agree >= 0.95  # random-init logits are nearly flat: allow an occasional argmax flip late in the sequence
    r = bench.benchmark_method([PROMPT], method="quant_mixed", max_new_tokens=4, fused_attention=True)
    assert r["total_new_tokens"] == 4 and r["est_kv_cache_mb_avg"] > 0


def test_composed_quant_plus_sliding_window(bench):
    """BASELINE C2 in miniature: mixed INT8-K / INT4-V store + sliding window in ONE cache object (fused attention)."""
    ev = {"strategy": "sliding_window", "window_size": 32}
    text, n_new, est_mb = bench.generate_with_quantized_kv(PROMPT, 6, mode="mixed", fused_attention=True, evict=ev)
    # the store holds the window (trimmed before each append) + the token of the last step
    assert n_new == 6 and est_mb == 3 * (4 * 33 * (64 + 32) + 2 * 33 * 2) / 1024 ** 2
    # a window larger than the sequence changes nothing
    big = bench.generate_with_quantized_kv(PROMPT, 6, mode="mixed", fused_attention=True, evict={"strategy": "sliding_window", "window_size": 4096})
    assert big[0] == bench.generate_with_quantized_kv(PROMPT, 6, mode="mixed", fused_attention=True)[0]
    with pytest.raises(ValueError):
        bench.generate_with_quantized_kv(PROMPT, 2, mode="mixed", evict=ev)


@pytest.mark.parametrize("kv_heads,head_dim", [(2, 64), (1, 128)])  # G = 4 at D = 64; G = 8 at D = 128 (Llama-3-70B's grouping)
def test_fused_attention_on_a_gqa_model(kv_heads, head_dim):
    """SURVEY 8f-1 on a grouped-query model (random-init Llama: RoPE keys, num_key_value_heads < num_attention_heads): the
    register-tiled GQA kernel behind the HF attention interface vs the reference-style dequantize-everything loop."""
    from transformers import DynamicCache, LlamaConfig, LlamaForCausalLM
    from kvq_b200.benchmarking import KVCacheBenchmarker
    from kvq_b200.hf_cache import KvqFusedCache, fused_attention
    from kvq_b200.quantization import QuantizedKVCache
    torch.manual_seed(1)
    cfg = LlamaConfig(hidden_size=8 * head_dim, intermediate_size=512, num_hidden_layers=2, num_attention_heads=8,
                      num_key_value_heads=kv_heads, head_dim=head_dim, vocab_size=1000, max_position_embeddings=512)
    model = LlamaForCausalLM(cfg).half().cuda().eval()
    ids = StubTokenizer(batch=2)(PROMPT * 2).input_ids.cuda()  # 176 tokens: more than one 128-token tile
    for mode in ("int8", "mixed"):
        out = model(input_ids=ids, use_cache=True)
        kv = tuple((l.keys, l.values) for l in out.past_key_values.layers)
        assert kv[0][0].shape[1] == kv_heads
        qc = QuantizedKVCache(n_layers=len(kv), mode=mode, device="cuda", compute_dtype=torch.float16)
        qc.init_from_prompt_past(kv)
        nxt = torch.argmax(out.logits[:, -1, :], dim=-1, keepdim=True)
        ref = model(input_ids=nxt, use_cache=True, past_key_values=DynamicCache(ddp_cache_data=qc.to_past_key_values()))
        cache = KvqFusedCache(len(kv), mode=mode, compute_dtype=torch.float16)
        with fused_attention(model):
            out2 = model(input_ids=ids, use_cache=True, past_key_values=cache)
            torch.testing.assert_close(out2.logits.float(), out.logits.float(), rtol=2e-2, atol=2e-2)
            for a, b in zip(cache.layers, qc.layers):
                assert torch.equal(a.store.quantized()[0], b.quantized()[0]) and torch.equal(a.store.quantized()[3], b.quantized()[3])
            got = model(input_ids=nxt, use_cache=True, past_key_values=cache)
        torch.testing.assert_close(got.logits.float(), ref.logits.float(), rtol=2e-2, atol=2e-2)
    b = KVCacheBenchmarker(model, StubTokenizer(), device="cuda")
    t_fused = b.generate_with_quantized_kv(PROMPT, 8, mode="mixed", fused_attention=True)
    t_plain = b.generate_with_quantized_kv(PROMPT, 8, mode="mixed")
    assert t_fused[1] == t_plain[1] == 8 and t_fused[2] == t_plain[2]
    agree = sum(x == y for x, y in zip(t_fused[0].split(), t_plain[0].split())) / len(t_plain[0].split())
    assert agree >= 0.95


def test_fused_attention_on_a_paged_store(bench):
    """Quantization composed with paging behind the HF cache interface: same codes, attention bit-identical to the contiguous
    store, so the generated tokens are identical; with a sliding window the paged store drops whole blocks."""
    for mode in ("int8", "mixed"):
        flat = bench.generate_with_quantized_kv(PROMPT, 12, mode=mode, fused_attention=True)
        paged = bench.generate_with_quantized_kv(PROMPT, 12, mode=mode, fused_attention=True, paged_block_size=16)
        assert paged[0] == flat[0] and paged[1] == flat[1] == 12
        assert paged[2] >= flat[2]  # blocks in use, including the unused tail of the last block
    ev = {"strategy": "sliding_window", "window_size": 32}
    text, n_new, est_mb = bench.generate_with_quantized_kv(PROMPT, 6, mode="mixed", fused_attention=True, evict=ev, paged_block_size=16)
    assert n_new == 6
    # 88 + 6 tokens, window 32, blocks of 16: after the last trim 32..47 cached tokens remain, + the token of the last step
    assert 3 * (4 * 32 * (64 + 32)) / 1024 ** 2 <= est_mb <= 3 * (4 * 64 * (64 + 32) + 2 * 64 * 2) / 1024 ** 2
    r = bench.benchmark_method([PROMPT], method="quant_mixed", max_new_tokens=4, fused_attention=True, paged_block_size=32)
    assert r["total_new_tokens"] == 4 and r["est_kv_cache_mb_avg"] > 0
    with pytest.raises(ValueError):
        bench.generate_with_quantized_kv(PROMPT, 2, mode="int8", paged_block_size=16)
    with pytest.raises(ValueError):
        bench.generate_with_quantized_kv(PROMPT, 2, mode="int8", fused_attention=True, paged_block_size=16, evict={"strategy": "strided"})


def test_fused_attention_honours_a_padding_mask(bench):
    """A left-padded batch: the mask hides cached tokens, so the step must not take the fused kernel (which sees every cached
    token); it takes the dequantize path with the mask and agrees with the reference-style loop."""
    from kvq_b200.hf_cache import KvqFusedCache, fused_attention
    from kvq_b200.quantization import QuantizedKVCache
    from transformers import DynamicCache
    model = bench.model
    ids = StubTokenizer(batch=2)(PROMPT[:40]).input_ids.cuda()
    mask = torch.ones_like(ids)
    mask[0, :7] = 0  # row 0 is left-padded
    out = model(input_ids=ids, attention_mask=mask, use_cache=True)
    kv = tuple((l.keys, l.values) for l in out.past_key_values.layers)
    qc = QuantizedKVCache(n_layers=len(kv), mode="mixed", device="cuda", compute_dtype=torch.float16)
    qc.init_from_prompt_past(kv)
    nxt = torch.argmax(out.logits[:, -1, :], dim=-1, keepdim=True)
    mask2 = torch.cat([mask, torch.ones_like(nxt)], dim=1)
    ref = model(input_ids=nxt, attention_mask=mask2, use_cache=True, past_key_values=DynamicCache(ddp_cache_data=qc.to_past_key_values()))
    cache = KvqFusedCache(len(kv), mode="mixed", compute_dtype=torch.float16)
    with fused_attention(model):
        model(input_ids=ids, attention_mask=mask, use_cache=True, past_key_values=cache)
        got = model(input_ids=nxt, attention_mask=mask2, use_cache=True, past_key_values=cache)
        # without padding the same call takes the fused kernel and must differ for the padded row only if the mask matters
        cache2 = KvqFusedCache(len(kv), mode="mixed", compute_dtype=torch.float16)
        model(input_ids=ids, use_cache=True, past_key_values=cache2)
        nomask = model(input_ids=nxt, use_cache=True, past_key_values=cache2)
    torch.testing.assert_close(got.logits.float(), ref.logits.float(), rtol=2e-2, atol=2e-2)
    assert not torch.allclose(got.logits[0].float(), nomask.logits[0].float(), rtol=2e-2, atol=2e-2)  # the mask was applied


def test_two_caches_in_flight_do_not_cross_read(bench):
    """_pending is keyed per thread and checked against the key tensor: interleaving two caches on one model is safe."""
    from kvq_b200.hf_cache import KvqFusedCache, fused_attention
    model = bench.model
    a_ids, b_ids = StubTokenizer()(PROMPT).input_ids.cuda(), StubTokenizer()(PROMPT[::-1]).input_ids.cuda()
    n = model.config.n_layer
    with fused_attention(model):
        ca, cb = KvqFusedCache(n, mode="int8"), KvqFusedCache(n, mode="int8")
        oa, ob = model(input_ids=a_ids, use_cache=True, past_key_values=ca), model(input_ids=b_ids, use_cache=True, past_key_values=cb)
        ta, tb = oa.logits[:, -1:].argmax(-1), ob.logits[:, -1:].argmax(-1)
        la = model(input_ids=ta, use_cache=True, past_key_values=ca).logits
        lb = model(input_ids=tb, use_cache=True, past_key_values=cb).logits
        solo = KvqFusedCache(n, mode="int8")
        model(input_ids=a_ids, use_cache=True, past_key_values=solo)
        la_solo = model(input_ids=ta, use_cache=True, past_key_values=solo).logits
    assert torch.equal(la, la_solo) and not torch.equal(la, lb)


def test_composed_window_sees_the_positions_of_trim_after_step(bench):
    """quant_mixed + sliding window as ONE cache vs composing the reference functions (quantize the store, trim to the window
    right after every step, benchmarker.py:199-205): the model must see the same cache length -- hence the same position ids
    -- and produce the same logits (the lazily trimmed store used to report W + 1)."""
    from kvq_b200.hf_cache import KvqFusedCache, fused_attention
    from kvq_b200.quantization import QuantizedKVCache
    from transformers import DynamicCache
    model, W = bench.model, 32
    ids = StubTokenizer()(PROMPT).input_ids.cuda()
    out = model(input_ids=ids, use_cache=True)
    kv = tuple((l.keys, l.values) for l in out.past_key_values.layers)
    qc = QuantizedKVCache(n_layers=len(kv), mode="mixed", device="cuda", compute_dtype=torch.float16)
    qc.init_from_prompt_past(kv)
    for l in qc.layers:
        l.evict("sliding_window", window_size=W)
    nxt = torch.argmax(out.logits[:, -1, :], dim=-1, keepdim=True)
    toks, ref_logits = [nxt], []
    for _ in range(3):
        o = model(input_ids=toks[-1], use_cache=True, past_key_values=DynamicCache(ddp_cache_data=qc.to_past_key_values()))
        ref_logits.append(o.logits.float())
        qc.append_from_past(tuple((l.keys, l.values) for l in o.past_key_values.layers))
        for l in qc.layers:
            l.evict("sliding_window", window_size=W)
        toks.append(torch.argmax(o.logits[:, -1, :], dim=-1, keepdim=True))
    cache = KvqFusedCache(len(kv), mode="mixed", compute_dtype=torch.float16, evict={"strategy": "sliding_window", "window_size": W})
    with fused_attention(model):
        model(input_ids=ids, use_cache=True, past_key_values=cache)
        for i in range(3):
            assert cache.get_seq_length() == W
            got = model(input_ids=toks[i], use_cache=True, past_key_values=cache).logits.float()
            torch.testing.assert_close(got, ref_logits[i], rtol=2e-2, atol=2e-2)


def test_golden_benchmarker_gpt2_small():
    """BASELINE C1's model: the reference's own generate_with_quantized_kv on a seeded random-init GPT-2 small (fp32, CPU;
    tests/golden/make_golden.py::make_benchmarker) vs ours on the GPU with the same weights and prompt.  est_mb must be the
    reference's number exactly (SURVEY section 4 KAT: 0.424072265625 / 0.213134765625 / 0.318603515625 MB -- codes plus
    fp32 scales of 24 tokens x 12 layers); n_new equal; greedy-token agreement is reported (a random-init model has nearly
    flat logits, CPU and GPU matmuls round differently, so an argmax can flip: the first new token -- produced by the
    exact prefill -- and at least half of the int8 continuation must agree)."""
    from _golden import load
    from transformers import GPT2Config, GPT2LMHeadModel
    from kvq_b200.benchmarking import KVCacheBenchmarker
    g = load("benchmarker.npz")
    ids = torch.from_numpy(g["prompt_ids"])
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config()).eval().cuda()   # same seeded init as the fixture, fp32

    class Tok(StubTokenizer):
        def __call__(self, prompt, return_tensors="pt", truncation=False, max_length=None):
            return type("Enc", (), {"input_ids": ids})()

    bench = KVCacheBenchmarker(model, Tok(), device="cuda")
    n = int(g["new_tokens"])
    for mode, kat in (("int8", 0.424072265625), ("int4", 0.213134765625), ("mixed", 0.318603515625)):
        for fused in (False, True):
            text, n_new, est_mb = bench.generate_with_quantized_kv("x", max_new_tokens=n, mode=mode, fused_attention=fused)
            assert n_new == int(g[f"{mode}/n_new"]) == n
            assert est_mb == float(g[f"{mode}/est_mb"]) == kat, (mode, fused, est_mb)
            got, want = [int(t) for t in text.split()], g[f"{mode}/tokens"].tolist()
            assert got[:16] == want[:16] == ids[0].tolist()
            agree = sum(a == b for a, b in zip(got[16:], want[16:])) / n
            print(f"golden benchmarker {mode} fused={fused}: est_mb {est_mb} == KAT, greedy-token agreement {agree:.2f}")
            assert got[16] == want[16]
            if mode == "int8" and not fused:
                assert agree >= 0.5
