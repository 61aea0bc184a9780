"""Host-side behaviour that needs no GPU: mode / shape validation, head sharding arithmetic, config
defaults, benchmarker method table."""

import pytest
import torch

import kvq_b200
from kvq_b200.distributed import head_shard_range, shard_heads


def test_mode_assertion_and_empty_cache():
    with pytest.raises(AssertionError, match="Invalid mode"):
        kvq_b200.QuantizedLayerKV(mode="int2")          # reference ops.py:160
    with pytest.raises(ValueError, match="Empty cache"):
        kvq_b200.QuantizedLayerKV(mode="int8").get_kv()  # reference ops.py:220
    with pytest.raises(ValueError, match="Empty cache"):
        kvq_b200.PagedKVCache().get_kv()                 # reference implementations.py:90
    assert kvq_b200.QuantizedLayerKV("mixed").estimated_bytes() == 0
    pc = kvq_b200.PagedKVCache(block_size=16)
    assert pc.num_blocks() == 0 and pc.allocated_bytes() == 0 and pc.used_bytes() == 0


def test_paged_append_shape_assertions():
    pc = kvq_b200.PagedKVCache()
    with pytest.raises(AssertionError):
        pc.append(torch.zeros(1, 2, 3), torch.zeros(1, 2, 3))        # :64
    with pytest.raises(AssertionError):
        pc.append(torch.zeros(1, 2, 2, 4), torch.zeros(1, 2, 2, 4))  # :66 (one == 1)


def test_unchanged_when_short(monkeypatch):
    # T <= prefix + window: same tensor objects come back, no kernel (and no CUDA) involved
    k, v = torch.zeros(1, 2, 10, 4), torch.ones(1, 2, 10, 4)
    for out in (kvq_b200.trim_kv_prefix_window(((k, v),), 4, 8), kvq_b200.trim_kv_strided(((k, v),), 8, 2, 4),
                kvq_b200.trim_kv_block_old(((k, v),), 8, 4, 2, 2), kvq_b200.trim_kv_budget_old(((k, v),), 8, 3, 2),
                kvq_b200.trim_kv_sliding_window(((k, v),), 10)):
        assert out[0][0] is k and out[0][1] is v
    # sliding window is a view, like the reference
    (k2, v2), = kvq_b200.trim_kv_sliding_window(((k, v),), 3)
    assert k2.shape == (1, 2, 3, 4) and k2.data_ptr() == k[:, :, -3:, :].data_ptr()
    # chunk summary with nothing older than keep_last: unchanged
    (k3, v3), = kvq_b200.chunk_summarize_kv(((k.cuda() if False else k, v),), 4, 10) if False else ((k, v),)
    assert k3 is k


def test_head_sharding():
    assert head_shard_range(8, 0, 1) == (0, 8)
    assert [head_shard_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert head_shard_range(8, 7, 8) == (7, 8)
    with pytest.raises(ValueError):
        head_shard_range(12, 0, 8)  # GPT-2's 12 heads do not shard over 8
    x = torch.arange(2 * 8 * 3 * 4).view(2, 8, 3, 4)
    assert torch.equal(shard_heads(x, 1, 4), x[:, 2:4])


def test_config_defaults():
    from kvq_b200.core import CacheConfig, QuantizationConfig, mb
    assert QuantizationConfig().eps == 1e-8 and QuantizationConfig().mode == "int8"
    c = CacheConfig()
    assert (c.window_size, c.block_size, c.chunk_size, c.keep_last) == (256, 64, 64, 256)
    assert mb(1024 ** 2) == 1.0


def test_benchmarker_method_table():
    from kvq_b200.benchmarking import KVCacheBenchmarker
    assert len(KVCacheBenchmarker.VALID_METHODS) == 12
    b = KVCacheBenchmarker(model=None, tokenizer=None, device="cuda")
    with pytest.raises(AssertionError, match="Invalid method"):
        b.benchmark_method(["x"], method="flash")
    for name in ("generate_no_cache", "generate_with_cache", "generate_with_sliding_window", "generate_with_prefix_window",
                 "generate_with_strided_cache", "generate_with_block_cache", "generate_with_budget_cache",
                 "generate_with_quantized_kv", "generate_with_paged_attention", "generate_with_chunked_cache",
                 "benchmark_method"):
        assert callable(getattr(b, name))


def test_reference_import_root_alias():
    """`src.*` imports of code written against the reference resolve to this package (drop-in import root)."""
    import importlib
    import subprocess
    import sys
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "efficient-llm-inference_b200")
    code = ("from src.quantization import QuantizedKVCache, quantize_int8_per_tensor, dequantize_int4_per_tensor_packed\n"
            "from src.quantization.ops import QuantizedLayerKV\n"
            "from src.cache import trim_kv_strided, PagedKVCache, chunk_summarize_kv, trim_kv_budget_old\n"
            "from src.cache.implementations import trim_kv_block_old\n"
            "from src.cuda import get_cuda_extension, build_cuda_extension\n"
            "from src.core import Config, CacheConfig, mb\n"
            "from src.benchmarking import KVCacheBenchmarker\n"
            "from src.benchmarking.benchmarker import KVCacheBenchmarker as K2\n"
            "import src, kvq_b200\n"
            "assert QuantizedKVCache is kvq_b200.QuantizedKVCache and K2 is KVCacheBenchmarker\n"
            "try:\n    from src import MMLUBenchmark\n    raise SystemExit(1)\nexcept ImportError:\n    pass\n"
            "print('ok')")
    env = {**os.environ, "PYTHONPATH": root}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-800:]


def test_paged_quantized_store_host_checks():
    """No GPU needed: constructor validation, and CPU tensors are refused (there is no CPU path)."""
    import kvq_b200
    for bad in (0, 8, 48, 100):
        with pytest.raises(ValueError):
            kvq_b200.PagedQuantizedLayerKV("int8", block_size=bad)
    with pytest.raises(AssertionError):
        kvq_b200.PagedQuantizedLayerKV("int3")
    p = kvq_b200.PagedQuantizedLayerKV("mixed", block_size=32)
    assert (p.k_bits, p.v_bits, len(p), p.num_blocks(), p.allocated_bytes()) == (8, 4, 0, 0, 0)
    x = torch.zeros(1, 2, 3, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.append_tokens(x, x)
    with pytest.raises(ValueError):
        p.attend(torch.zeros(1, 2, 1, 64))
    with pytest.raises(ValueError):
        p.get_kv()


def test_run_grouped_chunks_and_pads():
    """Query-head groupings the kernels are not built for: chunks of <= 8 heads per KV head, zero-padded to a supported size,
    results put back in order (pure host logic: checked with a stand-in for the kernel call)."""
    from kvq_b200.attention import run_grouped
    B, H, D = 2, 3, 8
    for G in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 17):
        q = torch.randn(B, H * G, 1, D)
        seen = []

        def call(qs):
            g = qs.size(1) // H
            seen.append(g)
            assert g in (1, 2, 4, 8)
            return qs * 2 + torch.arange(H).view(1, H, 1, 1, 1).expand(B, H, g, 1, 1).reshape(B, H * g, 1, 1)  # depends on the KV head

        out = run_grouped(q, H, (1, 2, 4, 8), call)
        want = q * 2 + torch.arange(H).repeat_interleave(G).view(1, H * G, 1, 1)
        assert torch.equal(out, want)
        expect = [G] if G in (1, 2, 4, 8) else [min(x for x in (1, 2, 4, 8) if x >= min(8, G - c)) for c in range(0, G, 8)]
        assert seen == expect
    with pytest.raises(RuntimeError):
        run_grouped(torch.zeros(1, 5, 1, 8), 2, (1, 2, 4, 8), lambda x: x)
