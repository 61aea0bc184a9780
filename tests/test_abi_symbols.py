"""The C-ABI library loads without a GPU and exports every symbol include/kvq.h declares; the ctypes
signature table covers exactly that set.  No compute calls here."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kvq.h")).read()
    return sorted(set(re.findall(r"KVQ_API\s+[\w\s\*]+?\b(kvq_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ("kvq_absmax", "kvq_quant_int8", "kvq_quant_int4_pack", "kvq_quant_per_token_fused", "kvq_dequant_int8",
                 "kvq_dequant_int4_unpack", "kvq_evict_plan_make", "kvq_evict_gather", "kvq_paged_append",
                 "kvq_paged_gather", "kvq_chunk_mean", "kvq_decode_attn", "kvq_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from kvq_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build the library first: make -C efficient-llm-inference_b200/csrc"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/kvq.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().kvq_version() == 100


def test_no_cpu_fallback():
    import torch
    import kvq_b200

    x = torch.randn(2, 2, 4, 16)
    for fn in (lambda: kvq_b200.quantize_int8_per_tensor(x), lambda: kvq_b200.quantize_int4_per_tensor_packed(x),
               lambda: kvq_b200.dequantize_int8_per_tensor(torch.zeros(4, dtype=torch.int8), torch.tensor(1.0), torch.float32),
               lambda: kvq_b200.trim_kv_strided(((x.repeat(1, 1, 100, 1), x.repeat(1, 1, 100, 1)),), 8, 2),
               lambda: kvq_b200.chunk_summarize_kv(((x, x),), 2, 1),
               lambda: kvq_b200.QuantizedLayerKV("int8").append(x[:, :, :1], x[:, :, :1])):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "efficient-llm-inference_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "kvq_oracle" not in text.replace("oracle kvqo_check_division_free", ""), f"{f} references the oracle"
                assert "import oracle" not in text and "from oracle" not in text
