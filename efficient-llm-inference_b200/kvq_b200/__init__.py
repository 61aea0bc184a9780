"""kvq_b200 -- B200-native (sm_100a) KV-cache hot path with the Python API of Efficient-LLM-Inference.

Drop-in for the reference's `src.quantization`, `src.cache`, `src.cuda` and
`src.benchmarking.KVCacheBenchmarker`; every operation launches hand-written CUDA kernels from
libkvq_sm100.so through the C ABI in include/kvq.h.  No CPU fallback.
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .quantization import (  # noqa: F401
    QuantizedKVCache,
    QuantizedLayerKV,
    dequantize_int4_per_tensor_packed,
    dequantize_int8_per_tensor,
    quantize_int4_per_tensor_packed,
    quantize_int8_per_tensor,
)
from .cache import (  # noqa: F401
    PagedKVCache,
    chunk_summarize_kv,
    trim_kv_block_old,
    trim_kv_budget_old,
    trim_kv_prefix_window,
    trim_kv_sliding_window,
    trim_kv_strided,
)
from .attention import decode_attention_paged, decode_attention_quantized  # noqa: F401
from .cache.paged_quantized import PagedQuantizedLayerKV  # noqa: F401
from .cuda import build_cuda_extension, get_cuda_extension  # noqa: F401
from .core import BenchmarkConfig, CacheConfig, Config, QuantizationConfig  # noqa: F401


def __getattr__(name):
    if name == "KVCacheBenchmarker":  # imports transformers: keep it lazy
        from .benchmarking import KVCacheBenchmarker
        return KVCacheBenchmarker
    raise AttributeError(name)
