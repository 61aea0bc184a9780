"""`src` -- the reference's import root, served by kvq_b200.

Put `efficient-llm-inference_b200/` on PYTHONPATH *instead of* the reference checkout and code written against the
reference keeps its imports:

    from src.quantization import QuantizedKVCache, quantize_int8_per_tensor
    from src.cache import trim_kv_strided, PagedKVCache, chunk_summarize_kv
    from src.cuda import get_cuda_extension
    from src.core import Config, CacheConfig, mb
    from src.benchmarking import KVCacheBenchmarker            # (imports transformers)

Only the KV-cache hot path exists here; the task harnesses (`MMLUBenchmark`,
`SummarizationBenchmark`), `src.datasets` and `src.evaluation` of the reference are out of scope and raise ImportError.
"""

import importlib
import sys

import kvq_b200 as _impl

__version__ = _impl.__version__

for _name in ("quantization", "cache", "cuda", "core"):
    _mod = importlib.import_module(f"kvq_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
for _sub in ("quantization.ops", "cache.implementations", "cuda.extensions", "core.config", "core.utils"):
    sys.modules[f"{__name__}.{_sub}"] = importlib.import_module(f"kvq_b200.{_sub}")

from kvq_b200.core import BenchmarkConfig, CacheConfig, Config, QuantizationConfig  # noqa: E402,F401


def __getattr__(name):
    if name == "KVCacheBenchmarker":  # lazy: pulls in transformers
        return importlib.import_module(f"{__name__}.benchmarking").KVCacheBenchmarker
    if name in ("MMLUBenchmark", "SummarizationBenchmark", "load_cnn_dailymail", "load_xsum", "load_samsum", "load_mmlu",
                "datasets", "evaluation"):
        raise ImportError(f"src.{name} is outside the KV-cache hot path rebuilt here (SURVEY.md section 8)")
    raise AttributeError(name)
