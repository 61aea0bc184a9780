"""Quantization API (reference: src/quantization/__init__.py:12-19)."""

from .ops import (
    QuantizedKVCache,
    QuantizedLayerKV,
    TensorTable,
    dequantize_per_tensor_batched,
    quantize_per_tensor_batched,
    dequantize_int4_per_tensor_packed,
    dequantize_int8_per_tensor,
    quantize_int4_per_tensor_packed,
    quantize_int8_per_tensor,
)

__all__ = [
    "quantize_int8_per_tensor",
    "quantize_int4_per_tensor_packed",
    "dequantize_int8_per_tensor",
    "dequantize_int4_per_tensor_packed",
    "QuantizedLayerKV",
    "QuantizedKVCache",
    "TensorTable",
    "quantize_per_tensor_batched",
    "dequantize_per_tensor_batched",
]
