/*
 * kvq_oracle.c -- CPU restatement of the Efficient-LLM-Inference KV-cache hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under efficient-llm-inference_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 * fixtures under tests/golden/ that were produced by importing the unmodified reference
 * (/root/reference/src/...) in the build container (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 * Plain C loops, OpenMP over independent rows; arithmetic is IEEE fp32 exactly as the torch
 * CPU kernels the reference calls (division is a true division, rounding is half-to-even).
 *
 * dtype codes: 0 = float32, 1 = float16, 2 = bfloat16.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KVQO_F32 0
#define KVQO_F16 1
#define KVQO_BF16 2

/* ------------------------------------------------------------------ element conversion */

static inline float bits_to_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f32_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float bf16_to_f32(uint16_t h) { return bits_to_f32((uint32_t)h << 16); }
static inline uint16_t f32_to_bf16(float f) { /* round-to-nearest-even, like torch .to(bfloat16) */
    uint32_t u = f32_to_bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) { _Float16 x; memcpy(&x, &h, 2); return (float)x; }
static inline uint16_t f32_to_f16(float f) { _Float16 x = (_Float16)f; uint16_t h; memcpy(&h, &x, 2); return h; }

static inline size_t esize(int dtype) { return dtype == KVQO_F32 ? 4 : 2; }

static inline float load_elem(const void* p, int dtype, int64_t i) {
    switch (dtype) {
        case KVQO_F32: return ((const float*)p)[i];
        case KVQO_F16: return f16_to_f32(((const uint16_t*)p)[i]);
        default: return bf16_to_f32(((const uint16_t*)p)[i]);
    }
}
static inline void store_elem(void* p, int dtype, int64_t i, float v) {
    switch (dtype) {
        case KVQO_F32: ((float*)p)[i] = v; break;
        case KVQO_F16: ((uint16_t*)p)[i] = f32_to_f16(v); break;
        default: ((uint16_t*)p)[i] = f32_to_bf16(v); break;
    }
}
/* value of v after a round trip through `dtype` (what `scale.to(x.dtype)` then `.float()` sees) */
static inline float round_through(float v, int dtype) {
    switch (dtype) {
        case KVQO_F32: return v;
        case KVQO_F16: return f16_to_f32(f32_to_f16(v));
        default: return bf16_to_f32(f32_to_bf16(v));
    }
}

int kvqo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void kvqo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ area 1: quantize / dequantize */

/* max |x| over n contiguous elements, fp32.  src/quantization/ops.py:26-27 and :47-48 */
float kvqo_absmax(const void* x, int dtype, int64_t n) {
    float m = 0.0f;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float a = fabsf(load_elem(x, dtype, i));
        if (a > m) m = a;
    }
    return m;
}

/* scale = (absmax / qmax).clamp(min=eps) in fp32.  ops.py:28 (qmax 127) and ops.py:49 (qmax 7) */
float kvqo_scale_from_absmax(float absmax, float qmax, float eps) {
    float s = absmax / qmax;
    return s < eps ? eps : s;
}

static inline float code_of(float x, float scale, float lo, float hi) {
    /* clamp(round(x32 / scale), lo, hi): ops.py:29 / ops.py:50; torch.round is half-to-even */
    float r = nearbyintf(x / scale);
    if (r < lo) r = lo;
    if (r > hi) r = hi;
    return r;
}

/* quantize_int8_per_tensor, ops.py:10-30, with the absmax supplied (so that a per-token or a
 * head-sharded caller can provide a max taken over a different extent).  Writes n int8 codes,
 * returns the fp32 scale actually divided by; *scale_out receives scale.to(x.dtype). */
float kvqo_quant_int8_given_absmax(const void* x, int dtype, int64_t n, float absmax, float eps,
                                   int8_t* q, void* scale_out) {
    const float scale = kvqo_scale_from_absmax(absmax, 127.0f, eps);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) q[i] = (int8_t)code_of(load_elem(x, dtype, i), scale, -127.0f, 127.0f);
    if (scale_out) store_elem(scale_out, dtype, 0, scale);
    return scale;
}

float kvqo_quant_int8(const void* x, int dtype, int64_t n, float eps, int8_t* q, void* scale_out) {
    return kvqo_quant_int8_given_absmax(x, dtype, n, kvqo_absmax(x, dtype, n), eps, q, scale_out);
}

/* quantize_int4_per_tensor_packed, ops.py:33-65.  x is `rows` rows of `d` elements (last dim d);
 * packed is rows x ceil(d/2) bytes: even index -> high nibble (ops.py:61-63), an odd last dim is
 * padded with one zero code, i.e. nibble 8 (ops.py:54-59). */
float kvqo_quant_int4_given_absmax(const void* x, int dtype, int64_t rows, int64_t d, float absmax,
                                   float eps, uint8_t* packed, void* scale_out) {
    const float scale = kvqo_scale_from_absmax(absmax, 7.0f, eps);
    const int64_t p = (d + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t j = 0; j < p; ++j) {
            int hi = (int)code_of(load_elem(x, dtype, r * d + 2 * j), scale, -8.0f, 7.0f) + 8;
            int lo = 8; /* padded zero code */
            if (2 * j + 1 < d) lo = (int)code_of(load_elem(x, dtype, r * d + 2 * j + 1), scale, -8.0f, 7.0f) + 8;
            packed[r * p + j] = (uint8_t)((hi << 4) | lo);
        }
    }
    if (scale_out) store_elem(scale_out, dtype, 0, scale);
    return scale;
}

float kvqo_quant_int4(const void* x, int dtype, int64_t rows, int64_t d, float eps, uint8_t* packed,
                      void* scale_out) {
    return kvqo_quant_int4_given_absmax(x, dtype, rows, d, kvqo_absmax(x, dtype, rows * d), eps, packed,
                                        scale_out);
}

/* dequantize_int8_per_tensor, ops.py:68-90 (the torch branch, :90): (q.float()*scale.float()).to(out) */
void kvqo_dequant_int8(const int8_t* q, int64_t n, const void* scale, int scale_dtype, void* out,
                       int out_dtype) {
    const float s = load_elem(scale, scale_dtype, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) store_elem(out, out_dtype, i, (float)q[i] * s);
}

/* dequantize_int4_per_tensor_packed, ops.py:93-133 (torch branch :122-133): high nibble -> even
 * index, low -> odd, minus 8, keep the first orig_last of each row, times scale. */
void kvqo_dequant_int4(const uint8_t* packed, int64_t rows, int64_t p, int64_t orig_last,
                       const void* scale, int scale_dtype, void* out, int out_dtype) {
    const float s = load_elem(scale, scale_dtype, 0);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t j = 0; j < orig_last; ++j) {
            uint8_t b = packed[r * p + (j >> 1)];
            int nib = (j & 1) ? (b & 0x0f) : (b >> 4);
            store_elem(out, out_dtype, r * orig_last + j, (float)(nib - 8) * s);
        }
    }
}

/* Per-token granularity: what QuantizedKVCache.init_from_prompt_past (ops.py:333-342) produces when
 * it calls QuantizedLayerKV.append (ops.py:174-210) on x[:, :, t:t+1, :] for every t -- one
 * "per-tensor" scale per token, shared by all B*H rows of that token.  x is contiguous [BH, T, D].
 * bits = 8: q is [BH, T, D] int8.  bits = 4: q is [BH, T, ceil(D/2)] uint8.
 * scales_out: T elements in x dtype (scale.to(x.dtype)); scales_f32 (optional): the unrounded scale. */
void kvqo_quant_per_token(const void* x, int dtype, int64_t BH, int64_t T, int64_t D, int bits, float eps,
                          void* q, void* scales_out, float* scales_f32) {
    const float qmax = bits == 8 ? 127.0f : 7.0f;
    const float lo = bits == 8 ? -127.0f : -8.0f, hi = qmax;
    const int64_t P = (D + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        float m = 0.0f;
        for (int64_t r = 0; r < BH; ++r)
            for (int64_t j = 0; j < D; ++j) {
                float a = fabsf(load_elem(x, dtype, (r * T + t) * D + j));
                if (a > m) m = a;
            }
        const float scale = kvqo_scale_from_absmax(m, qmax, eps);
        if (scales_f32) scales_f32[t] = scale;
        if (scales_out) store_elem(scales_out, dtype, t, scale);
        for (int64_t r = 0; r < BH; ++r) {
            const int64_t base = (r * T + t) * D;
            if (bits == 8) {
                int8_t* qo = (int8_t*)q + base;
                for (int64_t j = 0; j < D; ++j) qo[j] = (int8_t)code_of(load_elem(x, dtype, base + j), scale, lo, hi);
            } else {
                uint8_t* po = (uint8_t*)q + (r * T + t) * P;
                for (int64_t j = 0; j < P; ++j) {
                    int h4 = (int)code_of(load_elem(x, dtype, base + 2 * j), scale, lo, hi) + 8;
                    int l4 = 8;
                    if (2 * j + 1 < D) l4 = (int)code_of(load_elem(x, dtype, base + 2 * j + 1), scale, lo, hi) + 8;
                    po[j] = (uint8_t)((h4 << 4) | l4);
                }
            }
        }
    }
}

/* QuantizedLayerKV.get_kv (ops.py:213-269): dequantize every stored token with its own scale and
 * concatenate along dim 2.  q as written by kvqo_quant_per_token; scales: T elements of scale_dtype. */
void kvqo_dequant_per_token(const void* q, int bits, int64_t BH, int64_t T, int64_t D, const void* scales,
                            int scale_dtype, void* out, int out_dtype) {
    const int64_t P = (D + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < BH; ++r)
        for (int64_t t = 0; t < T; ++t) {
            const float s = load_elem(scales, scale_dtype, t);
            const int64_t ob = (r * T + t) * D;
            if (bits == 8) {
                const int8_t* qi = (const int8_t*)q + ob;
                for (int64_t j = 0; j < D; ++j) store_elem(out, out_dtype, ob + j, (float)qi[j] * s);
            } else {
                const uint8_t* pi = (const uint8_t*)q + (r * T + t) * P;
                for (int64_t j = 0; j < D; ++j) {
                    uint8_t b = pi[j >> 1];
                    int nib = (j & 1) ? (b & 0x0f) : (b >> 4);
                    store_elem(out, out_dtype, ob + j, (float)(nib - 8) * s);
                }
            }
        }
}

/* ------------------------------------------------------------------ area 2: eviction index sets */

enum { KVQO_SLIDING = 0, KVQO_PREFIX_WINDOW = 1, KVQO_STRIDED = 2, KVQO_BLOCK = 3, KVQO_BUDGET = 4 };

/* torch.linspace(start, end, steps) on CPU in float32, element i.  ATen computes
 * step = (end - start) / (steps - 1) and fills the first half as start + step*i and the second half
 * as end - step*(steps-1-i), each with a fused multiply-add (verified against torch in
 * tests/golden/make_golden.py over the budget fixtures).  Used by implementations.py:279-281. */
static inline float linspace_f32(float start, float end, int64_t steps, int64_t i) {
    if (steps == 1) return start;
    const float step = (end - start) / (float)(steps - 1);
    const int64_t half = steps / 2;
    if (i < half) return fmaf(step, (float)i, start);
    return fmaf(-step, (float)(steps - 1 - i), end);
}

/* Index set kept by trim_kv_* for a cache of length T.  Writes at most T (+prefix for the window=0
 * quirk) int64 indices into idx (may be NULL to only count) and returns how many.
 *   sliding        implementations.py:124-140
 *   prefix+window  implementations.py:143-154
 *   strided        implementations.py:157-190
 *   block          implementations.py:193-245
 *   budget         implementations.py:248-292
 * Python slice quirk kept: k[:, :, -0:, :] is the whole tensor, so window 0 keeps all T tokens for
 * sliding (no-op) and appends all T after the prefix for prefix+window. */
int64_t kvqo_evict_indices(int strategy, int64_t T, int64_t window, int64_t prefix, int64_t stride,
                           int64_t block, int64_t keep, int64_t budget, int64_t* idx) {
    int64_t n = 0;
#define PUSH(v) do { if (idx) idx[n] = (v); ++n; } while (0)
    if (strategy == KVQO_SLIDING) {
        if (T > window && window > 0) { for (int64_t t = T - window; t < T; ++t) PUSH(t); }
        else { for (int64_t t = 0; t < T; ++t) PUSH(t); }
        return n;
    }
    if (T <= prefix + window) { for (int64_t t = 0; t < T; ++t) PUSH(t); return n; }
    if (strategy == KVQO_PREFIX_WINDOW) {
        for (int64_t t = 0; t < prefix; ++t) PUSH(t);
        const int64_t ts = window > 0 ? T - window : 0;
        for (int64_t t = ts; t < T; ++t) PUSH(t);
        return n;
    }
    const int64_t tail_start = prefix > T - window ? prefix : T - window;
    for (int64_t t = 0; t < prefix; ++t) PUSH(t);
    if (strategy == KVQO_STRIDED) {
        for (int64_t t = prefix; t < tail_start; t += stride) PUSH(t);
    } else if (strategy == KVQO_BLOCK) {
        int64_t start = prefix;
        while (start < tail_start) {
            int64_t end = start + block < tail_start ? start + block : tail_start;
            int64_t ks = end - keep > start ? end - keep : start;
            for (int64_t t = ks; t < end; ++t) PUSH(t);
            start = end;
        }
    } else if (strategy == KVQO_BUDGET) {
        const int64_t old_len = tail_start - prefix;
        if (old_len > 0 && budget > 0) {
            if (old_len <= budget) {
                for (int64_t t = prefix; t < tail_start; ++t) PUSH(t);
            } else {
                int64_t prev = -1;
                for (int64_t i = 0; i < budget; ++i) {
                    int64_t v = (int64_t)linspace_f32((float)prefix, (float)(tail_start - 1), budget, i);
                    if (i == 0 || v != prev) PUSH(v); /* unique_consecutive, :282 */
                    prev = v;
                }
            }
        }
    }
    for (int64_t t = tail_start; t < T; ++t) PUSH(t);
#undef PUSH
    return n;
}

/* index_select / cat along dim 2: dst[r, j, :] = src[r, idx[j], :], rows of row_bytes bytes. */
void kvqo_gather_rows(const void* src, int64_t BH, int64_t T, int64_t row_bytes, const int64_t* idx,
                      int64_t n, void* dst) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t r = 0; r < BH; ++r)
        for (int64_t j = 0; j < n; ++j)
            memcpy((char*)dst + (r * n + j) * row_bytes, (const char*)src + (r * T + idx[j]) * row_bytes,
                   (size_t)row_bytes);
}

/* ------------------------------------------------------------------ area 3: chunk-summary pooling */

/* chunk_summarize_kv for one tensor, implementations.py:295-346.  x contiguous [BH, T, D]; out is
 * [BH, ceil(old_len/chunk) + keep_eff, D] in the same dtype.  The zero-padded last chunk is divided
 * by chunk (not by its token count), :326-339.  Sum in double, one rounding to fp32, fp32 divide,
 * one rounding to dtype (torch: cast fp32 -> sum -> div -> cast).  Returns the new length, or T with
 * `out` untouched when nothing is compressed (:316-318). */
int64_t kvqo_chunk_mean(const void* x, int dtype, int64_t BH, int64_t T, int64_t D, int64_t chunk,
                        int64_t keep_last, void* out) {
    const int64_t keep_eff = keep_last < T ? keep_last : T;
    const int64_t old_len = T - keep_eff;
    if (old_len <= 0) return T;
    const int64_t n_chunks = (old_len + chunk - 1) / chunk;
    const int64_t Tn = n_chunks + keep_eff;
    if (!out) return Tn;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t r = 0; r < BH; ++r)
        for (int64_t c = 0; c < Tn; ++c) {
            if (c >= n_chunks) { /* recent tokens copied exactly, :322-323,:342 */
                const int64_t t = old_len + (c - n_chunks);
                memcpy((char*)out + (r * Tn + c) * D * esize(dtype),
                       (const char*)x + (r * T + t) * D * esize(dtype), (size_t)(D * esize(dtype)));
                continue;
            }
            const int64_t t0 = c * chunk, t1 = t0 + chunk < old_len ? t0 + chunk : old_len;
            for (int64_t j = 0; j < D; ++j) {
                double acc = 0.0;
                for (int64_t t = t0; t < t1; ++t) acc += (double)load_elem(x, dtype, (r * T + t) * D + j);
                store_elem(out, dtype, (r * Tn + c) * D + j, (float)acc / (float)chunk);
            }
        }
    return Tn;
}

/* ------------------------------------------------------------------ area 4: decode attention */

/* One decode step of eager attention over a quantized cache, fp32 throughout.
 * The reference does not own this arithmetic: benchmarker.py:470-477 hands the dequantized cache to
 * transformers' attention (softmax(q k^T * D^-1/2) v).  This is that formula applied to
 * dequant(codes, scale[t]) -- PARITY UNPINNED against the reference (no fixture exists for it);
 * it is pinned against torch's own fp32 eager attention in tests/golden/make_golden.py instead.
 *   q      [B, Hq, D] fp32           (Hq = G * H; query head hq reads kv head hq / G)
 *   kq     [B, H, T, D] int8  or [B, H, T, ceil(D/2)] uint8 (k_bits 8 / 4); k_scales [T] fp32
 *   vq     likewise with v_bits; v_scales [T] fp32
 *   out    [B, Hq, D] fp32 */
void kvqo_decode_attn(const float* q, const void* kq, int k_bits, const float* k_scales, const void* vq,
                      int v_bits, const float* v_scales, int64_t B, int64_t Hq, int64_t H, int64_t T,
                      int64_t D, float sm_scale, float* out) {
    const int64_t G = Hq / H, P = (D + 1) / 2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t hq = 0; hq < Hq; ++hq) {
            const int64_t h = hq / G;
            const float* qv = q + (b * Hq + hq) * D;
            double* acc = (double*)calloc((size_t)D, sizeof(double));
            float* sc = (float*)malloc((size_t)T * sizeof(float));
            float m = -INFINITY;
            for (int64_t t = 0; t < T; ++t) {
                double dot = 0.0;
                for (int64_t j = 0; j < D; ++j) {
                    float kv;
                    if (k_bits == 8) kv = (float)((const int8_t*)kq)[((b * H + h) * T + t) * D + j];
                    else {
                        uint8_t by = ((const uint8_t*)kq)[((b * H + h) * T + t) * P + (j >> 1)];
                        kv = (float)(((j & 1) ? (by & 0x0f) : (by >> 4)) - 8);
                    }
                    dot += (double)qv[j] * (double)(kv * k_scales[t]);
                }
                sc[t] = (float)dot * sm_scale;
                if (sc[t] > m) m = sc[t];
            }
            double l = 0.0;
            for (int64_t t = 0; t < T; ++t) {
                const double p = exp((double)sc[t] - (double)m);
                l += p;
                for (int64_t j = 0; j < D; ++j) {
                    float vv;
                    if (v_bits == 8) vv = (float)((const int8_t*)vq)[((b * H + h) * T + t) * D + j];
                    else {
                        uint8_t by = ((const uint8_t*)vq)[((b * H + h) * T + t) * P + (j >> 1)];
                        vv = (float)(((j & 1) ? (by & 0x0f) : (by >> 4)) - 8);
                    }
                    acc[j] += p * (double)(vv * v_scales[t]);
                }
            }
            for (int64_t j = 0; j < D; ++j) out[(b * Hq + hq) * D + j] = (float)(acc[j] / l);
            free(acc);
            free(sc);
        }
}

/* ------------------------------------------------------------------ exact-division proof helper */

/* The CUDA quantize kernels replace x / scale by a Markstein-style sequence that needs no division
 * per element:  y = RN(1/scale);  q0 = RN(x*y);  r = fma(-q0, scale, x);  q1 = fma(r, y, q0).
 * This routine checks, for EVERY 16-bit input value (all 65536 bit patterns of `dtype` 1 or 2) and a
 * given absmax bit pattern, that rint(q1) == rint(x / scale) after clamping -- i.e. that the codes
 * are bit-identical to the reference's true division.  Returns the number of mismatching inputs.
 * tests/test_exact_division.py sweeps absmax over all finite 16-bit magnitudes. */
int64_t kvqo_check_division_free(int dtype, uint16_t absmax_bits, float qmax, float eps) {
    const float absmax = dtype == KVQO_F16 ? f16_to_f32(absmax_bits) : bf16_to_f32(absmax_bits);
    const float scale = kvqo_scale_from_absmax(absmax, qmax, eps);
    const float y = 1.0f / scale; /* IEEE RN reciprocal == __frcp_rn */
    int64_t bad = 0;
    for (uint32_t u = 0; u < 65536u; ++u) {
        const float x = dtype == KVQO_F16 ? f16_to_f32((uint16_t)u) : bf16_to_f32((uint16_t)u);
        if (!(fabsf(x) <= absmax)) continue; /* cannot occur in a tensor whose max |x| is absmax */
        const float q0 = x * y;
        const float r = fmaf(-q0, scale, x);
        const float q1 = fmaf(r, y, q0);
        const float lo = qmax == 127.0f ? -127.0f : -8.0f;
        float a = nearbyintf(q1), b = nearbyintf(x / scale);
        if (a < lo) a = lo; if (a > qmax) a = qmax;
        if (b < lo) b = lo; if (b > qmax) b = qmax;
        if (a != b) ++bad;
    }
    return bad;
}
