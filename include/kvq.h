/*
 * kvq.h -- C ABI of libkvq_sm100.so: the B200-native (sm_100a) KV-cache hot path of
 * Efficient-LLM-Inference.
 *
 * This is the drop-in boundary.  The reference's only native interface is the pybind11 module
 * `kvq_ext` JIT-built from an inline string (src/cuda/extensions.py:116-130, reached through
 * get_cuda_extension() :138-147 and called at src/quantization/ops.py:87 and :116-118).  Every entry
 * point below names the reference interface it replaces.  Paths are relative to /root/reference.
 *
 * Conventions
 *  - plain pointers and sizes only: no torch / pybind types.  All data pointers are DEVICE pointers
 *    unless the parameter name ends in `_host`.
 *  - the caller owns all memory; the library never allocates or frees device memory and keeps no
 *    global device state.  Every launch goes to the caller's `stream` (the reference launches on the
 *    legacy default stream, extensions.py:79,105 -- a bug this ABI does not inherit).
 *  - every function returns KVQ_OK (0) or a negative kvq_status; kvq_last_error() returns the
 *    thread-local message of the last failure (the reference raises RuntimeError through TORCH_CHECK,
 *    extensions.py:33-35,72,90,93; the Python wrapper turns non-zero into RuntimeError).
 *  - tensors are 4-D [B, H, T, D] ("batch, kv head, token, head_dim", ops.py:178-179) given as
 *    shape[4] and stride[4] in ELEMENTS; inputs may be views (k[:, :, t:t+1, :], transposed HF
 *    layouts).  Code / output tensors are given with stride3 = {sB, sH, sT} in elements of that
 *    tensor and a contiguous last dimension; stride3 == NULL means contiguous.
 *  - dtype codes: KVQ_F32 / KVQ_F16 / KVQ_BF16.
 *  - scales never leave the device (the reference syncs with float(scale) per call, ops.py:87,117).
 */
#ifndef KVQ_H_
#define KVQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVQ_VERSION 100

#if defined(__GNUC__)
#define KVQ_API __attribute__((visibility("default")))
#else
#define KVQ_API
#endif

typedef void* kvq_stream_t; /* cudaStream_t */

enum kvq_dtype { KVQ_F32 = 0, KVQ_F16 = 1, KVQ_BF16 = 2 };

enum kvq_status {
    KVQ_OK = 0,
    KVQ_ERR_INVALID = -1,     /* bad argument (shape, dtype, alignment the path cannot take) */
    KVQ_ERR_CUDA = -2,        /* a CUDA runtime call / launch failed; message has the CUDA error */
    KVQ_ERR_UNSUPPORTED = -3, /* valid request this build does not implement */
    KVQ_ERR_NO_DEVICE = -4    /* no sm_100 device visible */
};

/* granularity of the quantization scale */
enum kvq_granularity {
    KVQ_PER_TENSOR = 0, /* one scale for the whole tensor: the helpers called directly, ops.py:10-65 */
    KVQ_PER_TOKEN = 1   /* one scale per token t shared by all (b, h, d): what QuantizedKVCache produces by
                           calling the helpers on x[:, :, t:t+1, :] (ops.py:323-342) */
};

/* OR-ed into the `granularity` argument of kvq_quant_int8 / kvq_quant_int4_pack (and into `bits` of kvq_quant_batched):
 * the caller guarantees that absmax_bits is >= max|x| of the data being quantized -- it is the kvq_absmax result of this
 * very data, possibly max-combined with other shards.  Then |x / scale| <= qmax (1 + 2^-23), the reference's clamp
 * (ops.py:29, :50) can never bind, and the kernels skip it (2 of 8 instructions per element).  Without the flag the
 * clamp is applied, so a smaller, "calibrated" absmax clips exactly like the reference formula would. */
#define KVQ_ABSMAX_COVERS_INPUT 0x100

/* eviction strategies of src/cache/implementations.py */
enum kvq_evict_strategy {
    KVQ_EVICT_SLIDING = 0,       /* trim_kv_sliding_window  :124-140 */
    KVQ_EVICT_PREFIX_WINDOW = 1, /* trim_kv_prefix_window   :143-154 */
    KVQ_EVICT_STRIDED = 2,       /* trim_kv_strided         :157-190 */
    KVQ_EVICT_BLOCK = 3,         /* trim_kv_block_old       :193-245 */
    KVQ_EVICT_BUDGET = 4         /* trim_kv_budget_old      :248-292 */
};

/* ---------------------------------------------------------------------------------- library */

KVQ_API int kvq_version(void);
KVQ_API const char* kvq_last_error(void);
/* SM count / compute capability of the current device; KVQ_ERR_NO_DEVICE when it is not sm_100. */
KVQ_API int kvq_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------- area 1: quantize
 * Two-phase pair (north_star area 1).  Phase 1 reduces max|x| as the BIT PATTERN of a non-negative
 * fp32 (uint32 order == float order), so that partial results combine with integer max: across
 * CTAs by atomicMax, across head-sharded GPUs by an allreduce(MAX).  Phase 2 turns the max into the
 * scale and writes codes.
 *
 *   scale  = max(absmax / qmax, eps)           fp32 division     ops.py:28 (qmax 127), :49 (qmax 7)
 *   code   = clamp(rint(x / scale), lo, hi)    half-to-even      ops.py:29, :50
 *   returned scale = scale rounded to x's dtype                  ops.py:30, :65
 */

/* Phase 1.  absmax_bits: 1 word (per tensor) or T words (per token).  The words are max-combined
 * with their current contents when accumulate != 0 (zero them first for a fresh reduction, or chain
 * several shards into one word); otherwise they are overwritten.
 * Replaces: x_fp32.abs().max()  ops.py:26-27, :47-48. */
KVQ_API int kvq_absmax(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
               uint32_t* absmax_bits, int accumulate, kvq_stream_t stream);

/* Phase 2, INT8.  q: int8 codes [B,H,T,D] with q_stride3.  scale_out (nullable): 1 or T elements of
 * x's dtype = scale.to(x.dtype).  scale_f32_out (nullable): the same rounded value widened to fp32
 * (what dequantize multiplies by, ops.py:90).
 * Replaces: quantize_int8_per_tensor  ops.py:10-30. */
KVQ_API int kvq_quant_int8(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
                   const uint32_t* absmax_bits, float eps, int8_t* q, const int64_t q_stride3[3],
                   void* scale_out, float* scale_f32_out, kvq_stream_t stream);

/* Phase 2, INT4 nibble-packed: byte j of a row holds element 2j in the HIGH nibble and 2j+1 in the low
 * nibble, each as code+8; an odd D is padded with one zero code (nibble 8).  packed: uint8
 * [B,H,T,ceil(D/2)].  Replaces: quantize_int4_per_tensor_packed  ops.py:33-65. */
KVQ_API int kvq_quant_int4_pack(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                        int granularity, const uint32_t* absmax_bits, float eps, uint8_t* packed,
                        const int64_t q_stride3[3], void* scale_out, float* scale_f32_out,
                        kvq_stream_t stream);

/* Single-pass per-token quantize: one CTA owns whole tokens, reduces max|x| over (b,h,d) on chip and
 * writes codes without re-reading HBM.  bits = 8 or 4.  Also writes absmax_bits_out[T] when non-NULL.
 * Replaces the per-token loop  QuantizedKVCache.init_from_prompt_past / append_from_past
 * (ops.py:323-342) -> QuantizedLayerKV.append (ops.py:174-210). */
KVQ_API int kvq_quant_per_token_fused(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int bits,
                              float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                              float* scale_f32_out, uint32_t* absmax_bits_out, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 1: dequantize
 *   out = (float(code) * float(scale)) rounded to out_dtype      ops.py:90, :122-133
 * scale: 1 (per tensor) or T (per token) elements of scale_dtype, on the device. */

/* Replaces: dequantize_int8_per_tensor ops.py:68-90 and kvq_ext.dequant_int8_to_fp16
 * (extensions.py:37-48,70-86). */
KVQ_API int kvq_dequant_int8(const int8_t* q, const int64_t shape[4], const int64_t q_stride3[3], const void* scale,
                     int scale_dtype, int granularity, void* out, int out_dtype, kvq_stream_t stream);

/* shape[3] = orig_last_dim (elements per output row); packed rows hold ceil(orig_last/2) bytes.
 * Replaces: dequantize_int4_per_tensor_packed ops.py:93-133 and kvq_ext.dequant_int4_packed_to_fp16
 * (extensions.py:50-68,88-114). */
KVQ_API int kvq_dequant_int4_unpack(const uint8_t* packed, const int64_t shape[4], const int64_t q_stride3[3],
                            const void* scale, int scale_dtype, int granularity, void* out, int out_dtype,
                            kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 1: batched forms
 * n contiguous tensors of `numel` elements each (all layers x {K, V} of a step) in ONE launch per phase, per-tensor
 * granularity.  xs / qs / codes / outs are DEVICE arrays of n device pointers; absmax_bits, scale_out, scale_f32_out and
 * scales hold one entry per tensor.  Same arithmetic as the single-tensor entry points (which they replace when the
 * reference loops `for layer ... quantize(k); quantize(v)`, ops.py:323-342 / benchmarker.py:461,486). */
KVQ_API int kvq_absmax_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, uint32_t* absmax_bits,
                               kvq_stream_t stream);
KVQ_API int kvq_quant_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, int bits,
                              const uint32_t* absmax_bits, float eps, void* const* qs, void* scale_out, float* scale_f32_out,
                              kvq_stream_t stream);
KVQ_API int kvq_dequant_batched(const void* const* codes, int bits, int64_t n_tensors, int64_t numel, const void* scales,
                                int scale_dtype, void* const* outs, int out_dtype, kvq_stream_t stream);

/* Per-token single-pass quantize of n same-shape tensors (all layers' K streams, or all V streams, of a step) in ONE
 * launch: tensor i is read from xs[i] ([B,H,T,D] with the common `stride`) and written to rows
 * [dst_token_offset, dst_token_offset + T) of the store qs[i] (common q_stride3, bytes), its scales to
 * scales_out[i][dst_token_offset + t] (x's dtype) and scales_f32_out[i][dst_token_offset + t].  xs / qs / scales_out /
 * scales_f32_out are DEVICE arrays of n device pointers (the last two nullable); every pointer 16-byte aligned.
 * KVQ_ERR_UNSUPPORTED for layouts the vector path cannot take (call kvq_quant_per_token_fused per tensor then).
 * Replaces the loop `for layer: append(k[..., -1:, :], v[..., -1:, :])` / `for t: append(...)`, ops.py:323-342. */
KVQ_API int kvq_quant_per_token_fused_batched(const void* const* xs, int dtype, int64_t n_tensors, const int64_t shape[4],
                                              const int64_t stride[4], int bits, float eps, void* const* qs,
                                              const int64_t q_stride3[3], int64_t dst_token_offset, void* const* scales_out,
                                              float* const* scales_f32_out, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- multi-GPU: the absmax exchange
 * A head-sharded cache exchanges one thing, between phase 1 and phase 2: the absmax words (max over ranks; reference scale
 * over all heads: ops.py:27,:48,:330,:342).  kvq_absmax_exchange does it without a collective library: peer_buffers is a
 * DEVICE array of world_size device pointers, entry r = rank r's exchange buffer (kvq_absmax_exchange_buffer_bytes() bytes,
 * zero-initialised once, writable by every peer: e.g. torch symmetric memory or cudaIpc / VMM mappings).  One small kernel
 * per rank writes its words to every peer, publishes `seq`, waits for all peers' `seq` and max-combines in place.  `seq`: a
 * non-zero step counter, the same on every rank, increasing by one per call.  n_words <= 1024, world_size <= 16.  Results
 * are bit-identical to an allreduce(MAX).  (NCCL's allreduce stays available: kvq_b200.distributed.allreduce_absmax_.) */
KVQ_API int64_t kvq_absmax_exchange_buffer_bytes(void);
KVQ_API int kvq_absmax_exchange(uint32_t* absmax_bits, int64_t n_words, void* const* peer_buffers, int rank, int world_size,
                                uint32_t seq, kvq_stream_t stream);

/* The same exchange FUSED into the two kernels it sits between (no launch, no stream hand-over for it): 
 *   kvq_absmax_batched_publish = kvq_absmax_batched whose last CTA per tensor stores the tensor's finished word, tagged with
 *     `seq`, into every rank's buffer (one 8-byte NVLink store per peer); absmax_bits receives the LOCAL words as before;
 *   kvq_quant_batched_gather   = kvq_quant_batched whose CTAs first wait for all world_size words of their tensor for step
 *     `seq` in local_buffer (= this rank's own exchange buffer) and max-combine them; absmax_bits_out (nullable) receives
 *     the combined words.
 * Same buffers as kvq_absmax_exchange (a separate region; the two `seq` counters are independent), same rules for `seq`
 * (non-zero, identical on all ranks, +1 per publish/gather pair), n_tensors <= 1024.  Every rank must call both, in this
 * order, every step; the wait is bounded (a peer that never publishes traps the kernel after ~10 s).  Codes and scales are
 * bit-identical to the allreduce form.  Replaces the same reference lines as kvq_absmax_batched / kvq_quant_batched for a
 * head-sharded cache (ops.py:27,:48 with the scale over all heads). */
KVQ_API int kvq_absmax_batched_publish(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, uint32_t* absmax_bits,
                                       void* const* peer_buffers, int rank, int world_size, uint32_t seq, kvq_stream_t stream);
KVQ_API int kvq_quant_batched_gather(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, int bits,
                                     uint32_t* absmax_bits_out, float eps, void* const* qs, void* scale_out, float* scale_f32_out,
                                     void* local_buffer, int world_size, uint32_t seq, kvq_stream_t stream);

/* The per-token append of a head-sharded cache in ONE kernel: kvq_quant_per_token_fused whose CTAs, between the token's on-chip
 * max and the quantization of the values they hold in registers, store the token's word (tagged with `seq`) into every rank's
 * buffer and wait for the world's words of the same token.  What a decode step appends (one token of K, one of V per layer)
 * costs one launch per tensor instead of absmax kernel + allreduce + quantize kernel.  At most 1024 tokens per call; every rank
 * calls it with the same token count and `seq`, in the same order, on one stream.  absmax_bits_out (nullable) receives the
 * combined words.  Scales and codes are bit-identical to the allreduce form (ops.py:330,:342 with the scale over all heads).
 * seq == 0: the step number is taken from a counter in the rank's own buffer (steps completed + 1; the last CTA of the launch
 * records the completion), so the call can be captured into a CUDA graph and replayed -- the form the cache stores use; such
 * calls use their own region of the buffer, so they may be mixed with seq != 0 calls and with the publish / gather pair.
 * KVQ_ERR_UNSUPPORTED for layouts the vector path cannot take (custom eps, rows that are not 16-byte multiples), and for
 * seq != 0 -- like every entry point that takes a host step number -- while the stream is being captured into a CUDA graph
 * (a replay would present the same number again). */
KVQ_API int kvq_quant_per_token_fused_exchange(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                                               int bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                                               float* scale_f32_out, uint32_t* absmax_bits_out, void* const* peer_buffers, int rank,
                                               int world_size, uint32_t seq, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 2: eviction
 * The kept index set is a closed form of (T, parameters): the gather kernel computes the source token
 * of every destination row in registers; no index tensor and no intermediate parts list exists. */

typedef struct kvq_evict_plan {
    int32_t strategy;
    int32_t identity;     /* 1: cache unchanged (T <= prefix + window, or sliding with T <= window) */
    int64_t T;            /* source length */
    int64_t n_prefix;     /* rows copied from [0, n_prefix) */
    int64_t n_old;        /* rows taken from the older region [old_start, tail_start) */
    int64_t old_start;    /* = prefix_len */
    int64_t tail_start;   /* rows [tail_start, tail_start + n_tail) copied at the end */
    int64_t n_tail;
    int64_t n_total;      /* = n_prefix + n_old + n_tail : the new cache length */
    int64_t stride;       /* strided: step; block: block_size */
    int64_t keep;         /* block: keep_per_block */
    int64_t budget;       /* budget: linspace steps (0 when the whole older region is kept) */
    int32_t needs_table;  /* budget only: float32 linspace produced duplicates that
                             unique_consecutive removed -> the older region needs kvq_evict_indices_host
                             + table mode (never seen for budget < old_len / 2) */
    int32_t reserved;
} kvq_evict_plan;

/* Host, pure arithmetic (no CUDA): fills *plan for a cache of length T.  Returns KVQ_ERR_INVALID for
 * stride < 1, block < 1, keep outside [1, block], budget < 0, negative window / prefix (the
 * reference asserts the first four: implementations.py:164,207-208,257). */
KVQ_API int kvq_evict_plan_make(int strategy, int64_t T, int64_t window, int64_t prefix, int64_t stride, int64_t block,
                        int64_t keep, int64_t budget, kvq_evict_plan* plan);

/* Host: writes plan->n_total kept source indices (the reference's idx / cat order). */
KVQ_API int kvq_evict_indices_host(const kvq_evict_plan* plan, int64_t* idx_host);

/* Device: the same indices computed by the device function the gather kernel uses (parity probe). */
KVQ_API int kvq_evict_indices_device(const kvq_evict_plan* plan, const int32_t* old_table, int64_t* idx,
                             kvq_stream_t stream);

/* dst[b,h,j,:] = src[b,h,index(j),:] for j < plan->n_total.  elem_bytes = bytes per element of the rows
 * (2 for fp16/bf16, 4 fp32, 1 for INT8 codes / packed INT4 bytes: the same kernel compacts a quantized
 * store).  old_table (nullable, int32[n_old]): source tokens of the older region when
 * plan->needs_table.  Replaces: trim_kv_* implementations.py:124-292 (slicing, arange,
 * index_select, linspace, unique_consecutive, cat). */
KVQ_API int kvq_evict_gather(const void* src, int elem_bytes, const int64_t shape[4], const int64_t stride[4],
                     const kvq_evict_plan* plan, const int32_t* old_table, void* dst,
                     const int64_t dst_stride3[3], kvq_stream_t stream);

/* kvq_evict_gather for n same-shape tensors (all layers x {K, V} of a step) in ONE launch: srcs / dsts are DEVICE arrays
 * of n device pointers, 16-byte aligned; shape, stride, plan and dst_stride3 are common.  The 256-320-row outputs of
 * budget / prefix+window are latency-bound as separate launches; batched they stream.  KVQ_ERR_UNSUPPORTED for layouts
 * the vector path cannot take.  Replaces the per-layer loops implementations.py:176-190, :218-245, :268-292. */
KVQ_API int kvq_evict_gather_batched(const void* const* srcs, void* const* dsts, int64_t n_tensors, int elem_bytes,
                                     const int64_t shape[4], const int64_t stride[4], const kvq_evict_plan* plan,
                                     const int32_t* old_table, const int64_t dst_stride3[3], kvq_stream_t stream);

/* Device: builds old_table (int32 [plan->n_old]) for a budget plan -- trunc(float32 linspace) with consecutive
 * duplicates removed (implementations.py:279-282) -- so that the indices never exist on the host. */
KVQ_API int kvq_evict_old_table_device(const kvq_evict_plan* plan, int32_t* old_table, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 2: paged store
 * Physical pool [n_blocks, B, H, block_size, D]; block_table[i] = physical block of logical block i.
 * The reference keeps a Python list of [B,H,block,D] tensors whose list index is the logical block
 * (implementations.py:36-54). */

/* Writes the n = shape[2] tokens of x ([B,H,n,D], strided) at logical positions t .. t+n-1: token t+j goes to slot
 * (t+j) % block_size of logical block (t+j) / block_size (every block touched must be in the table).  n = 1 is
 * PagedKVCache.append (implementations.py:57-80); n > 1 replaces the per-token prompt loop benchmarker.py:528-530. */
KVQ_API int kvq_paged_append(const void* x, int elem_bytes, const int64_t shape[4], const int64_t stride[4], void* pool,
                     const int32_t* block_table, int64_t block_size, int64_t t, kvq_stream_t stream);

/* Stitches the first T tokens into contiguous out [B,H,T,D].  shape = {B,H,T,D}.
 * Replaces: PagedKVCache.get_kv implementations.py:83-106. */
KVQ_API int kvq_paged_gather(const void* pool, int elem_bytes, const int64_t shape[4], const int32_t* block_table,
                     int64_t block_size, void* out, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 3: chunk summary
 * out[b,h,c,:]            = sum_{t in chunk c of [0,old_len)} x[b,h,t,:] / chunk_size   (c < n_chunks)
 * out[b,h,n_chunks + r,:] = x[b,h,old_len + r,:]                                        (r < keep_eff)
 * with keep_eff = min(keep_last, T), old_len = T - keep_eff, n_chunks = ceil(old_len / chunk).  The last,
 * zero-padded chunk is divided by chunk_size like the others (implementations.py:326-339).  fp32
 * accumulation, one rounding to the dtype.  out is contiguous [B,H,n_chunks+keep_eff,D].
 * Replaces: chunk_summarize_kv implementations.py:295-346 (pad-cat, view, mean, cat). */
KVQ_API int kvq_chunk_mean(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int64_t chunk,
                   int64_t keep_last, void* out, kvq_stream_t stream);

/* kvq_chunk_mean for n same-shape tensors in ONE launch (xs / outs: DEVICE arrays of n 16-byte-aligned device pointers).
 * Replaces the per-layer loop implementations.py:316-346. */
KVQ_API int kvq_chunk_mean_batched(const void* const* xs, void* const* outs, int64_t n_tensors, int dtype, const int64_t shape[4],
                                   const int64_t stride[4], int64_t chunk, int64_t keep_last, kvq_stream_t stream);

/* ---------------------------------------------------------------------------------- area 4: decode attention
 * One decode step over a per-token-quantized cache, dequantizing K/V in registers (split-K GEMV, no
 * tensor cores):   out[b,hq,:] = softmax_t( q[b,hq,:] . K[b,hq/G,t,:] * sm_scale ) V[b,hq/G,t,:]
 * with K[.,t,:] = k_codes * k_scale[t], V likewise.  The reference hands the dequantized cache to
 * transformers' attention (benchmarker.py:470-477); this entry replaces dequantize-everything +
 * DynamicCache + attention for the cached tokens.
 *   q        [B, Hq, D]   q_dtype, contiguous
 *   k_codes  k_bits 8: int8 [B,H,T,D];  4: uint8 [B,H,T,D/2]  (k_stride3 in bytes of that tensor)
 *   k_scale  fp32 [T]  (the rounded scales, see kvq_quant_* scale_f32_out)
 *   v_codes / v_scale likewise
 *   out      [B, Hq, D]   out_dtype, contiguous
 *   workspace: kvq_decode_attn_workspace_bytes(...) bytes, n_splits as returned there
 * The tiled kernels take head_dim 64 / 128, Hq / H in {1, 2, 4, 8} and 16-byte-aligned code rows; any other head_dim (1..512),
 * grouping or alignment runs a generic (correct, not bandwidth-bound) kernel, so the entry accepts every model the reference's
 * path accepts (benchmarker.py:473-477 hands the cache to transformers' attention).  INT4 rows hold ceil(D / 2) bytes. */
KVQ_API int64_t kvq_decode_attn_workspace_bytes(int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, int* n_splits);

KVQ_API int kvq_decode_attn(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                    const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                    const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                    void* out, int out_dtype, void* workspace, int n_splits, kvq_stream_t stream);

/* kvq_decode_attn plus the decode step's own token, exact: k_cur / v_cur are [B, H, D] of q's dtype (both NULL = plain
 * kvq_decode_attn).  The reference's step attends over dequant(cached tokens) ++ the new token's unquantized K/V
 * (benchmarker.py:473-477) and only then quantizes it (:486); the new token enters the softmax merge as one more split. */
KVQ_API int kvq_decode_attn_ex(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                               const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                               const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                               const void* k_cur, const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits,
                               kvq_stream_t stream);

/* Kernel families of the decode attention.  KVQ_ATTN_CUDA_CORE: the split-K GEMV on CUDA cores as north_star words it
 * (per-warp DP4A for Hq/H = 1, register-tiled FFMA2 for 2, 4, 8).  KVQ_ATTN_IMMA: the same split / merge protocol with the
 * two contractions on the INT8 tensor-core MMA, tiles staged by tensor-map TMA (csrc/attn_mma.cu): at Hq/H > 1 the CUDA-core
 * kernels are FMA-bound, not HBM-bound, so KVQ_ATTN_AUTO picks IMMA there (kvq_decode_attn_default_variant).  Same results
 * within the stated attention tolerance; both are always available for A/B measurements. */
enum kvq_attn_variant { KVQ_ATTN_AUTO = 0, KVQ_ATTN_CUDA_CORE = 1, KVQ_ATTN_IMMA = 2 };
KVQ_API int kvq_decode_attn_default_variant(int64_t q_heads_per_kv_head);

/* kvq_decode_attn_ex with an explicit kernel family. */
KVQ_API int kvq_decode_attn_variant(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                                    const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                                    const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                                    const void* k_cur, const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits,
                                    int variant, kvq_stream_t stream);

/* Same, over a PAGED quantized store (the reference's PagedKVCache only simulates paging and re-stitches the
 * blocks with torch.cat before every use, src/cache/implementations.py:83-106): the kernel reads K/V blocks through the
 * block table, no stitch copy.
 *   k_pool / v_pool     : code pools [n_phys][B][H][block_size][D*bits/8]; *_pool_stride3 = byte strides {block, batch, head}
 *                         (NULL: contiguous); the rows of one (block, b, h) are back to back
 *   k/v_scale_pool      : float [n_phys][block_size], one scale per token
 *   block_table         : device int32 [ceil(T / block_size)]: logical block -> physical block
 *   block_size          : power of two >= 16
 * Hq / H in {1, 2, 4, 8}.  Workspace / n_splits as for kvq_decode_attn.  Results are bit-identical to kvq_decode_attn_ex on the
 * same tokens stored contiguously. */
KVQ_API int kvq_decode_attn_paged(const void* q, int q_dtype, const void* k_pool, int k_bits, const int64_t k_pool_stride3[3],
                          const float* k_scale_pool, const void* v_pool, int v_bits, const int64_t v_pool_stride3[3],
                          const float* v_scale_pool, const int32_t* block_table, int64_t block_size, int64_t B, int64_t Hq,
                          int64_t H, int64_t T, int64_t D, float sm_scale, const void* k_cur, const void* v_cur, void* out,
                          int out_dtype, void* workspace, int n_splits, kvq_stream_t stream);

/* kvq_decode_attn_paged with an explicit kernel family. */
KVQ_API int kvq_decode_attn_paged_variant(const void* q, int q_dtype, const void* k_pool, int k_bits, const int64_t k_pool_stride3[3],
                                          const float* k_scale_pool, const void* v_pool, int v_bits, const int64_t v_pool_stride3[3],
                                          const float* v_scale_pool, const int32_t* block_table, int64_t block_size, int64_t B,
                                          int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale, const void* k_cur,
                                          const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits, int variant,
                                          kvq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KVQ_H_ */
