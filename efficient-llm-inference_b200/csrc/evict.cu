// evict.cu -- area 2: eviction strategies as index-computing gather/compaction kernels, and the
// paged block-table store (sm_100a).
//
// Reference: src/cache/implementations.py (/root/reference)
//   trim_kv_sliding_window :124-140, trim_kv_prefix_window :143-154, trim_kv_strided :157-190,
//   trim_kv_block_old :193-245, trim_kv_budget_old :248-292, PagedKVCache :10-121.
//
// The kept set is  [0, n_prefix)  ++  old(j)  ++  [tail_start, tail_start + n_tail)  where old(j) is a
// closed form of the strategy.  The gather kernel evaluates it per destination row in registers, so no
// index tensor, arange, linspace or parts list ever exists in memory and the source cache is read
// exactly once for the kept rows only: 2 * kept * row_bytes of HBM traffic.
#include <cmath>
#include <cstdlib>

#include "kvq_common.cuh"

namespace kvq {
namespace {

constexpr int kThreads = 256;
#ifndef KVQ_GATHER_CTAS_PER_SM
#define KVQ_GATHER_CTAS_PER_SM 128
#endif
constexpr int kGatherCtasPerSm = KVQ_GATHER_CTAS_PER_SM;  // compile-time tuning knob (r1c sweep)

// device copy of the plan, 32-bit where the kernel needs speed (T <= INT32_MAX is enforced)
struct PlanDev {
    int strategy;
    int n_prefix, n_old, old_start, tail_start, n_tail, n_total;
    int stride, keep, budget;
    int n_full_rows;  // block: rows contributed by full blocks = (old_len / block) * keep
    int n_full;       // block: number of full blocks
    float lin_start, lin_end, lin_step;  // budget: float32 linspace(prefix, tail_start-1, budget)
    int use_table;
};

// torch.linspace(start, end, steps)[i] in float32 (ATen RangeFactories: symmetric fill with fused
// multiply-add from both ends), then .long() truncation.  implementations.py:279-281.
__host__ __device__ inline int linspace_index(float start, float end, float step, int steps, int i) {
    if (steps == 1) return (int)start;
    const int half = steps / 2;
#ifdef __CUDA_ARCH__
    const float v = i < half ? __fmaf_rn(step, (float)i, start) : __fmaf_rn(-step, (float)(steps - 1 - i), end);
    return __float2int_rz(v);
#else
    const float v = i < half ? std::fmaf(step, (float)i, start) : std::fmaf(-step, (float)(steps - 1 - i), end);
    return (int)v;
#endif
}

__host__ __device__ inline int old_token(const PlanDev& p, int j, const int32_t* __restrict__ table) {
    switch (p.strategy) {
        case KVQ_EVICT_STRIDED:
            return p.old_start + j * p.stride;
        case KVQ_EVICT_BLOCK: {
            if (j < p.n_full_rows) {
                const int blk = j / p.keep, r = j - blk * p.keep;
                return p.old_start + blk * p.stride + (p.stride - p.keep) + r;  // last `keep` of a full block
            }
            const int start = p.old_start + p.n_full * p.stride, end = p.tail_start;
            const int ks = end - p.keep > start ? end - p.keep : start;
            return ks + (j - p.n_full_rows);
        }
        case KVQ_EVICT_BUDGET:
            if (p.budget == 0) return p.old_start + j;  // old_len <= budget: whole older region
            if (p.use_table) return table[j];
            return linspace_index(p.lin_start, p.lin_end, p.lin_step, p.budget, j);
        default:
            return p.old_start + j;
    }
}

__host__ __device__ inline int src_token(const PlanDev& p, int j, const int32_t* __restrict__ table) {
    if (j < p.n_prefix) return j;
    j -= p.n_prefix;
    if (j >= p.n_old) return p.tail_start + (j - p.n_old);
    return old_token(p, j, table);
}

PlanDev to_dev(const kvq_evict_plan& pl, bool have_table) {
    PlanDev d{};
    d.strategy = pl.strategy;
    d.n_prefix = (int)pl.n_prefix; d.n_old = (int)pl.n_old; d.old_start = (int)pl.old_start;
    d.tail_start = (int)pl.tail_start; d.n_tail = (int)pl.n_tail; d.n_total = (int)pl.n_total;
    d.stride = (int)pl.stride; d.keep = (int)pl.keep; d.budget = (int)pl.budget;
    if (pl.strategy == KVQ_EVICT_BLOCK && pl.stride > 0) {
        const int64_t old_len = pl.tail_start - pl.old_start;
        d.n_full = (int)(old_len / pl.stride);
        d.n_full_rows = d.n_full * d.keep;
    }
    if (pl.strategy == KVQ_EVICT_BUDGET && pl.budget > 1) {
        d.lin_start = (float)pl.old_start;
        d.lin_end = (float)(pl.tail_start - 1);
        d.lin_step = (d.lin_end - d.lin_start) / (float)(pl.budget - 1);
    } else if (pl.strategy == KVQ_EVICT_BUDGET) {
        d.lin_start = (float)pl.old_start; d.lin_end = (float)(pl.tail_start - 1); d.lin_step = 0.f;
    }
    d.use_table = pl.needs_table && have_table;
    return d;
}

// ---- kernels ---------------------------------------------------------------------------------------------

struct CopyGeom {
    int nP, H;            // planes, heads (to split p into b,h)
    int64_t sB, sH, sT;   // source strides in BYTES
    int64_t dB, dH, dT;   // destination strides in BYTES
    int cpr;              // 16-byte chunks per row
    int shift;            // log2(cpr) when it is a power of two, else -1
};

// dst[p, j, :] = src[p, src_token(j), :].  A CTA owns a tile of destination rows: their source tokens are
// evaluated once (one thread per row) into shared memory and reused for every plane (b, h) the CTA walks,
// so the copy loop is LDS + two address IMADs + LDG.128 + STG.128 per 16 bytes, UNROLL loads in flight.
// Batched form (srcs != nullptr): blockIdx.z is the tensor (all layers x {K, V} of a step share shape, strides and plan), so
// the 256-320-row outputs of budget / prefix+window -- latency-bound as 64 separate 6 us launches -- become one launch.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads) evict_gather_kernel(const char* __restrict__ src, char* __restrict__ dst,
                                                               CopyGeom g, PlanDev pl, const int32_t* __restrict__ table,
                                                               const void* const* __restrict__ srcs = nullptr,
                                                               void* const* __restrict__ dsts = nullptr) {
    __shared__ int src_row[kThreads * UNROLL];
    if (srcs) { src = reinterpret_cast<const char*>(srcs[blockIdx.z]); dst = reinterpret_cast<char*>(dsts[blockIdx.z]); }
    int c, r0, rpp;
    if (g.shift >= 0) { c = threadIdx.x & (g.cpr - 1); r0 = threadIdx.x >> g.shift; rpp = kThreads >> g.shift; }
    else { c = threadIdx.x % g.cpr; r0 = threadIdx.x / g.cpr; rpp = kThreads / g.cpr; }
    const bool active = r0 < rpp;
    const int tile = rpp * UNROLL;
    for (int j0 = blockIdx.x * tile; j0 < pl.n_total; j0 += gridDim.x * tile) {
        __syncthreads();
        for (int i = threadIdx.x; i < tile; i += kThreads) src_row[i] = j0 + i < pl.n_total ? src_token(pl, j0 + i, table) : -1;
        __syncthreads();
        if (!active) continue;
        int st[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st[u] = src_row[r0 + u * rpp];
        for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
            const int b = g.H == 1 ? p : p / g.H, h = p - b * g.H;
            const char* s = src + b * g.sB + h * g.sH + (int64_t)c * 16;
            char* d = dst + b * g.dB + h * g.dH + (int64_t)c * 16 + (int64_t)(j0 + r0) * g.dT;
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (st[u] >= 0) v[u] = ld_stream(s + (int64_t)st[u] * g.sT);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (st[u] >= 0) st_stream(d + (int64_t)(u * rpp) * g.dT, v[u]);
        }
    }
}

// any row size / alignment / innermost stride: one thread per element of `ES` bytes
template <typename E>
__global__ void __launch_bounds__(kThreads) evict_gather_generic_kernel(const E* __restrict__ src, E* __restrict__ dst,
                                                                       int64_t B, int64_t H, int64_t D, int64_t sB,
                                                                       int64_t sH, int64_t sT, int64_t sD, int64_t dB,
                                                                       int64_t dH, int64_t dT, PlanDev pl,
                                                                       const int32_t* __restrict__ table) {
    const int64_t n = B * H * (int64_t)pl.n_total * D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % D, r = i / D, j = r % pl.n_total, ph = r / pl.n_total, h = ph % H, b = ph / H;
        dst[b * dB + h * dH + j * dT + d] = src[b * sB + h * sH + (int64_t)src_token(pl, (int)j, table) * sT + d * sD];
    }
}

__global__ void evict_indices_kernel(PlanDev pl, const int32_t* __restrict__ table, int64_t* __restrict__ idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < pl.n_total) idx[j] = src_token(pl, j, table);
}

// ---- paged store -------------------------------------------------------------------------------------------
// pool [n_blocks, B, H, block, D]; logical token t lives in physical block table[t / block], slot t % block.

struct PagedGeom {
    int nP;               // B*H
    int64_t T;            // tokens to stitch
    int block;            // tokens per block
    int64_t row_bytes;    // D * elem
    int cpr;
};

template <int UNROLL>
__global__ void __launch_bounds__(kThreads) paged_gather_kernel(const char* __restrict__ pool, char* __restrict__ out,
                                                               PagedGeom g, const int32_t* __restrict__ table) {
    const int c = threadIdx.x % g.cpr, r0 = threadIdx.x / g.cpr, rpp = kThreads / g.cpr;
    if (r0 >= rpp) return;
    const int64_t plane_bytes = (int64_t)g.block * g.row_bytes;      // one (b,h) inside a block
    const int64_t block_bytes = (int64_t)g.nP * plane_bytes;
    for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
        const int64_t t0 = (int64_t)blockIdx.x * rpp * UNROLL + r0;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t < g.T) {
                const int64_t lb = t / g.block, slot = t - lb * g.block;
                v[u] = ld_stream(pool + (int64_t)table[lb] * block_bytes + p * plane_bytes + slot * g.row_bytes + (int64_t)c * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t < g.T) st_stream(out + ((int64_t)p * g.T + t) * g.row_bytes + (int64_t)c * 16, v[u]);
        }
    }
}

// generic element-wise variants (any D / alignment)
template <typename E>
__global__ void __launch_bounds__(kThreads) paged_gather_generic_kernel(const E* __restrict__ pool, E* __restrict__ out,
                                                                       int64_t nP, int64_t T, int64_t D, int64_t block,
                                                                       const int32_t* __restrict__ table) {
    const int64_t n = nP * T * D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % D, r = i / D, t = r % T, p = r / T, lb = t / block, slot = t - lb * block;
        out[i] = pool[(((int64_t)table[lb] * nP + p) * block + slot) * D + d];
    }
}

unsigned gen_grid(int64_t n) {
    const int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 32;
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace kvq

using namespace kvq;

// ------------------------------------------------------------------------------------------------ host: plan

extern "C" int kvq_evict_plan_make(int strategy, int64_t T, int64_t window, int64_t prefix, int64_t stride, int64_t block,
                                   int64_t keep, int64_t budget, kvq_evict_plan* plan) {
    KVQ_REQUIRE(plan != nullptr, "plan is null");
    KVQ_REQUIRE(strategy >= KVQ_EVICT_SLIDING && strategy <= KVQ_EVICT_BUDGET, "unknown strategy %d", strategy);
    KVQ_REQUIRE(T >= 0 && T <= 0x7fffffff, "T out of range");
    KVQ_REQUIRE(window >= 0 && prefix >= 0, "window and prefix_len must be >= 0");
    KVQ_REQUIRE(window <= 0x7fffffff && prefix <= 0x7fffffff, "window / prefix_len out of range");
    if (strategy == KVQ_EVICT_STRIDED) KVQ_REQUIRE(stride >= 1, "stride must be >= 1");  // implementations.py:164
    if (strategy == KVQ_EVICT_BLOCK) {
        KVQ_REQUIRE(block >= 1, "block_size must be >= 1");                               // :207
        KVQ_REQUIRE(keep >= 1 && keep <= block, "need 1 <= keep_per_block <= block_size"); // :208
    }
    if (strategy == KVQ_EVICT_BUDGET) KVQ_REQUIRE(budget >= 0, "old_budget must be >= 0");  // :257
    kvq_evict_plan p{};
    p.strategy = strategy;
    p.T = T;
    auto identity = [&]() {
        p.identity = 1; p.n_prefix = 0; p.n_old = 0; p.old_start = 0; p.tail_start = 0; p.n_tail = T; p.n_total = T;
        *plan = p;
        return KVQ_OK;
    };
    if (strategy == KVQ_EVICT_SLIDING) {
        // k[:, :, -w:, :] only when T > w (:136-138); -0: is the whole tensor
        if (!(T > window) || window == 0) return identity();
        p.tail_start = T - window; p.n_tail = window; p.n_total = window;
        *plan = p;
        return KVQ_OK;
    }
    if (T <= prefix + window) return identity();  // :148, :168, :213, :262
    if (strategy == KVQ_EVICT_PREFIX_WINDOW) {
        p.n_prefix = prefix; p.old_start = prefix;
        if (window == 0) { p.tail_start = 0; p.n_tail = T; }  // cat(k[:, :, :p], k[:, :, -0:]) :151
        else { p.tail_start = T - window; p.n_tail = window; }
        p.n_total = p.n_prefix + p.n_tail;
        *plan = p;
        return KVQ_OK;
    }
    const int64_t tail_start = prefix > T - window ? prefix : T - window;  // :172
    const int64_t old_len = tail_start - prefix;
    p.n_prefix = prefix; p.old_start = prefix; p.tail_start = tail_start; p.n_tail = T - tail_start;
    if (strategy == KVQ_EVICT_STRIDED) {
        p.stride = stride;
        p.n_old = old_len > 0 ? (old_len + stride - 1) / stride : 0;  // arange(p, tail_start, stride) :181
    } else if (strategy == KVQ_EVICT_BLOCK) {
        p.stride = block; p.keep = keep;
        const int64_t rem = old_len % block;
        p.n_old = (old_len / block) * keep + (rem < keep ? rem : keep);  // :229-233
    } else {  // budget
        if (old_len > 0 && budget > 0) {
            if (old_len <= budget) { p.n_old = old_len; p.budget = 0; }  // :276-277
            else {
                p.budget = budget;
                const float s = (float)prefix, e = (float)(tail_start - 1);
                const float step = budget > 1 ? (e - s) / (float)(budget - 1) : 0.f;
                int64_t n = 0; int prev = -1;
                for (int64_t i = 0; i < budget; ++i) {
                    const int v = linspace_index(s, e, step, (int)budget, (int)i);
                    if (i == 0 || v != prev) ++n;  // unique_consecutive :282
                    prev = v;
                }
                p.n_old = n;
                p.needs_table = n != budget;
            }
        }
    }
    p.n_total = p.n_prefix + p.n_old + p.n_tail;
    *plan = p;
    return KVQ_OK;
}

extern "C" int kvq_evict_indices_host(const kvq_evict_plan* plan, int64_t* idx_host) {
    KVQ_REQUIRE(plan && idx_host, "null pointer");
    PlanDev d = to_dev(*plan, false);
    int64_t n = 0;
    for (int64_t j = 0; j < plan->n_prefix; ++j) idx_host[n++] = j;
    if (plan->strategy == KVQ_EVICT_BUDGET && plan->budget > 0) {
        int prev = -1;
        for (int64_t i = 0; i < plan->budget; ++i) {
            const int v = linspace_index(d.lin_start, d.lin_end, d.lin_step, d.budget, (int)i);
            if (i == 0 || v != prev) idx_host[n++] = v;
            prev = v;
        }
    } else {
        for (int64_t j = 0; j < plan->n_old; ++j) idx_host[n++] = old_token(d, (int)j, nullptr);
    }
    for (int64_t j = 0; j < plan->n_tail; ++j) idx_host[n++] = plan->tail_start + j;
    if (n != plan->n_total) return fail(KVQ_ERR_INVALID, "inconsistent plan: produced %lld of %lld indices", (long long)n, (long long)plan->n_total);
    return KVQ_OK;
}

extern "C" int kvq_evict_indices_device(const kvq_evict_plan* plan, const int32_t* old_table, int64_t* idx, kvq_stream_t stream) {
    KVQ_REQUIRE(plan && idx, "null pointer");
    KVQ_REQUIRE(!plan->needs_table || old_table, "plan needs old_table (float32 linspace produced duplicates)");
    if (plan->n_total == 0) return KVQ_OK;
    const PlanDev d = to_dev(*plan, old_table != nullptr);
    evict_indices_kernel<<<(unsigned)((plan->n_total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, old_table, idx);
    return check_launch("evict_indices");
}

namespace kvq { namespace {
// common body of kvq_evict_gather (n_tensors == 1, src / dst) and kvq_evict_gather_batched (device pointer tables)
int run_gather(const void* src, void* dst, const void* const* srcs, void* const* dsts, int64_t n_tensors, bool ptrs_aligned,
               int elem_bytes, const int64_t shape[4], const int64_t stride[4], const kvq_evict_plan* plan,
               const int32_t* old_table, const int64_t dst_stride3[3], cudaStream_t st) {
    KVQ_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 1, 2 or 4");
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && D >= 0, "negative shape");
    KVQ_REQUIRE(T == plan->T, "plan was made for T=%lld, tensor has T=%lld", (long long)plan->T, (long long)T);
    KVQ_REQUIRE(!plan->needs_table || old_table, "plan needs old_table (float32 linspace produced duplicates)");
    const int64_t n_out = plan->n_total;
    if (B * H * n_out * D == 0 || n_tensors == 0) return KVQ_OK;
    const Stride3 ds = make_stride3(dst_stride3, H, n_out, D);
    const PlanDev pd = to_dev(*plan, old_table != nullptr);
    const int64_t rb = D * elem_bytes;
    auto a16 = [&](int64_t s) { return (s * elem_bytes) % 16 == 0; };
    const bool fast = stride[3] == 1 && rb % 16 == 0 && rb / 16 <= kThreads && ptrs_aligned && a16(stride[0]) &&
                      a16(stride[1]) && a16(stride[2]) && a16(ds.sB) && a16(ds.sH) && a16(ds.sT) && B * H <= 0x7fffffff;
    if (fast) {
        CopyGeom g{(int)(B * H), (int)H, stride[0] * elem_bytes, stride[1] * elem_bytes, stride[2] * elem_bytes,
                   ds.sB * elem_bytes, ds.sH * elem_bytes, ds.sT * elem_bytes, (int)(rb / 16), -1};
        for (int sft = 0; sft < 9; ++sft) if ((1 << sft) == g.cpr) g.shift = sft;
        if (B == 1) { g.H = 1; g.sB = g.sH; g.dB = g.dH; }
        else if (H == 1) { g.H = 1; }
        else if (stride[0] == H * stride[1] && ds.sB == H * ds.sH) { g.H = 1; g.sB = g.sH; g.dB = g.dH; }  // uniform planes
        constexpr int U = 4;
        const int rpp = kThreads / g.cpr;
        const int64_t tiles = (n_out + (int64_t)rpp * U - 1) / ((int64_t)rpp * U);
        // 128 short CTAs per SM: 6.25 TB/s vs 5.96 at 16 (profiles/r1c_grid_sweep.log); batched: that many over ALL tensors
        int64_t total = ((int64_t)sm_count() * kGatherCtasPerSm + n_tensors - 1) / n_tensors;
        if (total < 1) total = 1;
        int64_t gx = tiles < total ? tiles : total, gy = (total + gx - 1) / gx;
        if (gy > g.nP) gy = g.nP;
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)gx, (unsigned)(gy > 0 ? gy : 1), (unsigned)n_tensors);
        evict_gather_kernel<U><<<grid, kThreads, 0, st>>>((const char*)src, (char*)dst, g, pd, old_table, srcs, dsts);
    } else {
        if (srcs) return fail(KVQ_ERR_UNSUPPORTED, "evict_gather_batched: rows must be 16-byte multiples with 16-byte-aligned strides and pointers "
                                                   "(use kvq_evict_gather per tensor for other layouts)");
        const unsigned grid = gen_grid(B * H * n_out * D);
#define KVQ_GG(E) evict_gather_generic_kernel<E><<<grid, kThreads, 0, st>>>((const E*)src, (E*)dst, B, H, D, stride[0], stride[1], \
                                                                           stride[2], stride[3], ds.sB, ds.sH, ds.sT, pd, old_table)
        if (elem_bytes == 1) KVQ_GG(uint8_t); else if (elem_bytes == 2) KVQ_GG(uint16_t); else KVQ_GG(uint32_t);
#undef KVQ_GG
    }
    return check_launch("evict_gather");
}
} }

extern "C" int kvq_evict_gather(const void* src, int elem_bytes, const int64_t shape[4], const int64_t stride[4],
                                const kvq_evict_plan* plan, const int32_t* old_table, void* dst, const int64_t dst_stride3[3],
                                kvq_stream_t stream) {
    KVQ_REQUIRE(src && dst && shape && stride && plan, "null pointer");
    return run_gather(src, dst, nullptr, nullptr, 1, aligned16(src) && aligned16(dst), elem_bytes, shape, stride, plan, old_table,
                      dst_stride3, (cudaStream_t)stream);
}

extern "C" int kvq_evict_gather_batched(const void* const* srcs, void* const* dsts, int64_t n_tensors, int elem_bytes,
                                        const int64_t shape[4], const int64_t stride[4], const kvq_evict_plan* plan,
                                        const int32_t* old_table, const int64_t dst_stride3[3], kvq_stream_t stream) {
    KVQ_REQUIRE(srcs && dsts && shape && stride && plan, "null pointer");
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535, "n_tensors out of range");
    return run_gather(nullptr, nullptr, srcs, dsts, n_tensors, true, elem_bytes, shape, stride, plan, old_table, dst_stride3,
                      (cudaStream_t)stream);
}

// The older region's tokens of a budget plan whose float32 linspace produced duplicates (plan->needs_table), built ON THE
// DEVICE: idx[i] = trunc(linspace(i)), keep i when idx[i] != idx[i-1] (torch.unique_consecutive, implementations.py:279-282),
// compacted by a block-wide scan.  table: int32 [plan->n_old].
namespace kvq { namespace {
__global__ void __launch_bounds__(1024) budget_table_kernel(PlanDev pl, int32_t* __restrict__ table) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < pl.budget; base += 1024) {
        const int i = base + threadIdx.x;
        int v = 0, keep = 0;
        if (i < pl.budget) {
            v = linspace_index(pl.lin_start, pl.lin_end, pl.lin_step, pl.budget, i);
            keep = i == 0 || v != linspace_index(pl.lin_start, pl.lin_end, pl.lin_step, pl.budget, i - 1);
        }
        int incl = keep;  // inclusive scan of `keep` over the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += n; }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int pos = carry + (warp ? warp_sum[warp - 1] : 0) + incl - keep;
        if (keep && pos < pl.n_old) table[pos] = v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
}
} }

extern "C" int kvq_evict_old_table_device(const kvq_evict_plan* plan, int32_t* old_table, kvq_stream_t stream) {
    KVQ_REQUIRE(plan && old_table, "null pointer");
    KVQ_REQUIRE(plan->strategy == KVQ_EVICT_BUDGET && plan->budget > 0, "old_table exists for budget plans with old_len > budget only");
    const PlanDev d = to_dev(*plan, false);
    budget_table_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d, old_table);
    return check_launch("evict_old_table");
}

// ------------------------------------------------------------------------------------------------ paged store



extern "C" int kvq_paged_gather(const void* pool, int elem_bytes, const int64_t shape[4], const int32_t* block_table,
                                int64_t block_size, void* out, kvq_stream_t stream) {
    KVQ_REQUIRE(pool && out && shape && block_table, "null pointer");
    KVQ_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 1, 2 or 4");
    KVQ_REQUIRE(block_size >= 1 && block_size <= 0x7fffffff, "block_size out of range");
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && T >= 0 && D >= 0 && B * H <= 0x7fffffff, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (B * H * T * D == 0) return KVQ_OK;
    const int64_t rb = D * elem_bytes;
    if (rb % 16 == 0 && rb / 16 <= kThreads && aligned16(pool) && aligned16(out)) {
        PagedGeom g{(int)(B * H), T, (int)block_size, rb, (int)(rb / 16)};
        constexpr int U = 4;
        const int rpp = kThreads / g.cpr;
        dim3 grid((unsigned)((T + (int64_t)rpp * U - 1) / ((int64_t)rpp * U)), (unsigned)(g.nP < 65535 ? g.nP : 65535));
        paged_gather_kernel<U><<<grid, kThreads, 0, st>>>((const char*)pool, (char*)out, g, block_table);
    } else {
        const unsigned grid = gen_grid(B * H * T * D);
#define KVQ_PG(E) paged_gather_generic_kernel<E><<<grid, kThreads, 0, st>>>((const E*)pool, (E*)out, B * H, T, D, block_size, block_table)
        if (elem_bytes == 1) KVQ_PG(uint8_t); else if (elem_bytes == 2) KVQ_PG(uint16_t); else KVQ_PG(uint32_t);
#undef KVQ_PG
    }
    return check_launch("paged_gather");
}

// n tokens of B*H rows each.  Fast path: one 16-byte vector per thread, 4 in flight, 32-bit index arithmetic
// (row -> (plane, token) costs one division per VECTOR, not four 64-bit div/mod per element); the block-table lookup is
// done by the device (the table stays resident on the GPU).
namespace kvq { namespace {
struct AppendGeom {
    int nP, H, N, cpr, block;     // planes, heads, tokens appended, 16-byte chunks per row, tokens per block
    int64_t sB, sH, sT;           // source strides in BYTES
    int64_t plane_bytes, block_bytes, row_bytes;
    int64_t t;                    // logical position of the first appended token
};
template <int UNROLL>
__global__ void __launch_bounds__(kThreads) paged_append_vec_kernel(const char* __restrict__ x, char* __restrict__ pool, AppendGeom g,
                                                                   const int32_t* __restrict__ table) {
    const unsigned n_vec = (unsigned)g.nP * g.N * g.cpr;
    const unsigned i0 = (blockIdx.x * UNROLL) * kThreads + threadIdx.x;
    uint4 v[UNROLL];
    char* dst[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const unsigned i = i0 + u * kThreads;
        dst[u] = nullptr;
        if (i < n_vec) {
            const unsigned r = i / g.cpr, c = i - r * g.cpr;   // r = p * N + j
            const unsigned p = r / g.N, j = r - p * g.N;
            const unsigned b = g.H == 1 ? p : p / g.H, h = p - b * g.H;
            const int64_t lt = g.t + j, lb = lt / g.block, slot = lt - lb * g.block;
            v[u] = ld_stream(x + b * g.sB + h * g.sH + (int64_t)j * g.sT + (int64_t)c * 16);
            dst[u] = pool + (int64_t)__ldg(table + lb) * g.block_bytes + (int64_t)p * g.plane_bytes + slot * g.row_bytes + (int64_t)c * 16;
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
        if (dst[u]) st_stream(dst[u], v[u]);
}

// n tokens starting at logical position t (n == 1: the reference's per-token append; n > 1: a whole prompt in one launch)
template <typename E>
__global__ void __launch_bounds__(kThreads) paged_append_tbl_kernel(const E* __restrict__ x, E* __restrict__ pool, int64_t B,
                                                                   int64_t H, int64_t N, int64_t D, int64_t sB, int64_t sH,
                                                                   int64_t sT, int64_t sD, int64_t block,
                                                                   const int32_t* __restrict__ table, int64_t t) {
    const int64_t n = B * H * N * D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % D, r = i / D, j = r % N, p = r / N, h = p % H, b = p / H;
        const int64_t lt = t + j, lb = lt / block, slot = lt - lb * block, phys = table[lb];
        pool[((phys * B * H + p) * block + slot) * D + d] = x[b * sB + h * sH + j * sT + d * sD];
    }
}
} }

extern "C" int kvq_paged_append(const void* x, int elem_bytes, const int64_t shape[4], const int64_t stride[4], void* pool,
                                const int32_t* block_table, int64_t block_size, int64_t t, kvq_stream_t stream) {
    KVQ_REQUIRE(x && pool && shape && stride && block_table, "null pointer");
    KVQ_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 1, 2 or 4");
    KVQ_REQUIRE(shape[2] >= 1, "paged append needs at least one token (shape[2] >= 1)");
    KVQ_REQUIRE(block_size >= 1 && t >= 0, "bad block_size / t");
    const int64_t B = shape[0], H = shape[1], N = shape[2], D = shape[3];
    if (B * H * D == 0) return KVQ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rb = D * elem_bytes;
    auto a16 = [&](int64_t s_) { return (s_ * elem_bytes) % 16 == 0; };
    if (stride[3] == 1 && rb % 16 == 0 && aligned16(x) && aligned16(pool) && a16(stride[0]) && a16(stride[1]) && a16(stride[2]) &&
        B * H * N * (rb / 16) < 0x7fffffffLL && block_size <= 0x7fffffff) {
        AppendGeom g{(int)(B * H), (int)H, (int)N, (int)(rb / 16), (int)block_size, stride[0] * elem_bytes, stride[1] * elem_bytes,
                     stride[2] * elem_bytes, block_size * rb, B * H * block_size * rb, rb, t};
        if (B == 1) { g.H = 1; g.sB = g.sH; }
        else if (H == 1) { g.H = 1; }
        else if (stride[0] == H * stride[1]) { g.H = 1; g.sB = g.sH; }
        constexpr int U = 4;
        const int64_t n_vec = B * H * N * (rb / 16);
        const unsigned grid = (unsigned)((n_vec + kThreads * U - 1) / (kThreads * U));
        paged_append_vec_kernel<U><<<grid, kThreads, 0, st>>>((const char*)x, (char*)pool, g, block_table);
        return check_launch("paged_append");
    }
    const unsigned grid = gen_grid(B * H * N * D);
#define KVQ_PA(E) paged_append_tbl_kernel<E><<<grid, kThreads, 0, st>>>((const E*)x, (E*)pool, B, H, N, D, stride[0], stride[1], stride[2], \
                                                                        stride[3], block_size, block_table, t)
    if (elem_bytes == 1) KVQ_PA(uint8_t); else if (elem_bytes == 2) KVQ_PA(uint16_t); else KVQ_PA(uint32_t);
#undef KVQ_PA
    return check_launch("paged_append");
}
