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
template <int UNROLL>
__global__ void __launch_bounds__(kThreads) evict_gather_kernel(const char* __restrict__ src, char* __restrict__ dst,
                                                               CopyGeom g, PlanDev pl, const int32_t* __restrict__ table) {
    __shared__ int src_row[kThreads * UNROLL];
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

extern "C" int kvq_evict_gather(const void* src, int elem_bytes, const int64_t shape[4], const int64_t stride[4],
                                const kvq_evict_plan* plan, const int32_t* old_table, void* dst, const int64_t dst_stride3[3],
                                kvq_stream_t stream) {
    KVQ_REQUIRE(src && dst && shape && stride && plan, "null pointer");
    KVQ_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 1, 2 or 4");
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && D >= 0, "negative shape");
    KVQ_REQUIRE(T == plan->T, "plan was made for T=%lld, tensor has T=%lld", (long long)plan->T, (long long)T);
    KVQ_REQUIRE(!plan->needs_table || old_table, "plan needs old_table (float32 linspace produced duplicates)");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_out = plan->n_total;
    if (B * H * n_out * D == 0) return KVQ_OK;
    const Stride3 ds = make_stride3(dst_stride3, H, n_out, D);
    const PlanDev pd = to_dev(*plan, old_table != nullptr);
    const int64_t rb = D * elem_bytes;
    auto a16 = [&](int64_t s) { return (s * elem_bytes) % 16 == 0; };
    const bool fast = stride[3] == 1 && rb % 16 == 0 && rb / 16 <= kThreads && aligned16(src) && aligned16(dst) && a16(stride[0]) &&
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
        static const int ctas_per_sm = [] { const char* e = getenv("KVQ_GATHER_CTAS_PER_SM"); const int n = e ? atoi(e) : 0; return n > 0 ? n : 128; }();
        const int64_t total = (int64_t)sm_count() * ctas_per_sm;  // 128 short CTAs per SM: 6.25 TB/s vs 5.96 at 16 (r1c sweep)
        int64_t gx = tiles < total ? tiles : total, gy = (total + gx - 1) / gx;
        if (gy > g.nP) gy = g.nP;
        dim3 grid((unsigned)gx, (unsigned)(gy > 0 ? gy : 1));
        evict_gather_kernel<U><<<grid, kThreads, 0, st>>>((const char*)src, (char*)dst, g, pd, old_table);
    } else {
        const unsigned grid = gen_grid(B * H * n_out * D);
#define KVQ_GG(E) evict_gather_generic_kernel<E><<<grid, kThreads, 0, st>>>((const E*)src, (E*)dst, B, H, D, stride[0], stride[1], \
                                                                           stride[2], stride[3], ds.sB, ds.sH, ds.sT, pd, old_table)
        if (elem_bytes == 1) KVQ_GG(uint8_t); else if (elem_bytes == 2) KVQ_GG(uint16_t); else KVQ_GG(uint32_t);
#undef KVQ_GG
    }
    return check_launch("evict_gather");
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

// One token is B*H rows of D elements (<= a few tens of KB): a single element-wise kernel; the block
// table lookup for the one logical block is done by the device (table stays resident on the GPU).
namespace kvq { namespace {
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
    const unsigned grid = gen_grid(B * H * N * D);
#define KVQ_PA(E) paged_append_tbl_kernel<E><<<grid, kThreads, 0, st>>>((const E*)x, (E*)pool, B, H, N, D, stride[0], stride[1], stride[2], \
                                                                        stride[3], block_size, block_table, t)
    if (elem_bytes == 1) KVQ_PA(uint8_t); else if (elem_bytes == 2) KVQ_PA(uint16_t); else KVQ_PA(uint32_t);
#undef KVQ_PA
    return check_launch("paged_append");
}
