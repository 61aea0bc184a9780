// pool.cu -- area 3: chunk-summary mean pooling as one fused segmented reduction (sm_100a).
//
// Reference: chunk_summarize_kv, src/cache/implementations.py:295-346 (/root/reference):
//   old = x[:, :, :old_len]; zero-pad to a multiple of chunk (:326-333); view(B,H,n,chunk,D).mean(3)
//   (:338-339, divides by chunk even for the padded last chunk); cat(summaries, recent) (:342-343).
// torch: pad-cat copy of the whole older region + mean + cat = ~3x the traffic.  Here one launch reads
// every old row once (old_len*R), writes the summaries (n_chunks*R) and copies the kept tail
// (2*keep*R); the zero padding is never materialised (missing rows simply are not added).
//
// Mapping: a thread owns one 16-byte column slice of one OUTPUT row and walks the chunk's tokens with
// 8 independent 128-bit loads in flight; fp32 accumulation in token order, one fp32 division by
// chunk, one rounding to the dtype (torch CPU: cast fp32 -> sum -> div -> cast).
#include "kvq_common.cuh"

namespace kvq {
namespace {

constexpr int kThreads = 256;

struct PoolGeom {
    int nP, H;
    int64_t xB, xH, xT;  // input strides, elements
    int64_t T, old_len, n_chunks, keep_eff, Tn;
    int chunk;
    int cpr;             // 16-byte column slices per row
};

template <int DT> struct Acc {
    static constexpr int kN = 16 / ElemT<DT>::kSize;  // elements per 16 bytes
    float a[kN];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < kN; ++i) a[i] = 0.f;
    }
    __device__ __forceinline__ void add(const uint4& v) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
        if (DT == KVQ_F32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = __fadd_rn(a[i], __uint_as_float(w[i]));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float lo, hi;
                widen2<DT>(w[i], lo, hi);
                a[2 * i] = __fadd_rn(a[2 * i], lo);
                a[2 * i + 1] = __fadd_rn(a[2 * i + 1], hi);
            }
        }
    }
    __device__ __forceinline__ uint4 mean(float n) const {
        uint4 o;
        uint32_t* w = reinterpret_cast<uint32_t*>(&o);
        if (DT == KVQ_F32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(__fdiv_rn(a[i], n));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = narrow2<DT>(__fdiv_rn(a[2 * i], n), __fdiv_rn(a[2 * i + 1], n));
        }
        return o;
    }
};

template <int DT>
__global__ void __launch_bounds__(kThreads) chunk_mean_kernel(const char* __restrict__ x, char* __restrict__ out, PoolGeom g,
                                                             const void* const* __restrict__ xs = nullptr,
                                                             void* const* __restrict__ outs = nullptr) {
    constexpr int ES = ElemT<DT>::kSize;
    if (xs) { x = reinterpret_cast<const char*>(xs[blockIdx.z]); out = reinterpret_cast<char*>(outs[blockIdx.z]); }  // batched: z = tensor
    constexpr int kInFlight = 8;
    const int c = threadIdx.x % g.cpr, r0 = threadIdx.x / g.cpr, rpp = kThreads / g.cpr;
    if (r0 >= rpp) return;
    const int64_t row_bytes = (int64_t)g.cpr * 16;
    for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
        const int b = p / g.H, h = p - b * g.H;
        const char* xp = x + (b * g.xB + h * g.xH) * ES + (int64_t)c * 16;
        const int64_t orow = (int64_t)blockIdx.x * rpp + r0;  // output row within the plane
        if (orow >= g.Tn) continue;
        char* o = out + ((int64_t)p * g.Tn + orow) * row_bytes + (int64_t)c * 16;
        if (orow >= g.n_chunks) {  // recent tokens: exact copy
            const int64_t t = g.old_len + (orow - g.n_chunks);
            st_stream(o, ld_stream(xp + t * g.xT * ES));
            continue;
        }
        const int64_t t0 = orow * g.chunk;
        const int64_t t1 = t0 + g.chunk < g.old_len ? t0 + g.chunk : g.old_len;
        Acc<DT> acc;
        acc.clear();
        int64_t t = t0;
        for (; t + kInFlight <= t1; t += kInFlight) {
            uint4 v[kInFlight];
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) v[u] = ld_stream(xp + (t + u) * g.xT * ES);
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) acc.add(v[u]);
        }
        for (; t < t1; ++t) acc.add(ld_stream(xp + t * g.xT * ES));
        st_stream(o, acc.mean((float)g.chunk));
    }
}

// any D / stride / alignment: one thread per output element
template <int DT>
__global__ void __launch_bounds__(kThreads) chunk_mean_generic_kernel(const void* __restrict__ x, void* __restrict__ out,
                                                                     int64_t B, int64_t H, int64_t D, int64_t xB, int64_t xH,
                                                                     int64_t xT, int64_t xD, int64_t old_len, int64_t n_chunks,
                                                                     int64_t Tn, int64_t chunk) {
    const int64_t n = B * H * Tn * D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % D, r = i / D, orow = r % Tn, ph = r / Tn, h = ph % H, b = ph / H;
        const int64_t base = b * xB + h * xH + d * xD;
        if (orow >= n_chunks) {
            store_scalar<DT>(out, i, load_as_float<DT>(x, base + (old_len + orow - n_chunks) * xT));
            continue;
        }
        const int64_t t0 = orow * chunk, t1 = t0 + chunk < old_len ? t0 + chunk : old_len;
        float acc = 0.f;
        for (int64_t t = t0; t < t1; ++t) acc = __fadd_rn(acc, load_as_float<DT>(x, base + t * xT));
        store_scalar<DT>(out, i, __fdiv_rn(acc, (float)chunk));
    }
}

}  // namespace
}  // namespace kvq

using namespace kvq;

namespace kvq { namespace {
int run_chunk_mean(const void* x, void* out, const void* const* xs, void* const* outs, int64_t n_tensors, bool ptrs_aligned, int dtype,
                   const int64_t shape[4], const int64_t stride[4], int64_t chunk, int64_t keep_last, cudaStream_t st) {
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    KVQ_REQUIRE(chunk >= 1 && chunk <= 0x7fffffff, "chunk_size must be >= 1");
    KVQ_REQUIRE(keep_last >= 0, "keep_last must be >= 0");
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && T >= 0 && D >= 0 && B * H <= 0x7fffffff, "bad shape");
    const int64_t keep_eff = keep_last < T ? keep_last : T, old_len = T - keep_eff;
    KVQ_REQUIRE(old_len > 0, "nothing to summarise (T <= keep_last): the caller returns the cache unchanged (implementations.py:316-318)");
    const int64_t n_chunks = (old_len + chunk - 1) / chunk, Tn = n_chunks + keep_eff;
    if (B * H * D == 0 || n_tensors == 0) return KVQ_OK;
    const int es = elem_size(dtype);
    const int64_t rb = D * es;
    auto a16 = [&](int64_t s) { return (s * es) % 16 == 0; };
    const bool fast = stride[3] == 1 && rb % 16 == 0 && rb / 16 <= kThreads && ptrs_aligned && a16(stride[0]) &&
                      a16(stride[1]) && a16(stride[2]);
    if (fast) {
        PoolGeom g{(int)(B * H), (int)H, stride[0], stride[1], stride[2], T, old_len, n_chunks, keep_eff, Tn, (int)chunk, (int)(rb / 16)};
        const int rpp = kThreads / g.cpr;
        dim3 grid((unsigned)((Tn + rpp - 1) / rpp), (unsigned)(g.nP < 65535 ? g.nP : 65535), (unsigned)n_tensors);
        switch (dtype) {
            case KVQ_F32: chunk_mean_kernel<KVQ_F32><<<grid, kThreads, 0, st>>>((const char*)x, (char*)out, g, xs, outs); break;
            case KVQ_F16: chunk_mean_kernel<KVQ_F16><<<grid, kThreads, 0, st>>>((const char*)x, (char*)out, g, xs, outs); break;
            default: chunk_mean_kernel<KVQ_BF16><<<grid, kThreads, 0, st>>>((const char*)x, (char*)out, g, xs, outs); break;
        }
    } else {
        if (xs) return fail(KVQ_ERR_UNSUPPORTED, "chunk_mean_batched: rows must be 16-byte multiples with 16-byte-aligned strides and pointers "
                                                 "(use kvq_chunk_mean per tensor for other layouts)");
        const int64_t n = B * H * Tn * D;
        const int64_t blocks = (n + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * 32;
        const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
#define KVQ_CM(DT) chunk_mean_generic_kernel<DT><<<grid, kThreads, 0, st>>>(x, out, B, H, D, stride[0], stride[1], stride[2], stride[3], \
                                                                            old_len, n_chunks, Tn, chunk)
        switch (dtype) {
            case KVQ_F32: KVQ_CM(KVQ_F32); break;
            case KVQ_F16: KVQ_CM(KVQ_F16); break;
            default: KVQ_CM(KVQ_BF16); break;
        }
#undef KVQ_CM
    }
    return check_launch("chunk_mean");
}
} }

extern "C" int kvq_chunk_mean(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int64_t chunk,
                              int64_t keep_last, void* out, kvq_stream_t stream) {
    KVQ_REQUIRE(x && out && shape && stride, "null pointer");
    return run_chunk_mean(x, out, nullptr, nullptr, 1, aligned16(x) && aligned16(out), dtype, shape, stride, chunk, keep_last,
                          (cudaStream_t)stream);
}

extern "C" int kvq_chunk_mean_batched(const void* const* xs, void* const* outs, int64_t n_tensors, int dtype, const int64_t shape[4],
                                      const int64_t stride[4], int64_t chunk, int64_t keep_last, kvq_stream_t stream) {
    KVQ_REQUIRE(xs && outs && shape && stride, "null pointer");
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535, "n_tensors out of range");
    return run_chunk_mean(nullptr, nullptr, xs, outs, n_tensors, true, dtype, shape, stride, chunk, keep_last, (cudaStream_t)stream);
}
