// quant.cu -- area 1: symmetric INT8 / nibble-packed INT4 quantize + dequantize (sm_100a).
//
// Reference semantics (src/quantization/ops.py in /root/reference):
//   scale = max(max|x| / qmax, eps)            fp32 division                 :28, :49
//   code  = clamp(rint(x / scale), lo, hi)     true division, half-to-even   :29, :50
//   INT4: byte j = ((code[2j]+8) << 4) | (code[2j+1]+8), odd D padded with nibble 8   :53-63
//   dequant = (float(code) * float(scale.to(x.dtype))) -> out dtype          :90, :122-133
//
// Everything here is HBM-bound streaming work: 128-bit coalesced loads/stores, several loads in
// flight per thread, warp REDUX + one atomicMax per CTA for the max.  The per-element division is
// replaced, for 16-bit inputs, by y = RN(1/scale); q0 = RN(x*y); r = fma(-q0,scale,x);
// q1 = fma(r,y,q0), which gives the same code as rint(x/scale) for EVERY fp16/bf16 x and every
// possible scale (proved by exhaustion: oracle kvqo_check_division_free, tests/test_exact_division.py).
// fp32 inputs use the IEEE division (__fdiv_rn).
#include "kvq_common.cuh"

namespace kvq {
namespace {

constexpr int kThreads = 256;
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: x + kMagic rounds x to an integer (half-to-even)
                                       // and leaves it in the low mantissa bits (two's complement)

template <int BITS> struct Q;
template <> struct Q<8> { static constexpr float qmax = 127.f, lo = -127.f, hi = 127.f, magic = kMagic; };
template <> struct Q<4> { static constexpr float qmax = 7.f, lo = -8.f, hi = 7.f, magic = kMagic + 8.f; };

struct Scale {
    float s;  // the fp32 scale divided by (not rounded to the input dtype)
    float y;  // RN(1/s)
};

template <int BITS> __device__ __forceinline__ Scale scale_from_bits(uint32_t absmax_bits, float eps) {
    Scale sc;
    sc.s = fmaxf(__fdiv_rn(__uint_as_float(absmax_bits), Q<BITS>::qmax), eps);
    sc.y = __frcp_rn(sc.s);
    return sc;
}

// float bits of clamp(x/scale) + magic: the low byte (BITS=8) is the int8 code, the low nibble
// (BITS=4) is code+8.
template <int DT, int BITS> __device__ __forceinline__ uint32_t code_word(float x, const Scale& sc) {
    float q;
    if (DT == KVQ_F32) {
        q = __fdiv_rn(x, sc.s);
    } else {
        const float q0 = __fmul_rn(x, sc.y);
        const float r = __fmaf_rn(-q0, sc.s, x);
        q = __fmaf_rn(r, sc.y, q0);
    }
    q = fminf(fmaxf(q, Q<BITS>::lo), Q<BITS>::hi);
    return __float_as_uint(__fadd_rn(q, Q<BITS>::magic));
}

__device__ __forceinline__ uint32_t pack4_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// ---- a 16-element item: raw registers, widening, absmax, quantize+pack ---------------------------------

template <int DT> struct Item16 {
    static constexpr int kVecs = DT == KVQ_F32 ? 4 : 2;  // 16-byte vectors per 16 elements
    uint4 v[kVecs];

    __device__ __forceinline__ void load(const void* p) {
#pragma unroll
        for (int i = 0; i < kVecs; ++i) v[i] = ld_stream(reinterpret_cast<const uint4*>(p) + i);
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < kVecs; ++i) v[i] = make_uint4(0, 0, 0, 0);
    }
    // max |x| as fp32 bits
    __device__ __forceinline__ uint32_t absmax_bits() const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
        if (DT == KVQ_F32) {
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) m = max(m, w[i] & 0x7fffffffu);
            return m;
        }
        // packed 16-bit: clear both sign bits, unsigned 16-bit SIMD max (bit order == value order)
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m = __vmaxu2(m, w[i] & 0x7fff7fffu);
        const uint32_t h = max(m & 0xffffu, m >> 16);
        if (DT == KVQ_BF16) return h << 16;
        return __float_as_uint(__half2float(__ushort_as_half((unsigned short)h)));
    }
    __device__ __forceinline__ void widen(float f[16]) const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
        if (DT == KVQ_F32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(w[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) widen2<DT>(w[i], f[2 * i], f[2 * i + 1]);
        }
    }
    // 16 int8 codes
    __device__ __forceinline__ uint4 quant8(const Scale& sc) const {
        float f[16];
        widen(f);
        uint32_t c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = code_word<DT, 8>(f[i], sc);
        return make_uint4(pack4_bytes(c[0], c[1], c[2], c[3]), pack4_bytes(c[4], c[5], c[6], c[7]),
                          pack4_bytes(c[8], c[9], c[10], c[11]), pack4_bytes(c[12], c[13], c[14], c[15]));
    }
    // 8 bytes of packed nibbles (even element -> high nibble)
    __device__ __forceinline__ uint2 quant4(const Scale& sc) const {
        float f[16];
        widen(f);
        uint32_t b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            b[i] = code_word<DT, 4>(f[2 * i], sc) * 16u + code_word<DT, 4>(f[2 * i + 1], sc);  // low byte valid
        return make_uint2(pack4_bytes(b[0], b[1], b[2], b[3]), pack4_bytes(b[4], b[5], b[6], b[7]));
    }
};

// ---- row-tiled addressing shared by the streaming kernels -------------------------------------------
// A "plane" is one (b, h); a CTA covers UNROLL * rows_per_pass tokens of one plane; a thread owns the
// 16-element item `c` of a row, so no per-item division is needed.

struct RowGeom {
    int64_t nP, H;          // planes = B*H (H only used to split p when planes are not uniform)
    int64_t T;              // rows per plane
    int64_t xB, xH, xT;     // input strides (elements)
    int64_t qB, qH, qT;     // output strides (elements of the output tensor)
    int ipr;                // 16-element items per row (D / 16)
};

__device__ __forceinline__ void plane_offsets(const RowGeom& g, int64_t p, int64_t& xo, int64_t& qo) {
    const int b = (int)p / (int)g.H, h = (int)p - b * (int)g.H;
    xo = b * g.xB + h * g.xH;
    qo = b * g.qB + h * g.qH;
}

// ---- phase 1: per-tensor absmax ---------------------------------------------------------------------

template <int DT, int UNROLL>
__global__ void __launch_bounds__(kThreads) absmax_rows_kernel(const void* __restrict__ x, RowGeom g,
                                                               uint32_t* __restrict__ absmax_bits) {
    __shared__ uint32_t red[kThreads / 32];
    const int c = threadIdx.x % g.ipr, r0 = threadIdx.x / g.ipr, rpp = kThreads / g.ipr;
    uint32_t m = 0;
    if (r0 < rpp) {
        for (int64_t p = blockIdx.y; p < g.nP; p += gridDim.y) {
            int64_t xo, qo;
            plane_offsets(g, p, xo, qo);
            const char* base = reinterpret_cast<const char*>(x) + (xo + (int64_t)c * 16) * ElemT<DT>::kSize;
            const int64_t t0 = (int64_t)blockIdx.x * rpp * UNROLL + r0;
            Item16<DT> it[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t t = t0 + (int64_t)u * rpp;
                if (t < g.T) it[u].load(base + t * g.xT * ElemT<DT>::kSize); else it[u].zero();
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) m = max(m, it[u].absmax_bits());
        }
    }
    m = block_max_u32<kThreads>(m, red);
    if (threadIdx.x == 0 && m != 0) atomicMax(absmax_bits, m);
}

// ---- phase 2: quantize rows given absmax (per tensor or per token) ---------------------------------------

template <int DT, int BITS, int UNROLL>
__global__ void __launch_bounds__(kThreads) quant_rows_kernel(const void* __restrict__ x, RowGeom g,
                                                              const uint32_t* __restrict__ absmax_bits,
                                                              int per_token, float eps, void* __restrict__ q,
                                                              void* __restrict__ scale_out,
                                                              float* __restrict__ scale_f32_out) {
    const int c = threadIdx.x % g.ipr, r0 = threadIdx.x / g.ipr, rpp = kThreads / g.ipr;
    constexpr int kOutBytes = BITS == 8 ? 16 : 8;  // bytes of codes per item
    if (r0 >= rpp) return;
    Scale sc = scale_from_bits<BITS>(per_token ? 0u : absmax_bits[0], eps);
    if (!per_token && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        if (scale_out) store_scalar<DT>(scale_out, 0, sc.s);
        if (scale_f32_out) scale_f32_out[0] = round_through<DT>(sc.s);
    }
    for (int64_t p = blockIdx.y; p < g.nP; p += gridDim.y) {
        int64_t xo, qo;
        plane_offsets(g, p, xo, qo);
        const char* xin = reinterpret_cast<const char*>(x) + (xo + (int64_t)c * 16) * ElemT<DT>::kSize;
        char* qout = reinterpret_cast<char*>(q) + qo + (int64_t)c * kOutBytes;
        const int64_t t0 = (int64_t)blockIdx.x * rpp * UNROLL + r0;
        Item16<DT> it[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t < g.T) it[u].load(xin + t * g.xT * ElemT<DT>::kSize);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t >= g.T) continue;
            if (per_token) {
                sc = scale_from_bits<BITS>(absmax_bits[t], eps);
                if (p == 0 && c == 0) {
                    if (scale_out) store_scalar<DT>(scale_out, t, sc.s);
                    if (scale_f32_out) scale_f32_out[t] = round_through<DT>(sc.s);
                }
            }
            if (BITS == 8) st_stream(qout + t * g.qT, it[u].quant8(sc));
            else st_stream8(qout + t * g.qT, it[u].quant4(sc));
        }
    }
}

// ---- single pass per token: one CTA owns token t across all planes -----------------------------------
// NI > 0: the token (nP * ipr items) is held in registers, NI items per thread: one HBM read.
// NI == 0: any size; two loops over the token, the second one hits L2.
// MODE 0: quantize (+ scales, + absmax out); MODE 1: absmax only (phase 1 of the head-sharded path).

template <int DT, int BITS, int NI, int MODE>
__global__ void __launch_bounds__(kThreads) token_fused_kernel(const void* __restrict__ x, RowGeom g, float eps,
                                                              void* __restrict__ q, void* __restrict__ scale_out,
                                                              float* __restrict__ scale_f32_out,
                                                              uint32_t* __restrict__ absmax_out, int accumulate) {
    __shared__ uint32_t red[kThreads / 32];
    constexpr int kOutBytes = BITS == 8 ? 16 : 8;
    constexpr int kNI = NI > 0 ? NI : 1;
    const int64_t t = blockIdx.x;
    const int items = (int)g.nP * g.ipr;
    const char* xt = reinterpret_cast<const char*>(x) + t * g.xT * ElemT<DT>::kSize;
    char* qt = reinterpret_cast<char*>(q) + t * g.qT;

    Item16<DT> it[kNI];
    int64_t qoff[kNI];
    uint32_t m = 0;
    if (NI > 0) {
#pragma unroll
        for (int k = 0; k < kNI; ++k) {
            const int j = threadIdx.x + k * kThreads;
            qoff[k] = -1;
            if (j < items) {
                const int p = j / g.ipr, c = j - p * g.ipr;
                int64_t xo, qo;
                plane_offsets(g, p, xo, qo);
                it[k].load(xt + (xo + (int64_t)c * 16) * ElemT<DT>::kSize);
                qoff[k] = qo + (int64_t)c * kOutBytes;
            } else {
                it[k].zero();
            }
        }
#pragma unroll
        for (int k = 0; k < kNI; ++k) m = max(m, it[k].absmax_bits());
    } else {
        for (int j = threadIdx.x; j < items; j += kThreads) {
            const int p = j / g.ipr, c = j - p * g.ipr;
            int64_t xo, qo;
            plane_offsets(g, p, xo, qo);
            it[0].load(xt + (xo + (int64_t)c * 16) * ElemT<DT>::kSize);
            m = max(m, it[0].absmax_bits());
        }
    }
    m = block_max_u32<kThreads>(m, red);
    if (threadIdx.x == 0 && absmax_out) {
        if (accumulate) atomicMax(absmax_out + t, m); else absmax_out[t] = m;
    }
    if (MODE == 1) return;

    const Scale sc = scale_from_bits<BITS>(m, eps);
    if (threadIdx.x == 0) {
        if (scale_out) store_scalar<DT>(scale_out, t, sc.s);
        if (scale_f32_out) scale_f32_out[t] = round_through<DT>(sc.s);
    }
    if (NI > 0) {
#pragma unroll
        for (int k = 0; k < kNI; ++k) {
            if (qoff[k] < 0) continue;
            if (BITS == 8) st_stream(qt + qoff[k], it[k].quant8(sc));
            else st_stream8(qt + qoff[k], it[k].quant4(sc));
        }
    } else {
        for (int j = threadIdx.x; j < items; j += kThreads) {
            const int p = j / g.ipr, c = j - p * g.ipr;
            int64_t xo, qo;
            plane_offsets(g, p, xo, qo);
            it[0].load(xt + (xo + (int64_t)c * 16) * ElemT<DT>::kSize);
            if (BITS == 8) st_stream(qt + qo + (int64_t)c * kOutBytes, it[0].quant8(sc));
            else st_stream8(qt + qo + (int64_t)c * kOutBytes, it[0].quant4(sc));
        }
    }
}

// ---- generic (any stride, any D incl. odd, unaligned) fallbacks: one thread per element / byte ----------

struct GenGeom {
    int64_t B, H, T, D;
    int64_t xB, xH, xT, xD;
    int64_t qB, qH, qT;  // output strides (elements of the output tensor), last dim contiguous
};

template <int DT>
__global__ void __launch_bounds__(kThreads) absmax_generic_kernel(const void* __restrict__ x, GenGeom g, int per_token,
                                                                  uint32_t* __restrict__ absmax_bits) {
    __shared__ uint32_t red[kThreads / 32];
    const int64_t n = g.B * g.H * g.T * g.D;
    uint32_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % g.D, r = i / g.D, t = r % g.T, ph = r / g.T, h = ph % g.H, b = ph / g.H;
        const uint32_t a = __float_as_uint(fabsf(load_as_float<DT>(x, b * g.xB + h * g.xH + t * g.xT + d * g.xD)));
        if (per_token) { if (a > absmax_bits[t]) atomicMax(absmax_bits + t, a); }
        else m = max(m, a);
    }
    if (!per_token) {
        m = block_max_u32<kThreads>(m, red);
        if (threadIdx.x == 0 && m != 0) atomicMax(absmax_bits, m);
    }
}

template <int DT, int BITS>
__global__ void __launch_bounds__(kThreads) quant_generic_kernel(const void* __restrict__ x, GenGeom g,
                                                                 const uint32_t* __restrict__ absmax_bits, int per_token,
                                                                 float eps, void* __restrict__ q, void* __restrict__ scale_out,
                                                                 float* __restrict__ scale_f32_out) {
    const int64_t W = BITS == 8 ? g.D : (g.D + 1) / 2;  // output bytes per row
    const int64_t n = g.B * g.H * g.T * W;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t j = i % W, r = i / W, t = r % g.T, ph = r / g.T, h = ph % g.H, b = ph / g.H;
        const Scale sc = scale_from_bits<BITS>(absmax_bits[per_token ? t : 0], eps);
        if (j == 0 && ph == 0 && (per_token || t == 0)) {
            if (scale_out) store_scalar<DT>(scale_out, per_token ? t : 0, sc.s);
            if (scale_f32_out) scale_f32_out[per_token ? t : 0] = round_through<DT>(sc.s);
        }
        const int64_t xrow = b * g.xB + h * g.xH + t * g.xT;
        uint8_t* out = reinterpret_cast<uint8_t*>(q) + b * g.qB + h * g.qH + t * g.qT + j;
        if (BITS == 8) {
            *out = (uint8_t)(code_word<DT, 8>(load_as_float<DT>(x, xrow + j * g.xD), sc) & 0xffu);
        } else {
            const uint32_t hi = code_word<DT, 4>(load_as_float<DT>(x, xrow + 2 * j * g.xD), sc) & 0xfu;
            const uint32_t lo = (2 * j + 1 < g.D) ? (code_word<DT, 4>(load_as_float<DT>(x, xrow + (2 * j + 1) * g.xD), sc) & 0xfu) : 8u;
            *out = (uint8_t)((hi << 4) | lo);
        }
    }
}

// ---- dequantize ----------------------------------------------------------------------------------------

// 16 codes -> 16 floats
template <int BITS> struct Codes16;
template <> struct Codes16<8> {
    uint4 v;
    __device__ __forceinline__ void load(const void* p) { v = ld_stream(p); }
    __device__ __forceinline__ void to_float(float f[16], float s) const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) f[4 * i + k] = __fmul_rn((float)(int)(int8_t)((w[i] >> (8 * k)) & 0xffu), s);
        }
    }
};
template <> struct Codes16<4> {
    uint2 v;
    __device__ __forceinline__ void load(const void* p) { v = ld_stream8(p); }
    __device__ __forceinline__ void to_float(float f[16], float s) const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t byte = (w[i] >> (8 * k)) & 0xffu;
                f[8 * i + 2 * k] = __fmul_rn((float)((int)(byte >> 4) - 8), s);
                f[8 * i + 2 * k + 1] = __fmul_rn((float)((int)(byte & 0xfu) - 8), s);
            }
        }
    }
};

template <int ODT> __device__ __forceinline__ void store16(void* p, const float f[16]) {
    if (ODT == KVQ_F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st_stream(reinterpret_cast<uint4*>(p) + i, make_uint4(__float_as_uint(f[4 * i]), __float_as_uint(f[4 * i + 1]),
                                                                  __float_as_uint(f[4 * i + 2]), __float_as_uint(f[4 * i + 3])));
    } else {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            st_stream(reinterpret_cast<uint4*>(p) + i,
                      make_uint4(narrow2<ODT>(f[8 * i], f[8 * i + 1]), narrow2<ODT>(f[8 * i + 2], f[8 * i + 3]),
                                 narrow2<ODT>(f[8 * i + 4], f[8 * i + 5]), narrow2<ODT>(f[8 * i + 6], f[8 * i + 7])));
    }
}

// g.x* describe the CODES (strides in bytes, rows of ipr items), g.q* the OUTPUT (elements, contiguous rows)
template <int ODT, int BITS, int UNROLL>
__global__ void __launch_bounds__(kThreads) dequant_rows_kernel(const void* __restrict__ codes, RowGeom g,
                                                                const void* __restrict__ scale, int scale_dtype,
                                                                int per_token, void* __restrict__ out) {
    const int c = threadIdx.x % g.ipr, r0 = threadIdx.x / g.ipr, rpp = kThreads / g.ipr;
    constexpr int kInBytes = BITS == 8 ? 16 : 8;
    if (r0 >= rpp) return;
    float s = per_token ? 0.f : load_scale(scale, scale_dtype, 0);
    for (int64_t p = blockIdx.y; p < g.nP; p += gridDim.y) {
        int64_t co, oo;
        plane_offsets(g, p, co, oo);
        const char* cin = reinterpret_cast<const char*>(codes) + co + (int64_t)c * kInBytes;
        char* o = reinterpret_cast<char*>(out) + (oo + (int64_t)c * 16) * ElemT<ODT>::kSize;
        const int64_t t0 = (int64_t)blockIdx.x * rpp * UNROLL + r0;
        Codes16<BITS> it[UNROLL];
        float sv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t < g.T) {
                it[u].load(cin + t * g.xT);
                sv[u] = per_token ? load_scale(scale, scale_dtype, t) : s;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t t = t0 + (int64_t)u * rpp;
            if (t >= g.T) continue;
            float f[16];
            it[u].to_float(f, sv[u]);
            store16<ODT>(o + t * g.qT * ElemT<ODT>::kSize, f);
        }
    }
}

// generic: one thread per output element; codes strides in bytes; out contiguous [B,H,T,D]
template <int ODT, int BITS>
__global__ void __launch_bounds__(kThreads) dequant_generic_kernel(const void* __restrict__ codes, GenGeom g,
                                                                   const void* __restrict__ scale, int scale_dtype,
                                                                   int per_token, void* __restrict__ out) {
    const int64_t n = g.B * g.H * g.T * g.D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t d = i % g.D, r = i / g.D, t = r % g.T, ph = r / g.T, h = ph % g.H, b = ph / g.H;
        const float s = load_scale(scale, scale_dtype, per_token ? t : 0);
        const uint8_t* row = reinterpret_cast<const uint8_t*>(codes) + b * g.xB + h * g.xH + t * g.xT;
        int code;
        if (BITS == 8) code = (int)(int8_t)row[d];
        else { const uint8_t by = row[d >> 1]; code = (int)((d & 1) ? (by & 0xf) : (by >> 4)) - 8; }
        store_scalar<ODT>(out, i, __fmul_rn((float)code, s));
    }
}

// ---- host helpers ------------------------------------------------------------------------------------------

bool fast_input_ok(const void* x, int dtype, const View4& v) {
    const int es = elem_size(dtype);
    auto ok = [&](int64_t s) { return (s * es) % 16 == 0; };
    return v.sD == 1 && v.D % 16 == 0 && v.D / 16 <= kThreads && aligned16(x) && ok(v.sB) && ok(v.sH) && ok(v.sT) &&
           v.B * v.H * v.T * v.D > 0;
}
bool fast_codes_ok(const void* q, const Stride3& s, int bits, int64_t D) {
    const int a = bits == 8 ? 16 : 8;
    return D % 16 == 0 && (reinterpret_cast<uintptr_t>(q) % a) == 0 && s.sB % a == 0 && s.sH % a == 0 && s.sT % a == 0;
}

// Collapse a view into planes x rows for a kernel whose work is row-independent.
// If `merge_rows` (per-tensor granularity) and the layouts allow, (B,H,T) become one run of rows.
RowGeom row_geom(const View4& v, const Stride3& q, int64_t q_row, bool merge_rows) {
    RowGeom g{};
    g.ipr = (int)(v.D / 16);
    g.H = v.H; g.nP = v.B * v.H; g.T = v.T;
    g.xB = v.sB; g.xH = v.sH; g.xT = v.sT;
    g.qB = q.sB; g.qH = q.sH; g.qT = q.sT;
    if (merge_rows) {
        const bool x_hm = v.sH == v.T * v.sT, q_hm = q.sH == v.T * q.sT;
        const bool x_bm = v.sB == v.H * v.sH, q_bm = q.sB == v.H * q.sH;
        if (x_hm && q_hm && x_bm && q_bm) { g.nP = 1; g.H = 1; g.T = v.B * v.H * v.T; }
        else if (x_hm && q_hm) { g.nP = v.B; g.H = 1; g.T = v.H * v.T; }
    }
    (void)q_row;
    return g;
}

dim3 row_grid(const RowGeom& g, int unroll) {
    const int rpp = kThreads / g.ipr;
    const int64_t tiles = (g.T + (int64_t)rpp * unroll - 1) / ((int64_t)rpp * unroll);
    return dim3((unsigned)tiles, (unsigned)(g.nP < 65535 ? g.nP : 65535), 1);
}

GenGeom gen_geom(const View4& v, const Stride3& q) {
    return GenGeom{v.B, v.H, v.T, v.D, v.sB, v.sH, v.sT, v.sD, q.sB, q.sH, q.sT};
}
unsigned gen_grid(int64_t n) {
    const int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 32;
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

#define KVQ_DISPATCH_DT(dtype, ...)                                   \
    switch (dtype) {                                                  \
        case KVQ_F32: { constexpr int DT = KVQ_F32; __VA_ARGS__; break; }   \
        case KVQ_F16: { constexpr int DT = KVQ_F16; __VA_ARGS__; break; }   \
        case KVQ_BF16: { constexpr int DT = KVQ_BF16; __VA_ARGS__; break; } \
        default: return fail(KVQ_ERR_INVALID, "unknown dtype %d", dtype);   \
    }

template <int BITS>
int quant_phase2(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
                 const uint32_t* absmax_bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                 float* scale_f32_out, cudaStream_t st) {
    KVQ_REQUIRE(x && q && absmax_bits && shape && stride, "null pointer");
    KVQ_REQUIRE(granularity == KVQ_PER_TENSOR || granularity == KVQ_PER_TOKEN, "bad granularity %d", granularity);
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int64_t W = BITS == 8 ? v.D : (v.D + 1) / 2;
    const Stride3 qs = make_stride3(q_stride3, v.H, v.T, W);
    const int64_t n = v.B * v.H * v.T * v.D;
    const int per_token = granularity == KVQ_PER_TOKEN;
    if (n == 0) {
        // empty tensor: absmax is 0 -> scale = eps; the reference would raise on max() of an empty tensor,
        // the Python wrapper keeps that behaviour; here only the scale is written.
        return KVQ_OK;
    }
    if (fast_input_ok(x, dtype, v) && fast_codes_ok(q, qs, BITS, v.D)) {
        const RowGeom g = row_geom(v, qs, W, !per_token);
        constexpr int U = 2;
        KVQ_DISPATCH_DT(dtype, quant_rows_kernel<DT, BITS, U><<<row_grid(g, U), kThreads, 0, st>>>(
                                   x, g, absmax_bits, per_token, eps, q, scale_out, scale_f32_out));
    } else {
        const GenGeom g = gen_geom(v, qs);
        KVQ_DISPATCH_DT(dtype, quant_generic_kernel<DT, BITS><<<gen_grid(v.B * v.H * v.T * W), kThreads, 0, st>>>(
                                   x, g, absmax_bits, per_token, eps, q, scale_out, scale_f32_out));
    }
    return check_launch("quant");
}

template <int BITS>
int dequant(const void* codes, const int64_t shape[4], const int64_t q_stride3[3], const void* scale, int scale_dtype,
            int granularity, void* out, int out_dtype, cudaStream_t st) {
    KVQ_REQUIRE(codes && scale && out && shape, "null pointer");
    KVQ_REQUIRE(granularity == KVQ_PER_TENSOR || granularity == KVQ_PER_TOKEN, "bad granularity %d", granularity);
    KVQ_REQUIRE(scale_dtype >= 0 && scale_dtype <= 2, "bad scale dtype %d", scale_dtype);
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && T >= 0 && D >= 0, "negative shape");
    const int64_t W = BITS == 8 ? D : (D + 1) / 2;
    const Stride3 cs = make_stride3(q_stride3, H, T, W);
    const int per_token = granularity == KVQ_PER_TOKEN;
    if (B * H * T * D == 0) return KVQ_OK;
    const int64_t contig[4] = {H * T * D, T * D, D, 1};
    const View4 ov = make_view(shape, contig);
    if (fast_codes_ok(codes, cs, BITS, D) && D / 16 <= kThreads && aligned16(out)) {
        // geometry: "x" side = codes (bytes), "q" side = output (elements)
        View4 cv = ov;
        cv.sB = cs.sB; cv.sH = cs.sH; cv.sT = cs.sT;
        RowGeom g = row_geom(cv, Stride3{contig[0], contig[1], contig[2]}, D, !per_token);
        constexpr int U = 2;
        KVQ_DISPATCH_DT(out_dtype, dequant_rows_kernel<DT, BITS, U><<<row_grid(g, U), kThreads, 0, st>>>(
                                       codes, g, scale, scale_dtype, per_token, out));
    } else {
        GenGeom g{B, H, T, D, cs.sB, cs.sH, cs.sT, 1, contig[0], contig[1], contig[2]};
        KVQ_DISPATCH_DT(out_dtype, dequant_generic_kernel<DT, BITS><<<gen_grid(B * H * T * D), kThreads, 0, st>>>(
                                       codes, g, scale, scale_dtype, per_token, out));
    }
    return check_launch("dequant");
}

template <int DT, int BITS, int MODE>
void launch_token_fused(int ni, unsigned T, cudaStream_t st, const void* x, const RowGeom& g, float eps, void* q,
                        void* scale_out, float* scale_f32_out, uint32_t* absmax_out, int accumulate) {
#define KVQ_TF(NI) token_fused_kernel<DT, BITS, NI, MODE><<<T, kThreads, 0, st>>>(x, g, eps, q, scale_out, scale_f32_out, absmax_out, accumulate)
    if (MODE == 1) { KVQ_TF(0); return; }
    switch (ni) {
        case 1: KVQ_TF(1); break;
        case 2: KVQ_TF(2); break;
        case 3: case 4: KVQ_TF(4); break;
        default: KVQ_TF(0); break;
    }
#undef KVQ_TF
}

}  // namespace
}  // namespace kvq

using namespace kvq;

extern "C" int kvq_absmax(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
                          uint32_t* absmax_bits, int accumulate, kvq_stream_t stream) {
    KVQ_REQUIRE(x && absmax_bits && shape && stride, "null pointer");
    KVQ_REQUIRE(granularity == KVQ_PER_TENSOR || granularity == KVQ_PER_TOKEN, "bad granularity %d", granularity);
    cudaStream_t st = (cudaStream_t)stream;
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int per_token = granularity == KVQ_PER_TOKEN;
    const int64_t words = per_token ? v.T : 1;
    const int64_t n = v.B * v.H * v.T * v.D;
    const bool fast = fast_input_ok(x, dtype, v);
    // the per-token fused kernel overwrites (or atomically accumulates); everything else max-accumulates
    if (!accumulate && !(per_token && fast) && words > 0) KVQ_CUDA(cudaMemsetAsync(absmax_bits, 0, words * 4, st));
    if (n == 0) return KVQ_OK;
    if (fast && per_token && v.B * v.H * (v.D / 16) < (1 << 30)) {
        RowGeom g = row_geom(v, Stride3{0, 0, 0}, 0, false);
        KVQ_REQUIRE(v.T <= 0x7fffffff, "T too large");
        KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 8, 1>(0, (unsigned)v.T, st, x, g, 0.f, nullptr, nullptr, nullptr,
                                                            absmax_bits, accumulate));
    } else if (fast) {
        RowGeom g = row_geom(v, Stride3{v.H * v.T * v.D, v.T * v.D, v.D}, v.D, true);
        if (!(v.sH == v.T * v.sT && v.sB == v.H * v.sH)) g = row_geom(v, Stride3{0, 0, 0}, 0, false);
        constexpr int U = 4;
        KVQ_DISPATCH_DT(dtype, absmax_rows_kernel<DT, U><<<row_grid(g, U), kThreads, 0, st>>>(x, g, absmax_bits));
    } else {
        const GenGeom g = gen_geom(v, Stride3{0, 0, 0});
        KVQ_DISPATCH_DT(dtype, absmax_generic_kernel<DT><<<gen_grid(n), kThreads, 0, st>>>(x, g, per_token, absmax_bits));
    }
    return check_launch("absmax");
}

extern "C" int kvq_quant_int8(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
                              const uint32_t* absmax_bits, float eps, int8_t* q, const int64_t q_stride3[3],
                              void* scale_out, float* scale_f32_out, kvq_stream_t stream) {
    return quant_phase2<8>(x, dtype, shape, stride, granularity, absmax_bits, eps, q, q_stride3, scale_out, scale_f32_out,
                           (cudaStream_t)stream);
}

extern "C" int kvq_quant_int4_pack(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                                   int granularity, const uint32_t* absmax_bits, float eps, uint8_t* packed,
                                   const int64_t q_stride3[3], void* scale_out, float* scale_f32_out,
                                   kvq_stream_t stream) {
    return quant_phase2<4>(x, dtype, shape, stride, granularity, absmax_bits, eps, packed, q_stride3, scale_out,
                           scale_f32_out, (cudaStream_t)stream);
}

extern "C" int kvq_quant_per_token_fused(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                                         int bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                                         float* scale_f32_out, uint32_t* absmax_bits_out, kvq_stream_t stream) {
    KVQ_REQUIRE(x && q && shape && stride, "null pointer");
    KVQ_REQUIRE(bits == 8 || bits == 4, "bits must be 8 or 4, got %d", bits);
    cudaStream_t st = (cudaStream_t)stream;
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    KVQ_REQUIRE(v.T <= 0x7fffffff, "T too large");
    const int64_t W = bits == 8 ? v.D : (v.D + 1) / 2;
    const Stride3 qs = make_stride3(q_stride3, v.H, v.T, W);
    if (v.B * v.H * v.T * v.D == 0) return KVQ_OK;
    const int64_t items = v.B * v.H * (v.D / 16);
    if (fast_input_ok(x, dtype, v) && fast_codes_ok(q, qs, bits, v.D) && items < (1 << 30)) {
        const RowGeom g = row_geom(v, qs, W, false);
        const int ni = (int)((items + kThreads - 1) / kThreads);
        if (bits == 8) {
            KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 8, 0>(ni, (unsigned)v.T, st, x, g, eps, q, scale_out,
                                                                scale_f32_out, absmax_bits_out, 0));
        } else {
            KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 4, 0>(ni, (unsigned)v.T, st, x, g, eps, q, scale_out,
                                                                scale_f32_out, absmax_bits_out, 0));
        }
        return check_launch("quant_per_token_fused");
    }
    // generic shapes: two generic phases through a caller-provided absmax buffer
    KVQ_REQUIRE(absmax_bits_out != nullptr,
                "quant_per_token_fused: this shape/stride needs the generic path, which requires absmax_bits_out");
    int rc = kvq_absmax(x, dtype, shape, stride, KVQ_PER_TOKEN, absmax_bits_out, 0, stream);
    if (rc != KVQ_OK) return rc;
    if (bits == 8)
        return quant_phase2<8>(x, dtype, shape, stride, KVQ_PER_TOKEN, absmax_bits_out, eps, q, q_stride3, scale_out,
                               scale_f32_out, st);
    return quant_phase2<4>(x, dtype, shape, stride, KVQ_PER_TOKEN, absmax_bits_out, eps, q, q_stride3, scale_out,
                           scale_f32_out, st);
}

extern "C" int kvq_dequant_int8(const int8_t* q, const int64_t shape[4], const int64_t q_stride3[3], const void* scale,
                                int scale_dtype, int granularity, void* out, int out_dtype, kvq_stream_t stream) {
    return dequant<8>(q, shape, q_stride3, scale, scale_dtype, granularity, out, out_dtype, (cudaStream_t)stream);
}

extern "C" int kvq_dequant_int4_unpack(const uint8_t* packed, const int64_t shape[4], const int64_t q_stride3[3],
                                       const void* scale, int scale_dtype, int granularity, void* out, int out_dtype,
                                       kvq_stream_t stream) {
    return dequant<4>(packed, shape, q_stride3, scale, scale_dtype, granularity, out, out_dtype, (cudaStream_t)stream);
}
