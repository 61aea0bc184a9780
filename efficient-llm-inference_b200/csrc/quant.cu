// quant.cu -- area 1: symmetric INT8 / nibble-packed INT4 quantize + dequantize (sm_100a).
//
// Reference semantics (src/quantization/ops.py in /root/reference):
//   scale = max(max|x| / qmax, eps)            fp32 division                 :28, :49
//   code  = clamp(rint(x / scale), lo, hi)     true division, half-to-even   :29, :50
//   INT4: byte j = ((code[2j]+8) << 4) | (code[2j+1]+8), odd D padded with nibble 8   :53-63
//   dequant = (float(code) * float(scale.to(x.dtype))) -> out dtype          :90, :122-133
//
// Everything here is HBM-bound streaming work.  The unit of work is one 16-byte vector per lane, so
// every warp-level load and store covers one contiguous, fully used run of sectors; a thread keeps
// several vectors in flight and per-thread setup is hoisted out of the inner loop (the first version
// spent half its issue slots on address arithmetic).  Launch shape: many short CTAs, not a few persistent
// ones -- measured faster on B200 by 5-19 % (profiles/r1c_grid_sweep.log).  The max is reduced with REDUX
// per warp and at most one atomicMax per CTA.
//
// The per-element division is replaced, for 16-bit inputs, by y = RN(1/scale); q0 = RN(x*y);
// r = fma(-q0,scale,x); q1 = fma(r,y,q0), which gives the same code as rint(x/scale) for EVERY
// fp16/bf16 x and every possible scale (proved by exhaustion: oracle kvqo_check_division_free,
// tests/test_exact_division.py).  fp32 inputs use the IEEE division (__fdiv_rn).
#include <cstdio>
#include <cstdlib>

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
    const float a = __fdiv_rn(__uint_as_float(absmax_bits), Q<BITS>::qmax);
    sc.s = a != a ? a : fmaxf(a, eps);  // clamp(min=eps) propagates NaN (torch semantics); fmaxf alone would return eps
    sc.y = __frcp_rn(sc.s);
    return sc;
}

// float bits of clamp(x/scale) + magic: the low byte (BITS=8) is the int8 code, the low nibble
// (BITS=4) is code+8.
// CLAMP = false: the caller guarantees max|x| <= the absmax the scale came from (it is the phase-1 result of this
// very data, possibly max-combined with other shards), so |x/scale| <= qmax * (1 + 2^-23) and the clamp is dead code.
// TRUE_DIV: IEEE division also for 16-bit inputs.  The division-free quotient is proved exact (by exhaustion) for scales that
// are absmax / qmax of a 16-bit-representable absmax or the default eps; callers that pass a calibrated fp32 absmax or a
// custom eps get the division.
template <int DT, int BITS, bool CLAMP = true, bool TRUE_DIV = false> __device__ __forceinline__ uint32_t code_word(float x, const Scale& sc) {
    float q;
    if (DT == KVQ_F32 || TRUE_DIV) {
        q = __fdiv_rn(x, sc.s);
    } else {
        const float q0 = __fmul_rn(x, sc.y);
        const float r = __fmaf_rn(-q0, sc.s, x);
        q = __fmaf_rn(r, sc.y, q0);
    }
    if (CLAMP) q = fminf(fmaxf(q, Q<BITS>::lo), Q<BITS>::hi);
    return __float_as_uint(__fadd_rn(q, Q<BITS>::magic));
}

// Two elements at once on the packed-fp32 pipe (16-bit inputs, division-free quotient): the same four IEEE operations per
// element as code_word -- q0 = x * y; r = fma(q0, -s, x) (== fma(-q0, s, x)); q = fma(r, y, q0); q + magic -- issued as FMUL2 /
// FFMA2 / FFMA2 / FADD2, i.e. HALF the issue slots of an issue-bound kernel; results are bit-identical.
template <int BITS, bool CLAMP> __device__ __forceinline__ void code_word2(float x0, float x1, const Scale& sc, uint32_t& c0, uint32_t& c1) {
    const f32x2 X = pack2(x0, x1), Y = pack2(sc.y, sc.y), NS = pack2(-sc.s, -sc.s);
    const f32x2 q0 = fmul2(X, Y);
    const f32x2 r = ffma2(q0, NS, X);
    f32x2 q = ffma2(r, Y, q0);
    if (CLAMP) {
        float a, b;
        unpack2(q, a, b);
        q = pack2(fminf(fmaxf(a, Q<BITS>::lo), Q<BITS>::hi), fminf(fmaxf(b, Q<BITS>::lo), Q<BITS>::hi));
    }
    float o0, o1;
    unpack2(fadd2(q, pack2(Q<BITS>::magic, Q<BITS>::magic)), o0, o1);
    c0 = __float_as_uint(o0); c1 = __float_as_uint(o1);
}

__device__ __forceinline__ uint32_t pack4_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// ---- one 16-byte input vector: 8 fp16/bf16 or 4 fp32 elements ------------------------------------------

template <int DT> struct InVec {
    static constexpr int kN = 16 / ElemT<DT>::kSize;

    // max |x| as fp32 bits
    static __device__ __forceinline__ uint32_t absmax_bits(const uint4& v) {
        if (DT == KVQ_F32) {
            return max(max(v.x & 0x7fffffffu, v.y & 0x7fffffffu), max(v.z & 0x7fffffffu, v.w & 0x7fffffffu));
        }
        // packed 16-bit: clear both sign bits, unsigned 16-bit SIMD max (bit order == value order)
        const uint32_t m = __vmaxu2(__vmaxu2(v.x & 0x7fff7fffu, v.y & 0x7fff7fffu), __vmaxu2(v.z & 0x7fff7fffu, v.w & 0x7fff7fffu));
        return m;  // still packed: finish with finish_absmax()
    }
    // packed running max -> fp32 bits (call once per thread, not per vector)
    static __device__ __forceinline__ uint32_t finish_absmax(uint32_t m) {
        if (DT == KVQ_F32) return m;
        const uint32_t h = max(m & 0xffffu, m >> 16);
        if (DT == KVQ_BF16) return h << 16;
        return __float_as_uint(__half2float(__ushort_as_half((unsigned short)h)));
    }
    static __device__ __forceinline__ uint32_t combine(uint32_t a, uint32_t b) {
        return DT == KVQ_F32 ? max(a, b) : __vmaxu2(a, b);
    }
    static __device__ __forceinline__ void widen(const uint4& v, float f[kN]) {
        if (DT == KVQ_F32) {
            f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
        } else {
            widen2<DT>(v.x, f[0], f[1]); widen2<DT>(v.y, f[2], f[3]); widen2<DT>(v.z, f[4], f[5]); widen2<DT>(v.w, f[6], f[7]);
        }
    }
};

// codes of one input vector, stored at `out` (naturally aligned): kN bytes (INT8) or kN/2 bytes (INT4)
template <int DT, int BITS, bool CLAMP = true, bool TRUE_DIV = false> __device__ __forceinline__ void quant_store(const uint4& v, const Scale& sc, void* out) {
    constexpr int kN = InVec<DT>::kN;
    float f[kN];
    InVec<DT>::widen(v, f);
    constexpr bool kPairs = DT != KVQ_F32 && !TRUE_DIV;   // division-free path: two elements per packed-fp32 instruction
    if constexpr (BITS == 8) {
        uint32_t c[kN];
        if constexpr (kPairs) {
#pragma unroll
            for (int i = 0; i < kN; i += 2) code_word2<8, CLAMP>(f[i], f[i + 1], sc, c[i], c[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < kN; ++i) c[i] = code_word<DT, 8, CLAMP, TRUE_DIV>(f[i], sc);
        }
        if constexpr (kN == 8) st_stream8(out, make_uint2(pack4_bytes(c[0], c[1], c[2], c[3]), pack4_bytes(c[4], c[5], c[6], c[7])));
        else st_stream4(out, pack4_bytes(c[0], c[1], c[2], c[3]));
    } else {
        uint32_t b[kN / 2];
#pragma unroll
        for (int i = 0; i < kN / 2; ++i) {
            if constexpr (kPairs) {
                uint32_t hi, lo;
                code_word2<4, CLAMP>(f[2 * i], f[2 * i + 1], sc, hi, lo);
                b[i] = hi * 16u + lo;  // low byte valid
            } else {
                b[i] = code_word<DT, 4, CLAMP, TRUE_DIV>(f[2 * i], sc) * 16u + code_word<DT, 4, CLAMP, TRUE_DIV>(f[2 * i + 1], sc);  // low byte valid
            }
        }
        if constexpr (kN == 8) st_stream4(out, pack4_bytes(b[0], b[1], b[2], b[3]));
        else *reinterpret_cast<unsigned short*>(out) = (unsigned short)(__byte_perm(b[0], b[1], 0x0040) & 0xffffu);
    }
}
template <int DT, int BITS> struct OutBytes { static constexpr int value = InVec<DT>::kN * BITS / 8; };

// ---- flat kernels: contiguous input and output, per-tensor scale ---------------------------------------------

// Flat kernels: NOT persistent.  A chunk is kThreads * UNROLL consecutive vectors (16-32 KB); a CTA owns a few chunks of one
// tensor (consecutive, or strided over the tensor: FlatWork), and the hardware hands CTAs out in index order, so the set of
// resident CTAs sweeps memory front to back.  On B200 this beats a grid-stride loop over 148 x k resident CTAs by up to 19 % on the write-heavy
// dequantize stream (5.5 -> 6.6 TB/s) and 5-8 % on the read-heavy ones (profiles/r1c_grid_sweep.log).
// Batched form (xs != nullptr): the grid covers n_tensors x blocks_per_tensor CTAs, tensor-major, so the sweep runs through
// one tensor after the other exactly like separate launches would, without the launch gaps.  Entry i of absmax_bits /
// scale_out / scale_f32_out belongs to tensor i.
struct FlatWork {
    int64_t chunks;   // chunks per tensor
    int bpt;          // CTAs (blocks) per tensor
    int cpc;          // chunks per CTA
    int cstep;        // distance between a CTA's chunks: 1 = consecutive chunks, bpt = strided over the tensor
};
// CTA -> tensor index and its chunks c0, c0 + cstep, ... (< c1)
__device__ __forceinline__ void flat_split(const FlatWork& w, int& ti, int64_t& c0, int64_t& c1) {
    ti = blockIdx.x / w.bpt;
    const int b = blockIdx.x - ti * w.bpt;
    if (w.cstep == 1) { c0 = (int64_t)b * w.cpc; c1 = c0 + w.cpc < w.chunks ? c0 + w.cpc : w.chunks; }
    else { c0 = b; c1 = w.chunks; }
}

// Head-sharded form of the absmax -> quantize pair WITHOUT a collective launch between them (the one exchange of the path,
// ops.py:27,:48: the scale is over all heads).  The LAST CTA of a tensor in the absmax kernel (arrival counter per tensor)
// stores the tensor's finished word, tagged with the step's sequence number, straight into every rank's peer-mapped buffer
// (one 8-byte NVLink store per peer); the quantize kernel's CTAs of that tensor wait in their prologue until all ranks' words
// for this step are there and max-combine them.  The absmax kernel waits for nobody, so there is no cycle; a peer may lag by
// almost a whole absmax phase before anyone stalls (tensor i's word is published when 1/n of the kernel has run, and is
// needed when the consumer's quantize reaches tensor i).  Two parities: a rank can be at most one step ahead of the slowest
// (its step s + 1 quantize needs everyone's s + 1 words, which are published after their step s quantize).
struct Publish {
    void* const* peers = nullptr;   // device array of `world` peer-mapped buffer pointers (nullptr: single rank, plain kernel)
    int rank = 0, world = 1;
    uint32_t seq = 0;
};
struct Gather {
    void* mine = nullptr;           // this rank's buffer (nullptr: absmax_bits already holds the final words)
    int world = 1;
    uint32_t seq = 0;
};

template <int DT, int UNROLL>
__global__ void __launch_bounds__(kThreads) absmax_flat_kernel(const uint4* __restrict__ x, int64_t n_vec,
                                                               uint32_t* __restrict__ absmax_bits, FlatWork w,
                                                               const void* const* __restrict__ xs = nullptr, Publish pub = Publish()) {
    __shared__ uint32_t red[kThreads / 32];
    int ti; int64_t c0, c1;
    flat_split(w, ti, c0, c1);
    const uint4* xt = xs ? reinterpret_cast<const uint4*>(xs[ti]) : x;
    uint32_t m = 0;
    for (int64_t c = c0; c < c1; c += w.cstep) {
        const int64_t i = c * (kThreads * UNROLL) + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + (int64_t)u * kThreads;
            v[u] = j < n_vec ? ld_stream(xt + j) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) m = InVec<DT>::combine(m, InVec<DT>::absmax_bits(v[u]));
    }
    m = block_max_u32<kThreads>(InVec<DT>::finish_absmax(m), red);
    // thousands of CTAs per tensor: only those that can still raise the max touch the atomic
    if (threadIdx.x == 0 && m > *reinterpret_cast<volatile uint32_t*>(absmax_bits + ti)) atomicMax(absmax_bits + ti, m);
    if (pub.peers) {
        __shared__ int last;
        uint32_t* arrive = xch_arrive(pub.peers[pub.rank]);
        if (threadIdx.x == 0) {
            __threadfence();   // my atomicMax before my arrival
            last = atomicAdd(arrive + ti, 1u) == (uint32_t)w.bpt - 1u;
        }
        __syncthreads();
        if (last && threadIdx.x < pub.world) {
            __threadfence();   // everyone's atomicMax before my read
            const uint32_t word = atomicOr(absmax_bits + ti, 0u);
            st_release_sys_u64(xch_tagged(pub.peers[threadIdx.x], (int)(pub.seq & 1u), pub.rank, ti), ((unsigned long long)pub.seq << 32) | word);
            if (threadIdx.x == 0) arrive[ti] = 0;   // ready for the next launch (stream order)
        }
    }
}

template <int DT, int BITS, int UNROLL, bool CLAMP>
__global__ void __launch_bounds__(kThreads) quant_flat_kernel(const uint4* __restrict__ x, int64_t n_vec,
                                                              const uint32_t* __restrict__ absmax_bits, float eps,
                                                              char* __restrict__ q, void* __restrict__ scale_out,
                                                              float* __restrict__ scale_f32_out, FlatWork w,
                                                              const void* const* __restrict__ xs = nullptr,
                                                              void* const* __restrict__ qs = nullptr, Gather gat = Gather(),
                                                              uint32_t* __restrict__ absmax_out = nullptr) {
    constexpr int kOut = OutBytes<DT, BITS>::value;
    int ti; int64_t c0, c1;
    flat_split(w, ti, c0, c1);
    const uint4* xt = xs ? reinterpret_cast<const uint4*>(xs[ti]) : x;
    char* qt = xs ? reinterpret_cast<char*>(qs[ti]) : q;
    uint32_t word;
    if (gat.mine) {
        // every rank's word of this tensor for this step (written by the peers' absmax kernels), max-combined
        __shared__ uint32_t combined;
        if (threadIdx.x < 32) {
            uint32_t v = 0;
            if ((int)threadIdx.x < gat.world) {
                const unsigned long long* slot = xch_tagged(gat.mine, (int)(gat.seq & 1u), (int)threadIdx.x, ti);
                unsigned long long t, spins = 0;
                while ((uint32_t)((t = ld_acquire_sys_u64(slot)) >> 32) != gat.seq) {
                    if (++spins > (1ull << 24)) __trap();   // a peer that never publishes: fail instead of hanging the GPU (~10 s)
                    __nanosleep(spins < 64 ? 0 : 500);
                }
                v = (uint32_t)t;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (threadIdx.x == 0) combined = v;
        }
        __syncthreads();
        word = combined;
        if (c0 == 0 && threadIdx.x == 0 && absmax_out) absmax_out[ti] = word;
    } else {
        word = absmax_bits[ti];
    }
    const Scale sc = scale_from_bits<BITS>(word, eps);
    if (c0 == 0 && threadIdx.x == 0) {
        if (scale_out) store_scalar<DT>(scale_out, ti, sc.s);
        if (scale_f32_out) scale_f32_out[ti] = round_through<DT>(sc.s);
    }
    for (int64_t c = c0; c < c1; c += w.cstep) {
        const int64_t i = c * (kThreads * UNROLL) + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + (int64_t)u * kThreads;
            if (j < n_vec) v[u] = ld_stream(xt + j);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + (int64_t)u * kThreads;
            if (j < n_vec) quant_store<DT, BITS, CLAMP, CLAMP>(v[u], sc, qt + j * kOut);
        }
    }
}

// ---- strided rows: a "plane" is one (b, h); a CTA walks tiles of UNROLL * rows-per-pass tokens --------------

struct RowGeom {
    int nP, H;              // planes, and the divisor that splits a plane index into (b, h); H == 1: p -> b
    int64_t T;              // rows (tokens) per plane
    int64_t xB, xH, xT;     // input strides in BYTES
    int64_t qB, qH, qT;     // output strides in BYTES
    int vpr;                // 16-byte vectors per row on the 16-byte side
    int shift;              // log2(vpr) when vpr is a power of two, else -1
};

__device__ __forceinline__ void plane_offsets(const RowGeom& g, int p, int64_t& xo, int64_t& qo) {
    const int b = g.H == 1 ? p : p / g.H, h = p - b * g.H;
    xo = b * g.xB + h * g.xH;
    qo = b * g.qB + h * g.qH;
}
__device__ __forceinline__ void split_thread(const RowGeom& g, int tid, int& c, int& r0, int& rpp) {
    if (g.shift >= 0) { c = tid & (g.vpr - 1); r0 = tid >> g.shift; rpp = kThreads >> g.shift; }
    else { c = tid % g.vpr; r0 = tid / g.vpr; rpp = kThreads / g.vpr; }
}

template <int DT, int UNROLL>
__global__ void __launch_bounds__(kThreads) absmax_rows_kernel(const char* __restrict__ x, RowGeom g,
                                                               uint32_t* __restrict__ absmax_bits) {
    __shared__ uint32_t red[kThreads / 32];
    int c, r0, rpp;
    split_thread(g, threadIdx.x, c, r0, rpp);
    uint32_t m = 0;
    if (r0 < rpp) {
        const int64_t tile = (int64_t)rpp * UNROLL;
        for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
            int64_t xo, qo;
            plane_offsets(g, p, xo, qo);
            const char* base = x + xo + (int64_t)c * 16;
            for (int64_t t0 = (int64_t)blockIdx.x * tile + r0; t0 < g.T; t0 += (int64_t)gridDim.x * tile) {
                uint4 v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int64_t t = t0 + (int64_t)u * rpp;
                    v[u] = t < g.T ? ld_stream(base + t * g.xT) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) m = InVec<DT>::combine(m, InVec<DT>::absmax_bits(v[u]));
            }
        }
    }
    m = block_max_u32<kThreads>(InVec<DT>::finish_absmax(m), red);
    if (threadIdx.x == 0 && m != 0) atomicMax(absmax_bits, m);
}

template <int DT, int BITS, int UNROLL, bool CLAMP>
__global__ void __launch_bounds__(kThreads) quant_rows_kernel(const char* __restrict__ x, RowGeom g,
                                                              const uint32_t* __restrict__ absmax_bits,
                                                              int per_token, float eps, char* __restrict__ q,
                                                              void* __restrict__ scale_out,
                                                              float* __restrict__ scale_f32_out) {
    constexpr int kOut = OutBytes<DT, BITS>::value;
    __shared__ float2 tile_scale[kThreads * UNROLL];  // per-token (s, y) of the current tile
    int c, r0, rpp;
    split_thread(g, threadIdx.x, c, r0, rpp);
    const bool active = r0 < rpp;
    Scale sc = scale_from_bits<BITS>(per_token ? 0u : absmax_bits[0], eps);
    if (!per_token && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        if (scale_out) store_scalar<DT>(scale_out, 0, sc.s);
        if (scale_f32_out) scale_f32_out[0] = round_through<DT>(sc.s);
    }
    const int tile = rpp * UNROLL;
    const int64_t in_step = (int64_t)rpp * g.xT, out_step = (int64_t)rpp * g.qT;
    for (int64_t tb = (int64_t)blockIdx.x * tile; tb < g.T; tb += (int64_t)gridDim.x * tile) {
        bool first = true;
        for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
            // issue this plane's loads before anything that waits (the scale words, the barrier)
            int64_t xo, qo;
            plane_offsets(g, p, xo, qo);
            const char* xin = x + xo + (int64_t)c * 16 + (tb + r0) * g.xT;
            char* qout = q + qo + (int64_t)c * kOut + (tb + r0) * g.qT;
            uint4 v[UNROLL];
            if (active) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
                    if (tb + r0 + u * rpp < g.T) v[u] = ld_stream(xin + u * in_step);
            }
            if (per_token && first) {  // once per tile: the tokens' scales, shared by every plane
                __syncthreads();       // readers of the previous tile are done
                for (int i = threadIdx.x; i < tile; i += kThreads) {
                    const int64_t t = tb + i;
                    if (t < g.T) {
                        const Scale s1 = scale_from_bits<BITS>(absmax_bits[t], eps);
                        tile_scale[i] = make_float2(s1.s, s1.y);
                        if (blockIdx.y == 0) {
                            if (scale_out) store_scalar<DT>(scale_out, t, s1.s);
                            if (scale_f32_out) scale_f32_out[t] = round_through<DT>(s1.s);
                        }
                    }
                }
                __syncthreads();
                first = false;
            }
            if (active) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (tb + r0 + u * rpp >= g.T) continue;
                    if (per_token) { const float2 s2 = tile_scale[r0 + u * rpp]; sc.s = s2.x; sc.y = s2.y; }
                    quant_store<DT, BITS, CLAMP, CLAMP>(v[u], sc, qout + u * out_step);
                }
            }
        }
    }
}

// ---- single pass per token: one CTA owns token t across all planes -----------------------------------
// NV > 0: the token (nP * vpr vectors) is held in registers, NV vectors per thread: one HBM read.
// NV == 0: any size; two loops over the token, the second one hits L2.
// MODE 0: quantize (+ scales, + absmax out); MODE 1: absmax only (phase 1 of the head-sharded path).

// Batched form (tb.xs != nullptr): blockIdx.y is the tensor -- all layers' K (or V) streams of a step share shape, strides and
// bit width; destination rows start `row_off` tokens into every store (QuantizedKVCache.append_from_past /
// init_from_prompt_past loop `for layer: quantize(k); quantize(v)`, ops.py:323-342: 2L launches -> one per bit width).
struct TokBatch {
    const void* const* xs; void* const* qs; void* const* scales; float* const* scales_f32;
    int64_t row_off;  // tokens
};

// TPT: threads per token.  256: one token per CTA at a time (block-wide max).  32 / 64 / 128: 8 / 4 / 2 tokens per CTA at a time,
// each owned by TPT threads -- for small tokens (a head shard at N = 8 is 4 KB per token: one vector per thread and two block
// barriers per 4 KB with TPT = 256; with TPT = 32 a warp owns the token, eight vectors per lane, and the max is a REDUX).
template <int DT, int BITS, int NV, int MODE, int TPT = kThreads, bool XCH = false>
__global__ void __launch_bounds__(kThreads) token_fused_kernel(const char* __restrict__ x, RowGeom g, float eps,
                                                              char* __restrict__ q, void* __restrict__ scale_out,
                                                              float* __restrict__ scale_f32_out,
                                                              uint32_t* __restrict__ absmax_out, int accumulate,
                                                              TokBatch tb = TokBatch{nullptr, nullptr, nullptr, nullptr, 0},
                                                              Publish xc = Publish()) {
    __shared__ uint32_t red[kThreads / 32];
    static_assert(TPT == kThreads || (NV > 0 && (TPT == 32 || TPT == 64 || TPT == 128)), "TPT");
    constexpr int TOKS = kThreads / TPT;   // tokens in flight per CTA
    if (tb.xs) {
        x = reinterpret_cast<const char*>(tb.xs[blockIdx.y]);
        q = reinterpret_cast<char*>(tb.qs[blockIdx.y]) + tb.row_off * g.qT;
        scale_out = tb.scales ? reinterpret_cast<char*>(tb.scales[blockIdx.y]) + tb.row_off * ElemT<DT>::kSize : nullptr;
        scale_f32_out = tb.scales_f32 ? tb.scales_f32[blockIdx.y] + tb.row_off : nullptr;
    }
    constexpr int kOut = OutBytes<DT, BITS>::value;
    constexpr int kNV = NV > 0 ? NV : 1;
    // The max is taken over this very data, so the clamp can never bind (see code_word).  Whether the two dead FMNMX
    // stay in is purely a scheduling choice, measured on B200 (1 GiB bf16 tensor): INT8 5 770 GB/s with them vs 5 340
    // without; INT4 5 410 with vs 5 950 without.
    constexpr bool kFusedClamp = BITS == 8;
    const int n_vec = g.nP * g.vpr;
    const int u = TPT == kThreads ? threadIdx.x : threadIdx.x % TPT, slot = TPT == kThreads ? 0 : threadIdx.x / TPT;

    // per-thread vector -> (plane, column) offsets do not depend on the token: compute them once
    // (32-bit byte offsets: the host only picks NV > 0 when they fit)
    constexpr uint32_t kNone = 0xffffffffu;
    uint32_t xoff[kNV], qoff[kNV];
    if (NV > 0) {
#pragma unroll
        for (int k = 0; k < kNV; ++k) {
            const int j = u + k * TPT;
            xoff[k] = kNone; qoff[k] = 0;
            if (j < n_vec) {
                const int p = g.shift >= 0 ? j >> g.shift : j / g.vpr, c = j - p * g.vpr;
                int64_t xo, qo;
                plane_offsets(g, p, xo, qo);
                xoff[k] = (uint32_t)(xo + (int64_t)c * 16);
                qoff[k] = (uint32_t)(qo + (int64_t)c * kOut);
            }
        }
    }
    // step number of the exchange form: the caller's (seq != 0), or -- capturable into a CUDA graph -- one more than the steps this
    // rank's buffer has completed; the last CTA to finish records the completion (all CTAs have read the counter by then)
    // (XCH is its own instantiation: the plain quantizer keeps its registers and schedule -- folding the exchange in as a run-time
    // branch cost the C3-shape per-token leg 8 %)
    uint32_t xseq = 0;
    if constexpr (XCH) {
        xseq = xc.seq;
        if (xseq == 0) xseq = *reinterpret_cast<volatile uint32_t*>(xch_step(xc.peers[xc.rank])) + 1u;
    }
    for (int64_t t0 = (int64_t)blockIdx.x * TOKS; t0 < g.T; t0 += (int64_t)gridDim.x * TOKS) {
        const int64_t t = t0 + slot;
        const bool live = t < g.T;   // (always true for TPT == kThreads)
        const char* xt = x + t * g.xT;
        char* qt = q + t * g.qT;
        uint4 v[kNV];
        uint32_t m = 0;
        if (NV > 0) {
#pragma unroll
            for (int k = 0; k < kNV; ++k) v[k] = (live && xoff[k] != kNone) ? ld_stream(xt + xoff[k]) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < kNV; ++k) m = InVec<DT>::combine(m, InVec<DT>::absmax_bits(v[k]));
        } else {
            for (int j = threadIdx.x; j < n_vec; j += kThreads) {
                const int p = g.shift >= 0 ? j >> g.shift : j / g.vpr, c = j - p * g.vpr;
                int64_t xo, qo;
                plane_offsets(g, p, xo, qo);
                m = InVec<DT>::combine(m, InVec<DT>::absmax_bits(ld_stream(xt + xo + (int64_t)c * 16)));
            }
        }
        if constexpr (TPT == kThreads) {
            m = block_max_u32<kThreads>(InVec<DT>::finish_absmax(m), red);
            if constexpr (XCH) {
                // head-sharded append in ONE kernel (kvq_quant_per_token_fused_exchange): the token's local word goes, tagged with
                // the step number, into every rank's peer-mapped buffer; then the CTA waits for the world's words of this token
                // and max-combines them -- the scale is over all heads (ops.py:330,:342).  Every CTA publishes before it waits and
                // both sides walk the tokens in the same order, so the oldest unfinished token of any rank always has its peers'
                // words published or about to be: no cycle.  The values stay in registers across the wait.
                __shared__ uint32_t combined;
                const int widx = (int)blockIdx.y * (int)g.T + (int)t, parity = (int)(xseq & 1u);
                if ((int)threadIdx.x < xc.world) {
                    void* peer = xc.peers[threadIdx.x];
                    st_release_sys_u64(xc.seq ? xch_tagged(peer, parity, xc.rank, widx) : xch_tagged2(peer, parity, xc.rank, widx),
                                       ((unsigned long long)xseq << 32) | m);
                }
                if (threadIdx.x < 32) {
                    uint32_t w = 0;
                    if ((int)threadIdx.x < xc.world) {
                        void* mine = xc.peers[xc.rank];
                        const unsigned long long* slot_p = xc.seq ? xch_tagged(mine, parity, (int)threadIdx.x, widx) : xch_tagged2(mine, parity, (int)threadIdx.x, widx);
                        unsigned long long tg, spins = 0;
                        while ((uint32_t)((tg = ld_acquire_sys_u64(slot_p)) >> 32) != xseq) {
                            if (++spins > (1ull << 24)) __trap();   // a peer that never arrives: fail instead of hanging the GPU
                            __nanosleep(spins < 64 ? 0 : 500);
                        }
                        w = (uint32_t)tg;
                    }
                    w = __reduce_max_sync(0xffffffffu, w);
                    if (threadIdx.x == 0) combined = w;
                }
                __syncthreads();
                m = combined;
            }
        } else {
            m = __reduce_max_sync(0xffffffffu, InVec<DT>::finish_absmax(m));
            if constexpr (TPT > 32) {   // the token's TPT / 32 warps combine through shared memory
                if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
                __syncthreads();
                const int w0 = slot * (TPT / 32);
#pragma unroll
                for (int w = 0; w < TPT / 32; ++w) m = max(m, red[w0 + w]);
            }
        }
        if (u == 0 && live && absmax_out) {
            if (accumulate) atomicMax(absmax_out + t, m); else absmax_out[t] = m;
        }
        if (MODE == 0) {
            const Scale sc = scale_from_bits<BITS>(m, eps);
            if (u == 0 && live) {
                if (scale_out) store_scalar<DT>(scale_out, t, sc.s);
                if (scale_f32_out) scale_f32_out[t] = round_through<DT>(sc.s);
            }
            if (NV > 0) {
#pragma unroll
                for (int k = 0; k < kNV; ++k)
                    if (live && xoff[k] != kNone) quant_store<DT, BITS, kFusedClamp>(v[k], sc, qt + qoff[k]);
            } else {
                for (int j = threadIdx.x; j < n_vec; j += kThreads) {
                    const int p = g.shift >= 0 ? j >> g.shift : j / g.vpr, c = j - p * g.vpr;
                    int64_t xo, qo;
                    plane_offsets(g, p, xo, qo);
                    quant_store<DT, BITS, kFusedClamp>(ld_stream(xt + xo + (int64_t)c * 16), sc, qt + qo + (int64_t)c * kOut);
                }
            }
        }
        if constexpr (TPT != 32) __syncthreads();  // `red` is reused by the next token
    }
    if constexpr (XCH) if (xc.seq == 0 && threadIdx.x == 0) {
        uint32_t* stp = xch_step(xc.peers[xc.rank]);
        __threadfence();
        if (atomicAdd(stp + 1, 1u) == gridDim.x * gridDim.y - 1u) {
            stp[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(stp) = xseq;
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
            *out = (uint8_t)(code_word<DT, 8, true, true>(load_as_float<DT>(x, xrow + j * g.xD), sc) & 0xffu);
        } else {
            const uint32_t hi = code_word<DT, 4, true, true>(load_as_float<DT>(x, xrow + 2 * j * g.xD), sc) & 0xfu;
            const uint32_t lo = (2 * j + 1 < g.D) ? (code_word<DT, 4, true, true>(load_as_float<DT>(x, xrow + (2 * j + 1) * g.xD), sc) & 0xfu) : 8u;
            *out = (uint8_t)((hi << 4) | lo);
        }
    }
}

// ---- dequantize: the unit is one 16-byte OUTPUT vector (8 fp16/bf16 or 4 fp32 values) --------------------------

template <int ODT, int BITS> struct CodeIn {
    static constexpr int kN = 16 / ElemT<ODT>::kSize;   // output elements per vector
    static constexpr int kBytes = kN * BITS / 8;        // code bytes feeding it: 8, 4 or 2
};

template <int BYTES> struct Raw;
template <> struct Raw<8> { uint2 v; __device__ __forceinline__ void load(const void* p) { v = ld_stream8(p); } };
template <> struct Raw<4> { uint32_t v; __device__ __forceinline__ void load(const void* p) { v = __ldg(reinterpret_cast<const uint32_t*>(p)); } };
template <> struct Raw<2> { uint32_t v; __device__ __forceinline__ void load(const void* p) { v = __ldg(reinterpret_cast<const unsigned short*>(p)); } };

// exact integer -> fp32 without the conversion pipe: drop the (biased) byte into the mantissa of 2^23
__device__ __forceinline__ float code8(uint32_t x /* w ^ 0x80808080 */, int k) {
    return __fadd_rn(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650 + k)), -8388736.f);  // 2^23 + 128
}
__device__ __forceinline__ float code4(uint32_t nib /* nibbles isolated per byte */, int k) {
    return __fadd_rn(__uint_as_float(__byte_perm(nib, 0x4B000000u, 0x7650 + k)), -8388616.f);  // 2^23 + 8
}

// Scale operand of the dequant kernels.  PK (packed) path: out dtype is 16-bit and the scale has that same
// dtype, so code * scale is ONE exact 16-bit multiply (code and scale are both exactly representable, the
// product is rounded once, like fp32-multiply-then-round whose fp32 product is exact): two elements per
// HMUL2 and no fp32 round trip.  General path: fp32 multiply, then one rounding to the out dtype.
template <bool PK> struct ScaleArg;
template <> struct ScaleArg<false> {
    float f;
    __device__ __forceinline__ void load(const void* p, int dtype, int64_t i) { f = load_scale(p, dtype, i); }
};
template <> struct ScaleArg<true> {
    uint32_t h2;  // the raw 16-bit scale in both halves
    __device__ __forceinline__ void load(const void* p, int, int64_t i) {
        h2 = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p) + i) * 0x10001u;
    }
};

template <int ODT> __device__ __forceinline__ uint32_t pk_sub_mul(uint32_t biased, uint32_t offset2, uint32_t s2) {
    if (ODT == KVQ_F16) {
        const __half2 c = __hsub2(*reinterpret_cast<const __half2*>(&biased), *reinterpret_cast<const __half2*>(&offset2));
        const __half2 r = __hmul2(c, *reinterpret_cast<const __half2*>(&s2));
        return *reinterpret_cast<const uint32_t*>(&r);
    }
    const __nv_bfloat162 c = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&biased), *reinterpret_cast<const __nv_bfloat162*>(&offset2));
    const __nv_bfloat162 r = __hmul2(c, *reinterpret_cast<const __nv_bfloat162*>(&s2));
    return *reinterpret_cast<const uint32_t*>(&r);
}

template <int ODT, int BITS, bool PK>
__device__ __forceinline__ void dequant_store(const Raw<CodeIn<ODT, BITS>::kBytes>& r, const ScaleArg<PK>& sa, void* out) {
    constexpr int kN = CodeIn<ODT, BITS>::kN;
    if constexpr (PK) {
        static_assert(ODT != KVQ_F32, "packed path is 16-bit only");
        uint32_t o[4];
        if constexpr (BITS == 4) {
            // word k = (hi nibble of byte k, lo nibble of byte k) as 16-bit floats MAGIC + nibble
            constexpr uint32_t magic2 = ODT == KVQ_F16 ? 0x64006400u : 0x43004300u;   // 1024 | 128
            constexpr uint32_t off2 = ODT == KVQ_F16 ? 0x64086408u : 0x43084308u;     // MAGIC + 8
            const uint32_t w = r.v, w4 = w >> 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = __byte_perm(w4, w, 0x0400 + k + (k << 8));  // [w4.b_k, w4.b0, w.b_k, w4.b0]
                o[k] = pk_sub_mul<ODT>((t & 0x000f000fu) | magic2, off2, sa.h2);
            }
        } else if constexpr (ODT == KVQ_F16) {
            // byte ^ 0x80 = code + 128 in [1, 255]; 0x6400 | u = 1024 + u exactly
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&r.v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = w[k >> 1] ^ 0x80808080u;
                const uint32_t t = __byte_perm(x, 0x64646464u, (k & 1) ? 0x4342 : 0x4140);  // [x.b0|2, 0x64, x.b1|3, 0x64]
                o[k] = pk_sub_mul<ODT>(t, 0x64806480u /* 1152 */, sa.h2);
            }
        } else {
            // bf16 holds 7 mantissa bits: sign-magnitude.  |code| <= 127 -> 0x4300 | a = 128 + a exactly
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&r.v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t wk = w[k >> 1];
                const uint32_t a = __vabsss4(wk), sg = wk & 0x80808080u;
                const uint32_t t = __byte_perm(a, 0x43434343u, (k & 1) ? 0x4342 : 0x4140);
                const uint32_t sgn = __byte_perm(sg, 0u, (k & 1) ? 0x3424 : 0x1404);  // [0, sg.b, 0, sg.b]
                o[k] = pk_sub_mul<ODT>(t, 0x43004300u /* 128 */, sa.h2) ^ sgn;
            }
        }
        st_stream(out, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
        const float s = sa.f;
        float f[kN];
        if constexpr (BITS == 8) {
            if constexpr (kN == 8) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&r.v);
                const uint32_t x0 = w[0] ^ 0x80808080u, x1 = w[1] ^ 0x80808080u;
#pragma unroll
                for (int i = 0; i < 4; ++i) { f[i] = __fmul_rn(code8(x0, i), s); f[4 + i] = __fmul_rn(code8(x1, i), s); }
            } else {
                const uint32_t x0 = r.v ^ 0x80808080u;
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = __fmul_rn(code8(x0, i), s);
            }
        } else {
            const uint32_t w = r.v, hi = (w >> 4) & 0x0f0f0f0fu, lo = w & 0x0f0f0f0fu;
#pragma unroll
            for (int i = 0; i < kN / 2; ++i) { f[2 * i] = __fmul_rn(code4(hi, i), s); f[2 * i + 1] = __fmul_rn(code4(lo, i), s); }
        }
        if constexpr (ODT == KVQ_F32)
            st_stream(out, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
        else
            st_stream(out, make_uint4(narrow2<ODT>(f[0], f[1]), narrow2<ODT>(f[2], f[3]), narrow2<ODT>(f[4 % kN], f[5 % kN]),
                                      narrow2<ODT>(f[6 % kN], f[7 % kN])));
    }
}

template <int ODT, int BITS, bool PK, int UNROLL>
__global__ void __launch_bounds__(kThreads) dequant_flat_kernel(const char* __restrict__ codes, int64_t n_vec,
                                                                const void* __restrict__ scale, int scale_dtype,
                                                                uint4* __restrict__ out, FlatWork w,
                                                                const void* const* __restrict__ cs = nullptr,
                                                                void* const* __restrict__ os = nullptr) {
    constexpr int kIn = CodeIn<ODT, BITS>::kBytes;
    int ti; int64_t c0, c1;
    flat_split(w, ti, c0, c1);  // batched form: tensor-major grid, one scale per tensor (see absmax_flat_kernel)
    const char* ct = cs ? reinterpret_cast<const char*>(cs[ti]) : codes;
    uint4* ot = cs ? reinterpret_cast<uint4*>(os[ti]) : out;
    ScaleArg<PK> sa;
    sa.load(scale, scale_dtype, ti);
    for (int64_t c = c0; c < c1; c += w.cstep) {
        const int64_t i = c * (kThreads * UNROLL) + threadIdx.x;
        Raw<kIn> r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + (int64_t)u * kThreads;
            if (j < n_vec) r[u].load(ct + j * kIn);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + (int64_t)u * kThreads;
            if (j < n_vec) dequant_store<ODT, BITS, PK>(r[u], sa, ot + j);
        }
    }
}

// Per-token scales on a [B,H,T,D] store whose (b,h) planes are T contiguous rows at a uniform plane stride (the
// QuantizedLayerKV buffer: capacity >= T per plane; get_kv() dequantizes the whole cache every decode step in the reference's
// loop, ops.py:213-269): the same short-CTA flat stream over the OUTPUT as above -- the strided-row kernel below is a persistent
// loop and 7-15 % slower on this write-heavy 1:2 / 1:4 stream.  Output vector j -> plane j / vpp, position j % vpp inside the
// plane (one 64-bit division per chunk, stepped per vector), token = position >> vpr_shift (vectors per row: a power of two).
template <int ODT, int BITS, bool PK, int UNROLL>
__global__ void __launch_bounds__(kThreads) dequant_flat_token_kernel(const char* __restrict__ codes, int64_t n_vec,
                                                                      const void* __restrict__ scale, int scale_dtype,
                                                                      uint4* __restrict__ out, FlatWork w, int vpr_shift,
                                                                      int64_t vpp, int64_t plane_stride) {
    constexpr int kIn = CodeIn<ODT, BITS>::kBytes;
    int ti; int64_t c0, c1;
    flat_split(w, ti, c0, c1);
    const uint32_t vpp32 = (uint32_t)vpp;  // the host takes this path only for planes of fewer than 2^31 vectors
    for (int64_t c = c0; c < c1; c += w.cstep) {
        const int64_t base = c * (kThreads * UNROLL);
        const int64_t plane0 = base / vpp;  // uniform over the CTA: one 64-bit division per chunk, 32-bit steps per vector
        const uint32_t within0 = (uint32_t)(base - plane0 * vpp);
        const char* cbase = codes + plane0 * plane_stride;
        Raw<kIn> r[UNROLL];
        ScaleArg<PK> sv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = base + u * kThreads + threadIdx.x;
            if (j < n_vec) {
                uint32_t pos = within0 + u * kThreads + threadIdx.x;
                const char* cp = cbase;
                while (pos >= vpp32) { pos -= vpp32; cp += plane_stride; }  // a chunk reaches into the next plane(s) only at a plane's end
                r[u].load(cp + (size_t)pos * kIn);
                sv[u].load(scale, scale_dtype, pos >> vpr_shift);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = base + u * kThreads + threadIdx.x;
            if (j < n_vec) dequant_store<ODT, BITS, PK>(r[u], sv[u], out + j);
        }
    }
}

// g.x* describe the CODES (bytes), g.q* the OUTPUT (bytes); vpr = 16-byte OUTPUT vectors per row
template <int ODT, int BITS, bool PK, int UNROLL>
__global__ void __launch_bounds__(kThreads) dequant_rows_kernel(const char* __restrict__ codes, RowGeom g,
                                                                const void* __restrict__ scale, int scale_dtype,
                                                                int per_token, char* __restrict__ out) {
    constexpr int kIn = CodeIn<ODT, BITS>::kBytes;
    int c, r0, rpp;
    split_thread(g, threadIdx.x, c, r0, rpp);
    if (r0 >= rpp) return;
    ScaleArg<PK> s0;
    s0.load(scale, scale_dtype, 0);
    const int64_t tile = (int64_t)rpp * UNROLL;
    const int64_t in_step = (int64_t)rpp * g.xT, out_step = (int64_t)rpp * g.qT;
    for (int p = blockIdx.y; p < g.nP; p += gridDim.y) {
        int64_t co, oo;
        plane_offsets(g, p, co, oo);
        for (int64_t t0 = (int64_t)blockIdx.x * tile + r0; t0 < g.T; t0 += (int64_t)gridDim.x * tile) {
            const char* cin = codes + co + (int64_t)c * kIn + t0 * g.xT;
            char* o = out + oo + (int64_t)c * 16 + t0 * g.qT;
            Raw<kIn> r[UNROLL];
            ScaleArg<PK> sv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (t0 + (int64_t)u * rpp < g.T) {
                    r[u].load(cin + u * in_step);
                    if (per_token) sv[u].load(scale, scale_dtype, t0 + (int64_t)u * rpp); else sv[u] = s0;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (t0 + (int64_t)u * rpp < g.T) dequant_store<ODT, BITS, PK>(r[u], sv[u], o + u * out_step);
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
    return v.sD == 1 && (v.D * es) % 16 == 0 && v.D * es / 16 <= kThreads && aligned16(x) && ok(v.sB) && ok(v.sH) && ok(v.sT) &&
           v.B * v.H * v.T * v.D > 0 && v.B * v.H <= 0x7fffffff;
}
// codes side of a quantize: rows of D*bits/8 bytes written in pieces of `piece` bytes
bool codes_ok(const void* q, const Stride3& s, int piece) {
    return (reinterpret_cast<uintptr_t>(q) % piece) == 0 && s.sB % piece == 0 && s.sH % piece == 0 && s.sT % piece == 0;
}
int ilog2_exact(int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; }

// x: strides in elements of size es; q: strides in bytes (codes) or elements of size qes
RowGeom row_geom(const View4& v, int es, const Stride3& q, int qes, int vpr) {
    RowGeom g{};
    g.nP = (int)(v.B * v.H); g.H = (int)v.H; g.T = v.T;
    g.xB = v.sB * es; g.xH = v.sH * es; g.xT = v.sT * es;
    g.qB = q.sB * qes; g.qH = q.sH * qes; g.qT = q.sT * qes;
    g.vpr = vpr; g.shift = ilog2_exact(vpr);
    if (v.B == 1) { g.H = 1; g.xB = g.xH; g.qB = g.qH; }
    else if (v.H == 1) { g.H = 1; }
    else if (v.sB == v.H * v.sH && q.sB == v.H * q.sH) { g.H = 1; g.xB = g.xH; g.qB = g.qH; }  // uniform planes
    return g;
}

// grid sizes of the strided-row and per-token kernels, in CTAs per SM
// (compile-time constants: -DKVQ_..._CTAS_PER_SM=n rebuilds for a sweep; a library does not read tuning knobs from the environment)
#ifndef KVQ_QROWS_CTAS_PER_SM
#define KVQ_QROWS_CTAS_PER_SM 32
#endif
#ifndef KVQ_ROWS_CTAS_PER_SM
#define KVQ_ROWS_CTAS_PER_SM 128
#endif
#ifndef KVQ_TOKEN_CTAS_PER_SM
#define KVQ_TOKEN_CTAS_PER_SM 32
#endif
#ifndef KVQ_FLAT_MODE_ABSMAX
#define KVQ_FLAT_MODE_ABSMAX -128
#endif
#ifndef KVQ_FLAT_MODE_QUANT
#define KVQ_FLAT_MODE_QUANT -128
#endif
#ifndef KVQ_FLAT_MODE_DEQUANT
#define KVQ_FLAT_MODE_DEQUANT 1
#endif
#ifndef KVQ_FLAT_MIN_CPC
#define KVQ_FLAT_MIN_CPC 3
#endif
// measured on B200 (profiles/r1c_grid_sweep.log): many short CTAs beat a few persistent ones here too; the quantize rows
// kernel (two block barriers per tile) peaks at 32 per SM, the barrier-free dequantize / absmax rows kernels at 128
constexpr int quant_rows_ctas_per_sm() { return KVQ_QROWS_CTAS_PER_SM; }
constexpr int rows_ctas_per_sm() { return KVQ_ROWS_CTAS_PER_SM; }
constexpr int token_ctas_per_sm() { return KVQ_TOKEN_CTAS_PER_SM; }

// planes are looped inside the CTA (the per-tile scales are reused): y covers at most 8 plane groups
dim3 row_grid_planes_inside(const RowGeom& g, int unroll) {
    const int rpp = kThreads / g.vpr;
    const int64_t tiles = (g.T + (int64_t)rpp * unroll - 1) / ((int64_t)rpp * unroll);
    const int64_t total = (int64_t)sm_count() * quant_rows_ctas_per_sm();
    int64_t gx = tiles < total ? tiles : total;
    int64_t gy = (total + gx - 1) / gx;
    if (gy > g.nP) gy = g.nP;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)(gx > 0 ? gx : 1), (unsigned)(gy > 0 ? gy : 1), 1);
}

dim3 row_grid(const RowGeom& g, int unroll) {
    const int rpp = kThreads / g.vpr;
    const int64_t tiles = (g.T + (int64_t)rpp * unroll - 1) / ((int64_t)rpp * unroll);
    const unsigned gy = (unsigned)(g.nP < 65535 ? g.nP : 65535);
    const int64_t want = ((int64_t)sm_count() * rows_ctas_per_sm() + gy - 1) / gy;
    const int64_t gx = tiles < want ? tiles : want;
    return dim3((unsigned)(gx > 0 ? gx : 1), gy, 1);
}
// Work split of the flat kernels per kernel kind (0 absmax, 1 quantize, 2 dequantize).  v > 0: v consecutive chunks per CTA;
// v < 0: |v| CTAs per SM per tensor, each striding over the tensor.  -DKVQ_FLAT_MODE_{ABSMAX,QUANT,DEQUANT}=v overrides at build time.
// Measured on B200 with 1 GiB bf16 tensors (profiles/r1c_grid_sweep.log): the write-heavy dequantize stream wants one
// 16 KB chunk per CTA (6.6 TB/s vs 5.5 with 16 persistent CTAs per SM), the issue-bound quantize and the read-only absmax
// want ~128 short CTAs per SM.
constexpr int flat_mode(int kind) { return kind == 0 ? KVQ_FLAT_MODE_ABSMAX : kind == 1 ? KVQ_FLAT_MODE_QUANT : KVQ_FLAT_MODE_DEQUANT; }
void flat_work(int64_t n_vec, int unroll, int kind, int64_t n_tensors, FlatWork& w, unsigned& grid) {
    const int64_t per = (int64_t)kThreads * unroll;
    w.chunks = (n_vec + per - 1) / per;
    const int mode = flat_mode(kind);
    const int64_t max_bpt = 0x7fffffffLL / (n_tensors > 0 ? n_tensors : 1);
    int64_t bpt, cpc;
    if (mode < 0) {
        bpt = (int64_t)sm_count() * -mode;
        // small tensors (a head shard at N = 8 is 128 MB): keep ~3 strided chunks per CTA rather than degenerating into one chunk
        // per CTA, which costs the issue-bound quantize kernel 15-20 % (profiles/r1c_grid_sweep.log, table 2 row cpc = 1)
        constexpr int min_cpc = KVQ_FLAT_MIN_CPC;
        if (kind != 2 && min_cpc > 1 && bpt * min_cpc > w.chunks) bpt = (w.chunks + min_cpc - 1) / min_cpc;
        if (bpt > w.chunks) bpt = w.chunks;
        if (bpt > max_bpt) bpt = max_bpt;
        if (bpt < 1) bpt = 1;
        cpc = (w.chunks + bpt - 1) / bpt;
        w.cstep = (int)bpt;
    } else {
        cpc = mode;
        bpt = (w.chunks + cpc - 1) / cpc;
        while (bpt > max_bpt) { cpc *= 2; bpt = (w.chunks + cpc - 1) / cpc; }
        if (bpt < 1) bpt = 1;
        w.cstep = 1;
    }
    w.bpt = (int)bpt; w.cpc = (int)cpc;
    if (w.cstep != 1 && bpt == 1) w.cstep = 1, w.cpc = (int)w.chunks;  // a single CTA walks consecutive chunks
    grid = (unsigned)(bpt * n_tensors);
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
    // the clamp-free, division-free instantiation is proved exact for the phase-1 absmax of this very data and the default eps only
    const bool covers = (granularity & KVQ_ABSMAX_COVERS_INPUT) != 0 && eps == 1e-8f;
    granularity &= ~KVQ_ABSMAX_COVERS_INPUT;
    KVQ_REQUIRE(granularity == KVQ_PER_TENSOR || granularity == KVQ_PER_TOKEN, "bad granularity %d", granularity);
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int64_t W = BITS == 8 ? v.D : (v.D + 1) / 2;
    const Stride3 qs = make_stride3(q_stride3, v.H, v.T, W);
    const int64_t n = v.B * v.H * v.T * v.D;
    const int per_token = granularity == KVQ_PER_TOKEN;
    if (n == 0) return KVQ_OK;  // the Python wrapper raises like torch's max() on an empty tensor
    const int es = elem_size(dtype);
    const int piece = (16 / es) * BITS / 8;  // code bytes produced per input vector
    const bool q_contig = qs.sT == W && qs.sH == v.T * W && qs.sB == v.H * v.T * W;
    if (!per_token && is_contiguous(v) && q_contig && aligned16(x) && (n * es) % 16 == 0 && v.D % 2 == 0 &&
        codes_ok(q, qs, piece)) {
        const int64_t n_vec = n * es / 16;
        constexpr int U = 4;
        FlatWork w; unsigned grid;
        flat_work(n_vec, U, 1, 1, w, grid);
        if (covers) {
            KVQ_DISPATCH_DT(dtype, quant_flat_kernel<DT, BITS, U, false><<<grid, kThreads, 0, st>>>(
                                       (const uint4*)x, n_vec, absmax_bits, eps, (char*)q, scale_out, scale_f32_out, w));
        } else {
            KVQ_DISPATCH_DT(dtype, quant_flat_kernel<DT, BITS, U, true><<<grid, kThreads, 0, st>>>(
                                       (const uint4*)x, n_vec, absmax_bits, eps, (char*)q, scale_out, scale_f32_out, w));
        }
    } else if (fast_input_ok(x, dtype, v) && codes_ok(q, qs, piece)) {
        const RowGeom g = row_geom(v, es, qs, 1, (int)(v.D * es / 16));
        constexpr int U = 4;
        if (covers) {
            KVQ_DISPATCH_DT(dtype, quant_rows_kernel<DT, BITS, U, false><<<row_grid_planes_inside(g, U), kThreads, 0, st>>>(
                                       (const char*)x, g, absmax_bits, per_token, eps, (char*)q, scale_out, scale_f32_out));
        } else {
            KVQ_DISPATCH_DT(dtype, quant_rows_kernel<DT, BITS, U, true><<<row_grid_planes_inside(g, U), kThreads, 0, st>>>(
                                       (const char*)x, g, absmax_bits, per_token, eps, (char*)q, scale_out, scale_f32_out));
        }
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
    KVQ_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "bad out dtype %d", out_dtype);
    const int64_t B = shape[0], H = shape[1], T = shape[2], D = shape[3];
    KVQ_REQUIRE(B >= 0 && H >= 0 && T >= 0 && D >= 0, "negative shape");
    const int64_t W = BITS == 8 ? D : (D + 1) / 2;
    const Stride3 cs = make_stride3(q_stride3, H, T, W);
    const int per_token = granularity == KVQ_PER_TOKEN;
    const int64_t n = B * H * T * D;
    if (n == 0) return KVQ_OK;
    const int oes = elem_size(out_dtype);
    const int piece = (16 / oes) * BITS / 8;  // code bytes consumed per 16-byte output vector
    const bool c_contig = cs.sT == W && cs.sH == T * W && cs.sB == H * T * W;
    const bool pk = scale_dtype == out_dtype && out_dtype != KVQ_F32 && (reinterpret_cast<uintptr_t>(scale) % 2) == 0;
    const bool rows_ok = (D * oes) % 16 == 0 && D * oes / 16 <= kThreads && aligned16(out) && codes_ok(codes, cs, piece) &&
                         B * H <= 0x7fffffff && (BITS == 8 || D % 2 == 0);
    if (!per_token && c_contig && rows_ok) {
        const int64_t n_vec = n * oes / 16;
        constexpr int U = 4;
        FlatWork w; unsigned grid;
        flat_work(n_vec, U, 2, 1, w, grid);
        if (pk) {
            if (out_dtype == KVQ_F16) dequant_flat_kernel<KVQ_F16, BITS, true, U><<<grid, kThreads, 0, st>>>((const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w);
            else dequant_flat_kernel<KVQ_BF16, BITS, true, U><<<grid, kThreads, 0, st>>>((const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w);
        } else {
            KVQ_DISPATCH_DT(out_dtype, dequant_flat_kernel<DT, BITS, false, U><<<grid, kThreads, 0, st>>>(
                                           (const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w));
        }
    } else if (per_token && rows_ok && cs.sT == W && cs.sB == H * cs.sH && ((D * oes / 16) & (D * oes / 16 - 1)) == 0 &&
               T * (D * oes / 16) < (1LL << 30)) {
        // planes of T contiguous rows at a uniform stride (the cache store: capacity >= T): flat stream over the output
        const int64_t n_vec = n * oes / 16;
        constexpr int U = 4;
        FlatWork w; unsigned grid;
        flat_work(n_vec, U, 2, 1, w, grid);
        int sh = 0;
        while ((1 << sh) < D * oes / 16) ++sh;
        const int64_t vpp = T * (D * oes / 16);
        if (pk) {
            if (out_dtype == KVQ_F16) dequant_flat_token_kernel<KVQ_F16, BITS, true, U><<<grid, kThreads, 0, st>>>((const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w, sh, vpp, cs.sH);
            else dequant_flat_token_kernel<KVQ_BF16, BITS, true, U><<<grid, kThreads, 0, st>>>((const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w, sh, vpp, cs.sH);
        } else {
            KVQ_DISPATCH_DT(out_dtype, dequant_flat_token_kernel<DT, BITS, false, U><<<grid, kThreads, 0, st>>>(
                                           (const char*)codes, n_vec, scale, scale_dtype, (uint4*)out, w, sh, vpp, cs.sH));
        }
    } else if (rows_ok) {
        const int64_t contig[4] = {H * T * D, T * D, D, 1};
        View4 cv = make_view(shape, contig);
        cv.sB = cs.sB; cv.sH = cs.sH; cv.sT = cs.sT;
        const RowGeom g = row_geom(cv, 1, Stride3{contig[0], contig[1], contig[2]}, oes, (int)(D * oes / 16));
        constexpr int U = 4;
        if (pk) {
            if (out_dtype == KVQ_F16) dequant_rows_kernel<KVQ_F16, BITS, true, U><<<row_grid(g, U), kThreads, 0, st>>>((const char*)codes, g, scale, scale_dtype, per_token, (char*)out);
            else dequant_rows_kernel<KVQ_BF16, BITS, true, U><<<row_grid(g, U), kThreads, 0, st>>>((const char*)codes, g, scale, scale_dtype, per_token, (char*)out);
        } else {
            KVQ_DISPATCH_DT(out_dtype, dequant_rows_kernel<DT, BITS, false, U><<<row_grid(g, U), kThreads, 0, st>>>(
                                           (const char*)codes, g, scale, scale_dtype, per_token, (char*)out));
        }
    } else {
        GenGeom g{B, H, T, D, cs.sB, cs.sH, cs.sT, 1, H * T * D, T * D, D};
        KVQ_DISPATCH_DT(out_dtype, dequant_generic_kernel<DT, BITS><<<gen_grid(n), kThreads, 0, st>>>(
                                       codes, g, scale, scale_dtype, per_token, out));
    }
    return check_launch("dequant");
}

#ifndef KVQ_TOKEN_TPT_MAX_VEC
// Tokens of at most this many 16-byte vectors (4 KB: the per-GPU token of an 8-way head shard at the C3 shape) take the
// several-tokens-per-CTA form.  Measured on one B200, batched INT4 / INT8, bf16 (tools/pertoken_profile.py):
//   256 vectors per token: 3.8 -> 5.9 TB/s;  512: 5.4 (this form) vs 5.9 (one token per CTA);  1024: 5.3 vs 6.4.
#define KVQ_TOKEN_TPT_MAX_VEC 256
#endif
template <int DT, int BITS, int MODE>
void launch_token_fused(int nv, dim3 grid, cudaStream_t st, const void* x, const RowGeom& g, float eps, void* q,
                        void* scale_out, float* scale_f32_out, uint32_t* absmax_out, int accumulate,
                        const TokBatch& tb = TokBatch{nullptr, nullptr, nullptr, nullptr, 0}, const Publish& xc = Publish()) {
#define KVQ_TF(NV) token_fused_kernel<DT, BITS, NV, MODE><<<grid, kThreads, 0, st>>>((const char*)x, g, eps, (char*)q, scale_out, scale_f32_out, absmax_out, accumulate, tb, xc)
#define KVQ_TFT(TPT) token_fused_kernel<DT, BITS, 8, MODE, TPT><<<dim3((unsigned)tgrid, grid.y), kThreads, 0, st>>>((const char*)x, g, eps, (char*)q, scale_out, scale_f32_out, absmax_out, accumulate, tb)
    // small tokens, many of them: several tokens per CTA, TPT threads each (eight vectors per thread)
    if (MODE == 0 && !xc.peers && nv >= 1 && (int64_t)g.nP * g.vpr <= KVQ_TOKEN_TPT_MAX_VEC && g.T >= 1024) {
        const int64_t n_vec = (int64_t)g.nP * g.vpr;
        const int tpt = n_vec <= 8 * 32 ? 32 : n_vec <= 8 * 64 ? 64 : 128;
        const int64_t groups = (g.T + (kThreads / tpt) - 1) / (kThreads / tpt);
        // as many BYTES per CTA as the one-token form walks (cap CTAs x one token each per sweep): cap / (tokens per CTA) CTAs
        int64_t cap = (int64_t)sm_count() * token_ctas_per_sm() / (kThreads / tpt);
        if (cap < 2 * sm_count()) cap = 2 * sm_count();
        const int64_t tgrid = groups < cap ? groups : cap;
        if (tpt == 32) KVQ_TFT(32); else if (tpt == 64) KVQ_TFT(64); else KVQ_TFT(128);
        return;
    }
    if (xc.peers) {
        if constexpr (MODE == 0) {
#define KVQ_TFX(NV) token_fused_kernel<DT, BITS, NV, 0, kThreads, true><<<grid, kThreads, 0, st>>>((const char*)x, g, eps, (char*)q, scale_out, scale_f32_out, absmax_out, accumulate, tb, xc)
            if (nv <= 1) KVQ_TFX(1);
            else if (nv <= 2) KVQ_TFX(2);
            else if (nv <= 4) KVQ_TFX(4);
            else if (nv <= 8) KVQ_TFX(8);
            else KVQ_TFX(0);
#undef KVQ_TFX
        }
        return;
    }
    if (nv <= 1) KVQ_TF(1);
    else if (nv <= 2) KVQ_TF(2);
    else if (nv <= 4) KVQ_TF(4);
    else if (nv <= 8) KVQ_TF(8);
    else KVQ_TF(0);
#undef KVQ_TF
#undef KVQ_TFT
}

dim3 token_grid(int64_t T) {
    const int64_t cap = (int64_t)sm_count() * token_ctas_per_sm();
    return dim3((unsigned)(T < cap ? T : cap));
}

}  // namespace
}  // namespace kvq

using namespace kvq;

extern "C" int kvq_absmax(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4], int granularity,
                          uint32_t* absmax_bits, int accumulate, kvq_stream_t stream) {
    KVQ_REQUIRE(x && absmax_bits && shape && stride, "null pointer");
    KVQ_REQUIRE(granularity == KVQ_PER_TENSOR || granularity == KVQ_PER_TOKEN, "bad granularity %d", granularity);
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int per_token = granularity == KVQ_PER_TOKEN;
    const int64_t words = per_token ? v.T : 1;
    const int64_t n = v.B * v.H * v.T * v.D;
    const int es = elem_size(dtype);
    const bool fast = fast_input_ok(x, dtype, v);
    // the per-token fused kernel overwrites (or atomically accumulates); everything else max-accumulates
    if (!accumulate && !(per_token && fast) && words > 0) KVQ_CUDA(cudaMemsetAsync(absmax_bits, 0, words * 4, st));
    if (n == 0) return KVQ_OK;
    if (fast && per_token) {
        const RowGeom g = row_geom(v, es, Stride3{0, 0, 0}, 1, (int)(v.D * es / 16));
        const int64_t n_vec = v.B * v.H * (v.D * es / 16);
        int nv = (int)((n_vec + kThreads - 1) / kThreads);
        const int64_t xmax = ((v.B - 1) * v.sB + (v.H - 1) * v.sH + v.D) * es;
        if (xmax >= 0xffffffffLL || xmax < 0 || n_vec >= (1 << 30)) nv = 9;
        KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 8, 1>(nv, token_grid(v.T), st, x, g, 0.f, nullptr, nullptr, nullptr,
                                                            absmax_bits, accumulate));
    } else if (is_contiguous(v) && aligned16(x) && (n * es) % 16 == 0 && !per_token) {
        const int64_t n_vec = n * es / 16;
        constexpr int U = 8;
        FlatWork w; unsigned grid;
        flat_work(n_vec, U, 0, 1, w, grid);
        KVQ_DISPATCH_DT(dtype, absmax_flat_kernel<DT, U><<<grid, kThreads, 0, st>>>((const uint4*)x, n_vec, absmax_bits, w));
    } else if (fast) {
        const RowGeom g = row_geom(v, es, Stride3{0, 0, 0}, 1, (int)(v.D * es / 16));
        constexpr int U = 8;
        KVQ_DISPATCH_DT(dtype, absmax_rows_kernel<DT, U><<<row_grid(g, U), kThreads, 0, st>>>((const char*)x, g, absmax_bits));
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

static int run_token_fused(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                           int bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                           float* scale_f32_out, uint32_t* absmax_bits_out, const Publish& xc, kvq_stream_t stream) {
    KVQ_REQUIRE(x && q && shape && stride, "null pointer");
    KVQ_REQUIRE(bits == 8 || bits == 4, "bits must be 8 or 4, got %d", bits);
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int64_t W = bits == 8 ? v.D : (v.D + 1) / 2;
    const Stride3 qs = make_stride3(q_stride3, v.H, v.T, W);
    if (v.B * v.H * v.T * v.D == 0) return KVQ_OK;
    const int es = elem_size(dtype);
    const int piece = (16 / es) * bits / 8;
    const int64_t n_vec = v.B * v.H * (v.D * es / 16);
    // (a custom eps takes the two-phase path below: its quantize kernels divide exactly, see code_word)
    if (eps == 1e-8f && fast_input_ok(x, dtype, v) && codes_ok(q, qs, piece) && n_vec < (1 << 30) && (bits == 8 || v.D % 2 == 0)) {
        const RowGeom g = row_geom(v, es, qs, 1, (int)(v.D * es / 16));
        int nv = (int)((n_vec + kThreads - 1) / kThreads);
        const int64_t xmax = ((v.B - 1) * v.sB + (v.H - 1) * v.sH + v.D) * es, qmaxo = (v.B - 1) * qs.sB + (v.H - 1) * qs.sH + W;
        if (xmax >= 0xffffffffLL || qmaxo >= 0xffffffffLL || xmax < 0 || qmaxo < 0) nv = 9;  // offsets need 64 bits: two-loop variant
        dim3 grid = token_grid(v.T);
        if (xc.peers) {
            KVQ_REQUIRE(v.T <= kXchMaxWords, "quant_per_token_fused_exchange: at most %d tokens per call (got %lld)", kXchMaxWords, (long long)v.T);
            grid = dim3((unsigned)v.T);   // one CTA per token: every token of the call is published without waiting for a slot of this grid
        }
        const TokBatch none{nullptr, nullptr, nullptr, nullptr, 0};
        if (bits == 8) {
            KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 8, 0>(nv, grid, st, x, g, eps, q, scale_out, scale_f32_out, absmax_bits_out, 0, none, xc));
        } else {
            KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 4, 0>(nv, grid, st, x, g, eps, q, scale_out, scale_f32_out, absmax_bits_out, 0, none, xc));
        }
        return check_launch("quant_per_token_fused");
    }
    if (xc.peers)
        return fail(KVQ_ERR_UNSUPPORTED, "quant_per_token_fused_exchange: needs the default eps and rows of 16-byte multiples with 16-byte-aligned "
                                         "strides (take the allreduce form for other layouts)");
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

extern "C" int kvq_quant_per_token_fused(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                                         int bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                                         float* scale_f32_out, uint32_t* absmax_bits_out, kvq_stream_t stream) {
    return run_token_fused(x, dtype, shape, stride, bits, eps, q, q_stride3, scale_out, scale_f32_out, absmax_bits_out, Publish(), stream);
}

static int check_exchange_args(void* const* peer_buffers, int rank, int world_size, uint32_t seq, int64_t n_tensors) {
    KVQ_REQUIRE(peer_buffers, "null pointer");
    KVQ_REQUIRE(world_size >= 1 && world_size <= kXchMaxWorld && rank >= 0 && rank < world_size, "bad rank / world size");
    KVQ_REQUIRE(n_tensors <= kXchMaxWords, "publish / gather: at most %d tensors per call (got %lld)", kXchMaxWords, (long long)n_tensors);
    KVQ_REQUIRE(seq != 0, "seq must be a non-zero step counter, the same on every rank (buffers start zeroed)");
    return KVQ_OK;
}

// The step counter of the exchange forms is a LAUNCH ARGUMENT: a captured graph would replay the same number and accept the previous
// replay's words.  Refuse instead (the caller falls back to the NCCL allreduce, which is capturable).
static int not_capturing(kvq_stream_t stream, const char* what) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    KVQ_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) return fail(KVQ_ERR_UNSUPPORTED, "%s: not capturable into a CUDA graph (the step counter is a launch argument)", what);
    return KVQ_OK;
}

extern "C" int kvq_quant_per_token_fused_exchange(const void* x, int dtype, const int64_t shape[4], const int64_t stride[4],
                                                  int bits, float eps, void* q, const int64_t q_stride3[3], void* scale_out,
                                                  float* scale_f32_out, uint32_t* absmax_bits_out, void* const* peer_buffers, int rank,
                                                  int world_size, uint32_t seq, kvq_stream_t stream) {
    if (int rc = check_exchange_args(peer_buffers, rank, world_size, seq ? seq : 1u, 0)) return rc;
    if (seq != 0)
        if (int rc = not_capturing(stream, "quant_per_token_fused_exchange")) return rc;
    KVQ_REQUIRE(shape && shape[0] * shape[1] * shape[2] * shape[3] != 0, "exchange: an empty input publishes nothing (every rank must publish every call)");
    Publish xc; xc.peers = peer_buffers; xc.rank = rank; xc.world = world_size; xc.seq = seq;
    return run_token_fused(x, dtype, shape, stride, bits, eps, q, q_stride3, scale_out, scale_f32_out, absmax_bits_out, xc, stream);
}

static int run_token_fused_batched(const void* const* xs, int dtype, int64_t n_tensors, const int64_t shape[4],
                                   const int64_t stride[4], int bits, float eps, void* const* qs,
                                   const int64_t q_stride3[3], int64_t dst_token_offset, void* const* scales_out,
                                   float* const* scales_f32_out, const Publish& xc, kvq_stream_t stream) {
    KVQ_REQUIRE(xs && qs && shape && stride, "null pointer");
    KVQ_REQUIRE(bits == 8 || bits == 4, "bits must be 8 or 4, got %d", bits);
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535 && dst_token_offset >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const View4 v = make_view(shape, stride);
    KVQ_REQUIRE(v.B >= 0 && v.H >= 0 && v.T >= 0 && v.D >= 0, "negative shape");
    const int64_t W = bits == 8 ? v.D : (v.D + 1) / 2;
    const Stride3 qst = make_stride3(q_stride3, v.H, v.T, W);
    if (v.B * v.H * v.T * v.D == 0 || n_tensors == 0) return KVQ_OK;
    const int es = elem_size(dtype);
    const int piece = (16 / es) * bits / 8;
    const int64_t n_vec = v.B * v.H * (v.D * es / 16);
    // (pointer alignment is the caller's promise here: the tables live on the device)
    if (!(eps == 1e-8f && fast_input_ok(reinterpret_cast<const void*>(uintptr_t(16)), dtype, v) && codes_ok(reinterpret_cast<const void*>(uintptr_t(16)), qst, piece) &&
          n_vec < (1 << 30) && (bits == 8 || v.D % 2 == 0)))
        return fail(KVQ_ERR_UNSUPPORTED, "quant_per_token_fused_batched: needs the default eps and rows of 16-byte multiples with 16-byte-aligned strides "
                                         "(use kvq_quant_per_token_fused per tensor for other layouts)");
    const RowGeom g = row_geom(v, es, qst, 1, (int)(v.D * es / 16));
    int nv = (int)((n_vec + kThreads - 1) / kThreads);
    const int64_t xmax = ((v.B - 1) * v.sB + (v.H - 1) * v.sH + v.D) * es, qmaxo = (v.B - 1) * qst.sB + (v.H - 1) * qst.sH + W;
    if (xmax >= 0xffffffffLL || qmaxo >= 0xffffffffLL || xmax < 0 || qmaxo < 0) nv = 9;
    const TokBatch tb{xs, qs, scales_out, scales_f32_out, dst_token_offset};
    // as many CTAs per tensor as a single-tensor launch: blocks are handed out x-fastest, so the resident set sweeps one tensor
    // after the other like separate launches would (64 interleaved streams of a few tokens each measured 8 % slower)
    int64_t gx = (int64_t)sm_count() * token_ctas_per_sm();
    if (gx > v.T) gx = v.T;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)n_tensors);
    if (bits == 8) {
        KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 8, 0>(nv, grid, st, nullptr, g, eps, nullptr, nullptr, nullptr, nullptr, 0, tb, xc));
    } else {
        KVQ_DISPATCH_DT(dtype, launch_token_fused<DT, 4, 0>(nv, grid, st, nullptr, g, eps, nullptr, nullptr, nullptr, nullptr, 0, tb, xc));
    }
    return check_launch("quant_per_token_fused_batched");
}

extern "C" int kvq_quant_per_token_fused_batched(const void* const* xs, int dtype, int64_t n_tensors, const int64_t shape[4],
                                                 const int64_t stride[4], int bits, float eps, void* const* qs,
                                                 const int64_t q_stride3[3], int64_t dst_token_offset, void* const* scales_out,
                                                 float* const* scales_f32_out, kvq_stream_t stream) {
    return run_token_fused_batched(xs, dtype, n_tensors, shape, stride, bits, eps, qs, q_stride3, dst_token_offset, scales_out, scales_f32_out,
                                   Publish(), stream);
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

// ---- batched (multi-tensor) forms: n contiguous tensors of `numel` elements each in ONE launch per phase ------------------
// One launch per phase for a whole step (all layers x {K,V}) instead of one per tensor: at 1/8 of a cache per GPU the
// per-tensor kernels last 20-30 us and launch gaps dominate.  Pointer tables live on the device.


static int run_absmax_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, uint32_t* absmax_bits, Publish pub,
                              kvq_stream_t stream) {
    KVQ_REQUIRE(xs && absmax_bits, "null pointer");
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535 && numel >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tensors == 0) return KVQ_OK;
    KVQ_CUDA(cudaMemsetAsync(absmax_bits, 0, n_tensors * 4, st));
    if (numel == 0) return KVQ_OK;
    const int es = elem_size(dtype);
    KVQ_REQUIRE((numel * es) % 16 == 0, "batched kernels need numel * elem_size to be a multiple of 16");
    const int64_t n_vec = numel * es / 16;
    constexpr int U = 8;
    FlatWork w; unsigned grid;
    flat_work(n_vec, U, 0, n_tensors, w, grid);
    KVQ_DISPATCH_DT(dtype, absmax_flat_kernel<DT, U><<<grid, kThreads, 0, st>>>(nullptr, n_vec, absmax_bits, w, xs, pub));
    return check_launch("absmax_batched");
}

extern "C" int kvq_absmax_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, uint32_t* absmax_bits,
                                  kvq_stream_t stream) {
    return run_absmax_batched(xs, dtype, n_tensors, numel, absmax_bits, Publish(), stream);
}

extern "C" int kvq_absmax_batched_publish(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, uint32_t* absmax_bits,
                                          void* const* peer_buffers, int rank, int world_size, uint32_t seq, kvq_stream_t stream) {
    if (int rc = check_exchange_args(peer_buffers, rank, world_size, seq, n_tensors)) return rc;
    if (int rc = not_capturing(stream, "absmax_batched_publish")) return rc;
    KVQ_REQUIRE(numel > 0 || n_tensors == 0, "publish: empty tensors publish nothing (every rank must publish every step)");
    Publish pub; pub.peers = peer_buffers; pub.rank = rank; pub.world = world_size; pub.seq = seq;
    return run_absmax_batched(xs, dtype, n_tensors, numel, absmax_bits, pub, stream);
}

static int run_quant_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, int bits,
                             const uint32_t* absmax_bits, float eps, void* const* qs, void* scale_out, float* scale_f32_out,
                             Gather gat, uint32_t* absmax_out, kvq_stream_t stream) {
    KVQ_REQUIRE(xs && qs && (absmax_bits || gat.mine), "null pointer");
    KVQ_REQUIRE(dtype >= 0 && dtype <= 2, "unknown dtype %d", dtype);
    const bool covers = (bits & KVQ_ABSMAX_COVERS_INPUT) != 0 && eps == 1e-8f;
    bits &= ~KVQ_ABSMAX_COVERS_INPUT;
    KVQ_REQUIRE(bits == 8 || bits == 4, "bits must be 8 or 4");
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535 && numel >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tensors == 0 || numel == 0) return KVQ_OK;
    const int es = elem_size(dtype);
    KVQ_REQUIRE((numel * es) % 16 == 0, "batched kernels need numel * elem_size to be a multiple of 16");
    const int64_t n_vec = numel * es / 16;
    constexpr int U = 4;
    FlatWork w; unsigned grid;
    flat_work(n_vec, U, 1, n_tensors, w, grid);
#define KVQ_QB(BITS, CL) KVQ_DISPATCH_DT(dtype, quant_flat_kernel<DT, BITS, U, CL><<<grid, kThreads, 0, st>>>( \
                                   nullptr, n_vec, absmax_bits, eps, nullptr, scale_out, scale_f32_out, w, xs, qs, gat, absmax_out))
    if (bits == 8) { if (covers) { KVQ_QB(8, false); } else { KVQ_QB(8, true); } }
    else { if (covers) { KVQ_QB(4, false); } else { KVQ_QB(4, true); } }
#undef KVQ_QB
    return check_launch("quant_batched");
}

extern "C" int kvq_quant_batched(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, int bits,
                                 const uint32_t* absmax_bits, float eps, void* const* qs, void* scale_out, float* scale_f32_out,
                                 kvq_stream_t stream) {
    return run_quant_batched(xs, dtype, n_tensors, numel, bits, absmax_bits, eps, qs, scale_out, scale_f32_out, Gather(), nullptr, stream);
}

extern "C" int kvq_quant_batched_gather(const void* const* xs, int dtype, int64_t n_tensors, int64_t numel, int bits,
                                        uint32_t* absmax_bits_out, float eps, void* const* qs, void* scale_out, float* scale_f32_out,
                                        void* local_buffer, int world_size, uint32_t seq, kvq_stream_t stream) {
    KVQ_REQUIRE(local_buffer, "null pointer");
    KVQ_REQUIRE(world_size >= 1 && world_size <= kXchMaxWorld, "bad world size");
    KVQ_REQUIRE(n_tensors <= kXchMaxWords, "publish / gather: at most %d tensors per call (got %lld)", kXchMaxWords, (long long)n_tensors);
    KVQ_REQUIRE(seq != 0, "seq must be the step counter given to kvq_absmax_batched_publish");
    if (int rc = not_capturing(stream, "quant_batched_gather")) return rc;
    Gather gat; gat.mine = local_buffer; gat.world = world_size; gat.seq = seq;
    return run_quant_batched(xs, dtype, n_tensors, numel, bits, nullptr, eps, qs, scale_out, scale_f32_out, gat, absmax_bits_out, stream);
}

extern "C" int kvq_dequant_batched(const void* const* codes, int bits, int64_t n_tensors, int64_t numel, const void* scales,
                                   int scale_dtype, void* const* outs, int out_dtype, kvq_stream_t stream) {
    KVQ_REQUIRE(codes && outs && scales, "null pointer");
    KVQ_REQUIRE(bits == 8 || bits == 4, "bits must be 8 or 4");
    KVQ_REQUIRE(scale_dtype >= 0 && scale_dtype <= 2 && out_dtype >= 0 && out_dtype <= 2, "bad dtype");
    KVQ_REQUIRE(n_tensors >= 0 && n_tensors <= 65535 && numel >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tensors == 0 || numel == 0) return KVQ_OK;
    const int oes = elem_size(out_dtype);
    KVQ_REQUIRE((numel * oes) % 16 == 0, "batched kernels need numel * elem_size to be a multiple of 16");
    const int64_t n_vec = numel * oes / 16;
    const bool pk = scale_dtype == out_dtype && out_dtype != KVQ_F32;
    constexpr int U = 4;
    FlatWork w; unsigned grid;
    flat_work(n_vec, U, 2, n_tensors, w, grid);
#define KVQ_DQB(ODT, BITS, PK) dequant_flat_kernel<ODT, BITS, PK, U><<<grid, kThreads, 0, st>>>(nullptr, n_vec, scales, scale_dtype, nullptr, w, codes, outs)
    if (bits == 8) {
        if (pk) { if (out_dtype == KVQ_F16) KVQ_DQB(KVQ_F16, 8, true); else KVQ_DQB(KVQ_BF16, 8, true); }
        else { KVQ_DISPATCH_DT(out_dtype, KVQ_DQB(DT, 8, false)); }
    } else {
        if (pk) { if (out_dtype == KVQ_F16) KVQ_DQB(KVQ_F16, 4, true); else KVQ_DQB(KVQ_BF16, 4, true); }
        else { KVQ_DISPATCH_DT(out_dtype, KVQ_DQB(DT, 4, false)); }
    }
#undef KVQ_DQB
    return check_launch("dequant_batched");
}
