// kvq_common.cuh -- shared device/host helpers of libkvq_sm100.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kvq.h"

namespace kvq {

// ----------------------------------------------------------------------------- host: errors

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> KVQ_ERR_CUDA

#define KVQ_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return ::kvq::fail(KVQ_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define KVQ_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return ::kvq::fail(KVQ_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int sm_count();  // cached per device; 148 on B200

// ----------------------------------------------------------------------------- host: 4-D views

// [B,H,T,D] view after merging (B,H) into "planes" where the strides allow it.
struct View4 {
    int64_t B, H, T, D;
    int64_t sB, sH, sT, sD;
    bool planes_uniform;  // sB == H * sH (or B == 1 / H == 1): plane p sits at p * sP
    int64_t sP;
};

inline View4 make_view(const int64_t shape[4], const int64_t stride[4]) {
    View4 v{shape[0], shape[1], shape[2], shape[3], stride[0], stride[1], stride[2], stride[3], false, 0};
    if (v.B == 1) { v.planes_uniform = true; v.sP = v.sH; }
    else if (v.H == 1) { v.planes_uniform = true; v.sP = v.sB; }
    else if (v.sB == v.H * v.sH) { v.planes_uniform = true; v.sP = v.sH; }
    return v;
}

inline bool is_contiguous(const View4& v) {
    return v.sD == 1 && v.sT == v.D && v.sH == v.T * v.D && v.sB == v.H * v.T * v.D;
}

// strides {sB,sH,sT} of a code/output tensor whose rows have `row` elements; NULL -> contiguous
struct Stride3 { int64_t sB, sH, sT; };
inline Stride3 make_stride3(const int64_t* s3, int64_t H, int64_t T, int64_t row) {
    if (s3) return Stride3{s3[0], s3[1], s3[2]};
    return Stride3{H * T * row, T * row, row};
}

inline int elem_size(int dtype) { return dtype == KVQ_F32 ? 4 : 2; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------- device: memory ops

#ifdef __CUDACC__

// streaming 128-bit load: read-only path, do not allocate in L1 (every byte is touched once)
__device__ __forceinline__ uint4 ld_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ----------------------------------------------------------------------------- device: element conversion

template <int DT> struct ElemT;
template <> struct ElemT<KVQ_F32> { static constexpr int kSize = 4; };
template <> struct ElemT<KVQ_F16> { static constexpr int kSize = 2; };
template <> struct ElemT<KVQ_BF16> { static constexpr int kSize = 2; };

// two packed 16-bit elements (one 32-bit word) -> two floats (exact widening)
template <int DT> __device__ __forceinline__ void widen2(uint32_t w, float& lo, float& hi) {
    if (DT == KVQ_BF16) {
        lo = __uint_as_float(w << 16);
        hi = __uint_as_float(w & 0xffff0000u);
    } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        lo = f.x; hi = f.y;
    }
}

// scalar load of element i as fp32 (generic path)
template <int DT> __device__ __forceinline__ float load_as_float(const void* p, int64_t i) {
    if (DT == KVQ_F32) return reinterpret_cast<const float*>(p)[i];
    if (DT == KVQ_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// fp32 -> dtype with round-to-nearest-even, as raw bits
template <int DT> __device__ __forceinline__ uint32_t narrow_bits(float v) {
    if (DT == KVQ_F32) return __float_as_uint(v);
    if (DT == KVQ_F16) return (uint32_t)__half_as_ushort(__float2half_rn(v));
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <int DT> __device__ __forceinline__ float round_through(float v) {
    if (DT == KVQ_F32) return v;
    if (DT == KVQ_F16) return __half2float(__float2half_rn(v));
    return __bfloat162float(__float2bfloat16_rn(v));
}
template <int DT> __device__ __forceinline__ void store_scalar(void* p, int64_t i, float v) {
    if (DT == KVQ_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (DT == KVQ_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
// two floats -> one packed word of two 16-bit elements (RN-even each)
template <int DT> __device__ __forceinline__ uint32_t narrow2(float lo, float hi) {
    if (DT == KVQ_F16) {
        const __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// load scale element i of dtype DT (runtime dtype) as fp32
__device__ __forceinline__ float load_scale(const void* p, int dtype, int64_t i) {
    if (dtype == KVQ_F32) return reinterpret_cast<const float*>(p)[i];
    if (dtype == KVQ_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// ----------------------------------------------------------------------------- device: packed fp32 pairs
// Blackwell's FFMA2 / FADD2 / FMUL2 (PTX fma.rn.f32x2 ...): two IEEE fp32 operations (round-to-nearest-even each, bit-identical to
// the scalar instructions) per issue slot; a pair lives in a 64-bit register.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ----------------------------------------------------------------------------- device: reductions

// max over the block of a uint32 (REDUX per warp, one smem round); result valid in every thread
template <int THREADS> __device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* smem /*THREADS/32*/) {
    v = __reduce_max_sync(0xffffffffu, v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    uint32_t r = lane < THREADS / 32 ? smem[lane] : 0u;
    r = __reduce_max_sync(0xffffffffu, r);
    return r;
}

#endif  // __CUDACC__


// ---- layout of one rank's peer-mapped exchange buffer (xchg.cu, and the publish / gather forms of the flat kernels) ----------
//   uint32 words  [2][kXchMaxWorld][kXchMaxWords]   kvq_absmax_exchange: plain words ...
//   uint32 flags  [2][kXchMaxWorld]                 ... and one sequence flag per (parity, source rank)
//   uint64 tagged [2][kXchMaxWorld][kXchMaxWords]   publish / gather: (seq << 32 | word), ONE 8-byte store per word, no flag
//   uint32 arrive [kXchMaxWords]                    local last-CTA counters of the publishing absmax kernel
//   uint64 tagged2[2][kXchMaxWorld][kXchMaxWords]   the same as `tagged` for calls that take their step number from the DEVICE
//   uint32 step   [4]                               counter below (seq == 0: capturable into CUDA graphs): [0] = steps completed,
//                                                   [1] = CTA arrivals of the running kernel
constexpr int kXchMaxWords = 1024;   // words per rank per parity
constexpr int kXchMaxWorld = 16;
__host__ __device__ inline size_t xch_word_off(int parity, int src_rank, int i) { return ((size_t)parity * kXchMaxWorld + src_rank) * kXchMaxWords + i; }
__host__ __device__ inline size_t xch_flag_off(int parity, int src_rank) { return (size_t)2 * kXchMaxWorld * kXchMaxWords + (size_t)parity * kXchMaxWorld + src_rank; }
constexpr size_t kXchTaggedByteOff = (size_t)(2 * kXchMaxWorld * kXchMaxWords + 2 * kXchMaxWorld) * 4;   // multiple of 8
constexpr size_t kXchArriveByteOff = kXchTaggedByteOff + (size_t)2 * kXchMaxWorld * kXchMaxWords * 8;
constexpr size_t kXchTagged2ByteOff = kXchArriveByteOff + (size_t)kXchMaxWords * 4;   // multiple of 8
constexpr size_t kXchStepByteOff = kXchTagged2ByteOff + (size_t)2 * kXchMaxWorld * kXchMaxWords * 8;
constexpr size_t kXchBufferBytes = kXchStepByteOff + 16;
__host__ __device__ inline unsigned long long* xch_tagged(void* buf, int parity, int src_rank, int i) {
    return reinterpret_cast<unsigned long long*>(static_cast<char*>(buf) + kXchTaggedByteOff) + xch_word_off(parity, src_rank, i);
}
__host__ __device__ inline unsigned long long* xch_tagged2(void* buf, int parity, int src_rank, int i) {
    return reinterpret_cast<unsigned long long*>(static_cast<char*>(buf) + kXchTagged2ByteOff) + xch_word_off(parity, src_rank, i);
}
__host__ __device__ inline uint32_t* xch_step(void* buf) { return reinterpret_cast<uint32_t*>(static_cast<char*>(buf) + kXchStepByteOff); }
__host__ __device__ inline uint32_t* xch_arrive(void* buf) { return reinterpret_cast<uint32_t*>(static_cast<char*>(buf) + kXchArriveByteOff); }
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace kvq
