// attn_common.cuh -- pieces shared by the decode-attention translation units (attn.cu: CUDA-core kernels, attn_mma.cu:
// tensor-map TMA + integer-MMA kernels).  sm_100a only.
#pragma once

#include "kvq_common.cuh"

namespace kvq {

constexpr int kTile = 128;  // tokens per tile

struct AttnParams {
    const void* q; int q_dtype;
    const uint8_t* k; int64_t kB, kH, kT; const float* k_scale;
    const uint8_t* v; int64_t vB, vH, vT; const float* v_scale;
    int B, Hq, H, T;
    float qk_scale_log2;  // sm_scale * log2(e)
    int tokens_per_split, n_splits;
    float* ws_acc;  // [B, Hq, n_splits, D]
    float* ws_ml;   // [B, Hq, n_splits, 2]
    // paged store (table != nullptr): k / v are block pools [n_phys][B][H][bs][row], k_scale / v_scale are [n_phys][bs];
    // logical block i of the sequence lives in physical block table[i]
    const int32_t* table; int bs, bs_shift; int64_t kBlk, vBlk;
};

// position of logical token t in the per-token scale arrays
template <bool PAGED> __device__ __forceinline__ int64_t tok_index(const AttnParams& P, int t) {
    if constexpr (!PAGED) return t;
    return (int64_t)__ldg(P.table + (t >> P.bs_shift)) * P.bs + (t & (P.bs - 1));
}


#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) { return __dp4a((int)a, (int)b, c); }
__device__ __forceinline__ int dp4a_us(uint32_t a_unsigned, uint32_t b_signed, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_unsigned), "r"(b_signed), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_su(uint32_t a_signed, uint32_t b_unsigned, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_signed), "r"(b_unsigned), "r"(c));
    return d;
}
// rows a,b,c,d of 4 bytes -> columns: col[j] = (a_j, b_j, c_j, d_j)
__device__ __forceinline__ void transpose4x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t col[4]) {
    const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
    const uint32_t t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
    col[0] = __byte_perm(t0, t1, 0x5410); col[1] = __byte_perm(t0, t1, 0x7632);
    col[2] = __byte_perm(t2, t3, 0x5410); col[3] = __byte_perm(t2, t3, 0x7632);
}

constexpr float kFixedMax = 8355000.f;  // < 127 * (65536 + 256 + 1): three balanced byte digits cannot overflow

__device__ __forceinline__ int float_order_key(float x) {  // monotone float -> int (an involution on the bit pattern)
    const int b = __float_as_int(x);
    return b >= 0 ? b : b ^ 0x7fffffff;
}


#endif  // __CUDACC__

// attn_mma.cu: the split kernel on tensor-map TMA + INT8 tensor-core MMA (variant KVQ_ATTN_IMMA).  Same grid, workspace and
// merge protocol as the CUDA-core kernels.  Returns KVQ_ERR_UNSUPPORTED when the layout cannot be described by a tensor map.
int launch_attn_mma(const AttnParams& P, int k_bits, int v_bits, int G, int D, dim3 grid, cudaStream_t st);
int attn_mma_ctas_per_sm(int G);

}  // namespace kvq
