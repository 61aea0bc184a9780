// attn.cu -- area 4: decode attention over a per-token-quantized KV cache (sm_100a): the CUDA-core kernel family, the
// generic any-shape kernel, the split merge, and the C entry points.  (The tensor-map TMA + INT8 MMA family: attn_mma.cu.)
//
// The reference dequantizes EVERY stored token, concatenates, wraps the result in a DynamicCache and
// lets transformers run softmax(q k^T / sqrt(D)) v (src/benchmarking/benchmarker.py:470-477,
// src/quantization/ops.py:213-269, :345-355).  Here the cache is read once, as codes: INT8 or packed
// INT4 K/V are dequantized in registers (code * scale[t]) inside a split-K pass with an online softmax.
//
//   decode_attn_mha_kernel     G = Hq/H = 1: per-warp DP4A over a cp.async.bulk ring (the bandwidth-bound GEMV of north_star)
//   decode_attn_gqa_kernel     G = 2, 4, 8: register-tiled packed FFMA2 over the same ring (FMA-bound on CUDA cores at G = 8)
//   decode_attn_split_kernel   register-staged fp32 fallback for code rows that are not back to back
//   decode_attn_generic_kernel any head_dim 1..512, any G, any alignment (coverage, not bandwidth)
//   decode_attn_merge_kernel   merges the splits' partials (m, l, acc[D]) and the step's own, exact token
//
// Grid (n_splits, H, B); n_splits is chosen for wave efficiency (choose_splits).  The register-staged fallback, per tile of
// 128 tokens:
//   K phase  : a lane owns 16 consecutive d of one token (16 B of INT8 / 8 B of INT4); partial dot
//              products are reduced over the D/16 lanes of the row by shuffles.
//   softmax  : per query head, tile max / running max / exp2 / running sum in one warp.
//   V phase  : a lane owns a fixed slice of 32/G consecutive d for every token of its row group and
//              accumulates G x (32/G) fp32 partial outputs; row groups are summed through shared
//              memory once per split.
#include <cmath>

#include "attn_common.cuh"

#ifndef KVQ_ATTN_IMMA_MIN_G
#define KVQ_ATTN_IMMA_MIN_G 2   // KVQ_ATTN_AUTO: smallest Hq / H that takes the tensor-core variant
#endif

namespace kvq {
namespace {

constexpr int kThreads = 128;

// 4 codes of one 32-bit word -> 4 floats, exact.  INT8: flip the sign bit so the byte is code+128,
// drop it into the mantissa of 2^23 (PRMT), subtract 2^23+128.
__device__ __forceinline__ void int8x4_to_float(uint32_t w, float f[4]) {
    const uint32_t x = w ^ 0x80808080u;
    f[0] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650)) - 8388736.f;
    f[1] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7651)) - 8388736.f;
    f[2] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7652)) - 8388736.f;
    f[3] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7653)) - 8388736.f;
}
// 8 nibbles of one 32-bit word (4 bytes; byte j = elements 2j (high nibble), 2j+1 (low)) -> 8 floats = code
__device__ __forceinline__ void int4x8_to_float(uint32_t w, float f[8]) {
    const uint32_t lo = w & 0x0f0f0f0fu, hi = (w >> 4) & 0x0f0f0f0fu;
    f[0] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7650)) - 8388616.f;
    f[1] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7650)) - 8388616.f;
    f[2] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7651)) - 8388616.f;
    f[3] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7651)) - 8388616.f;
    f[4] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7652)) - 8388616.f;
    f[5] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7652)) - 8388616.f;
    f[6] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7653)) - 8388616.f;
    f[7] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7653)) - 8388616.f;
}

// N consecutive codes starting at a byte pointer, as floats.  N*BITS/8 bytes, naturally aligned.
template <int BITS, int N> struct CodeVec {
    static constexpr int kBytes = N * BITS / 8;
    static constexpr int kWords = kBytes >= 4 ? kBytes / 4 : 1;
    uint32_t w[kWords];
    __device__ __forceinline__ void load(const uint8_t* p) {
        if (kBytes == 2) w[0] = *reinterpret_cast<const unsigned short*>(p);
        else if (kBytes == 4) w[0] = *reinterpret_cast<const uint32_t*>(p);
        else if (kBytes == 8) { const uint2 t = ld_stream8(p); w[0] = t.x; w[1] = t.y; }
        else {
#pragma unroll
            for (int i = 0; i < kBytes / 16; ++i) {
                const uint4 t = ld_stream(p + 16 * i);
                w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
            }
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < kWords; ++i) w[i] = BITS == 8 ? 0u : 0x88888888u;
    }
    __device__ __forceinline__ void to_float(float f[N]) const {
        if (BITS == 8) {
            if (N >= 4) {
#pragma unroll
                for (int i = 0; i < N / 4; ++i) int8x4_to_float(w[i], f + 4 * i);
            }
        } else {
            if (N >= 8) {
#pragma unroll
                for (int i = 0; i < N / 8; ++i) int4x8_to_float(w[i], f + 8 * i);
            } else {  // N == 4: two bytes
                float t[8];
                int4x8_to_float(w[0], t);
#pragma unroll
                for (int i = 0; i < N; ++i) f[i] = t[i];
            }
        }
    }
};

template <int KBITS, int VBITS, int G, int D>
__global__ void __launch_bounds__(kThreads) decode_attn_split_kernel(AttnParams P) {
    constexpr int LPR = D / 16;            // K: lanes per row
    constexpr int KROWS = kThreads / LPR;  // K rows per pass
    constexpr int KPASS = kTile / KROWS;
    constexpr int EV = 32 / G;             // V: elements per lane
    constexpr int LPRV = D / EV;
    constexpr int VROWS = kThreads / LPRV;
    constexpr int VPASS = kTile / VROWS;
    constexpr bool kQRegs = G <= 2;
    static_assert(G == 1 || G == 2 || G == 4 || G == 8, "G");
    static_assert(D == 64 || D == 128, "D");

    __shared__ float q_s[G][D];
    __shared__ float p_s[G][kTile];
    __shared__ float alpha_s[G], m_s[G], l_s[G];
    __shared__ float red_s[kThreads * 32];

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = split * P.tokens_per_split;
    const int t_end = min(P.T, t_begin + P.tokens_per_split);

    // q (G heads of this kv head), pre-scaled into the log2 domain
    for (int i = tid; i < G * D; i += kThreads) {
        const int g = i / D, d = i - g * D;
        const int64_t qi = ((int64_t)b * P.Hq + (int64_t)h * G + g) * D + d;
        q_s[g][d] = load_scale(P.q, P.q_dtype, qi) * P.qk_scale_log2;
    }
    if (tid < G) { m_s[tid] = -INFINITY; l_s[tid] = 0.f; alpha_s[tid] = 0.f; }
    __syncthreads();

    const int kc = tid % LPR, kr = tid / LPR;
    const int vc = tid % LPRV, vr = tid / LPRV;
    float qr[kQRegs ? G : 1][16];
    if (kQRegs) {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int e = 0; e < 16; ++e) qr[g][e] = q_s[g][kc * 16 + e];
    }
    float acc[G][EV];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int e = 0; e < EV; ++e) acc[g][e] = 0.f;

    const uint8_t* kbase = P.k + (int64_t)b * P.kB + (int64_t)h * P.kH + (int64_t)kc * (16 * KBITS / 8);
    const uint8_t* vbase = P.v + (int64_t)b * P.vB + (int64_t)h * P.vH + (int64_t)vc * (EV * VBITS / 8);

    for (int t0 = t_begin; t0 < t_end; t0 += kTile) {
        // ---- issue all loads of the tile (K then V) before any math
        CodeVec<KBITS, 16> kv[KPASS];
        float ks[KPASS];
#pragma unroll
        for (int ps = 0; ps < KPASS; ++ps) {
            const int t = t0 + ps * KROWS + kr;
            if (t < t_end) { kv[ps].load(kbase + (int64_t)t * P.kT); ks[ps] = P.k_scale[t]; }
            else { kv[ps].zero(); ks[ps] = 0.f; }
        }
        CodeVec<VBITS, EV> vv[VPASS];
#pragma unroll
        for (int ps = 0; ps < VPASS; ++ps) {
            const int t = t0 + ps * VROWS + vr;
            if (t < t_end) vv[ps].load(vbase + (int64_t)t * P.vT); else vv[ps].zero();
        }
        // ---- K phase: scores (log2 domain) into p_s
#pragma unroll
        for (int ps = 0; ps < KPASS; ++ps) {
            float f[16];
            kv[ps].to_float(f);
            const int row = ps * KROWS + kr;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float d = 0.f;
                if (kQRegs) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) d = fmaf(qr[g][e], f[e], d);
                } else {
                    const float4* qp = reinterpret_cast<const float4*>(&q_s[g][kc * 16]);
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 qq = qp[e4];
                        d = fmaf(qq.x, f[4 * e4], d); d = fmaf(qq.y, f[4 * e4 + 1], d);
                        d = fmaf(qq.z, f[4 * e4 + 2], d); d = fmaf(qq.w, f[4 * e4 + 3], d);
                    }
                }
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (kc == 0) p_s[g][row] = (t0 + row < t_end) ? d * ks[ps] : -INFINITY;
            }
        }
        __syncthreads();
        // ---- softmax bookkeeping: warp w owns heads w, w+4
        for (int g = warp; g < G; g += kThreads / 32) {
            float s[kTile / 32], mt = -INFINITY;
#pragma unroll
            for (int i = 0; i < kTile / 32; ++i) { s[i] = p_s[g][lane + 32 * i]; mt = fmaxf(mt, s[i]); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
            const float m_old = m_s[g], m_new = fmaxf(m_old, mt);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < kTile / 32; ++i) {
                const int t = t0 + lane + 32 * i;
                const float p = (t < t_end) ? exp2f(s[i] - m_new) : 0.f;
                sum += p;
                p_s[g][lane + 32 * i] = (t < t_end) ? p * P.v_scale[t] : 0.f;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncwarp();
            if (lane == 0) {
                const float a = (m_old == -INFINITY) ? 0.f : exp2f(m_old - m_new);
                alpha_s[g] = a; m_s[g] = m_new; l_s[g] = l_s[g] * a + sum;
            }
        }
        __syncthreads();
        // ---- V phase
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float a = alpha_s[g];
#pragma unroll
            for (int e = 0; e < EV; ++e) acc[g][e] *= a;
        }
#pragma unroll
        for (int ps = 0; ps < VPASS; ++ps) {
            float f[EV];
            vv[ps].to_float(f);
            const int row = ps * VROWS + vr;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float w = p_s[g][row];
#pragma unroll
                for (int e = 0; e < EV; ++e) acc[g][e] = fmaf(w, f[e], acc[g][e]);
            }
        }
        __syncthreads();
    }

    // ---- sum the VROWS row groups through shared memory; write the split's partial
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int e = 0; e < EV; ++e) red_s[(vr * G + g) * D + vc * EV + e] = acc[g][e];
    __syncthreads();
    for (int o = tid; o < G * D; o += kThreads) {
        float s = 0.f;
#pragma unroll 4
        for (int r = 0; r < VROWS; ++r) s += red_s[r * G * D + o];
        const int g = o / D, d = o - g * D;
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + g;
        P.ws_acc[(hq * P.n_splits + split) * D + d] = s;
    }
    if (tid < G) {
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + tid;
        P.ws_ml[(hq * P.n_splits + split) * 2] = m_s[tid];
        P.ws_ml[(hq * P.n_splits + split) * 2 + 1] = l_s[tid];
    }
}

// =====================================================================================================
// TMA-pipelined variant (used whenever the code rows of a (b, h) plane are contiguous, i.e. always for the
// cache store of this package).  The split-K loop is latency-bound when every tile's loads are issued only
// after the previous tile's math: here one elected thread streams tiles of 128 tokens -- K codes, V codes and
// both scale vectors -- into a ring of shared-memory stages with cp.async.bulk (the TMA bulk-copy engine,
// SASS UBLKCP) completing on an mbarrier per stage, 2-3 tiles ahead of the math.  All 256 threads are
// consumers; the softmax bookkeeping is folded into the two math phases (per-warp maxima during the K phase,
// exp2 on the fly in the V phase), so a tile costs two block barriers.
// =====================================================================================================

#ifndef KVQ_ATTN_RING_KB
#define KVQ_ATTN_RING_KB 100
#endif
#ifndef KVQ_ATTN_CTAS_PER_SM
#define KVQ_ATTN_CTAS_PER_SM 2
#endif
constexpr int kThreads2 = 256;

// one thread: arm the stage's mbarrier and start the bulk copies of a tile (rows tokens from t0) of plane (b, h).  Contiguous
// store: one copy per tensor.  Paged store: one copy per tensor per block piece, source address through the block table.
template <int KROW, int VROW, int KTILE, bool PAGED>
__device__ __forceinline__ void issue_tile(const AttnParams& P, int b, int h, int t0, int rows, uint8_t* sb, uint64_t* bar) {
    mbar_expect_tx(bar, (uint32_t)rows * (KROW + VROW));
    const uint8_t* kplane = P.k + (int64_t)b * P.kB + (int64_t)h * P.kH;
    const uint8_t* vplane = P.v + (int64_t)b * P.vB + (int64_t)h * P.vH;
    if constexpr (!PAGED) {
        bulk_g2s(sb, kplane + (int64_t)t0 * KROW, (uint32_t)rows * KROW, bar);
        bulk_g2s(sb + KTILE, vplane + (int64_t)t0 * VROW, (uint32_t)rows * VROW, bar);
        return;
    }
    // t0 is a multiple of the 128-token tile and block_size a power of two: a tile is one piece of a large block, or
    // 128 / block_size whole blocks.  All table entries are loaded before the first copy is issued (one latency, not eight),
    // and the entries of the next tile are pulled into L1 for the next call.
    const int blk0 = t0 >> P.bs_shift;
    if (P.bs >= kTile) {
        const int64_t phys = __ldg(P.table + blk0);
        const int w = t0 & (P.bs - 1);
        bulk_g2s(sb, kplane + phys * P.kBlk + (int64_t)w * KROW, (uint32_t)rows * KROW, bar);
        bulk_g2s(sb + KTILE, vplane + phys * P.vBlk + (int64_t)w * VROW, (uint32_t)rows * VROW, bar);
    } else {
        const int np = (rows + P.bs - 1) >> P.bs_shift;
        int phys[kTile / 16];
#pragma unroll
        for (int i = 0; i < kTile / 16; ++i) phys[i] = i < np ? __ldg(P.table + blk0 + i) : 0;
#pragma unroll
        for (int i = 0; i < kTile / 16; ++i) {
            if (i < np) {
                const int r0 = i << P.bs_shift, len = min(P.bs, rows - r0);
                bulk_g2s(sb + r0 * KROW, kplane + (int64_t)phys[i] * P.kBlk, (uint32_t)len * KROW, bar);
                bulk_g2s(sb + KTILE + r0 * VROW, vplane + (int64_t)phys[i] * P.vBlk, (uint32_t)len * VROW, bar);
            }
        }
    }
    if (t0 + kTile < P.T) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.table + ((t0 + kTile) >> P.bs_shift)));
}

// =====================================================================================================
// Register-tiled variant for GQA groups (G = 4, 8: Llama-3-70B's 64 query heads on 8 KV heads is the C5 shape).
// Every K/V element now feeds G query heads, so per KV head the step is a small [T x D] . [D x G] product
// followed by [G x T] . [T x D]: on CUDA cores that is FFMA-bound, not HBM-bound (G FFMA per code against one
// byte or nibble of traffic), and what matters is instructions per FFMA and keeping the FMA pipe fed.  The
// G <= 2 kernels above spend ~10 instructions per (code, head); this one is laid out like a register-blocked
// SGEMM, and every WARP runs its own online softmax over its own 16 tokens of each 128-token tile, so the main
// loop has no block-wide barrier at all (the first tiled version had three per tile: 27 % of its stall samples):
//   K phase : a lane owns 16 channels of RT tokens and accumulates a [RT tokens x G heads] block; a code is
//             converted to fp32 once and used G times, a float4 of q (LDS.128, conflict-free padded layout)
//             is used for RT tokens.  The D/16 lanes of a row are combined by a reduce-SCATTER: lane kc keeps
//             its accumulators in the order head = j ^ swz(kc), so every halving step exchanges statically
//             indexed registers (no selects) and lane kc ends with the full score of head kc.
//   softmax : in the registers that hold the scores: warp max per head by shuffles, exp2 once per (token, head)
//             -- the G <= 2 kernels recompute it in every V lane --, weights p * v_scale[t] to the warp's own
//             [16][G] patch of shared memory (__syncwarp only).
//   V phase : a lane owns 4 channels, the warp covers D; [G heads x 4 channels] accumulators per lane, weights
//             read back as LDS.128 broadcasts.
// Ring: cp.async.bulk + a `full` mbarrier per stage as in the other variants, plus an `empty` mbarrier per
// stage (one arrival per warp) so that the producer thread only ever waits for the tile before last.
// The eight per-warp partial results are merged through shared memory once per CTA; split / merge protocol
// with the second kernel unchanged.
// =====================================================================================================

#ifndef KVQ_ATTN_GQA_STAGES
#define KVQ_ATTN_GQA_STAGES 3
#endif
#ifndef KVQ_ATTN_GQA_CTAS_PER_SM
#define KVQ_ATTN_GQA_CTAS_PER_SM 2
#endif
#ifndef KVQ_ATTN_GQA_RT
#define KVQ_ATTN_GQA_RT 4
#endif
#ifndef KVQ_ATTN_GQA_VUNROLL
#define KVQ_ATTN_GQA_VUNROLL 8
#endif

constexpr float kLazyLog2 = 8.f;  // see the softmax step of decode_attn_gqa_kernel

// packed fp32 pairs: Blackwell's FFMA2 / FADD2 / FMUL2 (PTX fma.rn.f32x2 ...) do two IEEE fp32 operations per issue slot;
// a pair lives in a 64-bit register, and a pair whose halves are the same register becomes SASS's scalar-broadcast operand

// lane bit o of the K-phase reduce-scatter -> head bit it halves away (bits >= 2 only; see the kernel)
__host__ __device__ constexpr int gqa_swz(int kc, int G, int LPR) {
    int s = 0, hb = G / 2;
    for (int o = LPR / 2; o >= 1; o >>= 1) {
        if (hb >= 2) { if (kc & o) s |= hb; hb >>= 1; }
    }
    return s;
}

template <int KBITS, int VBITS, int G, int D> struct GqaCfg {
    static constexpr int NT = kThreads2, kWarps = NT / 32, WR = kTile / kWarps;  // WR = 16 tokens per warp per tile
    static constexpr int kKRow = D * KBITS / 8, kVRow = D * VBITS / 8;
    static constexpr int kKTile = kTile * kKRow, kVTile = kTile * kVRow, kStage = kKTile + kVTile;
    static constexpr int kStages = KVQ_ATTN_GQA_STAGES;
    static constexpr int LPR = D / 16;                          // K: lanes per token row, 16 channels each
    // q in shared memory: [variant 2][chunk LPR][channel 16][G heads], padded so that the LPR lanes of a row, which read
    // different chunks / variants / head groups in one LDS.128, hit 32 distinct banks (pads found by exhaustive search)
    static constexpr int CP = 16 * G + (G == 8 && LPR == 8 ? 8 : 4);  // G = 2: 36, chunk kc starts at bank 4 kc (LDS.128 of two channels)
    static constexpr int VS = LPR * CP + (G == 8 && LPR == 4 ? 4 : 0);
    static constexpr int LPRV = D / 4, VRW = 32 / LPRV;         // V: lanes per token row (4 channels each), rows per warp pass
    static constexpr int kRed = kWarps * VRW * G * D * 4;       // final merge of the warps' partials aliases the ring
    static constexpr int kBuf = kStages * kStage > kRed ? kStages * kStage : kRed;
    // q_s, per-warp weight patches, per-warp alpha / m / l, barriers
    static constexpr int kSmem = kBuf + 2 * VS * 4 + kWarps * WR * G * 4 + 3 * kWarps * G * 4 + 2 * kStages * 8 + 16;
};

// 4 consecutive codes -> 4 floats (two PRMT-built pairs, one FADD2 each).  INT8: word e4 of the lane's words; INT4: half
// (e4 & 1) of word e4 >> 1
template <int BITS> __device__ __forceinline__ void cvt4(const uint32_t* w, int e4, float f[4]) {
    if constexpr (BITS == 8) {
        const uint32_t x = w[e4] ^ 0x80808080u;
        const f32x2 c = pack2(-8388736.f, -8388736.f);
        const f32x2 lo = fadd2(pack2(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650)), __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7651))), c);
        const f32x2 hi = fadd2(pack2(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7652)), __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7653))), c);
        unpack2(lo, f[0], f[1]);
        unpack2(hi, f[2], f[3]);
    } else {
        const uint32_t x = w[e4 >> 1] >> (16 * (e4 & 1));
        const uint32_t hn = (x >> 4) & 0x0f0fu, ln = x & 0x0f0fu;
        const f32x2 c = pack2(-8388616.f, -8388616.f);
        const f32x2 lo = fadd2(pack2(__uint_as_float(__byte_perm(hn, 0x4B000000u, 0x7650)), __uint_as_float(__byte_perm(ln, 0x4B000000u, 0x7650))), c);
        const f32x2 hi = fadd2(pack2(__uint_as_float(__byte_perm(hn, 0x4B000000u, 0x7651)), __uint_as_float(__byte_perm(ln, 0x4B000000u, 0x7651))), c);
        unpack2(lo, f[0], f[1]);
        unpack2(hi, f[2], f[3]);
    }
}

template <int KBITS, int VBITS, int G, int D, bool PAGED>
__global__ void __launch_bounds__(kThreads2, KVQ_ATTN_GQA_CTAS_PER_SM) decode_attn_gqa_kernel(AttnParams P) {
    using C = GqaCfg<KBITS, VBITS, G, D>;
    constexpr int NT = C::NT, kWarps = C::kWarps, WR = C::WR;
    constexpr int LPR = C::LPR, RPP = 32 / LPR, RTT = WR / RPP;  // K: rows per warp pass, tokens per lane per tile
    constexpr int RT = RTT < KVQ_ATTN_GQA_RT ? RTT : KVQ_ATTN_GQA_RT, KSUB = RTT / RT;  // register tile: RT tokens x G heads
    constexpr int KW = 16 * KBITS / 32;                         // code words per lane per token
    constexpr int LPRV = C::LPRV, VRW = C::VRW, VP = WR / VRW;
    constexpr int NREM = G > LPR ? G / LPR : 1;                 // scores a lane holds per token after the reduce-scatter
    constexpr int S = C::kStages, PD = S > 2 ? S - 2 : 1;       // tiles in flight ahead of the math
    constexpr int GP = G / 2;                                   // head pairs
    constexpr int kVUnroll = KVQ_ATTN_GQA_VUNROLL;
    static_assert(G == 2 || G == 4 || G == 8, "G");
    static_assert(RT >= 1 && RT * KSUB * RPP == WR, "tiling");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* q_s = reinterpret_cast<float*>(smem_raw + C::kBuf);  // [2][LPR chunks][16][G] (padded): variant v holds head pos ^ 2v
    float* p_s = q_s + 2 * C::VS;                               // [kWarps][WR][G] weights p * v_scale
    float* alpha_s = p_s + kWarps * WR * G;                     // [kWarps][G]
    float* m_s = alpha_s + kWarps * G;                          // [kWarps][G]  (end of kernel)
    float* l_s = m_s + kWarps * G;                              // [kWarps][G]  (end of kernel)
    uint64_t* full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(l_s + kWarps * G) + 7) & ~uintptr_t(7));
    uint64_t* empty = full + S;

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = split * P.tokens_per_split;
    const int t_end = min(P.T, t_begin + P.tokens_per_split);
    const int n_tiles = (t_end - t_begin + kTile - 1) / kTile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * G * D; i += NT) {
        const int v = i / (G * D), r = i - v * G * D, g = r / D, d = r - g * D;
        q_s[v * C::VS + (d >> 4) * C::CP + (d & 15) * G + (g ^ (2 * v))] =
            load_scale(P.q, P.q_dtype, ((int64_t)b * P.Hq + (int64_t)h * G + g) * D + d) * P.qk_scale_log2;
    }
    __syncthreads();

    auto issue = [&](int tile_idx) {
        const int s = tile_idx % S;
        const int t0 = t_begin + tile_idx * kTile;
        issue_tile<C::kKRow, C::kVRow, C::kKTile, PAGED>(P, b, h, t0, min(kTile, t_end - t0), smem_raw + s * C::kStage, &full[s]);
    };
    if (tid == 0) {
        for (int i = 0; i < PD && i < n_tiles; ++i) issue(i);
    }

    // K-phase lane roles.  Register j of lane kc holds head j ^ swz: swz maps every lane bit whose reduce-scatter step halves away a
    // head bit >= 2 to that head bit, so those steps exchange statically indexed register PAIRS; head bit 1 is scattered with a select.
    const int kc = lane % LPR, kr = lane / LPR;
    const int swz = gqa_swz(kc, G, LPR);
    int sel_bit = 0, bfly_mask = 0;
    {
        int hb = G / 2;
#pragma unroll
        for (int o = LPR / 2; o >= 1; o >>= 1) {
            if (hb >= 2) hb >>= 1;
            else if (hb == 1) { sel_bit = o; hb = 0; }
            else bfly_mask |= o;
        }
    }
    const bool k_writer = (kc & bfly_mask) == 0;
    const int head0 = swz | ((sel_bit && (kc & sel_bit)) ? 1 : 0);  // head of score 0 (score 1, if NREM == 2, is head0 + 1)
    const int vc = lane % LPRV, vrow = lane / LPRV;
    const float* q_lane = q_s + ((swz >> 1) & 1) * C::VS + kc * C::CP;
    const int mm = swz >> 2;
    float* p_w = p_s + warp * WR * G;
    float* alpha_w = alpha_s + warp * G;

    f32x2 acc[GP][4];  // (head 2gp, head 2gp + 1) x 4 channels
    float m_run[NREM], l_run[NREM];
#pragma unroll
    for (int g = 0; g < GP; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[g][e] = 0ull;
#pragma unroll
    for (int j = 0; j < NREM; ++j) { m_run[j] = -INFINITY; l_run[j] = 0.f; }

    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % S;
        if (tid == 0 && it + PD < n_tiles) {
            const int nt = it + PD;  // goes into the stage tile nt - S used: every warp must have left it
            if (nt >= S) { while (!mbar_try_wait(&empty[nt % S], ((nt / S) - 1) & 1)) {} }
            issue(nt);
        }
        const int t0 = t_begin + it * kTile;
        const int rows = min(kTile, t_end - t0);
        // per-token scales of this lane's RT tokens (L1/L2-resident plain loads, consumed after the K phase)
        float ksr[RTT], vsr[RTT];
#pragma unroll
        for (int r = 0; r < RTT; ++r) {
            const int row = warp * WR + r * RPP + kr;
            const int64_t ti = row < rows ? tok_index<PAGED>(P, t0 + row) : 0;
            ksr[r] = row < rows ? __ldg(P.k_scale + ti) : 0.f;
            vsr[r] = row < rows ? __ldg(P.v_scale + ti) : 0.f;
        }
        while (!mbar_try_wait(&full[s], (it / S) & 1)) {}
        const uint8_t* sb = smem_raw + s * C::kStage;

        // ---- K phase: [RT tokens x G heads] register blocks, two heads per FFMA2
        float sc[RTT][NREM];
#pragma unroll
        for (int sub = 0; sub < KSUB; ++sub) {
            uint32_t kw[RT][KW];
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int row = warp * WR + (sub * RT + r) * RPP + kr;
                const uint8_t* src = sb + row * C::kKRow + kc * (2 * KBITS);
                if constexpr (KW == 4) {
                    const uint4 t = *reinterpret_cast<const uint4*>(src);
                    kw[r][0] = t.x; kw[r][1] = t.y; kw[r][2] = t.z; kw[r][3] = t.w;
                } else {
                    const uint2 t = *reinterpret_cast<const uint2*>(src);
                    kw[r][0] = t.x; kw[r][1] = t.y;
                }
            }
            f32x2 a[RT][GP];
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < GP; ++j) a[r][j] = 0ull;
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                float f[RT][4];
#pragma unroll
                for (int r = 0; r < RT; ++r) cvt4<KBITS>(kw[r], e4, f[r]);
                if constexpr (G == 2) {  // one head pair: an LDS.128 of q covers two channels
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const ulonglong2 qq = *reinterpret_cast<const ulonglong2*>(q_lane + (4 * e4 + 2 * e2) * G);
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            a[r][0] = ffma2(qq.x, pack2(f[r][2 * e2], f[r][2 * e2]), a[r][0]);
                            a[r][0] = ffma2(qq.y, pack2(f[r][2 * e2 + 1], f[r][2 * e2 + 1]), a[r][0]);
                        }
                    }
                } else {
#pragma unroll
                    for (int ee = 0; ee < 4; ++ee) {
#pragma unroll
                        for (int m = 0; m < G / 4; ++m) {
                            const ulonglong2 qq = *reinterpret_cast<const ulonglong2*>(q_lane + (4 * e4 + ee) * G + 4 * (m ^ mm));
#pragma unroll
                            for (int r = 0; r < RT; ++r) {
                                const f32x2 fb = pack2(f[r][ee], f[r][ee]);
                                a[r][2 * m] = ffma2(qq.x, fb, a[r][2 * m]);
                                a[r][2 * m + 1] = ffma2(qq.y, fb, a[r][2 * m + 1]);
                            }
                        }
                    }
                }
            }
            // reduce-scatter over the LPR lanes of a row
            int hb = G / 2;
            bool scalar = false;
#pragma unroll
            for (int o = LPR / 2; o >= 1; o >>= 1) {
                if (hb >= 2) {         // partner's upper half of the pairs lands on my lower half
                    hb >>= 1;
#pragma unroll
                    for (int r = 0; r < RT; ++r)
#pragma unroll
                        for (int j = 0; j < GP / 2; ++j)
                            if (j < hb) a[r][j] = fadd2(a[r][j], __shfl_xor_sync(0xffffffffu, a[r][j + hb], o));
                } else if (hb == 1) {  // last pair: keep one head, send the other
                    hb = 0; scalar = true;
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        float x, y;
                        unpack2(a[r][0], x, y);
                        const bool up = (kc & o) != 0;
                        sc[sub * RT + r][0] = (up ? y : x) + __shfl_xor_sync(0xffffffffu, up ? x : y, o);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RT; ++r) sc[sub * RT + r][0] += __shfl_xor_sync(0xffffffffu, sc[sub * RT + r][0], o);
                }
            }
            if (!scalar) {  // G > LPR: a pair of heads per lane
#pragma unroll
                for (int r = 0; r < RT; ++r) unpack2(a[r][0], sc[sub * RT + r][0], sc[sub * RT + r][NREM - 1]);
            }
        }
        // ---- softmax in registers: lane (kc, kr) holds the scores of head0 (+ j) for its RT tokens
        bool moved = false;
        float alj[NREM];
#pragma unroll
        for (int j = 0; j < NREM; ++j) {
            float mt = -INFINITY;
#pragma unroll
            for (int r = 0; r < RTT; ++r) {
                const int row = warp * WR + r * RPP + kr;
                sc[r][j] = row < rows ? sc[r][j] * ksr[r] : -INFINITY;
                mt = fmaxf(mt, sc[r][j]);
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
            // lazy reference: the running reference only moves when the tile maximum exceeds it by more than 2^kLazyLog2 (weights
            // stay <= 2^kLazyLog2, harmless in fp32; the result does not depend on the reference), so the accumulators are
            // rescaled a handful of times per split instead of after ~70 % of the tiles
            const float m_old = m_run[j];
            const bool up = mt > m_old + kLazyLog2;  // false when the warp's rows are all masked (mt == -inf)
            const float m_new = up ? mt : m_old;
            const float al = up ? (m_old == -INFINITY ? 0.f : exp2f(m_old - m_new)) : 1.f;
            moved |= up;
            m_run[j] = m_new;
            float lsum = 0.f;
#pragma unroll
            for (int r = 0; r < RTT; ++r) {
                const float pw = sc[r][j] == -INFINITY ? 0.f : exp2f(sc[r][j] - m_new);
                lsum += pw;
                if (k_writer) p_w[(r * RPP + kr) * G + head0 + j] = pw * vsr[r];
            }
            l_run[j] = fmaf(l_run[j], al, lsum);
            alj[j] = al;
        }
        moved = __any_sync(0xffffffffu, moved);
        if (moved && kr == 0 && k_writer) {
#pragma unroll
            for (int j = 0; j < NREM; ++j) alpha_w[head0 + j] = alj[j];
        }
        __syncwarp();

        // ---- V phase: [G heads x 4 channels] accumulators, two heads per FFMA2
        if (moved) {  // warp-uniform: the maxima settle after the first tiles
            if constexpr (G == 2) {
                const f32x2 al = *reinterpret_cast<const f32x2*>(alpha_w);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[0][e] = fmul2(acc[0][e], al);
            }
#pragma unroll
            for (int g4 = 0; g4 < G / 4; ++g4) {
                const ulonglong2 al = *reinterpret_cast<const ulonglong2*>(alpha_w + 4 * g4);
#pragma unroll
                for (int e = 0; e < 4; ++e) { acc[2 * g4][e] = fmul2(acc[2 * g4][e], al.x); acc[2 * g4 + 1][e] = fmul2(acc[2 * g4 + 1][e], al.y); }
            }
        }
        const uint8_t* vsb = sb + C::kKTile + (warp * WR + vrow) * C::kVRow + vc * (4 * VBITS / 8);
#pragma unroll kVUnroll
        for (int ps = 0; ps < VP; ++ps) {
            uint32_t vw[1];
            if constexpr (VBITS == 8) vw[0] = *reinterpret_cast<const uint32_t*>(vsb + ps * VRW * C::kVRow);
            else vw[0] = *reinterpret_cast<const unsigned short*>(vsb + ps * VRW * C::kVRow);
            float f[4];
            cvt4<VBITS>(vw, 0, f);
            if constexpr (G == 2) {
                const f32x2 pp = *reinterpret_cast<const f32x2*>(p_w + (ps * VRW + vrow) * G);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[0][e] = ffma2(pp, pack2(f[e], f[e]), acc[0][e]);
            }
            const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(p_w + (ps * VRW + vrow) * G);
#pragma unroll
            for (int g4 = 0; g4 < G / 4; ++g4) {
                const ulonglong2 pp = pr[g4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const f32x2 fb = pack2(f[e], f[e]);
                    acc[2 * g4][e] = ffma2(pp.x, fb, acc[2 * g4][e]);
                    acc[2 * g4 + 1][e] = ffma2(pp.y, fb, acc[2 * g4 + 1][e]);
                }
            }
        }
        __syncwarp();                           // p_w / alpha_w are rewritten by the next tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the stage before its async-proxy (bulk copy) refill
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
    }

    // ---- merge the warps' partials (own running max each) through shared memory; the ring is idle by now
    float* red_s = reinterpret_cast<float*>(smem_raw);
#pragma unroll
    for (int j = 0; j < NREM; ++j) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) l_run[j] += __shfl_xor_sync(0xffffffffu, l_run[j], o);
        if (kr == 0 && k_writer) { m_s[warp * G + head0 + j] = m_run[j]; l_s[warp * G + head0 + j] = l_run[j]; }
    }
    __syncthreads();  // also: every warp has finished reading the ring
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) {
        float fac[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int g = 2 * gp + u;
            float M = -INFINITY;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) M = fmaxf(M, m_s[w * G + g]);
            const float mw = m_s[warp * G + g];
            fac[u] = mw == -INFINITY ? 0.f : exp2f(mw - M);
        }
        float lo[4], hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) unpack2(acc[gp][e], lo[e], hi[e]);
        float* dst = red_s + ((warp * VRW + vrow) * G + 2 * gp) * D + vc * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(lo[0] * fac[0], lo[1] * fac[0], lo[2] * fac[0], lo[3] * fac[0]);
        *reinterpret_cast<float4*>(dst + D) = make_float4(hi[0] * fac[1], hi[1] * fac[1], hi[2] * fac[1], hi[3] * fac[1]);
    }
    __syncthreads();
    for (int o = tid; o < G * D; o += NT) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < kWarps * VRW; ++r) sum += red_s[r * G * D + o];
        const int g = o / D, d = o - g * D;
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + g;
        P.ws_acc[(hq * P.n_splits + split) * D + d] = sum;
    }
    if (tid < G) {
        float M = -INFINITY, L = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) M = fmaxf(M, m_s[w * G + tid]);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const float mw = m_s[w * G + tid];
            L += mw == -INFINITY ? 0.f : l_s[w * G + tid] * exp2f(mw - M);
        }
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + tid;
        P.ws_ml[(hq * P.n_splits + split) * 2] = M;
        P.ws_ml[(hq * P.n_splits + split) * 2 + 1] = L;
    }
}

// =====================================================================================================
// Per-warp integer-dot variant for G == 1 (replaces decode_attn_dp4a_kernel as the default MHA-shape kernel).
// Same arithmetic as the block-wide DP4A kernel above -- q as three signed byte planes, weights p * v_scale[t] as
// three unsigned byte planes, int32 accumulators flushed when the softmax reference moves -- restructured like the
// GQA kernel: every warp runs its own online softmax over its own 16 tokens of each 128-token tile, so the main loop
// has no block-wide barrier (the block-wide kernel: 2 per tile, no eligible warp in 50 % of the cycles), and
//   * the D/16 partial dot products of a row are combined by a reduce-scatter over the FOUR rows a lane group works on
//     (register j of lane kc holds pass j ^ (kc >> 1): 4 SHFL + 4 FADD instead of 12 + 12);
//   * each token's score ends in two lanes: one k_scale / v_scale load per lane per tile, warp max by REDUX on the
//     order-preserving integer image of the float;
//   * lazy reference (moves only when exceeded by 2^kMhaLazyLog2; weights are scaled so that 2^kMhaLazyLog2 * max(v_scale)
//     still fits the 24-bit fixed point: 21 bits left for weights <= 1), so the int32 -> fp32 flush is rare.
// =====================================================================================================

#ifndef KVQ_ATTN_MHA_STAGES
#define KVQ_ATTN_MHA_STAGES 3
#endif
#ifndef KVQ_ATTN_MHA_CTAS_PER_SM
#define KVQ_ATTN_MHA_CTAS_PER_SM 2
#endif
constexpr float kMhaLazyLog2 = 3.f;

template <int KBITS, int VBITS, int D> struct MhaCfg {
    static constexpr int NT = kThreads2, kWarps = NT / 32, WR = kTile / kWarps;  // 16 tokens per warp per tile
    static constexpr int kKRow = D * KBITS / 8, kVRow = D * VBITS / 8;
    static constexpr int kKTile = kTile * kKRow, kVTile = kTile * kVRow, kStage = kKTile + kVTile;
    static constexpr int kStages = KVQ_ATTN_MHA_STAGES;
    static constexpr int WPR = D / 8, HW = 32 / WPR;            // V: lanes per token row (8 channels each); row groups per warp
    static constexpr int kRed = kWarps * HW * D * 4;            // final merge of the partials aliases the ring
    static constexpr int kBuf = kStages * kStage > kRed ? kStages * kStage : kRed;
    // weight digit patches [kWarps][3][16], q digits [3][D], per-warp m / l, block scratch, barriers
    static constexpr int kSmem = kBuf + kWarps * 48 + 3 * D + 2 * kWarps * 4 + 2 * kWarps * 4 + 2 * kStages * 8 + 32;
};

template <int KBITS, int VBITS, int D, bool PAGED>
__global__ void __launch_bounds__(kThreads2, KVQ_ATTN_MHA_CTAS_PER_SM) decode_attn_mha_kernel(AttnParams P) {
    using C = MhaCfg<KBITS, VBITS, D>;
    constexpr int NT = C::NT, kWarps = C::kWarps, WR = C::WR;
    constexpr int LPR = D / 16, RPP = 32 / LPR, RTT = WR / RPP;  // K: lanes per row, rows per warp pass, passes (4 / 2)
    constexpr int WPR = C::WPR, HW = C::HW, GPH = 4 / HW;       // V: groups of four tokens per row group (2 / 1)
    constexpr int VB = 8 * VBITS / 8;                           // bytes a lane reads per V row: 8 (INT8) / 4 (INT4)
    constexpr int S = C::kStages, PD = S > 2 ? S - 2 : 1;
    static_assert(RTT == 4 || RTT == 2, "passes");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* wq_s = smem_raw + C::kBuf;                                  // [kWarps][3][16] weight digits
    int8_t* qd_s = reinterpret_cast<int8_t*>(wq_s + kWarps * 48);        // [3][D] q digits
    float* m_s = reinterpret_cast<float*>(qd_s + 3 * D);                 // [kWarps]
    float* l_s = m_s + kWarps;                                           // [kWarps]
    float* scr_s = l_s + kWarps;                                         // [2][kWarps] setup scratch
    uint64_t* full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(scr_s + 2 * kWarps) + 7) & ~uintptr_t(7));
    uint64_t* empty = full + S;

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = split * P.tokens_per_split;
    const int t_end = min(P.T, t_begin + P.tokens_per_split);
    const int n_tiles = (t_end - t_begin + kTile - 1) / kTile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto issue = [&](int tile_idx) {
        const int s = tile_idx % S;
        const int t0 = t_begin + tile_idx * kTile;
        issue_tile<C::kKRow, C::kVRow, C::kKTile, PAGED>(P, b, h, t0, min(kTile, t_end - t0), smem_raw + s * C::kStage, &full[s]);
    };
    __syncthreads();  // barriers initialised before the first bulk copy can complete on them
    if (tid == 0) {
        for (int i = 0; i < PD && i < n_tiles; ++i) issue(i);
    }

    // ---- setup: q -> 24-bit fixed point (three signed byte planes); max v_scale of the split
    float qv = 0.f;
    if (tid < D) qv = load_scale(P.q, P.q_dtype, ((int64_t)b * P.Hq + h) * D + tid) * P.qk_scale_log2;
    float vmax = 0.f;
    for (int t = t_begin + tid; t < t_end; t += NT * 8) {  // 8 independent loads in flight
        float v8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v8[u] = t + u * NT < t_end ? __ldg(P.v_scale + tok_index<PAGED>(P, t + u * NT)) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) vmax = fmaxf(vmax, v8[u]);
    }
    float qa = fabsf(qv);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { qa = fmaxf(qa, __shfl_xor_sync(0xffffffffu, qa, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o)); }
    if (lane == 0) { scr_s[warp] = qa; scr_s[kWarps + warp] = vmax; }
    __syncthreads();
    float qmax = 0.f, vsmax = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kWarps; ++w8) { qmax = fmaxf(qmax, scr_s[w8]); vsmax = fmaxf(vsmax, scr_s[kWarps + w8]); }
    const float dq = qmax > 0.f ? qmax / kFixedMax : 1.f;                                    // q = Q * dq
    const float dw = vsmax > 0.f ? vsmax * exp2f(kMhaLazyLog2) / kFixedMax : 1.f;            // weight = W * dw, weights <= 2^lazy * vsmax
    const float inv_dw = 1.f / dw;
    if (tid < D) {
        const int Q = __float2int_rn(qv / dq);
        const int d0 = ((Q + 128) & 255) - 128, Q1 = (Q - d0) >> 8;
        const int d1 = ((Q1 + 128) & 255) - 128, d2 = (Q1 - d1) >> 8;
        qd_s[tid] = (int8_t)d0; qd_s[D + tid] = (int8_t)d1; qd_s[2 * D + tid] = (int8_t)d2;
    }
    __syncthreads();

    // K-phase lane roles: kc = 16-channel chunk, kr = row of the pass; register j of the lane is pass j ^ swzp
    const int kc = lane % LPR, kr = lane / LPR;
    const int swzp = kc >> 1;
    const int my_row = warp * WR + swzp * RPP + kr;  // the token (row of the tile) whose score this lane ends up with
    const bool s_writer = (kc & 1) == 0;             // lanes kc and kc ^ 1 hold the same token
    // V-phase lane roles: 8 channels of a row; row group hw takes the token groups hw * GPH ..
    const int vc = lane % WPR, hw = lane / WPR;
    uint8_t* wq_w = wq_s + warp * 48;

    uint32_t qk[3][4];  // this lane's 16 channels of q, three planes x four words, in the element order of the K words
    int kinit[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const uint4 t = *reinterpret_cast<const uint4*>(qd_s + pl * D + kc * 16);
        if (KBITS == 8) { qk[pl][0] = t.x; qk[pl][1] = t.y; qk[pl][2] = t.z; qk[pl][3] = t.w; kinit[pl] = 0; }
        else {
            // INT4 rows: word j holds elements 8j..8j+7, high nibbles = even elements, low = odd
            qk[pl][0] = __byte_perm(t.x, t.y, 0x6420); qk[pl][1] = __byte_perm(t.x, t.y, 0x7531);
            qk[pl][2] = __byte_perm(t.z, t.w, 0x6420); qk[pl][3] = __byte_perm(t.z, t.w, 0x7531);
            // nibbles are used unsigned (code + 8): start the accumulator at -8 * sum(q digits)
            kinit[pl] = -8 * (dp4a_ss(0x01010101u, t.x, 0) + dp4a_ss(0x01010101u, t.y, 0) + dp4a_ss(0x01010101u, t.z, 0) +
                              dp4a_ss(0x01010101u, t.w, 0));
        }
    }

    int iacc[8][3];
    float facc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) { facc[d] = 0.f; iacc[d][0] = iacc[d][1] = iacc[d][2] = 0; }
    float m_run = -INFINITY, l_run = 0.f;
    auto flush = [&](float a) {  // int32 accumulators -> fp32, then move to the new reference
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const float v = fmaf((float)iacc[d][2], 65536.f, fmaf((float)iacc[d][1], 256.f, (float)iacc[d][0]));
            facc[d] = fmaf(v, dw, facc[d]) * a;
            iacc[d][0] = iacc[d][1] = iacc[d][2] = 0;
        }
    };

    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % S;
        if (tid == 0 && it + PD < n_tiles) {
            const int nt = it + PD;  // goes into the stage tile nt - S used: every warp must have left it
            if (nt >= S) { while (!mbar_try_wait(&empty[nt % S], ((nt / S) - 1) & 1)) {} }
            issue(nt);
        }
        const int t0 = t_begin + it * kTile;
        const int rows = min(kTile, t_end - t0);
        const bool valid = my_row < rows;
        const int64_t ti = valid ? tok_index<PAGED>(P, t0 + my_row) : 0;
        const float ks = valid ? __ldg(P.k_scale + ti) : 0.f;
        const float vs = valid ? __ldg(P.v_scale + ti) : 0.f;
        while (!mbar_try_wait(&full[s], (it / S) & 1)) {}
        const uint8_t* sb = smem_raw + s * C::kStage;

        // ---- K phase: RTT rows per lane, 12 DP4A each; reduce-scatter over the LPR lanes of a row group
        float dsum[RTT];
#pragma unroll
        for (int j = 0; j < RTT; ++j) {
            const int row = warp * WR + (j ^ swzp) * RPP + kr;
            int a0 = kinit[0], a1 = kinit[1], a2 = kinit[2];
            if (KBITS == 8) {
                const uint4 kw = *reinterpret_cast<const uint4*>(sb + row * C::kKRow + kc * 16);
                const uint32_t w[4] = {kw.x, kw.y, kw.z, kw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) { a0 = dp4a_ss(w[i], qk[0][i], a0); a1 = dp4a_ss(w[i], qk[1][i], a1); a2 = dp4a_ss(w[i], qk[2][i], a2); }
            } else {
                const uint2 kw = *reinterpret_cast<const uint2*>(sb + row * C::kKRow + kc * 8);
                const uint32_t w[4] = {(kw.x >> 4) & 0x0f0f0f0fu, kw.x & 0x0f0f0f0fu, (kw.y >> 4) & 0x0f0f0f0fu, kw.y & 0x0f0f0f0fu};
#pragma unroll
                for (int i = 0; i < 4; ++i) { a0 = dp4a_us(w[i], qk[0][i], a0); a1 = dp4a_us(w[i], qk[1][i], a1); a2 = dp4a_us(w[i], qk[2][i], a2); }
            }
            dsum[j] = fmaf((float)a2, 65536.f, fmaf((float)a1, 256.f, (float)a0));
        }
        {
            int n = RTT;
#pragma unroll
            for (int o = LPR / 2; o >= 1; o >>= 1) {
                if (n > 1) {
                    n >>= 1;
#pragma unroll
                    for (int j = 0; j < RTT / 2; ++j)
                        if (j < n) dsum[j] += __shfl_xor_sync(0xffffffffu, dsum[j + n], o);
                } else {
                    dsum[0] += __shfl_xor_sync(0xffffffffu, dsum[0], o);
                }
            }
        }
        // ---- softmax: one token per lane pair; warp max by REDUX; lazy reference
        const float sv = valid ? dsum[0] * (dq * ks) : -INFINITY;
        const float mt = __int_as_float(float_order_key(__int_as_float(__reduce_max_sync(0xffffffffu, float_order_key(sv)))));
        const bool up = mt > m_run + kMhaLazyLog2;  // warp-uniform; false when every row of the warp is masked
        if (up) {
            flush(m_run == -INFINITY ? 0.f : exp2f(m_run - mt));
            l_run *= m_run == -INFINITY ? 0.f : exp2f(m_run - mt);
            m_run = mt;
        } else if ((it & 1023) == 1023) {
            flush(1.f);  // bound the int32 accumulators on very long splits
        }
        const float pw = valid ? exp2f(sv - m_run) : 0.f;
        if (s_writer) {
            l_run += pw;
            const uint32_t W = (uint32_t)__float2int_rn(pw * vs * inv_dw);
            const int tok = swzp * RPP + kr;  // token of the warp's 16
            wq_w[tok] = (uint8_t)(W & 255u); wq_w[16 + tok] = (uint8_t)((W >> 8) & 255u); wq_w[32 + tok] = (uint8_t)(W >> 16);
        }
        __syncwarp();

        // ---- V phase: byte-transpose four tokens' words, one DP4A per (channel, weight plane)
#pragma unroll
        for (int gi = 0; gi < GPH; ++gi) {
            const int g = hw * GPH + gi;  // tokens 4g .. 4g + 3 of the warp's 16
            uint32_t wd[3];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) wd[pl] = *reinterpret_cast<const uint32_t*>(wq_w + pl * 16 + 4 * g);
            uint32_t lo[4], hi[4];  // INT8: lo = channels 0..3, hi = channels 4..7 (signed bytes); INT4: odd / even channels
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint8_t* src = sb + C::kKTile + (warp * WR + 4 * g + i) * C::kVRow + vc * VB;
                if (VBITS == 8) { const uint2 t = *reinterpret_cast<const uint2*>(src); lo[i] = t.x; hi[i] = t.y; }
                else {
                    const uint32_t t = *reinterpret_cast<const uint32_t*>(src);
                    lo[i] = __vsub4(t & 0x0f0f0f0fu, 0x08080808u);          // elements 1,3,5,7
                    hi[i] = __vsub4((t >> 4) & 0x0f0f0f0fu, 0x08080808u);   // elements 0,2,4,6
                }
            }
            uint32_t cl[4], ch[4];
            transpose4x4(lo[0], lo[1], lo[2], lo[3], cl);
            transpose4x4(hi[0], hi[1], hi[2], hi[3], ch);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                // channel index inside the lane's 8: INT8 -> lo: c4, hi: 4 + c4; INT4 -> lo: 2*c4 + 1, hi: 2*c4
                const int dl = VBITS == 8 ? c4 : 2 * c4 + 1, dh = VBITS == 8 ? 4 + c4 : 2 * c4;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    iacc[dl][pl] = dp4a_su(cl[c4], wd[pl], iacc[dl][pl]);
                    iacc[dh][pl] = dp4a_su(ch[c4], wd[pl], iacc[dh][pl]);
                }
            }
        }
        __syncwarp();                           // wq_w is rewritten by the next tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the stage before its async-proxy (bulk copy) refill
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
    }

    // ---- merge the warps' partials (own reference each) through shared memory; the ring is idle after the barrier
    flush(1.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l_run += __shfl_xor_sync(0xffffffffu, l_run, o);
    if (lane == 0) { m_s[warp] = m_run; l_s[warp] = l_run; }
    __syncthreads();
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) M = fmaxf(M, m_s[w]);
    const float fac = m_run == -INFINITY ? 0.f : exp2f(m_run - M);
    float* red_s = reinterpret_cast<float*>(smem_raw);  // [kWarps * HW][D]
    {
        float* dst = red_s + (warp * HW + hw) * D + vc * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(facc[0] * fac, facc[1] * fac, facc[2] * fac, facc[3] * fac);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(facc[4] * fac, facc[5] * fac, facc[6] * fac, facc[7] * fac);
    }
    __syncthreads();
    const int64_t hq = (int64_t)b * P.Hq + h;
    if (tid < D) {
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < kWarps * HW; ++r) sum += red_s[r * D + tid];
        P.ws_acc[(hq * P.n_splits + split) * D + tid] = sum;
    }
    if (tid == 0) {
        float L = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) L += m_s[w] == -INFINITY ? 0.f : l_s[w] * exp2f(m_s[w] - M);
        P.ws_ml[(hq * P.n_splits + split) * 2] = M;
        P.ws_ml[(hq * P.n_splits + split) * 2 + 1] = L;
    }
}


// =====================================================================================================
// Any head_dim (1 .. 512), any Hq / H, any strides and alignment: the coverage kernel behind the tiled ones (the reference
// path takes any model: benchmarker.py:473-477 hands the cache to transformers' attention).  Grid (n_splits, Hq, B), four
// warps; a warp walks the split's tokens four apart, a lane owns channels lane, lane + 32, ...; byte loads, fp32 math, online
// softmax per warp, warps merged through shared memory.  Not a bandwidth kernel.
// =====================================================================================================
template <int KBITS, int VBITS>
__global__ void __launch_bounds__(128) decode_attn_generic_kernel(AttnParams P, int D, int G) {
    constexpr int kMaxPerLane = 16;  // D <= 512
    __shared__ float red_s[4][512];
    __shared__ float ml_s[4][2];
    const int split = blockIdx.x, hq_in = blockIdx.y, b = blockIdx.z, h = hq_in / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = split * P.tokens_per_split, t_end = min(P.T, t_begin + P.tokens_per_split);
    const int64_t hq = (int64_t)b * P.Hq + hq_in;
    const int npl = (D + 31) / 32;
    float qv[kMaxPerLane], acc[kMaxPerLane];
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int d = lane + 32 * i;
        qv[i] = (i < npl && d < D) ? load_scale(P.q, P.q_dtype, hq * D + d) * P.qk_scale_log2 : 0.f;
        acc[i] = 0.f;
    }
    const uint8_t* kp = P.k + (int64_t)b * P.kB + (int64_t)h * P.kH;
    const uint8_t* vp = P.v + (int64_t)b * P.vB + (int64_t)h * P.vH;
    auto code = [](const uint8_t* row, int d, int bits) -> float {
        if (bits == 8) return (float)reinterpret_cast<const int8_t*>(row)[d];
        const uint8_t byte = row[d >> 1];
        return (float)((d & 1) ? (byte & 15) : (byte >> 4)) - 8.f;
    };
    float m = -INFINITY, l = 0.f;
    for (int t = t_begin + warp; t < t_end; t += 4) {
        const uint8_t* kr = kp + (int64_t)t * P.kT;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int d = lane + 32 * i;
            if (i < npl && d < D) dot = fmaf(qv[i], code(kr, d, KBITS), dot);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float sc = dot * P.k_scale[t];
        const float mn = fmaxf(m, sc);
        const float a = m == -INFINITY ? 0.f : exp2f(m - mn), p = exp2f(sc - mn);
        const float w = p * P.v_scale[t];
        const uint8_t* vr = vp + (int64_t)t * P.vT;
        l = l * a + p;
        m = mn;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int d = lane + 32 * i;
            if (i < npl && d < D) acc[i] = fmaf(w, code(vr, d, VBITS), acc[i] * a);
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int d = lane + 32 * i;
        if (i < npl && d < D) red_s[warp][d] = acc[i];
    }
    if (lane == 0) { ml_s[warp][0] = m; ml_s[warp][1] = l; }
    __syncthreads();
    float M = -INFINITY;
    for (int w = 0; w < 4; ++w) M = fmaxf(M, ml_s[w][0]);
    for (int d = threadIdx.x; d < D; d += 128) {
        float sum = 0.f;
        for (int w = 0; w < 4; ++w) sum += ml_s[w][0] == -INFINITY ? 0.f : red_s[w][d] * exp2f(ml_s[w][0] - M);
        P.ws_acc[(hq * P.n_splits + split) * D + d] = sum;
    }
    if (threadIdx.x == 0) {
        float L = 0.f;
        for (int w = 0; w < 4; ++w) L += ml_s[w][0] == -INFINITY ? 0.f : ml_s[w][1] * exp2f(ml_s[w][0] - M);
        P.ws_ml[(hq * P.n_splits + split) * 2] = M;
        P.ws_ml[(hq * P.n_splits + split) * 2 + 1] = L;
    }
}

// Optional exact "current token" (the reference's decode step sees the new token's K/V unquantized:
// benchmarker.py:473-477 concatenates them inside the model before they are quantized at :486): it is merged here
// as one more split with m = q.k_cur * scale, l = 1, acc = v_cur.
struct CurToken {
    const void* q; const void* k; const void* v;  // q [B,Hq,D]; k, v [B,H,D]; all of dtype `dtype`
    int dtype, G;
    float qk_scale_log2;
};

template <int ODT>
__global__ void __launch_bounds__(1024) decode_attn_merge_kernel(const float* __restrict__ ws_acc, const float* __restrict__ ws_ml, int n_splits, int D,
                                                                 void* __restrict__ out, CurToken cur) {
    // blockDim = (D rounded up to a multiple of 32, SEG): thread (d, seg) merges the splits seg, seg + SEG, ... of output d, eight
    // independent loads per round -- with SEG = ceil(n_splits / 8) the whole workspace read is ONE round trip after the dependency
    // wait instead of n_splits / 8 dependent ones (this kernel is on the critical path of every layer: 3-4 us of a 25 us layer at the
    // per-GPU shapes of an 8-way head shard) -- then segment 0 folds the SEG partial results and the step's own token.
    __shared__ float dot_s[16];
    extern __shared__ float seg_s[];  // [3][SEG - 1][blockDim.x]: M, num, den of the segments 1 ..
    const int64_t hq = blockIdx.x;
    const int d = threadIdx.x, seg = threadIdx.y, SEG = blockDim.y, Dp = blockDim.x;
    const bool live = d < D;
    float s_cur = -INFINITY, v_cur = 0.f;
    if (cur.k) {  // independent of the split kernel: runs before the dependency wait
        const int64_t kv_row = (hq / cur.G);  // (b * Hq + hq_in) / G == b * H + h  because Hq = G * H
        float p = live && seg == 0 ? load_scale(cur.q, cur.dtype, hq * D + d) * load_scale(cur.k, cur.dtype, kv_row * D + d) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (seg == 0 && (d & 31) == 0) dot_s[d >> 5] = p;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < Dp / 32; ++w) tot += dot_s[w];
        s_cur = tot * cur.qk_scale_log2;
        v_cur = live ? load_scale(cur.v, cur.dtype, kv_row * D + d) : 0.f;
    }
    // launched with programmatic stream serialization: everything above this line overlaps the split kernel's tail
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const float2* ml = reinterpret_cast<const float2*>(ws_ml) + hq * n_splits;
    const float* acc = ws_acc + hq * n_splits * D + (live ? d : 0);
    float M = -INFINITY, num = 0.f, den = 0.f;
    for (int s0 = seg; s0 < n_splits; s0 += 8 * SEG) {  // 8 splits per round: independent loads, then an online-softmax merge
        float2 m8[8];
        float a8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool ok = s0 + u * SEG < n_splits;
            m8[u] = ok ? ml[s0 + u * SEG] : make_float2(-INFINITY, 0.f);
            a8[u] = ok ? acc[(int64_t)(s0 + u * SEG) * D] : 0.f;
        }
        float Mn = M;
#pragma unroll
        for (int u = 0; u < 8; ++u) Mn = fmaxf(Mn, m8[u].x);
        const float r = M == -INFINITY ? 0.f : exp2f(M - Mn);
        num *= r; den *= r; M = Mn;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float w = m8[u].x == -INFINITY ? 0.f : exp2f(m8[u].x - M);
            num = fmaf(a8[u], w, num);
            den = fmaf(m8[u].y, w, den);
        }
    }
    if (seg > 0) {
        float* o = seg_s + (seg - 1) * Dp + d;
        o[0] = M; o[(SEG - 1) * Dp] = num; o[2 * (SEG - 1) * Dp] = den;
    }
    if (SEG > 1) __syncthreads();
    if (seg > 0 || !live) return;
    float Mt = fmaxf(M, s_cur);
    for (int g = 0; g < SEG - 1; ++g) Mt = fmaxf(Mt, seg_s[g * Dp + d]);
    float w0 = M == -INFINITY ? 0.f : exp2f(M - Mt);
    num *= w0; den *= w0;
    if (cur.k) { const float w = exp2f(s_cur - Mt); num = fmaf(v_cur, w, num); den += w; }
    for (int g = 0; g < SEG - 1; ++g) {
        const float mg = seg_s[g * Dp + d];
        const float w = mg == -INFINITY ? 0.f : exp2f(mg - Mt);
        num = fmaf(seg_s[((SEG - 1) + g) * Dp + d], w, num);
        den = fmaf(seg_s[(2 * (SEG - 1) + g) * Dp + d], w, den);
    }
    store_scalar<ODT>(out, hq * D + d, num / den);
}

template <int ODT>
int launch_merge(unsigned blocks, unsigned threads, cudaStream_t st, const float* ws_acc, const float* ws_ml, int n_splits, int D, void* out,
                 const CurToken& cur) {
    unsigned seg = (unsigned)(n_splits + 7) / 8;
    const unsigned seg_max = 1024u / threads < 8u ? 1024u / threads : 8u;
    seg = seg < 1 ? 1 : seg > seg_max ? seg_max : seg;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads, seg); cfg.dynamicSmemBytes = 3 * (seg - 1) * threads * sizeof(float); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    KVQ_CUDA(cudaLaunchKernelEx(&cfg, decode_attn_merge_kernel<ODT>, ws_acc, ws_ml, n_splits, D, out, cur));
    return KVQ_OK;
}

template <int KBITS, int VBITS, int D>
int launch_mha(dim3 grid, cudaStream_t st, const AttnParams& P) {
    using C = MhaCfg<KBITS, VBITS, D>;
    auto kern = decode_attn_mha_kernel<KBITS, VBITS, D, false>;
    auto kern_paged = decode_attn_mha_kernel<KBITS, VBITS, D, true>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    KVQ_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        KVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        KVQ_CUDA(cudaFuncSetAttribute(kern_paged, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        configured_dev = dev;
    }
    if (P.table) kern_paged<<<grid, C::NT, C::kSmem, st>>>(P);
    else kern<<<grid, C::NT, C::kSmem, st>>>(P);
    return KVQ_OK;
}

template <int KBITS, int VBITS, int G, int D>
int launch_gqa(dim3 grid, cudaStream_t st, const AttnParams& P) {
    using C = GqaCfg<KBITS, VBITS, G, D>;
    constexpr int NT = C::NT;
    auto kern = decode_attn_gqa_kernel<KBITS, VBITS, G, D, false>;
    auto kern_paged = decode_attn_gqa_kernel<KBITS, VBITS, G, D, true>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    KVQ_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        KVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        KVQ_CUDA(cudaFuncSetAttribute(kern_paged, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        configured_dev = dev;
    }
    if (P.table) kern_paged<<<grid, NT, C::kSmem, st>>>(P);
    else kern<<<grid, NT, C::kSmem, st>>>(P);
    return KVQ_OK;
}

template <int KBITS, int VBITS, int G>
int launch_split(int D, dim3 grid, cudaStream_t st, const AttnParams& P, bool ring) {
    // `ring`: every plane's code rows are back to back, so tiles can be staged with bulk copies -- always true for the cache
    // stores of this package.  G == 1: per-warp DP4A kernel; G in {2, 4, 8}: register-tiled FFMA2 kernel.  Strided rows take
    // the register-staged fallback.
    if (ring) {
        if constexpr (G == 1) return D == 128 ? launch_mha<KBITS, VBITS, 128>(grid, st, P) : launch_mha<KBITS, VBITS, 64>(grid, st, P);
        else return D == 128 ? launch_gqa<KBITS, VBITS, G, 128>(grid, st, P) : launch_gqa<KBITS, VBITS, G, 64>(grid, st, P);
    }
    if (D == 128) decode_attn_split_kernel<KBITS, VBITS, G, 128><<<grid, kThreads, 0, st>>>(P);
    else decode_attn_split_kernel<KBITS, VBITS, G, 64><<<grid, kThreads, 0, st>>>(P);
    return KVQ_OK;
}
template <int KBITS, int VBITS>
int launch_g(int G, int D, dim3 grid, cudaStream_t st, const AttnParams& P, bool tma) {
    switch (G) {
        case 1: return launch_split<KBITS, VBITS, 1>(D, grid, st, P, tma);
        case 2: return launch_split<KBITS, VBITS, 2>(D, grid, st, P, tma);
        case 4: return launch_split<KBITS, VBITS, 4>(D, grid, st, P, tma);
        case 8: return launch_split<KBITS, VBITS, 8>(D, grid, st, P, tma);
    }
    return fail(KVQ_ERR_UNSUPPORTED, "decode_attn: Hq/H must be 1, 2, 4 or 8 (got %d)", G);
}

int choose_splits(int64_t B, int64_t H, int64_t T, int64_t G) {
    // The grid is B * H * n CTAs on `slots` resident CTAs (2 per SM for every variant).  A grid of 1.08 waves -- what "one split
    // more than fits" gives, e.g. 64 planes x 5 splits on 296 slots -- runs its last 24 CTAs on an idle machine and costs almost
    // a whole extra wave (measured: C5 at N = 1).  Pick the split count with the best wave efficiency
    //     planes * n / (ceil(planes * n / slots) * slots)
    // among those that keep a split at >= 4 tiles, charging a little for every extra split (setup + merge work).
    (void)G;
    const int64_t slots = 2 * (int64_t)sm_count(), planes = B * H;
    int64_t max_n = (T + 511) / 512;
    if (max_n > 64) max_n = 64;
    if (max_n < 1) max_n = 1;
    int64_t best = 1;
    double best_score = -1.0;
    for (int64_t n = 1; n <= max_n; ++n) {
        const int64_t ctas = planes * n, waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        const double score = eff - 0.004 * (double)n - (waves > 1 ? 0.01 * (double)(waves - 1) : 0.0);
        if (score > best_score) { best_score = score; best = n; }
    }
    return (int)best;
}

}  // namespace
}  // namespace kvq

using namespace kvq;

extern "C" int64_t kvq_decode_attn_workspace_bytes(int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, int* n_splits) {
    if (B <= 0 || Hq <= 0 || H <= 0 || T <= 0 || D <= 0) { if (n_splits) *n_splits = 1; return 0; }
    const int n = choose_splits(B, H, T, Hq / H);
    if (n_splits) *n_splits = n;
    return B * Hq * n * (D + 2) * (int64_t)sizeof(float);
}

extern "C" int kvq_decode_attn(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                               const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                               const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                               void* out, int out_dtype, void* workspace, int n_splits, kvq_stream_t stream) {
    return kvq_decode_attn_ex(q, q_dtype, k_codes, k_bits, k_stride3, k_scale, v_codes, v_bits, v_stride3, v_scale, B, Hq, H, T, D,
                              sm_scale, nullptr, nullptr, out, out_dtype, workspace, n_splits, stream);
}

namespace {
// common tail of the contiguous and paged entry points: split geometry, kernel choice, merge
int run_decode_attn(AttnParams& P, int k_bits, int v_bits, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                    const void* k_cur, const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits, bool tma,
                    int variant, cudaStream_t st, bool generic = false) {
    P.B = (int)B; P.Hq = (int)Hq; P.H = (int)H; P.T = (int)T;
    P.qk_scale_log2 = sm_scale * 1.4426950408889634f;
    int64_t tps = (T + n_splits - 1) / n_splits;
    tps = (tps + kTile - 1) / kTile * kTile;
    P.tokens_per_split = (int)tps;
    n_splits = (int)((T + tps - 1) / tps);  // splits that actually contain tokens
    P.n_splits = n_splits;
    P.ws_acc = (float*)workspace;
    P.ws_ml = P.ws_acc + B * Hq * (int64_t)n_splits * D;
    const dim3 grid((unsigned)n_splits, (unsigned)H, (unsigned)B);
    const int G = (int)(Hq / H);
    int rc;
    if (generic) {   // head_dim other than 64 / 128, a grouping the tiled kernels are not built for, or unaligned rows
        const dim3 gg((unsigned)n_splits, (unsigned)Hq, (unsigned)B);
        if (k_bits == 8 && v_bits == 8) decode_attn_generic_kernel<8, 8><<<gg, 128, 0, st>>>(P, (int)D, G);
        else if (k_bits == 8 && v_bits == 4) decode_attn_generic_kernel<8, 4><<<gg, 128, 0, st>>>(P, (int)D, G);
        else if (k_bits == 4 && v_bits == 4) decode_attn_generic_kernel<4, 4><<<gg, 128, 0, st>>>(P, (int)D, G);
        else decode_attn_generic_kernel<4, 8><<<gg, 128, 0, st>>>(P, (int)D, G);
        rc = KVQ_OK;
    } else {
    if (variant == KVQ_ATTN_AUTO) variant = kvq_decode_attn_default_variant((int64_t)G);
    if (variant == KVQ_ATTN_IMMA) rc = launch_attn_mma(P, k_bits, v_bits, G, (int)D, grid, st);
    else if (k_bits == 8 && v_bits == 8) rc = launch_g<8, 8>(G, (int)D, grid, st, P, tma);
    else if (k_bits == 8 && v_bits == 4) rc = launch_g<8, 4>(G, (int)D, grid, st, P, tma);
    else if (k_bits == 4 && v_bits == 4) rc = launch_g<4, 4>(G, (int)D, grid, st, P, tma);
    else rc = launch_g<4, 8>(G, (int)D, grid, st, P, tma);
    }
    if (rc != KVQ_OK) return rc;
    rc = check_launch("decode_attn_split");
    if (rc != KVQ_OK) return rc;
    const unsigned mt = (unsigned)((D + 31) / 32 * 32);
    const CurToken cur{P.q, k_cur, v_cur, P.q_dtype, G, P.qk_scale_log2};
    switch (out_dtype) {
        case KVQ_F32: rc = launch_merge<KVQ_F32>((unsigned)(B * Hq), mt, st, P.ws_acc, P.ws_ml, n_splits, (int)D, out, cur); break;
        case KVQ_F16: rc = launch_merge<KVQ_F16>((unsigned)(B * Hq), mt, st, P.ws_acc, P.ws_ml, n_splits, (int)D, out, cur); break;
        default: rc = launch_merge<KVQ_BF16>((unsigned)(B * Hq), mt, st, P.ws_acc, P.ws_ml, n_splits, (int)D, out, cur); break;
    }
    if (rc != KVQ_OK) return rc;
    return check_launch("decode_attn_merge");
}
}  // namespace

extern "C" int kvq_decode_attn_ex(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                                  const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                                  const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                                  const void* k_cur, const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits,
                                  kvq_stream_t stream) {
    return kvq_decode_attn_variant(q, q_dtype, k_codes, k_bits, k_stride3, k_scale, v_codes, v_bits, v_stride3, v_scale, B, Hq, H, T, D,
                                   sm_scale, k_cur, v_cur, out, out_dtype, workspace, n_splits, KVQ_ATTN_AUTO, stream);
}

extern "C" int kvq_decode_attn_default_variant(int64_t q_heads_per_kv_head) {
    // G >= 2: every code feeds several query heads and the CUDA-core kernels are FMA-bound, not HBM-bound
    return q_heads_per_kv_head >= KVQ_ATTN_IMMA_MIN_G ? KVQ_ATTN_IMMA : KVQ_ATTN_CUDA_CORE;
}

extern "C" int kvq_decode_attn_variant(const void* q, int q_dtype, const void* k_codes, int k_bits, const int64_t k_stride3[3],
                                       const float* k_scale, const void* v_codes, int v_bits, const int64_t v_stride3[3],
                                       const float* v_scale, int64_t B, int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale,
                                       const void* k_cur, const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits,
                                       int variant, kvq_stream_t stream) {
    KVQ_REQUIRE(variant >= KVQ_ATTN_AUTO && variant <= KVQ_ATTN_IMMA, "unknown attention variant %d", variant);
    KVQ_REQUIRE((k_cur == nullptr) == (v_cur == nullptr), "k_cur and v_cur must be given together");
    KVQ_REQUIRE(q && k_codes && v_codes && k_scale && v_scale && out && workspace, "null pointer");
    KVQ_REQUIRE(q_dtype >= 0 && q_dtype <= 2 && out_dtype >= 0 && out_dtype <= 2, "bad dtype");
    KVQ_REQUIRE((k_bits == 8 || k_bits == 4) && (v_bits == 8 || v_bits == 4), "k_bits / v_bits must be 8 or 4");
    KVQ_REQUIRE(D >= 1 && D <= 512, "decode_attn: head_dim must be in [1, 512] (got %lld)", (long long)D);
    KVQ_REQUIRE(B >= 1 && H >= 1 && Hq >= 1 && T >= 1 && T <= 0x7fffffff && B <= 65535 && H <= 65535 && Hq <= 65535, "bad shape");
    KVQ_REQUIRE(Hq % H == 0, "Hq must be a multiple of H");
    KVQ_REQUIRE(n_splits >= 1, "n_splits must be >= 1");
    const int64_t kw = k_bits == 8 ? D : (D + 1) / 2, vw = v_bits == 8 ? D : (D + 1) / 2;
    const Stride3 ks = make_stride3(k_stride3, H, T, kw), vs = make_stride3(v_stride3, H, T, vw);
    auto al = [](const void* p, const Stride3& s, int a) {
        return reinterpret_cast<uintptr_t>(p) % a == 0 && s.sB % a == 0 && s.sH % a == 0 && s.sT % a == 0;
    };
    // the tiled kernels: head_dim 64 / 128, 1 / 2 / 4 / 8 query heads per KV head, 16-byte-aligned rows; everything else -> generic kernel
    const int64_t Gq = Hq / H;
    const bool generic = !((D == 64 || D == 128) && (Gq == 1 || Gq == 2 || Gq == 4 || Gq == 8) && al(k_codes, ks, 16) && al(v_codes, vs, 16));
    AttnParams P{};
    P.q = q; P.q_dtype = q_dtype;
    P.k = (const uint8_t*)k_codes; P.kB = ks.sB; P.kH = ks.sH; P.kT = ks.sT; P.k_scale = k_scale;
    P.v = (const uint8_t*)v_codes; P.vB = vs.sB; P.vH = vs.sH; P.vT = vs.sT; P.v_scale = v_scale;
    // the bulk-copy pipeline needs each plane's rows back to back (true for the cache store); scales are read with plain loads
    const bool tma = ks.sT == kw && vs.sT == vw;
    return run_decode_attn(P, k_bits, v_bits, B, Hq, H, T, D, sm_scale, k_cur, v_cur, out, out_dtype, workspace, n_splits, tma,
                           variant, (cudaStream_t)stream, generic);
}

extern "C" int kvq_decode_attn_paged(const void* q, int q_dtype, const void* k_pool, int k_bits, const int64_t k_pool_stride3[3],
                                     const float* k_scale_pool, const void* v_pool, int v_bits, const int64_t v_pool_stride3[3],
                                     const float* v_scale_pool, const int32_t* block_table, int64_t block_size, int64_t B,
                                     int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale, const void* k_cur,
                                     const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits,
                                     kvq_stream_t stream) {
    return kvq_decode_attn_paged_variant(q, q_dtype, k_pool, k_bits, k_pool_stride3, k_scale_pool, v_pool, v_bits, v_pool_stride3,
                                         v_scale_pool, block_table, block_size, B, Hq, H, T, D, sm_scale, k_cur, v_cur, out, out_dtype,
                                         workspace, n_splits, KVQ_ATTN_AUTO, stream);
}

extern "C" int kvq_decode_attn_paged_variant(const void* q, int q_dtype, const void* k_pool, int k_bits, const int64_t k_pool_stride3[3],
                                             const float* k_scale_pool, const void* v_pool, int v_bits, const int64_t v_pool_stride3[3],
                                             const float* v_scale_pool, const int32_t* block_table, int64_t block_size, int64_t B,
                                             int64_t Hq, int64_t H, int64_t T, int64_t D, float sm_scale, const void* k_cur,
                                             const void* v_cur, void* out, int out_dtype, void* workspace, int n_splits, int variant,
                                             kvq_stream_t stream) {
    KVQ_REQUIRE(variant >= KVQ_ATTN_AUTO && variant <= KVQ_ATTN_IMMA, "unknown attention variant %d", variant);
    KVQ_REQUIRE((k_cur == nullptr) == (v_cur == nullptr), "k_cur and v_cur must be given together");
    KVQ_REQUIRE(q && k_pool && v_pool && k_scale_pool && v_scale_pool && block_table && out && workspace, "null pointer");
    KVQ_REQUIRE(q_dtype >= 0 && q_dtype <= 2 && out_dtype >= 0 && out_dtype <= 2, "bad dtype");
    KVQ_REQUIRE((k_bits == 8 || k_bits == 4) && (v_bits == 8 || v_bits == 4), "k_bits / v_bits must be 8 or 4");
    KVQ_REQUIRE(D == 64 || D == 128, "decode_attn_paged: head_dim must be 64 or 128 (got %lld)", (long long)D);
    KVQ_REQUIRE(B >= 1 && H >= 1 && Hq >= 1 && T >= 1 && T <= 0x7fffffff && B <= 65535 && H <= 65535, "bad shape");
    KVQ_REQUIRE(Hq % H == 0, "Hq must be a multiple of H");
    const int64_t G = Hq / H;
    KVQ_REQUIRE(G == 1 || G == 2 || G == 4 || G == 8, "decode_attn_paged: Hq/H must be 1, 2, 4 or 8 (got %lld)", (long long)G);
    KVQ_REQUIRE(block_size >= 16 && (block_size & (block_size - 1)) == 0, "decode_attn_paged: block_size must be a power of two >= 16 (got %lld)",
                (long long)block_size);
    KVQ_REQUIRE(n_splits >= 1, "n_splits must be >= 1");
    const int64_t kw = k_bits == 8 ? D : D / 2, vw = v_bits == 8 ? D : D / 2;
    // pool strides in bytes: {physical block, batch, head}; rows of a (block, b, h) are back to back
    int64_t ks[3] = {B * H * block_size * kw, H * block_size * kw, block_size * kw};
    int64_t vs[3] = {B * H * block_size * vw, H * block_size * vw, block_size * vw};
    if (k_pool_stride3) { ks[0] = k_pool_stride3[0]; ks[1] = k_pool_stride3[1]; ks[2] = k_pool_stride3[2]; }
    if (v_pool_stride3) { vs[0] = v_pool_stride3[0]; vs[1] = v_pool_stride3[1]; vs[2] = v_pool_stride3[2]; }
    auto al16 = [](const void* p, const int64_t* s) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && s[0] % 16 == 0 && s[1] % 16 == 0 && s[2] % 16 == 0; };
    KVQ_REQUIRE(al16(k_pool, ks) && al16(v_pool, vs), "decode_attn_paged: pools must be 16-byte aligned with 16-byte-multiple strides");
    AttnParams P{};
    P.q = q; P.q_dtype = q_dtype;
    P.k = (const uint8_t*)k_pool; P.kBlk = ks[0]; P.kB = ks[1]; P.kH = ks[2]; P.kT = kw; P.k_scale = k_scale_pool;
    P.v = (const uint8_t*)v_pool; P.vBlk = vs[0]; P.vB = vs[1]; P.vH = vs[2]; P.vT = vw; P.v_scale = v_scale_pool;
    P.table = block_table; P.bs = (int)block_size;
    int sh = 0;
    while ((1LL << sh) < block_size) ++sh;
    P.bs_shift = sh;
    return run_decode_attn(P, k_bits, v_bits, B, Hq, H, T, D, sm_scale, k_cur, v_cur, out, out_dtype, workspace, n_splits, true,
                           variant, (cudaStream_t)stream);
}
