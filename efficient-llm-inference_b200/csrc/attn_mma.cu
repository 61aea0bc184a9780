// attn_mma.cu -- area 4, variant KVQ_ATTN_IMMA: the split kernel of the fused dequant decode attention on tensor-map TMA
// staging and the INT8 tensor-core MMA (sm_100a).
//
// Why it exists.  The CUDA-core kernels of attn.cu convert or multiply every code once per query head: at Llama-3-70B's
// GQA-8 that is 16 FMA per 1.5 bytes and the FMA pipe, not HBM, bounds the step (0.21 of the roofline, r1); at G = 1 the
// DP4A kernel still issues 5.85 instructions per (token, channel) pair (0.63).  But the codes ARE integers and both
// contractions are tiny GEMMs per KV head --  scores[t, g] = sum_d K[t, d] q[g, d]  and  out[g, d] = sum_t w[g, t] V[t, d]
// -- whose second operand can be made integer exactly the way the DP4A kernel already does it: q as three signed byte
// planes (24-bit fixed point per head), the softmax weights w = p * v_scale[t] as four unsigned byte planes (32-bit fixed
// point against 2^8 * max(v_scale), lazy reference).  Then one  mma.sync.m16n8k32.s32.{s8|u8}.{s8|u8}  does 4096 MACs where the
// DP4A kernel needs 1024 instructions, for any G <= 8, and the kernel is back on the HBM roofline.  north_star words decode
// attention as "a bandwidth-bound split-K GEMV with no tensor cores, because decode at batch <= 16 is not a dense
// contraction": at G = 8 that premise is measurably false, so this variant is offered
// next to the CUDA-core kernels, selected per call (kvq_decode_attn_variant), and DESIGN.md states the deviation.
//
// Data movement.  One elected thread streams tiles of 128 tokens with cp.async.bulk.tensor (SASS UTMALDG) from 4-D tensor maps
// over the code tensors [row bytes, T, H, B] (5-D [row, block, H, B, n_phys] through the block table for a paged store) into
// a ring of shared-memory stages, hardware-swizzled (128B / 64B / 32B swizzle = the row size), completing on an mbarrier
// per stage, together with the tile's k_scale / v_scale (two 512-byte bulk copies).  There is no producer thread: the LAST
// warp to leave a stage (an acq_rel counter per stage) refills it at once, so the whole ring stays in flight.  The swizzle is what makes
// the MMA fragment loads conflict-free: a fragment wants the SAME 16-byte column of 8 different token rows, and unswizzled
// rows of 128 bytes would put all eight in the same banks.
//
// Per CTA: 4 warps; a warp owns 32 tokens of each tile, runs its own online softmax over them (no block barrier in the main
// loop) and keeps int32 accumulators for all of D in registers:
//   K phase : A = K codes straight from shared memory (ldmatrix; INT4 rows split into even / odd nibbles, q's fragment
//             permuted to match), B = q byte planes, columns = (plane, head); 2 m-tiles x D/32 k-steps x ceil(3G/8) MMAs.
//   softmax : plane sums -> fp32 score, warp max per head by shuffles, lazy reference (moves only when exceeded by 2^8),
//             exp2, weight -> 32-bit fixed point = four bytes, exchanged through a 1 KB per-warp patch of shared memory in
//             A-fragment order.
//   V phase : B = V codes, four tokens' words byte-transposed in registers (8 PRMT per 16 bytes; packed INT4 is
//             transposed first and split into nibbles after, the high nibble used in place as 16 x code), A = weight planes,
//             rows = (plane, head); D/8 n-tiles x ceil(4G/16) MMAs per 32 tokens.
//   flush   : int32 -> fp32 (per-warp accumulators in shared memory) only when a reference moves or the split ends.
// The token <-> fragment-slot assignment (tok_of) is chosen so that every shared-memory access of the main loop is
// conflict-free under the hardware swizzle and the K-phase lane that computes a weight is the lane that packs it.
// Partials (m, l, acc[D]) per (b, hq, split) go to the same workspace as the CUDA-core kernels; the merge kernel is shared.
#include <cuda.h>  // CUtensorMap and the driver enums only: the encode entry point is fetched at run time (no libcuda link)

#include <cstring>

#include "attn_common.cuh"

namespace kvq {
namespace {

#ifndef KVQ_ATTN_MMA_CTAS_PER_SM
#define KVQ_ATTN_MMA_CTAS_PER_SM 2
#endif
#ifndef KVQ_ATTN_MMA_MAX_STAGES
#define KVQ_ATTN_MMA_MAX_STAGES 4
#endif
constexpr int kMT = 128;            // threads per CTA
constexpr int kMW = kMT / 32;       // warps
// The softmax reference of a head moves only when a tile maximum exceeds it by 2^kLazy (a move costs a flush of every
// accumulator of the warp).  The weights p * v_scale[t] are 32-bit fixed point against 2^kLazy * max(v_scale): FOUR unsigned byte
// planes -- which fill the MMA's 16-row tiles exactly for G = 4 and 8, so the fourth plane is free -- i.e. 24 bits below a
// weight of 1 and 8 bits of headroom above it: after the first tile a reference practically never moves.
constexpr float kLazy = 8.f;
constexpr float kWMax = 4294967040.f;   // largest weight word: 2^32 - 256 (exact in fp32, so the conversion cannot overflow)

// byte offset inside a 1024-byte-aligned tile whose rows have ROW bytes -> offset after the TMA swizzle of that row size
// (CUTLASS Swizzle<log2(ROW/16), 4, 3>: address bits [7, 7 + B) are XOR-ed into bits [4, 4 + B))
template <int ROW> __device__ __forceinline__ uint32_t swz(uint32_t o) { return o ^ (((o >> 7) & (ROW / 16 - 1)) << 4); }

// Token (0..31 inside the warp's slice) that sits in k-slot r (0..3) of V-phase group (c, half).  K-phase lane g = lane / 4 owns
// group (g & 3, g >> 2) and its four tokens in the same order.  The bit assignment depends on the V row size because the
// V-phase load of (token, 16 / 8 / 4-byte piece g') must spread the four c-lanes of a piece over different banks under the
// swizzle of that row size; for every choice the 8 groups' r-th tokens are distinct mod 8, which is what ldmatrix needs.
template <int VROW> __device__ __forceinline__ int tok_of(int c, int half, int r) {
    const int r0 = (r & 1) ^ half, r1 = r >> 1;
    if (VROW == 128) return (half << 4) | (r1 << 3) | (c << 1) | r0;
    if (VROW == 64) return (half << 4) | (r1 << 3) | ((c >> 1) << 2) | (r0 << 1) | (c & 1);
    return (half << 4) | (r1 << 3) | (r0 << 2) | c;  // VROW == 32
}

template <int KBITS, int VBITS, int G, int D> struct MmaCfg {
    static constexpr int KROW = D * KBITS / 8, VROW = D * VBITS / 8;
    static constexpr int kKTile = kTile * KROW, kVTile = kTile * VROW;
    static constexpr int kScl = 2 * kTile * 4;                          // k_scale[128] ++ v_scale[128] of the tile, fp32
    static constexpr int kStage = kKTile + kVTile + kScl;               // (a multiple of 1024: swizzle atoms stay aligned)
    static constexpr int NCOL = 3 * G;                 // K phase: (q plane, head) pairs = B columns
    static constexpr int U = (NCOL + 7) / 8;           // K phase: MMAs per k-step
    static constexpr int NROW = 4 * G;                 // V phase: (weight plane, head) pairs = A rows (4 planes fill the m-tiles exactly)
    static constexpr int MTV = (NROW + 15) / 16;       // V phase: m-tiles
    static constexpr int NKM = D / 32, NT = D / 8;     // K phase k-steps; V phase n-tiles
    static constexpr int HS = G == 1 ? 1 : 2;          // heads whose score a lane ends up with
    static constexpr int VB = VROW / 8;                // V bytes a lane loads per token: 16 / 8 / 4
    static constexpr int VW = VB / 4;                  // ... in words
    // shared memory after the ring
    static constexpr int oFacc = 0;                                     // float [kMW][G][D]
    static constexpr int oWp = oFacc + kMW * G * D * 4;                 // u32   [kMW][16 * MTV][8]   weight planes, A-fragment order
    static constexpr int oQf = oWp + kMW * 16 * MTV * 8 * 4;            // u32   [U][NKM][2][32]      q planes, B-fragment order
    static constexpr int oQd = oQf + U * NKM * 2 * 32 * 4;              // int8  [8 * U][D]           q digits (setup only)
    static constexpr int oDq = oQd + 8 * U * D;                         // float [G]  dq, int [8 * U] kinit, float [kMW][G] fa
    static constexpr int oKi = oDq + G * 4;
    static constexpr int oFa = oKi + 8 * U * 4;
    static constexpr int oMl = oFa + kMW * G * 4;                       // float [kMW][G][2]
    static constexpr int oScr = oMl + kMW * G * 8;                      // float [kMW]
    static constexpr int oBar = (oScr + kMW * 4 + 7) & ~7;              // full[S] (mbarriers), then done[S] (u32 counters)
    static constexpr int kFixed = oBar + KVQ_ATTN_MMA_MAX_STAGES * 8 + KVQ_ATTN_MMA_MAX_STAGES * 4 + 8;
    static constexpr int kBudget = 227 * 1024 / KVQ_ATTN_MMA_CTAS_PER_SM - 1024 /*per-CTA reservation*/ - 1024 /*alignment slack*/ - kFixed;
    static constexpr int kStagesRaw = kBudget / kStage;
    static constexpr int S = kStagesRaw > KVQ_ATTN_MMA_MAX_STAGES ? KVQ_ATTN_MMA_MAX_STAGES : (kStagesRaw < 2 ? 2 : kStagesRaw);
    static constexpr int kSmem = 1024 + S * kStage + kFixed;
};

// mbarrier wait with a bounded spin: a transaction-count mismatch (a bug) traps after a few seconds instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D = A (16 x 32, row) * B (32 x 8, col) + C, int8 operands, int32 accumulate.  AS / BS: operand is signed
template <bool AS, bool BS>
__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if constexpr (AS && BS)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else if constexpr (!AS && BS)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else if constexpr (AS && !BS)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int KBITS, int VBITS, int G, int D, bool PAGED>
__global__ void __launch_bounds__(kMT, KVQ_ATTN_MMA_CTAS_PER_SM)
decode_attn_mma_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap, AttnParams P) {
    using C = MmaCfg<KBITS, VBITS, G, D>;
    constexpr int KROW = C::KROW, VROW = C::VROW, U = C::U, MTV = C::MTV, NKM = C::NKM, NT = C::NT, HS = C::HS, VB = C::VB, VW = C::VW;
    constexpr int S = C::S;
    constexpr int NCOL = C::NCOL, NROW = C::NROW;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms are 1024-byte aligned
    uint8_t* fx = ring + S * C::kStage;
    float* facc_s = reinterpret_cast<float*>(fx + C::oFacc);
    uint32_t* wp_s = reinterpret_cast<uint32_t*>(fx + C::oWp);
    uint32_t* qf_s = reinterpret_cast<uint32_t*>(fx + C::oQf);
    int8_t* qd_s = reinterpret_cast<int8_t*>(fx + C::oQd);
    float* dq_s = reinterpret_cast<float*>(fx + C::oDq);
    int* ki_s = reinterpret_cast<int*>(fx + C::oKi);
    float* fa_s = reinterpret_cast<float*>(fx + C::oFa);
    float* ml_s = reinterpret_cast<float*>(fx + C::oMl);
    float* scr_s = reinterpret_cast<float*>(fx + C::oScr);
    uint64_t* full = reinterpret_cast<uint64_t*>(fx + C::oBar);
    unsigned* done = reinterpret_cast<unsigned*>(full + KVQ_ATTN_MMA_MAX_STAGES);   // warps that have left stage s

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    const int t_begin = split * P.tokens_per_split;
    const int t_end = min(P.T, t_begin + P.tokens_per_split);
    const int n_tiles = (t_end - t_begin + kTile - 1) / kTile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); done[s] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&kmap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&vmap) : "memory");
    }
    // the tile's scales travel with it (two 512-byte bulk copies) when the scale vectors are 16-byte aligned; otherwise,
    // for a tile that ends at T, and for a paged store the lanes load their scales themselves
    const bool scl_bulk = !PAGED && ((reinterpret_cast<uintptr_t>(P.k_scale) | reinterpret_cast<uintptr_t>(P.v_scale)) & 15u) == 0;
    auto issue = [&](int tile_idx) {   // one thread
        const int s = tile_idx % S;
        const int t0 = t_begin + tile_idx * kTile;
        uint8_t* sb = ring + s * C::kStage;
        if constexpr (!PAGED) {
            const bool scl = scl_bulk && t0 + kTile <= t_end;
            mbar_expect_tx(&full[s], C::kKTile + C::kVTile + (scl ? C::kScl : 0));   // out-of-range rows of the last tile are zero-filled and counted
            tma_load_4d(sb, &kmap, 0, t0, h, b, &full[s]);
            tma_load_4d(sb + C::kKTile, &vmap, 0, t0, h, b, &full[s]);
            if (scl) {
                bulk_g2s(sb + C::kKTile + C::kVTile, P.k_scale + t0, kTile * 4, &full[s]);
                bulk_g2s(sb + C::kKTile + C::kVTile + kTile * 4, P.v_scale + t0, kTile * 4, &full[s]);
            }
        } else {
            const int rows = min(kTile, t_end - t0);
            const int blk0 = t0 >> P.bs_shift;
            if (P.bs >= kTile) {   // the tile is one 128-row piece of a block
                const int phys = __ldg(P.table + blk0), w = t0 & (P.bs - 1);
                mbar_expect_tx(&full[s], C::kKTile + C::kVTile);
                tma_load_5d(sb, &kmap, 0, w, h, b, phys, &full[s]);
                tma_load_5d(sb + C::kKTile, &vmap, 0, w, h, b, phys, &full[s]);
            } else {               // 128 / bs whole blocks, one box each
                const int np = (rows + P.bs - 1) >> P.bs_shift;
                int phys[kTile / 16];
#pragma unroll
                for (int i = 0; i < kTile / 16; ++i) phys[i] = i < np ? __ldg(P.table + blk0 + i) : 0;
                mbar_expect_tx(&full[s], (uint32_t)(np << P.bs_shift) * (KROW + VROW));
#pragma unroll
                for (int i = 0; i < kTile / 16; ++i) {
                    if (i < np) {
                        tma_load_5d(sb + (i << P.bs_shift) * KROW, &kmap, 0, 0, h, b, phys[i], &full[s]);
                        tma_load_5d(sb + C::kKTile + (i << P.bs_shift) * VROW, &vmap, 0, 0, h, b, phys[i], &full[s]);
                    }
                }
            }
            if (t0 + kTile < P.T) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.table + ((t0 + kTile) >> P.bs_shift)));
        }
    };
    __syncthreads();   // barriers initialised before the first copy can complete on them
    if (tid == 0) {
        for (int i = 0; i < S && i < n_tiles; ++i) issue(i);   // the whole ring; later tiles are issued by whoever frees a stage
    }

    // ---------------------------------------------------------------- setup: q planes, max v_scale, per-warp state
    for (int i = tid; i < kMW * G * D; i += kMT) facc_s[i] = 0.f;
    for (int i = tid; i < kMW * 16 * MTV * 8; i += kMT) wp_s[i] = 0u;   // rows >= 3G stay zero for the whole kernel
    for (int i = tid; i < 8 * U * D / 4; i += kMT) reinterpret_cast<uint32_t*>(qd_s)[i] = 0u;
    __syncthreads();
    for (int hd = warp; hd < G; hd += kMW) {   // a warp per head: max |q|, then three balanced signed byte digits per channel
        float qv[D / 32], qa = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            qv[i] = load_scale(P.q, P.q_dtype, ((int64_t)b * P.Hq + (int64_t)h * G + hd) * D + lane + 32 * i) * P.qk_scale_log2;
            qa = fmaxf(qa, fabsf(qv[i]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qa = fmaxf(qa, __shfl_xor_sync(0xffffffffu, qa, o));
        const float dq = qa > 0.f ? qa / kFixedMax : 1.f;
        if (lane == 0) dq_s[hd] = dq;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            const int Q = __float2int_rn(qv[i] / dq);
            const int d0 = ((Q + 128) & 255) - 128, Q1 = (Q - d0) >> 8;
            const int d1 = ((Q1 + 128) & 255) - 128, d2 = (Q1 - d1) >> 8;
            const int d = lane + 32 * i;
            qd_s[(0 * G + hd) * D + d] = (int8_t)d0; qd_s[(1 * G + hd) * D + d] = (int8_t)d1; qd_s[(2 * G + hd) * D + d] = (int8_t)d2;
        }
    }
    float vmax = 0.f;
    if (scl_bulk) {   // 16-byte-aligned, contiguous scales: 128-bit loads (t_begin is a multiple of the tile)
        const int n4 = (t_end - t_begin) >> 2;
        const float4* v4p = reinterpret_cast<const float4*>(P.v_scale + t_begin);
        for (int i = tid; i < n4; i += kMT * 4) {
            float4 x[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) x[u4] = i + u4 * kMT < n4 ? __ldg(v4p + i + u4 * kMT) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) vmax = fmaxf(vmax, fmaxf(fmaxf(x[u4].x, x[u4].y), fmaxf(x[u4].z, x[u4].w)));
        }
        for (int t = t_begin + 4 * n4 + tid; t < t_end; t += kMT) vmax = fmaxf(vmax, __ldg(P.v_scale + t));
    } else
    for (int t = t_begin + tid; t < t_end; t += kMT * 8) {   // 8 independent loads in flight
        float v8[8];
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) v8[u8] = t + u8 * kMT < t_end ? __ldg(P.v_scale + tok_index<PAGED>(P, t + u8 * kMT)) : 0.f;
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) vmax = fmaxf(vmax, v8[u8]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) scr_s[warp] = vmax;
    __syncthreads();
    // q planes in B-fragment order: word [u][km][w][ln] is what lane ln feeds MMA (k-step km, column block u) as b0 (w = 0) / b1
    for (int i = tid; i < U * NKM * 2 * 32; i += kMT) {
        const int ln = i & 31, w = (i >> 5) & 1, km = (i >> 6) % NKM, u = (i >> 6) / NKM;
        const int col = 8 * u + (ln >> 2), cc = ln & 3;
        uint32_t word = 0u;
        if (col < NCOL) {
            const uint8_t* row = reinterpret_cast<const uint8_t*>(qd_s) + col * D;
            if (KBITS == 8) {
                word = *reinterpret_cast<const uint32_t*>(row + 32 * km + 16 * w + 4 * cc);
            } else {   // packed rows: word (chunk km, lane cc) holds channels 32 km + 8 cc + 0..7; even ones are the high nibbles
                const uint8_t* p8 = row + 32 * km + 8 * cc + w;
                word = (uint32_t)p8[0] | ((uint32_t)p8[2] << 8) | ((uint32_t)p8[4] << 16) | ((uint32_t)p8[6] << 24);
            }
        }
        qf_s[i] = word;
    }
    if (tid < 8 * U) {   // INT4 K codes are used as unsigned nibbles (code + 8): start the accumulator at -8 * sum(q digits)
        int sum = 0;
        if (KBITS == 4 && tid < NCOL) {
            for (int d = 0; d < D; ++d) sum += qd_s[tid * D + d];
        }
        ki_s[tid] = -8 * sum;
    }
    __syncthreads();
    float vsmax = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kMW; ++w8) vsmax = fmaxf(vsmax, scr_s[w8]);
    const float dw = vsmax > 0.f ? vsmax * exp2f(kLazy) / kWMax : 1.f;   // weight = W * dw, W < 2^32
    const float inv_dw = 1.f / dw;

    // this lane's K-phase columns: idx = 8u + 2c + j  ->  (plane, head); weight of the column in the fp32 score; C init
    float colw[U][2];
    int cinit[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = 8 * u + 2 * c + j;
            const bool ok = idx < NCOL;
            const int pl = ok ? idx / G : 0, hd = ok ? idx % G : 0;
            colw[u][j] = ok ? dq_s[hd] * (pl == 0 ? 1.f : pl == 1 ? 256.f : 65536.f) : 0.f;
            cinit[u][j] = ok ? ki_s[idx] : 0;
        }
    // head of score slot hs in this lane, and whether this lane publishes it (the others hold duplicates)
    const int head0 = G == 8 ? 2 * c : G == 4 ? 2 * (c & 1) : 0;
    const bool writer = G == 8 ? true : G == 4 ? c < 2 : c == 0;
    // K-phase roles: lane g owns V-phase group (g & 3, g >> 2); ldmatrix row addresses: lane -> (matrix lane >> 3, row lane & 7)
    const int lm_gi = lane & 7, lm_rh = (lane >> 3) & 1, lm_cs = lane >> 4;
    uint32_t lm_row[2], lm_x[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row = warp * 32 + tok_of<VROW>(lm_gi & 3, lm_gi >> 2, 2 * mt + lm_rh);
        lm_row[mt] = (uint32_t)row * KROW;
        lm_x[mt] = (((uint32_t)row * KROW >> 7) & (KROW / 16 - 1)) << 4;
    }
    int my_tok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) my_tok[r] = warp * 32 + tok_of<VROW>(g & 3, g >> 2, r);
    // V-phase roles: lane (g, c) loads piece g of tokens tok_of(c, half, r); A rows g and g + 8 of every m-tile
    uint32_t v_off[2][4];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int r = 0; r < 4; ++r) v_off[hf][r] = swz<VROW>((uint32_t)(warp * 32 + tok_of<VROW>(c, hf, r)) * VROW + g * VB);
    uint32_t* wp_w = wp_s + warp * 16 * MTV * 8;
    float* facc_w = facc_s + warp * G * D;
    float* fa_w = fa_s + warp * G;

    int acc[MTV][NT][4];
#pragma unroll
    for (int mt = 0; mt < MTV; ++mt)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0;
    int srow[MTV][2];   // INT4 V: per-row sums of the weight bytes this lane fed (for the -8 offset of the unsigned nibbles)
#pragma unroll
    for (int mt = 0; mt < MTV; ++mt) srow[mt][0] = srow[mt][1] = 0;
    float m_ref[HS], l_run[HS];
#pragma unroll
    for (int hs = 0; hs < HS; ++hs) { m_ref[hs] = -INFINITY; l_run[hs] = 0.f; }

    // int32 accumulators -> the warp's fp32 accumulators in shared memory, then scale head hd by fa_w[hd].  Planes go in three
    // passes (different lanes hold the planes of one head when G < 8): same order every time, so results are reproducible.
    auto flush = [&]() {
        int sfull[MTV][2];
        if (VBITS == 4) {
#pragma unroll
            for (int mt = 0; mt < MTV; ++mt)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    int v = srow[mt][rh];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    sfull[mt][rh] = v;
                    srow[mt][rh] = 0;
                }
        }
        __syncwarp();
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
            const float wpl = dw * (pl == 0 ? 1.f : pl == 1 ? 256.f : pl == 2 ? 65536.f : 16777216.f);
#pragma unroll
            for (int mt = 0; mt < MTV; ++mt)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    if (16 * mt + 8 * rh >= NROW) continue;   // rows that do not exist for this G
                    const int row = 16 * mt + 8 * rh + g;
                    if (row < NROW && row / G == pl) {
                        // fragment-native layout [j][e][head][c] (channel (2c + e) * NT + j of head `row % G`): the lanes of a pass
                        // touch consecutive words.  (Head-major [head][channel] put all 32 lanes of every access in ONE bank: the
                        // flush cost ~15 us per split, 15 % of the kernel at 57 tiles per split and 40 % at 14.)
                        float* dst = facc_w + (row % G) * 4 + c;
#pragma unroll
                        for (int j = 0; j < NT; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                int v = acc[mt][j][2 * rh + e];
                                float sc = wpl;
                                if (VBITS == 4) {   // tile j: even = high nibbles used as 16 * (code + 8), odd = low nibbles (code + 8)
                                    v -= ((j & 1) ? 8 : 128) * sfull[mt][rh];
                                    if (!(j & 1)) sc *= 0.0625f;
                                }
                                const int ix = (j * 2 + e) * 4 * G;
                                dst[ix] = fmaf((float)v, sc, dst[ix]);
                            }
                    }
                }
            __syncwarp();
        }
#pragma unroll
        for (int mt = 0; mt < MTV; ++mt)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0;
        for (int i = lane; i < G * D; i += 32) facc_w[i] *= fa_w[(i >> 2) % G];
        __syncwarp();
    };

    // scales of the lane's four tokens when they do not come with the tile: fetched one tile ahead
    float ks_n[4], vs_n[4];
    auto fetch_scales = [&](int tile_idx) {
        const int t0 = t_begin + tile_idx * kTile;
        if (tile_idx >= n_tiles || (scl_bulk && t0 + kTile <= t_end)) return;   // (block-uniform) nothing to fetch: comes with the tile
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = t0 + my_tok[r];
            const bool ok = t < t_end;
            const int64_t ti = ok ? tok_index<PAGED>(P, t) : 0;
            ks_n[r] = ok ? __ldg(P.k_scale + ti) : 0.f;
            vs_n[r] = ok ? __ldg(P.v_scale + ti) : 0.f;
        }
    };
#pragma unroll
    for (int r = 0; r < 4; ++r) { ks_n[r] = 0.f; vs_n[r] = 0.f; }
    fetch_scales(0);
    constexpr bool kQfRegs = U == 1;   // G <= 2: the lane's q fragments (2 words per k-step) live in registers
    uint32_t qfr[kQfRegs ? NKM : 1][2];
    if constexpr (kQfRegs) {
#pragma unroll
        for (int km = 0; km < NKM; ++km) { qfr[km][0] = qf_s[(km * 2 + 0) * 32 + lane]; qfr[km][1] = qf_s[(km * 2 + 1) * 32 + lane]; }
    }
    bool started = false;

    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % S;
        const int t0 = t_begin + it * kTile;
        const int rows = min(kTile, t_end - t0);
        const bool scl_here = scl_bulk && rows == kTile;   // this tile's scales are in the stage
        float ks[4], vs[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { ks[r] = ks_n[r]; vs[r] = vs_n[r]; }
        fetch_scales(it + 1);
        mbar_wait(&full[s], (it / S) & 1);
        __syncwarp();   // lanes leave the spin loop one by one: reconverge before the .aligned instructions
        const uint32_t kb = smem_u32(ring + s * C::kStage);
        const uint8_t* vbp = ring + s * C::kStage + C::kKTile;
        if (scl_here) {
            const float* sclp = reinterpret_cast<const float*>(vbp + C::kVTile);
#pragma unroll
            for (int r = 0; r < 4; ++r) { ks[r] = sclp[my_tok[r]]; vs[r] = sclp[kTile + my_tok[r]]; }
        }

        // ---------------------------------------------------------------- K phase
        int sacc[2][U][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < U; ++u) { sacc[mt][u][0] = sacc[mt][u][2] = cinit[u][0]; sacc[mt][u][1] = sacc[mt][u][3] = cinit[u][1]; }
        if constexpr (KBITS == 8) {
#pragma unroll
            for (int km = 0; km < NKM; ++km) {
                uint32_t b0[U], b1[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if constexpr (kQfRegs) { b0[u] = qfr[km][0]; b1[u] = qfr[km][1]; }
                    else { b0[u] = qf_s[((u * NKM + km) * 2 + 0) * 32 + lane]; b1[u] = qf_s[((u * NKM + km) * 2 + 1) * 32 + lane]; }
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t a0, a1, a2, a3;
                    ldsm_x4(kb + lm_row[mt] + ((uint32_t)((2 * km + lm_cs) * 16) ^ lm_x[mt]), a0, a1, a2, a3);
#pragma unroll
                    for (int u = 0; u < U; ++u) imma<true, true>(sacc[mt][u], a0, a1, a2, a3, b0[u], b1[u]);
                }
            }
        } else {
#pragma unroll
            for (int kp = 0; kp < NKM / 2; ++kp) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t w0, w1, w2, w3;   // (rows g, chunk 2kp), (rows g + 8, chunk 2kp), (rows g, chunk 2kp + 1), (rows g + 8, chunk 2kp + 1)
                    ldsm_x4(kb + lm_row[mt] + ((uint32_t)((2 * kp + lm_cs) * 16) ^ lm_x[mt]), w0, w1, w2, w3);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int km = 2 * kp + sub;
                        const uint32_t x = sub ? w2 : w0, y = sub ? w3 : w1;
                        const uint32_t a0 = (x >> 4) & 0x0f0f0f0fu, a1 = (y >> 4) & 0x0f0f0f0fu, a2 = x & 0x0f0f0f0fu, a3 = y & 0x0f0f0f0fu;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if constexpr (kQfRegs) imma<false, true>(sacc[mt][u], a0, a1, a2, a3, qfr[km][0], qfr[km][1]);
                            else imma<false, true>(sacc[mt][u], a0, a1, a2, a3, qf_s[((u * NKM + km) * 2 + 0) * 32 + lane], qf_s[((u * NKM + km) * 2 + 1) * 32 + lane]);
                        }
                    }
                }
            }
        }
        // ---------------------------------------------------------------- scores, online softmax, weights
        float sc[4][HS];   // token r (m-tile r >> 1, row half r & 1) x head slot
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float part[HS];
#pragma unroll
            for (int hs = 0; hs < HS; ++hs) part[hs] = 0.f;
            if constexpr (G == 8) {   // column block u IS plane u: planes 0 and 1 combine exactly in int32 (|sum| < 2^21 each)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int lo = sacc[r >> 1][0][2 * (r & 1) + j] + (sacc[r >> 1][1][2 * (r & 1) + j] << 8);
                    part[j] = fmaf((float)sacc[r >> 1][2][2 * (r & 1) + j], colw[2][j], (float)lo * colw[0][j]);
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < 2; ++j) part[G == 1 ? 0 : j] = fmaf((float)sacc[r >> 1][u][2 * (r & 1) + j], colw[u][j], part[G == 1 ? 0 : j]);
            }
#pragma unroll
            for (int hs = 0; hs < HS; ++hs) {
                if (G <= 2) part[hs] += __shfl_xor_sync(0xffffffffu, part[hs], 1);
                if (G <= 4) part[hs] += __shfl_xor_sync(0xffffffffu, part[hs], 2);
                sc[r][hs] = my_tok[r] < rows ? part[hs] * ks[r] : -INFINITY;
            }
        }
        float mt_[HS];
        bool up = false;
#pragma unroll
        for (int hs = 0; hs < HS; ++hs) {
            float m = fmaxf(fmaxf(sc[0][hs], sc[1][hs]), fmaxf(sc[2][hs], sc[3][hs]));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            mt_[hs] = m;
            up |= m > m_ref[hs] + kLazy;
        }
        if (__any_sync(0xffffffffu, up)) {   // rare: some head's reference moves
            float a[HS];
#pragma unroll
            for (int hs = 0; hs < HS; ++hs) {
                const bool mv = mt_[hs] > m_ref[hs] + kLazy;
                a[hs] = !mv ? 1.f : (m_ref[hs] == -INFINITY ? 0.f : exp2f(m_ref[hs] - mt_[hs]));
                if (mv) { l_run[hs] *= a[hs]; m_ref[hs] = mt_[hs]; }
                if (writer && g == 0) fa_w[head0 + hs] = a[hs];
            }
            if (started) flush();   // (the first tile starts from zero accumulators: nothing to rescale)
            started = true;
        } else if ((it & 511) == 511) {   // bound the int32 accumulators on very long splits
            if (lane < G) fa_w[lane] = 1.f;
            flush();
        }
#pragma unroll
        for (int hs = 0; hs < HS; ++hs) {
            uint32_t W[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float p = exp2f(sc[r][hs] - m_ref[hs]);   // -inf score -> 0
                l_run[hs] += p;
                W[r] = __float2uint_rn(p * vs[r] * inv_dw);
            }
            if (writer) {   // byte plane pl of the four tokens, in k-slot order, at A-fragment position (row (pl, head), group g)
                const int hd = head0 + hs, slot = (g & 3) * 2 + (g >> 2);
                const uint32_t lo01 = __byte_perm(W[0], W[1], 0x5140), lo23 = __byte_perm(W[2], W[3], 0x5140);   // bytes 0,0',1,1'
                const uint32_t hi01 = __byte_perm(W[0], W[1], 0x7362), hi23 = __byte_perm(W[2], W[3], 0x7362);   // bytes 2,2',3,3'
                wp_w[(0 * G + hd) * 8 + slot] = __byte_perm(lo01, lo23, 0x5410);
                wp_w[(1 * G + hd) * 8 + slot] = __byte_perm(lo01, lo23, 0x7632);
                wp_w[(2 * G + hd) * 8 + slot] = __byte_perm(hi01, hi23, 0x5410);
                wp_w[(3 * G + hd) * 8 + slot] = __byte_perm(hi01, hi23, 0x7632);
            }
        }
        __syncwarp();

        // ---------------------------------------------------------------- V phase
        uint32_t af[MTV][4];
#pragma unroll
        for (int mt = 0; mt < MTV; ++mt) {
            const uint2 lo = *reinterpret_cast<const uint2*>(wp_w + (16 * mt + g) * 8 + 2 * c);   // (group (c, 0), group (c, 1)) of row g
            af[mt][0] = lo.x; af[mt][2] = lo.y;
            if (16 * mt + 8 < NROW) {
                const uint2 hi = *reinterpret_cast<const uint2*>(wp_w + (16 * mt + 8 + g) * 8 + 2 * c);
                af[mt][1] = hi.x; af[mt][3] = hi.y;
            } else { af[mt][1] = 0u; af[mt][3] = 0u; }
            if (VBITS == 4) {
                srow[mt][0] = dp4a_us(af[mt][2], 0x01010101u, dp4a_us(af[mt][0], 0x01010101u, srow[mt][0]));   // weight bytes are unsigned
                srow[mt][1] = dp4a_us(af[mt][3], 0x01010101u, dp4a_us(af[mt][1], 0x01010101u, srow[mt][1]));
            }
        }
        uint32_t vw[2][4][VW];   // the lane's piece (VB bytes) of its eight tokens, one vector load each
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint8_t* src = vbp + v_off[hf][r];
                if constexpr (VW == 4) { const uint4 x = *reinterpret_cast<const uint4*>(src); vw[hf][r][0] = x.x; vw[hf][r][1] = x.y; vw[hf][r][2] = x.z; vw[hf][r][3] = x.w; }
                else if constexpr (VW == 2) { const uint2 x = *reinterpret_cast<const uint2*>(src); vw[hf][r][0] = x.x; vw[hf][r][1] = x.y; }
                else vw[hf][r][0] = *reinterpret_cast<const uint32_t*>(src);
            }
#pragma unroll
        for (int i = 0; i < VW; ++i) {   // word column i of the lane's piece: 4 byte columns = 4 (INT8) / 8 (INT4) n-tiles
            uint32_t t[2][4];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) transpose4x4(vw[hf][0][i], vw[hf][1][i], vw[hf][2][i], vw[hf][3][i], t[hf]);
#pragma unroll
            for (int bcol = 0; bcol < 4; ++bcol) {
                if constexpr (VBITS == 8) {
                    const int j = 4 * i + bcol;
#pragma unroll
                    for (int mt = 0; mt < MTV; ++mt) imma<false, true>(acc[mt][j], af[mt][0], af[mt][1], af[mt][2], af[mt][3], t[0][bcol], t[1][bcol]);
                } else {
                    const int j = 2 * (4 * i + bcol);
                    const uint32_t h0 = t[0][bcol] & 0xf0f0f0f0u, h1 = t[1][bcol] & 0xf0f0f0f0u;   // even channel: 16 * (code + 8)
                    const uint32_t l0 = t[0][bcol] & 0x0f0f0f0fu, l1 = t[1][bcol] & 0x0f0f0f0fu;   // odd channel: code + 8
#pragma unroll
                    for (int mt = 0; mt < MTV; ++mt) {
                        imma<false, false>(acc[mt][j], af[mt][0], af[mt][1], af[mt][2], af[mt][3], h0, h1);
                        imma<false, false>(acc[mt][j + 1], af[mt][0], af[mt][1], af[mt][2], af[mt][3], l0, l1);
                    }
                }
            }
        }
        // The stage was read through the generic proxy (ldmatrix / ld.shared) and is overwritten through the async proxy (TMA)
        // as soon as the fourth warp has arrived.  Without this cross-proxy fence the refill can overtake reads that were issued
        // but not yet performed (ptxas also hoists the arrive above the last MMAs): measured on B200 as rare wrong V rows when the
        // tiles come from a warm L2 (two in three runs at G = 8 before the fence, none in 80 after).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();                           // the weight patch is rewritten by the next tile
        if (lane == 0) {                        // the LAST warp to leave stage s refills it at once: the whole ring stays in flight
            unsigned old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done[s])) : "memory");
            if (old == kMW - 1) {
                done[s] = 0u;   // next touched only after the refilled tile has been consumed
                if (it + S < n_tiles) issue(it + S);
            }
        }
    }

    // ---------------------------------------------------------------- end of the split: flush, merge the four warps, write the partial
    if (lane < G) fa_w[lane] = 1.f;
    __syncwarp();
    flush();
#pragma unroll
    for (int hs = 0; hs < HS; ++hs) {
        float l = l_run[hs];
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        if (writer && g == 0) { ml_s[(warp * G + head0 + hs) * 2] = m_ref[hs]; ml_s[(warp * G + head0 + hs) * 2 + 1] = l; }
    }
    __syncthreads();
    for (int o = tid; o < G * D; o += kMT) {   // o walks the fragment-native layout [j][e][head][c]
        const int hd = (o >> 2) % G, je = o / (4 * G), ch = (2 * (o & 3) + (je & 1)) * NT + (je >> 1);
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kMW; ++w) M = fmaxf(M, ml_s[(w * G + hd) * 2]);
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kMW; ++w) {
            const float mw = ml_s[(w * G + hd) * 2];
            sum = fmaf(facc_s[w * G * D + o], mw == -INFINITY ? 0.f : exp2f(mw - M), sum);
        }
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + hd;
        P.ws_acc[(hq * P.n_splits + split) * D + ch] = sum;
    }
    if (tid < G) {
        float M = -INFINITY, Lsum = 0.f;
#pragma unroll
        for (int w = 0; w < kMW; ++w) M = fmaxf(M, ml_s[(w * G + tid) * 2]);
#pragma unroll
        for (int w = 0; w < kMW; ++w) {
            const float mw = ml_s[(w * G + tid) * 2];
            Lsum += mw == -INFINITY ? 0.f : ml_s[(w * G + tid) * 2 + 1] * exp2f(mw - M);
        }
        const int64_t hq = (int64_t)b * P.Hq + (int64_t)h * G + tid;
        P.ws_ml[(hq * P.n_splits + split) * 2] = M;
        P.ws_ml[(hq * P.n_splits + split) * 2 + 1] = Lsum;
    }
}

// ---------------------------------------------------------------------------------------------------- host: tensor maps

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct MapKey { const void* base; uint64_t dims[5], strides[4]; uint32_t box_rows, row, rank; };

// uint8 tensor map over code rows: dims (innermost first) {row bytes, tokens, ...}, box {row, box_rows, 1, ...}, swizzle = row size.
// A small per-thread cache keeps the last maps: a decode loop asks for the same stores every step.
int make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, uint32_t row, uint32_t box_rows) {
    static thread_local struct { MapKey key; CUtensorMap map; bool used; } cache[16];
    static thread_local int next = 0;
    MapKey k{};
    k.base = base; k.box_rows = box_rows; k.row = row; k.rank = (uint32_t)rank;
    for (int i = 0; i < rank; ++i) k.dims[i] = dims[i];
    for (int i = 0; i < rank - 1; ++i) k.strides[i] = strides[i];
    for (auto& e : cache)
        if (e.used && memcmp(&e.key, &k, sizeof(k)) == 0) { *out = e.map; return KVQ_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(KVQ_ERR_UNSUPPORTED, "decode_attn (IMMA variant): cuTensorMapEncodeTiled is not available from this driver");
    cuuint32_t box[5] = {row, box_rows, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = row == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims,
                          (const cuuint64_t*)strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(KVQ_ERR_UNSUPPORTED, "decode_attn (IMMA variant): cuTensorMapEncodeTiled failed (%d) for row=%u", (int)r, row);
    auto& e = cache[next];
    next = (next + 1) % 16;
    e.key = k; e.map = *out; e.used = true;
    return KVQ_OK;
}

template <int KBITS, int VBITS, int G, int D>
int launch_cfg(const AttnParams& P, dim3 grid, cudaStream_t st) {
    using C = MmaCfg<KBITS, VBITS, G, D>;
    CUtensorMap kmap, vmap;
    int rc;
    if (!P.table) {
        const uint64_t kd[4] = {(uint64_t)C::KROW, (uint64_t)P.T, (uint64_t)P.H, (uint64_t)P.B}, ks[3] = {(uint64_t)P.kT, (uint64_t)P.kH, (uint64_t)P.kB};
        const uint64_t vd[4] = {(uint64_t)C::VROW, (uint64_t)P.T, (uint64_t)P.H, (uint64_t)P.B}, vs[3] = {(uint64_t)P.vT, (uint64_t)P.vH, (uint64_t)P.vB};
        if ((rc = make_map(&kmap, P.k, 4, kd, ks, C::KROW, kTile)) != KVQ_OK) return rc;
        if ((rc = make_map(&vmap, P.v, 4, vd, vs, C::VROW, kTile)) != KVQ_OK) return rc;
    } else {
        const uint32_t box_rows = (uint32_t)(P.bs < kTile ? P.bs : kTile);
        const uint64_t nphys = 1ull << 31;   // the ABI does not carry the pool size: the table only holds valid blocks
        const uint64_t kd[5] = {(uint64_t)C::KROW, (uint64_t)P.bs, (uint64_t)P.H, (uint64_t)P.B, nphys}, ks[4] = {(uint64_t)C::KROW, (uint64_t)P.kH, (uint64_t)P.kB, (uint64_t)P.kBlk};
        const uint64_t vd[5] = {(uint64_t)C::VROW, (uint64_t)P.bs, (uint64_t)P.H, (uint64_t)P.B, nphys}, vs[4] = {(uint64_t)C::VROW, (uint64_t)P.vH, (uint64_t)P.vB, (uint64_t)P.vBlk};
        if ((rc = make_map(&kmap, P.k, 5, kd, ks, C::KROW, box_rows)) != KVQ_OK) return rc;
        if ((rc = make_map(&vmap, P.v, 5, vd, vs, C::VROW, box_rows)) != KVQ_OK) return rc;
    }
    auto kern = decode_attn_mma_kernel<KBITS, VBITS, G, D, false>;
    auto kern_paged = decode_attn_mma_kernel<KBITS, VBITS, G, D, true>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    KVQ_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        KVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        KVQ_CUDA(cudaFuncSetAttribute(kern_paged, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        configured_dev = dev;
    }
    if (P.table) kern_paged<<<grid, kMT, C::kSmem, st>>>(kmap, vmap, P);
    else kern<<<grid, kMT, C::kSmem, st>>>(kmap, vmap, P);
    return KVQ_OK;
}

template <int KBITS, int VBITS>
int launch_gd(const AttnParams& P, int G, int D, dim3 grid, cudaStream_t st) {
#define KVQ_MMA_CASE(GG)                                                                     \
    case GG: return D == 128 ? launch_cfg<KBITS, VBITS, GG, 128>(P, grid, st) : launch_cfg<KBITS, VBITS, GG, 64>(P, grid, st);
    switch (G) {
        KVQ_MMA_CASE(1) KVQ_MMA_CASE(2) KVQ_MMA_CASE(4) KVQ_MMA_CASE(8)
    }
#undef KVQ_MMA_CASE
    return fail(KVQ_ERR_UNSUPPORTED, "decode_attn (IMMA variant): Hq/H must be 1, 2, 4 or 8 (got %d)", G);
}

}  // namespace

int attn_mma_ctas_per_sm(int) { return KVQ_ATTN_MMA_CTAS_PER_SM; }

int launch_attn_mma(const AttnParams& P, int k_bits, int v_bits, int G, int D, dim3 grid, cudaStream_t st) {
    if (D != 64 && D != 128) return fail(KVQ_ERR_UNSUPPORTED, "decode_attn (IMMA variant): head_dim must be 64 or 128 (got %d)", D);
    if (k_bits == 8 && v_bits == 8) return launch_gd<8, 8>(P, G, D, grid, st);
    if (k_bits == 8 && v_bits == 4) return launch_gd<8, 4>(P, G, D, grid, st);
    if (k_bits == 4 && v_bits == 4) return launch_gd<4, 4>(P, G, D, grid, st);
    return launch_gd<4, 8>(P, G, D, grid, st);
}

}  // namespace kvq
