// xchg.cu -- the one collective on the path, without NCCL's rendezvous: an all-to-all max of the absmax words through
// peer-mapped (symmetric) memory over NVLink (sm_100a).
//
// A head-sharded cache exchanges exactly one thing: the absmax bit patterns (uint32; integer max == float max) between
// phase 1 (absmax) and phase 2 (quantize) -- reference scale over all heads, src/quantization/ops.py:27,:48,:330,:342.  NCCL
// does that as an allreduce(MAX) of 4 .. 256 bytes: correct, but each call costs a host-side enqueue, a proxy rendezvous
// and its own kernel (~20-30 us at the pace of the slowest rank, twice per step in the overlapped form of r1).  Here every
// rank owns a small buffer that all peers can write (torch symmetric memory: CUDA VMM handles mapped into every process);
// ONE tiny kernel per rank
//     1. stores its n words into slot [parity][rank] of EVERY rank's buffer (direct NVLink stores),
//     2. publishes a sequence number per peer (st.release.sys after a system fence),
//     3. spins until all peers' sequence numbers for this step have arrived (ld.acquire.sys),
//     4. max-combines the world's words into the caller's array.
// MAX is exact and order-independent: the combined words -- hence scales and codes -- are bit-identical to the 1-GPU run and
// to the NCCL path.  Two parities: a fast rank may already write step s + 1 while a slow one still reads step s; it cannot be
// two steps ahead because step s + 1 completes only with everyone's s + 1 flags.  The spin is bounded (a peer that never
// arrives traps after ~2 s instead of hanging the GPU).  The kernels of different ranks run on different GPUs.
#include "kvq_common.cuh"

namespace kvq {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) absmax_exchange_kernel(uint32_t* __restrict__ bits, int n, uint32_t* const* __restrict__ peers, int rank,
                                                             int world, uint32_t seq) {
    const int parity = (int)(seq & 1u);
    // 1. my words into every rank's buffer (including my own)
    for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
        const int p = i / n, j = i - p * n;
        peers[p][xch_word_off(parity, rank, j)] = bits[j];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, 3. wait for everyone
    if (threadIdx.x < world) {
        st_release_sys(peers[threadIdx.x] + xch_flag_off(parity, rank), seq);
        const uint32_t* mine = peers[rank] + xch_flag_off(parity, threadIdx.x);
        unsigned long long spins = 0;
        while (ld_acquire_sys(mine) != seq) {
            if (++spins > (1ull << 31)) __trap();
        }
    }
    __syncthreads();
    // 4. combine
    const uint32_t* buf = peers[rank];
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        uint32_t m = 0;
        for (int p = 0; p < world; ++p) m = max(m, __ldcv(buf + xch_word_off(parity, p, j)));
        bits[j] = m;
    }
}

}  // namespace
}  // namespace kvq

using namespace kvq;

extern "C" int64_t kvq_absmax_exchange_buffer_bytes(void) { return (int64_t)kXchBufferBytes; }

extern "C" int kvq_absmax_exchange(uint32_t* absmax_bits, int64_t n_words, void* const* peer_buffers, int rank, int world_size, uint32_t seq,
                                   kvq_stream_t stream) {
    KVQ_REQUIRE(absmax_bits && peer_buffers, "null pointer");
    KVQ_REQUIRE(n_words >= 1 && n_words <= kXchMaxWords, "absmax_exchange: 1 .. %d words per call (got %lld)", kXchMaxWords, (long long)n_words);
    KVQ_REQUIRE(world_size >= 1 && world_size <= kXchMaxWorld && rank >= 0 && rank < world_size, "bad rank / world size");
    KVQ_REQUIRE(seq != 0, "seq must be a non-zero, increasing step counter (buffers start zeroed)");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    KVQ_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
    if (cs != cudaStreamCaptureStatusNone)  // a replayed graph would present the same step number again and accept stale words
        return fail(KVQ_ERR_UNSUPPORTED, "absmax_exchange: not capturable into a CUDA graph (the step counter is a launch argument)");
    absmax_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(absmax_bits, (int)n_words, reinterpret_cast<uint32_t* const*>(peer_buffers), rank,
                                                               world_size, seq);
    return check_launch("absmax_exchange");
}
