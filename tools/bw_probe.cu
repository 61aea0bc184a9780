// bw_probe.cu -- what HBM3e on this B200 sustains for read-only / write-only / mixed streams, per store policy.
// Development probe behind the roofline notes in DESIGN.md (nvcc -arch=sm_100a tools/bw_probe.cu -o bw_probe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __device__ __forceinline__ void st16(uint4* p, uint4 v) {
    if (MODE == 0) *p = v;
    else if (MODE == 1) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// RD reads per WR writes of 16 B (per thread per iteration); RD or WR may be 0
template <int RD, int WR, int MODE>
__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n_units, uint32_t* sink) {
    const int64_t step = (int64_t)gridDim.x * 256;
    uint32_t acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_units; i += step * 4) {
        uint4 v[4][RD > 0 ? RD : 1];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i + u * step;
            if (j < n_units) {
#pragma unroll
                for (int r = 0; r < RD; ++r) v[u][r] = ld16(in + j * RD + r * 0 + (int64_t)r * n_units * 0 + 0) , v[u][r] = ld16(in + (int64_t)r * n_units + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i + u * step;
            if (j < n_units) {
                uint4 x = make_uint4(j, 1, 2, 3);
#pragma unroll
                for (int r = 0; r < RD; ++r) { x.x ^= v[u][r].x; x.y += v[u][r].y; x.z ^= v[u][r].z; x.w += v[u][r].w; }
                if (WR == 0) acc += x.x + x.y + x.z + x.w;
#pragma unroll
                for (int w = 0; w < WR; ++w) st16<MODE>(out + (int64_t)w * n_units + j, x);
            }
        }
    }
    if (WR == 0 && acc == 0x12345678u) *sink = acc;
}

template <int RD, int WR, int MODE>
int run(const char* name, const uint4* in, uint4* out, int64_t n_units, uint32_t* sink, int grid) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) stream_kernel<RD, WR, MODE><<<grid, 256>>>(in, out, n_units, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e9, sum = 0;
    const int iters = 10;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a));
        stream_kernel<RD, WR, MODE><<<grid, 256>>>(in, out, n_units, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best; sum += ms;
    }
    const double bytes = (double)n_units * 16 * (RD + WR);
    printf("%-44s rd=%d wr=%d  avg %7.3f ms  %7.1f GB/s  (best %7.1f)\n", name, RD, WR, sum / iters, bytes / (sum / iters) / 1e6, bytes / best / 1e6);
    return 0;
}

int main() {
    const int64_t n_units = (int64_t)1 << 26;  // 1 GiB per stream of 16-byte units
    uint4 *in, *out; uint32_t* sink;
    CK(cudaMalloc(&in, n_units * 16 * 4)); CK(cudaMalloc(&out, n_units * 16 * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, n_units * 16 * 4)); CK(cudaMemset(out, 0, n_units * 16 * 4));
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int g : {sms * 8, sms * 16, sms * 32}) {
        printf("grid %d\n", g);
        run<1, 0, 0>("read only", in, out, n_units, sink, g);
        run<0, 1, 0>("write only (st default)", in, out, n_units, sink, g);
        run<0, 1, 1>("write only (L1::no_allocate)", in, out, n_units, sink, g);
        run<0, 1, 2>("write only (.cs)", in, out, n_units, sink, g);
        run<0, 1, 3>("write only (.wt)", in, out, n_units, sink, g);
        run<1, 1, 0>("copy 1:1 (st default)", in, out, n_units, sink, g);
        run<1, 1, 1>("copy 1:1 (no_allocate)", in, out, n_units, sink, g);
        run<1, 2, 1>("1 read : 2 write (int8->bf16 mix)", in, out, n_units, sink, g);
        run<1, 4, 1>("1 read : 4 write (int4->bf16 mix)", in, out, n_units, sink, g);
        run<1, 4, 2>("1 read : 4 write (.cs)", in, out, n_units, sink, g);
        run<2, 1, 1>("2 read : 1 write (bf16->int8 mix)", in, out, n_units, sink, g);
        run<4, 1, 1>("4 read : 1 write (bf16->int4 mix)", in, out, n_units, sink, g);
    }
    return 0;
}
