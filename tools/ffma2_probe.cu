// Issue / pipe throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, alone and mixed with ALU work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu && tools/ffma2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
    constexpr int N = 16;
    float a[N]; u64 p[N]; uint32_t x[N];
    for (int i = 0; i < N; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = (u64)__float_as_uint(a[i]) * 0x100000001ull; x[i] = threadIdx.x + i; }
    const float m = 1.0001f; const u64 m2 = (u64)__float_as_uint(m) * 0x100000001ull;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) { a[i] = ffma(a[i], m, m); }                                   // FFMA
            if (MODE == 1) { p[i] = ffma2(p[i], m2, m2); }                                // FFMA2
            if (MODE == 2) { p[i] = ffma2(p[i], m2, m2); x[i] = prmt(x[i], x[(i + 1) % N], 0x3210 + i); }  // FFMA2 + PRMT 1:1
            if (MODE == 3) { a[i] = ffma(a[i], m, m); x[i] = prmt(x[i], x[(i + 1) % N], 0x3210 + i); }     // FFMA + PRMT 1:1
            if (MODE == 4) { p[i] = ffma2(p[i], m2, m2); a[i] = ffma(a[i], m, m); }       // FFMA2 + FFMA 1:1
            if (MODE == 5) { p[i] = ffma2(p[i], m2, m2); a[i] = ffma(a[i], m, m); x[i] = prmt(x[i], x[(i + 1) % N], 0x3210 + i); }
            if (MODE == 6) { x[i] = prmt(x[i], x[(i + 1) % N], 0x3210 + i); }             // PRMT
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < N; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int warps_per_sm, double fma_per_iter_elem, double instr_per_iter_elem) {
    const int iters = 4096, threads = warps_per_sm * 32, blocks = 148;
    float* out; long long* cyc; cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    probe<MODE><<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    probe<MODE><<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
    const double elems = (double)iters * 16 * warps_per_sm;  // warp-level loop bodies per SM
    printf("%-28s warps/SM %2d  cycles %9.0f  warp-instr/clk/SM %.2f  fp32 FMA lanes/clk/SM %.1f\n", name, warps_per_sm, c,
           elems * instr_per_iter_elem / c, elems * fma_per_iter_elem * 32 / c);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("FFMA", w, 1, 1); run<1>("FFMA2", w, 2, 1); run<2>("FFMA2+PRMT", w, 2, 2); run<3>("FFMA+PRMT", w, 1, 2);
        run<4>("FFMA2+FFMA", w, 3, 2); run<5>("FFMA2+FFMA+PRMT", w, 3, 3); run<6>("PRMT", w, 0, 1);
    }
    return 0;
}
