// Micro-benchmark: issue rates of FFMA2 / FFMA / FMNMX / IMAD.U32 / LOP3 / LDS.32 on sm_100a (per SMSP, warp-instr per clk).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float x) {
    u64 a[8]; float f[16]; unsigned u[8];
    for (int i = 0; i < 8; ++i) { a[i] = ((u64)__float_as_uint(x + i) << 32) | __float_as_uint(x - i); u[i] = threadIdx.x * 7 + i; }
    for (int i = 0; i < 16; ++i) f[i] = x + i;
    u64 b = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    u64 c = ((u64)__float_as_uint(0.001f) << 32) | __float_as_uint(-0.001f);
    __shared__ unsigned sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {        // 16 FFMA2 (8 chains x 2)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b, c);
        } else if (MODE == 1) { // 16 FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(x), "f"(0.5f * x));
        } else if (MODE == 2) { // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(x));
        } else if (MODE == 3) { // 8 FFMA2 + 8 FMNMX interleaved
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fma2(a[i], b, c); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(x)); }
        } else if (MODE == 4) { // 16 shl (unpack)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
        } else if (MODE == 5) { // 16 LDS.32
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = sm[u[i] & 1023];
        } else if (MODE == 6) { // 8 FFMA2 + 8 IMAD-style unpack (mul.lo by 65536)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fma2(a[i], b, c); asm volatile("mul.lo.u32 %0, %0, 65537;" : "+r"(u[i])); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { s += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32)) + (float)u[i]; }
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 12345.678f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 + MODE] = (float)(t1 - t0);
}
int main() {
    float* d; cudaMalloc(&d, 64 * 4); cudaMemset(d, 0, 256);
    const int iters = 4096;
    const char* names[] = {"16xFFMA2", "16xFFMA", "16xFMNMX", "8xFFMA2+8xFMNMX", "16xSHL", "16xLDS32", "8xFFMA2+8xIMUL"};
    for (int warps = 4; warps <= 16; warps *= 2) {
#define RUN(M) { k<M><<<148, 32 * warps>>>(d, iters, 1.0f); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("mode %d failed: %s\n", M, cudaGetErrorString(e)); return 1; } }
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(6)
        float h[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        for (int m = 0; m < 7; ++m) { if (m == 5) continue;
            const double cyc = h[1 + m];
            const double winstr_per_smsp = 16.0 * iters * (warps / 4.0);
            printf("warps/SM=%2d %-18s cycles=%9.0f  -> %.3f warp-instr/clk/SMSP\n", warps, names[m], cyc, winstr_per_smsp / cyc);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
