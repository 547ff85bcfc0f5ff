// Microbenchmark: issue/pipe rate of the packed FP32 instructions of sm_100 (add/mul/fma.f32x2)
// against their scalar forms.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE> __global__ void k(float* out, int iters, float s)
{
    const int ILP = 8;
    float a[ILP], b[ILP];
    u64 p[ILP];
    for (int q = 0; q < ILP; ++q) { a[q] = s + q + threadIdx.x; b[q] = s * q; p[q] = pk(a[q], b[q]); }
    const u64 m = pk(1.0000001f, 0.9999999f), c = pk(1e-9f, -1e-9f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) {
            if (MODE == 0) { a[q] = __fmul_rn(a[q], 1.0000001f); a[q] = __fadd_rn(a[q], 1e-9f); b[q] = __fmul_rn(b[q], 0.9999999f); b[q] = __fadd_rn(b[q], -1e-9f); }
            if (MODE == 1) { p[q] = mul2(p[q], m); p[q] = add2(p[q], c); }
            if (MODE == 2) { a[q] = __fmaf_rn(a[q], 1.0000001f, 1e-9f); b[q] = __fmaf_rn(b[q], 0.9999999f, -1e-9f); }
            if (MODE == 3) { p[q] = fma2(p[q], m, c); }
            // MODE 4: packed FP + integer work on the freed issue slots
            if (MODE == 4) { p[q] = mul2(p[q], m); p[q] = add2(p[q], c); a[q] = __int_as_float((__float_as_int(a[q]) ^ it) + q); b[q] = __int_as_float(__float_as_int(b[q]) * 3 + it); }
            if (MODE == 5) { a[q] = __fmul_rn(a[q], 1.0000001f); a[q] = __fadd_rn(a[q], 1e-9f); b[q] = __int_as_float((__float_as_int(b[q]) ^ it) + q); }
        }
    }
    float r = 0;
    for (int q = 0; q < ILP; ++q) r += a[q] + b[q] + __uint_as_float((unsigned)p[q]) + __uint_as_float((unsigned)(p[q] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, double laneops_per_iter_thread)
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = laneops_per_iter_thread * iters * 148.0 * 8 * 256;
    printf("%-40s %8.3f ms  %7.2f T lane-op/s  (err %s)\n", name, ms, ops / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main()
{
    run<0>("scalar FMUL+FADD", 8 * 4);
    run<1>("packed mul.f32x2+add.f32x2", 8 * 4);
    run<2>("scalar FFMA (ops=1 per fma)", 8 * 2);
    run<3>("packed fma.f32x2 (ops=1 per fma lane)", 8 * 2);
    run<4>("packed mul2+add2 + 2 int ops (fp ops counted)", 8 * 4);
    run<5>("scalar mul+add + 1 int (fp ops counted)", 8 * 2);
    return 0;
}
