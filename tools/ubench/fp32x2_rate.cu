// Micro-benchmark: issue rate of packed fp32x2 (FFMA2/FADD2/FMUL2) vs scalar FFMA on sm_100a, per SM sub-partition.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp32x2_rate tools/ubench/fp32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float2* out, int iters, float2 s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 m = make_float2(1.0001f + s.x, 0.9999f + s.y), c = make_float2(0.001f + s.x, -0.001f + s.y);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], m, c);
            if (MODE == 1) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            if (MODE == 2) a[i] = __fadd2_rn(a[i], c);
            if (MODE == 3) a[i] = __fmul2_rn(a[i], m);
            if (MODE == 4) { a[i] = __ffma2_rn(a[i], m, c); a[i] = __fadd2_rn(a[i], c); }   // mixed
        }
    }
    float2 r = a[0];
    for (int i = 1; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    if (r.x == 12345.678f) out[0] = r;
}

template <int MODE> void run(const char* name, int nthreads, float per_iter_instr) {
    float2* out; cudaMalloc(&out, 64);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, nthreads>>>(out, 100, make_float2(0, 0));
    cudaEventRecord(e0);
    k<MODE><<<148, nthreads>>>(out, iters, make_float2(0, 0));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double warp_instr = (double)iters * per_iter_instr * (nthreads / 32);     // per SM
    double cycles = ms * 1e-3 * clk_khz * 1e3;                                // at nominal max clock
    printf("%-28s threads/SM %4d  %.3f ms  warp-instr/clk/SMSP (at %d MHz) = %.3f\n", name, nthreads, ms, clk_khz / 1000,
           warp_instr / cycles / 4.0);
    cudaFree(out);
}

int main() {
    for (int nt : {128, 256, 512, 1024}) {
        run<0>("FFMA2", nt, 8);
        run<1>("FFMA x2 (scalar)", nt, 16);
        run<2>("FADD2", nt, 8);
        run<3>("FMUL2", nt, 8);
        run<4>("FFMA2+FADD2", nt, 16);
    }
    return 0;
}
