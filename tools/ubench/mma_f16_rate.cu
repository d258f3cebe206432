// Micro-benchmark: issue rate of legacy warp-level mma.sync.m16n8k16 f16 (fp32 accumulate) on sm_100a (per SM sub-partition), alone and
// interleaved with packed FFMA2 (do they overlap?).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/mma_f16_rate tools/ubench/mma_f16_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_f16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters) {
    float d[6][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
    for (int i = 0; i < 2; ++i) b[i] = 0x38003800u + i;
    for (int j = 0; j < 6; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    float2 f[6];
    for (int j = 0; j < 6; ++j) f[j] = make_float2(j, threadIdx.x);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (MODE != 1) mma_f16(d[j], a, b);
            if (MODE != 0) { f[j] = __ffma2_rn(f[j], m, c); f[j] = __ffma2_rn(f[j], m, c); }
        }
    }
    float r = 0.f;
    for (int j = 0; j < 6; ++j) { for (int i = 0; i < 4; ++i) r += d[j][i]; r += f[j].x + f[j].y; }
    if (r == 12345.678f) out[0] = r;
}

template <int MODE> void run(const char* name, int nthreads) {
    float* out; cudaMalloc(&out, 64);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, nthreads>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148, nthreads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double per_smsp = (double)iters * 6 * (nthreads / 32) / 4.0;     // mma (MODE 0/2) or FFMA2 pairs (MODE 1) per SMSP
    printf("%-22s threads/SM %4d  %.3f ms  cycles per (mma | 2xFFMA2) per SMSP = %.2f\n", name, nthreads, ms, cycles / per_smsp);
    cudaFree(out);
}

int main() {
    for (int nt : {128, 256, 512, 1024}) {
        run<0>("mma.m16n8k16.f16", nt);
        run<1>("2x FFMA2", nt);
        run<2>("mma + 2x FFMA2", nt);
    }
    return 0;
}
