// Microbenchmark + numerical check of the tensor-core form of SimpleFilter's first layer (16 -> 24, DockingModels.py:29)
// exactly as the epilogue kernel runs it: one thread per translation of a 2x2x160 tile, two rotations per complex value,
// A operand staged in TMEM (tcgen05.st, 3xTF32 split), weights in shared memory, tcgen05.mma.kind::tf32, tcgen05.ld.
//   MODE 0: the FP32 FFMA2 chain (what the round-1 kernel does)      MODE 1: tcgen05 / TMEM
// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/tc_filter tools/ubench/tc_filter.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../deeplocalproteindocking_b200/csrc/dlpd_tc.h"

constexpr int C0 = 16, H = 24, XY = 4, NZ = 160, NL0 = 68, NVOX = XY * NZ, HB = (NZ / 2) * H, NPAD = 32, NT = 640;

struct Weights {
    float W1t[C0 * H];   // [c][j]
    float W2[H];
    float b2;
};

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// MODE 0: FP32 FFMA2 chain.  MODE 1: 3xTF32, the two rotations of a thread's translation share one A staging area (sequential).
// MODE 2: TF32 hi (round to nearest) + BF16 lo, both rotations staged at once: D = lo w(bf16) + hi w_lo + hi w_hi, N = 24.
// ABL (timing only, wrong results): bit 0 = no MMA / no waits, bit 1 = no tcgen05.ld, bit 2 = no tcgen05.st
template <int MODE, int NACC, int ABL>
__global__ void __launch_bounds__(NT, 1) filter_kernel(const float2* __restrict__ tiles, const float2* __restrict__ h1g,
                                                        const __grid_constant__ Weights w, float2* __restrict__ out, int ntiles,
                                                        float* __restrict__ sink) {
    extern __shared__ __align__(1024) char smem[];
    float2* b0 = reinterpret_cast<float2*>(smem);                    // [2][NZ * NL0]
    float2* hb = b0 + 2 * NZ * NL0;                                  // [2][HB]
    float* Bw = reinterpret_cast<float*>(hb + 2 * HB);               // [2 (hi, lo)][4 kc][NPAD][4] tf32
    unsigned short* Bh = reinterpret_cast<unsigned short*>(Bw + 2 * 4 * NPAD * 4);   // [2 kc][NPAD][8] bf16
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bh + 2 * NPAD * 8);       // [5][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * NZ * NL0; i += NT) b0[i] = tiles[i];
    for (int i = tid; i < 2 * HB; i += NT) hb[i] = h1g[i];
    uint32_t tmem_base = 0;
    if (MODE >= 1) {
        for (int i = tid; i < 2 * 4 * NPAD * 4; i += NT) {
            const int e = i & 3, n = (i >> 2) % NPAD, kc = (i / (4 * NPAD)) & 3, part = i / (16 * NPAD);
            const int c = 4 * kc + e;
            const float v = n < H ? w.W1t[c * H + n] : 0.0f;
            const float hi = tc::tf32_hi(v);
            Bw[i] = part == 0 ? hi : (v - hi);
        }
        for (int i = tid; i < 2 * NPAD * 8; i += NT) {
            const int e = i & 7, n = (i >> 3) % NPAD, kc = i / (8 * NPAD);
            const int c = 8 * kc + e;
            const float v = n < H ? w.W1t[c * H + n] : 0.0f;
            unsigned short hbits;
            asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(hbits) : "f"(v));
            Bh[i] = hbits;
        }
        if (tid < 32) tc::tmem_alloc(tmem_slot, 512);
        if (tid == 0) {
            for (int i = 0; i < 10; ++i) tc::mbar_init(bars + i, 1);
            tc::mbar_fence_init();
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
    }
    __syncthreads();
    if (MODE >= 1) {
        tc::fence_after_sync();
        tmem_base = *tmem_slot;
    }
    const int v = tid, xy = v % XY, z = v / XY;
    const int g = tid >> 7, q = (tid >> 5) & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t bhi = tc::smem_u32(Bw), blo = bhi + 4 * NPAD * 16, bbf = tc::smem_u32(Bh);
    constexpr uint32_t IDESC = tc::idesc_tf32(128, NACC);
    constexpr uint32_t IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NACC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;
    float2 running = make_float2(1e30f, 1e30f);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int slot = tile & 1;
        __syncthreads();
        const float2* v0 = b0 + slot * NZ * NL0 + z * NL0 + xy;
        const float2* hz = hb + slot * HB + (z >> 1) * H;
        float2 sc;
        if (MODE == 0) {
            float2 h[H];
#pragma unroll
            for (int j = 0; j < H; ++j) h[j] = hz[j];
#pragma unroll
            for (int c = 0; c < C0; ++c) {
                const float2 x = v0[c * XY];
#pragma unroll
                for (int j = 0; j < H; ++j) h[j] = __ffma2_rn(make_float2(w.W1t[c * H + j], w.W1t[c * H + j]), x, h[j]);
            }
            sc = make_float2(w.b2, w.b2);
#pragma unroll
            for (int j = 0; j < H; ++j)
                sc = __ffma2_rn(make_float2(w.W2[j], w.W2[j]), make_float2(fmaxf(h[j].x, 0.f), fmaxf(h[j].y, 0.f)), sc);
        } else if (MODE == 1) {
            const uint32_t colA = g * 96, colDA = colA + 32, colDB = colA + 64;
            float2 x[C0];
#pragma unroll
            for (int c = 0; c < C0; ++c) x[c] = v0[c * XY];
#pragma unroll
            for (int rot = 0; rot < 2; ++rot) {
                uint32_t hi[C0], lo[C0];
#pragma unroll
                for (int c = 0; c < C0; ++c) {
                    const float a = rot ? x[c].y : x[c].x, ah = tc::tf32_hi(a);
                    hi[c] = __float_as_uint(ah);
                    lo[c] = __float_as_uint(a - ah);
                }
                if (rot == 1 && !(ABL & 1)) {            // the A area is free once the MMAs of rotation A have completed
                    tc::mbar_wait(bars + 2 * g, phase);
                    tc::fence_after_sync();
                }
                if (!(ABL & 4)) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t t0[8], t1[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { t0[i] = hi[8 * half + i]; t1[i] = lo[8 * half + i]; }
                        tc::st8(lane_addr + colA + 8 * half, t0);
                        tc::st8(lane_addr + colA + 16 + 8 * half, t1);
                    }
                    tc::wait_st();
                }
                tc::fence_before_sync();
                bar_sync(1 + g, 128);
                if (!(ABL & 1) && q == 0 && tc::elect_one()) {
                    tc::fence_after_sync();
                    const uint32_t colD = rot ? colDB : colDA;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint64_t dhi = tc::smem_desc_kmajor(bhi + s * 2 * NPAD * 16, NPAD * 16, 128);
                        const uint64_t dlo = tc::smem_desc_kmajor(blo + s * 2 * NPAD * 16, NPAD * 16, 128);
                        tc::mma_tf32_ts(tmem_base + colD, tmem_base + colA + 16 + 8 * s, dhi, IDESC, s > 0);   // a_lo w_hi
                        tc::mma_tf32_ts(tmem_base + colD, tmem_base + colA + 8 * s, dlo, IDESC, 1);             // a_hi w_lo
                        tc::mma_tf32_ts(tmem_base + colD, tmem_base + colA + 8 * s, dhi, IDESC, 1);             // a_hi w_hi
                    }
                    tc::mma_commit(bars + 2 * g + rot);
                }
                __syncwarp();
            }
            float acc[2][H];
#pragma unroll
            for (int rot = 0; rot < 2; ++rot) {
                if (rot == 1 && !(ABL & 1)) {
                    tc::mbar_wait(bars + 2 * g + 1, phase);
                    tc::fence_after_sync();
                }
                uint32_t r[3][8];
                if (!(ABL & 2)) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) tc::ld8(lane_addr + (rot ? colDB : colDA) + 8 * t, r[t]);
                    tc::wait_ld();
                } else {
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int i = 0; i < 8; ++i) r[t][i] = __float_as_uint(x[(8 * t + i) & 15].x);
                }
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[rot][8 * t + i] = __uint_as_float(r[t][i]);
            }
            sc = make_float2(w.b2, w.b2);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 hh = __fadd2_rn(make_float2(acc[0][j], acc[1][j]), hz[j]);
                sc = __ffma2_rn(make_float2(w.W2[j], w.W2[j]), make_float2(fmaxf(hh.x, 0.f), fmaxf(hh.y, 0.f)), sc);
            }
            phase ^= 1;
        } else {
            // per group: A_rotA [hi 16 | lo 8], A_rotB [24], D_rotA [24], D_rotB [24]
            const uint32_t colA = g * 96;
            float2 x[C0];
#pragma unroll
            for (int c = 0; c < C0; ++c) x[c] = v0[c * XY];
#pragma unroll
            for (int rot = 0; rot < 2; ++rot) {
                uint32_t hi[C0], lp[C0 / 2];
                float lo[C0];
#pragma unroll
                for (int c = 0; c < C0; ++c) {
                    const float a = rot ? x[c].y : x[c].x;
                    hi[c] = (__float_as_uint(a) + 0x1000u) & 0xffffe000u;      // TF32, round to nearest
                    lo[c] = a - __uint_as_float(hi[c]);
                }
#pragma unroll
                for (int c = 0; c < C0; c += 2) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp[c / 2]) : "f"(lo[c + 1]), "f"(lo[c]));
                if (!(ABL & 4)) {
                    uint32_t t0[8], t1[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { t0[i] = hi[i]; t1[i] = hi[8 + i]; }
                    tc::st8(lane_addr + colA + 24 * rot, t0);
                    tc::st8(lane_addr + colA + 24 * rot + 8, t1);
                    tc::st8(lane_addr + colA + 24 * rot + 16, lp);
                }
            }
            if (!(ABL & 4)) tc::wait_st();
            tc::fence_before_sync();
            bar_sync(1 + g, 128);
            if (!(ABL & 1) && q == 0 && tc::elect_one()) {
                tc::fence_after_sync();
#pragma unroll
                for (int rot = 0; rot < 2; ++rot) {
                    const uint32_t a0 = tmem_base + colA + 24 * rot, d0 = tmem_base + colA + 48 + 24 * rot;
                    const uint64_t dbf = tc::smem_desc_kmajor(bbf, NPAD * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d0), "r"(a0 + 16), "l"(dbf),
                        "r"(IDESC_BF), "r"(0u)
                        : "memory");                                                                          // a_lo w
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint64_t dhi = tc::smem_desc_kmajor(bhi + s * 2 * NPAD * 16, NPAD * 16, 128);
                        const uint64_t dlo = tc::smem_desc_kmajor(blo + s * 2 * NPAD * 16, NPAD * 16, 128);
                        tc::mma_tf32_ts(d0, a0 + 8 * s, dlo, IDESC, 1);             // a_hi w_lo
                        tc::mma_tf32_ts(d0, a0 + 8 * s, dhi, IDESC, 1);             // a_hi w_hi
                    }
                    tc::mma_commit(bars + 2 * g + rot);
                }
            }
            __syncwarp();
            float acc[2][H];
#pragma unroll
            for (int rot = 0; rot < 2; ++rot) {
                if (!(ABL & 1)) {
                    tc::mbar_wait(bars + 2 * g + rot, phase);
                    tc::fence_after_sync();
                }
                uint32_t r[3][8];
                if (!(ABL & 2)) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) tc::ld8(lane_addr + colA + 48 + 24 * rot + 8 * t, r[t]);
                    tc::wait_ld();
                } else {
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int i = 0; i < 8; ++i) r[t][i] = __float_as_uint(x[(8 * t + i) & 15].x);
                }
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[rot][8 * t + i] = __uint_as_float(r[t][i]);
            }
            sc = make_float2(w.b2, w.b2);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 hh = __fadd2_rn(make_float2(acc[0][j], acc[1][j]), hz[j]);
                sc = __ffma2_rn(make_float2(w.W2[j], w.W2[j]), make_float2(fmaxf(hh.x, 0.f), fmaxf(hh.y, 0.f)), sc);
            }
            phase ^= 1;
        }
        if (tile < 2 && blockIdx.x == 0) out[tile * NVOX + v] = sc;
        running.x = fminf(running.x, sc.x);
        running.y = fminf(running.y, sc.y);
    }
    if (running.x + running.y == 1.2345e-30f) sink[0] = running.x;
    if (MODE >= 1) {
        tc::fence_before_sync();
        __syncthreads();
        if (tid < 32) tc::tmem_dealloc(tmem_base, 512);
    }
}

static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }
static float nrand() {
    float s = 0.f;
    for (int i = 0; i < 6; ++i) s += frand();
    return s;
}

template <int MODE, int NACC, int ABL = 0>
static int run(const char* name, const float2* d_tiles, const float2* d_h1, const Weights& w, const std::vector<float2>& tiles,
               const std::vector<float2>& h1, int ntiles) {
    const size_t smem = sizeof(float2) * (2 * NZ * NL0 + 2 * HB) + 2 * 4 * NPAD * 16 + 2 * NPAD * 16 + 10 * 8 + 16;
    cudaFuncSetAttribute(filter_kernel<MODE, NACC, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float2* d_out;
    float* d_sink;
    cudaMalloc(&d_out, sizeof(float2) * 2 * NVOX);
    cudaMalloc(&d_sink, 4);
    cudaMemset(d_out, 0, sizeof(float2) * 2 * NVOX);
    filter_kernel<MODE, NACC, ABL><<<148, NT, smem>>>(d_tiles, d_h1, w, d_out, 4, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: FAILED %s\n", name, cudaGetErrorString(e));
        return 1;
    }
    std::vector<float2> out(2 * NVOX);
    cudaMemcpy(out.data(), d_out, sizeof(float2) * 2 * NVOX, cudaMemcpyDeviceToHost);
    double maxref = 0, maxerr = 0, maxrel = 0;
    for (int tile = 0; tile < 2; ++tile)
        for (int v = 0; v < NVOX; ++v)
            for (int rot = 0; rot < 2; ++rot) {
                const int xy = v % XY, z = v / XY;
                double s = w.b2;
                for (int j = 0; j < H; ++j) {
                    const float2 hv = h1[tile * HB + (z >> 1) * H + j];
                    double h = rot ? hv.y : hv.x;
                    for (int c = 0; c < C0; ++c) {
                        const float2 a = tiles[tile * NZ * NL0 + z * NL0 + c * XY + xy];
                        h += (double)w.W1t[c * H + j] * (double)(rot ? a.y : a.x);
                    }
                    s += (double)w.W2[j] * (h > 0 ? h : 0);
                }
                const float2 o = out[tile * NVOX + v];
                const double got = rot ? o.y : o.x, err = fabs(got - s);
                maxref = fmax(maxref, fabs(s));
                maxerr = fmax(maxerr, err);
                if (fabs(s) > 1e-3) maxrel = fmax(maxrel, err / fabs(s));
            }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    filter_kernel<MODE, NACC, ABL><<<148, NT, smem>>>(d_tiles, d_h1, w, d_out, ntiles, d_sink);
    cudaEventRecord(e0);
    filter_kernel<MODE, NACC, ABL><<<148, NT, smem>>>(d_tiles, d_h1, w, d_out, ntiles, d_sink);
    cudaEventRecord(e1);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: FAILED (timing) %s\n", name, cudaGetErrorString(e));
        return 1;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: max|ref| %.4g  max abs err %.3g (%.3g of range)  max rel err %.3g   %.3f us/tile (%d tiles/CTA, %.3f ms)\n", name, maxref,
           maxerr, maxerr / maxref, maxrel, 1e3 * ms / ntiles, ntiles, ms);
    cudaFree(d_out);
    cudaFree(d_sink);
    return 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    const float scale = argc > 2 ? (float)atof(argv[2]) : 50.f;
    srand(1234);
    std::vector<float2> tiles(2 * NZ * NL0), h1(2 * HB);
    for (auto& t : tiles) t = make_float2(scale * nrand(), scale * nrand());
    for (auto& t : h1) t = make_float2(scale * nrand(), scale * nrand());
    Weights w;
    for (int i = 0; i < C0 * H; ++i) w.W1t[i] = 0.29f * frand();
    for (int j = 0; j < H; ++j) w.W2[j] = 0.49f * frand();
    w.b2 = 0.1f;
    float2 *d_tiles, *d_h1;
    cudaMalloc(&d_tiles, sizeof(float2) * tiles.size());
    cudaMalloc(&d_h1, sizeof(float2) * h1.size());
    cudaMemcpy(d_tiles, tiles.data(), sizeof(float2) * tiles.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_h1, h1.data(), sizeof(float2) * h1.size(), cudaMemcpyHostToDevice);
    const int ntiles = 700;
    if (which == 0) return run<0, 32>("fp32 FFMA2 chain  ", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 1) return run<1, 32>("tcgen05 3xTF32 N=32", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 2) return run<1, 24>("tcgen05 3xTF32 N=24", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 3) return run<2, 24>("tcgen05 TF32 hi + BF16 lo, N=24", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 11) return run<1, 24, 1>("ablate: 3xTF32 no MMA / waits", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 12) return run<1, 24, 2>("ablate: 3xTF32 no tcgen05.ld", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 14) return run<1, 24, 4>("ablate: 3xTF32 no tcgen05.st", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 17) return run<1, 24, 7>("ablate: 3xTF32 none of the three", d_tiles, d_h1, w, tiles, h1, ntiles);
    if (which == 21) return run<2, 24, 1>("ablate: hi+bf16 no MMA / waits", d_tiles, d_h1, w, tiles, h1, ntiles);
    return 2;
}
