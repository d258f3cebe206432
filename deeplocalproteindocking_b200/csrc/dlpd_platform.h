// Platform glue: the kernel BODIES in dlpd_kernels.h are written once against the small set of
// macros below.  The product build (nvcc, sm_100a) maps them to CUDA; the emulation build
// (g++ -DDLPD_EMU, tests/emu only -- TEST INFRASTRUCTURE, never loaded by the package) runs every
// "parallel for over tasks" sequentially on the host so that index math, twiddles and layouts can
// be checked against the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <type_traits>

#if defined(DLPD_EMU)
  #define DLPD_HD inline
  #define DLPD_DEV inline
  #define DLPD_CE constexpr
  struct float2 { float x, y; };
  struct float4 { float x, y, z, w; };
  static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
  static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
  static inline float __ldg(const float* p) { return *p; }
  static inline float2 __ldg(const float2* p) { return *p; }
  static inline float4 __ldg(const float4* p) { return *p; }
  static inline float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
  // packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100)
  static inline float2 ffma2_(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
  static inline float2 fadd2_(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
  static inline float2 fmul2_(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
  // one emulated "thread": tasks run in order
  #define DLPD_FOR_TASKS(t, n, NT) for (int t = 0; t < (n); ++t)
  #define DLPD_SYNC() do {} while (0)
  // asynchronous global->shared copies (cp.async on the device)
  static inline void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
  static inline void cp_async8(void* dst, const void* src) { memcpy(dst, src, 8); }
  static inline void cp_async_wait_all() {}
  #define DLPD_UNROLL
  #define DLPD_RESTRICT
  typedef void* dlpd_stream_t;
#else
  #include <cuda_runtime.h>
  #define DLPD_HD __host__ __device__ __forceinline__
  #define DLPD_DEV __device__ __forceinline__
  #define DLPD_CE __host__ __device__ constexpr
  #define DLPD_FOR_TASKS(t, n, NT) for (int t = threadIdx.x; t < (n); t += (NT))
  #define DLPD_SYNC() __syncthreads()
  #define DLPD_UNROLL _Pragma("unroll")
  #define DLPD_RESTRICT __restrict__
  DLPD_DEV float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
  // packed fp32x2 arithmetic: one FFMA2 / FADD2 instruction does two fp32 operations per lane (Blackwell)
  DLPD_DEV float2 ffma2_(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
  DLPD_DEV float2 fadd2_(float2 a, float2 b) { return __fadd2_rn(a, b); }
  DLPD_DEV float2 fmul2_(float2 a, float2 b) { return __fmul2_rn(a, b); }
  // asynchronous global->shared copies: no register staging, every copy of a thread is in flight at once
  DLPD_DEV void cp_async16(void* dst, const void* src) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
  DLPD_DEV void cp_async8(void* dst, const void* src) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  }
  DLPD_DEV void cp_async_wait_all() {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  typedef cudaStream_t dlpd_stream_t;
#endif

typedef float2 cpx;
