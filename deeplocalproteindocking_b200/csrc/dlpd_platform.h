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
  // tasks for the SUB threads [off, off + SUB) of the CTA only
  #define DLPD_FOR_TASKS_SUB(t, n, SUB, off) for (int t = 0; t < (n); ++t)
  #define DLPD_SYNC() do {} while (0)
  // asynchronous global->shared copies (cp.async on the device)
  static inline void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
  static inline void cp_async8(void* dst, const void* src) { memcpy(dst, src, 8); }
  static inline void cp_async_wait_all() {}
  static inline void cp_async_commit() {}
  static inline void cp_async_wait_pending() {}
  static inline void cp_async_wait_allow1() {}
  static inline void prefetch_l2(const void*) {}
  // teams: the emulation runs the teams of a CTA one after the other (DLPD_FOR_TEAMS), each as one "thread"
  #define DLPD_FOR_TEAMS(g, G, TNT) for (int g = 0; g < (G); ++g)
  #define DLPD_TEAM_TID(TNT) 0
  static inline void named_barrier_sync(int, int) {}
  static inline void named_barrier_arrive(int, int) {}
  static inline void store_cpx2(float2* dst, float2 a, float2 b) { dst[0] = a; dst[1] = b; }
  static inline void cluster_arrive() {}
  static inline void cluster_arrive_relaxed() {}
  static inline void cluster_wait() {}
  static inline unsigned int atomic_inc_u32(unsigned int* p) { return (*p)++; }
  static inline void atomic_add_f32(float* p, float v) { *p += v; }
  #define DLPD_UNROLL
  #define DLPD_RESTRICT
  typedef void* dlpd_stream_t;
#else
  #include <cuda_runtime.h>
  #define DLPD_HD __host__ __device__ __forceinline__
  #define DLPD_DEV __device__ __forceinline__
  #define DLPD_CE __host__ __device__ constexpr
  #define DLPD_FOR_TASKS(t, n, NT) for (int t = threadIdx.x; t < (n); t += (NT))
  #define DLPD_FOR_TASKS_SUB(t, n, SUB, off) for (int t = (int)threadIdx.x - (off); t >= 0 && t < (n); t += (SUB))
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
  DLPD_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  DLPD_DEV void cp_async_wait_pending() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  DLPD_DEV void cp_async_wait_allow1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }   // all but the newest group
  DLPD_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
  // teams: G contiguous groups of TNT threads, each working on its own sub-problem with its own named barrier
  #define DLPD_FOR_TEAMS(g, G, TNT) for (int g = threadIdx.x / (TNT), g##_once = 1; g##_once; g##_once = 0)
  #define DLPD_TEAM_TID(TNT) ((int)(threadIdx.x % (TNT)))
  DLPD_DEV void named_barrier_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
  DLPD_DEV void named_barrier_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
  DLPD_DEV void store_cpx2(float2* dst, float2 a, float2 b) {       // 16-byte aligned destination
      *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
  }
  // split cluster barrier: arrive (release) when this CTA no longer reads its peers' shared memory, wait (acquire) only
  // right before it overwrites its own
  DLPD_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
  // arrive without memory ordering: for a barrier that only says "I have finished READING my peers' shared memory" (the values are
  // already in registers), so that the arrive does not wait for this thread's outstanding global stores
  DLPD_DEV void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
  DLPD_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  DLPD_DEV unsigned int atomic_inc_u32(unsigned int* p) { return atomicAdd(p, 1u); }
  DLPD_DEV void atomic_add_f32(float* p, float v) { atomicAdd(p, v); }
  typedef cudaStream_t dlpd_stream_t;
#endif

// A team is the set of threads that share one "parallel for over tasks" and its barrier.
//   CtaTeam<NT>  : the whole CTA (NT threads, __syncthreads)
//   SubTeam<NT>  : NT threads of a larger CTA with named barrier `bar` (1..15); lets independent sub-problems of one
//                  CTA drift apart instead of marching through every pass in lock step
// DUAL: butterfly passes with <= 10 values per task take two tasks per iteration (loads of both before either butterfly);
// pays off in the gather / level-1 kernels (FwdZC 1.52 -> 1.41 ms, HiddenL1 1.16 -> 1.11 ms), costs registers elsewhere
template <int NT_, bool DUAL_ = false> struct CtaTeam {
    static constexpr int NT = NT_;
    static constexpr bool DUAL = DUAL_;
#if defined(DLPD_EMU)
    DLPD_DEV int tid() const { return 0; }
#else
    DLPD_DEV int tid() const { return (int)threadIdx.x; }
#endif
    DLPD_DEV void sync() const { DLPD_SYNC(); }
};
template <int NT_, bool DUAL_ = false> struct SubTeam {
    static constexpr int NT = NT_;
    static constexpr bool DUAL = DUAL_;
    int tid_, bar_;
    DLPD_DEV int tid() const { return tid_; }
    DLPD_DEV void sync() const { named_barrier_sync(bar_, NT_); }
};
// the same team for a pass whose closing barrier is subsumed by a wider one that follows immediately
template <class TM> struct NoSyncTeam {
    static constexpr int NT = TM::NT;
    static constexpr bool DUAL = TM::DUAL;
    TM tm;
    DLPD_DEV int tid() const { return tm.tid(); }
    DLPD_DEV void sync() const {}
};
#if defined(DLPD_EMU)
  #define DLPD_TEAM_TASKS(tm, t, n) for (int t = 0; t < (n); ++t)
  // pairs (t, t + NT): the emulated thread plays every team thread of the even rounds
  #define DLPD_TEAM_TASKS2(tm, t, n) for (int t = 0; t < (n); ++t) if (((t / (tm).NT) & 1) == 0)
#else
  #define DLPD_TEAM_TASKS(tm, t, n) for (int t = (tm).tid(); t < (n); t += (tm).NT)
  #define DLPD_TEAM_TASKS2(tm, t, n) for (int t = (tm).tid(); t < (n); t += 2 * (tm).NT)
#endif

typedef float2 cpx;
