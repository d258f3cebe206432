// Host side of libdlpd.so: the C ABI declared in include/dlpd.h, kernel launchers and workspace management.
// Built by nvcc for sm_100a (product).  The same file compiles with `g++ -x c++ -DDLPD_EMU` into the
// host emulation used by tests/emu (TEST INFRASTRUCTURE: never loaded by the package).
#include "../../include/dlpd.h"
#if !defined(DLPD_EMU)
#include <cooperative_groups.h>
#include <cuda.h>               // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#endif
#include "dlpd_kernels.h"

#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifndef DLPD_PLANEX_PERSISTENT
#define DLPD_PLANEX_PERSISTENT 1      // cluster plane kernel: resident clusters loop over the planes
#endif

// ------------------------------------------------------------------------------------------------------
// runtime shim
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};     // kernels launched by this library (bench.py reports it as gpu_launches)
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#if defined(DLPD_EMU)
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static int rt_set_device(int) { return 0; }
static int rt_sm_count(int) { return 3; }       // the emulation walks the persistent kernels with 3 "CTAs"
static int rt_malloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : fail(DLPD_ECUDA, "emu: out of memory"); }
static int rt_malloc_host(void** p, size_t n) { return rt_malloc(p, n); }
static void rt_free(void* p) { free(p); }
static void rt_free_host(void* p) { free(p); }
static int rt_copy(void* d, const void* s, size_t n, int, dlpd_stream_t) { memcpy(d, s, n); return 0; }
static int rt_memset(void* d, int v, size_t n, dlpd_stream_t) { memset(d, v, n); return 0; }
static int rt_sync(dlpd_stream_t) { return 0; }
enum { H2D = 0, D2H = 1, D2D = 2 };
template <class K> static int launch(dim3 g, const typename K::Params& p, dlpd_stream_t) {
    std::vector<char> smem((size_t)K::SMEM + 16);
    for (unsigned z = 0; z < g.z; ++z)
        for (unsigned y = 0; y < g.y; ++y)
            for (unsigned x = 0; x < g.x; ++x) K::body(p, (int)x, (int)y, (int)z, smem.data());
    ++g_launches;
    return 0;
}
template <class K, bool CARVE = true> static int launch_mb(dim3 g, const typename K::Params& p, dlpd_stream_t st) { return launch<K>(g, p, st); }
// thread-block clusters: the CL CTAs of a cluster run phase A one after the other, then phase B with every CTA's
// shared memory visible (distributed shared memory on the device)
template <class K> static int launch_cluster(unsigned clusters, const typename K::Params& p, dlpd_stream_t) {
    std::vector<std::vector<char> > smem(K::CL, std::vector<char>((size_t)K::SMEM + 16));
    cpx* slabs[K::CL];
    for (int r = 0; r < K::CL; ++r) slabs[r] = reinterpret_cast<cpx*>(smem[r].data());
    for (unsigned cid = 0; cid < clusters; ++cid) {
        for (int r = 0; r < K::CL; ++r) K::phase_a(p, (int)cid, r, smem[r].data(), cid + 1 < clusters ? (int)cid + 1 : -1);
        if (K::exchanges(p))
            for (int r = 0; r < K::CL; ++r) K::phase_b(p, (int)cid, r, slabs, smem[r].data());
    }
    ++g_launches;
    return 0;
}
#else
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(DLPD_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)
static int rt_set_device(int d) { CU(cudaSetDevice(d)); return 0; }
static int rt_sm_count(int d) {
    int n = 0;
    return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) == cudaSuccess && n > 0 ? n : 148;
}
static int rt_malloc(void** p, size_t n) { CU(cudaMalloc(p, n ? n : 1)); return 0; }
static int rt_malloc_host(void** p, size_t n) { CU(cudaMallocHost(p, n ? n : 1)); return 0; }
static void rt_free(void* p) { if (p) cudaFree(p); }
static void rt_free_host(void* p) { if (p) cudaFreeHost(p); }
enum { H2D = 0, D2H = 1, D2D = 2 };
static int rt_copy(void* d, const void* s, size_t n, int kind, dlpd_stream_t st) {
    cudaMemcpyKind k = kind == H2D ? cudaMemcpyHostToDevice : kind == D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CU(cudaMemcpyAsync(d, s, n, k, st));
    return 0;
}
static int rt_memset(void* d, int v, size_t n, dlpd_stream_t st) { CU(cudaMemsetAsync(d, v, n, st)); return 0; }
static int rt_sync(dlpd_stream_t st) { CU(cudaStreamSynchronize(st)); return 0; }

template <class K>
__global__ void __launch_bounds__(K::NT) dlpd_kernel(const __grid_constant__ typename K::Params p) {
    extern __shared__ __align__(128) char smem[];
    K::body(p, blockIdx.x, blockIdx.y, blockIdx.z, smem);
}
template <class K> static int launch(dim3 g, const typename K::Params& p, dlpd_stream_t st) {
    static std::atomic<bool> configured[64];     // zero-initialised; a concurrent first call only repeats the idempotent attribute set
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        if (K::SMEM > 48 * 1024)
            CU(cudaFuncSetAttribute(dlpd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        configured[dev] = true;
    }
    dlpd_kernel<K><<<g, K::NT, K::SMEM, st>>>(p);
    CU(cudaGetLastError());
    ++g_launches;
    return 0;
}
// same for kernels that state how many CTAs per SM their register budget is cut for (K::MINB)
template <class K>
__global__ void __launch_bounds__(K::NT, K::MINB) dlpd_kernel_mb(const __grid_constant__ typename K::Params p) {
    extern __shared__ __align__(128) char smem[];
    K::body(p, blockIdx.x, blockIdx.y, blockIdx.z, smem);
}
// CARVE: ask for the maximum shared-memory carve-out (kernels that live in shared memory); false for kernels whose
// global gathers need the L1 cache (FwdZC)
template <class K, bool CARVE = true> static int launch_mb(dim3 g, const typename K::Params& p, dlpd_stream_t st) {
    static std::atomic<bool> configured[64];     // zero-initialised; a concurrent first call only repeats the idempotent attribute set
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        if (K::SMEM > 48 * 1024)
            CU(cudaFuncSetAttribute(dlpd_kernel_mb<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        if (CARVE)
            CU(cudaFuncSetAttribute(dlpd_kernel_mb<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[dev] = true;
    }
    dlpd_kernel_mb<K><<<g, K::NT, K::SMEM, st>>>(p);
    CU(cudaGetLastError());
    ++g_launches;
    return 0;
}

// kernels that span a thread-block cluster of K::CL CTAs (distributed shared memory)
// persistent: the resident clusters walk over the work items cid0, cid0 + n_resident, ... (no cluster launch per item)
template <class K>
__global__ void __launch_bounds__(K::NT, K::MINB) dlpd_cluster_kernel(const __grid_constant__ typename K::Params p, int n_items) {
    extern __shared__ __align__(128) char smem[];
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank(), step = (int)(gridDim.x / K::CL);
    cpx* slabs[K::CL];
#pragma unroll
    for (int r = 0; r < K::CL; ++r) slabs[r] = cl.map_shared_rank(reinterpret_cast<cpx*>(smem), r);
    const bool exch = K::exchanges(p);     // uniform over the grid
    bool pending = false;
    for (int cid = (int)(blockIdx.x / K::CL); cid < n_items; cid += step) {
        K::phase_a(p, cid, rank, smem, cid + step < n_items ? cid + step : -1, pending);
        if (exch) {
#if DLPD_PLANEX_LEAD_RELEASE
            if (threadIdx.x < 32) cluster_arrive(); else cluster_arrive_relaxed();
            cluster_wait();
#else
            cl.sync();
#endif
            K::phase_b(p, cid, rank, slabs, smem);
            // nobody re-fills (or leaves) while its shared memory is still being read: arrive now, wait inside the next
            // item's phase A (after its prefetches) or before leaving
#if DLPD_PLANEX_RELAXED_ARRIVE
            cluster_arrive_relaxed();
#else
            cluster_arrive();
#endif
            pending = true;
        } else {
            __syncthreads();
        }
    }
    if (pending) cluster_wait();
}
template <class K> static int launch_cluster(unsigned clusters, const typename K::Params& p, dlpd_stream_t st) {
    static std::atomic<bool> configured[64];     // zero-initialised; a concurrent first call only repeats the idempotent attribute set
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        if (K::SMEM > 48 * 1024)
            CU(cudaFuncSetAttribute(dlpd_cluster_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        CU(cudaFuncSetAttribute(dlpd_cluster_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * K::CL);
    cfg.blockDim = dim3(K::NT);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = K::CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static std::atomic<int> resident[64];     // clusters of this kernel the device can hold at once (0 = not asked yet)
    if (dev < 64 && resident[dev] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, dlpd_cluster_kernel<K>, &cfg) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 1 << 30; }
        resident[dev] = n;
    }
    const unsigned grid_clusters = (DLPD_PLANEX_PERSISTENT && dev < 64 && (unsigned)resident[dev] < clusters) ? (unsigned)resident[dev] : clusters;
    cfg.gridDim = dim3(grid_clusters * K::CL);
    CU(cudaLaunchKernelEx(&cfg, dlpd_cluster_kernel<K>, p, (int)clusters));
    ++g_launches;
    return 0;
}
#endif

#define TRY(call)                  \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != 0) return rc_;  \
    } while (0)

// ------------------------------------------------------------------------------------------------------
// compiled sizes
// ------------------------------------------------------------------------------------------------------
template <int L> struct Tile {
    static constexpr int FWD_YT = (L <= 40) ? L : ((L % 40 == 0) ? 40 : 32);      // y lines per FwdZ CTA
    static constexpr int INV_YT = ((2 * L) % 8 == 0) ? 8 : (((2 * L) % 4 == 0) ? 4 : 2);
    // y lines per FwdZC CTA.  At L = 128: 8 lines = one 132 KB CTA of 1024 threads per SM, 7.88 ms; 4 lines (68 KB, three
    // resident CTAs of 256 threads) 8.55; 4 x 512 threads 8.60; 4 x 384 8.75; 2 lines 11.6  (profiles/r02/cfg5_variants)
#ifdef DLPD_CL_YT
    static constexpr int CL_YT = (L % DLPD_CL_YT == 0) ? DLPD_CL_YT : 2;
#else
    static constexpr int CL_YT = (L % 8 == 0) ? 8 : ((L % 4 == 0) ? 4 : 2);
#endif
#ifdef DLPD_CL_CG
    static constexpr int CL_CG = DLPD_CL_CG;
#else
    static constexpr int CL_CG = 8;                                           // channels per FwdZC CTA
#endif
};

// box sizes the single-level kernels are compiled for
#define DLPD_OP_SIZES(X) X(4) X(8) X(10) X(16) X(20) X(40) X(64) X(80) X(128)
// (L, C0, C1, H) the fused search is compiled for
// (16, 32, 24) is the reference's model (multiplier = 8, local_train.py:23,30); (32, 64, 48) is the representation classes' own
// default multiplier = 16 (ProteinRepresentationModels.py:24,79); (4, 6, 5) is the golden-fixture size
#define DLPD_SEARCH_CONFIGS(X) \
    X(8, 4, 6, 5) X(16, 4, 6, 5) X(20, 4, 6, 5) X(8, 16, 32, 24) X(16, 16, 32, 24) X(20, 16, 32, 24) X(80, 16, 32, 24) \
    X(128, 16, 32, 24) X(16, 32, 64, 48) X(20, 32, 64, 48) X(80, 32, 64, 48)

static bool op_size_ok(int L) {
    switch (L) {
#define X(L_) case L_:
        DLPD_OP_SIZES(X)
#undef X
        return true;
    }
    return false;
}

// twiddle table e^{+2 pi i k / n}, computed in double
static int make_twiddles(cpx** dev, int n, dlpd_stream_t st) {
    std::vector<cpx> h(n);
    for (int k = 0; k < n; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        h[k] = make_float2((float)cos(a), (float)sin(a));
    }
    TRY(rt_malloc((void**)dev, sizeof(cpx) * n));
    TRY(rt_copy(*dev, h.data(), sizeof(cpx) * n, H2D, st));
    TRY(rt_sync(st));
    return 0;
}

static float center_of(int L, int mode) { return mode == DLPD_CENTER_GRID ? 0.5f * (float)(L - 1) : 0.5f * (float)L; }

// ---- single-level launch helpers (dispatch on L) -----------------------------------------------------------
template <int L> static int launch_fwdz_t(const FwdZParams& p, int P, dlpd_stream_t st) {
    typedef FwdZ<L, Tile<L>::FWD_YT> K;
    return launch<K>(dim3(L * K::YTILES, p.C, P), p, st);
}
template <int L> static int launch_fwdzc_t(const FwdZCParams& p, int P, dlpd_stream_t st) {
    typedef FwdZC<L, Tile<L>::CL_CG, Tile<L>::CL_YT> K;
    return launch_mb<K, false>(dim3(L * K::YTILES, p.C / K::CG, P), p, st);
}
// Plane<L> needs the whole N x N plane in one SM's shared memory; PlaneX<L, CL> spreads it over a cluster of CL = 4 or 8 CTAs.
static constexpr int SMEM_MAX = 227 * 1024;
#ifndef DLPD_PLANE_WAVES
#define DLPD_PLANE_WAVES 1      // look-ahead distance of Plane's S1 prefetch in grid waves
#endif
// Kernel-selection options.  A plan carries its own copy (dlpd_plan_set_option), initialised from the process-wide defaults
// below at dlpd_plan_create; dlpd_set_option changes the defaults (plans created afterwards, and the plan-less standalone
// operators).  Nothing a running plan depends on lives in a global.
struct Options {
    int plane_mode = DLPD_PLANE_AUTO;
    int epi_tile = 0;                    // 0 = automatic, 1 / 2 = force 2x1 / 2x2 tiles in the epilogue
    int epi_kind = DLPD_EPILOGUE_AUTO;
    int filter_kind = DLPD_FILTER_AUTO;
};
static Options g_defaults;
static std::mutex g_mutex;               // guards g_defaults and the standalone operators' workspaces
#ifndef DLPD_FILTER_DEFAULT
#define DLPD_FILTER_DEFAULT DLPD_FILTER_TCGEN05      // what DLPD_FILTER_AUTO means where the tcgen05 kernel exists (else FP32)
#endif
#ifndef DLPD_HID_TCGEN05
#define DLPD_HID_TCGEN05 0            // level-1 part of the first filter layer on the tensor cores too (HiddenL1<..., FK = 2>):
                                      // parity-green, measured on B200 and REJECTED -- 1.30 ms vs 1.10 ms per 32 rotations.  K = 32
                                      // needs two staging rounds per rotation through the warpgroup's 32 A columns (two CTAs per SM
                                      // share the 512 TMEM columns), i.e. four exposed MMA round trips per CTA, and the FP32
                                      // contraction it replaces is only a third of that kernel (profiles/r02)
#endif
#ifndef DLPD_EPILOGUE_DEFAULT_DUO
#define DLPD_EPILOGUE_DEFAULT_DUO 0
#endif
#ifndef DLPD_EPILOGUE_DEFAULT_TC_DUO
#define DLPD_EPILOGUE_DEFAULT_TC_DUO 0
#endif
#ifndef DLPD_EPILOGUE_DEFAULT_FUSED
#define DLPD_EPILOGUE_DEFAULT_FUSED 0
#endif
#ifndef DLPD_EPILOGUE_DEFAULT_PIPELINED
#define DLPD_EPILOGUE_DEFAULT_PIPELINED 0
#endif
template <int L> struct HasPlaneX { static constexpr bool value = (L == 8 || L == 16 || L == 20 || L == 40 || L == 64 || L == 80 || L == 128); };
// CTAs per cluster of the cluster plane kernel: 8 at N = 256 (three planes resident per SM), else 4
#ifndef DLPD_PLANEX_CL_BIG
#define DLPD_PLANEX_CL_BIG 8
#endif
#ifndef DLPD_PLANEX_CL_16
#define DLPD_PLANEX_CL_16 4          // test builds set 8 to run the 8-CTA code at a box size the emulation can afford
#endif
template <int L> struct PlaneXCL { static constexpr int value = (L == 128) ? DLPD_PLANEX_CL_BIG : ((L == 16) ? DLPD_PLANEX_CL_16 : 4); };
template <int L> struct HasPlane { static constexpr bool value = ((2 * L) * (2 * L + 1) + 2 * L) * 8 <= SMEM_MAX; };
// which plane kernel `mode` selects at box size L (the receptor spectra are laid out for it)
template <int L> static bool plane_is_cluster_t(int mode) {
    bool cluster = !HasPlane<L>::value;
    if (mode == DLPD_PLANE_CLUSTER && HasPlaneX<L>::value) cluster = true;
    if (mode == DLPD_PLANE_SINGLE) cluster = false;
    return cluster;
}
template <int L> static int launch_plane_t(const PlaneParams& p, int mode, dlpd_stream_t st) {
    const bool cluster = plane_is_cluster_t<L>(mode);
    if (cluster) {
        if constexpr (HasPlaneX<L>::value) return launch_cluster<PlaneX<L, PlaneXCL<L>::value> >((unsigned)(p.P * 2 * L * p.C), p, st);
        else return fail(DLPD_EINVAL, "the cluster plane kernel is not compiled for this box size");
    } else {
        if constexpr (HasPlane<L>::value) {
            const int total = p.P * 2 * L * p.C;
            PlaneParams q = p;
            q.grid = 0;
            // fused search on the big plane: one persistent CTA per SM walking over the planes
            if (DLPD_PLANE_PERSISTENT && Plane<L>::G > 1 && p.spec != nullptr && p.wave > 0 && p.wave < total) {
                q.grid = p.wave;
                return launch_mb<Plane<L> >(dim3((unsigned)q.grid), q, st);
            }
            return launch_mb<Plane<L> >(dim3((unsigned)total), q, st);
        }
        else return fail(DLPD_EINVAL, "the plane of this box size does not fit one SM: use the cluster plane kernel");
    }
}
template <int L> static int launch_invz_t(const InvZParams& p, int Q, dlpd_stream_t st) {
    typedef InvZStore<L, Tile<L>::INV_YT> K;
    return launch<K>(dim3(2 * L * K::YTILES, Q), p, st);
}
static int launch_fwdz(int L, const FwdZParams& p, int P, dlpd_stream_t st) {
    switch (L) {
#define X(L_) case L_: return launch_fwdz_t<L_>(p, P, st);
        DLPD_OP_SIZES(X)
#undef X
    }
    return fail(DLPD_EINVAL, "unsupported box size");
}
static int launch_fwdzc(int L, const FwdZCParams& p, int P, dlpd_stream_t st) {
    switch (L) {
#define X(L_) case L_: return launch_fwdzc_t<L_>(p, P, st);
        DLPD_OP_SIZES(X)
#undef X
    }
    return fail(DLPD_EINVAL, "unsupported box size");
}
static bool channel_last_ok(int C) { return C % 16 == 0; }
// channels per FwdZC CTA at box size L (= the group size of the channel-last ligand copy)
static int cl_group(int L) {
    switch (L) {
#define X(L_) case L_: return Tile<L_>::CL_CG;
        DLPD_OP_SIZES(X)
#undef X
    }
    return 8;
}     // multiple of every Tile<L>::CL_CG in use
static int launch_plane(int L, const PlaneParams& p, int mode, dlpd_stream_t st) {
    switch (L) {
#define X(L_) case L_: return launch_plane_t<L_>(p, mode, st);
        DLPD_OP_SIZES(X)
#undef X
    }
    return fail(DLPD_EINVAL, "unsupported box size");
}
static int plane_is_cluster(int L, int mode) {
    switch (L) {
#define X(L_) case L_: return plane_is_cluster_t<L_>(mode) ? 1 : 0;
        DLPD_OP_SIZES(X)
#undef X
    }
    return 0;
}
static int launch_invz(int L, const InvZParams& p, int Q, dlpd_stream_t st) {
    switch (L) {
#define X(L_) case L_: return launch_invz_t<L_>(p, Q, st);
        DLPD_OP_SIZES(X)
#undef X
    }
    return fail(DLPD_EINVAL, "unsupported box size");
}

// ------------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------------
struct dlpd_plan {
    int L, C0, C1, H, max_rot, center_mode, device;
    int P;                       // pairs per chunk
    cpx *tw0, *tw1, *tw2;        // tables of size 2L, L, L/2... (tw2 unused by kernels; kept for symmetry)
    float *lig0, *lig1;          // [(C0+1)][L^3], [C1][(L/2)^3]
    float *lig0_cl, *lig1_cl;    // channel-last copies [L^3][C0], [(L/2)^3][C1] of the feature volumes (FwdZC)
    cpx *rspec0, *rspec1;        // receptor spectra [(C0+1)][N^3], [C1][L^3]
    cpx *s1_0, *s1_1, *s2_0, *s2_1;
    cpx* h1;                     // [P][L^3][H] level-1 part of the hidden layer (HiddenL1 -> Epilogue)
    float *rec_ss, *lig_ss;      // [C0] squared norms of the level-0 feature volumes (bound the correlations: MlpDev)
    void* mlp;                   // MlpDev<C0, H>: fp16-split first-layer weights for the tensor-core filter
    int sm_count;
    unsigned long long* best;    // [max_rot]
    float *pin_R, *pin_best;     // pinned staging for dlpd_search_host
    float *dev_R, *dev_best;
    int host_cap;
    std::vector<float> W1, b1, W2;
    float b2;
    bool has_receptor, has_ligand, lig_has_clash, has_filter;
    Options opt;                 // kernel selection of THIS plan
    int rspec_cluster[2];        // plane kernel the receptor spectra of level 0 / 1 were laid out for (-1: none yet)
    // ligand atoms for the in-loop clash re-projection inside the library (dlpd_set_ligand_atoms)
    double* atoms;               // [nat][3] bbox-centred coordinates
    int nat;
    float resolution;
    float* clash_rot;            // [max_rot][L^3] re-projected clash volumes of the current chunk
    float* R32;                  // [max_rot][9] fp32 copies of the chunk's rotation matrices
    double* dev_R64;             // staging of dlpd_search_host_atoms
    long long bytes;
    // optional per-kernel timing with CUDA events on the launching stream (dlpd_profile_*)
    bool profile;
    double prof_ms[DLPD_PROFILE_KINDS];
    int prof_n[DLPD_PROFILE_KINDS];
#if !defined(DLPD_EMU)
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
#endif
};

#if defined(DLPD_EMU)
struct ProfScope { ProfScope(dlpd_plan*, int, dlpd_stream_t) {} };
static int prof_collect(dlpd_plan*, dlpd_stream_t) { return 0; }
#else
static cudaEvent_t prof_event(dlpd_plan* p) {
    cudaEvent_t e;
    if (!p->pool.empty()) { e = p->pool.back(); p->pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    dlpd_plan* p; int kind; dlpd_stream_t st; cudaEvent_t a, b;
    ProfScope(dlpd_plan* p_, int kind_, dlpd_stream_t st_) : p(p_), kind(kind_), st(st_) {
        if (p->profile) { a = prof_event(p); b = prof_event(p); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (p->profile) { cudaEventRecord(b, st); p->spans.push_back({kind, a, b}); }
    }
};
static int prof_collect(dlpd_plan* p, dlpd_stream_t st) {
    CU(cudaStreamSynchronize(st));
    for (auto& s : p->spans) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.a, s.b));
        p->prof_ms[s.kind] += ms;
        p->prof_n[s.kind] += 1;
        p->pool.push_back(s.a);
        p->pool.push_back(s.b);
    }
    p->spans.clear();
    return 0;
}
#endif

static size_t n3(int n) { return (size_t)n * n * n; }

extern "C" int dlpd_version(void) { return 100; }
extern "C" const char* dlpd_last_error(void) { return g_err.c_str(); }
extern "C" int dlpd_op_supported(int L) { return op_size_ok(L) ? 1 : 0; }
extern "C" int dlpd_search_supported(int L, int C0, int C1, int H) {
#define X(L_, A, B, C) if (L == L_ && C0 == A && C1 == B && H == C) return 1;
    DLPD_SEARCH_CONFIGS(X)
#undef X
    return 0;
}

extern "C" long long dlpd_plan_bytes(const dlpd_plan* p) { return p ? p->bytes : 0; }

extern "C" int dlpd_plan_destroy(dlpd_plan* p) {
    if (!p) return 0;
    int prev_dev = -1;
#if !defined(DLPD_EMU)
    if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
#endif
    rt_set_device(p->device);
    rt_free(p->atoms); rt_free(p->clash_rot); rt_free(p->R32); rt_free(p->dev_R64);
    rt_free(p->tw0); rt_free(p->tw1); rt_free(p->tw2);
    rt_free(p->lig0); rt_free(p->lig1); rt_free(p->lig0_cl); rt_free(p->lig1_cl); rt_free(p->rspec0); rt_free(p->rspec1);
    rt_free(p->s1_0); rt_free(p->s1_1); rt_free(p->s2_0); rt_free(p->s2_1); rt_free(p->h1);
    rt_free(p->rec_ss); rt_free(p->lig_ss); rt_free(p->mlp);
    rt_free(p->best); rt_free(p->dev_R); rt_free(p->dev_best);
    rt_free_host(p->pin_R); rt_free_host(p->pin_best);
#if !defined(DLPD_EMU)
    for (auto& sp : p->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto& e : p->pool) cudaEventDestroy(e);
#endif
    delete p;
    if (prev_dev >= 0) rt_set_device(prev_dev);     // leave the caller's current device as it was
    return 0;
}

extern "C" int dlpd_plan_create(dlpd_plan** out, int L, int C0, int C1, int H, int max_rot, int center_mode, int device) {
    if (!out) return fail(DLPD_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (!dlpd_search_supported(L, C0, C1, H)) {
        char b[160];
        snprintf(b, sizeof b, "search configuration L=%d C0=%d C1=%d H=%d is not compiled", L, C0, C1, H);
        return fail(DLPD_EINVAL, b);
    }
    if (max_rot < 1) return fail(DLPD_EINVAL, "max_rot must be >= 1");
    if (center_mode != DLPD_CENTER_BOX && center_mode != DLPD_CENTER_GRID) return fail(DLPD_EINVAL, "bad center_mode");
    TRY(rt_set_device(device));
    dlpd_plan* p = new dlpd_plan();
    p->L = L; p->C0 = C0; p->C1 = C1; p->H = H; p->max_rot = max_rot; p->center_mode = center_mode; p->device = device;
    p->P = (max_rot + 1) / 2;
    p->tw0 = p->tw1 = p->tw2 = nullptr; p->lig0 = p->lig1 = nullptr; p->rspec0 = p->rspec1 = nullptr;
    p->lig0_cl = p->lig1_cl = nullptr;
    p->s1_0 = p->s1_1 = p->s2_0 = p->s2_1 = nullptr; p->best = nullptr; p->h1 = nullptr;
    p->rec_ss = p->lig_ss = nullptr; p->mlp = nullptr;
    p->sm_count = rt_sm_count(device);
    p->pin_R = p->pin_best = p->dev_R = p->dev_best = nullptr; p->host_cap = 0;
    p->has_receptor = p->has_ligand = p->lig_has_clash = p->has_filter = false;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        p->opt = g_defaults;
    }
    p->rspec_cluster[0] = p->rspec_cluster[1] = -1;
    p->atoms = nullptr; p->nat = 0; p->resolution = 0.f; p->clash_rot = nullptr; p->R32 = nullptr; p->dev_R64 = nullptr;
    p->b2 = 0.f;
    p->profile = false;
    for (int i = 0; i < DLPD_PROFILE_KINDS; ++i) { p->prof_ms[i] = 0.0; p->prof_n[i] = 0; }
    const int N = 2 * L, L1 = L / 2, N1 = L;
    const size_t P = (size_t)p->P;
    struct { void** ptr; size_t bytes; } allocs[] = {
        {(void**)&p->lig0, sizeof(float) * (C0 + 1) * n3(L)},
        {(void**)&p->lig1, sizeof(float) * C1 * n3(L1)},
        {(void**)&p->lig0_cl, sizeof(float) * C0 * n3(L)},
        {(void**)&p->lig1_cl, sizeof(float) * C1 * n3(L1)},
        {(void**)&p->rspec0, sizeof(cpx) * (C0 + 1) * n3(N)},
        {(void**)&p->rspec1, sizeof(cpx) * C1 * n3(N1)},
        {(void**)&p->s1_0, sizeof(cpx) * P * (C0 + 1) * N * L * L},
        {(void**)&p->s1_1, sizeof(cpx) * P * C1 * N1 * L1 * L1},
        {(void**)&p->s2_0, sizeof(cpx) * P * (C0 + 1) * n3(N)},
        {(void**)&p->s2_1, sizeof(cpx) * P * C1 * n3(N1)},
        {(void**)&p->h1, sizeof(cpx) * P * n3(N1) * H},
        {(void**)&p->best, sizeof(unsigned long long) * (size_t)max_rot},
        {(void**)&p->rec_ss, sizeof(float) * C0},
        {(void**)&p->lig_ss, sizeof(float) * C0},
        {(void**)&p->mlp, 8192},
        {(void**)&p->R32, sizeof(float) * 9 * (size_t)max_rot},
    };
    p->bytes = 0;
    for (auto& a : allocs) {
        int rc = rt_malloc(a.ptr, a.bytes);
        if (rc) { dlpd_plan_destroy(p); return rc; }
        p->bytes += (long long)a.bytes;
    }
    int rc = make_twiddles(&p->tw0, N, 0);
    if (!rc) rc = make_twiddles(&p->tw1, N1, 0);
    if (!rc) rc = make_twiddles(&p->tw2, L1 > 1 ? L1 : 2, 0);
    if (rc) { dlpd_plan_destroy(p); return rc; }
    *out = p;
    return 0;
}

// forward transform of `C` real volumes (R = nullptr) into spectra: FwdZ (identity) + Plane (spectrum mode)
static int real_spectra(int L, const float* vols, int C, cpx* s1, cpx* spec, const cpx* tw, int plane_mode, dlpd_stream_t st) {
    const int N = 2 * L;
    FwdZParams f;
    f.src = vols; f.src_pair_stride = 0; f.src_rot = nullptr; f.R = nullptr; f.out = s1; f.tw = tw;
    f.C = C; f.Ctot = C; f.c_out0 = 0; f.c_ext = -1; f.ext_all = 0; f.n_rot = 1; f.center = 0.f;
    TRY(launch_fwdz(L, f, 1, st));
    PlaneParams q;
    q.in = s1; q.out = spec; q.spec = nullptr; q.spec_pair_stride = 0; q.tw = tw; q.C = C; q.P = 1; q.tiled = 0; q.wave = 1 << 30; q.grid = 0;
    q.scale = (float)(1.0 / ((double)N * N * N));
    TRY(launch_plane(L, q, plane_mode, st));
    return 0;
}

// how this plan evaluates the first filter layer: FP32 / TENSOR (mma.sync, experimental builds) / TCGEN05
static int filter_kind_of(const dlpd_plan* p) {
    const int k = p->opt.filter_kind;
    return k == DLPD_FILTER_AUTO ? DLPD_FILTER_DEFAULT : k;
}
#if DLPD_EXPERIMENTAL
static bool filter_tensor_wanted(const dlpd_plan* p) { return filter_kind_of(p) == DLPD_FILTER_TENSOR; }
// squared L2 norm per level-0 feature channel (max over the n volume sets): bounds the correlations for the fp16-split filter
static int channel_sumsq(const float* vols, int C, int V, int n, float* out, dlpd_stream_t st) {
    TRY(rt_memset(out, 0, sizeof(float) * C, st));
    SumSqParams q;
    q.v = vols; q.out = out; q.C = C; q.V = V;
    return launch<SumSq>(dim3(C, n), q, st);
}
#endif

extern "C" int dlpd_set_receptor(dlpd_plan* p, const float* vol0, const float* vol1, const float* clash, void* stream) {
    if (!p || !vol0 || !vol1 || !clash) return fail(DLPD_EINVAL, "dlpd_set_receptor: NULL argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    const int L = p->L, L1 = L / 2;
    // stage [C0 feature volumes | clash] contiguously in the (not yet needed) ligand buffer
    TRY(rt_copy(p->lig0, vol0, sizeof(float) * p->C0 * n3(L), D2D, st));
    TRY(rt_copy(p->lig0 + p->C0 * n3(L), clash, sizeof(float) * n3(L), D2D, st));
    TRY(real_spectra(L, p->lig0, p->C0 + 1, p->s1_0, p->rspec0, p->tw0, p->opt.plane_mode, st));
#if DLPD_EXPERIMENTAL
    TRY(channel_sumsq(vol0, p->C0, (int)n3(L), 1, p->rec_ss, st));
#endif
    TRY(real_spectra(L1, vol1, p->C1, p->s1_1, p->rspec1, p->tw1, p->opt.plane_mode, st));
    p->rspec_cluster[0] = plane_is_cluster(L, p->opt.plane_mode);
    p->rspec_cluster[1] = plane_is_cluster(L1, p->opt.plane_mode);
    p->has_receptor = true;
    p->has_ligand = false;      // lig0 was used as staging
    return 0;
}

extern "C" int dlpd_set_ligand(dlpd_plan* p, const float* vol0, const float* vol1, const float* clash, void* stream) {
    if (!p || !vol0 || !vol1) return fail(DLPD_EINVAL, "dlpd_set_ligand: NULL argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    const int L = p->L, L1 = L / 2;
    TRY(rt_copy(p->lig0, vol0, sizeof(float) * p->C0 * n3(L), D2D, st));
    if (clash) TRY(rt_copy(p->lig0 + p->C0 * n3(L), clash, sizeof(float) * n3(L), D2D, st));
    else TRY(rt_memset(p->lig0 + p->C0 * n3(L), 0, sizeof(float) * n3(L), st));
    TRY(rt_copy(p->lig1, vol1, sizeof(float) * p->C1 * n3(L1), D2D, st));
#if DLPD_EXPERIMENTAL
    TRY(channel_sumsq(vol0, p->C0, (int)n3(L), 1, p->lig_ss, st));
#endif
    ChannelLastParams cl;
    cl.in = p->lig0; cl.out = p->lig0_cl; cl.C = p->C0; cl.V = (int)n3(L); cl.G = cl_group(L);
    TRY(launch<ToChannelLast>(dim3((cl.V + ToChannelLast::VOX_PER_CTA - 1) / ToChannelLast::VOX_PER_CTA), cl, st));
    cl.in = p->lig1; cl.out = p->lig1_cl; cl.C = p->C1; cl.V = (int)n3(L1); cl.G = cl_group(L1);
    TRY(launch<ToChannelLast>(dim3((cl.V + ToChannelLast::VOX_PER_CTA - 1) / ToChannelLast::VOX_PER_CTA), cl, st));
    p->has_ligand = true;
    p->lig_has_clash = clash != nullptr;
    return 0;
}

extern "C" int dlpd_set_filter(dlpd_plan* p, const float* W1, const float* b1, const float* W2, const float* b2) {
    if (!p || !W1 || !b1 || !W2 || !b2) return fail(DLPD_EINVAL, "dlpd_set_filter: NULL argument");
    const int CT = p->C0 + p->C1;
    p->W1.assign(W1, W1 + (size_t)p->H * CT);
    p->b1.assign(b1, b1 + p->H);
    p->W2.assign(W2, W2 + p->H);
    p->b2 = b2[0];
    p->has_filter = true;
    return 0;
}

template <int C0, int C1, int H> static void fill_filter(const dlpd_plan* p, FilterW<C0, C1, H>& f) {
    for (int j = 0; j < H; ++j)
        for (int c = 0; c < C0 + C1; ++c) f.W1t[c * H + j] = p->W1[(size_t)j * (C0 + C1) + c];
    memcpy(f.b1, p->b1.data(), sizeof(float) * H);
    memcpy(f.W2, p->W2.data(), sizeof(float) * H);
    f.b2 = p->b2;
}

static int filter_kind_of(const dlpd_plan* p);
template <int L, int C0, int C1, int H> static int launch_hidden(dlpd_plan* p, int P, dlpd_stream_t st) {
    typedef HiddenL1<L, C0, C1, H> K;
    typename K::Params e;
    e.s2_1 = p->s2_1; e.h1 = p->h1; e.tw1 = p->tw1;
    fill_filter(p, e.filt);
#if !defined(DLPD_EMU)
    typedef HiddenL1<L, C0, C1, H, 2> KT;             // C1 -> H contraction on tcgen05 / TMEM
    if constexpr (KT::TC_OK && DLPD_HID_TCGEN05) {
        if (filter_kind_of(p) == DLPD_FILTER_TCGEN05) return launch<KT>(dim3(KT::N1 / KT::YT, KT::N1, P), e, st);
    }
#endif
    return launch<K>(dim3(K::N1 / K::YT, K::N1, P), e, st);
}

struct CandArgs {
    int* cand = nullptr;
    unsigned int* count = nullptr;
    unsigned int cap = 0, key = 0;
    int rot_base = 0;
};
#if DLPD_EXPERIMENTAL
// tensor-core filter: on unless switched off, wherever the epilogue kernel has the path (C0 = 16, H % 8 = 0)
template <int L, int C0, int C1, int H> static int launch_filter_prep(dlpd_plan* p, dlpd_stream_t st) {
#if !defined(DLPD_EMU)
    if constexpr (C0 == 16 && H % 8 == 0) {
        static_assert(sizeof(MlpDev<C0, H>) <= 8192, "plan->mlp allocation");
        typedef FilterPrep<C0, C1, H> K;
        typename K::Params e;
        e.rec_ss = p->rec_ss; e.lig_ss = p->lig_ss; e.out = (MlpDev<C0, H>*)p->mlp;
        fill_filter(p, e.filt);
        return launch<K>(dim3(1), e, st);
    }
#endif
    (void)p; (void)st;
    return 0;
}
#endif   // DLPD_EXPERIMENTAL
#if !defined(DLPD_EMU)
// TMA descriptor of the tiled level-0 S2 array [P][C0+1][N][T*T][8 floats] as a rank-5 fp32 tensor with the dimensions
// ordered (payload, channel, kz', tile, pair): a box (8 | 4, C0+1, N, 1, 1) is one epilogue tile and lands in shared memory as
// [kz'][channel][payload], the layout the transform passes work on.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<void*> fn{nullptr};
    void* f = fn.load();
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        fn.store(f);
    }
    return (EncodeTiledFn)f;
}
static int make_s2_tensor_map(const cpx* s2, int C, int N, int T, int P, int payload, unsigned long long (&out)[16]) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(DLPD_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    const cuuint64_t tile = 32;                                  // bytes of one 2x2 tile of one (c, kz') row
    cuuint64_t dims[5] = {8, (cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)T * T, (cuuint64_t)P};
    cuuint64_t strides[4] = {(cuuint64_t)N * T * T * tile, (cuuint64_t)T * T * tile, tile, (cuuint64_t)C * N * T * T * tile};
    cuuint32_t box[5] = {(cuuint32_t)payload, (cuuint32_t)C, (cuuint32_t)N, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)s2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[96];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(DLPD_ECUDA, b);
    }
    memcpy(out, &m, 128);
    return 0;
}
#endif

template <int L, int C0, int C1, int H, int TY, int FK = 0, bool DUO = false>
static int launch_epilogue_t(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    typedef Epilogue<L, C0, C1, H, TY, FK, DUO> K;
    if constexpr (K::SMEM <= SMEM_MAX) {
        typename K::Params e;
        e.s2_0 = p->s2_0; e.h1 = p->h1; e.best = p->best; e.score_out = out_scores;
        e.tw0 = p->tw0; e.n_rot = n; e.tau = tau;
        e.cand = ca.cand; e.cand_count = ca.count; e.cand_cap = ca.cap; e.cand_key = ca.key; e.rot_base = ca.rot_base;
        e.n_tiles = P * K::TILES;
        e.grid = e.n_tiles < K::MINB * p->sm_count ? e.n_tiles : K::MINB * p->sm_count;
        fill_filter(p, e.filt);
        e.mlp = nullptr;
#if !defined(DLPD_EMU)
        if constexpr (K::USE_TMA) TRY(make_s2_tensor_map(p->s2_0, C0 + 1, K::N, K::T, p->P, TY == 2 ? 8 : 4, e.tmap));
#endif
#if DLPD_EXPERIMENTAL
        if constexpr (K::MMA_OK) {
            TRY((launch_filter_prep<L, C0, C1, H>(p, st)));
            e.mlp = (const MlpDev<C0, H>*)p->mlp;
        }
#endif
        if constexpr (K::MINB > 1) return launch_mb<K>(dim3(e.grid), e, st);
        else return launch<K>(dim3(e.grid), e, st);
    } else {
        return fail(DLPD_EINVAL, "epilogue tile does not fit shared memory");
    }
}
#if DLPD_EXPERIMENTAL
template <int L, int C0, int C1, int H>
static int launch_epilogue_pipe(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    typedef EpiloguePipe<L, C0, C1, H> K;
    if constexpr (K::SMEM <= SMEM_MAX) {
        typename K::Params e;
        e.s2_0 = p->s2_0; e.h1 = p->h1; e.best = p->best; e.score_out = out_scores;
        e.tw0 = p->tw0; e.n_rot = n; e.tau = tau;
        e.cand = ca.cand; e.cand_count = ca.count; e.cand_cap = ca.cap; e.cand_key = ca.key; e.rot_base = ca.rot_base;
        e.n_tiles = P * L * L;                      // 2x2 groups
        e.mlp = nullptr;
        e.grid = e.n_tiles < p->sm_count ? e.n_tiles : p->sm_count;
        fill_filter(p, e.filt);
        return launch<K>(dim3(e.grid), e, st);
    } else {
        return fail(DLPD_EINVAL, "the pipelined epilogue does not fit shared memory at this box size");
    }
}
#if !defined(DLPD_EMU)
template <int L, int C0, int C1, int H>
static int launch_epilogue_pipe_tc(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    typedef EpiloguePipeTC<L, C0, C1, H> K;
    if constexpr (K::OK && K::SMEM <= SMEM_MAX) {
        typename K::Params e;
        e.s2_0 = p->s2_0; e.h1 = p->h1; e.best = p->best; e.score_out = out_scores;
        e.tw0 = p->tw0; e.n_rot = n; e.tau = tau;
        e.cand = ca.cand; e.cand_count = ca.count; e.cand_cap = ca.cap; e.cand_key = ca.key; e.rot_base = ca.rot_base;
        e.n_tiles = P * L * L;                      // 2x2 groups
        e.mlp = nullptr;
        e.grid = e.n_tiles < p->sm_count ? e.n_tiles : p->sm_count;
        fill_filter(p, e.filt);
        return launch<K>(dim3(e.grid), e, st);
    } else {
        return fail(DLPD_EINVAL, "the pipelined tcgen05 epilogue is not available for this configuration");
    }
}
#endif
template <int L, int C0, int C1, int H>
static int launch_epilogue_fused(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    typedef EpilogueFused<L, C0, C1, H> K;
    if constexpr (K::SMEM <= SMEM_MAX) {
        typename K::Params e;
        e.s2_0 = p->s2_0; e.h1 = p->h1; e.best = p->best; e.score_out = out_scores;
        e.tw0 = p->tw0; e.n_rot = n; e.tau = tau;
        e.cand = ca.cand; e.cand_count = ca.count; e.cand_cap = ca.cap; e.cand_key = ca.key; e.rot_base = ca.rot_base;
        e.n_tiles = P * L * L;                      // 2x2 groups
        e.mlp = nullptr;
        e.grid = e.n_tiles < p->sm_count ? e.n_tiles : p->sm_count;
        fill_filter(p, e.filt);
        return launch<K>(dim3(e.grid), e, st);
    } else {
        return fail(DLPD_EINVAL, "the fused epilogue does not fit shared memory at this box size");
    }
}
template <int L, int C0, int C1, int H>
static int launch_epilogue_duo(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    typedef EpilogueDuo<L, C0, C1, H> K;
    if constexpr (2 * (K::SMEM + 1024) <= 228 * 1024) {
        typename K::Params e;
        e.s2_0 = p->s2_0; e.h1 = p->h1; e.best = p->best; e.score_out = out_scores;
        e.tw0 = p->tw0; e.n_rot = n; e.tau = tau;
        e.cand = ca.cand; e.cand_count = ca.count; e.cand_cap = ca.cap; e.cand_key = ca.key; e.rot_base = ca.rot_base;
        e.n_tiles = P * L * L;                      // 2x2 groups
        e.mlp = nullptr;
        e.grid = e.n_tiles < 2 * p->sm_count ? e.n_tiles : 2 * p->sm_count;
        fill_filter(p, e.filt);
#if defined(DLPD_EMU)
        return launch<K>(dim3(e.grid), e, st);
#else
        return launch_mb<K>(dim3(e.grid), e, st);
#endif
    } else {
        return fail(DLPD_EINVAL, "two epilogue CTAs per SM do not fit shared memory at this box size");
    }
}
#endif   // DLPD_EXPERIMENTAL
template <int L, int C0, int C1, int H>
static int launch_epilogue(dlpd_plan* p, int n, int P, float tau, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
#if DLPD_EXPERIMENTAL
    const int ek = p->opt.epi_kind, et = p->opt.epi_tile;
    const bool duo_fits = 2 * (EpilogueDuo<L, C0, C1, H>::SMEM + 1024) <= 228 * 1024;
    if (ek == DLPD_EPILOGUE_DUO || (ek == DLPD_EPILOGUE_AUTO && DLPD_EPILOGUE_DEFAULT_DUO && duo_fits && !et))
        return launch_epilogue_duo<L, C0, C1, H>(p, n, P, tau, out_scores, ca, st);
    const bool fused_fits = EpilogueFused<L, C0, C1, H>::SMEM <= SMEM_MAX;
    if (ek == DLPD_EPILOGUE_FUSED || (ek == DLPD_EPILOGUE_AUTO && DLPD_EPILOGUE_DEFAULT_FUSED && fused_fits && !et))
        return launch_epilogue_fused<L, C0, C1, H>(p, n, P, tau, out_scores, ca, st);
#if !defined(DLPD_EMU)
    if (ek == DLPD_EPILOGUE_PIPE_TC) return launch_epilogue_pipe_tc<L, C0, C1, H>(p, n, P, tau, out_scores, ca, st);
#endif
    const bool pipe_fits = EpiloguePipe<L, C0, C1, H>::SMEM <= SMEM_MAX;
    if (ek == DLPD_EPILOGUE_PIPELINED || (ek == DLPD_EPILOGUE_AUTO && DLPD_EPILOGUE_DEFAULT_PIPELINED && pipe_fits && !et))
        return launch_epilogue_pipe<L, C0, C1, H>(p, n, P, tau, out_scores, ca, st);
#endif
    const bool fits2 = Epilogue<L, C0, C1, H, 2>::SMEM <= SMEM_MAX;
    const int ty = p->opt.epi_tile ? p->opt.epi_tile : (fits2 ? 2 : 1);
#if !defined(DLPD_EMU)
    // first filter layer on tcgen05 / TMEM wherever the kernel has the path (C0 % 8 == 0, H % 8 == 0, the tile's
    // warpgroups fit the 512 TMEM columns) unless the plan asks for the FP32 chain
    const int fk = filter_kind_of(p);
    if (fk == DLPD_FILTER_TCGEN05) {
#if DLPD_EXPERIMENTAL
        // two half-tile CTAs per SM (transform of one overlaps the TMEM / MMA phase of the other): parity-green, measured on
        // B200 and rejected -- 4.38 ms vs 3.30 ms per 32 rotations (profiles/r02/bench_tc_duo_rejected.log)
        const bool want_duo = p->opt.epi_kind == DLPD_EPILOGUE_TC_DUO || (p->opt.epi_kind == DLPD_EPILOGUE_AUTO && DLPD_EPILOGUE_DEFAULT_TC_DUO && !p->opt.epi_tile);
        if (want_duo) {
            typedef Epilogue<L, C0, C1, H, 1, 2, true> KD;
            if constexpr (KD::DUO && 2 * (KD::SMEM + 1024) <= 228 * 1024)
                return launch_epilogue_t<L, C0, C1, H, 1, 2, true>(p, n, P, tau, out_scores, ca, st);
            if (p->opt.epi_kind == DLPD_EPILOGUE_TC_DUO)
                return fail(DLPD_EINVAL, "the two-CTA tcgen05 epilogue is not available for this configuration");
        }
#endif
        if (ty == 2) {
            if constexpr (Epilogue<L, C0, C1, H, 2, 2>::TC_OK && Epilogue<L, C0, C1, H, 2, 2>::SMEM <= SMEM_MAX)
                return launch_epilogue_t<L, C0, C1, H, 2, 2>(p, n, P, tau, out_scores, ca, st);
        } else {
            if constexpr (Epilogue<L, C0, C1, H, 1, 2>::TC_OK && Epilogue<L, C0, C1, H, 1, 2>::SMEM <= SMEM_MAX)
                return launch_epilogue_t<L, C0, C1, H, 1, 2>(p, n, P, tau, out_scores, ca, st);
        }
        if (p->opt.filter_kind == DLPD_FILTER_TCGEN05)
            return fail(DLPD_EINVAL, "the tcgen05 filter is not available for this configuration (needs C0 % 8 == 0, H % 8 == 0)");
    }
#if DLPD_EXPERIMENTAL
    if constexpr (C0 == 16 && H % 8 == 0) {
        if (fk == DLPD_FILTER_TENSOR)
            return ty == 2 ? launch_epilogue_t<L, C0, C1, H, 2, 1>(p, n, P, tau, out_scores, ca, st)
                           : launch_epilogue_t<L, C0, C1, H, 1, 1>(p, n, P, tau, out_scores, ca, st);
    }
#endif
#endif
    return ty == 2 ? launch_epilogue_t<L, C0, C1, H, 2>(p, n, P, tau, out_scores, ca, st)
                   : launch_epilogue_t<L, C0, C1, H, 1>(p, n, P, tau, out_scores, ca, st);
}

// one chunk of n <= max_rot rotations
// lig0_rot / lig1_rot != nullptr: every ligand volume arrives already rotated, one set per rotation (dockE3)
static int search_chunk(dlpd_plan* p, const float* R, const float* clash_rot, const float* lig0_rot, const float* lig1_rot,
                        int n, float tau, float* out_best, float* out_scores, const CandArgs& ca, dlpd_stream_t st) {
    const int L = p->L, L1 = L / 2, N = 2 * L;
    const int P = (n + 1) / 2;
    TRY(rt_memset(p->best, 0xff, sizeof(unsigned long long) * n, st));
    {   // level 0: C0 rotated feature channels + the clash channel (rotated here, or supplied already rotated)
        ProfScope ps(p, DLPD_PROFILE_FWDZ0, st);
        FwdZParams f;
        f.src_pair_stride = 0; f.R = R; f.out = p->s1_0; f.tw = p->tw0; f.Ctot = p->C0 + 1; f.n_rot = n; f.ext_all = 0;
        f.center = center_of(L, p->center_mode);
        if (lig0_rot) {
#if DLPD_EXPERIMENTAL
            if (filter_tensor_wanted(p)) TRY(channel_sumsq(lig0_rot, p->C0, (int)n3(L), n, p->lig_ss, st));
#endif
            f.src = p->lig0; f.src_rot = lig0_rot; f.C = p->C0; f.c_out0 = 0; f.c_ext = -1; f.ext_all = 1;
            TRY(launch_fwdz(L, f, P, st));
            f.ext_all = 0;
        } else if (channel_last_ok(p->C0)) {
            FwdZCParams g;
            g.src = p->lig0_cl; g.R = R; g.out = p->s1_0; g.tw = p->tw0; g.C = p->C0; g.Ctot = p->C0 + 1; g.n_rot = n;
            g.center = f.center;
            TRY(launch_fwdzc(L, g, P, st));
        } else {
            f.src = p->lig0; f.src_rot = nullptr; f.C = p->C0; f.c_out0 = 0; f.c_ext = -1;
            TRY(launch_fwdz(L, f, P, st));
        }
        f.src = p->lig0 + (size_t)p->C0 * n3(L); f.src_rot = clash_rot; f.C = 1; f.c_out0 = p->C0; f.c_ext = 0;
        TRY(launch_fwdz(L, f, P, st));
    }
    {   // level 1
        ProfScope ps(p, DLPD_PROFILE_FWDZ1, st);
        if (lig1_rot) {
            FwdZParams f;
            f.src = p->lig1; f.src_pair_stride = 0; f.src_rot = lig1_rot; f.R = R; f.out = p->s1_1; f.tw = p->tw1;
            f.C = p->C1; f.Ctot = p->C1; f.c_out0 = 0; f.c_ext = -1; f.ext_all = 1; f.n_rot = n; f.center = center_of(L1, p->center_mode);
            TRY(launch_fwdz(L1, f, P, st));
        } else if (channel_last_ok(p->C1)) {
            FwdZCParams g;
            g.src = p->lig1_cl; g.R = R; g.out = p->s1_1; g.tw = p->tw1; g.C = p->C1; g.Ctot = p->C1; g.n_rot = n;
            g.center = center_of(L1, p->center_mode);
            TRY(launch_fwdzc(L1, g, P, st));
        } else {
            FwdZParams f;
            f.src = p->lig1; f.src_pair_stride = 0; f.src_rot = nullptr; f.R = R; f.out = p->s1_1; f.tw = p->tw1;
            f.C = p->C1; f.Ctot = p->C1; f.c_out0 = 0; f.c_ext = -1; f.ext_all = 0; f.n_rot = n; f.center = center_of(L1, p->center_mode);
            TRY(launch_fwdz(L1, f, P, st));
        }
    }
    PlaneParams q;
    q.in = p->s1_0; q.out = p->s2_0; q.spec = p->rspec0; q.spec_pair_stride = 0; q.tw = p->tw0; q.C = p->C0 + 1; q.P = P;
    q.scale = 1.f; q.tiled = 1; q.wave = DLPD_PLANE_WAVES * p->sm_count; q.grid = 0;
    { ProfScope ps(p, DLPD_PROFILE_PLANE0, st); TRY(launch_plane(L, q, p->opt.plane_mode, st)); }
    q.in = p->s1_1; q.out = p->s2_1; q.spec = p->rspec1; q.tw = p->tw1; q.C = p->C1; q.tiled = 0; q.wave = 2 * p->sm_count;
    { ProfScope ps(p, DLPD_PROFILE_PLANE1, st); TRY(launch_plane(L1, q, p->opt.plane_mode, st)); }
    int rc = DLPD_EINVAL;
    {
        ProfScope ps(p, DLPD_PROFILE_HIDDEN1, st);
#define X(L_, A, B, C) if (p->L == L_ && p->C0 == A && p->C1 == B && p->H == C) rc = launch_hidden<L_, A, B, C>(p, P, st);
        DLPD_SEARCH_CONFIGS(X)
#undef X
    }
    TRY(rc);
    rc = DLPD_EINVAL;
    {
        ProfScope ps(p, DLPD_PROFILE_EPILOGUE, st);
#define X(L_, A, B, C) if (p->L == L_ && p->C0 == A && p->C1 == B && p->H == C) rc = launch_epilogue<L_, A, B, C>(p, n, P, tau, out_scores, ca, st);
        DLPD_SEARCH_CONFIGS(X)
#undef X
    }
    TRY(rc);
    DecodeParams d;
    d.best = p->best; d.out = out_best; d.n_rot = n; d.N = N;
    TRY(launch<DecodeBest>(dim3((n + DecodeBest::NT - 1) / DecodeBest::NT), d, st));
    return 0;
}

static int project_clash_launch(const double* coords, const double* R, int nat, int n, int L, float resolution, float* out,
                                dlpd_stream_t st);

// R64 != nullptr: float64 rotation matrices; the ligand clash volume of every rotation is re-projected from the plan's
// atoms (dlpd_set_ligand_atoms) chunk by chunk, exactly as Docker.py:221-224 does per rotation
static int search_all(dlpd_plan* p, const float* R, const double* R64, const float* lig_clash_rot, const float* lig0_rot,
                      const float* lig1_rot, int n, float tau, float* out_best, float* out_scores, CandArgs ca, void* stream) {
    const bool prerot = lig0_rot != nullptr;
    if (!p || !out_best || n < 1 || (!prerot && !R && !R64)) return fail(DLPD_EINVAL, "dlpd_search: bad argument");
    if (!p->has_receptor || !p->has_filter || (!prerot && !p->has_ligand))
        return fail(DLPD_ESTATE, "dlpd_search: set_receptor, set_ligand and set_filter must be called first");
    if (R64 && (!p->atoms || !p->clash_rot))
        return fail(DLPD_ESTATE, "dlpd_search_atoms: dlpd_set_ligand_atoms must be called first");
    if (!R64 && !lig_clash_rot && !p->lig_has_clash)
        return fail(DLPD_ESTATE, "dlpd_search: no ligand clash volume (pass one to set_ligand or per rotation)");
    if (plane_is_cluster(p->L, p->opt.plane_mode) != p->rspec_cluster[0] || plane_is_cluster(p->L / 2, p->opt.plane_mode) != p->rspec_cluster[1])
        return fail(DLPD_ESTATE, "dlpd_search: the plane mode was changed after dlpd_set_receptor -- the receptor spectra are laid "
                                 "out for the other plane kernel; call dlpd_set_receptor again");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    const size_t L3 = n3(p->L), N3 = n3(2 * p->L);
    for (int i0 = 0; i0 < n; i0 += p->max_rot) {
        const int m = n - i0 < p->max_rot ? n - i0 : p->max_rot;
        ca.rot_base = i0;
        const size_t L13 = n3(p->L / 2);
        if (R64) {
            CastParams cp;
            cp.in = R64 + 9 * (size_t)i0; cp.out = p->R32; cp.n = 9 * m;
            TRY(launch<CastF64F32>(dim3((cp.n + CastF64F32::NT - 1) / CastF64F32::NT), cp, st));
            TRY(project_clash_launch(p->atoms, R64 + 9 * (size_t)i0, p->nat, m, p->L, p->resolution, p->clash_rot, st));
            TRY(search_chunk(p, p->R32, p->clash_rot, nullptr, nullptr, m, tau, out_best + 4 * (size_t)i0,
                             out_scores ? out_scores + N3 * i0 : nullptr, ca, st));
            continue;
        }
        TRY(search_chunk(p, R ? R + 9 * (size_t)i0 : nullptr, lig_clash_rot ? lig_clash_rot + L3 * i0 : nullptr,
                         prerot ? lig0_rot + L3 * p->C0 * i0 : nullptr, prerot ? lig1_rot + L13 * p->C1 * i0 : nullptr, m, tau,
                         out_best + 4 * (size_t)i0, out_scores ? out_scores + N3 * i0 : nullptr, ca, st));
    }
    return 0;
}

extern "C" int dlpd_search(dlpd_plan* p, const float* R, const float* lig_clash_rot, int n, float tau, float* out_best,
                           float* out_scores, void* stream) {
    return search_all(p, R, nullptr, lig_clash_rot, nullptr, nullptr, n, tau, out_best, out_scores, CandArgs(), stream);
}

extern "C" int dlpd_set_ligand_atoms(dlpd_plan* p, const double* coords, int nat, float resolution, void* stream) {
    if (!p || !coords || nat < 1 || !(resolution > 0.f)) return fail(DLPD_EINVAL, "dlpd_set_ligand_atoms: bad argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    if (nat > p->nat) {
        TRY(rt_sync(st));
        rt_free(p->atoms);
        p->atoms = nullptr; p->nat = 0;
        TRY(rt_malloc((void**)&p->atoms, sizeof(double) * 3 * (size_t)nat));
    }
    if (!p->clash_rot) {
        TRY(rt_malloc((void**)&p->clash_rot, sizeof(float) * n3(p->L) * (size_t)p->max_rot));
        p->bytes += (long long)(sizeof(float) * n3(p->L) * (size_t)p->max_rot);
    }
    TRY(rt_copy(p->atoms, coords, sizeof(double) * 3 * (size_t)nat, D2D, st));
    p->nat = nat;
    p->resolution = resolution;
    return 0;
}

extern "C" int dlpd_search_atoms(dlpd_plan* p, const double* R64, int n, float tau, float* out_best, float* out_scores, void* stream) {
    if (!R64) return fail(DLPD_EINVAL, "dlpd_search_atoms: NULL rotation matrices");
    return search_all(p, nullptr, R64, nullptr, nullptr, nullptr, n, tau, out_best, out_scores, CandArgs(), stream);
}

extern "C" int dlpd_search_prerotated(dlpd_plan* p, const float* lig0_rot, const float* lig1_rot, const float* lig_clash_rot,
                                      int n, float tau, float* out_best, float* out_scores, void* stream) {
    if (!lig0_rot || !lig1_rot || !lig_clash_rot) return fail(DLPD_EINVAL, "dlpd_search_prerotated: NULL volume pointer");
    return search_all(p, nullptr, nullptr, lig_clash_rot, lig0_rot, lig1_rot, n, tau, out_best, out_scores, CandArgs(), stream);
}

static int search_candidates_impl(dlpd_plan* p, const float* R, const double* R64, const float* lig_clash_rot, int n, float tau,
                                  float score_max, float* out_best, int* cand, int cand_cap, unsigned int* cand_count, void* stream) {
    if (!cand || !cand_count || cand_cap < 1) return fail(DLPD_EINVAL, "dlpd_search_candidates: bad candidate buffer");
    if (score_max != score_max) return fail(DLPD_EINVAL, "dlpd_search_candidates: score_max is NaN");
    if (!p) return fail(DLPD_EINVAL, "dlpd_search_candidates: NULL plan");
    TRY(rt_set_device(p->device));
    TRY(rt_memset(cand_count, 0, sizeof(unsigned int), (dlpd_stream_t)stream));
    CandArgs ca;
    ca.cand = cand; ca.count = cand_count; ca.cap = (unsigned int)cand_cap; ca.key = score_key(score_max);
    return search_all(p, R, R64, lig_clash_rot, nullptr, nullptr, n, tau, out_best, nullptr, ca, stream);
}
extern "C" int dlpd_search_candidates(dlpd_plan* p, const float* R, const float* lig_clash_rot, int n, float tau,
                                      float score_max, float* out_best, int* cand, int cand_cap, unsigned int* cand_count,
                                      void* stream) {
    return search_candidates_impl(p, R, nullptr, lig_clash_rot, n, tau, score_max, out_best, cand, cand_cap, cand_count, stream);
}
extern "C" int dlpd_search_candidates_atoms(dlpd_plan* p, const double* R64, int n, float tau, float score_max, float* out_best,
                                            int* cand, int cand_cap, unsigned int* cand_count, void* stream) {
    if (!R64) return fail(DLPD_EINVAL, "dlpd_search_candidates_atoms: NULL rotation matrices");
    return search_candidates_impl(p, nullptr, R64, nullptr, n, tau, score_max, out_best, cand, cand_cap, cand_count, stream);
}

// pinned / device staging of the host-buffer entry points, grown on demand
static int host_staging(dlpd_plan* p, int n) {
    if (n <= p->host_cap) return 0;
    rt_free_host(p->pin_R); rt_free_host(p->pin_best); rt_free(p->dev_R); rt_free(p->dev_best); rt_free(p->dev_R64);
    p->pin_R = p->pin_best = p->dev_R = p->dev_best = nullptr; p->dev_R64 = nullptr;
    p->host_cap = 0;
    TRY(rt_malloc_host((void**)&p->pin_R, sizeof(double) * 9 * n));      // holds fp32 or fp64 matrices
    TRY(rt_malloc_host((void**)&p->pin_best, sizeof(float) * 4 * n));
    TRY(rt_malloc((void**)&p->dev_R, sizeof(float) * 9 * n));
    TRY(rt_malloc((void**)&p->dev_R64, sizeof(double) * 9 * n));
    TRY(rt_malloc((void**)&p->dev_best, sizeof(float) * 4 * n));
    p->host_cap = n;
    return 0;
}

extern "C" int dlpd_search_host(dlpd_plan* p, const float* R_host, int n, float tau, float* best_host, void* stream) {
    if (!p || !R_host || !best_host || n < 1) return fail(DLPD_EINVAL, "dlpd_search_host: bad argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    TRY(host_staging(p, n));
    memcpy(p->pin_R, R_host, sizeof(float) * 9 * n);
    TRY(rt_copy(p->dev_R, p->pin_R, sizeof(float) * 9 * n, H2D, st));
    TRY(dlpd_search(p, p->dev_R, nullptr, n, tau, p->dev_best, nullptr, stream));
    TRY(rt_copy(p->pin_best, p->dev_best, sizeof(float) * 4 * n, D2H, st));
    TRY(rt_sync(st));
    memcpy(best_host, p->pin_best, sizeof(float) * 4 * n);
    return 0;
}

extern "C" int dlpd_search_host_atoms(dlpd_plan* p, const double* R_host, int n, float tau, float* best_host, void* stream) {
    if (!p || !R_host || !best_host || n < 1) return fail(DLPD_EINVAL, "dlpd_search_host_atoms: bad argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_set_device(p->device));
    TRY(host_staging(p, n));
    memcpy(p->pin_R, R_host, sizeof(double) * 9 * n);
    TRY(rt_copy(p->dev_R64, p->pin_R, sizeof(double) * 9 * n, H2D, st));
    TRY(dlpd_search_atoms(p, p->dev_R64, n, tau, p->dev_best, nullptr, stream));
    TRY(rt_copy(p->pin_best, p->dev_best, sizeof(float) * 4 * n, D2H, st));
    TRY(rt_sync(st));
    memcpy(best_host, p->pin_best, sizeof(float) * 4 * n);
    return 0;
}

extern "C" long long dlpd_launch_count(void) { return g_launches.load(); }

static int apply_option(Options& o, int key, int value) {
    switch (key) {
        case DLPD_OPT_PLANE_MODE:
            if (value < DLPD_PLANE_AUTO || value > DLPD_PLANE_CLUSTER) return fail(DLPD_EINVAL, "bad plane mode");
            o.plane_mode = value;
            return 0;
        case DLPD_OPT_EPILOGUE_KIND:
            if (value < DLPD_EPILOGUE_AUTO || value > DLPD_EPILOGUE_PIPE_TC) return fail(DLPD_EINVAL, "bad epilogue kind");
#if !DLPD_EXPERIMENTAL
            if (value > DLPD_EPILOGUE_CLASSIC)
                return fail(DLPD_EINVAL, "this epilogue variant is only compiled with -DDLPD_EXPERIMENTAL=1 (see dlpd_build_flags)");
#endif
            o.epi_kind = value;
            return 0;
        case DLPD_OPT_EPILOGUE_TILE:
            if (value < 0 || value > 2) return fail(DLPD_EINVAL, "epilogue tile must be 0 (auto), 1 (2x1) or 2 (2x2)");
            o.epi_tile = value;
            return 0;
        case DLPD_OPT_FILTER_KIND:
            if (value < DLPD_FILTER_AUTO || value > DLPD_FILTER_TCGEN05) return fail(DLPD_EINVAL, "bad filter kind");
#if !DLPD_EXPERIMENTAL
            if (value == DLPD_FILTER_TENSOR)
                return fail(DLPD_EINVAL, "the mma.sync filter is only compiled with -DDLPD_EXPERIMENTAL=1 (see dlpd_build_flags)");
#endif
#if defined(DLPD_EMU)
            if (value == DLPD_FILTER_TCGEN05) return fail(DLPD_EINVAL, "the tcgen05 filter needs the device");
#endif
            o.filter_kind = value;
            return 0;
    }
    return fail(DLPD_EINVAL, "unknown option");
}

extern "C" int dlpd_set_option(int key, int value) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return apply_option(g_defaults, key, value);
}

extern "C" int dlpd_plan_set_option(dlpd_plan* p, int key, int value) {
    if (!p) return fail(DLPD_EINVAL, "dlpd_plan_set_option: NULL plan");
    return apply_option(p->opt, key, value);
}

extern "C" int dlpd_build_flags(void) {
    int f = 0;
#if DLPD_EXPERIMENTAL
    f |= DLPD_BUILD_EXPERIMENTAL;
#endif
#if defined(DLPD_EMU)
    f |= DLPD_BUILD_EMULATION;
#endif
    return f;
}

extern "C" int dlpd_profile_enable(dlpd_plan* p, int on) {
    if (!p) return fail(DLPD_EINVAL, "dlpd_profile_enable: NULL plan");
    p->profile = on != 0;
    return 0;
}

extern "C" int dlpd_profile_read(dlpd_plan* p, double* ms, int* launches, void* stream) {
    if (!p || !ms || !launches) return fail(DLPD_EINVAL, "dlpd_profile_read: NULL argument");
    TRY(rt_set_device(p->device));
    TRY(prof_collect(p, (dlpd_stream_t)stream));
    for (int i = 0; i < DLPD_PROFILE_KINDS; ++i) {
        ms[i] = p->prof_ms[i]; launches[i] = p->prof_n[i];
        p->prof_ms[i] = 0.0; p->prof_n[i] = 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// standalone operators
// ------------------------------------------------------------------------------------------------------
extern "C" int dlpd_volume_rotate(const float* vol, const float* R, float* out, int B, int C, int L, int center_mode,
                                  void* stream) {
    if (!vol || !R || !out || B < 1 || C < 1 || L < 2) return fail(DLPD_EINVAL, "dlpd_volume_rotate: bad argument");
    if (vol == out) return fail(DLPD_EINVAL, "dlpd_volume_rotate: in-place rotation is not supported");
    RotateParams p;
    p.vol = vol; p.R = R; p.out = out; p.B = B; p.C = C; p.L = L; p.center = center_of(L, center_mode);
    const unsigned gx = (unsigned)((n3(L) + RotateVolume::VOX_PER_CTA - 1) / RotateVolume::VOX_PER_CTA);
    return launch<RotateVolume>(dim3(gx, C, B), p, (dlpd_stream_t)stream);
}

static int project_clash_launch(const double* coords, const double* R, int nat, int n, int L, float resolution, float* out,
                                dlpd_stream_t st) {
    TRY(rt_memset(out, 0, sizeof(float) * n3(L) * (size_t)n, st));
    ProjectParams p;
    p.coords = coords; p.R = R; p.out = out; p.nat = nat; p.n = n; p.L = L; p.res = resolution;
    p.centre = 0.5 * (double)L * (double)resolution;
    return launch<ProjectClash>(dim3((nat + ProjectClash::NT - 1) / ProjectClash::NT, n), p, st);
}

extern "C" int dlpd_project_clash(const double* coords, const double* R, int nat, int n, int L, float resolution, float* out,
                                  void* stream) {
    if (!coords || !R || !out || nat < 1 || n < 1 || L < 1 || !(resolution > 0.f))
        return fail(DLPD_EINVAL, "dlpd_project_clash: bad argument");
    return project_clash_launch(coords, R, nat, n, L, resolution, out, (dlpd_stream_t)stream);
}

extern "C" int dlpd_project_typed(const float* coords, const int* num_atoms_of_type, const int* offsets, int B, int T, int nmax,
                                  int L, float resolution, float* out, void* stream) {
    if (!coords || !num_atoms_of_type || !offsets || !out || B < 1 || T < 1 || nmax < 1 || L < 1 || !(resolution > 0.f))
        return fail(DLPD_EINVAL, "dlpd_project_typed: bad argument");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    TRY(rt_memset(out, 0, sizeof(float) * n3(L) * (size_t)B * (size_t)T, st));
    ProjectTypedParams p;
    p.coords = coords; p.num = num_atoms_of_type; p.off = offsets; p.out = out; p.B = B; p.T = T; p.nmax = nmax; p.L = L;
    p.res = resolution;
    return launch<ProjectTyped>(dim3((nmax + ProjectTyped::NT - 1) / ProjectTyped::NT, B), p, st);
}

struct OpWorkspace {
    cpx *tw = nullptr, *s1 = nullptr, *spec = nullptr, *s2 = nullptr;
    int L = 0, Q = 0;
};
static std::map<int, OpWorkspace> g_ws;     // per device; guarded by g_mutex (held for the whole call: launches on one
                                            // workspace must not interleave with another thread's)
extern "C" int dlpd_release_workspace(void) {
    std::lock_guard<std::mutex> lk(g_mutex);
    int prev = -1;
#if !defined(DLPD_EMU)
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
#endif
    for (auto& kv : g_ws) {
        rt_set_device(kv.first);
#if !defined(DLPD_EMU)
        cudaDeviceSynchronize();
#endif
        rt_free(kv.second.tw); rt_free(kv.second.s1); rt_free(kv.second.spec); rt_free(kv.second.s2);
    }
    g_ws.clear();
    if (prev >= 0) rt_set_device(prev);
    return 0;
}

extern "C" int dlpd_volume_correlate(const float* v1, const float* v2, float* out, int B, int C, int L, void* stream) {
    if (!v1 || !v2 || !out || B < 1 || C < 1) return fail(DLPD_EINVAL, "dlpd_volume_correlate: bad argument");
    if (!op_size_ok(L)) return fail(DLPD_EINVAL, "dlpd_volume_correlate: box size not compiled");
    dlpd_stream_t st = (dlpd_stream_t)stream;
    int dev = 0;
#if !defined(DLPD_EMU)
    CU(cudaGetDevice(&dev));
#endif
    const int N = 2 * L, items = B * C;
    // workspace: ~2 * 8 * N^3 + 8 * N L^2 bytes per item; keep chunks under ~1.5 GB
    const size_t per_item = sizeof(cpx) * (2 * n3(N) + (size_t)N * L * L);
    int Q = (int)((size_t)1536 * 1024 * 1024 / per_item);
    if (Q < 1) Q = 1;
    if (Q > items) Q = items;
    std::lock_guard<std::mutex> lk(g_mutex);
    const int mode = g_defaults.plane_mode;
    OpWorkspace& w = g_ws[dev];
    if (w.L != L || w.Q < Q) {
        TRY(rt_sync(st));
        rt_free(w.tw); rt_free(w.s1); rt_free(w.spec); rt_free(w.s2);
        w = OpWorkspace();
        TRY(make_twiddles(&w.tw, N, st));
        TRY(rt_malloc((void**)&w.s1, sizeof(cpx) * Q * N * L * L));
        TRY(rt_malloc((void**)&w.spec, sizeof(cpx) * Q * n3(N)));
        TRY(rt_malloc((void**)&w.s2, sizeof(cpx) * Q * n3(N)));
        w.L = L; w.Q = Q;
    }
    for (int i0 = 0; i0 < items; i0 += w.Q) {
        const int q = items - i0 < w.Q ? items - i0 : w.Q;
        FwdZParams f;
        f.src = v1 + n3(L) * i0; f.src_pair_stride = (long)n3(L); f.src_rot = nullptr; f.R = nullptr; f.out = w.s1; f.tw = w.tw;
        f.C = 1; f.Ctot = 1; f.c_out0 = 0; f.c_ext = -1; f.ext_all = 0; f.n_rot = 1; f.center = 0.f;
        TRY(launch_fwdz(L, f, q, st));
        PlaneParams pl;
        pl.in = w.s1; pl.out = w.spec; pl.spec = nullptr; pl.spec_pair_stride = 0; pl.tw = w.tw; pl.C = 1; pl.P = q; pl.tiled = 0; pl.wave = 1 << 30; pl.grid = 0;
        pl.scale = (float)(1.0 / ((double)N * N * N));
        TRY(launch_plane(L, pl, mode, st));
        f.src = v2 + n3(L) * i0;
        TRY(launch_fwdz(L, f, q, st));
        pl.out = w.s2; pl.spec = w.spec; pl.spec_pair_stride = (long)n3(N); pl.scale = 1.f;
        TRY(launch_plane(L, pl, mode, st));
        InvZParams iz;
        iz.in = w.s2; iz.out = out + n3(N) * i0; iz.tw = w.tw;
        TRY(launch_invz(L, iz, q, st));
    }
    // two host threads may use different streams: the next caller must not overwrite the workspace under this call
    TRY(rt_sync(st));
    return 0;
}
