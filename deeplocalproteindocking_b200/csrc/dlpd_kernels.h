// Kernel bodies of the docking search (reference path: Docker.dockSE3, /root/reference/src/Docker/Docker.py:211-236).
//
// Two rotations are processed per complex volume ("pair packing"): u = B^{rotA} + i B^{rotB}.  The receptor A is
// real, so corr(A, u) = corr(A, B^{rotA}) + i corr(A, B^{rotB}) and one complex N^3 transform chain serves two
// rotations with no Hermitian bookkeeping.  With G[k] = sum_r u[r] e^{+2 pi i k r/N} (sign + forward) and
// Ahat[k] = sum_r A[r] e^{-2 pi i k r/N} = conj(sign+ transform of A):
//     sum_r A[r+d] u[r] = (1/N^3) sum_k Ahat[k] G[k] e^{+2 pi i k d/N}        (sign + inverse)
// which is the reference's VolumeConvolution output at circular index d mod N (DockingModels.py:71).
//
// Data flow per pair (L = box, N = 2L, ' = digit-reversed frequency order of dlpd_fft.h):
//   FwdZ    : ligand volume --trilinear gather (Docker.py:218)--> z-lines --pruned fwd FFT--> S1[c][kz'][x<L][y<L]
//   Plane   : per (c,kz'): pruned fwd FFT along y, x; multiply by receptor spectrum [c][kz'][kx'][ky'];
//             inverse FFT along x, y                                            --> S2[c][kz'][x][y]
//   Epilogue: per 2x2 (x,y) tile, all channels: inverse FFT along z, nearest-upsampled level 1 (DockingModels.py:74-76),
//             48->24->1 filter (DockingModels.py:28-32,79-82), clash mask (Docker.py:226,232), argmin (Docker.py:91-97)
//             -- the correlation and score volumes are never written to HBM.
#pragma once
#include "dlpd_fft.h"
#include "dlpd_tc.h"

#ifndef DLPD_EXPERIMENTAL
#define DLPD_EXPERIMENTAL 0        // 1: also compile the measured-and-rejected variants (EpiloguePipe / Fused / Duo, mma.sync filter)
#endif

// tunables (overridable with -D for experiments; the defaults are what ships)
#ifndef DLPD_PLANE_NT
#define DLPD_PLANE_NT 640          // with two tasks per iteration (DLPD_PLANE_DUAL): 5.41 ms vs 5.95 ms at 800 threads, single tasks
#endif
#ifndef DLPD_PLANE_NT_MID
#define DLPD_PLANE_NT_MID 256      // Plane<40>: four resident CTAs of 256 threads beat two of 512 (1.27 -> 1.01 ms)
#endif
#ifndef DLPD_HID_YT
#define DLPD_HID_YT 4
#endif
#ifndef DLPD_EPI_NT
#define DLPD_EPI_NT 768
#endif
#ifndef DLPD_EPI_NPROD
#define DLPD_EPI_NPROD 64
#endif
#ifndef DLPD_EPI_DUAL
#define DLPD_EPI_DUAL 0
#endif
#ifndef DLPD_PLANE_MID_DUAL
#define DLPD_PLANE_MID_DUAL 0
#endif
#ifndef DLPD_HID_DIRECT
#define DLPD_HID_DIRECT 0
#endif
#ifndef DLPD_HID_NT
#define DLPD_HID_NT 512
#endif
#ifndef DLPD_HID_YT_BIG
#define DLPD_HID_YT_BIG 2            // HiddenL1 at N1 >= 128 (level 1 of the 128^3 box): 2-line tiles of 70 KB, three CTAs of 256 threads per SM
#endif                               // instead of one 135 KB CTA of 512: 6.16 -> 5.21 ms (2 lines x 512 threads 5.70, x 384 5.25)
#ifndef DLPD_HID_NT_BIG
#define DLPD_HID_NT_BIG 256
#endif
#ifndef DLPD_PLANE_TEAMS
#define DLPD_PLANE_TEAMS 5
#endif
#ifndef DLPD_PLANE_TEAMS_128
#define DLPD_PLANE_TEAMS_128 4       // teams of the 128 x 128 plane (Plane<64>: level 1 of the 128^3 box): 5.83 ms with whole-CTA passes of 640
                                     // threads (1) -> 4.97 with 4 x 128 threads; 2 x 256: 5.13, 4 x 160: 5.19, 8 x 64: 5.73
#endif
#ifndef DLPD_PLANE_NT_128
#define DLPD_PLANE_NT_128 512        // threads of Plane<64> when it runs in teams
#endif
#ifndef DLPD_PLANE_DUAL
#define DLPD_PLANE_DUAL 1            // two butterfly tasks per iteration in Plane's plain passes (needs registers: 640 threads x 96)
#endif
#ifndef DLPD_PLANE_PRELOAD
#define DLPD_PLANE_PRELOAD 0         // persistent Plane: next plane's first-pass inputs fetched into registers early (measured: 5.81 vs 5.38 ms without -- the 20 registers cost more than the hidden latency)
#endif
#ifndef DLPD_PLANE_PERSISTENT
#define DLPD_PLANE_PERSISTENT 1
#endif
#ifndef DLPD_PLANE_COLBLOCKS
#define DLPD_PLANE_COLBLOCKS 1       // inverse-x column ownership that makes the barrier between planes unnecessary
#endif
#ifndef DLPD_EPI_PACKED_TW
#define DLPD_EPI_PACKED_TW 0          // the same for the epilogue's inverse head pass (classic kernel) -- prepared, not measured yet
#endif
#ifndef DLPD_FWDZ_PACKED_TW
#define DLPD_FWDZ_PACKED_TW 0         // ... FwdZC's forward head pass
#endif
#ifndef DLPD_HID_PACKED_TW
#define DLPD_HID_PACKED_TW 0          // ... HiddenL1's inverse head pass
#endif
#ifndef DLPD_PLANE_PACKED_TW
#define DLPD_PLANE_PACKED_TW 1        // head-pass twiddles from a packed table with 16-byte loads (5.11 -> 5.04 ms)
#endif
#ifndef DLPD_PLANE_PREFETCH
#define DLPD_PLANE_PREFETCH 2
#endif
#ifndef DLPD_FWDZC_NT
#define DLPD_FWDZC_NT 256
#endif
#ifndef DLPD_EPI_TMA
#define DLPD_EPI_TMA 0               // tcgen05 epilogue: tiles arrive by TMA (one cp.async.bulk.tensor box + one bulk copy per tile,
                                     // mbarrier) instead of the producer warps' cp.async.  Parity-green, measured on B200 and
                                     // REJECTED: a tile is 2720 rows of 32 contiguous bytes and the TMA unit processes such short
                                     // rows far below HBM speed -- epilogue 7.90 ms vs 3.48 ms per 32 rotations (profiles/r02e)
#endif
#ifndef DLPD_EPI_BULK_H1
#define DLPD_EPI_BULK_H1 1           // tcgen05 epilogue: the level-1 hidden column of a tile (N1 * H complex, 15 KB at 80^3, contiguous
                                     // in HBM) arrives by ONE cp.async.bulk + mbarrier instead of 960 16-byte cp.async of the producers
#endif
#if defined(DLPD_ABLATE_EPI_LOAD)
#undef DLPD_EPI_BULK_H1
#define DLPD_EPI_BULK_H1 0           // the load ablation issues nothing, so nobody may wait for a bulk copy either
#endif
#ifndef DLPD_EPI_TC_EXTRA
#define DLPD_EPI_TC_EXTRA 64         // tcgen05 epilogue: this many extra threads join the inverse-z transform passes (they do not score)
#endif

template <int S> struct LineMul {
    DLPD_DEV int operator()(int l) const { return l * S; }
};

// ------------------------------------------------------------------------------------------------------
// trilinear sample with zero padding:  value of v (L^3, z fastest) at (sx,sy,sz)
// ------------------------------------------------------------------------------------------------------
DLPD_DEV float trilinear_zero(const float* DLPD_RESTRICT v, int L, float sx, float sy, float sz) {
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    if (ix < -1 || iy < -1 || iz < -1 || ix >= L || iy >= L || iz >= L) return 0.0f;
    const float wx1 = sx - fx, wy1 = sy - fy, wz1 = sz - fz;
    const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1, wz0 = 1.0f - wz1;
    const bool x0 = ix >= 0, x1 = ix + 1 < L;
    const bool y0 = iy >= 0, y1 = iy + 1 < L;
    const bool z0 = iz >= 0, z1 = iz + 1 < L;
    const float* p = v + ((long)ix * L + iy) * L + iz;
    const int sxs = L * L, sys = L;
    float acc = 0.0f;
    if (x0) {
        if (y0) {
            if (z0) acc = fmaf_(__ldg(p), wx0 * wy0 * wz0, acc);
            if (z1) acc = fmaf_(__ldg(p + 1), wx0 * wy0 * wz1, acc);
        }
        if (y1) {
            if (z0) acc = fmaf_(__ldg(p + sys), wx0 * wy1 * wz0, acc);
            if (z1) acc = fmaf_(__ldg(p + sys + 1), wx0 * wy1 * wz1, acc);
        }
    }
    if (x1) {
        if (y0) {
            if (z0) acc = fmaf_(__ldg(p + sxs), wx1 * wy0 * wz0, acc);
            if (z1) acc = fmaf_(__ldg(p + sxs + 1), wx1 * wy0 * wz1, acc);
        }
        if (y1) {
            if (z0) acc = fmaf_(__ldg(p + sxs + sys), wx1 * wy1 * wz0, acc);
            if (z1) acc = fmaf_(__ldg(p + sxs + sys + 1), wx1 * wy1 * wz1, acc);
        }
    }
    return acc;
}

// source position of output voxel (x,y,z):  s = ctr + R^T (p - ctr)   (R row-major 3x3)
DLPD_DEV void rotate_point(const float* R, float ctr, int x, int y, int z, float& sx, float& sy, float& sz) {
    const float dx = (float)x - ctr, dy = (float)y - ctr, dz = (float)z - ctr;
    sx = fmaf_(R[6], dz, fmaf_(R[3], dy, fmaf_(R[0], dx, ctr)));
    sy = fmaf_(R[7], dz, fmaf_(R[4], dy, fmaf_(R[1], dx, ctr)));
    sz = fmaf_(R[8], dz, fmaf_(R[5], dy, fmaf_(R[2], dx, ctr)));
}

// ------------------------------------------------------------------------------------------------------
// VolumeRotation as a standalone operator (Docker.py:40,218): out[b,c,p] = trilinear(vol[b,c], ctr + R_b^T (p-ctr))
// ------------------------------------------------------------------------------------------------------
struct RotateParams {
    const float* vol;   // [B][C][L^3]
    const float* R;     // [B][9]
    float* out;         // [B][C][L^3]
    int B, C, L;
    float center;
};

struct RotateVolume {
    typedef RotateParams Params;
    static constexpr int NT = 256;
    static constexpr int SMEM = 0;
    static constexpr int VOX_PER_CTA = 1024;
    // grid: x = ceil(L^3 / VOX_PER_CTA), y = C, z = B
    static DLPD_DEV void body(const Params& p, int bx, int by, int bz, char*) {
        const int L = p.L, L3 = L * L * L;
        const float* R = p.R + 9 * bz;
        float r[9];
        for (int i = 0; i < 9; ++i) r[i] = R[i];
        const float* src = p.vol + ((long)bz * p.C + by) * L3;
        float* dst = p.out + ((long)bz * p.C + by) * L3;
        DLPD_FOR_TASKS(t, VOX_PER_CTA, NT) {
            const int i = bx * VOX_PER_CTA + t;
            if (i < L3) {
                const int z = i % L, y = (i / L) % L, x = i / (L * L);
                float sx, sy, sz;
                rotate_point(r, p.center, x, y, z, sx, sy, sz);
                dst[i] = trilinear_zero(src, L, sx, sy, sz);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// ProjectClash: the in-loop ligand re-projection of dockSE3 (Docker.py:221-224) for a batch of rotations:
//   x = (float)(R_i r + box_centre)   -- rotation and translation in float64 like the reference's CPU CoordsRotate /
//                                        CoordsTranslate, then the cast of `.to(dtype=torch.float)`
//   out[i][v] += exp(-|x - v res|^2 / 2) for the 5^3 voxels v within +-2 cells of floor(x / res) inside the box
// i.e. `TypedCoords2Volume(...)(...).sum(dim=1)` (atom types do not matter once summed).  One thread per
// (atom, rotation); the Gaussian is evaluated per axis (15 exponentials) and scattered with atomic adds.
// ------------------------------------------------------------------------------------------------------
struct ProjectParams {
    const double* coords;    // [nat][3] ligand atoms, bbox-centred (Angstrom)
    const double* R;         // [n][9] row-major
    float* out;              // [n][L^3], zeroed by the launcher
    int nat, n, L;
    float res;
    double centre;           // box_length / 2
};
struct ProjectClash {
    typedef ProjectParams Params;
    static constexpr int NT = 128;
    static constexpr int SMEM = 0;
    // grid: x = ceil(nat / NT), y = n
    static DLPD_DEV void body(const Params& p, int bx, int by, int, char*) {
        const double* R = p.R + 9 * by;
        float* out = p.out + (long)by * p.L * p.L * p.L;
        DLPD_FOR_TASKS(t, NT, NT) {
            const int a = bx * NT + t;
            if (a < p.nat) {
                const double rx = p.coords[3 * a], ry = p.coords[3 * a + 1], rz = p.coords[3 * a + 2];
                const float x[3] = {(float)(R[0] * rx + R[1] * ry + R[2] * rz + p.centre),
                                    (float)(R[3] * rx + R[4] * ry + R[5] * rz + p.centre),
                                    (float)(R[6] * rx + R[7] * ry + R[8] * rz + p.centre)};
                int base[3];
                float g[3][5];
                DLPD_UNROLL
                for (int d = 0; d < 3; ++d) {
                    base[d] = (int)floorf(x[d] / p.res) - 2;
                    DLPD_UNROLL
                    for (int k = 0; k < 5; ++k) {
                        const int i = base[d] + k;
                        const float dx = x[d] - (float)i * p.res;
                        g[d][k] = (i >= 0 && i < p.L) ? expf(-0.5f * dx * dx) : 0.0f;
                    }
                }
                for (int i = 0; i < 5; ++i) {
                    if (g[0][i] == 0.0f) continue;
                    for (int j = 0; j < 5; ++j) {
                        const float gxy = g[0][i] * g[1][j];
                        if (gxy == 0.0f) continue;
                        float* row = out + ((long)(base[0] + i) * p.L + (base[1] + j)) * p.L;
                        DLPD_UNROLL
                        for (int k = 0; k < 5; ++k)
                            if (g[2][k] != 0.0f) atomic_add_f32(row + base[2] + k, gxy * g[2][k]);
                    }
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// ProjectTyped: TorchProteinLibrary.Volume.TypedCoords2Volume (Docker.py:31,204,208; in the dockE3 loop Docker.py:164):
//   out[b][t][v] = sum over the atoms a of type t of structure b of exp(-|x_a - v res|^2 / 2), 5^3 voxels around floor(x / res).
// coords [B][3 nmax] fp32 in TPL's typed layout: the atoms of type t of structure b are a = offsets[b][t] .. + num[b][t] - 1.
// One thread per (atom slot, structure); the same per-axis Gaussian factors and atomic scatter as ProjectClash.
// ------------------------------------------------------------------------------------------------------
struct ProjectTypedParams {
    const float* coords;     // [B][3 nmax]
    const int* num;          // [B][T]
    const int* off;          // [B][T]
    float* out;              // [B][T][L^3], zeroed by the launcher
    int B, T, nmax, L;
    float res;
};
struct ProjectTyped {
    typedef ProjectTypedParams Params;
    static constexpr int NT = 128;
    static constexpr int SMEM = 0;
    // grid: x = ceil(nmax / NT), y = B
    static DLPD_DEV void body(const Params& p, int bx, int by, int, char*) {
        const int* num = p.num + (long)by * p.T;
        const int* off = p.off + (long)by * p.T;
        DLPD_FOR_TASKS(t, NT, NT) {
            const int a = bx * NT + t;
            if (a >= p.nmax) continue;
            int type = -1;
            for (int k = 0; k < p.T; ++k)
                if (a >= off[k] && a < off[k] + num[k]) type = k;       // the last matching type wins, like the reference's sequential fill
            if (type < 0) continue;                                     // padding slot
            const float* c = p.coords + ((long)by * p.nmax + a) * 3;
            float* out = p.out + ((long)by * p.T + type) * p.L * p.L * p.L;
            const float x[3] = {c[0], c[1], c[2]};
            int base[3];
            float g[3][5];
            DLPD_UNROLL
            for (int d = 0; d < 3; ++d) {
                base[d] = (int)floorf(x[d] / p.res) - 2;
                DLPD_UNROLL
                for (int k = 0; k < 5; ++k) {
                    const int i = base[d] + k;
                    const float dx = x[d] - (float)i * p.res;
                    g[d][k] = (i >= 0 && i < p.L) ? expf(-0.5f * dx * dx) : 0.0f;
                }
            }
            for (int i = 0; i < 5; ++i) {
                if (g[0][i] == 0.0f) continue;
                for (int j = 0; j < 5; ++j) {
                    const float gxy = g[0][i] * g[1][j];
                    if (gxy == 0.0f) continue;
                    float* row = out + ((long)(base[0] + i) * p.L + (base[1] + j)) * p.L;
                    DLPD_UNROLL
                    for (int k = 0; k < 5; ++k)
                        if (g[2][k] != 0.0f) atomic_add_f32(row + base[2] + k, gxy * g[2][k]);
                }
            }
        }
    }
};

// float64 -> float32 copy of a batch of rotation matrices (r_batch_gpu = r_batch_cpu.to(float32), Docker.py:215)
struct CastParams {
    const double* in;
    float* out;
    int n;
};
struct CastF64F32 {
    typedef CastParams Params;
    static constexpr int NT = 128;
    static constexpr int SMEM = 0;
    static DLPD_DEV void body(const Params& p, int bx, int, int, char*) {
        DLPD_FOR_TASKS(t, NT, NT) {
            const int i = bx * NT + t;
            if (i < p.n) p.out[i] = (float)p.in[i];
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// FwdZ: gather (rotate) + pruned forward FFT along z, one channel per CTA.
// Used for the clash channel, the receptor / standalone operators and channel counts the channel-last kernel
// (FwdZC below) does not take.
// ------------------------------------------------------------------------------------------------------
struct FwdZParams {
    const float* src;        // [.. pair_stride ..][C][L^3] source volumes
    long src_pair_stride;    // 0: every pair reads the same volumes (search); C*L^3: one real volume per "pair"
    const float* src_rot;    // optional [n_rot][C_ext][L^3]: volumes that are ALREADY rotated -- for channel c_ext only
                             // (C_ext = 1: re-projected clash volume, Docker.py:221-224) or, with ext_all, for every
                             // channel of the launch (C_ext = C: dockE3 re-runs the representation per rotation, Docker.py:165-169)
    const float* R;          // [n_rot][9] rotation matrices, or nullptr: identity, imaginary part = 0
    cpx* out;                // S1 [P][Ctot][N][L][L]
    const cpx* tw;           // [N] e^{+2 pi i k/N}
    int C;                   // channels of this launch (grid.y)
    int Ctot, c_out0;        // channel count of S1 and index of this launch's first channel in it
    int c_ext, n_rot;
    int ext_all;
    float center;
};

template <int L_, int YT_> struct FwdZ {
    typedef FwdZParams Params;
    static constexpr int L = L_, N = 2 * L_, YT = YT_, YTILES = L_ / YT_;
    static constexpr int NT = 256;
    static constexpr int SMEM = (N * YT + N) * (int)sizeof(cpx);
    static_assert(L_ % YT_ == 0, "y tile must divide L");
    // grid: x = L * YTILES, y = C, z = P
    static DLPD_DEV void body(const Params& p, int bx, int by, int bz, char* smem) {
        cpx* buf = reinterpret_cast<cpx*>(smem);       // [N][YT]  (kz, y)
        cpx* tw = buf + N * YT;
        const int x = bx / YTILES, y0 = (bx % YTILES) * YT, c = by, pair = bz;
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        const bool ext = (p.src_rot != nullptr) && (p.ext_all || c == p.c_ext);
        const long ext_c = p.ext_all ? p.C : 1, ext_i = p.ext_all ? c : 0;
        const bool rotate = (p.R != nullptr) && !ext;
        constexpr int L3 = L * L * L;
        float ra[9], rb[9];
        for (int i = 0; i < 9; ++i) {
            ra[i] = rotate ? p.R[9 * rotA + i] : 0.0f;
            rb[i] = (rotate && hasB) ? p.R[9 * rotB + i] : 0.0f;
        }
        const float* src = p.src + (long)pair * p.src_pair_stride + (long)c * L3;
        DLPD_FOR_TASKS(t, N, NT) tw[t] = p.tw[t];
        DLPD_FOR_TASKS(t, YT * L, NT) {
            const int yl = t % YT, z = t / YT, y = y0 + yl;
            const int idx = (x * L + y) * L + z;
            float a, b = 0.0f;
            if (rotate) {
                float sx, sy, sz;
                rotate_point(ra, p.center, x, y, z, sx, sy, sz);
                a = trilinear_zero(src, L, sx, sy, sz);
                if (hasB) {
                    rotate_point(rb, p.center, x, y, z, sx, sy, sz);
                    b = trilinear_zero(src, L, sx, sy, sz);
                }
            } else if (ext) {
                a = __ldg(p.src_rot + ((long)rotA * ext_c + ext_i) * L3 + idx);
                if (hasB) b = __ldg(p.src_rot + ((long)rotB * ext_c + ext_i) * L3 + idx);
            } else {
                a = __ldg(src + idx);                          // plain real volume (receptor / standalone operator)
            }
            buf[z * YT + yl] = make_float2(a, b);              // lower half only: the pruned head never reads z >= L
        }
        DLPD_SYNC();
        fft_fwd_head<N, YT, YT, NT>(buf, tw, LineMul<1>());
        fft_fwd_tail<N, YT, YT, NT>(buf, tw, LineMul<1>());
        cpx* out = p.out + ((((long)pair * p.Ctot + p.c_out0 + c) * N) * L + x) * L + y0;
        DLPD_FOR_TASKS(t, N * YT, NT) {
            const int kz = t / YT, yl = t % YT;
            out[(long)kz * L * L + yl] = buf[t];
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// ToChannelLast: [C][V] -> [C/G][V][G]  (ligand feature volumes, once per pair): channel-last within groups of G
// channels, one group per FwdZC CTA -- every byte of a cache line the gather pulls in belongs to the CTA's own channels
// (with all C channels in one record half of each line was another group's data), and one group of a 128^3 volume
// (67 MB) fits the L2 where the whole volume (134 MB) does not.  G = 0 or C: plain [V][C].
// ------------------------------------------------------------------------------------------------------
struct ChannelLastParams {
    const float* in;
    float* out;
    int C, V, G;
};
struct ToChannelLast {
    typedef ChannelLastParams Params;
    static constexpr int NT = 256;
    static constexpr int SMEM = 0;
    static constexpr int VOX_PER_CTA = 256;
    // grid: x = ceil(V / VOX_PER_CTA)
    static DLPD_DEV void body(const Params& p, int bx, int, int, char*) {
        const int G = (p.G > 0 && p.C % p.G == 0) ? p.G : p.C;
        DLPD_FOR_TASKS(t, VOX_PER_CTA * p.C, NT) {
            const int c = t % p.C, v = bx * VOX_PER_CTA + t / p.C;
            if (v < p.V) p.out[((long)(c / G) * p.V + v) * G + (c % G)] = p.in[(long)c * p.V + v];
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// FwdZC: same as FwdZ for the rotated feature channels, reading a CHANNEL-LAST ligand volume [x][y][z][C]:
// the source position and the 8 interpolation weights of a voxel are computed once for CG channels and every tap is
// one 16-byte load per 4 channels (full 32-byte sectors for CG = 8) instead of CG scattered 4-byte loads.
// ------------------------------------------------------------------------------------------------------
struct FwdZCParams {
    const float* src;        // [C/CG][L^3][CG] channel-last within groups of CG channels (ToChannelLast)
    const float* R;          // [n_rot][9]
    cpx* out;                // S1 [P][Ctot][N][L][L]
    const cpx* tw;
    int C, Ctot, n_rot;
    float center;
};

DLPD_DEV float4 f4_fma(float w, float4 v, float4 acc) {
    return make_float4(fmaf_(w, v.x, acc.x), fmaf_(w, v.y, acc.y), fmaf_(w, v.z, acc.z), fmaf_(w, v.w, acc.w));
}

// trilinear sample of 4 consecutive channels of a channel-last volume, zero outside
DLPD_DEV float4 trilinear_zero_cl4(const float* DLPD_RESTRICT v, int L, int C, float sx, float sy, float sz) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    if (ix < -1 || iy < -1 || iz < -1 || ix >= L || iy >= L || iz >= L) return acc;
    const float wx1 = sx - fx, wy1 = sy - fy, wz1 = sz - fz;
    const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1, wz0 = 1.0f - wz1;
    const bool x0 = ix >= 0, x1 = ix + 1 < L;
    const bool y0 = iy >= 0, y1 = iy + 1 < L;
    const bool z0 = iz >= 0, z1 = iz + 1 < L;
    const float* p = v + (((long)ix * L + iy) * L + iz) * C;
    const long sxs = (long)L * L * C, sys = (long)L * C, szs = C;
#define DLPD_TAP(cond, off, w) if (cond) acc = f4_fma(w, __ldg(reinterpret_cast<const float4*>(p + (off))), acc);
    DLPD_TAP(x0 && y0 && z0, 0, wx0 * wy0 * wz0)
    DLPD_TAP(x0 && y0 && z1, szs, wx0 * wy0 * wz1)
    DLPD_TAP(x0 && y1 && z0, sys, wx0 * wy1 * wz0)
    DLPD_TAP(x0 && y1 && z1, sys + szs, wx0 * wy1 * wz1)
    DLPD_TAP(x1 && y0 && z0, sxs, wx1 * wy0 * wz0)
    DLPD_TAP(x1 && y0 && z1, sxs + szs, wx1 * wy0 * wz1)
    DLPD_TAP(x1 && y1 && z0, sxs + sys, wx1 * wy1 * wz0)
    DLPD_TAP(x1 && y1 && z1, sxs + sys + szs, wx1 * wy1 * wz1)
#undef DLPD_TAP
    return acc;
}

template <int L_, int CG_, int YT_> struct FwdZC {
    typedef FwdZCParams Params;
    static constexpr int L = L_, N = 2 * L_, CG = CG_, YT = YT_, YTILES = L_ / YT_, CQ = CG_ / 4;
    static constexpr int CS = N * YT + 8;          // per-channel stride: +8 keeps 4 channel groups of a warp on distinct banks
    static constexpr int NT = (L_ >= 128 && YT_ >= 8) ? 4 * DLPD_FWDZC_NT                      // 1024 at L = 128 (one 132 KB CTA per SM): 7.88 ms vs 8.25 at 512
                            : ((L_ >= 80 && YT_ >= 8) ? 2 * DLPD_FWDZC_NT : DLPD_FWDZC_NT);   // 512 at L = 80: 1.51 vs 1.62 ms; worse at L = 40
    static constexpr bool PKT = DLPD_FWDZ_PACKED_TW && (N / (2 * LinePlan<N>::R0) > 1);
    static constexpr int SMEM = (CG * CS + N + (PKT ? N : 0)) * (int)sizeof(cpx);
    static constexpr int MINB = 1024 / NT;         // resident CTAs the register budget is cut for (64 registers per thread)
    static_assert(L_ % YT_ == 0 && CG_ % 4 == 0, "tile shape");
    struct Lines {
        DLPD_DEV int operator()(int l) const { return (l / YT) * CS + (l % YT); }
    };
    // grid: x = L * YTILES, y = C / CG, z = P
    static DLPD_DEV void body(const Params& p, int bx, int by, int bz, char* smem) {
        cpx* buf = reinterpret_cast<cpx*>(smem);       // [CG][N][YT]
        cpx* tw = buf + CG * CS;
        const int x = bx / YTILES, y0 = (bx % YTILES) * YT, c0 = by * CG, pair = bz;
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        float ra[9], rb[9];
        for (int i = 0; i < 9; ++i) {
            ra[i] = p.R[9 * rotA + i];
            rb[i] = hasB ? p.R[9 * rotB + i] : 0.0f;
        }
        DLPD_FOR_TASKS(t, N, NT) tw[t] = p.tw[t];
        const cpx* htw = tw;
        if constexpr (PKT) {
            cpx* pk = tw + N;
            DLPD_FOR_TASKS(t, N, NT) pk[t] = packed_head_twiddle<N, LinePlan<N>::R0>(p.tw, t);
            htw = pk;
        }
        const float* src = p.src + (long)by * L * L * L * CG;       // this CTA's channel group
        DLPD_FOR_TASKS(t, CQ * YT * L, NT) {
            const int cq = t % CQ, yl = (t / CQ) % YT, z = t / (CQ * YT), y = y0 + yl;
            float sx, sy, sz;
            rotate_point(ra, p.center, x, y, z, sx, sy, sz);
            const float4 a = trilinear_zero_cl4(src + 4 * cq, L, CG, sx, sy, sz);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hasB) {
                rotate_point(rb, p.center, x, y, z, sx, sy, sz);
                b = trilinear_zero_cl4(src + 4 * cq, L, CG, sx, sy, sz);
            }
            cpx* d = buf + (4 * cq) * CS + z * YT + yl;
            d[0] = make_float2(a.x, b.x);
            d[CS] = make_float2(a.y, b.y);
            d[2 * CS] = make_float2(a.z, b.z);
            d[3 * CS] = make_float2(a.w, b.w);
        }
        DLPD_SYNC();
        const CtaTeam<NT, true> tm;
        team_fwd_head<N, CG * YT, YT, LinePlan<N>, PKT>(tm, buf, htw, Lines());
        team_fwd_tail<N, CG * YT, YT>(tm, buf, tw, Lines());
        cpx* out = p.out + ((((long)pair * p.Ctot + c0) * N) * L + x) * L + y0;
        DLPD_FOR_TASKS(t, CG * N * YT, NT) {
            const int yl = t % YT, kz = (t / YT) % N, c = t / (YT * N);
            out[((long)c * N + kz) * L * L + yl] = buf[c * CS + kz * YT + yl];
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// Plane: for one (pair, channel, kz') the whole (x,y) plane lives in shared memory.
//   forward x (head reads S1 from global, columns y < L) -> forward y head (rows) -> [forward y tail . receptor
//   multiply . inverse y tail] in registers -> inverse y head -> inverse x tail -> inverse x head (writes S2 to global).
// The receptor spectrum is stored TRANSPOSED ([kz'][ky'][kx']) so that the multiply, whose lanes run over rows kx',
// reads it coalesced.
// ------------------------------------------------------------------------------------------------------
struct PlaneParams {
    const cpx* in;           // S1 [P][C][N][L][L]
    cpx* out;                // S2 [P][C][N][N][N]   (or, in spectrum mode, [P][C][N][ky'][kx'] spectra)
    const cpx* spec;         // receptor spectrum [..][C][N][ky'][kx']; nullptr = spectrum mode (write conj(fwd) * scale)
    long spec_pair_stride;   // 0: all pairs share the receptor (search); C*N^3: one spectrum per "pair"
    const cpx* tw;           // [N]
    int C, P;
    int wave;                // CTAs per wave (SM count x resident CTAs): distance of the S1 look-ahead prefetch
    int grid;                // > 0: persistent launch with this many CTAs, each walking planes bx, bx + grid, ... (fused search,
                             //      team path); 0: one CTA per plane
    int tiled;               // S2 rows in natural [x][y] order (0) or in 2x2 tiles [x/2][y/2][x&1][y&1] (1, fused search)
    float scale;
};

template <int L_> struct Plane {
    typedef PlaneParams Params;
    static constexpr int L = L_, N = 2 * L_, NP = N + 1;     // odd pitch: conflict-free along both axes
    typedef PlanePlan<N> PP;                                 // factorisation of the x / y lines
    static constexpr int TEAMS = (N == 128) ? DLPD_PLANE_TEAMS_128 : DLPD_PLANE_TEAMS;
    static constexpr int NT = (N == 128 && TEAMS > 1) ? DLPD_PLANE_NT_128
                            : ((N >= 128) ? DLPD_PLANE_NT : ((N >= 64) ? DLPD_PLANE_NT_MID : 256));   // 640 = 5 teams x 128: whole head passes per round, <= 102 registers
    static constexpr bool PKT = DLPD_PLANE_PACKED_TW && (N >= 128);       // packed head-pass twiddles (team path)
    static constexpr int SMEM = (N * NP + N + (PKT ? N : 0)) * (int)sizeof(cpx);
    static constexpr int MINB = (SMEM <= 56 * 1024) ? 4 : 1;   // resident CTAs per SM the register budget is cut for
    // teams (see body): only for the big plane, and only when rows, columns and threads split evenly
    static constexpr int G = (N >= 128 && N % TEAMS == 0 && L_ % TEAMS == 0 && NT % (32 * TEAMS) == 0) ? TEAMS : 1;
    static constexpr int TNT = NT / G;
    static constexpr int PFT = (G > 1) ? TNT : NT;           // threads that issue the prefetches
    // DLPD_ABLATE_* (never defined in the shipped build): timing experiments that remove one memory stream at a time
    struct LoadS1 {          // element e = x of column l = y
        const cpx* in;
#if defined(DLPD_ABLATE_PLANE_LOAD)
        DLPD_DEV cpx operator()(int l, int e) const { return make_float2((float)l, (float)e); }
#else
        DLPD_DEV cpx operator()(int l, int e) const { return in[e * L + l]; }
#endif
    };
    struct MulSpec {         // element e = ky' of row l = kx'
        const cpx* spec;
#if defined(DLPD_ABLATE_PLANE_MUL)
        DLPD_DEV cpx operator()(int l, int e) const { return make_float2(1.0f + (float)l, (float)e); }
#else
        DLPD_DEV cpx operator()(int l, int e) const { return __ldg(spec + e * N + l); }
#endif
    };
    struct StoreS2 {         // element e = x of column l = y, natural [x][y] rows
        cpx* out;
        DLPD_DEV void operator()(int l, int e, cpx v) const { out[e * N + l] = v; }
    };
    // tiled S2 (fused search): [x/2][y/2][y&1][x&1] -- the 2x2 (x,y) tile of the epilogue is 32 contiguous bytes
    struct StoreS2Tiled {
        cpx* out;
        DLPD_DEV void operator()(int l, int e, cpx v) const { out[((((e >> 1) * L) + (l >> 1)) << 2) + ((l & 1) << 1) + (e & 1)] = v; }
    };
    struct StoreS2TiledPair {   // elements x = e (even) and e + 1 of column y = l: one 16-byte store, 512 contiguous bytes per warp
        cpx* out;
        DLPD_DEV void operator()(int l, int e, cpx v0, cpx v1) const {
#if defined(DLPD_ABLATE_PLANE_STORE)
            if (v0.x == 1.2345e30f) out[0] = v1;
#else
            store_cpx2(out + ((((e >> 1) * L) + (l >> 1)) << 2) + ((l & 1) << 1), v0, v1);
#endif
        }
    };
    // Column ownership of the inverse-x passes in the team path: team g takes the two blocks of L/G columns that start at
    // y = g L/G and y = L + g L/G.  The first block is exactly the set of columns the same team fills in the forward-x
    // passes of the NEXT plane, so a persistent CTA needs no CTA-wide barrier between planes: nobody else reads them.
    static constexpr int CB = (G > 1) ? L / G : 1;
    struct ColBlocks {       // local line l in [0, 2 CB) -> column offset from the team's first column
        DLPD_DEV int operator()(int l) const { return l < CB ? l : l + (L - CB); }
    };
    struct StoreS2TiledPairCB {  // StoreS2TiledPair for a team's two column blocks; y0 = g CB
        cpx* out;
        int y0;
        DLPD_DEV void operator()(int l, int e, cpx v0, cpx v1) const {
            const int y = y0 + (l < CB ? l : l + (L - CB));
#if defined(DLPD_ABLATE_PLANE_STORE)
            if (v0.x == 1.2345e30f) out[y] = v1;
#else
            store_cpx2(out + ((((e >> 1) * L) + (y >> 1)) << 2) + ((y & 1) << 1), v0, v1);
#endif
        }
    };
    struct StoreS2CB {
        cpx* out;
        int y0;
        DLPD_DEV void operator()(int l, int e, cpx v) const { out[e * N + y0 + (l < CB ? l : l + (L - CB))] = v; }
    };
    // grid: x = P * N * C  with the pair index fastest (CTAs that share a receptor plane run together -> L2 reuse)
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* pl = reinterpret_cast<cpx*>(smem);        // [x][y], pitch NP
        cpx* tw = pl + N * NP;
        const int pair = bx % p.P, kz = (bx / p.P) % N, c = bx / (p.P * N);
        DLPD_FOR_TASKS(t, N, NT) tw[t] = p.tw[t];
        DLPD_SYNC();
        const cpx* htw = tw;                           // twiddles of the head passes
        if constexpr (PKT) {
            cpx* pk = tw + N;
            DLPD_FOR_TASKS(t, N, NT) pk[t] = packed_head_twiddle<N, PP::R0>(tw, t);
            DLPD_SYNC();
            htw = pk;
        }
        const long plane = ((long)pair * p.C + c) * N + kz;
        const cpx* spec = p.spec ? p.spec + (long)pair * p.spec_pair_stride + ((long)c * N + kz) * N * N : nullptr;
        if constexpr (G > 1) {
            // G teams of TNT threads: each owns L/G columns (x passes) or N/G rows (y passes) and synchronises on its own
            // named barrier between the passes of one direction, so the teams drift apart and their shared-memory and
            // arithmetic phases overlap; only the two transpositions (x -> y -> x) need the whole CTA.
            // Persistent form (p.grid > 0): the CTA walks planes bx, bx + grid, ...; while the last two passes of a plane run,
            // every thread already holds the S1 values of ITS tasks of the next plane's first pass in registers, so neither
            // the CTA launch nor the L2 round trip of the first loads is exposed between planes.
            if (spec != nullptr) {
                const int total = p.P * N * p.C, stride = p.grid > 0 ? p.grid : total;
                typedef SubTeam<TNT, DLPD_PLANE_DUAL != 0> Team;
#if !defined(DLPD_EMU)
                constexpr int NPRE = (L / G) * (N / (2 * PP::R0)) / TNT;            // first-pass tasks per thread
                constexpr bool PRE = DLPD_PLANE_PRELOAD && ((L / G) * (N / (2 * PP::R0)) == NPRE * TNT) && NPRE >= 1 && NPRE <= 2;
                typedef PrunedHeadPre<PP::R0, L / G, N, N, NP, PRE ? NPRE : 1, Team, LineMul<1> > Pre;
                cpx nx[PRE ? NPRE : 1][PP::R0];
                bool have = false;
#endif
                for (int b = bx; b < total; b += stride) {
                    const int pair_b = b % p.P, kz_b = (b / p.P) % N, c_b = b / (p.P * N);
                    const long plane_b = ((long)pair_b * p.C + c_b) * N + kz_b;
                    const cpx* spec_b = p.spec + (long)pair_b * p.spec_pair_stride + ((long)c_b * N + kz_b) * N * N;
                    DLPD_FOR_TEAMS(g, G, TNT) {
                        const Team tm{DLPD_TEAM_TID(TNT), g + 1};
                        GlobalIn<LoadS1> in{LoadS1{p.in + plane_b * L * L + g * (L / G)}};
                        bool first_done = false;                                                         // along x, columns y < L
#if !defined(DLPD_EMU)
                        if constexpr (PRE) {
                            if (have) {
                                Pre::run(tm, pl + g * (L / G), tw, LineMul<1>(), nx);
                                first_done = true;
                            }
                        }
#endif
                        if (!first_done) team_fwd_head<N, L / G, NP, PP, PKT>(tm, pl + g * (L / G), htw, LineMul<1>(), in);
                        if (g == 0) prefetch(p, b, spec_b);
                        team_fwd_tail<N, L / G, NP, PP>(NoSyncTeam<Team>{tm}, pl + g * (L / G), tw, LineMul<1>());     // CTA barrier next
                    }
                    DLPD_SYNC();
                    DLPD_FOR_TEAMS(g, G, TNT) {
                        const Team tm{DLPD_TEAM_TID(TNT), g + 1};
                        cpx* rows = pl + g * (N / G) * NP;
                        PointMul<MulSpec> mul{MulSpec{spec_b + g * (N / G)}};
#if defined(DLPD_ABLATE_PLANE_YBAR)      // timing experiment (wrong results): the y phase without its two team barriers
                        team_fwd_head<N, N / G, 1, PP, PKT>(NoSyncTeam<Team>{tm}, rows, htw, LineMul<NP>());
                        team_mid<N, N / G, 1, PP>(NoSyncTeam<Team>{tm}, rows, tw, LineMul<NP>(), mul);
#else
                        team_fwd_head<N, N / G, 1, PP, PKT>(tm, rows, htw, LineMul<NP>());                   // along y, rows kx'
                        team_mid<N, N / G, 1, PP>(tm, rows, tw, LineMul<NP>(), mul);                        // y tail . multiply . inverse y tail
#endif
                        team_inv_head<N, N / G, 1, PP, PKT>(NoSyncTeam<Team>{tm}, rows, htw, LineMul<NP>());  // inverse y; CTA barrier next
                    }
                    DLPD_SYNC();
                    DLPD_FOR_TEAMS(g, G, TNT) {
                        const Team tm{DLPD_TEAM_TID(TNT), g + 1};
#if DLPD_PLANE_COLBLOCKS
                        cpx* cols = pl + g * CB;
                        typedef ColBlocks InvOff;
                        static_assert(N / G == 2 * CB && CB % 2 == 0, "two column blocks per team");
#else
                        cpx* cols = pl + g * (N / G);
                        typedef LineMul<1> InvOff;
#endif
#if !defined(DLPD_EMU)
                        if constexpr (PRE) {
                            have = b + stride < total;
                            if (have) {
                                const int b2 = b + stride;
                                const long plane2 = ((long)(b2 % p.P) * p.C + b2 / (p.P * N)) * N + (b2 / p.P) % N;
                                Pre::fetch(tm, nx, GlobalIn<LoadS1>{LoadS1{p.in + plane2 * L * L + g * (L / G)}});
                            }
                        }
#endif
                        team_inv_tail<N, N / G, NP, PP>(tm, cols, tw, InvOff());                            // inverse x
                        static_assert((N / G) % 2 == 0 && (N / 2 / PP::R0) % 2 == 0, "tiled pair stores");
#if DLPD_PLANE_COLBLOCKS
                        if (p.tiled) {
                            GlobalOutPair<StoreS2TiledPairCB> out{StoreS2TiledPairCB{p.out + plane_b * N * N, g * CB}};
                            team_inv_head<N, N / G, NP, PP, PKT>(tm, cols, htw, InvOff(), out);
                        } else {
                            GlobalOut<StoreS2CB> out{StoreS2CB{p.out + plane_b * N * N, g * CB}};
                            team_inv_head<N, N / G, NP, PP, PKT>(tm, cols, htw, InvOff(), out);
                        }
#else
                        if (p.tiled) {
                            GlobalOutPair<StoreS2TiledPair> out{StoreS2TiledPair{p.out + plane_b * N * N + g * (N / G / 2) * 4}};
                            team_inv_head<N, N / G, NP, PP, PKT>(tm, cols, htw, InvOff(), out);
                        } else {
                            GlobalOut<StoreS2> out{StoreS2{p.out + plane_b * N * N + g * (N / G)}};
                            team_inv_head<N, N / G, NP, PP, PKT>(tm, cols, htw, InvOff(), out);
                        }
#endif
                    }
#if !DLPD_PLANE_COLBLOCKS
                    if (b + stride < total) DLPD_SYNC();      // the next plane's first pass overwrites columns other teams still read
#endif
                }
                return;
            }
        }
        const CtaTeam<NT, (N < 128) && (DLPD_PLANE_MID_DUAL != 0)> tmc;
        GlobalIn<LoadS1> in{LoadS1{p.in + plane * L * L}};
        team_fwd_head<N, L, NP, PP>(tmc, pl, tw, LineMul<1>(), in);                   // along x, columns y < L
        prefetch(p, bx, spec);
        team_fwd_tail<N, L, NP, PP>(tmc, pl, tw, LineMul<1>());
        team_fwd_head<N, N, 1, PP>(tmc, pl, tw, LineMul<NP>());                       // along y, all rows kx'
        if (p.spec == nullptr) {
            team_fwd_tail<N, N, 1, PP>(tmc, pl, tw, LineMul<NP>());
            cpx* out = p.out + plane * N * N;
            DLPD_FOR_TASKS(t, N * N, NT) {
                const int kx = t % N, ky = t / N;                               // transposed store, coalesced over kx
                const cpx v = pl[kx * NP + ky];
                out[t] = make_float2(v.x * p.scale, -v.y * p.scale);
            }
            return;
        }
        PointMul<MulSpec> mul{MulSpec{spec}};
        team_mid<N, N, 1, PP>(tmc, pl, tw, LineMul<NP>(), mul);                       // y tail . multiply . inverse y tail
        team_inv_head<N, N, 1, PP>(tmc, pl, tw, LineMul<NP>());                       // inverse y
        team_inv_tail<N, N, NP, PP>(tmc, pl, tw, LineMul<1>());                       // inverse x
        if (p.tiled) {
            if constexpr ((N / 2 / PP::R0) % 2 == 0) {
                GlobalOutPair<StoreS2TiledPair> out{StoreS2TiledPair{p.out + plane * N * N}};
                team_inv_head<N, N, NP, PP>(tmc, pl, tw, LineMul<1>(), out);
            } else {
                GlobalOut<StoreS2Tiled> out{StoreS2Tiled{p.out + plane * N * N}};
                team_inv_head<N, N, NP, PP>(tmc, pl, tw, LineMul<1>(), out);
            }
        } else {
            GlobalOut<StoreS2> out{StoreS2{p.out + plane * N * N}};
            team_inv_head<N, N, NP, PP>(tmc, pl, tw, LineMul<1>(), out);
        }
    }

    // L2 look-ahead: the receptor plane (consumed two passes later) and the S1 plane of the CTA that will follow this
    // one on the SM (about one grid wave ahead)
    static DLPD_DEV void prefetch(const Params& p, int bx, const cpx* spec) {
        (void)p; (void)bx; (void)spec;
#if DLPD_PLANE_PREFETCH
        if (spec) DLPD_FOR_TASKS_SUB(t, N * N / 16, PFT, 0) prefetch_l2(reinterpret_cast<const char*>(spec) + 128 * t);
#endif
#if DLPD_PLANE_PREFETCH >= 2
        if (bx + p.wave < p.P * N * p.C) {
            const int b2 = bx + p.wave;
            const long plane2 = ((long)(b2 % p.P) * p.C + b2 / (p.P * N)) * N + (b2 / p.P) % N;
            DLPD_FOR_TASKS_SUB(t, L * L / 16, PFT, 0) prefetch_l2(reinterpret_cast<const char*>(p.in + plane2 * L * L) + 128 * t);
        }
#endif
    }
};

// ------------------------------------------------------------------------------------------------------
// PlaneX: the same (c, kz') plane pipeline spread over a thread-block CLUSTER of CL CTAs, for planes that do not fit
// one SM's shared memory (N = 256: 526 KB) or that should leave room for several resident CTAs per SM (N = 160).
//
// The x transform is split by its outermost radix-CL step.  With kx = q + CL k2 (q < CL, k2 < M = N / CL):
//     forward:  X[q + CL k2] = FFT_M( y_q )[k2],   y_q[j] = w_N^{q j} sum_{m < CL/2} x[j + m M] e^{2 pi i q m / CL}
//               (the zero-padded upper half drops out of the sum) -- every slab q needs only the INPUT column, so the
//               forward direction has no exchange at all: CTA q folds the S1 plane into its own M rows;
//     inverse:  A_q = IFFT_M(slab q) is local as well; only the last step
//               out[j + m M] = sum_q e^{2 pi i q m / CL} (w_N^{q j} A_q[j])
//               mixes the slabs.  Each CTA gathers its share of rows j from all CL slabs through distributed shared
//               memory (one coalesced read per value), applies the radix-CL butterfly and stores S2 straight to global.
// Between the two, the y transforms, the receptor multiply and the inverse y transform work on rows, i.e. inside a slab.
// The frequency order along kx' differs from Plane<L>'s; it is private to this kernel (the receptor spectrum is
// produced by the same code in spectrum mode).
// ------------------------------------------------------------------------------------------------------
template <int M> struct XPlan;            // M = XA * XB: two single-radix shared-memory passes
#define DLPD_XPLAN(M_, A_, B_)                                            \
    template <> struct XPlan<M_> {                                        \
        static constexpr int XA = A_, XB = B_;                            \
        static_assert(A_ * B_ == M_, "bad x factorisation");              \
    };
DLPD_XPLAN(4, 4, 1)
DLPD_XPLAN(8, 8, 1)
DLPD_XPLAN(10, 5, 2)
DLPD_XPLAN(20, 5, 4)
DLPD_XPLAN(32, 8, 4)
DLPD_XPLAN(40, 5, 8)
DLPD_XPLAN(64, 8, 8)

#ifndef DLPD_PLANEX_NT
#define DLPD_PLANEX_NT 256
#endif
#ifndef DLPD_PLANEX_NT_CL8
#define DLPD_PLANEX_NT_CL8 256
#endif
#ifndef DLPD_PLANEX_MINB_CL8
#define DLPD_PLANEX_MINB_CL8 3
#endif
#ifndef DLPD_PLANEX_RELAXED_ARRIVE
#define DLPD_PLANEX_RELAXED_ARRIVE 1   // the barrier after the exchange phase arrives .relaxed: it only says "I have finished reading my
#endif                                 // peers' slabs" and must not wait for the S2 stores in flight (MEMBAR.ALL.GPU): 38.0-38.6 -> 37.3 ms
#ifndef DLPD_PLANEX_KEEP_TW
#define DLPD_PLANEX_KEEP_TW 1          // persistent clusters copy the twiddle table once instead of once per plane (one CTA barrier less): 37.4 -> 37.05 ms
#endif
#ifndef DLPD_PLANEX_SKIP_LAST_SYNC
#define DLPD_PLANEX_SKIP_LAST_SYNC 0
#endif
#ifndef DLPD_PLANEX_LEAD_RELEASE
#define DLPD_PLANEX_LEAD_RELEASE 0     // barrier before the exchange phase: only warp 0 arrives .release (the CTA barrier that ends phase A
#endif                                 // has already ordered every thread's shared-memory writes before it), the others .relaxed
#ifndef DLPD_PLANEX_NT_BIG
#define DLPD_PLANEX_NT_BIG 512
#endif

template <int L_, int CL_> struct PlaneX {
    typedef PlaneParams Params;
    static constexpr int L = L_, N = 2 * L_, CL = CL_, M = N / CL_, NP = N + 1;
    static constexpr int XA = XPlan<M>::XA, XB = XPlan<M>::XB;
    // 8-CTA clusters hold N/8 rows each (68 KB at N = 256): three CTAs = three planes in flight per SM instead of one, which
    // is what the latency-bound passes of this kernel need; 256 threads each so that the register file still fits them
    static constexpr int NT = (CL_ == 8 && N > 160) ? DLPD_PLANEX_NT_CL8 : ((N > 160) ? DLPD_PLANEX_NT_BIG : ((N >= 128) ? DLPD_PLANEX_NT : 128));
    static constexpr int SMEM = (M * NP + N) * (int)sizeof(cpx);
    static_assert((CL_ == 4 || CL_ == 8) && N % CL_ == 0 && M % 2 == 0 && N % XB == 0, "cluster shape");
    // the S1 plane (L x L) can be staged in the slab (M x N) up to CL = 4; larger clusters fold straight from global memory
    static constexpr bool STAGE = (L * L <= M * N);

    static constexpr int MINB = (SMEM <= 56 * 1024) ? 4 : ((SMEM <= 75 * 1024) ? DLPD_PLANEX_MINB_CL8 : 1);   // resident CTAs per SM the register budget is cut for
    struct FoldS1 {          // element e = j of column l = y of slab q: w_N^{q j} (x[j] + i^q x[j + M])
        const cpx* pl;       // staged S1 plane: pl[j][y] = x[j][y], pl[j][L + y] = x[j + M][y]   (j < M, y < L)
        const cpx* tw;
        int q;
        cpx iq;              // i^q
        DLPD_DEV cpx operator()(int l, int e) const {
            const cpx a = pl[e * NP + l], b = pl[e * NP + L + l];
            const cpx w = tw[e * q];                                   // q = 0: tw[0] = 1
            return cadd(cmul(a, w), cmul(b, cmul(w, iq)));
        }
    };
    struct FoldG {           // the same fold for any CL, reading the S1 plane from global memory (L2: all CL CTAs read it):
        const cpx* in;       //   y_q[j] = w_N^{q j} sum_{m < CL/2} x[j + m M] e^{2 pi i q m / CL}
        const cpx* tw;
        int q;
        DLPD_DEV cpx operator()(int l, int e) const {
            cpx acc = __ldg(in + e * L + l);
            DLPD_UNROLL
            for (int m = 1; m < CL / 2; ++m) acc = cadd(acc, cmul(__ldg(in + (e + m * M) * L + l), tw[(N / CL) * ((q * m) & (CL - 1))]));
            return cmul(acc, tw[e * q]);
        }
    };
    struct MulSpec {         // element e = ky' of local row l (global row kx' = q M + l); spectrum stored [ky'][kx']
        const cpx* spec;
        DLPD_DEV cpx operator()(int l, int e) const { return __ldg(spec + e * N + l); }
    };

    static DLPD_DEV bool exchanges(const Params& p) { return p.spec != nullptr; }

    // everything up to the partial inverse along x: local to the slab
    // cid_next: the work item this cluster handles next (-1: none) -- its S1 plane is pulled into L2 meanwhile
    // pending: the cluster barrier that ended the previous item's phase B has been arrived at but not waited for -- the wait
    // is placed after this item's prefetches, right before the slab is written again
    static DLPD_DEV void phase_a(const Params& p, int cid, int q, char* smem, int cid_next, bool pending = false) {
        cpx* pl = reinterpret_cast<cpx*>(smem);        // [M local rows kx'][y], pitch NP
        cpx* tw = pl + M * NP;
        const int pair = cid % p.P, kz = (cid / p.P) % N, c = cid / (p.P * N);
        const long plane = ((long)pair * p.C + c) * N + kz;
        if (STAGE && pending) cluster_wait();
        // stage the whole S1 plane (L x L, read by all CL CTAs of the cluster) in the still empty slab with asynchronous
        // copies -- every load is in flight at once instead of one dependent round trip per butterfly task: row j < M
        // holds x = j in its left half (columns y < L) and x = j + M in its right half, which the y transform only
        // fills later
        if constexpr (STAGE) {
            const cpx* s1 = p.in + plane * L * L;
            DLPD_FOR_TASKS(t, L * L, NT) {
                const int x = t / L, y = t - x * L;
                cp_async8(pl + (x % M) * NP + (x / M) * L + y, s1 + t);
            }
        }
        const cpx* spec = p.spec ? p.spec + (long)pair * p.spec_pair_stride + ((long)c * N + kz) * N * N + q * M : nullptr;
        if (spec != nullptr) {           // my quarter of the receptor plane (M x 8 bytes per ky' row) -> L2, used 3 passes later
            constexpr int LPR = (M * (int)sizeof(cpx) + 127) / 128 + 1;
            DLPD_FOR_TASKS(t, N * LPR, NT)
                prefetch_l2(reinterpret_cast<const char*>(spec + (long)(t / LPR) * N) + 128 * (t % LPR));
        }
        if (cid_next >= 0) {             // next S1 plane -> L2: each CTA of the cluster takes a quarter
            const long plane2 = ((long)(cid_next % p.P) * p.C + cid_next / (p.P * N)) * N + (cid_next / p.P) % N;
            const char* s1n = reinterpret_cast<const char*>(p.in + plane2 * L * L) + (long)q * (L * L * (int)sizeof(cpx) / CL);
            DLPD_FOR_TASKS(t, L * L * (int)sizeof(cpx) / CL / 128, NT) prefetch_l2(s1n + 128 * t);
        }
        if (!STAGE && pending) cluster_wait();
        // the twiddle table sits behind the slab and survives from item to item of a persistent cluster: copy it (and pay the CTA
        // barrier that publishes it) only for the first item; the STAGE path needs the barrier for its asynchronous copies anyway
        if (STAGE || !pending || !DLPD_PLANEX_KEEP_TW) {
            DLPD_FOR_TASKS(t, N, NT) tw[t] = p.tw[t];
            cp_async_wait_all();
            DLPD_SYNC();
        }
        // forward x on the columns y < L: radix-CL fold in place, then the M-point transform
        if constexpr (STAGE && CL == 4) {
            const cpx iq = q == 0 ? make_float2(1.f, 0.f) : q == 1 ? make_float2(0.f, 1.f) : q == 2 ? make_float2(-1.f, 0.f) : make_float2(0.f, -1.f);
            GlobalIn<FoldS1> in{FoldS1{pl, tw, q, iq}};
            fft_pass<PASS_DIF, XA, 1, L, M, M, N, NP, NT>(pl, tw, LineMul<1>(), in);
        } else {
            GlobalIn<FoldG> in{FoldG{p.in + plane * L * L, tw, q}};
            fft_pass<PASS_DIF, XA, 1, L, M, M, N, NP, NT>(pl, tw, LineMul<1>(), in);
        }
        if constexpr (XB > 1) fft_pass<PASS_DIF, XB, 1, L, M, XB, N, NP, NT>(pl, tw, LineMul<1>());
        fft_fwd_head<N, M, 1, NT>(pl, tw, LineMul<NP>());                       // along y, my rows
        if (p.spec == nullptr) {
            fft_fwd_tail<N, M, 1, NT>(pl, tw, LineMul<NP>());
            cpx* out = p.out + plane * N * N + q * M;
            DLPD_FOR_TASKS(t, M * N, NT) {
                const int l = t % M, ky = t / M;                                // transposed store [ky'][kx']
                const cpx v = pl[l * NP + ky];
                out[ky * N + l] = make_float2(v.x * p.scale, -v.y * p.scale);
            }
            return;
        }
        PointMul<MulSpec> mul{MulSpec{spec}};
        fft_mid<N, M, 1, NT>(pl, tw, LineMul<NP>(), mul);                       // y tail . multiply . inverse y tail
        fft_inv_head<N, M, 1, NT>(pl, tw, LineMul<NP>());                       // inverse y
        if constexpr (XB > 1) fft_pass<PASS_DIT, XB, 1, N, M, XB, N, NP, NT>(pl, tw, LineMul<1>());
#if DLPD_PLANEX_SKIP_LAST_SYNC      // the cluster barrier that follows phase A is a CTA barrier as well: the pass needs none of its own
        fft_pass_team<PASS_DIT, XA, 1, N, M, M, N, NP>(NoSyncTeam<CtaTeam<NT> >{CtaTeam<NT>()}, pl, tw, LineMul<1>());
#else
        fft_pass<PASS_DIT, XA, 1, N, M, M, N, NP, NT>(pl, tw, LineMul<1>());    // A_q = IFFT_M along x, all columns
#endif
    }

    // last inverse-x step across the slabs of the cluster (slabs[r] = CTA r's plane buffer) for the row pairs
    // jp = rank, rank + CL, ...: lanes run along y, so slab reads and S2 stores are contiguous
    static DLPD_DEV void phase_b(const Params& p, int cid, int rank, cpx* const* slabs, char* smem) {
        const cpx* tw = reinterpret_cast<const cpx*>(smem) + M * NP;
        const int pair = cid % p.P, kz = (cid / p.P) % N, c = cid / (p.P * N);
        cpx* out = p.out + (((long)pair * p.C + c) * N + kz) * N * N;
        const int npairs = (M / 2 - rank + CL - 1) / CL;
        DLPD_FOR_TASKS(t, npairs * N, NT) {
            const int y = t % N, j = 2 * (rank + CL * (t / N));
            cpx a0[CL], a1[CL];
            DLPD_UNROLL
            for (int r = 0; r < CL; ++r) {
                a0[r] = slabs[r][j * NP + y];
                a1[r] = slabs[r][(j + 1) * NP + y];
            }
            DLPD_UNROLL
            for (int r = 1; r < CL; ++r) {
                a0[r] = cmul(a0[r], tw[r * j]);
                a1[r] = cmul(a1[r], tw[r * (j + 1)]);
            }
            Dft<CL>::run(a0);
            Dft<CL>::run(a1);
            DLPD_UNROLL
            for (int m = 0; m < CL; ++m) {
                const int x = j + m * M;                                        // even
                if (p.tiled) {
                    store_cpx2(out + ((((x >> 1) * L) + (y >> 1)) << 2) + ((y & 1) << 1), a0[m], a1[m]);
                } else {
                    out[x * N + y] = a0[m];
                    out[(x + 1) * N + y] = a1[m];
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// InvZStore: inverse FFT along z, real part -> real volume (standalone VolumeConvolution operator only).
// ------------------------------------------------------------------------------------------------------
struct InvZParams {
    const cpx* in;     // S2 [Q][N][N][N]  (kz', x, y)
    float* out;        // [Q][N][N][N]     (x, y, z)
    const cpx* tw;
};

template <int L_, int YT_> struct InvZStore {
    typedef InvZParams Params;
    static constexpr int L = L_, N = 2 * L_, YT = YT_, YP = YT_ + 1, YTILES = N / YT_;
    static constexpr int NT = 256;
    static constexpr int SMEM = (N * YP + N) * (int)sizeof(cpx);
    static_assert(N % YT_ == 0, "y tile must divide N");
    // grid: x = N * YTILES, y = Q
    static DLPD_DEV void body(const Params& p, int bx, int by, int, char* smem) {
        cpx* buf = reinterpret_cast<cpx*>(smem);       // [kz][YP]
        cpx* tw = buf + N * YP;
        const int x = bx / YTILES, y0 = (bx % YTILES) * YT;
        DLPD_FOR_TASKS(t, N, NT) tw[t] = p.tw[t];
        const cpx* in = p.in + (long)by * N * N * N + (long)x * N + y0;
        DLPD_FOR_TASKS(t, N * YT, NT) {
            const int kz = t / YT, yl = t % YT;
            buf[kz * YP + yl] = in[(long)kz * N * N + yl];
        }
        DLPD_SYNC();
        fft_inv_tail<N, YT, YP, NT>(buf, tw, LineMul<1>());
        fft_inv_head<N, YT, YP, NT>(buf, tw, LineMul<1>());
        float* out = p.out + (long)by * N * N * N + ((long)x * N + y0) * N;
        DLPD_FOR_TASKS(t, N * YT, NT) {
            const int z = t % N, yl = t / N;
            out[yl * N + z] = buf[z * YP + yl].x;
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// Epilogue
// ------------------------------------------------------------------------------------------------------
// SimpleFilter weights as kernel-parameter constants.  The MLP applies each weight to the two rotations of a pair
// (real / imaginary part of the correlation) with one packed FFMA2 whose weight operand is a broadcast scalar.
template <int C0, int C1, int H> struct FilterW {
    float W1t[(C0 + C1) * H];  // TRANSPOSED nn.Linear weight (DockingModels.py:29): W1t[c][j] = W1[j][c], so the H
                               // weights of one input channel are contiguous (one 128-bit constant load per 4)
    float b1[H];
    float W2[H];               // [1][H]                (DockingModels.py:31)
    float b2;
};
DLPD_DEV float2 dup2(float w) { return make_float2(w, w); }

// ------------------------------------------------------------------------------------------------------
// Tensor-core form of the filter's first layer (level-0 half: C0 = 16 correlation channels -> H hidden units).
// mma.sync.m16n8k16 on fp16 operands with fp32 accumulation, each fp32 factor split into two fp16 terms
//     a = a_hi + a_lo,  w = w_hi + w_lo,   a w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo        (22 significant bits)
// so the result agrees with the FFMA chain to fp32 rounding level.  fp16 has a 5-bit exponent, so every correlation
// channel is first scaled by a power of two 2^-e_c chosen from the Cauchy-Schwarz bound
//     |corr_c[t]| <= ||A_c||_2 ||B_c^rot||_2   (x4 head room for the trilinear resampling of B)
// to |a'| <= 2^15, and the weights by 2^(e_c - f) with f such that max |w'| <= 2^14: all scalings are exact, 2^-f is
// carried through the ReLU into the second layer's weights.  SumSq measures the norms (set_receptor / set_ligand /
// per chunk for pre-rotated ligands), FilterPrep turns them plus the weights into the per-lane B fragments.
// ------------------------------------------------------------------------------------------------------
template <int C0, int H> struct MlpDev {
    float ascale[C0];                        // 2^-e_c
    float W2s[H];                            // W2[j] 2^f
    float hscale, pad_[3];                   // 2^-f: applied to the level-1 hidden contribution (h1) on load
    unsigned int wfrag[(H + 7) / 8][2][2][32];   // [n-tile][hi | lo][B register][lane] packed half2
};

#if DLPD_EXPERIMENTAL
struct SumSqParams {
    const float* v;          // [n][C][V]
    float* out;              // [C]: max over n of sum_v v^2  (pre-zeroed; non-negative floats order like their bits)
    int C, V;
};
struct SumSq {
    typedef SumSqParams Params;
    static constexpr int NT = 256;
    static constexpr int SMEM = 0;
    // grid: x = C, y = n
    static DLPD_DEV void body(const Params& p, int bx, int by, int, char*) {
        const float* v = p.v + ((long)by * p.C + bx) * p.V;
        float acc = 0.0f;
        DLPD_FOR_TASKS(t, p.V, NT) acc = fmaf_(v[t], v[t], acc);
#if defined(DLPD_EMU)
        if (acc > p.out[bx]) p.out[bx] = acc;
#else
        __shared__ float red[NT / 32];
        DLPD_UNROLL
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.0f;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            atomicMax(reinterpret_cast<unsigned int*>(p.out) + bx, __float_as_uint(s));
        }
#endif
    }
};

#if !defined(DLPD_EMU)
DLPD_DEV unsigned int pack_h2(float lo, float hi) {       // low half <- lo, high half <- hi (round to nearest)
    unsigned int r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
DLPD_DEV float h_round(float w) {                          // fp16(w) as fp32
    unsigned short h;
    float r;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(w));
    asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
    return r;
}
DLPD_DEV void mma_f16_16816(float (&d)[4], const unsigned int (&a)[4], unsigned int b0, unsigned int b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// k index of the MMA -> correlation channel: lane q (= lane & 3) holds k = 2q, 2q+1, 2q+8, 2q+9 and reads channels
// q, q+4, q+8, q+12 -- its four loads of a voxel row then hit distinct shared-memory banks across the warp
DLPD_DEV int mlp_channel(int q, int kk) { return q + 4 * kk; }

template <int C0, int C1, int H> struct PrepParams {
    const float* rec_ss;     // [C0] sum of squares of the receptor's level-0 feature volumes
    const float* lig_ss;     // [C0] same for the ligand (max over rotations when it arrives pre-rotated)
    MlpDev<C0, H>* out;
    FilterW<C0, C1, H> filt;
};
template <int C0, int C1, int H> struct FilterPrep {
    typedef PrepParams<C0, C1, H> Params;
    static constexpr int NT = 128;
    static constexpr int SMEM = 0;
    static_assert(C0 == 16 && H % 8 == 0, "m16n8k16 tiling of the first filter layer");
    static DLPD_DEV void body(const Params& p, int, int, int, char*) {
        int e[C0];
        DLPD_UNROLL
        for (int c = 0; c < C0; ++c) {
            const float rs = p.rec_ss[c], ls = p.lig_ss[c];
            int ec = 0;
            if (rs > 0.0f && ls > 0.0f && rs < 3.0e38f && ls < 3.0e38f)
                ec = (int)ceilf(2.0f + 0.5f * (log2f(rs) + log2f(ls))) - 15;
            e[c] = ec < -60 ? -60 : (ec > 60 ? 60 : ec);
        }
        float m = 0.0f;
        for (int c = 0; c < C0; ++c)
            for (int j = 0; j < H; ++j) m = fmaxf(m, fabsf(ldexpf(p.filt.W1t[c * H + j], e[c])));
        int f = 0;
        if (m > 0.0f && m < 3.0e38f) f = (int)ceilf(log2f(m)) - 14;
        f = f < -100 ? -100 : (f > 100 ? 100 : f);
        MlpDev<C0, H>& o = *p.out;
        DLPD_FOR_TASKS(t, C0, NT) o.ascale[t] = ldexpf(1.0f, -e[t]);
        DLPD_FOR_TASKS(t, H, NT) o.W2s[t] = ldexpf(p.filt.W2[t], f);
        if (threadIdx.x == 0) o.hscale = ldexpf(1.0f, -f);
        DLPD_FOR_TASKS(t, (H / 8) * 2 * 32, NT) {
            const int lane = t % 32, r = (t / 32) % 2, nt = t / 64;
            const int g = lane >> 2, q = lane & 3, j = 8 * nt + g;
            const int ca = mlp_channel(q, 2 * r), cb = mlp_channel(q, 2 * r + 1);
            const float wa = ldexpf(p.filt.W1t[ca * H + j], e[ca] - f), wb = ldexpf(p.filt.W1t[cb * H + j], e[cb] - f);
            const float ha = h_round(wa), hb = h_round(wb);
            o.wfrag[nt][0][r][lane] = pack_h2(ha, hb);
            o.wfrag[nt][1][r][lane] = pack_h2(wa - ha, wb - hb);
        }
    }
};
#endif
#endif   // DLPD_EXPERIMENTAL

// order-preserving map float -> uint32 (a < b  <=>  key(a) < key(b)); +-0 collapse to +0 like torch.min
DLPD_HD unsigned int score_key(float s) {
    if (s == 0.0f) s = 0.0f;
    unsigned int u;
    memcpy(&u, &s, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
DLPD_HD float key_score(unsigned int k) {
    unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float s;
    memcpy(&s, &u, 4);
    return s;
}

#if !defined(DLPD_EMU)
template <int NT>
DLPD_DEV unsigned long long block_min_u64(unsigned long long v, unsigned long long* red) {
    DLPD_UNROLL
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < NT / 32 ? red[lane] : ~0ull;
        DLPD_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = w < v ? w : v;
        }
    }
    __syncthreads();
    return v;      // valid in thread 0
}
#endif

// hidden-layer contribution of the level-1 (half resolution) channels for JG hidden units starting at J0, both
// rotations of the pair at once:  out[jj] = b1[J0+jj] + sum_c W1[J0+jj][C0+c] * v[c * CS]
template <int C0, int C1, int H, int JG, int J0, int CS> struct Level1Hidden {
    static DLPD_DEV void run(const FilterW<C0, C1, H>& f, const cpx* v, cpx* out) {
        float2 acc[JG];
        DLPD_UNROLL
        for (int jj = 0; jj < JG; ++jj) acc[jj] = dup2(f.b1[J0 + jj]);
        DLPD_UNROLL
        for (int c = 0; c < C1; ++c) {
            const float2 x = v[c * CS];
            DLPD_UNROLL
            for (int jj = 0; jj < JG; ++jj) acc[jj] = ffma2_(dup2(f.W1t[(C0 + c) * H + J0 + jj]), x, acc[jj]);
        }
        DLPD_UNROLL
        for (int jj = 0; jj < JG; ++jj) out[jj] = acc[jj];
    }
};
template <int C0, int C1, int H, int JG, int G, int CS> struct Level1Dispatch {      // warp-uniform switch over hidden groups
    static DLPD_DEV void run(const FilterW<C0, C1, H>& f, int jg, const cpx* v, cpx* out) {
        if (jg == G) Level1Hidden<C0, C1, H, JG, G * JG, CS>::run(f, v, out);
        else if constexpr (G + 1 < H / JG) Level1Dispatch<C0, C1, H, JG, G + 1, CS>::run(f, jg, v, out);
    }
};

// ------------------------------------------------------------------------------------------------------
// HiddenL1: level-1 (half resolution) part of the filter's hidden layer.
// The first filter layer is linear, so the C1 nearest-up-sampled channels (U[c,t] = corr1[c, t//2], DockingModels.py:74-76)
// contribute ONE hidden pre-activation vector per level-1 voxel, shared by the 8 level-0 voxels above it:
//     h1[x1,y1,z1][j] = b1[j] + sum_c W1[j][C0+c] * corr1[c][x1,y1,z1]
// One CTA: inverse z-FFT of the C1 lines of YT level-1 columns (x1, y0..y0+YT-1), then the C1 -> H contraction.
// The level-1 correlation volumes never reach HBM; only h1 (H complex = 2 rotations per level-1 voxel) does.
// ------------------------------------------------------------------------------------------------------
template <int C0, int C1, int H> struct HidParams {
    const cpx* s2_1;              // [P][C1][N1][N1][N1]  (kz', x1, y1)
    cpx* h1;                      // [P][N1][N1][N1][H]   (x1, y1, z1, j)
    const cpx* tw1;               // [N1]
    FilterW<C0, C1, H> filt;
};

// FK_ = 2: the C1 -> H contraction on tcgen05.mma.kind::tf32 (3 x TF32 split, operand and accumulators in TMEM) like the
// epilogue's level-0 half: one thread per level-1 voxel (yl, z1) of the CTA, K staged 16 channels at a time into the
// warpgroup's 32 A columns (hi | lo), accumulating over the C1 / 16 rounds; two CTAs per SM share the 512 TMEM columns.
template <int L_, int C0, int C1, int H, int FK_ = 0> struct HiddenL1 {
    typedef HidParams<C0, C1, H> Params;
    static constexpr int N1 = L_, YT = (L_ >= 128) ? DLPD_HID_YT_BIG : DLPD_HID_YT, NL = C1 * YT;
    static constexpr int P1 = NL + 4;                        // row pitch: 16-byte rows, conflict-free (yl, z1) reads
    static constexpr int JG = (H % 8 == 0) ? 8 : H;          // hidden units per task
    static constexpr int VP = (YT * N1 + 31) / 32 * 32;      // (yl, z1) padded to whole warps (warp-uniform hidden group)
    static constexpr int NT = (N1 >= 128) ? DLPD_HID_NT_BIG : ((N1 >= 64) ? DLPD_HID_NT : 128);
    static constexpr bool PKT = DLPD_HID_PACKED_TW && (N1 / (2 * LinePlan<N1>::R0) > 1);
    // tcgen05 path
    static constexpr int NVOX = YT * N1, NSC = VP, NWG = (NVOX + 127) / 128, HP = (H + 7) / 8 * 8, KH = 16;
    static constexpr int TC_COLS_WG = 2 * KH + 2 * HP;
    static constexpr int TC_COLS = (NWG * TC_COLS_WG <= 32) ? 32 : (NWG * TC_COLS_WG <= 64) ? 64 : (NWG * TC_COLS_WG <= 128) ? 128 : 256;
    static constexpr bool TC_OK = (FK_ == 2) && (C1 % KH == 0) && (H % 8 == 0) && (NWG * TC_COLS_WG <= 256) && (NSC <= NT) && (NWG <= 6);
    static constexpr int TC_BW = 2 * C1 * HP;                // floats: [hi | lo][C1 / 4 k-chunks][HP rows][4]
    static constexpr int TC_BYTES = TC_OK ? (TC_BW * 4 + 8 * 8 + 16) : 0;
    static constexpr int SMEM = (N1 * P1 + N1 + (PKT ? N1 : 0)) * (int)sizeof(cpx) + TC_BYTES;
    static_assert(N1 % YT == 0, "level-1 y tile");
    struct LoadS2 {          // element e = kz' of line l = c*YT + yl
        const cpx* s;
        DLPD_DEV cpx operator()(int l, int e) const { return __ldg(s + ((long)(l / YT) * N1 + e) * N1 * N1 + (l % YT)); }
    };
#if !defined(DLPD_EMU)
    // h1[voxel][j] = b1[j] + sum_c W1[j][C0 + c] corr1[c][voxel] for the thread's voxel, both rotations of the pair
    static DLPD_DEV void contract_tc(const Params& p, const cpx* b1, cpx* h1, uint32_t tmem_base, uint32_t bw_addr, uint64_t* bars) {
        const int t = (int)threadIdx.x;
        const bool valid = t < NVOX;
        const int tv = valid ? t : 0;
        const int yl = tv % YT, z1 = tv / YT;
        const int g = t >> 7, q = (t >> 5) & 3;
        const int gthreads = (NSC - 128 * g) < 128 ? (NSC - 128 * g) : 128;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t colA = (uint32_t)(g * TC_COLS_WG), colD0 = colA + 2 * KH, colD1 = colD0 + HP;
        constexpr uint32_t IDESC = tc::idesc_tf32(128, HP);
        constexpr uint32_t LBO = HP * 16, SBO = 128, KSTEP = 2 * HP * 16, BLO = C1 * HP * 4;   // bytes
        const cpx* v0 = b1 + z1 * P1 + yl;
        uint64_t* bar = bars + g;
        uint32_t phase = 0;
        bool first = true;
        DLPD_UNROLL
        for (int rot = 0; rot < 2; ++rot) {
            DLPD_UNROLL
            for (int kr = 0; kr < C1 / KH; ++kr) {
                uint32_t hi[KH], lo[KH];
                DLPD_UNROLL
                for (int c = 0; c < KH; ++c) {
                    const cpx xv = v0[(kr * KH + c) * YT];
                    const float a = rot ? xv.y : xv.x, ah = tc::tf32_hi(a);
                    hi[c] = __float_as_uint(ah);
                    lo[c] = __float_as_uint(a - ah);
                }
                if (!first) {                                 // the A columns are free once the previous round's MMAs have completed
                    tc::mbar_wait(bar, phase);
                    tc::fence_after_sync();
                    phase ^= 1;
                }
                first = false;
                DLPD_UNROLL
                for (int h8 = 0; h8 < KH / 8; ++h8) {
                    uint32_t t0[8], t1[8];
                    DLPD_UNROLL
                    for (int i = 0; i < 8; ++i) { t0[i] = hi[8 * h8 + i]; t1[i] = lo[8 * h8 + i]; }
                    tc::st8(lane_addr + colA + 8 * h8, t0);
                    tc::st8(lane_addr + colA + KH + 8 * h8, t1);
                }
                tc::wait_st();
                tc::fence_before_sync();
                named_barrier_sync(2 + g, gthreads);
                if (q == 0 && tc::elect_one()) {
                    tc::fence_after_sync();
                    const uint32_t d = tmem_base + (rot ? colD1 : colD0), a = tmem_base + colA;
                    DLPD_UNROLL
                    for (int s = 0; s < KH / 8; ++s) {
                        const int ks = kr * (KH / 8) + s;                          // 8-channel step within the C1 channels
                        const uint64_t dhi = tc::smem_desc_kmajor(bw_addr + ks * KSTEP, LBO, SBO);
                        const uint64_t dlo = tc::smem_desc_kmajor(bw_addr + BLO + ks * KSTEP, LBO, SBO);
                        tc::mma_tf32_ts(d, a + KH + 8 * s, dhi, IDESC, (kr > 0 || s > 0));      // a_lo w_hi
                        tc::mma_tf32_ts(d, a + 8 * s, dlo, IDESC, 1);                           // a_hi w_lo
                        tc::mma_tf32_ts(d, a + 8 * s, dhi, IDESC, 1);                           // a_hi w_hi
                    }
                    tc::mma_commit(bar);
                }
                __syncwarp();
            }
        }
        tc::mbar_wait(bar, phase);                            // the last round: both accumulators are complete
        tc::fence_after_sync();
        float acc[2][HP];
        DLPD_UNROLL
        for (int rot = 0; rot < 2; ++rot) {
            uint32_t r[HP / 8][8];
            DLPD_UNROLL
            for (int u = 0; u < HP / 8; ++u) tc::ld8(lane_addr + (rot ? colD1 : colD0) + 8 * u, r[u]);
            tc::wait_ld();
            DLPD_UNROLL
            for (int u = 0; u < HP / 8; ++u) {
                DLPD_UNROLL
                for (int i = 0; i < 8; ++i) acc[rot][8 * u + i] = __uint_as_float(r[u][i]);
            }
        }
        if (valid) {
            cpx* d = h1 + ((long)yl * N1 + z1) * H;
            DLPD_UNROLL
            for (int j = 0; j < H; j += 2)
                store_cpx2(d + j, make_float2(acc[0][j] + p.filt.b1[j], acc[1][j] + p.filt.b1[j]),
                           make_float2(acc[0][j + 1] + p.filt.b1[j + 1], acc[1][j + 1] + p.filt.b1[j + 1]));
        }
    }
#endif

    // grid: x = N1 / YT, y = N1 (x1), z = P
    static DLPD_DEV void body(const Params& p, int bx, int by, int bz, char* smem) {
        cpx* b1 = reinterpret_cast<cpx*>(smem);        // [kz1][P1]   line = c*YT + yl
        cpx* tw1 = b1 + N1 * P1;
        const int y0 = YT * bx, x1 = by, pair = bz;
        DLPD_FOR_TASKS(t, N1, NT) tw1[t] = p.tw1[t];
        const cpx* htw1 = tw1;
        if constexpr (PKT) {
            cpx* pk = tw1 + N1;
            DLPD_FOR_TASKS(t, N1, NT) pk[t] = packed_head_twiddle<N1, LinePlan<N1>::R0>(p.tw1, t);
            htw1 = pk;
        }
        const CtaTeam<NT, true> tm;
#if !defined(DLPD_EMU)
        uint32_t tmem_base = 0;
        float* Bw = reinterpret_cast<float*>(tw1 + N1 + (PKT ? N1 : 0));
        uint64_t* bars = reinterpret_cast<uint64_t*>(Bw + TC_BW);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
        if constexpr (TC_OK) {
            DLPD_FOR_TASKS(i, TC_BW, NT) {
                const int e = i & 3, n = (i >> 2) % HP, kc = (i / (4 * HP)) % (C1 / 4), part = i / (C1 * HP);
                const int c = 4 * kc + e;
                const float v = n < H ? p.filt.W1t[(C0 + c) * H + n] : 0.0f;
                const float hi = tc::tf32_hi(v);
                Bw[i] = part == 0 ? hi : (v - hi);
            }
            if (threadIdx.x < 32) tc::tmem_alloc(tmem_slot, TC_COLS);
            if (threadIdx.x == 0) {
                for (int i = 0; i < NWG; ++i) tc::mbar_init(bars + i, 1);
                tc::mbar_fence_init();
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();           // published by the barrier that follows the tile load below
        }
#endif
#if DLPD_HID_DIRECT
        {
            // the first inverse pass reads its inputs straight from global memory (two tasks = 2 x TAIL loads in flight per
            // thread) instead of staging the tile in shared memory first: no cp.async issue loop, one shared-memory
            // round trip less
            DLPD_SYNC();
            GlobalIn<LoadS2> in{LoadS2{p.s2_1 + (long)pair * C1 * N1 * N1 * N1 + (long)x1 * N1 + y0}};
            team_inv_tail<N1, NL, P1>(tm, b1, tw1, LineMul<1>(), in);
        }
#else
        {
            const cpx* s = p.s2_1 + (long)pair * C1 * N1 * N1 * N1 + (long)x1 * N1 + y0;
#if defined(DLPD_EMU)
            DLPD_FOR_TASKS(t, C1 * N1 * (YT / 2), NT) {
                const int h = t % (YT / 2), q = t / (YT / 2), kz = q % N1, c = q / N1;
                cp_async16(b1 + kz * P1 + c * YT + 2 * h, s + (long)q * N1 * N1 + 2 * h);     // q = c*N1 + kz
            }
#else
            if constexpr (NT % (YT / 2) == 0) {
                // thread t copies rows q = t / (YT/2), + QS, + 2 QS, ...: (c, kz) and both addresses advance by constants
                constexpr int HT = YT / 2, QS = NT / HT;
                const int h = (int)threadIdx.x % HT;
                int q = (int)threadIdx.x / HT, c = q / N1, kz = q - c * N1;
                const cpx* src = s + (long)q * N1 * N1 + 2 * h;
                for (; q < C1 * N1; q += QS) {
                    cp_async16(b1 + kz * P1 + c * YT + 2 * h, src);
                    src += (long)QS * N1 * N1;
                    kz += QS % N1;
                    c += QS / N1;
                    if (kz >= N1) { kz -= N1; ++c; }
                }
            } else {
                DLPD_FOR_TASKS(t, C1 * N1 * (YT / 2), NT) {
                    const int h = t % (YT / 2), q = t / (YT / 2), kz = q % N1, c = q / N1;
                    cp_async16(b1 + kz * P1 + c * YT + 2 * h, s + (long)q * N1 * N1 + 2 * h);
                }
            }
#endif
        }
        cp_async_wait_all();
        DLPD_SYNC();
        team_inv_tail<N1, NL, P1>(tm, b1, tw1, LineMul<1>());
#endif
        team_inv_head<N1, NL, P1, LinePlan<N1>, PKT>(tm, b1, htw1, LineMul<1>());
        cpx* h1 = p.h1 + (((long)pair * N1 + x1) * N1 + y0) * N1 * H;
#if !defined(DLPD_EMU)
        if constexpr (TC_OK) {
            tc::fence_after_sync();
            tmem_base = *tmem_slot;
            if (threadIdx.x < NSC) contract_tc(p, b1, h1, tmem_base, tc::smem_u32(Bw), bars);
            tc::fence_before_sync();
            __syncthreads();
            if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base, TC_COLS);
            return;
        }
#endif
        DLPD_FOR_TASKS(t, (H / JG) * VP, NT) {
            const int v = t % VP, jg = t / VP;
            if (v < YT * N1) {
                const int yl = v % YT, z1 = v / YT;
                cpx out[JG];
                Level1Dispatch<C0, C1, H, JG, 0, YT>::run(p.filt, jg, b1 + z1 * P1 + yl, out);
                cpx* d = h1 + ((long)yl * N1 + z1) * H + jg * JG;
                if constexpr (JG % 2 == 0 && H % 2 == 0) {           // 16-byte stores: half as many L2 write transactions
                    DLPD_UNROLL
                    for (int jj = 0; jj < JG; jj += 2) store_cpx2(d + jj, out[jj], out[jj + 1]);
                } else {
                    DLPD_UNROLL
                    for (int jj = 0; jj < JG; ++jj) d[jj] = out[jj];
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// Epilogue (persistent): per 2x2 (x,y) tile and all z: inverse z-FFT of the 4 (C0+1) level-0 lines, filter
// (DockingModels.py:28-32,79-82) with the level-1 hidden contribution from HiddenL1, clash mask (Docker.py:226,232),
// running argmin (Docker.py:91-97).  One CTA per SM walks over tiles; the lines of the NEXT tile are copied into the
// second shared-memory buffer (cp.async) while the current tile is transformed and scored, so no load is exposed.
// ------------------------------------------------------------------------------------------------------
template <int C0, int C1, int H> struct EpiParams {
    // TMA descriptor (CUtensorMap, 128 bytes) of the tiled S2 array viewed as a rank-5 fp32 tensor
    //   (8 floats of a 2x2 tile, channel, kz', tile, pair): one box = all channels and kz' of one tile, landing in shared
    // memory as [kz'][channel][8] -- exactly the line-interleaved layout the transform passes use.  Built by the host.
    alignas(64) unsigned long long tmap[16];
    const cpx* s2_0;              // [P][C0+1][N][N/2][N/2][2][2]   (kz', x/2, y/2, y&1, x&1); channel C0 = clash
    const cpx* h1;                // [P][N1][N1][N1][H]
    unsigned long long* best;     // [n_rot] packed (score_key << 32 | linear index), pre-set to ~0
    float* score_out;             // optional [n_rot][N^3] masked score volume (module-level drop-in / tests)
    // optional candidate list for the global top-K (Docker.update_top, Docker.py:86-105): every translation whose masked
    // score is <= the caller's current K-th best is appended as (score bits, rotation, linear index, 0)
    int* cand;                    // [cand_cap][4] or nullptr
    unsigned int* cand_count;     // total number of qualifying translations (may exceed cand_cap: caller re-runs exactly)
    unsigned int cand_cap, cand_key;
    int rot_base;                 // index of this chunk's first rotation within the caller's batch
    const MlpDev<C0, H>* mlp;     // tensor-core form of the first filter layer (FilterPrep), or nullptr: FP32 FFMA chain
    const cpx* tw0;               // [N]
    int n_rot, n_tiles, grid;     // tiles = P * (N/2)^2; grid = number of CTAs walking over them
    float tau;
    FilterW<C0, C1, H> filt;
};

// FK_: how the level-0 half of the filter's first layer (C0 -> H, DockingModels.py:29) is evaluated
//   0 = packed FP32 FFMA2 chain;  1 = mma.sync fp16 split (DLPD_EXPERIMENTAL only);
//   2 = tcgen05.mma kind::tf32 (3 x TF32 split) with the A operand and the accumulators in TMEM (dlpd_tc.h)
// DUO_ (tcgen05 + 2x1 tiles only): TWO persistent CTAs per SM, each on its own stream of half-tiles.  The transform passes of
//   one CTA (FP32 pipe, shared memory) then overlap the TMEM staging / MMA waits of the other -- the overlap the round-1
//   EpilogueDuo could not get while the filter itself was FP32-bound.  The level-1 hidden part is read straight from global
//   memory (L2) instead of being staged, so that two CTAs fit the SM's shared memory.
template <int L_, int C0, int C1, int H, int TY_ = 2, int FK_ = 0, bool DUO_ = false> struct Epilogue {
    typedef EpiParams<C0, C1, H> Params;
    static constexpr int L = L_, N = 2 * L_, N1 = L_, T = L_;      // T = N/2 tile rows along x
    // tile = 2 x TY (x,y) columns: 2x2 normally; 2x1 when two 2x2 buffers do not fit shared memory (N = 256) -- the
    // level-1 column (x/2, y/2) is then fetched by both y-halves
    static constexpr int TY = TY_, XY = 2 * TY_, TPR = (TY_ == 2) ? T : N;     // TPR tiles per x-pair row
    static constexpr int NL0 = XY * (C0 + 1);      // level-0 lines of a tile: line = c*XY + (y&1)*2 + (x&1)
    static constexpr bool TC_ = (FK_ == 1);
    // tcgen05 path: ONE consumer thread per translation of the tile, in warpgroups of 128 (= the 128 TMEM lanes of an
    // M = 128 instruction).  Per warpgroup the TMEM columns hold [A: C0 hi | C0 lo][D rotation A: HP][D rotation B: HP].
    static constexpr int NVOX = XY * N, NWG = (NVOX + 127) / 128, HP = (H + 7) / 8 * 8;
    static constexpr int TC_COLS_WG = 2 * C0 + 2 * HP;
    static constexpr int TC_COLS = (NWG * TC_COLS_WG <= 32) ? 32 : (NWG * TC_COLS_WG <= 64) ? 64 : (NWG * TC_COLS_WG <= 128) ? 128
                                 : (NWG * TC_COLS_WG <= 256) ? 256 : 512;
    static constexpr bool TC_OK = (FK_ == 2) && (C0 % 8 == 0) && (H % 8 == 0) && (H <= 256) && (NWG * TC_COLS_WG <= 512) && (NWG <= 7);
    static constexpr bool DUO = DUO_ && TC_OK && (TY_ == 1) && (NWG * TC_COLS_WG <= 256);
    static constexpr int MINB = DUO ? 2 : 1;                       // resident CTAs per SM the register budget is cut for
    // warp specialisation: the last NPROD threads only issue the cp.async copies of the NEXT tile (each copy touches
    // its own cache line, so issuing them is slow -- but it now overlaps the transform of the current tile instead of
    // stalling every warp); the first NC threads transform and score
    // USE_TMA (tcgen05 kernels): no producer threads at all -- thread 0 issues one tensor-map box load and one bulk copy per
    // tile, completion is counted on an mbarrier; NX extra threads only help with the transform passes
    static constexpr bool USE_TMA = TC_OK && !DUO && (DLPD_EPI_TMA != 0);
    static constexpr int NPROD = USE_TMA ? 0 : (DUO ? 32 : ((N >= 128) ? DLPD_EPI_NPROD : 32));
    static constexpr int NX = DUO ? ((N >= 128) ? 32 : 0) : ((TC_OK && N >= 128) ? DLPD_EPI_TC_EXTRA : 0);
    // tcgen05: the scoring threads are the whole warps that hold a translation (the last warpgroup may be partial)
    static constexpr int NC = TC_OK ? (NVOX + 31) / 32 * 32 : ((N >= 128) ? DLPD_EPI_NT : 128) - NPROD;
    static constexpr int NT = NC + NX + NPROD;
    static constexpr int BUF = N * NL0, HB = N1 * H;
    static constexpr int HBS = DUO ? 0 : HB;                       // staged level-1 hidden column (none in DUO: read from global)
    static constexpr bool BULK_H1 = TC_OK && !USE_TMA && (HBS > 0) && (DLPD_EPI_BULK_H1 != 0) && ((HB * (int)sizeof(cpx)) % 16 == 0);
    // level-1 column buffers: two (the next tile's column is prefetched) or, where two do not fit next to the tile buffers
    // (wide models), ONE that the bulk copy refills at the top of every tile while the transform passes run
    static constexpr int SMEM2 = (2 * BUF + 2 * HBS + N) * (int)sizeof(cpx) + 64 * 8 + 16 * 1024;      // estimate with two
    static constexpr int NHB = (BULK_H1 && SMEM2 > 227 * 1024) ? 1 : 2;
#if DLPD_EXPERIMENTAL
    static constexpr int MLP_BYTES = TC_ ? ((int)sizeof(MlpDev<C0, H>) + 15) / 16 * 16 : 0;      // shared-memory copy of the MMA fragments
#else
    static constexpr int MLP_BYTES = 0;
#endif
    // tcgen05: first-layer weights as the B operand, [hi | lo][C0 / 4 k-chunks][HP rows][4] TF32 (K-major, no swizzle),
    // then 2 mbarriers per warpgroup and the TMEM base address
    static constexpr int TC_BW = 2 * C0 * HP;                      // floats
    static constexpr int TC_BYTES = TC_OK ? (TC_BW * 4 + 16 * 8 + 16) : 0;       // bars[0 .. 2 NWG) MMA, bars[14], bars[15] tile landed
    static constexpr bool PKT = DLPD_EPI_PACKED_TW && (N / (2 * LinePlan<N>::R0) > 1);      // packed twiddles of the inverse head pass
    static constexpr int SMEM = (2 * BUF + NHB * HBS + N + (PKT ? N : 0)) * (int)sizeof(cpx) + 64 * 8 + MLP_BYTES + TC_BYTES;
    static constexpr int QPT = NPROD / TY;         // (c, kz) rows covered per sweep of the producer threads
    static constexpr int TILES = T * TPR;          // tiles per pair
    static_assert(HB % 2 == 0 && NPROD % 32 == 0 && NC % 32 == 0 && (TY_ == 1 || TY_ == 2), "team shapes");

    // tile r of a pair -> x/2, first y
    static DLPD_DEV int tile_xh(int r) { return r / TPR; }
    static DLPD_DEV int tile_y0(int r) { return (TY == 2) ? 2 * (r % TPR) : r % TPR; }

    static DLPD_DEV void issue_loads(const Params& p, int tile, cpx* b0, cpx* hb) {
        const int pair = tile / TILES, r = tile % TILES;
        const int r4 = tile_xh(r) * T + (tile_y0(r) >> 1);              // 2x2 group (x/2) * T + (y/2) of the tiled S2
        const cpx* s = p.s2_0 + ((long)pair * (C0 + 1) * N * T * T + r4) * 4 + ((TY == 1) ? 2 * (tile_y0(r) & 1) : 0);
        DLPD_FOR_TASKS_SUB(t, NPROD, NPROD, NC + NX) {
            const int h = (TY == 2) ? (t & 1) : 0, q0 = t / TY;
            int c = q0 / N, kz = q0 % N;
            const cpx* src = s + ((long)q0 * (T * T) * 4 + 2 * h);      // q = c*N + kz: 16 TY contiguous bytes per row
            for (int q = q0; q < (C0 + 1) * N; q += QPT) {
                cp_async16(b0 + kz * NL0 + c * XY + 2 * h, src);
                src += (long)QPT * T * T * 4;
                kz += QPT % N;
                c += QPT / N;
                if (kz >= N) { kz -= N; ++c; }
            }
        }
        if constexpr (HBS > 0 && !BULK_H1) {
            const cpx* hs = p.h1 + ((long)pair * T * T + r4) * HB;      // level-1 column (x/2, y/2): N1*H contiguous
            DLPD_FOR_TASKS_SUB(t, HB / 2, NPROD, NC + NX) cp_async16(hb + 2 * t, hs + 2 * t);
        }
    }

#if !defined(DLPD_EMU)
    // level-1 hidden column of a tile by one bulk copy (TMA engine, contiguous 15 KB): issued by ONE thread, completes on `bar`
    static DLPD_DEV void issue_h1_bulk(const Params& p, int tile, cpx* hb, uint64_t* bar) {
        const int pair = tile / TILES, r = tile % TILES;
        const int r4 = tile_xh(r) * T + (tile_y0(r) >> 1);
        tc::mbar_expect_tx(bar, (uint32_t)(HB * sizeof(cpx)));
        tc::bulk_g2s(hb, p.h1 + ((long)pair * T * T + r4) * HB, (uint32_t)(HB * sizeof(cpx)), bar);
    }
    // the same loads as ONE tensor-map box (all channels and kz' of the tile -> [kz'][c][8 floats]) and ONE bulk copy (the
    // level-1 column), issued by a single thread; both complete on `bar`
    static DLPD_DEV void issue_tma(const Params& p, int tile, cpx* b0, cpx* hb, uint64_t* bar) {
        const int pair = tile / TILES, r = tile % TILES;
        const int r4 = tile_xh(r) * T + (tile_y0(r) >> 1);
        tc::mbar_expect_tx(bar, (uint32_t)((BUF + HB) * sizeof(cpx)));
        tc::tma_load_5d(b0, p.tmap, bar, (TY == 1) ? 4 * (tile_y0(r) & 1) : 0, 0, 0, r4, pair);
        tc::bulk_g2s(hb, p.h1 + ((long)pair * T * T + r4) * HB, (uint32_t)(HB * sizeof(cpx)), bar);
    }
#endif

    static DLPD_DEV void flush_best(const Params& p, int pair, unsigned long long bestA, unsigned long long bestB,
                                    unsigned long long* red) {
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
#if defined(DLPD_EMU)
        (void)red;
        if (bestA < p.best[rotA]) p.best[rotA] = bestA;
        if (hasB && bestB < p.best[rotB]) p.best[rotB] = bestB;
#else
        bestA = block_min_u64<NT>(bestA, red);
        bestB = block_min_u64<NT>(bestB, red + 32);
        if (threadIdx.x == 0) {
            atomicMin(p.best + rotA, bestA);
            if (hasB) atomicMin(p.best + rotB, bestB);
        }
#endif
    }

    static DLPD_DEV void emit_candidate(const Params& p, float v, int rot, unsigned int lin) {
        const unsigned int slot = atomic_inc_u32(p.cand_count);
        if (slot < p.cand_cap) {
            int bits;
            memcpy(&bits, &v, 4);
            int* d = p.cand + 4 * (long)slot;
            d[0] = bits; d[1] = p.rot_base + rot; d[2] = (int)lin; d[3] = 0;
        }
    }

    // transform + score one tile (consumer team)
    template <class TM>
    static DLPD_DEV void score_tile(const Params& p, const TM& tm, cpx* b0, const cpx* h1, const cpx* tw0, const MlpDev<C0, H>* ms,
                                    int pair, int x0, int y0, unsigned long long& bestA, unsigned long long& bestB) {
        if constexpr (PKT) transform_tile_pk(tm, b0, tw0, tw0 + N);
        else transform_tile(tm, b0, tw0);
#if !defined(DLPD_EMU) && DLPD_EXPERIMENTAL
        if constexpr (MMA_OK) {
            if (p.mlp != nullptr) {
                score_voxels_mma(p, tm, b0, h1, *ms, pair, x0, y0, bestA, bestB);
                return;
            }
        }
#endif
        (void)ms;
        score_voxels(p, tm, b0, h1, pair, x0, y0, bestA, bestB);
    }
    // TC_: this instantiation carries the mma.sync filter path (a separate kernel, so the FP32 kernel keeps its registers)
    static constexpr bool MMA_OK = TC_ && (C0 == 16) && (H % 8 == 0) && ((XY * N) % 16 == 0) && (NC % 32 == 0);

#if !defined(DLPD_EMU) && DLPD_EXPERIMENTAL
    // The same filter with the level-0 half of the first layer on the tensor cores (see MlpDev).  One warp per 16 voxels
    // (M) x 2 rotations; K = 16 channels; N = H hidden units in tiles of 8.  Lane (g = lane / 4, q = lane % 4) owns voxel
    // rows g and g + 8 in the A / C fragments; after the layer-2 partial sums are reduced over the quad, lane q finishes
    // (row q / 2, rotation q % 2): clash mask, argmin key, optional score / candidate output.
    template <class TM>
    static DLPD_DEV void score_voxels_mma(const Params& p, const TM& tm, const cpx* b0, const cpx* h1, const MlpDev<C0, H>& m,
                                          int pair, int x0, int y0, unsigned long long& bestA, unsigned long long& bestB) {
        constexpr int NTL = H / 8, NMT = XY * N / 16;
        const int lane = tm.tid() & 31, warp = tm.tid() >> 5, g = lane >> 2, q = lane & 3;
        // m is the CTA's shared-memory copy: the B fragments stay in registers for the whole tile, the scales and the
        // second-layer weights are re-read where they are used (short live ranges: no spills at 768 threads)
        unsigned int wf[NTL][2][2];
        DLPD_UNROLL
        for (int nt = 0; nt < NTL; ++nt) {
            DLPD_UNROLL
            for (int s = 0; s < 2; ++s) {
                wf[nt][s][0] = m.wfrag[nt][s][0][lane];
                wf[nt][s][1] = m.wfrag[nt][s][1][lane];
            }
        }
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        for (int mt = warp; mt < NMT; mt += TM::NT / 32) {
            int zr[2], xyr[2];
            float2 x[2][4];
            DLPD_UNROLL
            for (int r = 0; r < 2; ++r) {
                const int v = 16 * mt + g + 8 * r;
                xyr[r] = v % XY;
                zr[r] = v / XY;
                DLPD_UNROLL
                for (int kk = 0; kk < 4; ++kk) x[r][kk] = b0[zr[r] * NL0 + mlp_channel(q, kk) * XY + xyr[r]];
            }
            // accumulators start from the level-1 hidden contribution (+ b1) of the voxel's level-1 parent
            float acc[2][NTL][4];
            DLPD_UNROLL
            for (int r = 0; r < 2; ++r) {
                const cpx* hz = h1 + (zr[r] >> 1) * H + 2 * q;
                DLPD_UNROLL
                for (int nt = 0; nt < NTL; ++nt) {
                    const float4 hv = *reinterpret_cast<const float4*>(hz + 8 * nt);     // (j: rotA, rotB), (j+1: rotA, rotB)
                    const float hs = m.hscale;
                    acc[0][nt][2 * r] = hv.x * hs;
                    acc[1][nt][2 * r] = hv.y * hs;
                    acc[0][nt][2 * r + 1] = hv.z * hs;
                    acc[1][nt][2 * r + 1] = hv.w * hs;
                }
            }
            // scale, split into fp16 hi + lo, pack the A fragments of both rotations
            unsigned int ah[2][4], al[2][4];
            {
                float2 xh[2][4], xl[2][4];
                DLPD_UNROLL
                for (int r = 0; r < 2; ++r) {
                    DLPD_UNROLL
                    for (int kk = 0; kk < 4; ++kk) {
                        const float2 xs = fmul2_(x[r][kk], dup2(m.ascale[mlp_channel(q, kk)]));
                        xh[r][kk] = make_float2(__uint_as_float(__float_as_uint(xs.x) & 0xffffe000u),
                                                __uint_as_float(__float_as_uint(xs.y) & 0xffffe000u));
                        xl[r][kk] = fadd2_(xs, make_float2(-xh[r][kk].x, -xh[r][kk].y));
                    }
                }
                DLPD_UNROLL
                for (int r = 0; r < 2; ++r) {
                    DLPD_UNROLL
                    for (int h = 0; h < 2; ++h) {          // register r + 2h: row g + 8r, k = 2q + 8h, 2q + 8h + 1
                        ah[0][r + 2 * h] = pack_h2(xh[r][2 * h].x, xh[r][2 * h + 1].x);
                        ah[1][r + 2 * h] = pack_h2(xh[r][2 * h].y, xh[r][2 * h + 1].y);
                        al[0][r + 2 * h] = pack_h2(xl[r][2 * h].x, xl[r][2 * h + 1].x);
                        al[1][r + 2 * h] = pack_h2(xl[r][2 * h].y, xl[r][2 * h + 1].y);
                    }
                }
            }
            // three products per term; the 2 * NTL accumulators are independent, so consecutive MMAs never wait on each other
            DLPD_UNROLL
            for (int rot = 0; rot < 2; ++rot) {
                DLPD_UNROLL
                for (int nt = 0; nt < NTL; ++nt) mma_f16_16816(acc[rot][nt], al[rot], wf[nt][0][0], wf[nt][0][1]);
            }
            DLPD_UNROLL
            for (int rot = 0; rot < 2; ++rot) {
                DLPD_UNROLL
                for (int nt = 0; nt < NTL; ++nt) mma_f16_16816(acc[rot][nt], ah[rot], wf[nt][1][0], wf[nt][1][1]);
            }
            DLPD_UNROLL
            for (int rot = 0; rot < 2; ++rot) {
                DLPD_UNROLL
                for (int nt = 0; nt < NTL; ++nt) mma_f16_16816(acc[rot][nt], ah[rot], wf[nt][0][0], wf[nt][0][1]);
            }
            // ReLU + second layer: partial sums over this lane's 2 * NTL hidden units, reduced over the quad
            float s[2][2];
            DLPD_UNROLL
            for (int rot = 0; rot < 2; ++rot) {
                DLPD_UNROLL
                for (int r = 0; r < 2; ++r) {
                    float t = 0.0f;
                    DLPD_UNROLL
                    for (int nt = 0; nt < NTL; ++nt) {
                        const float2 w2 = *reinterpret_cast<const float2*>(&m.W2s[8 * nt + 2 * q]);
                        t = fmaf_(w2.x, fmaxf(acc[rot][nt][2 * r], 0.0f), t);
                        t = fmaf_(w2.y, fmaxf(acc[rot][nt][2 * r + 1], 0.0f), t);
                    }
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    s[rot][r] = t;
                }
            }
            const int r = q >> 1, rot = q & 1;
            const float sc = p.filt.b2 + (r ? (rot ? s[1][1] : s[0][1]) : (rot ? s[1][0] : s[0][0]));
            const int z = r ? zr[1] : zr[0], xy = r ? xyr[1] : xyr[0];
            const cpx clash = b0[z * NL0 + C0 * XY + xy];
            const float V = (rot ? clash.y : clash.x) < p.tau ? sc : 0.0f;     // (clash < tau).float() * score   (Docker.py:226,232)
            const unsigned int lin = (unsigned int)(((x0 + (xy & 1)) * N + (y0 + (xy >> 1))) * N + z);
            const unsigned long long key = ((unsigned long long)score_key(V) << 32) | lin;
            if (rot) bestB = key < bestB ? key : bestB;
            else bestA = key < bestA ? key : bestA;
            if (rot == 0 || hasB) {
                if (p.score_out != nullptr) p.score_out[(long)(rot ? rotB : rotA) * N * N * N + lin] = V;
                if (p.cand != nullptr && (unsigned int)(key >> 32) <= p.cand_key) emit_candidate(p, V, rot ? rotB : rotA, lin);
            }
        }
    }
#endif

    // inverse z transform of the NL0 lines of a tile, in place
    template <class TM> static DLPD_DEV void transform_tile(const TM& tm, cpx* b0, const cpx* tw0) {
#if !defined(DLPD_ABLATE_EPI_FFT)
        team_inv_tail<N, NL0, NL0>(tm, b0, tw0, LineMul<1>());
        team_inv_head<N, NL0, NL0>(tm, b0, tw0, LineMul<1>());
#endif
    }
    // the same with the head pass reading its twiddles from the packed table htw (classic kernel, PKT)
    template <class TM> static DLPD_DEV void transform_tile_pk(const TM& tm, cpx* b0, const cpx* tw0, const cpx* htw) {
#if !defined(DLPD_ABLATE_EPI_FFT)
        team_inv_tail<N, NL0, NL0>(tm, b0, tw0, LineMul<1>());
        team_inv_head<N, NL0, NL0, LinePlan<N>, PKT>(tm, b0, htw, LineMul<1>());
#endif
    }

    // filter, clash mask and running argmin over the XY * N voxels of a transformed tile
    template <class TM>
    static DLPD_DEV void score_voxels(const Params& p, const TM& tm, const cpx* b0, const cpx* h1, int pair, int x0, int y0,
                                      unsigned long long& bestA, unsigned long long& bestB) {
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        DLPD_TEAM_TASKS(tm, t, XY * N) {
            const int xy = t % XY, z = t / XY;
            const cpx* v0 = b0 + z * NL0 + xy;
            const cpx* hz = h1 + (z >> 1) * H;
            float2 h[H];                                        // (rotA, rotB) per hidden unit
            DLPD_UNROLL
            for (int j = 0; j < H; ++j) h[j] = hz[j];
#if defined(DLPD_ABLATE_EPI_MLP)
            h[0] = v0[0];
#else
            DLPD_UNROLL
            for (int c = 0; c < C0; ++c) {
                const float2 v = v0[c * XY];
                DLPD_UNROLL
                for (int j = 0; j < H; ++j) h[j] = ffma2_(dup2(p.filt.W1t[c * H + j]), v, h[j]);
            }
#endif
            float2 sc = dup2(p.filt.b2);
            DLPD_UNROLL
            for (int j = 0; j < H; ++j) sc = ffma2_(dup2(p.filt.W2[j]), make_float2(fmaxf(h[j].x, 0.0f), fmaxf(h[j].y, 0.0f)), sc);
            const unsigned int lin = (unsigned int)(((x0 + (xy & 1)) * N + (y0 + (xy >> 1))) * N + z);
            finish_voxel(p, sc, v0[C0 * XY], lin, rotA, rotB, hasB, bestA, bestB);
        }
    }

    // clash mask (Docker.py:226,232), running argmin (Docker.py:91-97) and the optional score-volume / candidate outputs of
    // one translation: sc = filter output of the two rotations of the pair, clash = their clash correlations
    static DLPD_DEV void finish_voxel(const Params& p, float2 sc, cpx clash, unsigned int lin, int rotA, int rotB, bool hasB,
                                      unsigned long long& bestA, unsigned long long& bestB) {
        const float Va = clash.x < p.tau ? sc.x : 0.0f;     // (clash < tau).float() * score
        const float Vb = clash.y < p.tau ? sc.y : 0.0f;
        const unsigned long long ka = ((unsigned long long)score_key(Va) << 32) | lin;
        const unsigned long long kb = ((unsigned long long)score_key(Vb) << 32) | lin;
        bestA = ka < bestA ? ka : bestA;
        bestB = kb < bestB ? kb : bestB;
        if (p.score_out != nullptr) {
            p.score_out[(long)rotA * N * N * N + lin] = Va;
            if (hasB) p.score_out[(long)rotB * N * N * N + lin] = Vb;
        }
        if (p.cand != nullptr) {
            if ((unsigned int)(ka >> 32) <= p.cand_key) emit_candidate(p, Va, rotA, lin);
            if (hasB && (unsigned int)(kb >> 32) <= p.cand_key) emit_candidate(p, Vb, rotB, lin);
        }
    }

#if !defined(DLPD_EMU)
    // ---- tcgen05 form of the same filter (FK_ == 2) --------------------------------------------------------------------
    // Thread t of the consumer team owns translation t of the tile (xy = t % XY, z = t / XY) and TMEM lane t % 128 of its
    // warpgroup g = t / 128.  Per rotation of the pair (the real / imaginary parts of the correlation values):
    //   1. split the C0 channel values into TF32 hi + lo (exact), tcgen05.st them into the warpgroup's A columns;
    //   2. one elected thread issues, per 8 channels, D (+)= A_lo W_hi^T, A_hi W_lo^T, A_hi W_hi^T (M = 128, N = H, K = 8) and
    //      commits to the warpgroup's mbarrier;
    //   3. the accumulator row comes back with tcgen05.ld; + level-1 hidden part (h1), ReLU, second layer, mask, argmin.
    // The A columns are shared by the two rotations: B is staged when A's MMAs have completed, A's epilogue overlaps B's MMAs.
    struct TcState {
        uint32_t tmem_base, bw_addr;     // TMEM base address, shared-memory address of the weight operand
        uint64_t* bars;                  // [NWG][2]
        uint32_t phase;
        int team_bar = 0, team_threads = 0;   // > 0: synchronise the staging on this named barrier (whole scoring team) instead of
                                              // one barrier per warpgroup (callers that have no barrier ids to spare)
    };
    static DLPD_DEV void tc_setup(const Params& p, float* Bw, uint64_t* bars, uint32_t* tmem_slot) {
        DLPD_FOR_TASKS(i, TC_BW, NT) {
            const int e = i & 3, n = (i >> 2) % HP, kc = (i / (4 * HP)) % (C0 / 4), part = i / (C0 * HP);
            const int c = 4 * kc + e;
            const float v = n < H ? p.filt.W1t[c * H + n] : 0.0f;
            const float hi = tc::tf32_hi(v);
            Bw[i] = part == 0 ? hi : (v - hi);
        }
        if (threadIdx.x < 32) tc::tmem_alloc(tmem_slot, TC_COLS);
        if (threadIdx.x == 0) {
            for (int i = 0; i < 2 * NWG; ++i) tc::mbar_init(bars + i, 1);
            tc::mbar_init(bars + 14, 1);
            tc::mbar_init(bars + 15, 1);
            tc::mbar_fence_init();
        }
        tc::fence_proxy_async();           // the weights were written with ordinary stores, the tensor core reads them through the async proxy
        tc::fence_before_sync();
    }
    template <class TM>
    static DLPD_DEV void score_voxels_tc(const Params& p, const TM& tm, const cpx* b0, const cpx* h1, TcState& ts, int pair, int x0,
                                         int y0, unsigned long long& bestA, unsigned long long& bestB) {
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        const int t = tm.tid();
        const bool valid = t < NVOX;
        const int tv = valid ? t : 0;                         // padding rows of the last warpgroup recompute translation 0 and drop it
        const int xy = tv % XY, z = tv / XY;
        const int g = t >> 7, q = (t >> 5) & 3;
        const int gthreads = (NC - 128 * g) < 128 ? (NC - 128 * g) : 128;          // scoring threads of my warpgroup
        const uint32_t lane_addr = ts.tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t colA = (uint32_t)(g * TC_COLS_WG), colD0 = colA + 2 * C0, colD1 = colD0 + HP;
        constexpr uint32_t IDESC = tc::idesc_tf32(128, HP);
        constexpr uint32_t LBO = HP * 16, SBO = 128, KSTEP = 2 * HP * 16, BLO = C0 * HP * 4;   // bytes
        const cpx* v0 = b0 + z * NL0 + xy;
        const cpx* hz = h1 + (z >> 1) * H;
        cpx x[C0];
        DLPD_UNROLL
        for (int c = 0; c < C0; ++c) x[c] = v0[c * XY];
        uint64_t* bar = ts.bars + 2 * g;
#if defined(DLPD_ABLATE_EPI_MLP)
        if (valid) {                                          // timing experiment: no staging, no MMA, no read-back
            const unsigned int lin0 = (unsigned int)(((x0 + (xy & 1)) * N + (y0 + (xy >> 1))) * N + z);
            finish_voxel(p, fadd2_(x[0], hz[0]), v0[C0 * XY], lin0, rotA, rotB, hasB, bestA, bestB);
        }
        (void)bar; (void)lane_addr; (void)colA; (void)colD0; (void)colD1; (void)gthreads;
        return;
#endif
        DLPD_UNROLL
        for (int rot = 0; rot < 2; ++rot) {
            uint32_t hi[C0], lo[C0];
            DLPD_UNROLL
            for (int c = 0; c < C0; ++c) {
                const float a = rot ? x[c].y : x[c].x, ah = tc::tf32_hi(a);
                hi[c] = __float_as_uint(ah);
                lo[c] = __float_as_uint(a - ah);
            }
            if (rot == 1) {                                   // the A columns are free once rotation A's MMAs have completed
                tc::mbar_wait(bar, ts.phase);
                tc::fence_after_sync();
            }
            DLPD_UNROLL
            for (int h8 = 0; h8 < C0 / 8; ++h8) {
                uint32_t t0[8], t1[8];
                DLPD_UNROLL
                for (int i = 0; i < 8; ++i) { t0[i] = hi[8 * h8 + i]; t1[i] = lo[8 * h8 + i]; }
                tc::st8(lane_addr + colA + 8 * h8, t0);
                tc::st8(lane_addr + colA + C0 + 8 * h8, t1);
            }
            tc::wait_st();
            tc::fence_before_sync();
            if (ts.team_bar > 0) named_barrier_sync(ts.team_bar, ts.team_threads);
            else named_barrier_sync(2 + g, gthreads);
            if (q == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t d = ts.tmem_base + (rot ? colD1 : colD0), a = ts.tmem_base + colA;
                DLPD_UNROLL
                for (int s = 0; s < C0 / 8; ++s) {
                    const uint64_t dhi = tc::smem_desc_kmajor(ts.bw_addr + s * KSTEP, LBO, SBO);
                    const uint64_t dlo = tc::smem_desc_kmajor(ts.bw_addr + BLO + s * KSTEP, LBO, SBO);
                    tc::mma_tf32_ts(d, a + C0 + 8 * s, dhi, IDESC, s > 0);       // a_lo w_hi
                    tc::mma_tf32_ts(d, a + 8 * s, dlo, IDESC, 1);                // a_hi w_lo
                    tc::mma_tf32_ts(d, a + 8 * s, dhi, IDESC, 1);                // a_hi w_hi
                }
                tc::mma_commit(bar + rot);
            }
            __syncwarp();
        }
        float acc[2][HP];
        DLPD_UNROLL
        for (int rot = 0; rot < 2; ++rot) {
            if (rot == 1) {
                tc::mbar_wait(bar + 1, ts.phase);
                tc::fence_after_sync();
            }
            uint32_t r[HP / 8][8];
            DLPD_UNROLL
            for (int u = 0; u < HP / 8; ++u) tc::ld8(lane_addr + (rot ? colD1 : colD0) + 8 * u, r[u]);
            tc::wait_ld();
            DLPD_UNROLL
            for (int u = 0; u < HP / 8; ++u) {
                DLPD_UNROLL
                for (int i = 0; i < 8; ++i) acc[rot][8 * u + i] = __uint_as_float(r[u][i]);
            }
        }
        tc::fence_before_sync();                               // the accumulators are overwritten by the next tile's MMAs
        ts.phase ^= 1;
        float2 sc = dup2(p.filt.b2);
        DLPD_UNROLL
        for (int j = 0; j < H; ++j) {
            const float2 hh = fadd2_(make_float2(acc[0][j], acc[1][j]), DUO ? __ldg(hz + j) : hz[j]);
            sc = ffma2_(dup2(p.filt.W2[j]), make_float2(fmaxf(hh.x, 0.0f), fmaxf(hh.y, 0.0f)), sc);
        }
        if (valid) {
            const unsigned int lin = (unsigned int)(((x0 + (xy & 1)) * N + (y0 + (xy >> 1))) * N + z);
            finish_voxel(p, sc, v0[C0 * XY], lin, rotA, rotB, hasB, bestA, bestB);
        }
    }
#endif

    // grid: x = p.grid persistent CTAs
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* bufs = reinterpret_cast<cpx*>(smem);      // [2][kz][NL0]
        cpx* hbufs = bufs + 2 * BUF;                   // [2][z'][H]  (rotA, rotB): b1 + level-1 part of the hidden layer
        cpx* tw0 = hbufs + NHB * HBS;
        unsigned long long* red = reinterpret_cast<unsigned long long*>(tw0 + N + (PKT ? N : 0));
        MlpDev<C0, H>* ms = reinterpret_cast<MlpDev<C0, H>*>(red + 64);
        DLPD_FOR_TASKS(t, N, NT) tw0[t] = p.tw0[t];
        if constexpr (PKT) {
            DLPD_FOR_TASKS(t, N, NT) tw0[N + t] = packed_head_twiddle<N, LinePlan<N>::R0>(p.tw0, t);
        }
#if DLPD_EXPERIMENTAL
        if (MMA_OK && p.mlp != nullptr) {
            DLPD_FOR_TASKS(t, (int)sizeof(MlpDev<C0, H>) / 4, NT)
                reinterpret_cast<unsigned int*>(ms)[t] = reinterpret_cast<const unsigned int*>(p.mlp)[t];
        }
#endif
#if !defined(DLPD_EMU)
        TcState ts;
        ts.tmem_base = ts.bw_addr = 0u; ts.bars = nullptr; ts.phase = 0u;
        if constexpr (TC_OK) {
            float* Bw = reinterpret_cast<float*>(reinterpret_cast<char*>(red + 64) + MLP_BYTES);
            ts.bars = reinterpret_cast<uint64_t*>(Bw + TC_BW);
            uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ts.bars + 16);
            tc_setup(p, Bw, ts.bars, tmem_slot);
            __syncthreads();
            tc::fence_after_sync();
            ts.tmem_base = *tmem_slot;
            ts.bw_addr = tc::smem_u32(Bw);
        }
#endif
#if !defined(DLPD_EMU)
        if constexpr (USE_TMA) {
            // ---- TMA-fed loop: buffers are filled by the async proxy, "tile landed" is an mbarrier phase ----------------
            uint64_t* full = ts.bars + 14;
            unsigned long long bestA = ~0ull, bestB = ~0ull;
            int cur_pair = -1, slot = 0;
            uint32_t fullphase = 0;
            const SubTeam<NC + NX> tmT{(int)threadIdx.x, 1};         // transform team (everybody)
            const SubTeam<NC> tmS{(int)threadIdx.x, 1};             // scoring threads: one per translation
            if (threadIdx.x == 0) {
                tc::tma_prefetch_desc(p.tmap);
                if (bx < p.n_tiles) issue_tma(p, bx, bufs, hbufs, full);
            }
            for (int tile = bx; tile < p.n_tiles; tile += p.grid, slot ^= 1) {
                tc::fence_proxy_async();                     // my ordinary stores into the other buffer precede its async refill
                __syncthreads();                             // everyone is done with the other buffer
                const int pair = tile / TILES, r = tile % TILES;
                if (pair != cur_pair) {
                    if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
                    cur_pair = pair;
                    bestA = bestB = ~0ull;
                }
                if (threadIdx.x == 0 && tile + p.grid < p.n_tiles)
                    issue_tma(p, tile + p.grid, bufs + (slot ^ 1) * BUF, hbufs + (slot ^ 1) * HB, full + (slot ^ 1));
                tc::mbar_wait(full + slot, (fullphase >> slot) & 1u);
                fullphase ^= 1u << slot;
                cpx* b0 = bufs + slot * BUF;
                if constexpr (PKT) transform_tile_pk(tmT, b0, tw0, tw0 + N);
                else transform_tile(tmT, b0, tw0);
                if (threadIdx.x < NC)
                    score_voxels_tc(p, tmS, b0, hbufs + slot * HB, ts, pair, 2 * tile_xh(r), tile_y0(r), bestA, bestB);
            }
            if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
            tc::fence_before_sync();
            __syncthreads();
            if (threadIdx.x < 32) tc::tmem_dealloc(ts.tmem_base, TC_COLS);
            return;
        }
#endif
        if (bx < p.n_tiles) issue_loads(p, bx, bufs, hbufs);
        cp_async_commit();
        cp_async_wait_pending();
#if !defined(DLPD_EMU)
        uint32_t h1phase = 0;
        if constexpr (BULK_H1 && NHB == 2) {
            if (threadIdx.x == NC + NX && bx < p.n_tiles) issue_h1_bulk(p, bx, hbufs, ts.bars + 14);
        }
#endif
        unsigned long long bestA = ~0ull, bestB = ~0ull;
        int cur_pair = -1, slot = 0;
#if defined(DLPD_EMU)
        const CtaTeam<NC> tm;
        const bool producer = true, consumer = true;   // one emulated thread plays both roles, copies complete at once
#else
        const SubTeam<NC, DLPD_EPI_DUAL != 0> tm{(int)threadIdx.x, 1};
        const SubTeam<NC + NX, DLPD_EPI_DUAL != 0> tmT{(int)threadIdx.x, 1};     // transform team: the scoring threads + NX helpers
        const bool producer = threadIdx.x >= NC + NX, consumer = !producer;
#endif
        for (int tile = bx; tile < p.n_tiles; tile += p.grid, slot ^= 1) {
            DLPD_SYNC();                               // this tile has landed; everyone is done with the other buffer
            const int pair = tile / TILES, r = tile % TILES;
            if (pair != cur_pair) {
                if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
                cur_pair = pair;
                bestA = bestB = ~0ull;
            }
            if (producer) {
#if !defined(DLPD_ABLATE_EPI_LOAD)
#if !defined(DLPD_EMU)
                if constexpr (BULK_H1 && NHB == 1) {             // single column buffer: everyone is done with the previous tile's
                    if (threadIdx.x == NC + NX) issue_h1_bulk(p, tile, hbufs, ts.bars + 14);
                }
#endif
                if (tile + p.grid < p.n_tiles) {
                    issue_loads(p, tile + p.grid, bufs + (slot ^ 1) * BUF, hbufs + (NHB == 2 ? (slot ^ 1) : 0) * HBS);
#if !defined(DLPD_EMU)
                    if constexpr (BULK_H1 && NHB == 2) {
                        if (threadIdx.x == NC + NX) issue_h1_bulk(p, tile + p.grid, hbufs + (slot ^ 1) * HBS, ts.bars + 14 + (slot ^ 1));
                    }
#endif
                }
#endif
                cp_async_commit();
                cp_async_wait_pending();
            }
            if (consumer) {
#if !defined(DLPD_EMU)
                if constexpr (TC_OK) {
                    cpx* b0 = bufs + slot * BUF;
                    if constexpr (PKT) transform_tile_pk(tmT, b0, tw0, tw0 + N);
                    else transform_tile(tmT, b0, tw0);
                    if (threadIdx.x < NC) {
                        constexpr int HS = (NHB == 2) ? 1 : 0;    // column buffer of this tile: its slot, or the single one
                        if constexpr (BULK_H1) {                 // the tile's level-1 column has landed (mbarrier phase of its buffer)
                            tc::mbar_wait(ts.bars + 14 + HS * slot, (h1phase >> (HS * slot)) & 1u);
                            h1phase ^= 1u << (HS * slot);
                        }
                        const cpx* hcol = DUO ? p.h1 + ((long)pair * T * T + tile_xh(r) * T + (tile_y0(r) >> 1)) * HB
                                              : hbufs + HS * slot * HBS;
                        score_voxels_tc(p, tm, b0, hcol, ts, pair, 2 * tile_xh(r), tile_y0(r), bestA, bestB);
                    }
                } else
#endif
                score_tile(p, tm, bufs + slot * BUF, hbufs + (NHB == 2 ? slot : 0) * HBS, tw0, ms, pair, 2 * tile_xh(r), tile_y0(r), bestA, bestB);
            }
        }
        if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
#if !defined(DLPD_EMU)
        if constexpr (TC_OK) {
            tc::fence_before_sync();
            __syncthreads();
            if (threadIdx.x < 32) tc::tmem_dealloc(ts.tmem_base, TC_COLS);
        }
#endif
    }
};

#if DLPD_EXPERIMENTAL
#include "dlpd_experimental.h"
#endif

// packed best -> [n_rot][4] floats (score, x, y, z)
struct DecodeParams {
    const unsigned long long* best;
    float* out;
    int n_rot, N;
};
struct DecodeBest {
    typedef DecodeParams Params;
    static constexpr int NT = 256;
    static constexpr int SMEM = 0;
    static DLPD_DEV void body(const Params& p, int bx, int, int, char*) {
        DLPD_FOR_TASKS(t, NT, NT) {
            const int i = bx * NT + t;
            if (i < p.n_rot) {
                const unsigned long long k = p.best[i];
                const unsigned int lin = (unsigned int)(k & 0xffffffffu);
                p.out[4 * i + 0] = key_score((unsigned int)(k >> 32));
                p.out[4 * i + 1] = (float)(lin / (unsigned int)(p.N * p.N));
                p.out[4 * i + 2] = (float)((lin / (unsigned int)p.N) % (unsigned int)p.N);
                p.out[4 * i + 3] = (float)(lin % (unsigned int)p.N);
            }
        }
    }
};
