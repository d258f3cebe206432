// Measured-and-rejected variants of the epilogue: compiled only with -DDLPD_EXPERIMENTAL=1 (the CPU emulation
// build of tests/emu sets it so that the bodies stay covered); the shipped libdlpd.so does not contain them.
#pragma once

// ------------------------------------------------------------------------------------------------------
// EpiloguePipe: the same per-tile work as Epilogue, organised as a three-stage pipeline inside one persistent CTA so
// that the inverse z transforms of one tile (shared-memory / latency bound, ~1/3 of the FP32 pipe) run WHILE the filter
// of the previous tile (pure FFMA2, pipe bound) executes:
//     producer warps : cp.async of half-tile it + 2.. into a ring of NB buffers            (waits FREE, signals FULL)
//     FFT team       : inverse z transform of half-tile it + 1, in place                   (waits FULL, signals XF)
//     filter team    : 48 -> 24 -> 1 filter, clash mask, argmin / candidates of half-tile it (waits XF,  signals FREE)
// A half-tile is 2 x 1 (x,y) columns (one y of a 2x2 group): four half-tile buffers fit where two 2x2 buffers did.  The
// two halves of a group are consecutive iterations of the same CTA and share the group's level-1 column h1, which
// travels with the first half in a ring of NHB = 3 buffers.  The stages hand buffers over with named barriers
// (bar.arrive by the finishing stage, bar.sync by the waiting one): FULL/XF/FREE x NB buffers + one barrier inside each team.
// ------------------------------------------------------------------------------------------------------
#ifndef DLPD_PIPE_NFFT
#define DLPD_PIPE_NFFT 352
#endif
#ifndef DLPD_PIPE_NMLP
#define DLPD_PIPE_NMLP 320
#endif
#ifndef DLPD_PIPE_NPROD
#define DLPD_PIPE_NPROD 128
#endif

template <int L_, int C0, int C1, int H> struct EpiloguePipe {
    typedef Epilogue<L_, C0, C1, H, 1> E;            // half-tile geometry, loads and scoring code
    typedef EpiParams<C0, C1, H> Params;
    static constexpr int L = L_, N = 2 * L_, N1 = L_, T = L_;
    static constexpr int NL0 = E::NL0, BUF = N * NL0, HB = N1 * H;
    static constexpr int NB = 4, NHB = 3;
    static constexpr bool BIG = N >= 128;
    static constexpr int NPROD = BIG ? DLPD_PIPE_NPROD : 64, PW = NPROD / 2;      // two producer groups: even / odd iterations
    static constexpr int NFFT = BIG ? DLPD_PIPE_NFFT : 64, NMLP = BIG ? DLPD_PIPE_NMLP : 64;
    static constexpr int NT = NPROD + NFFT + NMLP;
    static constexpr int SMEM = (NB * BUF + NHB * HB + N) * (int)sizeof(cpx) + 64 * 8;
    static constexpr int BAR_FFT = 1, BAR_MLP = 2, BAR_FULL = 3, BAR_XF = 3 + NB, BAR_FREE = 3 + 2 * NB;
    static_assert(3 + 3 * NB <= 16 && PW % 32 == 0 && NFFT % 32 == 0 && NMLP % 32 == 0 && HB % 2 == 0, "pipeline shape");

    struct Item { int pair, r4, half; };
    // iteration `it` of CTA bx: half `it & 1` of 2x2 group bx + (it >> 1) * grid
    static DLPD_DEV Item item(const Params& p, int bx, int it) {
        const int g = bx + (it >> 1) * p.grid;
        return Item{g / (T * T), g % (T * T), it & 1};
    }

    // copies of one half-tile (and, with its first half, the group's h1 column) by PW threads; tp = index among them
    static DLPD_DEV void issue_loads(const Params& p, const Item& w, cpx* b0, cpx* hb, int tp) {
        const cpx* s = p.s2_0 + ((long)w.pair * (C0 + 1) * N * T * T + w.r4) * 4 + 2 * w.half;
        int c = tp / N, kz = tp % N;
        const cpx* src = s + (long)tp * (T * T) * 4;                     // q = c*N + kz: 16 contiguous bytes per row
        for (int q = tp; q < (C0 + 1) * N; q += PW) {
            cp_async16(b0 + kz * NL0 + c * 2, src);
            src += (long)PW * T * T * 4;
            kz += PW % N;
            c += PW / N;
            if (kz >= N) { kz -= N; ++c; }
        }
        if (w.half == 0) {
            const cpx* hs = p.h1 + ((long)w.pair * T * T + w.r4) * HB;
            for (int t = tp; t < HB / 2; t += PW) cp_async16(hb + 2 * t, hs + 2 * t);
        }
    }

#if !defined(DLPD_EMU)
    static DLPD_DEV unsigned long long team_min_u64(unsigned long long v, unsigned long long* red, int tmid) {
        DLPD_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = w < v ? w : v;
        }
        const int lane = tmid & 31, warp = tmid >> 5;
        if (lane == 0) red[warp] = v;
        named_barrier_sync(BAR_MLP, NMLP);
        if (warp == 0) {
            v = lane < NMLP / 32 ? red[lane] : ~0ull;
            DLPD_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
                v = w < v ? w : v;
            }
        }
        named_barrier_sync(BAR_MLP, NMLP);
        return v;      // valid in team thread 0
    }
#endif

    static DLPD_DEV void flush_best(const Params& p, int pair, unsigned long long bestA, unsigned long long bestB,
                                    unsigned long long* red, int tmid) {
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
#if defined(DLPD_EMU)
        (void)red; (void)tmid;
        if (bestA < p.best[rotA]) p.best[rotA] = bestA;
        if (hasB && bestB < p.best[rotB]) p.best[rotB] = bestB;
#else
        bestA = team_min_u64(bestA, red, tmid);
        bestB = team_min_u64(bestB, red, tmid);
        if (tmid == 0) {
            atomicMin(p.best + rotA, bestA);
            if (hasB) atomicMin(p.best + rotB, bestB);
        }
#endif
    }

    // grid: x = p.grid persistent CTAs; p.n_tiles = number of 2x2 groups = P * T * T
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* bufs = reinterpret_cast<cpx*>(smem);      // [NB][kz][NL0]
        cpx* hbufs = bufs + NB * BUF;                  // [NHB][z'][H]
        cpx* tw0 = hbufs + NHB * HB;
        unsigned long long* red = reinterpret_cast<unsigned long long*>(tw0 + N);
        const int my_groups = bx < p.n_tiles ? (p.n_tiles - bx + p.grid - 1) / p.grid : 0;
        const int n_it = 2 * my_groups;
        DLPD_FOR_TASKS(t, N, NT) tw0[t] = p.tw0[t];
        DLPD_SYNC();
#if defined(DLPD_EMU)
        // one emulated thread plays the three roles in turn, copies complete at once
        unsigned long long bestA = ~0ull, bestB = ~0ull;
        int cur_pair = -1;
        const CtaTeam<1> tm;
        for (int it = 0; it < n_it; ++it) {
            const Item w = item(p, bx, it);
            cpx* b0 = bufs + (it % NB) * BUF;
            cpx* hb = hbufs + ((it >> 1) % NHB) * HB;
            for (int tp = 0; tp < PW; ++tp) issue_loads(p, w, b0, hb, tp);
            E::transform_tile(tm, b0, tw0);
            if (w.pair != cur_pair) {
                if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, 0);
                cur_pair = w.pair;
                bestA = bestB = ~0ull;
            }
            E::score_voxels(p, tm, b0, hb, w.pair, 2 * (w.r4 / T), 2 * (w.r4 % T) + w.half, bestA, bestB);
        }
        if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, 0);
#else
        const int tid = (int)threadIdx.x;
        if (tid < NPROD) {
            // ---- producers: group `par` serves the iterations of its parity (buffer index has the same parity)
            const int par = tid / PW, tp = tid % PW;
            for (int it = par; it < n_it; it += 2) {
                const int b = it % NB;
                if (it >= NB) named_barrier_sync(BAR_FREE + b, PW + NMLP);          // filter team is done with buffer b
#if !defined(DLPD_ABLATE_EPI_LOAD)
                issue_loads(p, item(p, bx, it), bufs + b * BUF, hbufs + ((it >> 1) % NHB) * HB, tp);
#endif
                cp_async_commit();
                cp_async_wait_pending();
                named_barrier_arrive(BAR_FULL + b, PW + NFFT);
            }
        } else if (tid < NPROD + NFFT) {
            // ---- FFT team
            const SubTeam<NFFT> tm{tid - NPROD, BAR_FFT};
            for (int it = 0; it < n_it; ++it) {
                const int b = it % NB;
                named_barrier_sync(BAR_FULL + b, PW + NFFT);
                E::transform_tile(tm, bufs + b * BUF, tw0);
                named_barrier_arrive(BAR_XF + b, NFFT + NMLP);
            }
        } else {
            // ---- filter team
            const int tmid = tid - NPROD - NFFT;
            const SubTeam<NMLP> tm{tmid, BAR_MLP};
            unsigned long long bestA = ~0ull, bestB = ~0ull;
            int cur_pair = -1;
            for (int it = 0; it < n_it; ++it) {
                const int b = it % NB;
                const Item w = item(p, bx, it);
                if (w.pair != cur_pair) {
                    if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, tmid);
                    cur_pair = w.pair;
                    bestA = bestB = ~0ull;
                }
                named_barrier_sync(BAR_XF + b, NFFT + NMLP);
                E::score_voxels(p, tm, bufs + b * BUF, hbufs + ((it >> 1) % NHB) * HB, w.pair, 2 * (w.r4 / T), 2 * (w.r4 % T) + w.half,
                                bestA, bestB);
                if (it + NB < n_it) named_barrier_arrive(BAR_FREE + b, PW + NMLP);
            }
            if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, tmid);
        }
#endif
    }
};

// ------------------------------------------------------------------------------------------------------
// EpiloguePipeTC (round 2): the three-stage pipeline of EpiloguePipe with the tcgen05 / TMEM filter in its last stage.  The
// round-1 pipeline lost because its filter stage was FP32-pipe bound like the transform stage; the tcgen05 scoring is TMEM /
// latency bound and leaves the FP32 pipe to the transform team.  Role order: scoring team first (its warps must start at a
// multiple of four: a warp reaches only the TMEM lanes 32 (warp % 4) ..), then the FFT team, then the producers.
// ------------------------------------------------------------------------------------------------------
#if !defined(DLPD_EMU)
template <int L_, int C0, int C1, int H> struct EpiloguePipeTC {
    typedef Epilogue<L_, C0, C1, H, 1, 2> E;         // half-tile geometry, tcgen05 scoring code
    typedef EpiloguePipe<L_, C0, C1, H> P0;          // loads, team reductions
    typedef EpiParams<C0, C1, H> Params;
    static constexpr int L = L_, N = 2 * L_, N1 = L_, T = L_;
    static constexpr int NL0 = E::NL0, BUF = N * NL0, HB = N1 * H;
    static constexpr int NB = 4, NHB = 3;
    static constexpr int NMLP = E::NC, NFFT = DLPD_PIPE_NFFT, NPROD = DLPD_PIPE_NPROD, PW = NPROD / 2;
    static constexpr int NT = NMLP + NFFT + NPROD;
    static constexpr bool OK = E::TC_OK && (N >= 128) && (NMLP == P0::NMLP) && (NFFT == P0::NFFT) && (NPROD == P0::NPROD);
    static constexpr int SMEM = (NB * BUF + NHB * HB + N) * (int)sizeof(cpx) + 64 * 8 + E::TC_BYTES;
    static constexpr int BAR_FFT = P0::BAR_FFT, BAR_MLP = P0::BAR_MLP, BAR_FULL = P0::BAR_FULL, BAR_XF = P0::BAR_XF, BAR_FREE = P0::BAR_FREE;

    // grid: x = p.grid persistent CTAs; p.n_tiles = number of 2x2 groups = P * T * T
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* bufs = reinterpret_cast<cpx*>(smem);      // [NB][kz][NL0]
        cpx* hbufs = bufs + NB * BUF;                  // [NHB][z'][H]
        cpx* tw0 = hbufs + NHB * HB;
        unsigned long long* red = reinterpret_cast<unsigned long long*>(tw0 + N);
        float* Bw = reinterpret_cast<float*>(red + 64);
        const int my_groups = bx < p.n_tiles ? (p.n_tiles - bx + p.grid - 1) / p.grid : 0;
        const int n_it = 2 * my_groups;
        DLPD_FOR_TASKS(t, N, NT) tw0[t] = p.tw0[t];
        typename E::TcState ts;
        ts.tmem_base = 0u; ts.phase = 0u;
        ts.bars = reinterpret_cast<uint64_t*>(Bw + E::TC_BW);
        ts.team_bar = BAR_MLP; ts.team_threads = NMLP;
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ts.bars + 16);
        E::tc_setup(p, Bw, ts.bars, tmem_slot);
        __syncthreads();
        tc::fence_after_sync();
        ts.tmem_base = *tmem_slot;
        ts.bw_addr = tc::smem_u32(Bw);
        const int tid = (int)threadIdx.x;
        if (tid >= NMLP + NFFT) {
            // ---- producers: group `par` serves the iterations of its parity (buffer index has the same parity)
            const int par = (tid - NMLP - NFFT) / PW, tp = (tid - NMLP - NFFT) % PW;
            for (int it = par; it < n_it; it += 2) {
                const int b = it % NB;
                if (it >= NB) named_barrier_sync(BAR_FREE + b, PW + NMLP);          // scoring team is done with buffer b
                P0::issue_loads(p, P0::item(p, bx, it), bufs + b * BUF, hbufs + ((it >> 1) % NHB) * HB, tp);
                cp_async_commit();
                cp_async_wait_pending();
                named_barrier_arrive(BAR_FULL + b, PW + NFFT);
            }
        } else if (tid >= NMLP) {
            // ---- FFT team
            const SubTeam<NFFT> tm{tid - NMLP, BAR_FFT};
            for (int it = 0; it < n_it; ++it) {
                const int b = it % NB;
                named_barrier_sync(BAR_FULL + b, PW + NFFT);
                E::transform_tile(tm, bufs + b * BUF, tw0);
                named_barrier_arrive(BAR_XF + b, NFFT + NMLP);
            }
        } else {
            // ---- scoring team (tcgen05): thread = translation of the half-tile, warpgroup = 128 TMEM lanes
            const SubTeam<NMLP> tm{tid, BAR_MLP};
            unsigned long long bestA = ~0ull, bestB = ~0ull;
            int cur_pair = -1;
            for (int it = 0; it < n_it; ++it) {
                const int b = it % NB;
                const typename P0::Item w = P0::item(p, bx, it);
                if (w.pair != cur_pair) {
                    if (cur_pair >= 0) P0::flush_best(p, cur_pair, bestA, bestB, red, tid);
                    cur_pair = w.pair;
                    bestA = bestB = ~0ull;
                }
                named_barrier_sync(BAR_XF + b, NFFT + NMLP);
                E::score_voxels_tc(p, tm, bufs + b * BUF, hbufs + ((it >> 1) % NHB) * HB, ts, w.pair, 2 * (w.r4 / T), 2 * (w.r4 % T) + w.half,
                                   bestA, bestB);
                if (it + NB < n_it) named_barrier_arrive(BAR_FREE + b, PW + NMLP);
            }
            if (cur_pair >= 0) P0::flush_best(p, cur_pair, bestA, bestB, red, tid);
        }
        tc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tc::tmem_dealloc(ts.tmem_base, E::TC_COLS);
    }
};
#endif

// ------------------------------------------------------------------------------------------------------
// EpilogueFused: the half-tile ring and producer warps of EpiloguePipe, but ONE consumer team whose every thread does, in
// the same straight-line code, a butterfly task of half-tile it (shared-memory latency bound) AND half of the filter of
// its voxel of half-tile it - 1 (a pure FFMA2 stream on 2 H accumulators): the instruction scheduler fills the load
// latencies of the transform with filter arithmetic inside each warp, so no warp-level hand-over, no unbalanced teams and no
// starvation of one role by the other (the failure modes of EpiloguePipe / EpilogueDuo).
//     phase 1:  inverse-tail task of half-tile it   +  filter channels [0, C0/2) of the voxel of half-tile it - 1
//     phase 2:  inverse-head task(s) of half-tile it +  channels [C0/2, C0), ReLU, second layer, mask, argmin
// ------------------------------------------------------------------------------------------------------
#ifndef DLPD_FUSED_FILTER_FIRST
#define DLPD_FUSED_FILTER_FIRST 1
#endif
#ifndef DLPD_FUSED_NC
#define DLPD_FUSED_NC 352
#endif
#ifndef DLPD_FUSED_NPROD
#define DLPD_FUSED_NPROD DLPD_PIPE_NPROD
#endif
template <int L_, int C0, int C1, int H> struct EpilogueFused {
    typedef Epilogue<L_, C0, C1, H, 1> E;            // half-tile geometry and scoring code
    typedef EpiloguePipe<L_, C0, C1, H> PP;          // items, loads
    typedef EpiParams<C0, C1, H> Params;
    typedef typename PP::Item Item;
    static constexpr int L = L_, N = 2 * L_, N1 = L_, T = L_;
    static constexpr int NL0 = E::NL0, XY = E::XY, BUF = N * NL0, HB = N1 * H;
    static constexpr int NB = 4, NHB = 3;
    static constexpr bool BIG = N >= 128;
    static constexpr int NPROD = BIG ? DLPD_FUSED_NPROD : 64, PW = NPROD / 2;
    static constexpr int NC = BIG ? DLPD_FUSED_NC : 64;
    static constexpr int NT = NPROD + NC;
    static constexpr int SMEM = (NB * BUF + NHB * HB + N) * (int)sizeof(cpx) + 64 * 8;
    static constexpr int BAR_TEAM = 1, BAR_FULL = 2, BAR_FREE = 2 + NB;
    typedef LinePlan<N> LP;
    static constexpr int NTAIL = (LP::TAIL > 1) ? NL0 * (N / LP::TAIL) : 0;      // tasks of the inverse tail / head pass
    static constexpr int NHEAD = NL0 * (N / (2 * LP::R0));
    static constexpr int NVOX = XY * N, CA = C0 / 2;
    static_assert(2 + 2 * NB <= 16 && PW % 32 == 0 && NC % 32 == 0 && HB % 2 == 0 && C0 >= 2, "pipeline shape");
    static_assert(PW == PP::PW, "the producers run EpiloguePipe::issue_loads, which strides by its own group width");

    static DLPD_DEV void tail_task(cpx* DLPD_RESTRICT b, const cpx* DLPD_RESTRICT tw, int t) {
        if constexpr (LP::TAIL > 1) fft_one_task<PASS_DIT, LP::TA, LP::TB, NL0, N, LP::TAIL, N, NL0>(b, tw, LineMul<1>(), t);
    }
    static DLPD_DEV void head_task(cpx* DLPD_RESTRICT b, const cpx* DLPD_RESTRICT tw, int t) {
        fft_one_task<PASS_DIT, 2, LP::R0, NL0, N, N, N, NL0>(b, tw, LineMul<1>(), t);
    }
    // filter of voxel v = z * XY + xy, split in two: same operation order as Epilogue::score_voxels (bit-identical scores)
    static DLPD_DEV void filter_a(const Params& p, const cpx* DLPD_RESTRICT b0, const cpx* DLPD_RESTRICT h1, int v, float2 (&h)[H]) {
        const int xy = v % XY, z = v / XY;
        const cpx* v0 = b0 + z * NL0 + xy;
        const cpx* hz = h1 + (z >> 1) * H;
        DLPD_UNROLL
        for (int j = 0; j < H; ++j) h[j] = hz[j];
        DLPD_UNROLL
        for (int c = 0; c < CA; ++c) {
            const float2 x = v0[c * XY];
            DLPD_UNROLL
            for (int j = 0; j < H; ++j) h[j] = ffma2_(dup2(p.filt.W1t[c * H + j]), x, h[j]);
        }
    }
    static DLPD_DEV void filter_b(const Params& p, const cpx* DLPD_RESTRICT b0, int v, float2 (&h)[H]) {
        const int xy = v % XY, z = v / XY;
        const cpx* v0 = b0 + z * NL0 + xy;
        DLPD_UNROLL
        for (int c = CA; c < C0; ++c) {
            const float2 x = v0[c * XY];
            DLPD_UNROLL
            for (int j = 0; j < H; ++j) h[j] = ffma2_(dup2(p.filt.W1t[c * H + j]), x, h[j]);
        }
    }
    static DLPD_DEV void filter_c(const Params& p, const cpx* b0, int v, const float2 (&h)[H], int pair, int x0, int y0,
                                  unsigned long long& bestA, unsigned long long& bestB) {
        const int xy = v % XY, z = v / XY;
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
        float2 sc = dup2(p.filt.b2);
        DLPD_UNROLL
        for (int j = 0; j < H; ++j) sc = ffma2_(dup2(p.filt.W2[j]), make_float2(fmaxf(h[j].x, 0.0f), fmaxf(h[j].y, 0.0f)), sc);
        const cpx clash = b0[z * NL0 + C0 * XY + xy];
        const float Va = clash.x < p.tau ? sc.x : 0.0f;     // (clash < tau).float() * score   (Docker.py:226,232)
        const float Vb = clash.y < p.tau ? sc.y : 0.0f;
        const unsigned int lin = (unsigned int)(((x0 + (xy & 1)) * N + (y0 + (xy >> 1))) * N + z);
        const unsigned long long ka = ((unsigned long long)score_key(Va) << 32) | lin;
        const unsigned long long kb = ((unsigned long long)score_key(Vb) << 32) | lin;
        bestA = ka < bestA ? ka : bestA;
        bestB = kb < bestB ? kb : bestB;
        if (p.score_out != nullptr) {
            p.score_out[(long)rotA * N * N * N + lin] = Va;
            if (hasB) p.score_out[(long)rotB * N * N * N + lin] = Vb;
        }
        if (p.cand != nullptr) {
            if ((unsigned int)(ka >> 32) <= p.cand_key) E::emit_candidate(p, Va, rotA, lin);
            if (hasB && (unsigned int)(kb >> 32) <= p.cand_key) E::emit_candidate(p, Vb, rotB, lin);
        }
    }

#if !defined(DLPD_EMU)
    static DLPD_DEV unsigned long long team_min_u64(unsigned long long v, unsigned long long* red, int t) {
        DLPD_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = w < v ? w : v;
        }
        const int lane = t & 31, warp = t >> 5;
        if (lane == 0) red[warp] = v;
        named_barrier_sync(BAR_TEAM, NC);
        if (warp == 0) {
            v = lane < NC / 32 ? red[lane] : ~0ull;
            DLPD_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
                v = w < v ? w : v;
            }
        }
        named_barrier_sync(BAR_TEAM, NC);
        return v;      // valid in team thread 0
    }
#endif
    static DLPD_DEV void flush_best(const Params& p, int pair, unsigned long long bestA, unsigned long long bestB,
                                    unsigned long long* red, int t) {
        const int rotA = 2 * pair, rotB = 2 * pair + 1;
        const bool hasB = rotB < p.n_rot;
#if defined(DLPD_EMU)
        (void)red; (void)t;
        if (bestA < p.best[rotA]) p.best[rotA] = bestA;
        if (hasB && bestB < p.best[rotB]) p.best[rotB] = bestB;
#else
        bestA = team_min_u64(bestA, red, t);
        bestB = team_min_u64(bestB, red, t);
        if (t == 0) {
            atomicMin(p.best + rotA, bestA);
            if (hasB) atomicMin(p.best + rotB, bestB);
        }
#endif
    }

    // grid: x = p.grid persistent CTAs; p.n_tiles = number of 2x2 groups = P * T * T
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* bufs = reinterpret_cast<cpx*>(smem);      // [NB][kz][NL0]
        cpx* hbufs = bufs + NB * BUF;                  // [NHB][z'][H]
        cpx* tw0 = hbufs + NHB * HB;
        unsigned long long* red = reinterpret_cast<unsigned long long*>(tw0 + N);
        const int my_groups = bx < p.n_tiles ? (p.n_tiles - bx + p.grid - 1) / p.grid : 0;
        const int n_it = 2 * my_groups;
        DLPD_FOR_TASKS(t, N, NT) tw0[t] = p.tw0[t];
        DLPD_SYNC();
#if defined(DLPD_EMU)
        // one emulated thread plays every role in turn, copies complete at once
        unsigned long long bestA = ~0ull, bestB = ~0ull;
        int cur_pair = -1;
        for (int it = 0; it < n_it; ++it) {
            const Item w = PP::item(p, bx, it);
            cpx* b0 = bufs + (it % NB) * BUF;
            cpx* hb = hbufs + ((it >> 1) % NHB) * HB;
            for (int tp = 0; tp < PW; ++tp) PP::issue_loads(p, w, b0, hb, tp);
            for (int t = 0; t < NTAIL; ++t) tail_task(b0, tw0, t);
            for (int t = 0; t < NHEAD; ++t) head_task(b0, tw0, t);
            if (w.pair != cur_pair) {
                if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, 0);
                cur_pair = w.pair;
                bestA = bestB = ~0ull;
            }
            for (int v = 0; v < NVOX; ++v) {
                float2 h[H];
                filter_a(p, b0, hb, v, h);
                filter_b(p, b0, v, h);
                filter_c(p, b0, v, h, w.pair, 2 * (w.r4 / T), 2 * (w.r4 % T) + w.half, bestA, bestB);
            }
        }
        if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, 0);
#else
        const int tid = (int)threadIdx.x;
        if (tid < NPROD) {
            // ---- producers: group `par` serves the iterations of its parity (buffer index has the same parity)
            const int par = tid / PW, tp = tid % PW;
            for (int it = par; it < n_it; it += 2) {
                const int b = it % NB;
                if (it >= NB) named_barrier_sync(BAR_FREE + b, PW + NC);            // the consumers are done with buffer b
                PP::issue_loads(p, PP::item(p, bx, it), bufs + b * BUF, hbufs + ((it >> 1) % NHB) * HB, tp);
                cp_async_commit();
                cp_async_wait_pending();
                named_barrier_arrive(BAR_FULL + b, PW + NC);
            }
        } else {
            // ---- consumers: iteration it transforms half-tile it and filters half-tile it - 1
            const int t = tid - NPROD;
            unsigned long long bestA = ~0ull, bestB = ~0ull;
            int cur_pair = -1;
            for (int it = 0; it <= n_it; ++it) {
                const bool hasT = it < n_it, hasF = it >= 1;
                cpx* bT = bufs + (it % NB) * BUF;
                const int itf = hasF ? it - 1 : 0;
                const cpx* bF = bufs + (itf % NB) * BUF;
                const cpx* hF = hbufs + ((itf >> 1) % NHB) * HB;
                const Item wf = PP::item(p, bx, itf);
                const int x0 = 2 * (wf.r4 / T), y0 = 2 * (wf.r4 % T) + wf.half;
                if (hasF && wf.pair != cur_pair) {
                    if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, t);
                    cur_pair = wf.pair;
                    bestA = bestB = ~0ull;
                }
                if (hasT) named_barrier_sync(BAR_FULL + it % NB, PW + NC);          // half-tile it has landed
                const bool vox = hasF && t < NVOX;
                float2 h[H];
                // phase 1: one inverse-tail task and the first half of the filter in the same block
                if (vox) {
                    if (hasT && t < NTAIL) {
#if DLPD_FUSED_FILTER_FIRST
                        filter_a(p, bF, hF, t, h);
                        tail_task(bT, tw0, t);
#else
                        tail_task(bT, tw0, t);
                        filter_a(p, bF, hF, t, h);
#endif
                    } else {
                        filter_a(p, bF, hF, t, h);
                    }
                } else if (hasT && t < NTAIL) {
                    tail_task(bT, tw0, t);
                }
                if (hasT) {
                    for (int t2 = t + NC; t2 < NTAIL; t2 += NC) tail_task(bT, tw0, t2);
                    named_barrier_sync(BAR_TEAM, NC);
                }
                // phase 2: inverse-head task and the second half of the filter
                if (vox) {
                    if (hasT && t < NHEAD) {
#if DLPD_FUSED_FILTER_FIRST
                        filter_b(p, bF, t, h);
                        head_task(bT, tw0, t);
#else
                        head_task(bT, tw0, t);
                        filter_b(p, bF, t, h);
#endif
                    } else {
                        filter_b(p, bF, t, h);
                    }
                    filter_c(p, bF, t, h, wf.pair, x0, y0, bestA, bestB);
                } else if (hasT && t < NHEAD) {
                    head_task(bT, tw0, t);
                }
                if (hasT) {
                    for (int t2 = t + NC; t2 < NHEAD; t2 += NC) head_task(bT, tw0, t2);
                }
                // voxels beyond one per thread (small boxes only)
                if (hasF) {
                    for (int v = t + NC; v < NVOX; v += NC) {
                        float2 g[H];
                        filter_a(p, bF, hF, v, g);
                        filter_b(p, bF, v, g);
                        filter_c(p, bF, v, g, wf.pair, x0, y0, bestA, bestB);
                    }
                    if (itf + NB < n_it) named_barrier_arrive(BAR_FREE + itf % NB, PW + NC);
                }
                if (hasT) named_barrier_sync(BAR_TEAM, NC);                         // half-tile it is transformed
            }
            if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red, t);
        }
#endif
    }
};

// ------------------------------------------------------------------------------------------------------
// EpilogueDuo: the epilogue as TWO independent persistent CTAs per SM (384 threads, ~104 KB each) instead of one CTA of
// 768 threads: while one CTA is in its shared-memory-bound transform passes the other is in its FFMA2-bound filter loop,
// with no hand-over protocol at all.  Half-tiles (2x1 columns: one y of a 2x2 group) halve the tile buffers; the two
// halves of a group are consecutive iterations of the same CTA and share ONE h1 buffer, which is fetched at the start of
// the first half and only needed after that half's transform passes.  Every thread issues its share of the cp.async
// copies (next half-tile + h1), so there are no producer warps.
// ------------------------------------------------------------------------------------------------------
#ifndef DLPD_DUO_SKEW
#define DLPD_DUO_SKEW 0
#endif
#ifndef DLPD_DUO_NT
#define DLPD_DUO_NT 384
#endif
template <int L_, int C0, int C1, int H> struct EpilogueDuo {
    typedef Epilogue<L_, C0, C1, H, 1> E;            // half-tile geometry and scoring code
    typedef EpiParams<C0, C1, H> Params;
    static constexpr int L = L_, N = 2 * L_, N1 = L_, T = L_;
    static constexpr int NL0 = E::NL0, BUF = N * NL0, HB = N1 * H;
    static constexpr int NT = (N >= 128) ? DLPD_DUO_NT : 128;
    static constexpr int MINB = 2;
    static constexpr int SMEM = (2 * BUF + HB + N) * (int)sizeof(cpx) + 64 * 8;
    static_assert(HB % 2 == 0 && NT % 32 == 0, "shapes");

    struct Item { int pair, r4, half; };
    static DLPD_DEV Item item(const Params& p, int bx, int it) {
        const int g = bx + (it >> 1) * p.grid;
        return Item{g / (T * T), g % (T * T), it & 1};
    }
    static DLPD_DEV void issue_data(const Params& p, const Item& w, cpx* b0) {
        const cpx* s = p.s2_0 + ((long)w.pair * (C0 + 1) * N * T * T + w.r4) * 4 + 2 * w.half;
        DLPD_FOR_TASKS(q, (C0 + 1) * N, NT) {                                // q = c*N + kz: 16 contiguous bytes per row
            const int c = q / N, kz = q - c * N;
            cp_async16(b0 + kz * NL0 + c * 2, s + (long)q * (T * T) * 4);
        }
    }
    static DLPD_DEV void issue_h1(const Params& p, const Item& w, cpx* hb) {
        const cpx* hs = p.h1 + ((long)w.pair * T * T + w.r4) * HB;
        DLPD_FOR_TASKS(t, HB / 2, NT) cp_async16(hb + 2 * t, hs + 2 * t);
    }
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

    // grid: x = p.grid persistent CTAs (two per SM); p.n_tiles = number of 2x2 groups = P * T * T
    static DLPD_DEV void body(const Params& p, int bx, int, int, char* smem) {
        cpx* bufs = reinterpret_cast<cpx*>(smem);      // [2][kz][NL0]
        cpx* hbuf = bufs + 2 * BUF;                    // [z'][H]: level-1 hidden column of the current group
        cpx* tw0 = hbuf + HB;
        unsigned long long* red = reinterpret_cast<unsigned long long*>(tw0 + N);
        const int my_groups = bx < p.n_tiles ? (p.n_tiles - bx + p.grid - 1) / p.grid : 0;
        const int n_it = 2 * my_groups;
        DLPD_FOR_TASKS(t, N, NT) tw0[t] = p.tw0[t];
        if (n_it > 0) issue_data(p, item(p, bx, 0), bufs);
        cp_async_commit();
        cp_async_wait_pending();
        const CtaTeam<NT> tm;
        unsigned long long bestA = ~0ull, bestB = ~0ull;
        int cur_pair = -1;
#if !defined(DLPD_EMU) && DLPD_DUO_SKEW > 0
        // the two CTAs of an SM start together and would stay in phase (both transforming, then both filtering): delay the
        // second one by about half a half-tile period so that one is in its FFMA-bound filter while the other transforms
        if (bx >= p.grid / 2) {
            const long long t0 = clock64();
            while (clock64() - t0 < DLPD_DUO_SKEW) {}
        }
#endif
        for (int it = 0, slot = 0; it < n_it; ++it, slot ^= 1) {
            DLPD_SYNC();                               // half-tile `it` has landed; everyone is done with the other buffers
            const Item w = item(p, bx, it);
            if (w.pair != cur_pair) {
                if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
                cur_pair = w.pair;
                bestA = bestB = ~0ull;
            }
            if (w.half == 0) issue_h1(p, w, hbuf);     // needed after the transform passes of this half
            cp_async_commit();
#if !defined(DLPD_ABLATE_EPI_LOAD)
            if (it + 1 < n_it) issue_data(p, item(p, bx, it + 1), bufs + (slot ^ 1) * BUF);
#endif
            cp_async_commit();
            E::transform_tile(tm, bufs + slot * BUF, tw0);
            if (w.half == 0) {
                cp_async_wait_allow1();                // my h1 copies are done (the prefetch may still be in flight)
                DLPD_SYNC();                           // ... and everybody else's
            }
            E::score_voxels(p, tm, bufs + slot * BUF, hbuf, w.pair, 2 * (w.r4 / T), 2 * (w.r4 % T) + w.half, bestA, bestB);
            cp_async_wait_pending();
        }
        if (cur_pair >= 0) flush_best(p, cur_pair, bestA, bestB, red);
    }
};

