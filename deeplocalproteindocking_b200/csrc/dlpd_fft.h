// In-shared-memory mixed-radix FFT building blocks (sign +:  X[k] = sum_t x[t] e^{+2 pi i k t / n}).
//
// Every transform of the docking search is a length-N = 2L line whose input occupies the first half
// (zero padding, Docker.py:225 / DockingModels.py:71 pad L -> 2L) or whose output is wanted in full:
//   forward  = decimation in frequency, natural order in -> digit-reversed order out; the outermost radix-2 step
//              has a zero upper half, so it degenerates to "copy + twiddle" and is fused with the next radix
//              ("pruned" pass: only the first half of the line is read);
//   inverse  = decimation in time, digit-reversed in -> natural out, with the mirrored factor list.
// Forward and inverse share the factor list N = 2 * R0 * TA * TB, so the permuted spectra never have to be
// reordered: the receptor spectrum is produced by the same forward code and multiplied pointwise.
//
// Two elementary radices are fused per pass in registers (RA x RB values per task), so a line needs at most two
// shared-memory passes per direction:   head = (2, R0) at block N,   tail = (TA, TB) at block N / (2 R0).
//
// Work is distributed as "tasks" = (butterfly, line) pairs with the LINE index fastest, so that
// consecutive lanes touch consecutive lines: shared-memory accesses are conflict free for any
// butterfly stride and twiddles are (almost) warp-uniform.
#pragma once
#include "dlpd_platform.h"

#ifndef DLPD_PLANE_ALT_PLAN
#define DLPD_PLANE_ALT_PLAN 0
#endif
#ifndef DLPD_MUL_EARLY
#define DLPD_MUL_EARLY 1        // pointwise multipliers of a MID pass are fetched before the task's shared-memory loads (5.37 -> 5.33 ms)
#endif
#ifndef DLPD_DUAL_TASKS
#define DLPD_DUAL_TASKS 0
#endif

// Complex arithmetic on packed fp32x2 instructions (sm_100 FADD2 / FMUL2 / FFMA2): a complex add is ONE instruction,
// a complex multiply TWO (the swap / negate of halves folds into operand modifiers).
DLPD_DEV cpx cadd(cpx a, cpx b) { return fadd2_(a, b); }
DLPD_DEV cpx csub(cpx a, cpx b) { return fadd2_(a, make_float2(-b.x, -b.y)); }
DLPD_DEV cpx cmul_i(cpx a) { return make_float2(-a.y, a.x); }          // a * (+i)
DLPD_DEV cpx cscale(cpx a, float s) { return fmul2_(make_float2(s, s), a); }
DLPD_DEV cpx cmul(cpx a, cpx w) {                                       // a.x * w + a.y * (i w)
    const cpx t = fmul2_(make_float2(a.y, a.y), make_float2(w.y, w.x));
    return ffma2_(make_float2(a.x, a.x), w, make_float2(-t.x, t.y));
}
DLPD_DEV cpx caxpy(float s, cpx x, cpx y) { return ffma2_(make_float2(s, s), x, y); }   // s * x + y

// a * e^{+2 pi i m / 16} for an index that is a compile-time constant once the butterfly loops are unrolled: the
// trivial rotations cost nothing, the others multiply by literal constants (no table load)
DLPD_DEV float w16_cos(int m) {
    return m == 0 ? 1.0f : m == 1 ? 0.92387953251128674f : m == 2 ? 0.70710678118654752f : m == 3 ? 0.38268343236508977f
         : m == 4 ? 0.0f : m == 5 ? -0.38268343236508977f : m == 6 ? -0.70710678118654752f : m == 7 ? -0.92387953251128674f
         : m == 8 ? -1.0f : m == 9 ? -0.92387953251128674f : m == 10 ? -0.70710678118654752f : m == 11 ? -0.38268343236508977f
         : m == 12 ? 0.0f : m == 13 ? 0.38268343236508977f : m == 14 ? 0.70710678118654752f : 0.92387953251128674f;
}
DLPD_DEV cpx mul_w16(cpx a, int m) {
    m &= 15;
    if (m == 0) return a;
    if (m == 4) return cmul_i(a);
    if (m == 8) return make_float2(-a.x, -a.y);
    if (m == 12) return make_float2(a.y, -a.x);
    return cmul(a, make_float2(w16_cos(m), w16_cos((m + 12) & 15)));       // sin(t) = cos(t - pi/2)
}

// the same for 20th roots of unity (innermost (5,4) pass of the 8 x 20 factorisation)
DLPD_DEV float w20_cos(int m) {
    return m == 0 ? 1.0f : m == 1 ? 0.95105651629515357f : m == 2 ? 0.80901699437494742f : m == 3 ? 0.58778525229247313f
         : m == 4 ? 0.30901699437494742f : m == 5 ? 0.0f : m == 6 ? -0.30901699437494742f : m == 7 ? -0.58778525229247313f
         : m == 8 ? -0.80901699437494742f : m == 9 ? -0.95105651629515357f : m == 10 ? -1.0f : m == 11 ? -0.95105651629515357f
         : m == 12 ? -0.80901699437494742f : m == 13 ? -0.58778525229247313f : m == 14 ? -0.30901699437494742f : m == 15 ? 0.0f
         : m == 16 ? 0.30901699437494742f : m == 17 ? 0.58778525229247313f : m == 18 ? 0.80901699437494742f : 0.95105651629515357f;
}
DLPD_DEV cpx mul_w20(cpx a, int m) {
    m %= 20;
    if (m == 0) return a;
    if (m == 5) return cmul_i(a);
    if (m == 10) return make_float2(-a.x, -a.y);
    if (m == 15) return make_float2(a.y, -a.x);
    return cmul(a, make_float2(w20_cos(m), w20_cos((m + 15) % 20)));       // sin(t) = cos(t - pi/2)
}
// a * w_NB^m with compile-time m, NB | 16 or NB == 20
template <int NB> DLPD_DEV cpx mul_wconst(cpx a, int m) {
    if constexpr (NB == 20) return mul_w20(a, m);
    else return mul_w16(a, (16 / NB) * m);
}

// ---- small DFTs, sign + ------------------------------------------------------------------------
template <int R> struct Dft;

template <> struct Dft<1> {
    static DLPD_DEV void run(cpx*) {}
};

template <> struct Dft<2> {
    static DLPD_DEV void run(cpx* a) {
        cpx t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};

template <> struct Dft<4> {
    static DLPD_DEV void run(cpx* a) {
        cpx s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]);
        cpx s13 = cadd(a[1], a[3]), d13 = cmul_i(csub(a[1], a[3]));
        a[0] = cadd(s02, s13);
        a[1] = cadd(d02, d13);
        a[2] = csub(s02, s13);
        a[3] = csub(d02, d13);
    }
};

template <> struct Dft<5> {
    static DLPD_DEV void run(cpx* a) {
        const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
        const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
        cpx t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
        cpx t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
        cpx m1 = caxpy(c2, t2, caxpy(c1, t1, a[0]));
        cpx m2 = caxpy(c1, t2, caxpy(c2, t1, a[0]));
        cpx n1 = cmul_i(caxpy(s2, t4, cscale(t3, s1)));
        cpx n2 = cmul_i(caxpy(-s1, t4, cscale(t3, s2)));
        a[0] = cadd(a[0], cadd(t1, t2));
        a[1] = cadd(m1, n1);
        a[4] = csub(m1, n1);
        a[2] = cadd(m2, n2);
        a[3] = csub(m2, n2);
    }
};

template <> struct Dft<8> {
    static DLPD_DEV void run(cpx* a) {
        const float h = 0.70710678118654752f;
        cpx e[4] = {a[0], a[2], a[4], a[6]};
        cpx o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        cpx w1 = cscale(cadd(o[1], cmul_i(o[1])), h);                             // o1 * e^{+i pi/4}
        cpx w2 = cmul_i(o[2]);                                                    // o2 * i
        cpx w3 = cscale(csub(cmul_i(o[3]), o[3]), h);                             // o3 * e^{+3i pi/4}
        a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
        a[1] = cadd(e[1], w1);   a[5] = csub(e[1], w1);
        a[2] = cadd(e[2], w2);   a[6] = csub(e[2], w2);
        a[3] = cadd(e[3], w3);   a[7] = csub(e[3], w3);
    }
};

// ---- factor lists: N = 2 * R0 * TA * TB ------------------------------------------------------------------
template <int N> struct LinePlan;
#define DLPD_LINEPLAN(N_, R0_, TA_, TB_)                                                    \
    template <> struct LinePlan<N_> {                                                       \
        static constexpr int R0 = R0_, TA = TA_, TB = TB_, TAIL = TA_ * TB_;                \
        static_assert(2 * R0_ * TA_ * TB_ == N_, "bad factorisation");                      \
    };
DLPD_LINEPLAN(8, 4, 1, 1)
DLPD_LINEPLAN(10, 5, 1, 1)
DLPD_LINEPLAN(16, 8, 1, 1)
DLPD_LINEPLAN(20, 5, 2, 1)
DLPD_LINEPLAN(32, 4, 4, 1)
DLPD_LINEPLAN(40, 5, 4, 1)
DLPD_LINEPLAN(64, 8, 4, 1)
DLPD_LINEPLAN(80, 5, 8, 1)
DLPD_LINEPLAN(128, 8, 8, 1)
#ifndef DLPD_LINE160_ALT
#define DLPD_LINE160_ALT 0      // 1: z lines of the 80^3 box as (2,4) x (5,4) = 8 x 20 instead of (2,5) x (4,4) = 10 x 16
#endif
#if DLPD_LINE160_ALT
DLPD_LINEPLAN(160, 4, 5, 4)
#else
DLPD_LINEPLAN(160, 5, 4, 4)
#endif
DLPD_LINEPLAN(256, 8, 4, 4)
// alternative factorisation for the plane kernel's x / y passes (private to it: the receptor spectrum is produced by the same
// code): 160 = (2,4) x (5,4) gives 20 / 8 tasks per line, i.e. whole rounds for teams of 128 threads on 32 lines
template <int N> struct PlanePlan : LinePlan<N> {};
#if DLPD_PLANE_ALT_PLAN
template <> struct PlanePlan<160> {
    static constexpr int R0 = 4, TA = 5, TB = 4, TAIL = 20;
};
#elif DLPD_LINE160_ALT
template <> struct PlanePlan<160> {          // the plane keeps 10 x 16 whatever the z lines use (its order is private to it)
    static constexpr int R0 = 5, TA = 4, TB = 4, TAIL = 16;
};
#endif

// hooks of fft_pass: where a task's values come from / go to, and an optional pointwise multiplier
struct NoHook {};
// element `e` (index along the line) of line `l`
template <class F> struct GlobalIn { F f; };      // f(l, e) -> cpx        : read from global instead of shared
template <class F> struct GlobalOut { F f; };     // f(l, e, value)        : write to global instead of shared
template <class F> struct GlobalOutPair { F f; }; // f(l, e, v_e, v_e1)    : elements e (even) and e+1 of line l together
template <class F> struct PointMul { F f; };      // f(l, e) -> cpx        : multiply between the two halves of a MID pass
template <class T> struct is_pair_out : std::false_type {};
template <class F> struct is_pair_out<GlobalOutPair<F> > : std::true_type {};

enum PassKind {
    PASS_DIF = 0,        // forward pair: outer radix RA at block NB, inner radix RB at block NB/RA
    PASS_DIF_PRUNED = 1, // same with RA == 2 and an all-zero upper half: reads only the lower half
    PASS_DIT = 2,        // inverse pair (inner first, then outer)
    PASS_MID = 3         // DIF pair, pointwise multiply, DIT pair on the same registers (no round trip)
};

// One fused pass over NLINES lines of LEN elements held in shared memory `s` (element stride ES, line offsets `off`).
//   NB   block size of the outer radix (LEN % NB == 0);  TWN size of the twiddle table tw[k] = e^{+2 pi i k/TWN}
//   In / Out / Mul: NoHook or the hook structs above
// PK: `tw` is a PACKED twiddle table for this (2, RB) head pass -- entry [j][0..RB) = w_NB^(j + k M2), [j][RB..2RB-1) = w_M^(j k)
// (k = 1..RB-1), one zero pad -- read with RB 16-byte loads instead of 2 RB - 1 8-byte ones (see pack_head_twiddles)
template <int KIND, int RA, int RB, int LEN, int NB, int TWN, int ES, bool PK = false> struct FftTask {
    static constexpr int M = NB / RA;             // outer sub-block
    static constexpr int M2 = M / RB;             // inner sub-block: range of the butterfly index j
    static constexpr int TPL = LEN / (RA * RB);   // tasks per line
    static constexpr int TWO = TWN / NB;          // outer twiddle w_NB^x = tw[TWO * x]
    static constexpr int TWI = TWN / M;           // inner twiddle w_M^x  = tw[TWI * x]
    static constexpr bool FWD = KIND == PASS_DIF || KIND == PASS_DIF_PRUNED || KIND == PASS_MID;
    static constexpr bool INV = KIND == PASS_DIT || KIND == PASS_MID;
    static_assert(KIND != PASS_DIF_PRUNED || RA == 2, "pruned pass is the outermost radix-2 step");
    // innermost pass (one butterfly per block, j == 0): every twiddle is a compile-time power of w_NB, NB | 16
    static constexpr bool CONST_TW = (M2 == 1) && (16 % NB == 0 || NB == 20);
    static_assert(!PK || (RA == 2 && M2 > 1 && !CONST_TW), "packed twiddles are for the (2, R0) head passes");
    static DLPD_DEV void load_packed(cpx (&w)[2 * RB], const cpx* tw, int j) {
        const float4* q = reinterpret_cast<const float4*>(tw + j * 2 * RB);
        DLPD_UNROLL
        for (int u = 0; u < RB; ++u) {
            const float4 v = q[u];
            w[2 * u] = make_float2(v.x, v.y);
            w[2 * u + 1] = make_float2(v.z, v.w);
        }
    }

    // load the RA x RB values of the task whose first element is e0 = blk*NB + j (line l at `p`) and transform them
    template <class In, class Mul>
    static DLPD_DEV void run(cpx (&a)[RA][RB], const cpx* p, const cpx* tw, int l, int e0, int j, In& in, Mul& mul) {
        if constexpr (DLPD_MUL_EARLY && !std::is_same<Mul, NoHook>::value && std::is_same<In, NoHook>::value) {
            // pointwise multipliers (global memory, L2 latency) first: in flight during the loads and the forward butterflies
            cpx mv[RA][RB];
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) mv[i][k] = mul.f(l, e0 + i * M + k * M2);
            }
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) a[i][k] = p[(i * M + k * M2) * ES];
            }
            compute_fwd(a, tw, j);
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) a[i][k] = cmul(a[i][k], mv[i][k]);
            }
            compute_inv(a, tw, j);
        } else {
            // ---- load: a[i][k] <- element e0 + i*M + k*M2
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                if (KIND == PASS_DIF_PRUNED && i > 0) break;
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) {
                    if constexpr (std::is_same<In, NoHook>::value) a[i][k] = p[(i * M + k * M2) * ES];
                    else a[i][k] = in.f(l, e0 + i * M + k * M2);
                }
            }
            compute(a, tw, l, e0, j, mul);
        }
    }
    // the butterflies of a task whose values are already in a[][] (pruned pass: a[0][*] only)
    template <class Mul>
    static DLPD_DEV void compute(cpx (&a)[RA][RB], const cpx* tw, int l, int e0, int j, Mul& mul) {
        compute_fwd(a, tw, j);
        if constexpr (!std::is_same<Mul, NoHook>::value) {
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) a[i][k] = cmul(a[i][k], mul.f(l, e0 + i * M + k * M2));
            }
        }
        compute_inv(a, tw, j);
    }
    static DLPD_DEV void compute_fwd(cpx (&a)[RA][RB], const cpx* tw, int j) {
        if (FWD) {
            cpx w[2 * RB];
            if constexpr (PK) load_packed(w, tw, j);
            // outer DFT over i for each k, then twiddle w_NB^{(j + k*M2) * pa}
            DLPD_UNROLL
            for (int k = 0; k < RB; ++k) {
                cpx c[RA];
                if (KIND == PASS_DIF_PRUNED) {
                    c[0] = a[0][k];
                    c[RA - 1] = a[0][k];
                } else {
                    DLPD_UNROLL
                    for (int i = 0; i < RA; ++i) c[i] = a[i][k];
                    Dft<RA>::run(c);
                }
                a[0][k] = c[0];
                DLPD_UNROLL
                for (int i = 1; i < RA; ++i) {
                    if constexpr (CONST_TW) a[i][k] = mul_wconst<NB>(c[i], k * i);
                    else if constexpr (PK) a[i][k] = cmul(c[i], w[k]);
                    else a[i][k] = cmul(c[i], tw[TWO * (j + k * M2) * i]);
                }
            }
            // inner DFT over k for each pa, then twiddle w_M^{j * s}  (j == 0 when M2 == 1: no twiddle)
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                Dft<RB>::run(a[i]);
                if constexpr (M2 > 1) {
                    DLPD_UNROLL
                    for (int k = 1; k < RB; ++k) a[i][k] = cmul(a[i][k], PK ? w[RB + k - 1] : tw[TWI * j * k]);
                }
            }
        }
    }
    static DLPD_DEV void compute_inv(cpx (&a)[RA][RB], const cpx* tw, int j) {
        if (INV) {
            cpx w[2 * RB];
            if constexpr (PK) load_packed(w, tw, j);
            // inner: twiddle w_M^{j * s}, DFT over s for each pa
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                if constexpr (M2 > 1) {
                    DLPD_UNROLL
                    for (int k = 1; k < RB; ++k) a[i][k] = cmul(a[i][k], PK ? w[RB + k - 1] : tw[TWI * j * k]);
                }
                Dft<RB>::run(a[i]);
            }
            // outer: twiddle w_NB^{(j + k*M2) * pa}, DFT over pa for each k
            DLPD_UNROLL
            for (int k = 0; k < RB; ++k) {
                cpx c[RA];
                c[0] = a[0][k];
                DLPD_UNROLL
                for (int i = 1; i < RA; ++i) {
                    if constexpr (CONST_TW) c[i] = mul_wconst<NB>(a[i][k], k * i);
                    else if constexpr (PK) c[i] = cmul(a[i][k], w[k]);
                    else c[i] = cmul(a[i][k], tw[TWO * (j + k * M2) * i]);
                }
                Dft<RA>::run(c);
                DLPD_UNROLL
                for (int i = 0; i < RA; ++i) a[i][k] = c[i];
            }
        }
    }
};

template <int KIND, int RA, int RB, int NLINES, int LEN, int NB, int TWN, int ES, bool PK = false, class TM, class LineOff,
          class In = NoHook, class Out = NoHook, class Mul = NoHook>
DLPD_DEV void fft_pass_team(const TM& tm, cpx* s, const cpx* tw, LineOff off, In in = In(), Out out = Out(), Mul mul = Mul()) {
    typedef FftTask<KIND, RA, RB, LEN, NB, TWN, ES, PK> T;
    constexpr int M = T::M, M2 = T::M2, TPL = T::TPL;
    if constexpr (is_pair_out<Out>::value) {
        // two neighbouring butterflies (elements e, e+1) per task: the hook stores 16-byte pairs
        static_assert(M2 % 2 == 0, "paired output needs an even inner sub-block");
        DLPD_TEAM_TASKS(tm, t, NLINES * (TPL / 2)) {
            const int l = t % NLINES;
            const int b = 2 * (t / NLINES);
            const int blk = b / M2;
            const int j = b - blk * M2;
            const int e0 = blk * NB + j;
            const cpx* p = s + off(l) + e0 * ES;
            cpx a0[RA][RB], a1[RA][RB];
            T::run(a0, p, tw, l, e0, j, in, mul);
            T::run(a1, p + ES, tw, l, e0 + 1, j + 1, in, mul);
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) out.f(l, e0 + i * M + k * M2, a0[i][k], a1[i][k]);
            }
        }
    } else if constexpr ((DLPD_DUAL_TASKS || TM::DUAL) && RA * RB <= 10 && std::is_same<Out, NoHook>::value && std::is_same<Mul, NoHook>::value) {
        // two tasks per iteration (t and t + NT): both tasks' loads are issued before either butterfly, so the
        // shared-memory latency of the second hides behind the arithmetic of the first (20 complex values in registers)
        constexpr int NTASK = NLINES * TPL;
        DLPD_TEAM_TASKS2(tm, t, NTASK) {
            const int t1 = (t + tm.NT < NTASK) ? t + tm.NT : t;       // no second task: redo the first, store once
            const int l0 = t % NLINES, b0 = t / NLINES, l1 = t1 % NLINES, b1 = t1 / NLINES;
            const int k0 = b0 / M2, j0 = b0 - k0 * M2, k1 = b1 / M2, j1 = b1 - k1 * M2;
            const int e0 = k0 * NB + j0, e1 = k1 * NB + j1;
            cpx* p0 = s + off(l0) + e0 * ES;
            cpx* p1 = s + off(l1) + e1 * ES;
            cpx a0[RA][RB], a1[RA][RB];
            T::run(a0, p0, tw, l0, e0, j0, in, mul);
            T::run(a1, p1, tw, l1, e1, j1, in, mul);
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) p0[(i * M + k * M2) * ES] = a0[i][k];
            }
            if (t1 != t) {
                DLPD_UNROLL
                for (int i = 0; i < RA; ++i) {
                    DLPD_UNROLL
                    for (int k = 0; k < RB; ++k) p1[(i * M + k * M2) * ES] = a1[i][k];
                }
            }
        }
    } else {
        DLPD_TEAM_TASKS(tm, t, NLINES * TPL) {
            const int l = t % NLINES;
            const int b = t / NLINES;
            const int blk = b / M2;
            const int j = b - blk * M2;
            const int e0 = blk * NB + j;       // element index of (pa = 0, s = 0)
            cpx* p = s + off(l) + e0 * ES;
            cpx a[RA][RB];
            T::run(a, p, tw, l, e0, j, in, mul);
            // ---- store to the same positions
            DLPD_UNROLL
            for (int i = 0; i < RA; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) {
                    if constexpr (std::is_same<Out, NoHook>::value) p[(i * M + k * M2) * ES] = a[i][k];
                    else out.f(l, e0 + i * M + k * M2, a[i][k]);
                }
            }
        }
    }
    tm.sync();
}
template <int KIND, int RA, int RB, int NLINES, int LEN, int NB, int TWN, int ES, int NT, class LineOff,
          class In = NoHook, class Out = NoHook, class Mul = NoHook>
DLPD_DEV void fft_pass(cpx* s, const cpx* tw, LineOff off, In in = In(), Out out = Out(), Mul mul = Mul()) {
    fft_pass_team<KIND, RA, RB, NLINES, LEN, NB, TWN, ES>(CtaTeam<NT>(), s, tw, off, in, out, mul);
}

// packed twiddles of the (2, R0) head pass of a length-N line (N entries): see FftTask<..., PK = true>
template <int N, int R0> DLPD_DEV cpx packed_head_twiddle(const cpx* tw, int t) {
    constexpr int M2 = N / (2 * R0);
    const int j = t / (2 * R0), k = t % (2 * R0);
    if (k < R0) return tw[j + k * M2];                      // w_N^(j + k M2)
    if (k < 2 * R0 - 1) return tw[(2 * j * (k - R0 + 1)) % N];  // w_(N/2)^(j k'),  k' = k - R0 + 1
    return make_float2(0.f, 0.f);
}

// ONE task of a plain in-place pass (no hooks, no loop, no barrier): lets a caller put a butterfly task and unrelated
// register-only work into the same basic block so that the scheduler interleaves them (EpilogueFused)
template <int KIND, int RA, int RB, int NLINES, int LEN, int NB, int TWN, int ES, class LineOff>
DLPD_DEV void fft_one_task(cpx* s, const cpx* tw, LineOff off, int t) {
    typedef FftTask<KIND, RA, RB, LEN, NB, TWN, ES> T;
    constexpr int M = T::M, M2 = T::M2;
    const int l = t % NLINES, b = t / NLINES, blk = b / M2, j = b - blk * M2, e0 = blk * NB + j;
    cpx* p = s + off(l) + e0 * ES;
    cpx a[RA][RB];
    NoHook none;
    T::run(a, p, tw, l, e0, j, none, none);
    DLPD_UNROLL
    for (int i = 0; i < RA; ++i) {
        DLPD_UNROLL
        for (int k = 0; k < RB; ++k) p[(i * M + k * M2) * ES] = a[i][k];
    }
}

// Pruned forward head whose inputs were fetched into registers ahead of time (persistent kernels: the next work item's
// global loads are issued while the current one finishes).  Thread `tid` owns tasks tid, tid + NT, ...: exactly NPRE each.
template <int RB, int NLINES, int LEN, int TWN, int ES, int NPRE, class TM, class LineOff>
struct PrunedHeadPre {
    typedef FftTask<PASS_DIF_PRUNED, 2, RB, LEN, LEN, TWN, ES> T;
    static_assert(NLINES * T::TPL == NPRE * TM::NT, "every thread owns exactly NPRE tasks");
    // fetch: v[u][k] <- element (j + k*M2) of line l of task u, through `in`
    template <class In> static DLPD_DEV void fetch(const TM& tm, cpx (&v)[NPRE][RB], In in) {
        DLPD_UNROLL
        for (int u = 0; u < NPRE; ++u) {
            const int t = tm.tid() + u * TM::NT, l = t % NLINES, j = t / NLINES;
            DLPD_UNROLL
            for (int k = 0; k < RB; ++k) v[u][k] = in.f(l, j + k * T::M2);
        }
    }
    static DLPD_DEV void run(const TM& tm, cpx* s, const cpx* tw, LineOff off, const cpx (&v)[NPRE][RB]) {
        NoHook none;
        DLPD_UNROLL
        for (int u = 0; u < NPRE; ++u) {
            const int t = tm.tid() + u * TM::NT, l = t % NLINES, j = t / NLINES;
            cpx a[2][RB];
            DLPD_UNROLL
            for (int k = 0; k < RB; ++k) a[0][k] = v[u][k];
            T::compute(a, tw, l, j, j, none);
            cpx* p = s + off(l) + j * ES;
            DLPD_UNROLL
            for (int i = 0; i < 2; ++i) {
                DLPD_UNROLL
                for (int k = 0; k < RB; ++k) p[(i * T::M + k * T::M2) * ES] = a[i][k];
            }
        }
        tm.sync();
    }
};

// ---- whole-line helpers -------------------------------------------------------------------------------------
// forward head: pruned (2, R0) pair at block N; reads elements [0, N/2) only (optionally from global via `in`)
template <int N, int NLINES, int ES, int NT, class P = LinePlan<N>, class LineOff, class In = NoHook>
DLPD_DEV void fft_fwd_head(cpx* s, const cpx* tw, LineOff off, In in = In()) {
    fft_pass<PASS_DIF_PRUNED, 2, P::R0, NLINES, N, N, N, ES, NT>(s, tw, off, in);
}
// forward tail: (TA, TB) pair at block N / (2 R0)   (nothing to do when TAIL == 1)
template <int N, int NLINES, int ES, int NT, class P = LinePlan<N>, class LineOff>
DLPD_DEV void fft_fwd_tail(cpx* s, const cpx* tw, LineOff off) {
    if constexpr (P::TAIL > 1) fft_pass<PASS_DIF, P::TA, P::TB, NLINES, N, P::TAIL, N, ES, NT>(s, tw, off);
}
// forward tail, pointwise multiply, inverse tail in one pass (a plain multiply pass when TAIL == 1)
template <int N, int NLINES, int ES, int NT, class P = LinePlan<N>, class LineOff, class Mul>
DLPD_DEV void fft_mid(cpx* s, const cpx* tw, LineOff off, Mul mul) {
    fft_pass<PASS_MID, P::TA, P::TB, NLINES, N, P::TAIL, N, ES, NT>(s, tw, off, NoHook(), NoHook(), mul);
}
template <int N, int NLINES, int ES, int NT, class P = LinePlan<N>, class LineOff, class In = NoHook>
DLPD_DEV void fft_inv_tail(cpx* s, const cpx* tw, LineOff off, In in = In()) {
    if constexpr (P::TAIL > 1 || !std::is_same<In, NoHook>::value)
        fft_pass<PASS_DIT, P::TA, P::TB, NLINES, N, P::TAIL, N, ES, NT>(s, tw, off, in);
}
// inverse head: (2, R0) pair at block N (optionally writing to global via `out`)
template <int N, int NLINES, int ES, int NT, class P = LinePlan<N>, class LineOff, class Out = NoHook>
DLPD_DEV void fft_inv_head(cpx* s, const cpx* tw, LineOff off, Out out = Out()) {
    fft_pass<PASS_DIT, 2, P::R0, NLINES, N, N, N, ES, NT>(s, tw, off, NoHook(), out);
}

// ---- the same helpers for an explicit team (SubTeam / CtaTeam) ------------------------------------------------
template <int N, int NLINES, int ES, class P = LinePlan<N>, bool PK = false, class TM, class LineOff, class In = NoHook>
DLPD_DEV void team_fwd_head(const TM& tm, cpx* s, const cpx* tw, LineOff off, In in = In()) {
    fft_pass_team<PASS_DIF_PRUNED, 2, P::R0, NLINES, N, N, N, ES, PK>(tm, s, tw, off, in);
}
template <int N, int NLINES, int ES, class P = LinePlan<N>, class TM, class LineOff>
DLPD_DEV void team_fwd_tail(const TM& tm, cpx* s, const cpx* tw, LineOff off) {
    if constexpr (P::TAIL > 1) fft_pass_team<PASS_DIF, P::TA, P::TB, NLINES, N, P::TAIL, N, ES>(tm, s, tw, off);
}
template <int N, int NLINES, int ES, class P = LinePlan<N>, class TM, class LineOff, class Mul>
DLPD_DEV void team_mid(const TM& tm, cpx* s, const cpx* tw, LineOff off, Mul mul) {
    fft_pass_team<PASS_MID, P::TA, P::TB, NLINES, N, P::TAIL, N, ES>(tm, s, tw, off, NoHook(), NoHook(), mul);
}
template <int N, int NLINES, int ES, class P = LinePlan<N>, class TM, class LineOff, class In = NoHook>
DLPD_DEV void team_inv_tail(const TM& tm, cpx* s, const cpx* tw, LineOff off, In in = In()) {
    if constexpr (P::TAIL > 1 || !std::is_same<In, NoHook>::value)
        fft_pass_team<PASS_DIT, P::TA, P::TB, NLINES, N, P::TAIL, N, ES>(tm, s, tw, off, in);
}
template <int N, int NLINES, int ES, class P = LinePlan<N>, bool PK = false, class TM, class LineOff, class Out = NoHook>
DLPD_DEV void team_inv_head(const TM& tm, cpx* s, const cpx* tw, LineOff off, Out out = Out()) {
    fft_pass_team<PASS_DIT, 2, P::R0, NLINES, N, N, N, ES, PK>(tm, s, tw, off, NoHook(), out);
}
