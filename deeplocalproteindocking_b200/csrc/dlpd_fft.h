// In-shared-memory mixed-radix FFT building blocks (sign +:  X[k] = sum_t x[t] e^{+2 pi i k t / n}).
//
// Every transform of the docking search is a length-N = 2L line whose input occupies the first half
// (zero padding, Docker.py:225 / DockingModels.py:71 pad L -> 2L) or whose output is wanted in full:
//   forward  = pruned radix-2 DIF step (copy + twiddle) followed by two independent length-L DIF FFTs
//              (natural order in, digit-reversed order out);
//   inverse  = two length-L DIT FFTs (digit-reversed in, natural out) followed by one radix-2 DIT step.
// Forward and inverse use the same factor list mirrored, so the permuted spectra never have to be
// reordered: the receptor spectrum is produced by the same forward code and multiplied pointwise.
//
// Work is distributed as "tasks" = (butterfly, line) pairs with the LINE index fastest, so that
// consecutive lanes touch consecutive lines: shared-memory accesses are conflict free for any
// butterfly stride and twiddles are (almost) warp-uniform.
#pragma once
#include "dlpd_platform.h"

DLPD_HD cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
DLPD_HD cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
DLPD_HD cpx cmul(cpx a, cpx b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DLPD_HD cpx cmul_i(cpx a) { return make_float2(-a.y, a.x); }          // a * (+i)
DLPD_HD cpx cscale(cpx a, float s) { return make_float2(a.x * s, a.y * s); }

// ---- small DFTs, sign + ------------------------------------------------------------------------
template <int R> struct Dft;

template <> struct Dft<2> {
    static DLPD_HD void run(cpx* a) {
        cpx t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};

template <> struct Dft<4> {
    static DLPD_HD void run(cpx* a) {
        cpx s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]);
        cpx s13 = cadd(a[1], a[3]), d13 = cmul_i(csub(a[1], a[3]));
        a[0] = cadd(s02, s13);
        a[1] = cadd(d02, d13);
        a[2] = csub(s02, s13);
        a[3] = csub(d02, d13);
    }
};

template <> struct Dft<5> {
    static DLPD_HD void run(cpx* a) {
        const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
        const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
        cpx t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
        cpx t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
        cpx m1 = make_float2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
        cpx m2 = make_float2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
        cpx n1 = cmul_i(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        cpx n2 = cmul_i(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        a[0] = make_float2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
        a[1] = cadd(m1, n1);
        a[4] = csub(m1, n1);
        a[2] = cadd(m2, n2);
        a[3] = csub(m2, n2);
    }
};

template <> struct Dft<8> {
    static DLPD_HD void run(cpx* a) {
        const float h = 0.70710678118654752f;
        cpx e[4] = {a[0], a[2], a[4], a[6]};
        cpx o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        cpx w1 = make_float2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y));      // o1 * e^{+i pi/4}
        cpx w2 = cmul_i(o[2]);                                                    // o2 * i
        cpx w3 = make_float2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y));     // o3 * e^{+3i pi/4}
        a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
        a[1] = cadd(e[1], w1);   a[5] = csub(e[1], w1);
        a[2] = cadd(e[2], w2);   a[6] = csub(e[2], w2);
        a[3] = cadd(e[3], w3);   a[7] = csub(e[3], w3);
    }
};

// ---- factor lists of the HALF length M = L (first = outermost DIF stage) ---------------------------
template <int M> struct HalfPlan;
#define DLPD_HALFPLAN(M_, NS_, ...)                                   \
    template <> struct HalfPlan<M_> {                                 \
        static constexpr int NS = NS_;                                \
        static DLPD_CE int radix(int i) { constexpr int r[] = {__VA_ARGS__}; return r[i]; } \
    };
DLPD_HALFPLAN(4, 1, 4)
DLPD_HALFPLAN(5, 1, 5)
DLPD_HALFPLAN(8, 1, 8)
DLPD_HALFPLAN(10, 2, 5, 2)
DLPD_HALFPLAN(16, 2, 4, 4)
DLPD_HALFPLAN(20, 2, 5, 4)
DLPD_HALFPLAN(32, 2, 8, 4)
DLPD_HALFPLAN(40, 2, 5, 8)
DLPD_HALFPLAN(64, 2, 8, 8)
DLPD_HALFPLAN(80, 3, 5, 4, 4)
DLPD_HALFPLAN(128, 3, 8, 4, 4)

// block size seen by DIF stage i of a length-M transform: M / (r0 * ... * r_{i-1})
template <int M, int I> struct StageBlock {
    static constexpr int value = StageBlock<M, I - 1>::value / HalfPlan<M>::radix(I - 1);
};
template <int M> struct StageBlock<M, 0> { static constexpr int value = M; };

// One stage over a set of lines held in shared memory.
//   R        radix;  DIT: twiddle-then-butterfly (inverse), else butterfly-then-twiddle (forward)
//   NLINES   number of lines, LEN line length, NB block size of this stage (LEN % NB == 0)
//   TWN      size of the twiddle table tw[k] = e^{+2 pi i k / TWN};  TWS = TWN / (transform size the
//            blocks belong to) * (that size / NB), i.e. twiddle index = TWMUL * j * p
//   LineOff  functor: line index -> offset (in cpx) of element 0;  ES element stride (in cpx)
template <int R, bool DIT, int NLINES, int LEN, int NB, int TWMUL, int ES, int NT, class LineOff>
DLPD_DEV void fft_stage(cpx* s, const cpx* tw, LineOff off) {
    constexpr int MB = NB / R;             // sub-block size = butterfly stride
    constexpr int BPL = LEN / R;           // butterflies per line
    DLPD_FOR_TASKS(t, NLINES * BPL, NT) {
        const int l = t % NLINES;
        const int b = t / NLINES;
        const int blk = b / MB;
        const int j = b - blk * MB;
        cpx* p = s + off(l) + (blk * NB + j) * ES;
        cpx a[R];
        DLPD_UNROLL
        for (int q = 0; q < R; ++q) a[q] = p[q * MB * ES];
        if (DIT) {
            DLPD_UNROLL
            for (int q = 1; q < R; ++q) a[q] = cmul(a[q], tw[TWMUL * j * q]);
        }
        Dft<R>::run(a);
        if (!DIT) {
            DLPD_UNROLL
            for (int q = 1; q < R; ++q) a[q] = cmul(a[q], tw[TWMUL * j * q]);
        }
        DLPD_UNROLL
        for (int q = 0; q < R; ++q) p[q * MB * ES] = a[q];
    }
    DLPD_SYNC();
}

// All DIF stages of NLINES length-M lines (forward, natural -> permuted).  tw has TWN = 2M entries.
template <int M, int NLINES, int ES, int NT, int I, class LineOff> struct DifStages {
    static DLPD_DEV void run(cpx* s, const cpx* tw, LineOff off) {
        constexpr int NB = StageBlock<M, I>::value;
        fft_stage<HalfPlan<M>::radix(I), false, NLINES, M, NB, (2 * M) / NB, ES, NT>(s, tw, off);
        if constexpr (I + 1 < HalfPlan<M>::NS) DifStages<M, NLINES, ES, NT, I + 1, LineOff>::run(s, tw, off);
    }
};

// All DIT stages (inverse, permuted -> natural): mirror order of DifStages.
template <int M, int NLINES, int ES, int NT, int I, class LineOff> struct DitStages {
    static DLPD_DEV void run(cpx* s, const cpx* tw, LineOff off) {
        constexpr int NB = StageBlock<M, I>::value;
        if constexpr (I + 1 < HalfPlan<M>::NS) DitStages<M, NLINES, ES, NT, I + 1, LineOff>::run(s, tw, off);
        fft_stage<HalfPlan<M>::radix(I), true, NLINES, M, NB, (2 * M) / NB, ES, NT>(s, tw, off);
    }
};

// Offsets of the 2*NLINES half lines (half h = l / NLINES, line = l % NLINES) given the base functor.
template <int NLINES, int HALF_OFF, class LineOff> struct HalfLines {
    LineOff base;
    DLPD_DEV int operator()(int l) const {
        const int h = l / NLINES;
        return base(l - h * NLINES) + h * HALF_OFF;
    }
};

// Forward sub-transforms: the caller has already applied the pruned radix-2 step
//   s[j + M] = s[j] * tw[j]   (j < M),  then both halves get a length-M DIF FFT.
template <int M, int NLINES, int ES, int NT, class LineOff>
DLPD_DEV void fft_forward_halves(cpx* s, const cpx* tw, LineOff off) {
    HalfLines<NLINES, M * ES, LineOff> h{off};
    DifStages<M, 2 * NLINES, ES, NT, 0, HalfLines<NLINES, M * ES, LineOff>>::run(s, tw, h);
}

// Full inverse of a length-2M line: DIT on the halves, then the radix-2 DIT step of block 2M.
template <int M, int NLINES, int ES, int NT, class LineOff>
DLPD_DEV void fft_inverse_full(cpx* s, const cpx* tw, LineOff off) {
    HalfLines<NLINES, M * ES, LineOff> h{off};
    DitStages<M, 2 * NLINES, ES, NT, 0, HalfLines<NLINES, M * ES, LineOff>>::run(s, tw, h);
    fft_stage<2, true, NLINES, 2 * M, 2 * M, 1, ES, NT>(s, tw, off);
}
