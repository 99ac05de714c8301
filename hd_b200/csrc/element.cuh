// element.cuh — feBeam2d element routines, evaluated exactly in the reference's order.
//
// Reference: feBeam2d.go:11-19 (distance), :22-140 (getStiffBeam2d incl. the literal pin
// condensation and its write-back index rule), :142-172 (getCoordTransStiffBeam2d),
// model.go:619-620 ((T^T k) T through gonum's Dgemm: inner index ascending, zero multipliers
// skipped, multiply then add).  Go on amd64 never contracts a*b+c, so every product and sum
// here is an explicitly rounded FP64 operation (__dmul_rn / __dadd_rn ... on the device,
// -ffp-contract=off on the host): the exact zeros that define the sparsity pattern
// (model.go:645-649) and the assembled values come out bit-identical on CPU and GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define HD_FN __device__ __forceinline__
#define HD_MUL(a, b) __dmul_rn((a), (b))
#define HD_ADD(a, b) __dadd_rn((a), (b))
#define HD_SUB(a, b) __dsub_rn((a), (b))
#define HD_DIV(a, b) __ddiv_rn((a), (b))
#define HD_SQRT(a) __dsqrt_rn((a))
#else
#define HD_FN static inline
#define HD_MUL(a, b) ((a) * (b))
#define HD_ADD(a, b) ((a) + (b))
#define HD_SUB(a, b) ((a) - (b))
#define HD_DIV(a, b) ((a) / (b))
#define HD_SQRT(a) sqrt((a))
#endif

namespace hdb {

// Go math.Hypot scaling formula (same in the pure-Go and the amd64 assembly versions)
HD_FN double go_hypot(double p, double q)
{
    p = fabs(p);
    q = fabs(q);
    if (isinf(p) || isinf(q)) return INFINITY;
    if (isnan(p) || isnan(q)) return NAN;
    if (p < q) { double t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = HD_DIV(q, p);
    return HD_MUL(p, HD_SQRT(HD_ADD(1.0, HD_MUL(q, q))));
}

// feBeam2d.go:11-19: start minus end
HD_FN double beam_length(double x0, double y0, double x1, double y1)
{
    return go_hypot(HD_SUB(x0, x1), HD_SUB(y0, y1));
}

// the five distinct values of the unpinned matrix, feBeam2d.go:28-56
struct BeamCoef {
    double EFL, EJL3, EJL2, EJL, EJL4;
};

HD_FN BeamCoef beam_coef(double A, double J, double E, double length)
{
    BeamCoef c;
    c.EFL = HD_DIV(HD_MUL(E, A), length);
    c.EJL3 = HD_DIV(HD_MUL(HD_MUL(12.0, E), J), HD_MUL(HD_MUL(length, length), length));
    c.EJL2 = HD_DIV(HD_MUL(HD_MUL(6.0, E), J), HD_MUL(length, length));
    c.EJL = HD_DIV(HD_MUL(HD_MUL(2.0, E), J), length);
    c.EJL4 = HD_MUL(2.0, c.EJL);
    return c;
}

// feBeam2d.go:22-140.  k is a 6x6 row-major matrix in registers / local storage.
// pinMask bit fr set <=> Pins[beam][fr].
HD_FN void stiff_beam2d(const BeamCoef &c, uint32_t pinMask, double k[36])
{
#pragma unroll
    for (int q = 0; q < 36; q++) k[q] = 0.0;
    k[0 * 6 + 0] = c.EFL;   k[0 * 6 + 3] = -c.EFL;  k[3 * 6 + 0] = -c.EFL;  k[3 * 6 + 3] = c.EFL;
    k[1 * 6 + 1] = c.EJL3;  k[4 * 6 + 4] = c.EJL3;  k[1 * 6 + 4] = -c.EJL3; k[4 * 6 + 1] = -c.EJL3;
    k[1 * 6 + 2] = c.EJL2;  k[2 * 6 + 1] = c.EJL2;  k[1 * 6 + 5] = c.EJL2;  k[5 * 6 + 1] = c.EJL2;
    k[2 * 6 + 4] = -c.EJL2; k[4 * 6 + 2] = -c.EJL2; k[4 * 6 + 5] = -c.EJL2; k[5 * 6 + 4] = -c.EJL2;
    k[2 * 6 + 5] = c.EJL;   k[5 * 6 + 2] = c.EJL;
    k[2 * 6 + 2] = c.EJL4;  k[5 * 6 + 5] = c.EJL4;
    if (pinMask == 0) return;
    // :58-137.  KS = KBB - (1/KAA) * (KBA * KAB) on the 5x5 complement of fr; the write-back
    // loop compares the REDUCED index with fr, so full row/column fr+1 keeps its old value.
#pragma unroll
    for (int fr = 0; fr < 6; fr++) {
        if (!((pinMask >> fr) & 1u)) continue;
        const double KAA = k[fr * 6 + fr];
        const double f = HD_DIV(1.0, KAA);
        double col[6], row[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            col[i] = k[i * 6 + fr]; // KBA
            row[i] = k[fr * 6 + i]; // KAB
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            if (i == fr || i == fr + 1) continue; // fr: zeroed below; fr+1: the write-back quirk
#pragma unroll
            for (int j = 0; j < 6; j++) {
                if (j == fr || j == fr + 1) continue;
                // gonum Dgemm skips a zero multiplier KBA[i] (leaves 0), then Scale multiplies by f
                const double prod = (col[i] != 0.0) ? HD_MUL(col[i], row[j]) : 0.0;
                k[i * 6 + j] = HD_SUB(k[i * 6 + j], HD_MUL(f, prod));
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            k[i * 6 + fr] = 0.0;
            k[fr * 6 + i] = 0.0;
        }
    }
}

// feBeam2d.go:175-275 without the axial-force factor: the geometric stiffness of a beam is
// axialForce * g(length, pins), axialForce = max(0, bf[0], -bf[3]) (:180-192, applied by the caller).
HD_FN void geom_unit_beam2d(double length, uint32_t pinMask, double k[36])
{
#pragma unroll
    for (int q = 0; q < 36; q++) k[q] = 0.0;
    const bool p2 = (pinMask >> 2) & 1u, p5 = (pinMask >> 5) & 1u;
    if (!p2 && !p5) { // :197
        const double G1 = HD_DIV(6., HD_MUL(5., length));
        k[1 * 6 + 1] = G1;  k[4 * 6 + 4] = G1;  k[1 * 6 + 4] = -G1; k[4 * 6 + 1] = -G1;
        k[1 * 6 + 2] = 0.1; k[2 * 6 + 1] = 0.1; k[1 * 6 + 5] = 0.1; k[5 * 6 + 1] = 0.1;
        k[2 * 6 + 4] = -0.1; k[4 * 6 + 2] = -0.1; k[5 * 6 + 4] = -0.1; k[4 * 6 + 5] = -0.1;
        const double G2 = HD_DIV(HD_MUL(2., length), 15.);
        k[2 * 6 + 2] = G2;  k[5 * 6 + 5] = G2;
        const double G3 = HD_DIV(-length, 30.);
        k[5 * 6 + 2] = G3;  k[2 * 6 + 5] = G3;
    } else if (p2 && !p5) { // :226
        const double G1 = HD_DIV(6., HD_MUL(5., length));
        k[1 * 6 + 1] = G1;  k[4 * 6 + 4] = G1;  k[1 * 6 + 4] = -G1; k[4 * 6 + 1] = -G1;
        k[1 * 6 + 5] = HD_DIV(1., 5.);  k[5 * 6 + 1] = HD_DIV(1., 5.);
        k[4 * 6 + 5] = HD_DIV(-1., 5.); k[5 * 6 + 4] = HD_DIV(-1., 5.);
        k[5 * 6 + 5] = HD_DIV(length, 5.0);
    } else if (!p2 && p5) { // :245
        const double G1 = HD_DIV(6., HD_MUL(5., length));
        k[1 * 6 + 1] = G1;  k[4 * 6 + 4] = G1;  k[1 * 6 + 4] = -G1; k[4 * 6 + 1] = -G1;
        k[1 * 6 + 2] = HD_DIV(1., 5.);  k[2 * 6 + 1] = HD_DIV(1., 5.);
        k[4 * 6 + 2] = HD_DIV(-1., 5.); k[2 * 6 + 4] = HD_DIV(-1., 5.);
        k[2 * 6 + 2] = HD_DIV(length, 5.0);
    } else { // :264
        const double g = HD_DIV(1.0, length);
        k[2 * 6 + 2] = g;  k[5 * 6 + 5] = g;
        k[2 * 6 + 5] = -g; k[5 * 6 + 2] = -g;
    }
}

// feBeam2d.go:142-172: direction cosines use end minus start
HD_FN void beam_cs(double x0, double y0, double x1, double y1, double length, double &c, double &s)
{
    c = HD_DIV(HD_SUB(x1, x0), length);
    s = HD_DIV(HD_SUB(y1, y0), length);
}

HD_FN void trans_beam2d(double c, double s, double t[36])
{
#pragma unroll
    for (int q = 0; q < 36; q++) t[q] = 0.0;
    t[0 * 6 + 0] = c;  t[3 * 6 + 3] = c;
    t[0 * 6 + 1] = s;  t[3 * 6 + 4] = s;
    t[1 * 6 + 0] = -s; t[4 * 6 + 3] = -s;
    t[1 * 6 + 1] = c;  t[4 * 6 + 4] = c;
    t[2 * 6 + 2] = 1.0;
    t[5 * 6 + 5] = 1.0;
}

// gonum Dgemm semantics for one output entry: sum over l ascending of a_l * b_l, skipping
// exact-zero a_l, starting from 0.0, multiply then add.
#define HD_ACC(acc, a, b)                                              \
    do {                                                               \
        const double a__ = (a);                                        \
        if (a__ != 0.0) acc = HD_ADD(acc, HD_MUL(a__, (b)));           \
    } while (0)

// model.go:619-620: kg = (T^T * k) * T.  T = blockdiag(R, R), R = [[c, s, 0], [-s, c, 0], [0, 0, 1]].
// Written with the block structure of T (only structurally non-zero terms of T are visited; a
// structural zero of T is an exact 0.0 multiplier that Dgemm skips or an exact-zero addend).
HD_FN void rotate_to_global(const double k[36], double c, double s, double kg[36])
{
    double a[36]; // a = T^T k : a[i][j] = sum_m T[m][i] k[m][j], m ascending
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
        const int o = 3 * blk;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            // row o+0 of T^T: T[o][o] = c, T[o+1][o] = -s
            HD_ACC(v0, c, k[(o + 0) * 6 + j]);
            HD_ACC(v0, -s, k[(o + 1) * 6 + j]);
            // row o+1: T[o][o+1] = s, T[o+1][o+1] = c
            HD_ACC(v1, s, k[(o + 0) * 6 + j]);
            HD_ACC(v1, c, k[(o + 1) * 6 + j]);
            // row o+2: T[o+2][o+2] = 1
            HD_ACC(v2, 1.0, k[(o + 2) * 6 + j]);
            a[(o + 0) * 6 + j] = v0;
            a[(o + 1) * 6 + j] = v1;
            a[(o + 2) * 6 + j] = v2;
        }
    }
    // kg = a T : kg[i][j] = sum_l a[i][l] T[l][j], l ascending, zero a[i][l] skipped
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int blk = 0; blk < 2; blk++) {
            const int o = 3 * blk;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            const double a0 = a[i * 6 + o], a1 = a[i * 6 + o + 1], a2 = a[i * 6 + o + 2];
            // column o+0 of T: T[o][o] = c, T[o+1][o] = -s
            HD_ACC(v0, a0, c);
            HD_ACC(v0, a1, -s);
            // column o+1: T[o][o+1] = s, T[o+1][o+1] = c
            HD_ACC(v1, a0, s);
            HD_ACC(v1, a1, c);
            // column o+2: T[o+2][o+2] = 1
            HD_ACC(v2, a2, 1.0);
            kg[i * 6 + o + 0] = v0;
            kg[i * 6 + o + 1] = v1;
            kg[i * 6 + o + 2] = v2;
        }
    }
}

// The same product for a beam without pins, written out: k has 18 non-zero entries built from five values, so the
// 36 results are eight distinct values and their negations.  Every sum below is the reference's own sequence of
// rounded operations (a zero term it adds or skips changes nothing, a sign flips exactly), hence bit-identical to
// rotate_to_global(stiff_beam2d(cf, 0), c, s).  d[8] = {x1, x2, x3, x4, r, w, EJL4, EJL}:
//   row 0: [ x1  x2 -r -x1 -x2 -r]   row 1: [ x3  x4  w -x3 -x4  w]   row 2: [-r  w EJL4  r -w EJL ]
//   row 3: -row 0                    row 4: -row 1                    row 5: [-r  w EJL   r -w EJL4]
HD_FN void rigid_global_values(const BeamCoef &cf, double c, double s, double d[8])
{
    const double p = HD_MUL(c, cf.EFL), t = HD_MUL(s, cf.EFL), q = HD_MUL(s, cf.EJL3), u = HD_MUL(c, cf.EJL3);
    d[0] = HD_ADD(HD_MUL(p, c), HD_MUL(q, s));
    d[1] = HD_SUB(HD_MUL(p, s), HD_MUL(q, c));
    d[2] = HD_SUB(HD_MUL(t, c), HD_MUL(u, s));
    d[3] = HD_ADD(HD_MUL(t, s), HD_MUL(u, c));
    d[4] = HD_MUL(s, cf.EJL2);
    d[5] = HD_MUL(c, cf.EJL2);
    d[6] = cf.EJL4;
    d[7] = cf.EJL;
}
// matrix positions (bit e = row * 6 + col) that hold +-d[i]
#define HD_RIGID_POS(i)                                                                                             \
    ((i) == 0 ? 0x240009ull : (i) == 1 ? 0x480012ull : (i) == 2 ? 0x9000240ull : (i) == 3 ? 0x12000480ull          \
     : (i) == 4 ? 0x240909024ull : (i) == 5 ? 0x4a4012900ull : (i) == 6 ? 0x800004000ull : 0x100020000ull)
// matrix positions that hold -d[i] rather than +d[i] (the minus signs of rigid_global_expand)
#define HD_RIGID_NEG 0x4670d163cull
HD_FN void rigid_global_expand(const double d[8], double kg[36])
{
    const double x1 = d[0], x2 = d[1], x3 = d[2], x4 = d[3], r = d[4], w = d[5], G = d[6], F = d[7];
    kg[0] = x1;   kg[1] = x2;   kg[2] = -r;  kg[3] = -x1;  kg[4] = -x2;  kg[5] = -r;
    kg[6] = x3;   kg[7] = x4;   kg[8] = w;   kg[9] = -x3;  kg[10] = -x4; kg[11] = w;
    kg[12] = -r;  kg[13] = w;   kg[14] = G;  kg[15] = r;   kg[16] = -w;  kg[17] = F;
    kg[18] = -x1; kg[19] = -x2; kg[20] = r;  kg[21] = x1;  kg[22] = x2;  kg[23] = r;
    kg[24] = -x3; kg[25] = -x4; kg[26] = -w; kg[27] = x3;  kg[28] = x4;  kg[29] = -w;
    kg[30] = -r;  kg[31] = w;   kg[32] = F;  kg[33] = r;   kg[34] = -w;  kg[35] = G;
}

} // namespace hdb
