// solve_ops.cuh — device helpers shared by the solve kernels (k_solve_pipe, k_solve_general, k_solve_bulk):
// explicit shared-window loads/stores, the beam end-force formulas (model.go:455-468) and the block substitution
// walked by one lane per load case (model.go:440).
#pragma once
#include "device_topo.hpp"
#include "element.cuh"
#include "hdb_internal.hpp"

namespace hdb {

struct BeamGeom {
    double c, s, length;
};

__device__ __forceinline__ BeamGeom beam_geom(const DevTopo &T, int b)
{ // feBeam2d.go:142-172 evaluated once per topology on the host with the routines of element.cuh
    BeamGeom g;
    g.c = __ldg(T.geom + 3 * b);
    g.s = __ldg(T.geom + 3 * b + 1);
    g.length = __ldg(T.geom + 3 * b + 2);
    return g;
}

// geometry of beam b of a model: the batch's own table (k_geometry) when coordinates vary per model
__device__ __forceinline__ BeamGeom beam_geom(const DevTopo &T, const StaticLaunch &L, int model, int b)
{
    if (!L.geom) return beam_geom(T, b);
    const double *q = L.geom + ((size_t)model * T.nB + b) * 3;
    BeamGeom g;
    g.c = q[0];
    g.s = q[1];
    g.length = q[2];
    return g;
}


struct SolveCfg {
    int casesPerPass; // <= 32
    int nPass;        // passes over the load cases of a model (1 unless shared memory is short)
    int lvmDoubles;   // one factor/table buffer
    int xsDoubles;    // one xs buffer
    int vec2;         // load and output arrays are 16-byte aligned and dof is even: 16-byte global accesses
};

// Status of one pass over the load cases of a model (model.go:415-437): the reference stops at the FIRST failing case
// — assemblyNodeLoad rejects a non-finite load (model.go:692-696) before the ignore-load check of the same case
// (model.go:424-437) — and the cases before it keep their results.  nf / ig = first case of this pass with a non-finite
// load / with a load on a DOF without stiffness (NO_FAIL_CASE = none); passes arrive in ascending case order, so a
// failure recorded by an earlier pass stands.  A model that failed in assembly or factorization keeps that status with
// failCase = -1.
constexpr unsigned NO_FAIL_CASE = 0x7fffffffu;
__device__ __forceinline__ void merge_load_status(const StaticLaunch &L, int model, int pass, int sOld, unsigned nf, unsigned ig)
{
    if (sOld == HDB_STATUS_OK) {
        int s = HDB_STATUS_OK, fc = -1;
        if (nf != NO_FAIL_CASE && nf <= ig) { s = HDB_STATUS_NONFINITE; fc = (int)nf; }
        else if (ig != NO_FAIL_CASE) { s = HDB_STATUS_IGNORE_LOAD; fc = (int)ig; }
        if (s != HDB_STATUS_OK) L.status[model] = s;
        if (L.failCase && (pass == 0 || fc >= 0)) L.failCase[model] = fc;
    } else if (pass == 0 && L.failCase) L.failCase[model] = -1;
}

// one block of xs = [cases][3] doubles, padded to an odd count: threads that read the same component of
// consecutive blocks (reactions, beam forces) then fall into different banks
__host__ __device__ inline int xs_block_doubles(int cp) { return (3 * cp) | 1; }


__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds_f64o(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void lds_v2o(unsigned a, double &x, double &y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(a), "n"(OFF));
}
__device__ __forceinline__ void sts_f64(unsigned a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

__device__ __forceinline__ void beam_forces_rigid(const BeamCoef &cf, const double gc, const double gs, const double *u,
                                                  double *o)
{ // rigid ends: the matrix has five distinct values (feBeam2d.go:28-56)
    const double z0 = fma(gs, u[1], gc * u[0]), z1 = fma(gc, u[1], -gs * u[0]), z2 = u[2];
    const double z3 = fma(gs, u[4], gc * u[3]), z4 = fma(gc, u[4], -gs * u[3]), z5 = u[5];
    const double ax = cf.EFL * (z0 - z3);
    const double sh = fma(cf.EJL2, z2 + z5, cf.EJL3 * (z1 - z4));
    const double m0 = fma(cf.EJL, z5, fma(cf.EJL4, z2, cf.EJL2 * (z1 - z4)));
    const double m1 = fma(cf.EJL4, z5, fma(cf.EJL, z2, cf.EJL2 * (z1 - z4)));
    // 0 - x, not -x: the reference accumulates every force from +0 (gonum Dgemv), so it never prints a negative zero
    o[0] = ax; o[1] = sh; o[2] = m0; o[3] = 0.0 - ax; o[4] = 0.0 - sh; o[5] = m1;
}

static __device__ __noinline__ void beam_forces_pinned(const BeamCoef cf, const uint32_t pm, const double gc, const double gs, // cf by value: a reference would pin the caller's copy to local memory
                                                const double *u, double *o)
{ // released ends: the condensed matrix is dense
    double k[36], z[6];
    stiff_beam2d(cf, pm, k);
    z[0] = fma(gs, u[1], gc * u[0]);
    z[1] = fma(gc, u[1], -gs * u[0]);
    z[2] = u[2];
    z[3] = fma(gs, u[4], gc * u[3]);
    z[4] = fma(gc, u[4], -gs * u[3]);
    z[5] = u[5];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double acc = k[i * 6] * z[0];
#pragma unroll
        for (int l = 1; l < 6; l++) acc = fma(k[i * 6 + l], z[l], acc);
        o[i] = acc + 0.0; // -0 -> +0, as a sum started from +0 gives (see beam_forces_rigid)
    }
}


__device__ __forceinline__ int4 lds_op(unsigned a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts_s32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void bar_stream(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); } // recover stage: streaming + helper warps
__device__ __forceinline__ void bar_load(int n) { asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); }   // load stage: streaming warps

// ---- the sequential part: one lane per load case walks the op runs (device_topo.hpp) -----------------
// A block column of the factor in shared memory: header [1/d0 1/d1 1/d2 v10 v20 v21], off blocks L_IJ
// row-major (+ 1 pad).  All table loads are warp-uniform 16-byte loads; the loads of op k+1 are issued
// before the result of op k is stored.
//
// CHAIN run, forward: z(tgt) = V_tgt (p(tgt) - L z(src)), src = target of the op before
__device__ __forceinline__ void chain_fwd(const unsigned pos, const int n, const unsigned lvA, const unsigned xA, const unsigned BS)
{
    int4 r = lds_op(pos), rN = lds_op(pos + 16u);
    double h0, h1, h2;
    {
        const unsigned sa = xA + (unsigned)((r.z >> OP_SRC_SHIFT) & OP_BLK_MASK) * BS;
        h0 = lds_f64o<0>(sa); h1 = lds_f64o<8>(sa); h2 = lds_f64o<16>(sa);
    }
    double m00, m01, m02, m10, m11, m12, m20, m21, m22, i2, v10, v20, v21, q0, q1, q2;
    unsigned ta;
#define CHAIN_LOAD(REC)                                                  \
    do {                                                                 \
        const unsigned a_ = lvA + (unsigned)(REC).x * 8u;                \
        const unsigned h_ = lvA + (unsigned)(REC).y * 8u;                \
        ta = xA + (unsigned)((REC).z & OP_BLK_MASK) * BS;                \
        lds_v2o<0>(a_, m00, m01);                                        \
        lds_v2o<16>(a_, m02, m10);                                       \
        lds_v2o<32>(a_, m11, m12);                                       \
        lds_v2o<48>(a_, m20, m21);                                       \
        m22 = lds_f64o<64>(a_);                                          \
        lds_v2o<16>(h_, i2, v10);                                        \
        lds_v2o<32>(h_, v20, v21);                                       \
        q0 = lds_f64o<0>(ta); q1 = lds_f64o<8>(ta); q2 = lds_f64o<16>(ta); \
    } while (0)
    CHAIN_LOAD(r);
    for (int k = 0; k < n; k++) {
        const int4 rNN = lds_op(pos + 16u * (unsigned)(k + 2)); // may run past the run: never used then
        const double a0 = fma(-m02, h2, fma(-m01, h1, fma(-m00, h0, q0)));
        const double a1 = fma(-m12, h2, fma(-m11, h1, fma(-m10, h0, q1)));
        const double a2 = fma(-m22, h2, fma(-m21, h1, fma(-m20, h0, q2)));
        const double z1 = fma(v10, a0, a1), z2 = fma(v21, a1, fma(v20, a0, a2));
        const unsigned tc = ta;
        CHAIN_LOAD(rN); // past the end of the run this reads a run header or the terminator: aligned, in range, unused
        sts_f64(tc, a0); sts_f64(tc + 8u, z1); sts_f64(tc + 16u, z2);
        h0 = a0; h1 = z1; h2 = z2;
        rN = rNN;
    }
    (void)i2;
#undef CHAIN_LOAD
}

// CHAIN run, backward: x(tgt) = V_tgt^T (D^-1 z(tgt) - L^T x(src))
__device__ __forceinline__ void chain_bwd(const unsigned pos, const int n, const unsigned lvA, const unsigned xA, const unsigned BS)
{
    int4 r = lds_op(pos), rN = lds_op(pos + 16u);
    double h0, h1, h2;
    {
        const unsigned sa = xA + (unsigned)((r.z >> OP_SRC_SHIFT) & OP_BLK_MASK) * BS;
        h0 = lds_f64o<0>(sa); h1 = lds_f64o<8>(sa); h2 = lds_f64o<16>(sa);
    }
    double m00, m01, m02, m10, m11, m12, m20, m21, m22, i0, i1, i2, v10, v20, v21, z0, z1, z2;
    unsigned ta;
#define CHAIN_LOAD(REC)                                                  \
    do {                                                                 \
        const unsigned a_ = lvA + (unsigned)(REC).x * 8u;                \
        const unsigned h_ = lvA + (unsigned)(REC).y * 8u;                \
        ta = xA + (unsigned)((REC).z & OP_BLK_MASK) * BS;                \
        lds_v2o<0>(a_, m00, m01);                                        \
        lds_v2o<16>(a_, m02, m10);                                       \
        lds_v2o<32>(a_, m11, m12);                                       \
        lds_v2o<48>(a_, m20, m21);                                       \
        m22 = lds_f64o<64>(a_);                                          \
        lds_v2o<0>(h_, i0, i1);                                          \
        lds_v2o<16>(h_, i2, v10);                                        \
        lds_v2o<32>(h_, v20, v21);                                       \
        z0 = lds_f64o<0>(ta); z1 = lds_f64o<8>(ta); z2 = lds_f64o<16>(ta); \
    } while (0)
    CHAIN_LOAD(r);
    for (int k = 0; k < n; k++) {
        const int4 rNN = lds_op(pos + 16u * (unsigned)(k + 2));
        // L^T: column j of the off block against x(src)
        const double a0 = fma(-m20, h2, fma(-m10, h1, fma(-m00, h0, z0 * i0)));
        const double a1 = fma(-m21, h2, fma(-m11, h1, fma(-m01, h0, z1 * i1)));
        const double a2 = fma(-m22, h2, fma(-m12, h1, fma(-m02, h0, z2 * i2)));
        const double x1 = fma(v21, a2, a1), x0 = fma(v20, a2, fma(v10, a1, a0));
        const unsigned tc = ta;
        CHAIN_LOAD(rN);
        sts_f64(tc, x0); sts_f64(tc + 8u, x1); sts_f64(tc + 16u, a2);
        h0 = x0; h1 = x1; h2 = a2;
        rN = rNN;
    }
#undef CHAIN_LOAD
}

// GENERIC run: per-op flags (blocks with several off blocks, blocks without any).  Same software pipeline as the
// chain loops: the off block, the source x (unless it is the block just finished, which lives in registers) and,
// at the first op of a block, its right-hand side are fetched one op ahead.
template <bool FWD>
static __device__ __noinline__ void generic_run(const unsigned pos, const int n, const unsigned lvA, const unsigned xA, const unsigned BS)
{
    double h0 = 0, h1 = 0, h2 = 0, a0 = 0, a1 = 0, a2 = 0;
    double m00, m01, m02, m10, m11, m12, m20, m21, m22, z0, z1, z2, q0, q1, q2;
#define GEN_LOAD(REC)                                                                          \
    do {                                                                                       \
        const unsigned a_ = lvA + (unsigned)(REC).x * 8u;                                      \
        const int y_ = (REC).z;                                                                \
        lds_v2o<0>(a_, m00, m01);                                                              \
        lds_v2o<16>(a_, m02, m10);                                                             \
        lds_v2o<32>(a_, m11, m12);                                                             \
        lds_v2o<48>(a_, m20, m21);                                                             \
        m22 = lds_f64o<64>(a_);                                                                \
        const unsigned s_ = xA + (unsigned)((y_ >> OP_SRC_SHIFT) & OP_BLK_MASK) * BS;          \
        z0 = lds_f64o<0>(s_); z1 = lds_f64o<8>(s_); z2 = lds_f64o<16>(s_);                     \
        const unsigned t_ = xA + (unsigned)(y_ & OP_BLK_MASK) * BS;                            \
        q0 = lds_f64o<0>(t_); q1 = lds_f64o<8>(t_); q2 = lds_f64o<16>(t_);                     \
    } while (0)
    int4 r = lds_op(pos), rN = lds_op(pos + 16u);
    GEN_LOAD(r);
    for (int k = 0; k < n; k++) {
        const int4 rNN = lds_op(pos + 16u * (unsigned)(k + 2)); // may run past the run: never used then
        const int y = r.z;
        const unsigned h_ = lvA + (unsigned)r.y * 8u;
        const unsigned ta = xA + (unsigned)(y & OP_BLK_MASK) * BS;
        if (y & OPF_FIRST) {
            a0 = q0; a1 = q1; a2 = q2;
            if (!FWD) { a0 *= lds_f64o<0>(h_); a1 *= lds_f64o<8>(h_); a2 *= lds_f64o<16>(h_); }
        }
        if (!(y & OPF_NOPULL)) {
            if (y & OPF_PREV) { z0 = h0; z1 = h1; z2 = h2; }
            if (FWD) {
                a0 = fma(-m02, z2, fma(-m01, z1, fma(-m00, z0, a0)));
                a1 = fma(-m12, z2, fma(-m11, z1, fma(-m10, z0, a1)));
                a2 = fma(-m22, z2, fma(-m21, z1, fma(-m20, z0, a2)));
            } else {
                a0 = fma(-m20, z2, fma(-m10, z1, fma(-m00, z0, a0)));
                a1 = fma(-m21, z2, fma(-m11, z1, fma(-m01, z0, a1)));
                a2 = fma(-m22, z2, fma(-m12, z1, fma(-m02, z0, a2)));
            }
        }
        double v10 = 0, v20 = 0, v21 = 0;
        if (y & OPF_LAST) { v10 = lds_f64o<24>(h_); v20 = lds_f64o<32>(h_); v21 = lds_f64o<40>(h_); }
        // the next op's values; its source or right-hand side may be the block this op is about to store, in
        // which case PREV is set (source) or it is a different block (right-hand side): both safe to read now
        GEN_LOAD(rN);
        if (y & OPF_LAST) {
            if (FWD) { h0 = a0; h1 = fma(v10, a0, a1); h2 = fma(v21, a1, fma(v20, a0, a2)); }
            else { h2 = a2; h1 = fma(v21, a2, a1); h0 = fma(v20, a2, fma(v10, a1, a0)); }
            sts_f64(ta, h0); sts_f64(ta + 8u, h1); sts_f64(ta + 16u, h2);
        }
        r = rN;
        rN = rNN;
    }
#undef GEN_LOAD
}

template <bool FWD>
__device__ __forceinline__ void walk_ops(unsigned pos, const unsigned lvA, const unsigned xA, const unsigned BS)
{
    for (;;) {
        const int4 hdr = lds_op(pos);
        const int n = hdr.x >> 1;
        if (n == 0) break;
        pos += 16u;
        if (hdr.z == RUN_CHAIN) {
            if (FWD) chain_fwd(pos, n, lvA, xA, BS);
            else chain_bwd(pos, n, lvA, xA, BS);
        } else generic_run<FWD>(pos, n, lvA, xA, BS);
        pos += 16u * (unsigned)n;
    }
}

} // namespace hdb
