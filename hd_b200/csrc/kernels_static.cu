// kernels_static.cu — sm_100a FP64 kernels of the batched LinearStatic path.
//
//   k_assemble       one CTA per model: per-beam 6x6 stiffness + transform + (T^T k) T in shared
//                    memory (feBeam2d.go:22-172, model.go:619-620), then a deterministic gather —
//                    one thread per matrix slot sums its element contributions in beam order (the
//                    duplicate-summation order of model.go:652), no atomics — into the interleaved
//                    factor workspace and the reaction block.
//   k_factor         one THREAD per model (SIMT over models, workspace interleaved so every access is
//                    coalesced): sparse LDL^T along the precomputed schedule; replaces
//                    sparse.LU.Factorize (model.go:326-331).
//   k_solve_recover  one CTA per model: load cases -> shared memory, forward / diagonal / backward
//                    substitution with one lane per load case (model.go:440), then displacement
//                    unpack (model.go:446-451), beam end forces k*(T*d) (model.go:455-468) and
//                    reactions -p + K[r,:]d (model.go:470-489), all written once, coalesced.
//
// Everything is HBM-bound integer-indexed FP64 work: no tensor-core shapes exist here (fronts are
// <= 6x6); see DESIGN.md for the roofline of each kernel.
#include <cstdio>

#include "device_topo.hpp"
#include "element.cuh"
#include "hdb_internal.hpp"

namespace hdb {

// ------------------------------------------------------------------------------------------------
// shared element evaluation
// ------------------------------------------------------------------------------------------------
struct BeamGeom {
    double c, s, length;
};

__device__ __forceinline__ BeamGeom beam_geom(const DevTopo &T, int b)
{
    const int n0 = T.beamN[2 * b], n1 = T.beamN[2 * b + 1];
    const double x0 = T.points[2 * n0], y0 = T.points[2 * n0 + 1];
    const double x1 = T.points[2 * n1], y1 = T.points[2 * n1 + 1];
    BeamGeom g;
    g.length = beam_length(x0, y0, x1, y1);
    beam_cs(x0, y0, x1, y1, g.length, g.c, g.s);
    return g;
}

__device__ __forceinline__ bool section_ok(double v)
{ // check.go:301-324: NaN, Inf, negative, zero are rejected
    return (v > 0.0) && !isinf(v);
}

// ------------------------------------------------------------------------------------------------
// k_assemble
// ------------------------------------------------------------------------------------------------
constexpr int ASM_THREADS = 256;

__global__ void __launch_bounds__(ASM_THREADS) k_assemble(const DevTopo T, const StaticLaunch L)
{
    extern __shared__ double smem_d[];
    __shared__ int s_flag;
    const int model = blockIdx.x;
    const int tid = threadIdx.x;
    double *ke = L.keScratch ? L.keScratch + (size_t)model * T.nB * KE_LD : smem_d;
    if (tid == 0) s_flag = 0;
    __syncthreads();

    // phase A: element matrices in global axes
    int flag = 0;
    const double *sec = L.sections + (size_t)model * T.nB * 3;
    for (int b = tid; b < T.nB; b += ASM_THREADS) {
        const double A = sec[3 * b], J = sec[3 * b + 1], E = sec[3 * b + 2];
        if (!(section_ok(A) && section_ok(J) && section_ok(E))) flag = max(flag, (int)HDB_STATUS_NONFINITE);
        const BeamGeom g = beam_geom(T, b);
        const BeamCoef cf = beam_coef(A, J, E, g.length);
        double k[36], kg[36];
        stiff_beam2d(cf, T.pinMask[b], k);
        rotate_to_global(k, g.c, g.s, kg);
        const uint8_t *dr = T.dropped + (size_t)b * 36;
#pragma unroll
        for (int e = 0; e < 36; e++) {
            const double v = kg[e];
            ke[b * KE_LD + e] = v;
            if (!isfinite(v)) flag = max(flag, (int)HDB_STATUS_NONFINITE);
            else if (dr[e] && v != 0.0) flag = max(flag, (int)HDB_STATUS_PATTERN);
        }
    }
    if (flag) atomicMax(&s_flag, flag == HDB_STATUS_NONFINITE ? 100 : flag); // NONFINITE wins
    __syncthreads();

    // phase B: gather; one thread per slot, contributions in beam order
    for (int s = tid; s < T.nSlots; s += ASM_THREADS) {
        double sum = T.slotInit[s] ? 1.0 : 0.0; // unit diagonal of a dummy unknown
        const int q1 = T.gptr[s + 1];
        for (int q = T.gptr[s]; q < q1; q++) sum = __dadd_rn(sum, ke[T.gsrc[q]]);
        L.W[(size_t)s * L.ld + model] = sum;
    }
    double *ksf = L.Ksf + (size_t)model * T.nnzR;
    for (int s = tid; s < T.nnzR; s += ASM_THREADS) {
        double sum = 0.0;
        const int q1 = T.rgptr[s + 1];
        for (int q = T.rgptr[s]; q < q1; q++) sum = __dadd_rn(sum, ke[T.rgsrc[q]]);
        ksf[s] = sum;
    }
    if (L.kval) {
        double *kv = L.kval + (size_t)model * T.nnzK;
        for (int s = tid; s < T.nnzK; s += ASM_THREADS) {
            double sum = 0.0;
            const int q1 = T.kptr[s + 1];
            for (int q = T.kptr[s]; q < q1; q++) sum = __dadd_rn(sum, ke[T.ksrc[q]]);
            kv[s] = sum;
        }
    }
    if (tid == 0) L.status[model] = (s_flag == 100) ? (int)HDB_STATUS_NONFINITE : s_flag;
}

bool assemble_fits_smem(const DevTopo &T, int smemOptin)
{
    return (size_t)T.nB * KE_LD * sizeof(double) + 1024 <= (size_t)smemOptin;
}

cudaError_t launch_assemble(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    const size_t smem = L.keScratch ? 0 : (size_t)T.nB * KE_LD * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_assemble<<<L.nModels, ASM_THREADS, smem, st>>>(T, L);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// element matrices only (tests)
// ------------------------------------------------------------------------------------------------
__global__ void k_elements(const DevTopo T, int nModels, const double *sections, double *klocal, double *tr,
                           double *kglobal)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nModels * T.nB) return;
    const int b = (int)(idx % T.nB);
    const double *sec = sections + idx * 3;
    const BeamGeom g = beam_geom(T, b);
    const BeamCoef cf = beam_coef(sec[0], sec[1], sec[2], g.length);
    double k[36], t[36], kg[36];
    stiff_beam2d(cf, T.pinMask[b], k);
    trans_beam2d(g.c, g.s, t);
    rotate_to_global(k, g.c, g.s, kg);
    for (int e = 0; e < 36; e++) {
        if (klocal) klocal[idx * 36 + e] = k[e];
        if (tr) tr[idx * 36 + e] = t[e];
        if (kglobal) kglobal[idx * 36 + e] = kg[e];
    }
}

cudaError_t launch_elements(const DevTopo &T, int nModels, const double *sections, double *klocal, double *tr,
                            double *kglobal, cudaStream_t st)
{
    const long long n = (long long)nModels * T.nB;
    k_elements<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(T, nModels, sections, klocal, tr, kglobal);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_factor: block LDL^T, one THREAD per model (a warp = 32 models), 3x3 node blocks.
//
// The factor workspace W[slot][ld] is interleaved over models, so lane = model makes every global
// access a coalesced 256-byte row.  The slots a block column touches form a contiguous range that
// slides forward with the elimination (the symbolic phase bounds it: T.winNeed), so the working set
// lives in a circular shared-memory window win[WS][32]: slots stream in with cp.async two block
// columns ahead, every dependent read-modify-write of the pivot chain is a shared-memory access,
// and each finished block column streams out once.  WS == 0: the window does not fit, work in
// global memory directly.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double *smem, const double *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int WS>
__global__ void __launch_bounds__(32) k_factor(const DevTopo T, const StaticLaunch L)
{
    extern __shared__ __align__(16) double smem_d[];
    const int lane = threadIdx.x;
    const int model = blockIdx.x * 32 + lane; // ld is a multiple of 32: padding lanes work on scratch columns
    double *wg = L.W + model;
    const size_t ld = (size_t)L.ld;
    double *win = smem_d + lane;
#define WIN(s) (*(WS ? (win + (((s) & (WS ? WS - 1 : 0)) << 5)) : (wg + (size_t)(s) * ld)))
    const int nBlk = T.nBlk;
    int loaded = 0;
    if (WS) {
        for (int pf = 0; pf < FACTOR_PREFETCH; pf++) {
            const int upto = __ldg(T.needUpto + min(pf, nBlk - 1));
            for (; loaded < upto; loaded++) cp_async8(&WIN(loaded), wg + (size_t)loaded * ld);
            cp_async_commit();
        }
    }
    bool bad = false;
    for (int J = 0; J < nBlk; J++) {
        if (WS) {
            const int upto = __ldg(T.needUpto + min(J + FACTOR_PREFETCH, nBlk - 1));
            for (; loaded < upto; loaded++) cp_async8(&WIN(loaded), wg + (size_t)loaded * ld);
            cp_async_commit();
            cp_async_wait<FACTOR_PREFETCH>();
        }
        const int base = __ldg(T.colOff + J);
        // ---- scalar LDL^T of the 3x3 diagonal block ----
        double d0 = WIN(base + 0), d1 = WIN(base + 1), d2 = WIN(base + 2);
        const double a10 = WIN(base + 3), a20 = WIN(base + 4), a21 = WIN(base + 5);
        // the reference's LU gives up only when a column has no non-zero pivot candidate (cs_lu: a <= 0):
        // tiny or negative round-off pivots of pin mechanisms pass there and must pass here (TestSplit)
        if (d0 == 0.0 || !isfinite(d0)) bad = true;
        const double inv0 = __drcp_rn(d0);
        const double l10 = a10 * inv0, l20 = a20 * inv0;
        d1 = fma(-l10, a10, d1);
        const double u21 = fma(-l20, a10, a21);
        if (d1 == 0.0 || !isfinite(d1)) bad = true;
        const double inv1 = __drcp_rn(d1);
        const double l21 = u21 * inv1;
        d2 = fma(-l21, u21, fma(-l20, a20, d2));
        if (d2 == 0.0 || !isfinite(d2)) bad = true;
        const double inv2 = __drcp_rn(d2);
        WIN(base + 0) = inv0; WIN(base + 1) = inv1; WIN(base + 2) = inv2;
        WIN(base + 3) = l10;  WIN(base + 4) = l20;  WIN(base + 5) = l21;
        const int q0 = __ldg(T.BLp + J), nb = __ldg(T.BLp + J + 1) - q0;
        // ---- off blocks: rows of L = A (L_JJ D_J)^-T ; stored scaled, u = l * d recovered when needed ----
        for (int a = 0; a < nb; a++) {
            const int sa = base + BLK_HDR + BLK_STRIDE * a;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double u0 = WIN(sa + 3 * r);
                const double u1 = fma(-l10, u0, WIN(sa + 3 * r + 1));
                const double u2 = fma(-l21, u1, fma(-l20, u0, WIN(sa + 3 * r + 2)));
                WIN(sa + 3 * r) = u0 * inv0;
                WIN(sa + 3 * r + 1) = u1 * inv1;
                WIN(sa + 3 * r + 2) = u2 * inv2;
            }
        }
        // ---- Schur updates of every pair of off blocks ----
        const int32_t *tg = T.pairTgt + __ldg(T.pairPtr + J);
        for (int a = 0; a < nb; a++) {
            const int sa = base + BLK_HDR + BLK_STRIDE * a;
            double la[9];
#pragma unroll
            for (int e = 0; e < 9; e++) la[e] = WIN(sa + e);
            for (int b = 0; b <= a; b++) {
                const int sb = base + BLK_HDR + BLK_STRIDE * b;
                double ub[9];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    ub[3 * r] = WIN(sb + 3 * r) * d0;
                    ub[3 * r + 1] = WIN(sb + 3 * r + 1) * d1;
                    ub[3 * r + 2] = WIN(sb + 3 * r + 2) * d2;
                }
                const int t = __ldg(tg++);
                if (a == b) { // diagonal block of row block I_a: header [d0 d1 d2 | (1,0) (2,0) (2,1)]
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        WIN(t + r) = fma(-la[3 * r + 2], ub[3 * r + 2], fma(-la[3 * r + 1], ub[3 * r + 1], fma(-la[3 * r], ub[3 * r], WIN(t + r))));
                    WIN(t + 3) = fma(-la[5], ub[2], fma(-la[4], ub[1], fma(-la[3], ub[0], WIN(t + 3))));
                    WIN(t + 4) = fma(-la[8], ub[2], fma(-la[7], ub[1], fma(-la[6], ub[0], WIN(t + 4))));
                    WIN(t + 5) = fma(-la[8], ub[5], fma(-la[7], ub[4], fma(-la[6], ub[3], WIN(t + 5))));
                } else { // block (I_a, I_b): rows from a, columns from b
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int rp = 0; rp < 3; rp++)
                            WIN(t + 3 * r + rp) = fma(-la[3 * r + 2], ub[3 * rp + 2],
                                                      fma(-la[3 * r + 1], ub[3 * rp + 1], fma(-la[3 * r], ub[3 * rp], WIN(t + 3 * r + rp))));
                }
            }
        }
        // ---- block column J is final: stream it out ----
        if (WS) {
            const int end = __ldg(T.colOff + J + 1);
            for (int s2 = base; s2 < end; s2++) wg[(size_t)s2 * ld] = WIN(s2);
        }
    }
#undef WIN
    if (bad && model < L.nModels && L.status[model] == HDB_STATUS_OK) L.status[model] = HDB_STATUS_SINGULAR;
}

cudaError_t launch_factor(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    const int grid = (L.nModels + 31) / 32;
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
#define LAUNCH_WIN(WS)                                                                                             \
    do {                                                                                                           \
        const size_t smem = (size_t)(WS) * 32 * sizeof(double);                                                    \
        if (smem > 48 * 1024) {                                                                                    \
            cudaError_t e = cudaFuncSetAttribute(k_factor<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                        \
        }                                                                                                          \
        k_factor<WS><<<grid, 32, smem, st>>>(T, L);                                                                \
    } while (0)
    if (T.winNeed <= 64) LAUNCH_WIN(64);
    else if (T.winNeed <= 128) LAUNCH_WIN(128);
    else if (T.winNeed <= 256) LAUNCH_WIN(256);
    else if (T.winNeed <= 512 && (size_t)512 * 32 * 8 + 1024 <= (size_t)optin) LAUNCH_WIN(512);
    else LAUNCH_WIN(0);
#undef LAUNCH_WIN
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_solve_recover
// ------------------------------------------------------------------------------------------------
constexpr int SOL_MAX_THREADS = 256;
constexpr int LV_PAD = 16; // the substitution reads 16 doubles around any slot: pad both ends of the factor

struct SolveCfg {
    int casesPerPass; // <= 32
    int cs;           // row stride of xs in doubles (odd)
};

__host__ __device__ inline int odd_stride(int c) { return c | 1; }

static size_t solve_smem_bytes(const DevTopo &T, int cp)
{
    return ((size_t)(T.nSlots + 2 * LV_PAD) + (size_t)T.nPad * odd_stride(cp)) * sizeof(double);
}

bool solve_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes)
{
    int cp = nLC < 32 ? nLC : 32;
    if (cp < 1) cp = 1;
    for (; cp >= 1; cp--) {
        const size_t bytes = solve_smem_bytes(T, cp);
        if (bytes + 2048 <= (size_t)smemOptin) {
            casesPerPass = cp;
            smemBytes = bytes;
            return true;
        }
    }
    return false;
}

// Block substitution, one lane per load case.  The warp walks an op list (device_topo.hpp): OP_PULL
// accumulates s += L_IJ z_J for one off block, OP_DIAG finishes a block with its diagonal part, OP_FUSED
// does both for the common banded case (single pull from the block finished just before, whose values
// are still in registers).  The 16 doubles around rec.x are laid out per pass so that every register
// index is static:
//   forward : [off block 9 + pad | 1/d0 1/d1 1/d2 l10 l20 l21]   (OP_DIAG points 10 before its header)
//   backward: [1/d0 1/d1 1/d2 l10 l20 l21 | off block 9 + pad]   (OP_PULL points 6 before its block)
// Software pipeline: the loads of op k+1 are issued in the middle of op k, right after its off-block
// values have been consumed, so their shared-memory latency hides behind the diagonal solve chain.
template <bool FWD>
struct OpVals {
    double L00, L01, L02, L10, L11, L12, L20, L21, L22;
    double i0, i1, i2; // 1/d (backward only)
    double l10, l20, l21;
    double p0, p1, p2; // x rows of rec's block: right-hand side of its own block, or a pulled source
};

template <bool FWD>
__device__ __forceinline__ void op_load(const int2 rec, const double *Lv, const double *xs, const int CS, const int c,
                                        OpVals<FWD> &o)
{
    const double *base = Lv + rec.x;
    const double2 *src = reinterpret_cast<const double2 *>(base);
    double2 t;
    if (FWD) {
        t = src[0]; o.L00 = t.x; o.L01 = t.y;
        t = src[1]; o.L02 = t.x; o.L10 = t.y;
        t = src[2]; o.L11 = t.x; o.L12 = t.y;
        t = src[3]; o.L20 = t.x; o.L21 = t.y;
        o.L22 = base[8];
        o.l10 = base[13];
        t = src[7]; o.l20 = t.x; o.l21 = t.y;
        o.i0 = o.i1 = o.i2 = 0.0;
    } else {
        t = src[0]; o.i0 = t.x; o.i1 = t.y;
        t = src[1]; o.i2 = t.x; o.l10 = t.y;
        t = src[2]; o.l20 = t.x; o.l21 = t.y;
        t = src[3]; o.L00 = t.x; o.L01 = t.y;
        t = src[4]; o.L02 = t.x; o.L10 = t.y;
        t = src[5]; o.L11 = t.x; o.L12 = t.y;
        t = src[6]; o.L20 = t.x; o.L21 = t.y;
        o.L22 = base[14];
    }
    const double *x = xs + (size_t)(3 * (rec.y & OP_BLK_MASK)) * CS + c;
    o.p0 = x[0];
    o.p1 = x[CS];
    o.p2 = x[2 * CS];
}

template <bool FWD>
__device__ __noinline__ void substitute(const int2 *__restrict__ ops, const int nOps, const double *Lv, double *xs,
                                        const int CS, const int c)
{
    if (nOps <= 0) return;
    double h0 = 0, h1 = 0, h2 = 0; // the block finished last
    double s0 = 0, s1 = 0, s2 = 0; // pulled sum of the block in progress
    int2 rec = __ldg(ops), recN = __ldg(ops + min(1, nOps - 1));
    OpVals<FWD> v;
    op_load<FWD>(rec, Lv, xs, CS, c, v);
    for (int k = 0; k < nOps; k++) {
        const int2 recNN = __ldg(ops + min(k + 2, nOps - 1));
        const int kind = (unsigned)rec.y >> OP_KIND_SHIFT;
        // ---- stage 1: off block ----
        double t0, t1, t2; // right-hand side of the diagonal stage
        if (kind == OP_FUSED) { // source = block finished just before (registers), fresh accumulator
            if (FWD) {
                t0 = fma(-v.L02, h2, fma(-v.L01, h1, fma(-v.L00, h0, v.p0)));
                t1 = fma(-v.L12, h2, fma(-v.L11, h1, fma(-v.L10, h0, v.p1)));
                t2 = fma(-v.L22, h2, fma(-v.L21, h1, fma(-v.L20, h0, v.p2)));
            } else {
                t0 = fma(-v.L20, h2, fma(-v.L10, h1, fma(-v.L00, h0, v.p0 * v.i0)));
                t1 = fma(-v.L21, h2, fma(-v.L11, h1, fma(-v.L01, h0, v.p1 * v.i1)));
                t2 = fma(-v.L22, h2, fma(-v.L12, h1, fma(-v.L02, h0, v.p2 * v.i2)));
            }
        } else if (kind == OP_PULL) {
            if (rec.y & OPF_FIRST) s0 = s1 = s2 = 0.0;
            double z0 = h0, z1 = h1, z2 = h2;
            if (!(rec.y & OPF_PREV)) { z0 = v.p0; z1 = v.p1; z2 = v.p2; }
            if (FWD) {
                s0 = fma(v.L02, z2, fma(v.L01, z1, fma(v.L00, z0, s0)));
                s1 = fma(v.L12, z2, fma(v.L11, z1, fma(v.L10, z0, s1)));
                s2 = fma(v.L22, z2, fma(v.L21, z1, fma(v.L20, z0, s2)));
            } else {
                s0 = fma(v.L20, z2, fma(v.L10, z1, fma(v.L00, z0, s0)));
                s1 = fma(v.L21, z2, fma(v.L11, z1, fma(v.L01, z0, s1)));
                s2 = fma(v.L22, z2, fma(v.L12, z1, fma(v.L02, z0, s2)));
            }
            t0 = t1 = t2 = 0.0;
        } else { // OP_DIAG
            if (rec.y & OPF_FIRST) s0 = s1 = s2 = 0.0;
            if (FWD) { t0 = v.p0 - s0; t1 = v.p1 - s1; t2 = v.p2 - s2; }
            else { t0 = fma(v.p0, v.i0, -s0); t1 = fma(v.p1, v.i1, -s1); t2 = fma(v.p2, v.i2, -s2); }
        }
        const double l10 = v.l10, l20 = v.l20, l21 = v.l21;
        const int blk = rec.y & OP_BLK_MASK;
        // ---- issue the loads of the next op: they complete behind the diagonal chain below ----
        if (k + 1 < nOps) op_load<FWD>(recN, Lv, xs, CS, c, v);
        // ---- stage 2: diagonal block ----
        if (kind != OP_PULL) {
            double r0, r1, r2;
            if (FWD) { // unit lower solve; the D scaling is applied by the backward pass
                r0 = t0;
                r1 = fma(-l10, r0, t1);
                r2 = fma(-l21, r1, fma(-l20, r0, t2));
            } else { // unit upper (transposed) solve
                r2 = t2;
                r1 = fma(-l21, r2, t1);
                r0 = fma(-l10, r1, fma(-l20, r2, t0));
            }
            double *x = xs + (size_t)(3 * blk) * CS + c;
            x[0] = r0;
            x[CS] = r1;
            x[2 * CS] = r2;
            h0 = r0; h1 = r1; h2 = r2;
        }
        rec = recN;
        recN = recNN;
    }
}

__global__ void __launch_bounds__(SOL_MAX_THREADS, 2) k_solve_recover(const DevTopo T, const StaticLaunch L, const SolveCfg cfg)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ int s_failCase, s_flag;
    const int model = blockIdx.x;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int dof = T.dof, nB = T.nB;
    const int CS = cfg.cs;
    double *Lv = smem_d + LV_PAD;                 // [-LV_PAD, nSlots + LV_PAD) factor of this model
    double *xs = smem_d + T.nSlots + 2 * LV_PAD;  // [nPad][CS]

    if (tid == 0) {
        s_failCase = 0x7fffffff;
        s_flag = 0;
    }
#define STAMP(i)                                                                              \
    do {                                                                                      \
        if (L.dbgClocks && tid == 0) L.dbgClocks[(size_t)model * 8 + (i)] = clock64();        \
    } while (0)
    STAMP(0);
    // phase 0: factor of this model -> shared memory (strided gather from the interleaved workspace);
    // eight independent loads in flight per thread
    {
        const double *w = L.W + model;
        const size_t ld = (size_t)L.ld;
        for (int s0 = 0; s0 < T.nSlots; s0 += NT * 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int s = s0 + u * NT + tid;
                v[u] = s < T.nSlots ? w[(size_t)s * ld] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int s = s0 + u * NT + tid;
                if (s < T.nSlots) Lv[s] = v[u];
            }
        }
        for (int s = tid; s < LV_PAD; s += NT) Lv[T.nSlots + s] = Lv[s - LV_PAD] = 0.0;
    }
    const double *ksf = L.Ksf + (size_t)model * T.nnzR;

    for (int c0 = 0; c0 < L.nLC; c0 += cfg.casesPerPass) {
        const int nc = min(cfg.casesPerPass, L.nLC - c0);
        const double *p = L.loads + ((size_t)model * L.nLC + c0) * dof;
        // phase 1: load cases -> xs (free rows), dummy rows = 0; model.go:692-696 and 424-437 checks.
        // A thread owns up to four DOFs (index maps read once) and streams the cases two at a time:
        // eight independent coalesced loads in flight.
        for (int g0 = 0; g0 < dof; g0 += NT * 4) {
            int xr[4];
            bool in[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int g = g0 + u * NT + tid;
                in[u] = g < dof;
                xr[u] = in[u] ? __ldg(T.xrow + g) : -1;
            }
            for (int c = 0; c < nc; c += 2) {
                double v[2][4];
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        v[j][u] = (in[u] && c + j < nc) ? p[(size_t)(c + j) * dof + g0 + u * NT + tid] : 0.0;
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (!(in[u] && c + j < nc)) continue;
                        const double val = v[j][u];
                        if (!isfinite(val)) atomicMax(&s_flag, 1);
                        if (xr[u] >= 0) xs[xr[u] * CS + c + j] = val;
                        else if (xr[u] == -2 && val != 0.0) atomicMin(&s_failCase, c0 + c + j);
                    }
            }
        }
        for (int idx = tid; idx < T.nDummy * nc; idx += NT) {
            const int q = idx / nc, c = idx - q * nc;
            xs[__ldg(T.dummyPos + q) * CS + c] = 0.0;
        }
        __syncthreads();
        STAMP(1);

        // phase 2: L z = p, then L^T x = D^-1 z; one lane per load case walks the block op lists
        if (tid < nc) {
            substitute<true>(T.fwdOps, T.nFwdOps, Lv, xs, CS, tid);
            STAMP(2);
            substitute<false>(T.bwdOps, T.nBwdOps, Lv, xs, CS, tid);
        }
        STAMP(3);
        __syncthreads();
        STAMP(4);

        // phase 3a+b: displacements (model.go:446-451) and reactions (model.go:470-489: columns ascending,
        // eliminated columns carry d = 0).  Same DOF ownership as phase 1; cases in groups of four.
        if (L.disp || L.react) {
            double *od = L.disp ? L.disp + ((size_t)model * L.nLC + c0) * dof : nullptr;
            double *orr = L.react ? L.react + ((size_t)model * L.nLC + c0) * dof : nullptr;
            for (int g0 = 0; g0 < dof; g0 += NT * 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int g = g0 + u * NT + tid;
                    if (g >= dof) continue;
                    const int xr = __ldg(T.xrow + g);
                    const int si = orr ? __ldg(T.supIndex + g) : -1;
                    int q0 = 0, q1 = 0;
                    if (si >= 0) {
                        q0 = __ldg(T.rptr + si);
                        q1 = __ldg(T.rptr + si + 1);
                    }
                    for (int c = 0; c < nc; c += 4) {
                        double acc[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[j] = (si >= 0 && c + j < nc) ? -p[(size_t)(c + j) * dof + g] : 0.0;
                        if (od) {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (c + j < nc) od[(size_t)(c + j) * dof + g] = xr >= 0 ? xs[xr * CS + c + j] : 0.0;
                        }
                        if (orr) {
                            for (int q = q0; q < q1; q++) {
                                const double kv = __ldg(ksf + q);
                                const double *xc = xs + __ldg(T.rcol + q) * CS + c;
#pragma unroll
                                for (int j = 0; j < 4; j++)
                                    if (c + j < nc) acc[j] = fma(kv, xc[j], acc[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (c + j < nc) orr[(size_t)(c + j) * dof + g] = acc[j];
                        }
                    }
                }
            }
        }
        // phase 3c: beam end forces s = k (T Zo), model.go:455-468.  One thread per beam for all cases; the
        // beams beyond the last full round of the block are spread case by case over the threads.
        if (L.forces) {
            double *out = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
            const bool vec16 = (reinterpret_cast<uintptr_t>(L.forces) & 15) == 0;
            const double *sec = L.sections + (size_t)model * nB * 3;
            const int nMain = (nB / NT) * NT, nRest = nB - nMain;
            const int nItems = nMain + nRest * nc;
            for (int it = tid; it < nItems; it += NT) {
                int b, cBeg, cEnd;
                if (it < nMain) { b = it; cBeg = 0; cEnd = nc; }
                else { b = nMain + (it - nMain) / nc; cBeg = (it - nMain) % nc; cEnd = cBeg + 1; }
                const BeamGeom g = beam_geom(T, b);
                const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
                const uint32_t pm = T.pinMask[b];
                int rows[6];
#pragma unroll
                for (int e = 0; e < 6; e++) rows[e] = __ldg(T.beamRows + b * 6 + e);
                if (pm == 0) { // rigid ends: the matrix has five distinct values (feBeam2d.go:28-56)
                    for (int c = cBeg; c < cEnd; c++) {
                        double u[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) u[e] = rows[e] >= 0 ? xs[rows[e] * CS + c] : 0.0;
                        const double z0 = fma(g.s, u[1], g.c * u[0]), z1 = fma(g.c, u[1], -g.s * u[0]), z2 = u[2];
                        const double z3 = fma(g.s, u[4], g.c * u[3]), z4 = fma(g.c, u[4], -g.s * u[3]), z5 = u[5];
                        const double ax = cf.EFL * (z0 - z3);
                        const double sh = fma(cf.EJL2, z2 + z5, cf.EJL3 * (z1 - z4));
                        const double m0 = fma(cf.EJL, z5, fma(cf.EJL4, z2, cf.EJL2 * (z1 - z4)));
                        const double m1 = fma(cf.EJL4, z5, fma(cf.EJL, z2, cf.EJL2 * (z1 - z4)));
                        double *of = out + ((size_t)c * nB + b) * 6;
                        if (vec16) { // 48-byte records: 16-byte aligned whenever the base pointer is
                            double2 *o = reinterpret_cast<double2 *>(of);
                            o[0] = make_double2(ax, sh);
                            o[1] = make_double2(m0, -ax);
                            o[2] = make_double2(-sh, m1);
                        } else {
                            of[0] = ax; of[1] = sh; of[2] = m0; of[3] = -ax; of[4] = -sh; of[5] = m1;
                        }
                    }
                } else { // released ends: the condensed matrix is dense
                    double k[36];
                    stiff_beam2d(cf, pm, k);
                    for (int c = cBeg; c < cEnd; c++) {
                        double u[6], z[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) u[e] = rows[e] >= 0 ? xs[rows[e] * CS + c] : 0.0;
                        z[0] = fma(g.s, u[1], g.c * u[0]);
                        z[1] = fma(g.c, u[1], -g.s * u[0]);
                        z[2] = u[2];
                        z[3] = fma(g.s, u[4], g.c * u[3]);
                        z[4] = fma(g.c, u[4], -g.s * u[3]);
                        z[5] = u[5];
                        double *o = out + ((size_t)c * nB + b) * 6;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            double acc = k[i * 6] * z[0];
#pragma unroll
                            for (int l = 1; l < 6; l++) acc = fma(k[i * 6 + l], z[l], acc);
                            o[i] = acc;
                        }
                    }
                }
            }
        }
        __syncthreads(); // xs is reused by the next pass
        STAMP(5);
    }
    if (tid == 0) {
        int st = L.status[model];
        if (st == HDB_STATUS_OK && s_flag) st = HDB_STATUS_NONFINITE;
        if (st == HDB_STATUS_OK && s_failCase != 0x7fffffff) st = HDB_STATUS_IGNORE_LOAD;
        L.status[model] = st;
        if (L.failCase) L.failCase[model] = (s_failCase != 0x7fffffff) ? s_failCase : -1;
    }
}

cudaError_t launch_solve_recover(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    SolveCfg cfg;
    size_t smem = 0;
    if (!solve_config(T, L.nLC, optin, cfg.casesPerPass, smem)) return cudaErrorInvalidConfiguration;
    cfg.cs = odd_stride(cfg.casesPerPass);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_solve_recover, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // one thread per beam in the force phase when the model is small enough, whole warps
    int threads = ((T.nB + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > SOL_MAX_THREADS) threads = 256;
    k_solve_recover<<<L.nModels, threads, smem, st>>>(T, L, cfg);
    return cudaGetLastError();
}

} // namespace hdb
