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
#include <algorithm>
#include <cstdio>
#include <cstdlib>

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
constexpr int ASM_MAX_THREADS = 512;

__global__ void __launch_bounds__(ASM_MAX_THREADS) k_assemble(const DevTopo T, const StaticLaunch L)
{
    const int ASM_THREADS = blockDim.x; // one thread per beam when nB <= 512: a single round of element work
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
    int threads = ((T.nB + 31) / 32) * 32;
    threads = threads < 128 ? 128 : (threads > ASM_MAX_THREADS ? ASM_MAX_THREADS : threads);
    k_assemble<<<L.nModels, threads, smem, st>>>(T, L);
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
    // one 16-byte descriptor per block column {first slot, off blocks, prefetch bound, first Schur target},
    // fetched one column ahead so that no index load sits on the pivot chain
    int4 dN = __ldg(T.colDesc);
    for (int J = 0; J < nBlk; J++) {
        const int4 d = dN;
        if (J + 1 < nBlk) dN = __ldg(T.colDesc + J + 1);
        if (WS) {
            for (; loaded < d.z; loaded++) cp_async8(&WIN(loaded), wg + (size_t)loaded * ld);
            cp_async_commit();
            cp_async_wait<FACTOR_PREFETCH>();
        }
        const int base = d.x;
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
        const int nb = d.y;
        if (nb == 1) {
            // ---- the banded case: one off block; its values and the Schur update stay in registers ----
            const int sa = base + BLK_HDR, t = d.w;
            double la[9], ua[9];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double u0 = WIN(sa + 3 * r);
                const double u1 = fma(-l10, u0, WIN(sa + 3 * r + 1));
                const double u2 = fma(-l21, u1, fma(-l20, u0, WIN(sa + 3 * r + 2)));
                ua[3 * r] = u0; ua[3 * r + 1] = u1; ua[3 * r + 2] = u2;
                la[3 * r] = u0 * inv0; la[3 * r + 1] = u1 * inv1; la[3 * r + 2] = u2 * inv2;
            }
#pragma unroll
            for (int e = 0; e < 9; e++) WIN(sa + e) = la[e];
#pragma unroll
            for (int r = 0; r < 3; r++)
                WIN(t + r) = fma(-la[3 * r + 2], ua[3 * r + 2], fma(-la[3 * r + 1], ua[3 * r + 1], fma(-la[3 * r], ua[3 * r], WIN(t + r))));
            WIN(t + 3) = fma(-la[5], ua[2], fma(-la[4], ua[1], fma(-la[3], ua[0], WIN(t + 3))));
            WIN(t + 4) = fma(-la[8], ua[2], fma(-la[7], ua[1], fma(-la[6], ua[0], WIN(t + 4))));
            WIN(t + 5) = fma(-la[8], ua[5], fma(-la[7], ua[4], fma(-la[6], ua[3], WIN(t + 5))));
        } else if (nb > 1) {
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
        }
        // ---- block column J is final: stream it out ----
        if (WS) {
            const int end = base + BLK_HDR + BLK_STRIDE * nb;
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
constexpr int LOAD_CASES = 8; // load cases streamed together by the load phase: 4 DOFs x 8 cases = 32 loads in flight per thread
constexpr int LV_PAD = 16; // the substitution reads 16 doubles around any slot: pad both ends of the factor

struct SolveCfg {
    int casesPerPass; // <= 32
    int cs;           // row stride of xs in doubles (odd)
    int nSM;
};

__host__ __device__ inline int odd_stride(int c) { return c | 1; }

static int solve_threads(const DevTopo &T)
{ // one thread per beam in the force phase when the model is small enough, whole warps
    int threads = ((T.nB + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > SOL_MAX_THREADS) threads = SOL_MAX_THREADS;
    return threads;
}

static size_t solve_smem_bytes(const DevTopo &T, int cp)
{
    return ((size_t)(T.nSlots + 2 * LV_PAD) + (size_t)T.nPad * odd_stride(cp) + (SOL_MAX_THREADS / 32) * 32 * 7 + (size_t)(T.nB % solve_threads(T)) * 8) *
           sizeof(double);
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
// A single warp is issue-bound here (about one instruction every two cycles), so the loop is written
// for instruction count: 32-bit shared-window addresses with immediate offsets, 16-byte value loads, a
// two-op unrolled body without register rotation.  Software pipeline: the loads of op k+1 are issued in
// the middle of op k, right after its off-block values have been consumed, so their latency hides
// behind the diagonal solve chain.
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

template <bool FWD>
__device__ __noinline__ void substitute(const int2 *__restrict__ ops, const int nOps, const unsigned lvA, const unsigned xsA,
                                        const int CS, const int c)
{
    if (nOps <= 0) return;
    const unsigned xA = xsA + (unsigned)c * 8u;
    const unsigned CS8 = (unsigned)CS * 8u, CS24 = 3u * CS8;
    double h0 = 0, h1 = 0, h2 = 0; // the block finished last
    double s0 = 0, s1 = 0, s2 = 0; // pulled sum of the block in progress
    // values of the current op that its first stage consumes; reloaded in the middle of the op
    double L00, L01, L02, L10, L11, L12, L20, L21, L22, p0, p1, p2, i0 = 0, i1 = 0, i2 = 0;
    unsigned xa;
#define SUB_LOAD_MID(REC)                                                          \
    do {                                                                           \
        const unsigned a_ = lvA + (unsigned)(REC).x * 8u;                          \
        if (FWD) {                                                                 \
            lds_v2o<0>(a_, L00, L01);                                              \
            lds_v2o<16>(a_, L02, L10);                                             \
            lds_v2o<32>(a_, L11, L12);                                             \
            lds_v2o<48>(a_, L20, L21);                                             \
            L22 = lds_f64o<64>(a_);                                                \
        } else {                                                                   \
            lds_v2o<0>(a_, i0, i1);                                                \
            i2 = lds_f64o<16>(a_);                                                 \
            lds_v2o<48>(a_, L00, L01);                                             \
            lds_v2o<64>(a_, L02, L10);                                             \
            lds_v2o<80>(a_, L11, L12);                                             \
            lds_v2o<96>(a_, L20, L21);                                             \
            L22 = lds_f64o<112>(a_);                                               \
        }                                                                          \
        xa = xA + (unsigned)((REC).y & OP_BLK_MASK) * CS24;                        \
        p0 = lds_f64(xa);                                                          \
        p1 = lds_f64(xa + CS8);                                                    \
        p2 = lds_f64(xa + 2u * CS8);                                               \
    } while (0)

    int2 rec = __ldg(ops), recN = __ldg(ops + min(1, nOps - 1));
    SUB_LOAD_MID(rec);
    for (int k = 0; k < nOps; k++) {
        const int2 recNN = __ldg(ops + min(k + 2, nOps - 1));
        const int kind = (unsigned)rec.y >> OP_KIND_SHIFT;
        // diagonal-stage values of this op: needed only after the first stage, so their latency is hidden
        double l10, l20, l21;
        {
            const unsigned a = lvA + (unsigned)rec.x * 8u;
            if (FWD) {
                l10 = lds_f64o<104>(a);
                lds_v2o<112>(a, l20, l21);
            } else {
                l10 = lds_f64o<24>(a);
                lds_v2o<32>(a, l20, l21);
            }
        }
        // ---- stage 1: off block ----
        double t0, t1, t2; // right-hand side of the diagonal stage
        if (kind == OP_FUSED) { // source = block finished just before (registers), fresh accumulator
            if (FWD) {
                t0 = fma(-L02, h2, fma(-L01, h1, fma(-L00, h0, p0)));
                t1 = fma(-L12, h2, fma(-L11, h1, fma(-L10, h0, p1)));
                t2 = fma(-L22, h2, fma(-L21, h1, fma(-L20, h0, p2)));
            } else {
                t0 = fma(-L20, h2, fma(-L10, h1, fma(-L00, h0, p0 * i0)));
                t1 = fma(-L21, h2, fma(-L11, h1, fma(-L01, h0, p1 * i1)));
                t2 = fma(-L22, h2, fma(-L12, h1, fma(-L02, h0, p2 * i2)));
            }
        } else if (kind == OP_PULL) {
            if (rec.y & OPF_FIRST) s0 = s1 = s2 = 0.0;
            double z0 = h0, z1 = h1, z2 = h2;
            if (!(rec.y & OPF_PREV)) { z0 = p0; z1 = p1; z2 = p2; }
            if (FWD) {
                s0 = fma(L02, z2, fma(L01, z1, fma(L00, z0, s0)));
                s1 = fma(L12, z2, fma(L11, z1, fma(L10, z0, s1)));
                s2 = fma(L22, z2, fma(L21, z1, fma(L20, z0, s2)));
            } else {
                s0 = fma(L20, z2, fma(L10, z1, fma(L00, z0, s0)));
                s1 = fma(L21, z2, fma(L11, z1, fma(L01, z0, s1)));
                s2 = fma(L22, z2, fma(L12, z1, fma(L02, z0, s2)));
            }
            t0 = t1 = t2 = 0.0;
        } else { // OP_DIAG
            if (rec.y & OPF_FIRST) s0 = s1 = s2 = 0.0;
            if (FWD) { t0 = p0 - s0; t1 = p1 - s1; t2 = p2 - s2; }
            else { t0 = fma(p0, i0, -s0); t1 = fma(p1, i1, -s1); t2 = fma(p2, i2, -s2); }
        }
        const unsigned xaCur = xa;
        // ---- the first-stage values are dead: reload them for the next op; the loads complete behind
        //      the diagonal chain below ----
        if (k + 1 < nOps) SUB_LOAD_MID(recN);
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
            sts_f64(xaCur, r0);
            sts_f64(xaCur + CS8, r1);
            sts_f64(xaCur + 2u * CS8, r2);
            h0 = r0; h1 = r1; h2 = r2;
        }
        rec = recN;
        recN = recNN;
    }
#undef SUB_LOAD_MID
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
    o[0] = ax; o[1] = sh; o[2] = m0; o[3] = -ax; o[4] = -sh; o[5] = m1;
}

__device__ __noinline__ void beam_forces_pinned(const BeamCoef &cf, const uint32_t pm, const double gc, const double gs,
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
        o[i] = acc;
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
    // Shared memory is addressed through 32-bit shared-window addresses and explicit ld/st.shared: generic
    // pointers into dynamic shared memory make the compiler re-derive the window base (S2R SR_CgaCtaId, an
    // XU-pipe instruction) inside every loop, which saturated the XU pipe in the first version of this kernel.
    const unsigned smA = (unsigned)__cvta_generic_to_shared(smem_d);
    const unsigned lvA = smA + LV_PAD * 8u;                       // [-LV_PAD, nSlots + LV_PAD) factor of this model
    const unsigned xsA = smA + (unsigned)(T.nSlots + 2 * LV_PAD) * 8u; // [nPad][CS]
    const unsigned CS8 = (unsigned)CS * 8u;
    const unsigned fstA = xsA + (unsigned)T.nPad * CS8;              // per-warp 32 x 7 staging tiles of the force phase

    if (tid == 0) {
        s_failCase = 0x7fffffff;
        s_flag = 0;
    }
#define STAMP(i)                                                                              \
    do {                                                                                      \
        if (L.dbgClocks && tid == stampTid) L.dbgClocks[(size_t)model * 32 + (i)] = clock64();  \
    } while (0)
    const int stampTid = (((int)blockIdx.x / (cfg.nSM > 0 ? cfg.nSM : 148)) & 3) * 32;
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
                if (s < T.nSlots) sts_f64(lvA + (unsigned)s * 8u, v[u]);
            }
        }
        for (int s = tid; s < LV_PAD; s += NT) {
            sts_f64(lvA + (unsigned)(T.nSlots + s) * 8u, 0.0);
            sts_f64(lvA - (unsigned)(LV_PAD - s) * 8u, 0.0);
        }
    }
    const double *ksf = L.Ksf + (size_t)model * T.nnzR;
    const int nSM = cfg.nSM > 0 ? cfg.nSM : 148;
    const int solver = ((blockIdx.x / nSM) & 3) * 32; // co-resident CTAs put their solver warp on different SMSPs
    // beams beyond the last full round of the block: their coefficients go to a small shared table once,
    // so that the force phase never recomputes an element inside its case loop
    const int nMain = (nB / NT) * NT, nRest = nB - nMain;
    const unsigned crA = fstA + (unsigned)(SOL_MAX_THREADS / 32) * (32u * 7u * 8u);
    if (L.forces && tid < nRest) {
        const int b = nMain + tid;
        const double *sec = L.sections + (size_t)model * nB * 3;
        const BeamGeom g = beam_geom(T, b);
        const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
        const unsigned a = crA + (unsigned)tid * 64u;
        sts_f64(a, cf.EFL); sts_f64(a + 8, cf.EJL3); sts_f64(a + 16, cf.EJL2); sts_f64(a + 24, cf.EJL);
        sts_f64(a + 32, cf.EJL4); sts_f64(a + 40, g.c); sts_f64(a + 48, g.s);
    }
    constexpr int LC = 4; // load cases handled together by one work item of the streaming phases
    bool badLoad = false;
    int failC = 0x7fffffff;

    for (int c0 = 0; c0 < L.nLC; c0 += cfg.casesPerPass) {
        const int nc = min(cfg.casesPerPass, L.nLC - c0);
        const double *p = L.loads + ((size_t)model * L.nLC + c0) * dof;
        double *orr = L.react ? L.react + ((size_t)model * L.nLC + c0) * dof : nullptr;
        const int nItems = ((nc + LC - 1) / LC) * dof; // item = (group of LC cases, DOF); DOF fastest: coalesced
        // phase 1: load cases -> xs (free rows), dummy rows = 0; model.go:692-696 and 424-437 checks.
        // Four items per thread are in flight (16 independent coalesced loads).  The reaction output starts
        // here as -p on supported DOFs (model.go:478) and 0 elsewhere; phase 3 adds K[r,:] d to it with one
        // fire-and-forget reduction per entry, so p is never read twice.
        for (int it0 = 0; it0 < nItems; it0 += NT * 4) {
            double v[4][LC];
            int gi[4], ci[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int it = it0 + u * NT + tid;
                const int cg = it / dof;
                gi[u] = it < nItems ? it - cg * dof : -1;
                ci[u] = cg * LC;
#pragma unroll
                for (int j = 0; j < LC; j++)
                    v[u][j] = (gi[u] >= 0 && ci[u] + j < nc) ? __ldcs(p + (size_t)(ci[u] + j) * dof + gi[u]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (gi[u] < 0) continue;
                const int g = gi[u], c = ci[u];
                const int xr = __ldg(T.xrow + g);
                const bool sup = orr && __ldg(T.supIndex + g) >= 0;
                const unsigned xb = xsA + (unsigned)(xr >= 0 ? xr : 0) * CS8 + (unsigned)c * 8u;
                double *po = orr ? orr + (size_t)c * dof + g : nullptr;
#pragma unroll
                for (int j = 0; j < LC; j++) {
                    if (c + j >= nc) break;
                    const double val = v[u][j];
                    badLoad |= !(fabs(val) <= 1.7976931348623157e308);
                    if (xr >= 0) sts_f64(xb + 8u * j, val);
                    if (po) po[(size_t)j * dof] = sup ? -val : 0.0;
                }
                if (xr == -2) {
#pragma unroll
                    for (int j = 0; j < LC; j++)
                        if (c + j < nc && v[u][j] != 0.0) failC = min(failC, c0 + c + j);
                }
            }
        }
        for (int idx = tid; idx < T.nDummy * nc; idx += NT) {
            const int q = idx / nc, c = idx - q * nc;
            sts_f64(xsA + (unsigned)__ldg(T.dummyPos + q) * CS8 + (unsigned)c * 8u, 0.0);
        }
        __syncthreads();
        STAMP(1);

        // phase 2: L z = p, then L^T x = D^-1 z; one lane per load case walks the block op lists
        if (tid >= solver && tid < solver + nc) {
            substitute<true>(T.fwdOps, T.nFwdOps, lvA, xsA, CS, tid - solver);
            STAMP(2);
            substitute<false>(T.bwdOps, T.nBwdOps, lvA, xsA, CS, tid - solver);
        }
        STAMP(3);
        __syncthreads();
        STAMP(4);

        // phase 3a+b: displacements (model.go:446-451) and reactions (model.go:470-489: columns ascending,
        // eliminated columns carry d = 0) added to the -p rows of phase 1 with one reduction per supported entry.
        if (L.disp || orr) {
            double *od0 = L.disp ? L.disp + ((size_t)model * L.nLC + c0) * dof : nullptr;
            for (int it = tid; it < nItems; it += NT) {
                const int cg = it / dof, g = it - cg * dof, c = cg * LC;
                const int xr = __ldg(T.xrow + g);
                if (od0) {
                    double *od = od0 + (size_t)c * dof + g;
                    const unsigned xb = xsA + (unsigned)(xr >= 0 ? xr : 0) * CS8 + (unsigned)c * 8u;
                    double d[LC];
#pragma unroll
                    for (int j = 0; j < LC; j++) d[j] = (xr >= 0 && c + j < nc) ? lds_f64(xb + 8u * j) : 0.0;
#pragma unroll
                    for (int j = 0; j < LC; j++)
                        if (c + j < nc) od[(size_t)j * dof] = d[j];
                }
                const int si = orr ? __ldg(T.supIndex + g) : -1;
                if (si >= 0) {
                    const int q0 = __ldg(T.rptr + si), q1 = __ldg(T.rptr + si + 1);
                    double acc[LC];
#pragma unroll
                    for (int j = 0; j < LC; j++) acc[j] = 0.0;
                    for (int q = q0; q < q1; q++) {
                        const double kv = __ldg(ksf + q);
                        const unsigned xa = xsA + (unsigned)__ldg(T.rcol + q) * CS8 + (unsigned)c * 8u;
#pragma unroll
                        for (int j = 0; j < LC; j++)
                            if (c + j < nc) acc[j] = fma(kv, lds_f64(xa + 8u * j), acc[j]);
                    }
                    double *pr = orr + (size_t)c * dof + g;
#pragma unroll
                    for (int j = 0; j < LC; j++)
                        if (c + j < nc) atomicAdd(pr + (size_t)j * dof, acc[j]); // RED: no return value
                }
            }
        }
        if (L.dbgClocks && (tid & 31) == 0 && (tid >> 5) < 8) L.dbgClocks[(size_t)model * 32 + 8 + (tid >> 5)] = clock64();
        // phase 3c: beam end forces s = k (T Zo), model.go:455-468.  One thread per beam keeps the element
        // coefficients in registers over all cases; the few beams beyond the last full round are spread case
        // by case over the threads and take their coefficients from the shared table.
        if (L.forces) {
            double *out = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
            const double *sec = L.sections + (size_t)model * nB * 3;
            for (int b = tid; b < nMain; b += NT) {
                const BeamGeom g = beam_geom(T, b);
                const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
                const uint32_t pm = T.pinMask[b];
                unsigned ra[6];
                bool rv[6];
#pragma unroll
                for (int e = 0; e < 6; e++) {
                    const int r = __ldg(T.beamRows + b * 6 + e);
                    rv[e] = r >= 0;
                    ra[e] = xsA + (unsigned)(r >= 0 ? r : 0) * CS8;
                }
                const unsigned wb = fstA + (unsigned)(tid >> 5) * (32u * 7u * 8u);
                const unsigned lane = tid & 31;
                for (int c = 0; c < nc; c++) {
                    double u[6], o[6];
#pragma unroll
                    for (int e = 0; e < 6; e++) u[e] = rv[e] ? lds_f64(ra[e] + (unsigned)c * 8u) : 0.0;
                    if (pm == 0) beam_forces_rigid(cf, g.c, g.s, u, o);
                    else { // keep u/o of the common path in registers: the out-of-line call gets its own copies
                        double uu[6], oo[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) uu[e] = u[e];
                        beam_forces_pinned(cf, pm, g.c, g.s, uu, oo);
#pragma unroll
                        for (int e = 0; e < 6; e++) o[e] = oo[e];
                    }
                    // a full warp holds 32 consecutive beams = 192 consecutive doubles of the output row:
                    // transpose through shared memory so that every store instruction writes one contiguous
                    // 256-byte run (full sectors) instead of 16-byte pieces 48 bytes apart
                    __syncwarp();
#pragma unroll
                    for (int e = 0; e < 6; e++) sts_f64(wb + (lane * 7u + e) * 8u, o[e]);
                    __syncwarp();
                    double *ow = out + ((size_t)c * nB + (b - (int)lane)) * 6;
#pragma unroll
                    for (int e = 0; e < 6; e++) {
                        const unsigned q = e * 32u + lane;
                        ow[q] = lds_f64(wb + ((q / 6u) * 7u + q % 6u) * 8u);
                    }
                }
            }
            for (int it = tid; it < nRest * nc; it += NT) {
                const int r = it / nc, c = it - r * nc, b = nMain + r;
                const unsigned a = crA + (unsigned)r * 64u;
                BeamCoef cf;
                cf.EFL = lds_f64(a); cf.EJL3 = lds_f64(a + 8); cf.EJL2 = lds_f64(a + 16); cf.EJL = lds_f64(a + 24);
                cf.EJL4 = lds_f64(a + 32);
                const double gc = lds_f64(a + 40), gs = lds_f64(a + 48);
                const uint32_t pm = T.pinMask[b];
                double u[6], o[6];
#pragma unroll
                for (int e = 0; e < 6; e++) {
                    const int row = __ldg(T.beamRows + b * 6 + e);
                    u[e] = row >= 0 ? lds_f64(xsA + (unsigned)row * CS8 + (unsigned)c * 8u) : 0.0;
                }
                if (pm == 0) beam_forces_rigid(cf, gc, gs, u, o);
                else {
                    double uu[6], oo[6];
#pragma unroll
                    for (int e = 0; e < 6; e++) uu[e] = u[e];
                    beam_forces_pinned(cf, pm, gc, gs, uu, oo);
#pragma unroll
                    for (int e = 0; e < 6; e++) o[e] = oo[e];
                }
                double *of = out + ((size_t)c * nB + b) * 6;
#pragma unroll
                for (int e = 0; e < 6; e++) of[e] = o[e];
            }
        }
        if (L.dbgClocks && (tid & 31) == 0 && (tid >> 5) < 8) L.dbgClocks[(size_t)model * 32 + 16 + (tid >> 5)] = clock64();
        __syncthreads(); // xs is reused by the next pass
        STAMP(5);
    }
    if (badLoad) atomicMax(&s_flag, 1);
    if (failC != 0x7fffffff) atomicMin(&s_failCase, failC);
    __syncthreads();
    if (tid == 0) {
        int st = L.status[model];
        if (st == HDB_STATUS_OK && s_flag) st = HDB_STATUS_NONFINITE;
        if (st == HDB_STATUS_OK && s_failCase != 0x7fffffff) st = HDB_STATUS_IGNORE_LOAD;
        L.status[model] = st;
        if (L.failCase) L.failCase[model] = (s_failCase != 0x7fffffff) ? s_failCase : -1;
    }
}

// ------------------------------------------------------------------------------------------------
// k_solve_recover_tma: same pipeline, rows moved by the TMA unit (cp.async.bulk + mbarrier).
//
// A model's load block [nLC][dof] and its outputs are long contiguous rows, so every row crosses the
// SM boundary as ONE bulk asynchronous copy: four load rows are always in flight into a shared-memory
// ring without costing a register or an issue slot per element, results are staged row by row in the
// same ring and leave with one bulk store each.  Preconditions (checked at launch, else the LDG/STG
// kernel above runs): dof even and 16-byte aligned arrays (bulk copies move multiples of 16 bytes),
// nB < 2 * blockDim (one beam per thread in registers, the rest through a small coefficient table).
// ------------------------------------------------------------------------------------------------
constexpr int TMA_DROWS = 4; // ring depth for rows of dof doubles (loads in, displacements out)
constexpr int TMA_FROWS = 2; // ring depth for rows of nB*6 doubles (beam forces out)

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TmaCfg {
    int casesPerPass, cs;
    int ringDoubles; // ring size in doubles: max(TMA_DROWS * dof, TMA_FROWS * nB * 6)
    int nSM;
};

__global__ void __launch_bounds__(SOL_MAX_THREADS, 2) k_solve_recover_tma(const DevTopo T, const StaticLaunch L, const TmaCfg cfg)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ int s_failCase, s_flag;
    __shared__ __align__(8) unsigned long long s_bar[TMA_DROWS];
    const int model = blockIdx.x;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int dof = T.dof, nB = T.nB;
    const int CS = cfg.cs;
    double *Lv = smem_d + LV_PAD;                // [-LV_PAD, nSlots + LV_PAD) factor of this model
    double *xs = smem_d + T.nSlots + 2 * LV_PAD; // [nPad][CS]
    double *ring = xs + (((size_t)T.nPad * CS + 1) & ~(size_t)1); // 16-byte aligned
    double *coefRest = ring + cfg.ringDoubles;   // [nRest][8] coefficients of the beams beyond blockDim
    const unsigned ringA = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned barA = (unsigned)__cvta_generic_to_shared(s_bar);
    const unsigned rowDB = (unsigned)dof * 8u, rowFB = (unsigned)nB * 48u;

    if (tid == 0) {
        s_failCase = 0x7fffffff;
        s_flag = 0;
        for (int i = 0; i < TMA_DROWS; i++) mbar_init(barA + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int stampTid = (((int)blockIdx.x / cfg.nSM) & 3) * 32;
    STAMP(0);
    // phase 0: factor of this model -> shared memory (strided gather from the interleaved workspace)
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
    // per-thread DOF ownership (up to four DOFs) and beam ownership (one beam in registers)
    int xr[4], si[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int g = tid + u * NT;
        xr[u] = g < dof ? __ldg(T.xrow + g) : -1;
        si[u] = g < dof ? __ldg(T.supIndex + g) : -1;
    }
    const double *sec = L.sections + (size_t)model * nB * 3;
    const bool haveBeam = tid < nB;
    BeamCoef cfB{};
    double gcB = 0, gsB = 0;
    uint32_t pmB = 0;
    int rowsB[6] = {-1, -1, -1, -1, -1, -1};
    if (haveBeam && L.forces) {
        const BeamGeom g = beam_geom(T, tid);
        cfB = beam_coef(sec[3 * tid], sec[3 * tid + 1], sec[3 * tid + 2], g.length);
        gcB = g.c;
        gsB = g.s;
        pmB = T.pinMask[tid];
#pragma unroll
        for (int e = 0; e < 6; e++) rowsB[e] = __ldg(T.beamRows + tid * 6 + e);
    }
    const int nRest = nB > NT ? nB - NT : 0;
    if (tid < nRest && L.forces) {
        const int b = NT + tid;
        const BeamGeom g = beam_geom(T, b);
        const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
        double *cr = coefRest + tid * 8;
        cr[0] = cf.EFL; cr[1] = cf.EJL3; cr[2] = cf.EJL2; cr[3] = cf.EJL; cr[4] = cf.EJL4; cr[5] = g.c; cr[6] = g.s;
    }
    __syncthreads(); // barriers initialised, factor and coefficient table in place
    const int solver = ((blockIdx.x / cfg.nSM) & 3) * 32; // co-resident CTAs put their solver warp on different SMSPs
    unsigned rowsDone = 0; // rows that went through the load ring so far (mbarrier phase bookkeeping)

    for (int c0 = 0; c0 < L.nLC; c0 += cfg.casesPerPass) {
        const int nc = min(cfg.casesPerPass, L.nLC - c0);
        const double *p = L.loads + ((size_t)model * L.nLC + c0) * dof;
        double *orr = L.react ? L.react + ((size_t)model * L.nLC + c0) * dof : nullptr;
        // ---- phase 1: load rows through the ring -> xs; reaction rows start as -p (model.go:478) ----
        if (tid == 0) {
            for (int c = 0; c < min(TMA_DROWS, nc); c++) {
                const unsigned s = (rowsDone + c) % TMA_DROWS;
                mbar_expect_tx(barA + 8u * s, rowDB);
                bulk_g2s(ringA + s * rowDB, p + (size_t)c * dof, rowDB, barA + 8u * s);
            }
        }
        for (int c = 0; c < nc; c++) {
            const unsigned rd = rowsDone + c, s = rd % TMA_DROWS;
            mbar_wait(barA + 8u * s, (rd / TMA_DROWS) & 1u);
            double *row = ring + (size_t)s * dof;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int g = tid + u * NT;
                if (g >= dof) continue;
                const double val = row[g];
                if (!isfinite(val)) atomicMax(&s_flag, 1); // model.go:692-696
                if (xr[u] >= 0) xs[xr[u] * CS + c] = val;
                else if (xr[u] == -2 && val != 0.0) atomicMin(&s_failCase, c0 + c); // model.go:424-437
                if (orr) row[g] = si[u] >= 0 ? -val : 0.0;
            }
            if (orr) fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (orr) {
                    bulk_s2g(orr + (size_t)c * dof, ringA + s * rowDB, rowDB);
                    bulk_commit();
                }
                if (c + TMA_DROWS < nc) { // refill this stage once the store has read it
                    if (orr) bulk_wait_read<0>();
                    mbar_expect_tx(barA + 8u * s, rowDB);
                    bulk_g2s(ringA + s * rowDB, p + (size_t)(c + TMA_DROWS) * dof, rowDB, barA + 8u * s);
                }
            }
        }
        rowsDone += nc;
        for (int idx = tid; idx < T.nDummy * nc; idx += NT) {
            const int q = idx / nc, c = idx - q * nc;
            xs[__ldg(T.dummyPos + q) * CS + c] = 0.0;
        }
        if (tid == 0 && orr) bulk_wait<0>(); // the -p rows are in memory before phase 3 adds to them
        __syncthreads();
        STAMP(1);

        // ---- phase 2: L z = p, then L^T x = D^-1 z; one lane per load case walks the block op lists ----
        if (tid >= solver && tid < solver + nc) {
            substitute<true>(T.fwdOps, T.nFwdOps, (unsigned)__cvta_generic_to_shared(Lv), (unsigned)__cvta_generic_to_shared(xs), CS, tid - solver);
            substitute<false>(T.bwdOps, T.nBwdOps, (unsigned)__cvta_generic_to_shared(Lv), (unsigned)__cvta_generic_to_shared(xs), CS, tid - solver);
        }
        STAMP(3);
        __syncthreads();
        STAMP(4);

        // ---- phase 3a: reactions, model.go:470-489: one fire-and-forget reduction per supported entry ----
        if (orr) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (si[u] < 0) continue;
                const int g = tid + u * NT;
                const int q0 = __ldg(T.rptr + si[u]), q1 = __ldg(T.rptr + si[u] + 1);
                for (int c = 0; c < nc; c += 4) {
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int q = q0; q < q1; q++) {
                        const double kv = __ldg(ksf + q);
                        const double *xc = xs + __ldg(T.rcol + q) * CS + c;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (c + j < nc) acc[j] = fma(kv, xc[j], acc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (c + j < nc) atomicAdd(orr + (size_t)(c + j) * dof + g, acc[j]); // RED: no return value
                }
            }
        }
        // ---- phase 3b: displacements (model.go:446-451), one bulk store per case row ----
        if (L.disp) {
            double *od = L.disp + ((size_t)model * L.nLC + c0) * dof;
            for (int c = 0; c < nc; c++) {
                const int s = c % TMA_DROWS;
                if (tid == 0 && c >= TMA_DROWS) bulk_wait_read<TMA_DROWS - 1>();
                __syncthreads();
                double *row = ring + (size_t)s * dof;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int g = tid + u * NT;
                    if (g < dof) row[g] = xr[u] >= 0 ? xs[xr[u] * CS + c] : 0.0;
                }
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    bulk_s2g(od + (size_t)c * dof, ringA + s * rowDB, rowDB);
                    bulk_commit();
                }
            }
            if (tid == 0) bulk_wait_read<0>(); // the ring changes row size below
        }
        // ---- phase 3c: beam end forces s = k (T Zo), model.go:455-468; one bulk store per case row ----
        if (L.forces) {
            double *of = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
            for (int c = 0; c < nc; c++) {
                const int s = c % TMA_FROWS;
                if (tid == 0 && c >= TMA_FROWS) bulk_wait_read<TMA_FROWS - 1>();
                __syncthreads();
                double *row = ring + (size_t)s * nB * 6;
                if (haveBeam) {
                    double u[6];
#pragma unroll
                    for (int e = 0; e < 6; e++) u[e] = rowsB[e] >= 0 ? xs[rowsB[e] * CS + c] : 0.0;
                    double o[6];
                    if (pmB == 0) beam_forces_rigid(cfB, gcB, gsB, u, o);
                    else {
                        double uu[6], oo[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) uu[e] = u[e];
                        beam_forces_pinned(cfB, pmB, gcB, gsB, uu, oo);
#pragma unroll
                        for (int e = 0; e < 6; e++) o[e] = oo[e];
                    }
#pragma unroll
                    for (int e = 0; e < 6; e++) row[tid * 6 + e] = o[e];
                }
                if (tid < nRest) {
                    const int b = NT + tid;
                    const double *cr = coefRest + tid * 8;
                    BeamCoef cf;
                    cf.EFL = cr[0]; cf.EJL3 = cr[1]; cf.EJL2 = cr[2]; cf.EJL = cr[3]; cf.EJL4 = cr[4];
                    double u[6];
#pragma unroll
                    for (int e = 0; e < 6; e++) {
                        const int r = __ldg(T.beamRows + b * 6 + e);
                        u[e] = r >= 0 ? xs[r * CS + c] : 0.0;
                    }
                    const uint32_t pm = T.pinMask[b];
                    double o[6];
                    if (pm == 0) beam_forces_rigid(cf, cr[5], cr[6], u, o);
                    else {
                        double uu[6], oo[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) uu[e] = u[e];
                        beam_forces_pinned(cf, pm, cr[5], cr[6], uu, oo);
#pragma unroll
                        for (int e = 0; e < 6; e++) o[e] = oo[e];
                    }
#pragma unroll
                    for (int e = 0; e < 6; e++) row[b * 6 + e] = o[e];
                }
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    bulk_s2g(of + (size_t)c * nB * 6, ringA + (unsigned)s * rowFB, rowFB);
                    bulk_commit();
                }
            }
        }
        if (tid == 0) bulk_wait_read<0>(); // ring and xs are reused by the next pass
        __syncthreads();
        STAMP(5);
    }
    if (tid == 0) {
        bulk_wait<0>();
        int st = L.status[model];
        if (st == HDB_STATUS_OK && s_flag) st = HDB_STATUS_NONFINITE;
        if (st == HDB_STATUS_OK && s_failCase != 0x7fffffff) st = HDB_STATUS_IGNORE_LOAD;
        L.status[model] = st;
        if (L.failCase) L.failCase[model] = (s_failCase != 0x7fffffff) ? s_failCase : -1;
    }
}

static bool tma_eligible(const DevTopo &T, const StaticLaunch &L, int threads)
{
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (T.dof % 2 != 0 || T.dof > 4 * threads || T.nB >= 2 * threads) return false;
    if (!al16(L.loads) || !al16(L.disp) || !al16(L.forces) || !al16(L.react)) return false;
    return true;
}

cudaError_t launch_solve_recover(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    int dev = 0, optin = 0, nSM = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev);
    SolveCfg cfg;
    size_t smem = 0;
    if (!solve_config(T, L.nLC, optin, cfg.casesPerPass, smem)) return cudaErrorInvalidConfiguration;
    cfg.cs = odd_stride(cfg.casesPerPass);
    cfg.nSM = nSM;
    const int threads = solve_threads(T);
    static const bool useTma = std::getenv("HDB_USE_TMA") != nullptr; // experimental: slower than the LDG/STG kernel so far
    if (useTma && tma_eligible(T, L, threads)) {
        TmaCfg tc;
        tc.casesPerPass = cfg.casesPerPass;
        tc.cs = cfg.cs;
        tc.nSM = nSM;
        const size_t ring = std::max<size_t>((size_t)TMA_DROWS * T.dof, (size_t)TMA_FROWS * T.nB * 6);
        tc.ringDoubles = (int)((ring + 1) & ~(size_t)1);
        const int nRest = T.nB > threads ? T.nB - threads : 0;
        const size_t xsD = (((size_t)T.nPad * cfg.cs + 1) & ~(size_t)1);
        const size_t bytes = ((size_t)(T.nSlots + 2 * LV_PAD) + xsD + tc.ringDoubles + (size_t)nRest * 8) * sizeof(double);
        if (bytes + 2048 <= (size_t)optin) {
            if (bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(k_solve_recover_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e != cudaSuccess) return e;
            }
            k_solve_recover_tma<<<L.nModels, threads, bytes, st>>>(T, L, tc);
            return cudaGetLastError();
        }
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_solve_recover, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_solve_recover<<<L.nModels, threads, smem, st>>>(T, L, cfg);
    return cudaGetLastError();
}

} // namespace hdb
