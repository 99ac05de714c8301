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
//   k_solve_pipe     persistent block per SM, three models in flight: streaming warps pull a model's factor and
//                    load cases into shared memory, a dedicated solver warp runs the forward / backward block
//                    substitution with one lane per load case (model.go:440), streaming warps write the
//                    displacements (model.go:446-451), beam end forces k*(T*d) (model.go:455-468) and
//                    reactions -p + K[r,:]d (model.go:470-489), all with wide contiguous accesses.
//   k_uniform_loads  mod.LoadUniform + assemblyNodeLoad on the device (mod/uniform.go:46-211, model.go:683-698).
//
// Everything is HBM-bound integer-indexed FP64 work: no tensor-core shapes exist here (fronts are
// 3x3 node blocks); see DESIGN.md for the roofline of each kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "device_topo.hpp"
#include "element.cuh"
#include "hdb_internal.hpp"
#include "solve_ops.cuh"

namespace hdb {

// ------------------------------------------------------------------------------------------------
// shared element evaluation
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// k_geometry: per-model coordinates -> (cos, sin, length) of every beam with the element routines of the reference
// (feBeam2d.go:11-19, 142-172).  This is also the geometry part of check.go for a batch: a NaN / Inf coordinate
// (check.go:223-245) or a beam of length zero (check.go:294-300) gives a non-finite entry here, which the assembly
// kernel reports as HDB_STATUS_NONFINITE for that model only.
// ------------------------------------------------------------------------------------------------
__global__ void k_geometry(const DevTopo T, const StaticLaunch L)
{
    const size_t n = (size_t)L.nModels * T.nB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i % T.nB);
        const double *p = L.points + (i / T.nB) * (size_t)T.nP * 2;
        const int n0 = T.beamN[2 * b], n1 = T.beamN[2 * b + 1];
        const double x0 = p[2 * n0], y0 = p[2 * n0 + 1], x1 = p[2 * n1], y1 = p[2 * n1 + 1];
        const double len = beam_length(x0, y0, x1, y1);
        double c, s;
        beam_cs(x0, y0, x1, y1, len, c, s);
        L.geom[3 * i] = c;
        L.geom[3 * i + 1] = s;
        L.geom[3 * i + 2] = len;
    }
}

cudaError_t launch_geometry(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    const size_t n = (size_t)L.nModels * T.nB;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_geometry<<<blocks, 256, 0, st>>>(T, L);
    return cudaGetLastError();
}

// Factor workspace W: tiles of W_TILE = 32 models, inside a tile [nSlots][32].  The factor kernel's warp (one tile)
// then works on one contiguous region with 256-byte rows; a model's slots are 256 bytes apart for the assembly.
constexpr int W_TILE = 32;
__device__ __forceinline__ size_t w_tile_base(const DevTopo &T, int model)
{
    return (size_t)(model >> 5) * T.nSlots * W_TILE + (size_t)(model & 31);
}

__device__ __forceinline__ bool section_ok(double v)
{ // check.go:301-324: NaN, Inf, negative, zero are rejected
    return (v > 0.0) && !isinf(v);
}

// ------------------------------------------------------------------------------------------------
// k_assemble
// ------------------------------------------------------------------------------------------------
constexpr int ASM_MAX_THREADS = 512;

// Gather of up to U slots (s, s + stride, ...) from a fixed-width list: U * W index loads, all independent and
// coalesced, before the first dependent shared-memory read.  Index -2 in row 0 = the slot starts at 1.0 (unit
// diagonal of a dummy unknown) and has no contribution.
template <int W, int U>
__device__ __forceinline__ void gather_slots(const int32_t *__restrict__ ell, const int n, const int s, const int stride, const double *ke,
                                             double *out, const size_t outStride)
{
    int idx[U][W];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int k = 0; k < W; k++) idx[u][k] = (s + u * stride < n) ? __ldg(ell + (size_t)k * n + s + u * stride) : -1;
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (s + u * stride >= n) break;
        double sum = idx[u][0] == -2 ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < W; k++)
            if (idx[u][k] >= 0) sum = __dadd_rn(sum, ke[idx[u][k]]);
        out[(size_t)(s + u * stride) * outStride] = sum;
    }
}
template <int U>
__device__ __forceinline__ void gather_slots_w(const int w, const int32_t *__restrict__ ell, const int n, const int s, const int stride,
                                               const double *ke, double *out, const size_t outStride)
{
    switch (w) { // warp-uniform
    case 1: gather_slots<1, U>(ell, n, s, stride, ke, out, outStride); break;
    case 2: gather_slots<2, U>(ell, n, s, stride, ke, out, outStride); break;
    case 3: gather_slots<3, U>(ell, n, s, stride, ke, out, outStride); break;
    case 4: gather_slots<4, U>(ell, n, s, stride, ke, out, outStride); break;
    case 5: gather_slots<5, U>(ell, n, s, stride, ke, out, outStride); break;
    case 6: gather_slots<6, U>(ell, n, s, stride, ke, out, outStride); break;
    case 7: gather_slots<7, U>(ell, n, s, stride, ke, out, outStride); break;
    default: gather_slots<8, U>(ell, n, s, stride, ke, out, outStride); break;
    }
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_assemble(const DevTopo T, const StaticLaunch L)
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
        const BeamGeom g = beam_geom(T, L, model, b);
        const BeamCoef cf = beam_coef(A, J, E, g.length);
        double kg[36];
        const unsigned long long dr = T.droppedMask[b]; // bit e: matrix position e is absent from the representative pattern
        const uint32_t pm = T.pinMask[b];
        if (pm == 0) { // rigid ends: eight distinct values (element.cuh), checked once each
            double d[8];
            rigid_global_values(cf, g.c, g.s, d);
            unsigned long long nz = 0;
            bool fin = true;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                fin = fin && isfinite(d[i]);
                if (d[i] != 0.0) nz |= HD_RIGID_POS(i);
            }
            if (!fin) flag = max(flag, (int)HDB_STATUS_NONFINITE);
            else if (dr & nz) flag = max(flag, (int)HDB_STATUS_PATTERN);
            rigid_global_expand(d, kg);
#pragma unroll
            for (int e = 0; e < 36; e++) ke[b * KE_LD + e] = kg[e];
        } else {
            double k[36];
            stiff_beam2d(cf, pm, k);
            rotate_to_global(k, g.c, g.s, kg);
#pragma unroll
            for (int e = 0; e < 36; e++) {
                const double v = kg[e];
                ke[b * KE_LD + e] = v;
                if (!isfinite(v)) flag = max(flag, (int)HDB_STATUS_NONFINITE);
                else if (((dr >> e) & 1ull) && v != 0.0) flag = max(flag, (int)HDB_STATUS_PATTERN);
            }
        }
    }
    if (flag) atomicMax(&s_flag, flag == HDB_STATUS_NONFINITE ? 100 : flag); // NONFINITE wins
    __syncthreads();

    // phase B: gather; one thread per slot, contributions in beam order.  The fixed-width lists make every
    // index load independent and coalesced (the CSR walk is a chain of dependent global loads per slot).
    if (T.gWidth > 0) {
        for (int s = tid; s < T.nSlots; s += 4 * ASM_THREADS)
            gather_slots_w<4>(T.gWidth, T.gEll, T.nSlots, s, ASM_THREADS, ke, L.W + w_tile_base(T, model), (size_t)W_TILE);
    } else {
        for (int s = tid; s < T.nSlots; s += ASM_THREADS) {
            double sum = T.slotInit[s] ? 1.0 : 0.0;
            const int q1 = T.gptr[s + 1];
            for (int q = T.gptr[s]; q < q1; q++) sum = __dadd_rn(sum, ke[T.gsrc[q]]);
            L.W[w_tile_base(T, model) + (size_t)s * W_TILE] = sum;
        }
    }
    double *ksf = L.Ksf + (size_t)model * T.nnzR;
    if (T.rgWidth > 0) {
        for (int s = tid; s < T.nnzR; s += 4 * ASM_THREADS) gather_slots_w<4>(T.rgWidth, T.rgEll, T.nnzR, s, ASM_THREADS, ke, ksf, 1);
    } else {
        for (int s = tid; s < T.nnzR; s += ASM_THREADS) {
            double sum = 0.0;
            const int q1 = T.rgptr[s + 1];
            for (int q = T.rgptr[s]; q < q1; q++) sum = __dadd_rn(sum, ke[T.rgsrc[q]]);
            ksf[s] = sum;
        }
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

// ------------------------------------------------------------------------------------------------
// k_assemble_quad: the same assembly for four models per block.  The element store is compact (eight distinct values
// of a beam without pins, element.cuh; 36 of a beam with pins) and interleaved over the four models, so one gather
// index load serves four models, the four values come with two 16-byte shared loads and a slot of the tiled
// workspace W is written as one full 32-byte sector.  Sums run in the same (beam) order with the same rounded
// additions as k_assemble: the results are bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int QUAD = 4;
#ifndef QUAD_T
#define QUAD_T 320
#endif
#ifndef QUAD_B
#define QUAD_B 3
#endif
#ifndef QUAD_U
#define QUAD_U 2
#endif
constexpr int QUAD_THREADS = QUAD_T;

__device__ __forceinline__ double flip_sign(double v, int sg) { return __hiloint2double(__double2hiint(v) ^ sg, __double2loint(v)); }

template <int W, int U>
__device__ __forceinline__ void gather_quad(const int32_t *__restrict__ ell, const int n, const int s, const int stride, const double *ke4,
                                            double (&sum)[U][QUAD])
{
    int idx[U][W];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int k = 0; k < W; k++) idx[u][k] = (s + u * stride < n) ? __ldg(ell + (size_t)k * n + s + u * stride) : -1;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const double init = idx[u][0] == -2 ? 1.0 : 0.0; // unit diagonal of a dummy unknown
#pragma unroll
        for (int m = 0; m < QUAD; m++) sum[u][m] = init;
#pragma unroll
        for (int k = 0; k < W; k++)
            if (idx[u][k] >= 0) {
                const double2 *p = reinterpret_cast<const double2 *>(ke4 + (size_t)(idx[u][k] >> 1) * QUAD);
                const double2 a = p[0], b = p[1];
                const int sg = (idx[u][k] & 1) << 31;
                sum[u][0] = __dadd_rn(sum[u][0], flip_sign(a.x, sg));
                sum[u][1] = __dadd_rn(sum[u][1], flip_sign(a.y, sg));
                sum[u][2] = __dadd_rn(sum[u][2], flip_sign(b.x, sg));
                sum[u][3] = __dadd_rn(sum[u][3], flip_sign(b.y, sg));
            }
    }
}

template <int W>
__device__ __forceinline__ void quad_slots(const DevTopo &T, const StaticLaunch &L, const double *ke4, const int model0, const int nm,
                                           const int tid, const int nthreads)
{
    constexpr int U = QUAD_U;
    double *wbase = L.W + w_tile_base(T, model0); // model0 is a multiple of 4: 32-byte aligned rows
    for (int s = tid; s < T.nSlots; s += U * nthreads) {
        double sum[U][QUAD];
        gather_quad<W, U>(T.cgEll, T.nSlots, s, nthreads, ke4, sum);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s2 = s + u * nthreads;
            if (s2 >= T.nSlots) break;
            double *o = wbase + (size_t)s2 * W_TILE;
            if (nm == QUAD) {
                *reinterpret_cast<double2 *>(o) = make_double2(sum[u][0], sum[u][1]);
                *reinterpret_cast<double2 *>(o + 2) = make_double2(sum[u][2], sum[u][3]);
            } else {
#pragma unroll
                for (int m = 0; m < QUAD; m++)
                    if (m < nm) o[m] = sum[u][m];
            }
        }
    }
}
template <int W>
__device__ __forceinline__ void quad_react(const DevTopo &T, const StaticLaunch &L, const double *ke4, const int model0, const int nm,
                                           const int tid, const int nthreads)
{
    constexpr int U = QUAD_U;
    for (int s = tid; s < T.nnzR; s += U * nthreads) {
        double sum[U][QUAD];
        gather_quad<W, U>(T.crgEll, T.nnzR, s, nthreads, ke4, sum);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s2 = s + u * nthreads;
            if (s2 >= T.nnzR) break;
#pragma unroll
            for (int m = 0; m < QUAD; m++)
                if (m < nm) L.Ksf[(size_t)(model0 + m) * T.nnzR + s2] = sum[u][m];
        }
    }
}

// a beam with pins: all 36 entries, literal product (kept out of line: its two 6x6 arrays would set the register
// budget of the whole kernel)
__device__ __noinline__ int quad_pinned_beam(const BeamCoef cf, const uint32_t pm, const double c, const double s, const unsigned long long dr,
                                             double *dst)
{
    double k[36], kg[36];
    int flag = 0;
    stiff_beam2d(cf, pm, k);
    rotate_to_global(k, c, s, kg);
#pragma unroll
    for (int e = 0; e < 36; e++) {
        const double v = kg[e];
        dst[e * QUAD] = v;
        if (!isfinite(v)) flag = max(flag, (int)HDB_STATUS_NONFINITE);
        else if (((dr >> e) & 1ull) && v != 0.0) flag = max(flag, (int)HDB_STATUS_PATTERN);
    }
    return flag;
}

__global__ void __launch_bounds__(QUAD_THREADS, QUAD_B) k_assemble_quad(const DevTopo T, const StaticLaunch L)
{
    extern __shared__ double smem_d[]; // ke4[keLen][QUAD]
    __shared__ int s_flag[QUAD];
    double *ke4 = smem_d;
    const int model0 = blockIdx.x * QUAD;
    const int nm = min(QUAD, L.nModels - model0);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (tid < QUAD) s_flag[tid] = 0;
    __syncthreads();

    // phase A: one thread per beam, two models at a time
    for (int b = tid; b < T.nB; b += nthreads) {
        const BeamGeom g = beam_geom(T, b);
        const unsigned long long dr = T.droppedMask[b];
        const uint32_t pm = T.pinMask[b];
        double *dst = ke4 + (size_t)T.keOff[b] * QUAD;
#pragma unroll 1
        for (int half = 0; half < QUAD; half += 2) {
            double d[2][8];
            int flag[2] = {0, 0};
            BeamCoef cf[2];
            BeamGeom gj[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int mm = min(half + j, nm - 1); // a block past the end of the batch repeats its last model
                const double *sec = L.sections + ((size_t)(model0 + mm) * T.nB + b) * 3;
                const double A = sec[0], J = sec[1], E = sec[2];
                if (!(section_ok(A) && section_ok(J) && section_ok(E))) flag[j] = HDB_STATUS_NONFINITE;
                gj[j] = L.geom ? beam_geom(T, L, model0 + mm, b) : g;
                cf[j] = beam_coef(A, J, E, gj[j].length);
            }
            if (pm == 0) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    rigid_global_values(cf[j], gj[j].c, gj[j].s, d[j]);
                    unsigned long long nz = 0;
                    bool fin = true;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        fin = fin && isfinite(d[j][i]);
                        if (d[j][i] != 0.0) nz |= HD_RIGID_POS(i);
                    }
                    if (!fin) flag[j] = max(flag[j], (int)HDB_STATUS_NONFINITE);
                    else if (dr & nz) flag[j] = max(flag[j], (int)HDB_STATUS_PATTERN);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) *reinterpret_cast<double2 *>(dst + i * QUAD + half) = make_double2(d[0][i], d[1][i]);
            } else {
#pragma unroll 1
                for (int j = 0; j < 2; j++) flag[j] = max(flag[j], quad_pinned_beam(cf[j], pm, gj[j].c, gj[j].s, dr, dst + half + j));
            }
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (flag[j]) atomicMax(&s_flag[half + j], flag[j] == HDB_STATUS_NONFINITE ? 100 : flag[j]); // NONFINITE wins
        }
    }
    __syncthreads();

    // phase B: one thread per slot and four models
    switch (T.gWidth) { // block-uniform
    case 1: quad_slots<1>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 2: quad_slots<2>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 3: quad_slots<3>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 4: quad_slots<4>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 5: quad_slots<5>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 6: quad_slots<6>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 7: quad_slots<7>(T, L, ke4, model0, nm, tid, nthreads); break;
    default: quad_slots<8>(T, L, ke4, model0, nm, tid, nthreads); break;
    }
    switch (T.rgWidth) {
    case 1: quad_react<1>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 2: quad_react<2>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 3: quad_react<3>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 4: quad_react<4>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 5: quad_react<5>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 6: quad_react<6>(T, L, ke4, model0, nm, tid, nthreads); break;
    case 7: quad_react<7>(T, L, ke4, model0, nm, tid, nthreads); break;
    default: quad_react<8>(T, L, ke4, model0, nm, tid, nthreads); break;
    }
    if (tid < nm) L.status[model0 + tid] = (s_flag[tid] == 100) ? (int)HDB_STATUS_NONFINITE : s_flag[tid];
}

bool assemble_fits_smem(const DevTopo &T, int smemOptin)
{
    return (size_t)T.nB * KE_LD * sizeof(double) + 1024 <= (size_t)smemOptin;
}

cudaError_t launch_assemble(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    { // four models per block over the compact element store, when two such blocks fit an SM
        const size_t quadSmem = (size_t)T.keLen * QUAD * sizeof(double);
        const char *off = std::getenv("HDB_ASSEMBLE_SINGLE"); // diagnostics: the one-model kernel
        if (T.keLen > 0 && !L.kval && !L.keScratch && quadSmem <= 110 * 1024 && !(off && atoi(off))) {
            if (quadSmem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(k_assemble_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)quadSmem);
                if (e != cudaSuccess) return e;
            }
            cudaFuncSetAttribute(k_assemble_quad, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int th = ((T.nB + 31) / 32) * 32;
            th = th < 128 ? 128 : (th > QUAD_THREADS ? QUAD_THREADS : th);
            k_assemble_quad<<<(L.nModels + QUAD - 1) / QUAD, th, quadSmem, st>>>(T, L);
            return cudaGetLastError();
        }
    }
    const size_t smem = L.keScratch ? 0 : (size_t)T.nB * KE_LD * sizeof(double);
    int threads = ((T.nB + 31) / 32) * 32;
    threads = threads < 128 ? 128 : (threads > ASM_MAX_THREADS ? ASM_MAX_THREADS : threads);
    // up to 320 threads two blocks share an SM (register budget set by the launch bounds); the carve-out
    // preference keeps the driver from picking a shared-memory split that fits only one
    auto kern = threads <= 320 ? k_assemble<320, 2> : k_assemble<ASM_MAX_THREADS, 1>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<L.nModels, threads, smem, st>>>(T, L);
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

// sixteen consecutive slots (one banded block column) that start on a multiple of 16: no wrap inside the window, so
// one address pair and immediate offsets instead of sixteen rounds of index arithmetic
template <int K>
__device__ __forceinline__ void cp_async8_run(const unsigned s, const double *g)
{
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%2], 8;" ::"r"(s), "l"(g), "n"((16 - K) * 256) : "memory");
    if constexpr (K > 1) cp_async8_run<K - 1>(s, g);
}

template <int WS>
__global__ void __launch_bounds__(32) k_factor(const DevTopo T, const StaticLaunch L)
{
    extern __shared__ __align__(16) double smem_d[];
    const int lane = threadIdx.x;
    const int model = blockIdx.x * 32 + lane; // ld is a multiple of 32: padding lanes work on scratch columns
    double *wg = L.W + w_tile_base(T, model); // the 32 models of this warp are one tile: its whole workspace is contiguous
    const size_t ld = (size_t)W_TILE;
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
            if (WS >= 32 && d.z - loaded == 16 && (loaded & 15) == 0) {
                cp_async8_run<16>((unsigned)__cvta_generic_to_shared(&WIN(loaded)), wg + (size_t)loaded * ld);
                loaded += 16;
            } else {
                for (; loaded < d.z; loaded++) cp_async8(&WIN(loaded), wg + (size_t)loaded * ld);
            }
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
        // header for the solve kernel: 1/d and the strictly lower part of V = Ldiag^-1 (both substitutions then
        // finish a block with independent multiply-adds instead of a dependent chain)
        const double v10 = -l10, v20 = fma(l10, l21, -l20), v21 = -l21;
        const int nb = d.y;
        double *lm = L.Lm + (size_t)model * T.nSlots; // finished columns leave model-major: the solve kernel reads a model's
                                                      // factor as one contiguous run (16-byte stores: base is even)
        *reinterpret_cast<double2 *>(lm + base) = make_double2(inv0, inv1);
        *reinterpret_cast<double2 *>(lm + base + 2) = make_double2(inv2, v10);
        *reinterpret_cast<double2 *>(lm + base + 4) = make_double2(v20, v21);
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
            // nobody reads this off block from the window again (the Schur update below has it in registers): it goes
            // straight to the model-major copy
#pragma unroll
            for (int e = 0; e < 8; e += 2) *reinterpret_cast<double2 *>(lm + sa + e) = make_double2(la[e], la[e + 1]);
            *reinterpret_cast<double2 *>(lm + sa + 8) = make_double2(la[8], 0.0);
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
                double w[9]; // read, compute, write back together (see the Schur updates below)
#pragma unroll
                for (int e = 0; e < 9; e++) w[e] = WIN(sa + e);
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double u0 = w[3 * r];
                    const double u1 = fma(-l10, u0, w[3 * r + 1]);
                    const double u2 = fma(-l21, u1, fma(-l20, u0, w[3 * r + 2]));
                    w[3 * r] = u0 * inv0;
                    w[3 * r + 1] = u1 * inv1;
                    w[3 * r + 2] = u2 * inv2;
                }
#pragma unroll
                for (int e = 0; e < 9; e++) WIN(sa + e) = w[e];
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
                    // targets are read into registers first and written back together: the window index is masked,
                    // so the compiler cannot tell the slots apart and would serialise nine load-FMA-store chains
                    if (a == b) { // diagonal block of row block I_a: header [d0 d1 d2 | (1,0) (2,0) (2,1)]
                        double tt[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) tt[e] = WIN(t + e);
#pragma unroll
                        for (int r = 0; r < 3; r++)
                            tt[r] = fma(-la[3 * r + 2], ub[3 * r + 2], fma(-la[3 * r + 1], ub[3 * r + 1], fma(-la[3 * r], ub[3 * r], tt[r])));
                        tt[3] = fma(-la[5], ub[2], fma(-la[4], ub[1], fma(-la[3], ub[0], tt[3])));
                        tt[4] = fma(-la[8], ub[2], fma(-la[7], ub[1], fma(-la[6], ub[0], tt[4])));
                        tt[5] = fma(-la[8], ub[5], fma(-la[7], ub[4], fma(-la[6], ub[3], tt[5])));
#pragma unroll
                        for (int e = 0; e < 6; e++) WIN(t + e) = tt[e];
                    } else { // block (I_a, I_b): rows from a, columns from b
                        double tt[9];
#pragma unroll
                        for (int e = 0; e < 9; e++) tt[e] = WIN(t + e);
#pragma unroll
                        for (int r = 0; r < 3; r++)
#pragma unroll
                            for (int rp = 0; rp < 3; rp++)
                                tt[3 * r + rp] = fma(-la[3 * r + 2], ub[3 * rp + 2],
                                                     fma(-la[3 * r + 1], ub[3 * rp + 1], fma(-la[3 * r], ub[3 * rp], tt[3 * r + rp])));
#pragma unroll
                        for (int e = 0; e < 9; e++) WIN(t + e) = tt[e];
                    }
                }
            }
        }
        // ---- off blocks of a column with several of them: stream them out from the window ----
        if (nb > 1) {
            const int end = base + BLK_HDR + BLK_STRIDE * nb;
            for (int s2 = base + BLK_HDR; s2 < end; s2 += 2) *reinterpret_cast<double2 *>(lm + s2) = make_double2(WIN(s2), WIN(s2 + 1));
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
// k_solve_pipe: load cases -> substitution -> displacements, reactions, beam forces (model.go:421-491)
//
// One persistent block per SM works through its models as a three-stage software pipeline over two
// shared-memory buffers.  In iteration i
//   * the solver warp runs the two triangular solves of model i-1: a sequential chain of 3x3 block ops,
//     one lane per load case.  It has SM sub-partition 0 to itself (warps 4, 8, 12 only take part in the
//     block-wide barrier): shared-memory loads that queue behind the global traffic of other warps in the
//     same sub-partition made the chain 2.5x slower;
//   * the streaming warps (sub-partitions 1-3) first turn model i-2's solution into its three outputs, then
//     pull model i in: factor and load cases,
// so the memory system always has twelve warps of loads or stores in flight while the chain runs.
//
// Shared memory of one buffer: the factor of the model [nSlots] (reused for the beam coefficients and the
// reaction block once the solve is over), and xs[nBlk][cases][3] (a block's three unknowns of one load
// case are adjacent, the lane stride of 24 bytes is bank-conflict free).  Integer topology tables (op
// lists, DOF info, beam rows, reaction rows) are staged once per block.
// ------------------------------------------------------------------------------------------------
constexpr int PIPE_THREADS = 512; // warp 0 = solver, warps 4, 8, 12 idle, the other twelve stream
constexpr int LC = 5;             // load cases handled together by one work item of the streaming stages
constexpr unsigned ZERO_BYTES = (24u * LC + 15u) & ~15u; // a zero "row" long enough for the LC cases of a work unit
constexpr int NS = PIPE_THREADS / 32 / 4 * 3 * 32; // streaming threads: the warps outside SM sub-partition 0
constexpr int NH = (PIPE_THREADS / 32 / 4 - 1) * 32; // helper threads: warps 4, 8, 12 of the solver's sub-partition
constexpr int REACT_HELPER_PCT = 57;               // share of the reaction items the helper warps take
constexpr int LOAD_U = 7;         // 16-byte loads in flight per streaming thread in the load stage
constexpr unsigned FST_BYTES = 32u * 6u * 8u; // per-warp 32 x 6 staging tile of the force stage

static size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

static size_t pipe_idx_bytes(const DevTopo &T)
{
    return (size_t)(T.nFwdOps + T.nBwdOps) * 16 +
           align16(((size_t)T.dof + (size_t)2 * T.nSup + 1 + (size_t)T.nB) * 4 + ((size_t)6 * T.nB + (size_t)T.nnzR) * 2 + 16) + (size_t)T.nB * 24 + 32 + ZERO_BYTES +
           (size_t)(NS / 32) * FST_BYTES + 32;
}

static size_t pipe_lvm_doubles(const DevTopo &T)
{
    // solve: factor of the model; recovery: beam coefficients + reaction block K[sup, free] of the model
    size_t d = std::max<size_t>((size_t)T.nSlots, (size_t)T.nB * 5 + (size_t)T.nnzR);
    return (d + 1) & ~(size_t)1;
}

static size_t solve_smem_bytes(const DevTopo &T, int cp)
{
    return pipe_idx_bytes(T) + 2 * (pipe_lvm_doubles(T) + (size_t)T.nBlk * xs_block_doubles(cp)) * sizeof(double);
}

bool solve_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes)
{
    int cp = nLC < 32 ? nLC : 32;
    if (const char *e = std::getenv("HDB_CASES_PER_PASS")) cp = std::min(cp, atoi(e)); // diagnostics: smaller passes
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

template <bool DBG> // DBG: per-phase cycle counters (HDB_DEBUG_CLOCKS); a separate instantiation keeps them out of the product's registers
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_solve_pipe(const DevTopo T, const StaticLaunch L, const SolveCfg cfg)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ unsigned s_fail[2], s_failNF[2];
    const int tid = threadIdx.x;
    // The solver warp keeps its scheduler (SM sub-partition 0 = warps 0, 4, 8, ...) to itself: its shared-memory
    // loads would otherwise queue behind the global loads and stores of streaming warps in the same
    // instruction queue.  Warps 4, 8, 12 only take part in the block-wide barriers.
    const int warp = tid >> 5;
    const bool streamer = (warp & 3) != 0;
    const int st = (warp - (warp >> 2) - 1) * 32 + (tid & 31); // streaming thread index
    const int dof = T.dof, nB = T.nB, CSP = cfg.casesPerPass;
    // Shared memory is addressed through 32-bit shared-window addresses and explicit ld/st.shared: generic
    // pointers into dynamic shared memory make the compiler re-derive the window base (S2R SR_CgaCtaId, an
    // XU-pipe instruction) inside every loop, which saturated the XU pipe in the first version of this kernel.
    const unsigned smA = (unsigned)__cvta_generic_to_shared(smem_d);
    const unsigned opsA = smA; // forward list, backward list
    const unsigned xbA = opsA + (unsigned)(T.nFwdOps + T.nBwdOps) * 16u;                   // [dof] DOF info, see below
    // 16-bit tables hold xs offsets in units of 8 bytes, 0xffff = none: they keep the block under the 196 kB
    // shared-memory configuration, which leaves L1 twice the room of the 228 kB one
    const unsigned bbA = xbA + (unsigned)dof * 4u;                                         // [nB*6] u16 xs offset / 8 or 0xffff
    const unsigned rpA = (bbA + (unsigned)nB * 12u + 3u) & ~3u;                            // [nSup+1]
    const unsigned rcA = rpA + (unsigned)(T.nSup + 1) * 4u;                                // [nnzR] u16 xs offset / 8
    const unsigned sdA = (rcA + (unsigned)T.nnzR * 2u + 3u) & ~3u;                         // [nSup] supported DOFs
    const unsigned pmA = sdA + (unsigned)T.nSup * 4u;                                      // [nB] pin masks
    const unsigned gmA = (pmA + (unsigned)nB * 4u + 15u) & ~15u;                           // [3][nB] beam cos, sin, length
    const unsigned zeroA = (gmA + (unsigned)nB * 24u + 15u) & ~15u;                        // ZERO_BYTES zeros: "row" of a supported DOF
    const unsigned fst0 = zeroA + ZERO_BYTES;                                                     // staging tiles, one per streaming warp
    const unsigned lvm0 = fst0 + (unsigned)(NS / 32) * FST_BYTES;
    const unsigned lvmBytes = (unsigned)cfg.lvmDoubles * 8u, xsBytes = (unsigned)cfg.xsDoubles * 8u;
    const unsigned xs0 = lvm0 + 2u * lvmBytes;
    const unsigned BS = 8u * (unsigned)xs_block_doubles(CSP); // bytes of one block of xs: [cases][3] (+ pad)

    // ---- once per block: integer tables -> shared memory ----
    auto xoff = [&](int r) { return r >= 0 ? (int)((unsigned)(r / 3) * BS + (unsigned)(r % 3) * 8u) : r; };
    for (int k = tid; k < T.nFwdOps + T.nBwdOps; k += blockDim.x) {
        const int4 r = k < T.nFwdOps ? __ldg(T.fwdOps + k) : __ldg(T.bwdOps + (k - T.nFwdOps));
        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(opsA + 16u * k), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
    }
    // DOF info: xs byte offset (a multiple of 8) for a free DOF, else negative; bit 0 = supported, bit 1 = ignored
    // (no stiffness: its load must be zero, model.go:424-437)
    for (int g = tid; g < dof; g += blockDim.x) {
        const int xr = __ldg(T.xrow + g);
        sts_s32(xbA + 4u * g, (xr >= 0 ? xoff(xr) : (int)0x80000000) | (__ldg(T.supIndex + g) >= 0 ? 1 : 0) | (xr == -2 ? 2 : 0));
    }
    for (int k = tid; k < nB * 6; k += blockDim.x) {
        const int r = xoff(__ldg(T.beamRows + k));
        sts_u16(bbA + 2u * k, r >= 0 ? (unsigned)r >> 3 : 0xffffu);
    }
    for (int k = tid; k <= T.nSup; k += blockDim.x) sts_s32(rpA + 4u * k, __ldg(T.rptr + k));
    for (int k = tid; k < T.nnzR; k += blockDim.x) sts_u16(rcA + 2u * k, (unsigned)xoff(__ldg(T.rcol + k)) >> 3);
    for (int k = tid; k < T.nSup; k += blockDim.x) sts_s32(sdA + 4u * k, __ldg(T.supDofs + k));
    for (int b = tid; b < nB; b += blockDim.x) { // geometry is shared by all models: once per block
        const BeamGeom g = beam_geom(T, b);
        sts_f64(gmA + (unsigned)b * 8u, g.c);
        sts_f64(gmA + (unsigned)(nB + b) * 8u, g.s);
        sts_f64(gmA + (unsigned)(2 * nB + b) * 8u, g.length);
        sts_s32(pmA + 4u * b, (int)T.pinMask[b]);
    }
    if (tid < (int)(ZERO_BYTES / 8)) sts_f64(zeroA + 8u * tid, 0.0);
    if (tid < 2) {
        s_fail[tid] = NO_FAIL_CASE;
        s_failNF[tid] = NO_FAIL_CASE;
    }
    __syncthreads();

    const int nMy = (L.nModels - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x; // models of this block
    const int nIt = nMy * cfg.nPass;
    long long tSolve = 0, tAll = DBG ? clock64() : 0;
    long long tPh[DBG ? 16 : 1] = {0}; // streaming sub-phases
#define TICK(k)                                  \
    do {                                         \
        if (DBG) {                               \
            const long long t1_ = clock64();     \
            tPh[k] += t1_ - t0;                  \
            t0 = t1_;                            \
        }                                        \
    } while (0)

    // displacements (model.go:446-451): the [cases][dof] block of a model is one contiguous run, walked in pairs with a
    // constant stride so that every store is base + immediate; (DOF, case) follow incrementally.  The helper warps take
    // their share (threads NS .. NS + NH - 1 of the walk): they would otherwise wait for the reaction block.
    constexpr int ND = NS + NH;
    auto disp_items = [&](const int model, const int c0, const int nc, const unsigned xsB, const int t) {
        double *ob = L.disp + ((size_t)model * L.nLC + c0) * dof + 2 * t;
        const int nFlat = nc * dof;
        int g = (2 * t) % dof;
        unsigned cOff = (unsigned)((2 * t) / dof) * 24u;
        for (int f0 = 0; f0 < nFlat; f0 += 2 * ND * 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int f = f0 + 2 * (u * ND + t);
                if (f < nFlat) {
                    int g1 = g + 1;
                    unsigned c1 = cOff;
                    if (g1 == dof) { g1 = 0; c1 += 24u; }
                    const int i0 = lds_s32(xbA + 4u * g), i1 = lds_s32(xbA + 4u * g1);
                    const double d0 = i0 >= 0 ? lds_f64(xsB + (unsigned)i0 + cOff) : 0.0;
                    const double d1 = i1 >= 0 ? lds_f64(xsB + (unsigned)i1 + c1) : 0.0;
                    double *o = ob + f0 + 2 * u * ND;
                    if (cfg.vec2) __stcs(reinterpret_cast<double2 *>(o), make_double2(d0, d1));
                    else {
                        __stcs(o, d0);
                        if (f + 1 < nFlat) __stcs(o + 1, d1);
                    }
                }
                g += (2 * ND) % dof;
                cOff += (unsigned)((2 * ND) / dof) * 24u;
                if (g >= dof) { g -= dof; cOff += 24u; }
            }
        }
    };

    // reactions (model.go:470-489: columns ascending, eliminated columns carry d = 0) added to the -p rows of the
    // load stage with one reduction per entry; item = (group of LC cases, supported DOF)
    auto react_items = [&](double *orr, const unsigned xsB, const unsigned ksA, const int nc, const int first, const int end, const int stride) {
        for (int w = first; w < end; w += stride) {
            const int cg = w / T.nSup, si = w - cg * T.nSup, c = cg * LC;
            const int q0 = lds_s32(rpA + 4u * si), q1 = lds_s32(rpA + 4u * si + 4u);
            double acc[LC];
#pragma unroll
            for (int j = 0; j < LC; j++) acc[j] = 0.0;
            for (int q = q0; q < q1; q++) {
                const double kv = lds_f64(ksA + 8u * q);
                const unsigned xa = xsB + lds_u16(rcA + 2u * q) * 8u + (unsigned)c * 24u;
#pragma unroll
                for (int j = 0; j < LC; j++)
                    if (c + j < nc) acc[j] = fma(kv, lds_f64(xa + 24u * j), acc[j]);
            }
            double *pr = orr + (size_t)c * dof + lds_s32(sdA + 4u * si);
#pragma unroll
            for (int j = 0; j < LC; j++)
                if (c + j < nc) atomicAdd(pr + (size_t)j * dof, acc[j]); // RED: no return value
        }
    };

    for (int i = 0; i < nIt + 2; i++) {
        if (tid < 32) {
            // ================= solver warp: item i-1 =================
            if (i >= 1 && i <= nIt) {
                const int it = i - 1, pass = it % cfg.nPass;
                const int nc = min(CSP, L.nLC - pass * CSP);
                const unsigned buf = (unsigned)(it & 1);
                if (tid < nc) {
                    const long long t0 = DBG ? clock64() : 0;
                    const unsigned tabA = lvm0 + buf * lvmBytes, xA = xs0 + buf * xsBytes + (unsigned)tid * 24u;
                    walk_ops<true>(opsA, tabA, xA, BS);
                    walk_ops<false>(opsA + 16u * (unsigned)T.nFwdOps, tabA, xA, BS);
                    if (DBG) tSolve += clock64() - t0;
                }
            }
        } else if (streamer) {
            // ================= streaming warps =================
            const unsigned buf = (unsigned)(i & 1);
            const unsigned lvB = lvm0 + buf * lvmBytes, xsB = xs0 + buf * xsBytes;
            long long t0 = DBG ? clock64() : 0;
            if (i >= 2) {
                // ---------- recover item i-2: displacements, reactions, beam forces ----------
                const int it = i - 2, pass = it % cfg.nPass;
                const int model = (int)blockIdx.x + (it / cfg.nPass) * (int)gridDim.x;
                const int c0 = pass * CSP, nc = min(CSP, L.nLC - c0);
                const int nGroups = (nc + LC - 1) / LC;
                double *orr = L.react ? L.react + ((size_t)model * L.nLC + c0) * dof : nullptr;
                const unsigned cs = (unsigned)nB * 8u, ksA = lvB + 5u * cs;
                // the factor buffer is dead once the solve is over: it takes the beam coefficients [5][nB] and the
                // reaction block K[sup, free] of the model (one coalesced read instead of a dependent global
                // load per entry)
                // their global reads (first round) are issued here and consumed after the displacement stores, which
                // hide the latency
                const double *ksf = L.Ksf + (size_t)model * T.nnzR;
                const double *sec = L.sections + (size_t)model * nB * 3;
                double kv[4], s3[3];
#pragma unroll
                for (int u = 0; u < 4; u++) kv[u] = (orr && u * NS + st < T.nnzR) ? __ldcs(ksf + u * NS + st) : 0.0;
#pragma unroll
                for (int e = 0; e < 3; e++) s3[e] = (L.forces && st < nB) ? __ldg(sec + 3 * st + e) : 1.0;
                TICK(0);
                // displacements (model.go:446-451): the [cases][dof] block of a model is one contiguous run, walked
                // in pairs with a constant stride so that every store is base + immediate; (DOF, case) follow
                // incrementally
                if (L.disp) disp_items(model, c0, nc, xsB, st);
                if (orr) {
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u * NS + st < T.nnzR) sts_f64(ksA + 8u * (unsigned)(u * NS + st), kv[u]);
                    for (int q = 4 * NS + st; q < T.nnzR; q += NS) sts_f64(ksA + 8u * q, __ldcs(ksf + q));
                }
                if (L.forces) {
                    for (int b = st; b < nB; b += NS) {
                        const bool first = b == st;
                        if (L.geom) { // coordinates vary per model: this model's row replaces the shared geometry table
                            const BeamGeom g = beam_geom(T, L, model, b);
                            sts_f64(gmA + (unsigned)b * 8u, g.c);
                            sts_f64(gmA + (unsigned)(nB + b) * 8u, g.s);
                            sts_f64(gmA + (unsigned)(2 * nB + b) * 8u, g.length);
                        }
                        const BeamCoef cf = beam_coef(first ? s3[0] : sec[3 * b], first ? s3[1] : sec[3 * b + 1], first ? s3[2] : sec[3 * b + 2],
                                                      lds_f64(gmA + (unsigned)(2 * nB + b) * 8u));
                        const unsigned a = lvB + (unsigned)b * 8u; // component-major: conflict-free
                        sts_f64(a, cf.EFL); sts_f64(a + cs, cf.EJL3); sts_f64(a + 2u * cs, cf.EJL2); sts_f64(a + 3u * cs, cf.EJL);
                        sts_f64(a + 4u * cs, cf.EJL4);
                    }
                }
                TICK(2);
                bar_stream(NS + NH); // coefficient table and reaction block complete
                TICK(3);
                // reactions (model.go:470-489: columns ascending, eliminated columns carry d = 0) added to the -p
                // rows of the load stage with one reduction per entry; item = (group of LC cases, supported DOF)
                if (orr) react_items(orr, xsB, ksA, nc, nGroups * T.nSup * REACT_HELPER_PCT / 100 + st, nGroups * T.nSup, NS);
                TICK(4);
                // beam end forces s = k (T Zo), model.go:455-468.  Unit of work of a warp = (32 consecutive beams,
                // group of LC cases); the 192 results of one case go through the warp's staging tile so that every
                // store instruction writes one contiguous run (512 bytes) instead of 8-byte pieces 48 bytes apart.
                if (L.forces) {
                    double *out = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
                    const int nBG = (nB + 31) / 32, sw = st >> 5, nSW = NS >> 5;
                    const unsigned lane = (unsigned)tid & 31u;
                    const unsigned wb = fst0 + (unsigned)sw * FST_BYTES;
                    for (int w = sw; w < nGroups * nBG; w += nSW) {
                        const int cg = w / nBG, bg = w - cg * nBG, c = cg * LC;
                        const int b = bg * 32 + (int)lane;
                        const bool live = b < nB;
                        const int bl = live ? b : 0;
                        const unsigned a = lvB + (unsigned)bl * 8u;
                        BeamCoef cf;
                        cf.EFL = lds_f64(a); cf.EJL3 = lds_f64(a + cs); cf.EJL2 = lds_f64(a + 2u * cs); cf.EJL = lds_f64(a + 3u * cs);
                        cf.EJL4 = lds_f64(a + 4u * cs);
                        const unsigned ga = gmA + (unsigned)bl * 8u;
                        const double gc = lds_f64(ga), gs = lds_f64(ga + cs);
                        const uint32_t pm = live ? (uint32_t)lds_s32(pmA + 4u * b) : 0u;
                        unsigned ra[6]; // xs address of the beam's six DOFs for case c; supported DOFs read the zero row
#pragma unroll
                        for (int e = 0; e < 6; e++) {
                            const unsigned r = lds_u16(bbA + 2u * (unsigned)(bl * 6 + e));
                            ra[e] = (live && r != 0xffffu) ? xsB + r * 8u + (unsigned)c * 24u : zeroA;
                        }
                        // released ends are rare: a warp with any takes the out-of-line general element for all its
                        // lanes, so that the rigid path never touches the local arrays that call needs
                        const bool anyPin = __any_sync(0xffffffffu, pm != 0u);
                        double *ow = out + ((size_t)c * nB + bg * 32) * 6 + 2 * lane;
                        const size_t caseStride = (size_t)nB * 6;
                        if (!anyPin && c + LC <= nc && bg * 32 + 32 <= nB && cfg.vec2) {
                            // the common unit: full warp, full case group, rigid ends, aligned output: no predicates
#pragma unroll
                            for (int j = 0; j < LC; j++) {
                                double u[6], o[6];
#pragma unroll
                                for (int e = 0; e < 6; e++) u[e] = lds_f64(ra[e] + 24u * j);
                                beam_forces_rigid(cf, gc, gs, u, o);
                                __syncwarp();
#pragma unroll
                                for (int e = 0; e < 6; e += 2)
                                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(wb + (lane * 6u + e) * 8u), "d"(o[e]), "d"(o[e + 1]) : "memory");
                                __syncwarp();
#pragma unroll
                                for (int e = 0; e < 3; e++) {
                                    double x, y;
                                    lds_v2o<0>(wb + (e * 64u + 2u * lane) * 8u, x, y);
                                    __stcs(reinterpret_cast<double2 *>(ow + e * 64), make_double2(x, y));
                                }
                                ow += caseStride;
                            }
                            continue;
                        }
                        const int nOut = min(32, nB - bg * 32) * 6; // doubles this warp writes per case
#pragma unroll
                        for (int j = 0; j < LC; j++) {
                            if (c + j >= nc) break;
                            double o[6];
                            if (!anyPin) {
                                double u[6];
#pragma unroll
                                for (int e = 0; e < 6; e++) u[e] = lds_f64(ra[e] + 24u * j);
                                beam_forces_rigid(cf, gc, gs, u, o);
                            } else {
                                double uu[6], oo[6];
#pragma unroll
                                for (int e = 0; e < 6; e++) uu[e] = lds_f64(ra[e] + 24u * j);
                                beam_forces_pinned(cf, pm, gc, gs, uu, oo);
#pragma unroll
                                for (int e = 0; e < 6; e++) o[e] = oo[e];
                            }
                            __syncwarp();
#pragma unroll
                            for (int e = 0; e < 6; e += 2)
                                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(wb + (lane * 6u + e) * 8u), "d"(o[e]), "d"(o[e + 1]) : "memory");
                            __syncwarp();
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                const unsigned q = 2u * (e * 32u + lane); // doubles q, q+1 of the 192
                                if ((int)q < nOut) {
                                    double x, y;
                                    lds_v2o<0>(wb + q * 8u, x, y);
                                    if (cfg.vec2) __stcs(reinterpret_cast<double2 *>(ow + e * 64), make_double2(x, y));
                                    else { __stcs(ow + e * 64, x); __stcs(ow + e * 64 + 1, y); }
                                }
                            }
                            ow += caseStride;
                        }
                    }
                }
                TICK(5);
                bar_stream(NS + NH); // every reader of this buffer is done: the load stage may overwrite it
                TICK(6);
            }
            if (i < nIt) {
                // ---------- load item i: factor, load cases ----------
                const int it = i, pass = it % cfg.nPass;
                const int model = (int)blockIdx.x + (it / cfg.nPass) * (int)gridDim.x;
                const int c0 = pass * CSP, nc = min(CSP, L.nLC - c0);
                int sOld = 0; // needed at the end of the stage only; volatile asm keeps the load up here
                if (st == 0) asm volatile("ld.global.s32 %0, [%1];" : "=r"(sOld) : "l"(L.status + model) : "memory");
                // factor of this model: one contiguous run; the first eight pieces of every thread stay in flight
                // across the load-case loop below
                const double *wf = L.Lm + (size_t)model * T.nSlots + 2 * st;
                double2 fv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int s = 2 * (u * NS + st);
                    fv[u] = s < T.nSlots ? __ldcs(reinterpret_cast<const double2 *>(wf + 2 * u * NS)) : make_double2(0.0, 0.0);
                }
                TICK(7);
                // load cases -> xs (free rows), dummy rows = 0; checks of model.go:692-696 and 424-437.  The
                // [cases][dof] block is one contiguous run walked in pairs with a constant stride: LOAD_U
                // independent base + immediate 16-byte loads per thread are in flight.  The reaction output
                // starts here as -p on supported DOFs (model.go:478) and 0 elsewhere, so p is never read twice.
                const double *pb = L.loads + ((size_t)model * L.nLC + c0) * dof + 2 * st;
                double *orr = L.react ? L.react + ((size_t)model * L.nLC + c0) * dof + 2 * st : nullptr;
                const int nFlat = nc * dof;
                unsigned failNF = NO_FAIL_CASE, failC = NO_FAIL_CASE;
                {
                    int g = (2 * st) % dof;
                    unsigned cOff = (unsigned)((2 * st) / dof) * 24u;
                    for (int f0 = 0; f0 < nFlat; f0 += 2 * NS * LOAD_U) {
                        double2 v[LOAD_U];
                        if (cfg.vec2 && f0 + 2 * NS * LOAD_U <= nFlat) {
#pragma unroll
                            for (int u = 0; u < LOAD_U; u++) v[u] = __ldcs(reinterpret_cast<const double2 *>(pb + f0 + 2 * u * NS));
                        } else {
#pragma unroll
                            for (int u = 0; u < LOAD_U; u++) {
                                const int f = f0 + 2 * (u * NS + st);
                                if (cfg.vec2) v[u] = f < nFlat ? __ldcs(reinterpret_cast<const double2 *>(pb + f0 + 2 * u * NS)) : make_double2(0.0, 0.0);
                                else {
                                    v[u].x = f < nFlat ? __ldcs(pb + f0 + 2 * u * NS) : 0.0;
                                    v[u].y = f + 1 < nFlat ? __ldcs(pb + f0 + 2 * u * NS + 1) : 0.0;
                                }
                            }
                        }
                        // DOF info of all pairs of the round first (independent shared-memory reads, issued while the
                        // global loads are in flight), then the stores: the ld/st.shared wrappers are ordered asm
                        // statements, so interleaving them per pair would serialise one shared-memory latency per pair
                        unsigned a0[LOAD_U], a1[LOAD_U], fl = 0; // xs address or 0 (not free); 4 flag bits per pair
                        static_assert(LOAD_U <= 8, "4 flag bits per pair");
#pragma unroll
                        for (int u = 0; u < LOAD_U; u++) {
                            const int f = f0 + 2 * (u * NS + st);
                            int g1 = g + 1;
                            unsigned c1 = cOff;
                            if (g1 == dof) { g1 = 0; c1 += 24u; }
                            const int i0 = f < nFlat ? lds_s32(xbA + 4u * g) : (int)0x80000000;
                            const int i1 = f + 1 < nFlat ? lds_s32(xbA + 4u * g1) : (int)0x80000000;
                            a0[u] = i0 >= 0 ? xsB + (unsigned)i0 + cOff : 0u;
                            a1[u] = i1 >= 0 ? xsB + (unsigned)i1 + c1 : 0u;
                            fl |= (unsigned)((i0 & 3) | ((i1 & 3) << 2)) << (4 * u);
                            g += (2 * NS) % dof;
                            cOff += (unsigned)((2 * NS) / dof) * 24u;
                            if (g >= dof) { g -= dof; cOff += 24u; }
                        }
#pragma unroll
                        for (int u = 0; u < LOAD_U; u++) {
                            const int f = f0 + 2 * (u * NS + st);
                            if (f < nFlat) {
                                const double v0 = v[u].x, v1 = v[u].y;
                                const unsigned b4 = fl >> (4 * u);
                                if (!(fabs(v0) <= 1.7976931348623157e308)) failNF = min(failNF, (unsigned)(c0 + f / dof)); // rare
                                if (!(fabs(v1) <= 1.7976931348623157e308) && f + 1 < nFlat) failNF = min(failNF, (unsigned)(c0 + (f + 1) / dof));
                                if (a0[u]) sts_f64(a0[u], v0);
                                if (a1[u]) sts_f64(a1[u], v1);
                                if (orr) {
                                    const double r0 = (b4 & 1u) ? -v0 : 0.0, r1 = (b4 & 4u) ? -v1 : 0.0;
                                    double *o = orr + f0 + 2 * u * NS;
                                    if (cfg.vec2) *reinterpret_cast<double2 *>(o) = make_double2(r0, r1);
                                    else {
                                        o[0] = r0;
                                        if (f + 1 < nFlat) o[1] = r1;
                                    }
                                }
                                if (b4 & 10u) { // a load on a DOF without stiffness (model.go:424-437); rare
                                    if ((b4 & 2u) && v0 != 0.0) failC = min(failC, (unsigned)(c0 + f / dof));
                                    if ((b4 & 8u) && v1 != 0.0) failC = min(failC, (unsigned)(c0 + (f + 1) / dof));
                                }
                            }
                        }
                    }
                }
                for (int idx = st; idx < T.nDummy * nc; idx += NS) {
                    const int q = idx / nc, c = idx - q * nc;
                    const int r = __ldg(T.dummyPos + q);
                    sts_f64(xsB + (unsigned)(r / 3) * BS + (unsigned)(r % 3) * 8u + (unsigned)c * 24u, 0.0);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int s = 2 * (u * NS + st);
                    if (s < T.nSlots) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(lvB + (unsigned)s * 8u), "d"(fv[u].x), "d"(fv[u].y) : "memory");
                }
                for (int s = 2 * (4 * NS + st); s < T.nSlots; s += 2 * NS) {
                    const double2 v = __ldcs(reinterpret_cast<const double2 *>(L.Lm + (size_t)model * T.nSlots + s));
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(lvB + (unsigned)s * 8u), "d"(v.x), "d"(v.y) : "memory");
                }
                if (failNF != NO_FAIL_CASE) atomicMin(&s_failNF[buf], failNF);
                if (failC != NO_FAIL_CASE) atomicMin(&s_fail[buf], failC);
                TICK(8);
                bar_load(NS);
                TICK(9);
                if (st == 0) { // status of this pass (solve_ops.cuh); passes of a model arrive in order
                    merge_load_status(L, model, pass, sOld, s_failNF[buf], s_fail[buf]);
                    s_fail[buf] = NO_FAIL_CASE;
                    s_failNF[buf] = NO_FAIL_CASE;
                }
                TICK(10);
            }
        } else if (i >= 2) {
            // ================= helper warps (4, 8, 12): a share of the reactions of item i-2 =================
            // They live in the solver's sub-partition, so they get work that is almost all shared memory.
            const int it = i - 2, pass = it % cfg.nPass;
            const int model = (int)blockIdx.x + (it / cfg.nPass) * (int)gridDim.x;
            const int c0 = pass * CSP, nc = min(CSP, L.nLC - c0);
            const unsigned buf = (unsigned)(i & 1);
            const unsigned lvB = lvm0 + buf * lvmBytes, xsB = xs0 + buf * xsBytes;
            const int hx = ((warp >> 2) - 1) * 32 + (tid & 31);
            if (L.disp) disp_items(model, c0, nc, xsB, NS + hx);
            bar_stream(NS + NH); // reaction block staged by the streaming warps
            if (L.react)
                react_items(L.react + ((size_t)model * L.nLC + c0) * dof, xsB, lvB + 5u * (unsigned)nB * 8u, nc, hx,
                            ((nc + LC - 1) / LC) * T.nSup * REACT_HELPER_PCT / 100, NH);
            bar_stream(NS + NH);
        }
        __syncthreads();
    }
    if (DBG) {
        long long *d = L.dbgClocks + (size_t)blockIdx.x * 32;
        if (tid == 0) { d[0] = clock64() - tAll; d[1] = tSolve; d[2] = nMy; }
        if (tid == 32) // streaming warp 1
            for (int k = 0; k < 16; k++) d[8 + k] = tPh[k];
    }
}

// ------------------------------------------------------------------------------------------------
// k_solve_general: the same step for models whose factor does not fit the shared-memory resident pipeline (tall
// multi-storey frames: bands of a storey width).  One block per model; only the right-hand sides live in shared
// memory ([padded position][case], cases per pass chosen to fit), the factor is read from global memory / L2 where
// the factor kernel left it.  Four lanes per load case walk the block columns of L directly (BLp / BLi), sharing a column's off blocks:
// forward z_J = V_J acc_J, acc_I -= L_IJ z_J; backward x_J = V_J^T (D_J^-1 z_J - sum_I L_IJ^T x_I).  Recovery and
// status rules are those of k_solve_pipe (model.go:421-491).  Correctness path, not a tuned one.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_THREADS = 128;
constexpr int GEN_SUB = 4; // lanes per load case in the substitutions

bool solve_general_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes)
{
    int cp = nLC < 32 ? nLC : 32;
    if (cp < 1) cp = 1;
    for (; cp >= 1; cp--) {
        const size_t bytes = (size_t)T.nPad * cp * sizeof(double);
        if (bytes + 2048 <= (size_t)smemOptin) {
            casesPerPass = cp;
            smemBytes = bytes;
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(GEN_THREADS) k_solve_general(const DevTopo T, const StaticLaunch L, const int cp)
{
    extern __shared__ double xs[]; // [nPad][cp]
    __shared__ unsigned s_bad, s_failc; // first case of the pass with a non-finite load / a load on a DOF without stiffness
    const int model = blockIdx.x, tid = threadIdx.x, dof = T.dof, nB = T.nB;
    const double *lm = L.Lm + (size_t)model * T.nSlots;
    const double *ksf = L.Ksf + (size_t)model * T.nnzR;
    const double *sec = L.sections + (size_t)model * nB * 3;
    const int nPass = (L.nLC + cp - 1) / cp;
    for (int pass = 0; pass < (nPass < 1 ? 1 : nPass); pass++) {
        const int c0 = pass * cp, nc = min(cp, L.nLC - c0);
        if (tid == 0) {
            s_bad = NO_FAIL_CASE;
            s_failc = NO_FAIL_CASE;
        }
        __syncthreads();
        // ---- load cases -> right-hand sides; checks of model.go:692-696 and 424-437; reactions start as -p ----
        const double *pb = L.loads + ((size_t)model * L.nLC + c0) * dof;
        unsigned bad = NO_FAIL_CASE, failC = NO_FAIL_CASE;
        for (int idx = tid; idx < nc * dof; idx += GEN_THREADS) {
            const int c = idx / dof, g = idx - c * dof;
            const double v = pb[idx];
            if (!(fabs(v) <= 1.7976931348623157e308)) bad = min(bad, (unsigned)(c0 + c));
            const int xr = T.xrow[g];
            if (xr >= 0) xs[xr * cp + c] = v;
            else if (xr == -2 && v != 0.0) failC = min(failC, (unsigned)(c0 + c));
            if (L.react) L.react[((size_t)model * L.nLC + c0) * dof + idx] = T.supIndex[g] >= 0 ? -v : 0.0;
        }
        for (int idx = tid; idx < T.nDummy * nc; idx += GEN_THREADS) {
            const int q = idx / nc, c = idx - q * nc;
            xs[T.dummyPos[q] * cp + c] = 0.0;
        }
        if (bad != NO_FAIL_CASE) atomicMin(&s_bad, bad);
        if (failC != NO_FAIL_CASE) atomicMin(&s_failc, failC);
        __syncthreads();
        // ---- both substitutions: four lanes per load case share the off blocks of a column ----
        for (int cb = 0; cb < nc; cb += GEN_THREADS / GEN_SUB) { // uniform trip count for the whole block
            const int c = cb + (tid >> 2), sub = tid & (GEN_SUB - 1);
            const bool act = c < nc;
            double *x = xs + (act ? c : nc - 1); // idle lanes shadow the last case and never write
            for (int J = 0; J < T.nBlk; J++) {
                const double *h = lm + T.colOff[J];
                const double a0 = x[(3 * J) * cp], a1 = x[(3 * J + 1) * cp], a2 = x[(3 * J + 2) * cp];
                const double z0 = a0, z1 = fma(h[3], a0, a1), z2 = fma(h[5], a1, fma(h[4], a0, a2));
                __syncwarp(); // everybody has read acc_J
                if (act && sub == 0) {
                    x[(3 * J + 1) * cp] = z1;
                    x[(3 * J + 2) * cp] = z2;
                }
                for (int q = T.BLp[J] + sub; q < T.BLp[J + 1]; q += GEN_SUB) { // distinct row blocks: no two lanes on one target
                    const int I = T.BLi[q];
                    const double *l = lm + T.blkSlot[q];
                    if (act) {
#pragma unroll
                        for (int r = 0; r < 3; r++)
                            x[(3 * I + r) * cp] -= fma(l[3 * r + 2], z2, fma(l[3 * r + 1], z1, l[3 * r] * z0));
                    }
                }
                __syncwarp(); // the updates of this column are visible before the next one reads its block
            }
            for (int J = T.nBlk - 1; J >= 0; J--) {
                const double *h = lm + T.colOff[J];
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                for (int q = T.BLp[J] + sub; q < T.BLp[J + 1]; q += GEN_SUB) {
                    const int I = T.BLi[q];
                    const double *l = lm + T.blkSlot[q];
                    const double y0 = x[(3 * I) * cp], y1 = x[(3 * I + 1) * cp], y2 = x[(3 * I + 2) * cp];
                    t0 -= fma(l[6], y2, fma(l[3], y1, l[0] * y0));
                    t1 -= fma(l[7], y2, fma(l[4], y1, l[1] * y0));
                    t2 -= fma(l[8], y2, fma(l[5], y1, l[2] * y0));
                }
#pragma unroll
                for (int o = 1; o < GEN_SUB; o <<= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
                }
                t0 = fma(h[0], x[(3 * J) * cp], t0);
                t1 = fma(h[1], x[(3 * J + 1) * cp], t1);
                t2 = fma(h[2], x[(3 * J + 2) * cp], t2);
                __syncwarp();
                if (act && sub == 0) {
                    x[(3 * J) * cp] = fma(h[4], t2, fma(h[3], t1, t0));
                    x[(3 * J + 1) * cp] = fma(h[5], t2, t1);
                    x[(3 * J + 2) * cp] = t2;
                }
                __syncwarp(); // x_J is in place before column J-1 pulls from it
            }
        }
        __syncthreads();
        // ---- displacements (model.go:446-451) ----
        if (L.disp) {
            double *ob = L.disp + ((size_t)model * L.nLC + c0) * dof;
            for (int idx = tid; idx < nc * dof; idx += GEN_THREADS) {
                const int c = idx / dof, g = idx - c * dof;
                const int xr = T.xrow[g];
                ob[idx] = xr >= 0 ? xs[xr * cp + c] : 0.0;
            }
        }
        // ---- reactions (model.go:470-489): -p (stored above by this thread set's own earlier loop) + K[r, free] d ----
        if (L.react) {
            double *orr = L.react + ((size_t)model * L.nLC + c0) * dof;
            for (int idx = tid; idx < nc * T.nSup; idx += GEN_THREADS) {
                const int c = idx / T.nSup, si = idx - c * T.nSup;
                double acc = 0.0;
                for (int q = T.rptr[si]; q < T.rptr[si + 1]; q++) acc = fma(ksf[q], xs[T.rcol[q] * cp + c], acc);
                const int g = T.supDofs[si];
                orr[(size_t)c * dof + g] = -pb[(size_t)c * dof + g] + acc;
            }
        }
        // ---- beam end forces s = k (T Zo), model.go:455-468 ----
        if (L.forces) {
            double *of = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
            for (int idx = tid; idx < nc * nB; idx += GEN_THREADS) {
                const int c = idx / nB, b = idx - c * nB;
                const BeamGeom g = beam_geom(T, L, model, b);
                const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
                double u[6], o[6];
#pragma unroll
                for (int e = 0; e < 6; e++) {
                    const int r = T.beamRows[b * 6 + e];
                    u[e] = r >= 0 ? xs[r * cp + c] : 0.0;
                }
                const uint32_t pm = T.pinMask[b];
                if (pm == 0) beam_forces_rigid(cf, g.c, g.s, u, o);
                else beam_forces_pinned(cf, pm, g.c, g.s, u, o);
#pragma unroll
                for (int e = 0; e < 6; e++) of[(size_t)idx * 6 + e] = o[e];
            }
        }
        __syncthreads();
        if (tid == 0) merge_load_status(L, model, pass, L.status[model], s_bad, s_failc); // as in k_solve_pipe
        __syncthreads();
    }
}

static cudaError_t launch_solve_general(const DevTopo &T, const StaticLaunch &L, cudaStream_t st, int optin)
{
    int cp = 1;
    size_t smem = 0;
    if (!solve_general_config(T, L.nLC, optin, cp, smem)) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_solve_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_solve_general<<<L.nModels, GEN_THREADS, smem, st>>>(T, L, cp);
    return cudaGetLastError();
}

cudaError_t launch_solve_pipe(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    int dev = 0, optin = 0, nSM = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev);
    SolveCfg cfg;
    size_t smem = 0;
    if (!(std::getenv("HDB_SOLVE_GENERAL") && atoi(std::getenv("HDB_SOLVE_GENERAL")))) { // the bulk-copy pipeline when the shapes allow it
        bool taken = false;
        cudaError_t e = launch_solve_bulk(T, L, st, optin, nSM, taken);
        if (taken || e != cudaSuccess) return e;
    }
    {
        const char *force = std::getenv("HDB_SOLVE_GENERAL"); // tests: the general kernel on a model the pipeline could take
        if ((force && atoi(force)) || !solve_config(T, L.nLC, optin, cfg.casesPerPass, smem)) return launch_solve_general(T, L, st, optin);
    }
    cfg.nPass = (L.nLC + cfg.casesPerPass - 1) / cfg.casesPerPass;
    if (cfg.nPass < 1) cfg.nPass = 1;
    cfg.lvmDoubles = (int)pipe_lvm_doubles(T);
    {
        auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        cfg.vec2 = (T.dof % 2 == 0) && al16(L.loads) && al16(L.disp) && al16(L.react) && al16(L.forces);
    }
    cfg.xsDoubles = T.nBlk * xs_block_doubles(cfg.casesPerPass);
    auto kern = L.dbgClocks ? k_solve_pipe<true> : k_solve_pipe<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int grid = L.nModels < nSM ? L.nModels : nSM;
    if (const char *e = std::getenv("HDB_SOLVE_BLOCKS")) grid = std::max(1, std::min(grid, atoi(e))); // diagnostics: fewer SMs
    if (L.dbgClocks) std::fprintf(stderr, "[hdb] k_solve_pipe: %d blocks, %zu bytes of shared memory, %d cases per pass\n", grid, smem, cfg.casesPerPass);
    kern<<<grid, PIPE_THREADS, smem, st>>>(T, L, cfg);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_uniform_loads: mod.LoadUniform (mod/uniform.go:46-211) for a fixed list of uniformly loaded beams with
// per-model, per-case intensities, followed by a fixed list of loaded nodes, flattened like assemblyNodeLoad
// (model.go:683-698): one thread per (model, case, DOF) adds its contributions in the reference's order, so the
// result is bit-identical to building the LoadNode list on the host.  Codes: 0/1 = +/-(q1*dx*dx)/12,
// 2/3 = +/-(q0*dx*dx)/12 (the reference uses dx for the dy branch too, uniform.go:146-153), 4 = (q0*ly)/2,
// 5 = (q1*lx)/2.
// ------------------------------------------------------------------------------------------------
__global__ void k_uniform_loads(const UniformLaunch U)
{
    const size_t n = (size_t)U.nModels * U.nLC * U.dof;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % U.dof);
        const size_t mc = i / U.dof;
        const double *q = U.q + mc * U.nU * 2;
        double sum = 0.0;
        for (int e = U.uPtr[g]; e < U.uPtr[g + 1]; e++) {
            const int code = U.uIdx[e] & 7, u = U.uIdx[e] >> 3;
            const double *geo = U.uGeo + 4 * u;
            double v;
            if (code < 4) {
                const double dx = geo[0];
                v = __ddiv_rn(__dmul_rn(__dmul_rn(q[2 * u + (code < 2 ? 1 : 0)], dx), dx), 12.0);
                if (code & 1) v = -v;
            } else if (code == 4) v = __ddiv_rn(__dmul_rn(q[2 * u], geo[3]), 2.0);
            else v = __ddiv_rn(__dmul_rn(q[2 * u + 1], geo[2]), 2.0);
            sum = __dadd_rn(sum, v);
        }
        if (U.f) {
            const double *f = U.f + mc * U.nN * 3;
            for (int e = U.nPtr[g]; e < U.nPtr[g + 1]; e++) sum = __dadd_rn(sum, f[U.nIdx[e]]);
        }
        U.loads[i] = sum;
    }
}

cudaError_t launch_uniform_loads(const UniformLaunch &U, cudaStream_t st)
{
    k_uniform_loads<<<148 * 8, 256, 0, st>>>(U);
    return cudaGetLastError();
}

} // namespace hdb
