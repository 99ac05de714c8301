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
        double sum = 0.0;
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
// k_factor: thread per model, interleaved workspace W[slot][ld]
// ------------------------------------------------------------------------------------------------
constexpr int FAC_THREADS = 64;

template <int CNT>
__device__ __forceinline__ void factor_column(double *__restrict__ w, const size_t ld, const int p0,
                                              const int32_t *__restrict__ tg, const double inv)
{
    constexpr int NU = CNT * (CNT + 1) / 2;
    double u[CNT], tv[NU];
    int ti[NU];
#pragma unroll
    for (int a = 0; a < CNT; a++) u[a] = w[(size_t)(p0 + a) * ld];
#pragma unroll
    for (int q = 0; q < NU; q++) ti[q] = __ldg(tg + q);
#pragma unroll
    for (int q = 0; q < NU; q++) tv[q] = w[(size_t)ti[q] * ld];
    int q = 0;
#pragma unroll
    for (int a = 0; a < CNT; a++) {
        const double la = u[a] * inv;
#pragma unroll
        for (int b = 0; b <= a; b++, q++) tv[q] = fma(-la, u[b], tv[q]);
        w[(size_t)(p0 + a) * ld] = la;
    }
#pragma unroll
    for (int q2 = 0; q2 < NU; q2++) w[(size_t)ti[q2] * ld] = tv[q2];
}

__device__ __noinline__ void factor_column_any(double *__restrict__ w, const size_t ld, const int p0, const int cnt,
                                               const int32_t *__restrict__ tg, const double inv)
{
    // generic column: entries stay unscaled until every update of this column is done
    int q = 0;
    for (int a = 0; a < cnt; a++) {
        const double la = w[(size_t)(p0 + a) * ld] * inv;
        for (int b = 0; b <= a; b++, q++) {
            const size_t t = (size_t)__ldg(tg + q) * ld;
            w[t] = fma(-la, w[(size_t)(p0 + b) * ld], w[t]);
        }
    }
    for (int a = 0; a < cnt; a++) w[(size_t)(p0 + a) * ld] *= inv;
}

__global__ void __launch_bounds__(FAC_THREADS) k_factor(const DevTopo T, const StaticLaunch L)
{
    const int model = blockIdx.x * FAC_THREADS + threadIdx.x;
    if (model >= L.nModels) return;
    double *w = L.W + model;
    const size_t ld = (size_t)L.ld;
    bool bad = false;
    const int n = T.nFree, nnzL = T.nnzL;
    for (int j = 0; j < n; j++) {
        const size_t dj = (size_t)(nnzL + j) * ld;
        const double d = w[dj];
        // the reference's LU gives up only when a column has no non-zero pivot candidate (cs_lu: a <= 0):
        // tiny or negative round-off pivots of pin mechanisms pass there and must pass here (TestSplit)
        if (d == 0.0 || !isfinite(d)) bad = true;
        const double inv = __drcp_rn(d);
        const int p0 = __ldg(T.Lp + j);
        const int cnt = __ldg(T.Lp + j + 1) - p0;
        const int32_t *tg = T.utgt + __ldg(T.uptr + j);
        switch (cnt) {
        case 0: break;
        case 1: factor_column<1>(w, ld, p0, tg, inv); break;
        case 2: factor_column<2>(w, ld, p0, tg, inv); break;
        case 3: factor_column<3>(w, ld, p0, tg, inv); break;
        case 4: factor_column<4>(w, ld, p0, tg, inv); break;
        case 5: factor_column<5>(w, ld, p0, tg, inv); break;
        case 6: factor_column<6>(w, ld, p0, tg, inv); break;
        default: factor_column_any(w, ld, p0, cnt, tg, inv); break;
        }
        w[dj] = inv;
    }
    if (bad && L.status[model] == HDB_STATUS_OK) L.status[model] = HDB_STATUS_SINGULAR;
}

cudaError_t launch_factor(const DevTopo &T, const StaticLaunch &L, cudaStream_t st)
{
    k_factor<<<(L.nModels + FAC_THREADS - 1) / FAC_THREADS, FAC_THREADS, 0, st>>>(T, L);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_solve_recover
// ------------------------------------------------------------------------------------------------
constexpr int SOL_THREADS = 256;

struct SolveCfg {
    int casesPerPass; // <= 32
    int cs;           // row stride of xs in doubles (odd)
};

__host__ __device__ inline int odd_stride(int c) { return c | 1; }

bool solve_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes)
{
    int cp = nLC < 32 ? nLC : 32;
    for (; cp >= 1; cp--) {
        const size_t bytes = ((size_t)T.nFree * odd_stride(cp) + 2 * (size_t)T.nnzL + T.nFree + T.nnzR) * sizeof(double);
        if (bytes + 2048 <= (size_t)smemOptin) {
            casesPerPass = cp;
            smemBytes = bytes;
            return true;
        }
        if (cp > 16) cp = 17; // try 16 next, then count down
    }
    return false;
}

__global__ void __launch_bounds__(SOL_THREADS) k_solve_recover(const DevTopo T, const StaticLaunch L, const SolveCfg cfg)
{
    extern __shared__ double smem_d[];
    __shared__ int s_failCase, s_flag;
    const int model = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = T.nFree, nnzL = T.nnzL, dof = T.dof, nB = T.nB;
    const int CS = cfg.cs;
    double *xs = smem_d;                   // [n][CS]
    double *Lv = xs + (size_t)n * CS;      // [nnzL] CSC order
    double *LvRow = Lv + nnzL;             // [nnzL] row order (forward pass)
    double *dinv = LvRow + nnzL;           // [n]
    double *Kr = dinv + n;                 // [nnzR]

    if (tid == 0) {
        s_failCase = 0x7fffffff;
        s_flag = 0;
    }
    // phase 0: factor of this model -> shared memory
    {
        const double *w = L.W + model;
        const size_t ld = (size_t)L.ld;
        for (int s = tid; s < nnzL; s += SOL_THREADS) {
            Lv[s] = w[(size_t)s * ld];
            LvRow[s] = w[(size_t)__ldg(T.rLslot + s) * ld];
        }
        for (int s = tid; s < n; s += SOL_THREADS) dinv[s] = w[(size_t)(nnzL + s) * ld];
        const double *ksf = L.Ksf + (size_t)model * T.nnzR;
        for (int s = tid; s < T.nnzR; s += SOL_THREADS) Kr[s] = ksf[s];
    }
    __syncthreads();

    for (int c0 = 0; c0 < L.nLC; c0 += cfg.casesPerPass) {
        const int nc = min(cfg.casesPerPass, L.nLC - c0);
        const double *p = L.loads + ((size_t)model * L.nLC + c0) * dof;
        // phase 1: load cases -> xs (free rows only); model.go:692-696 and 424-437 checks
        for (int idx = tid; idx < nc * dof; idx += SOL_THREADS) {
            const int c = idx / dof, g = idx - c * dof;
            const double v = p[idx];
            const int xr = __ldg(T.xrow + g);
            if (!isfinite(v)) atomicMax(&s_flag, 1);
            if (xr >= 0) xs[xr * CS + c] = v;
            else if (xr == -2 && v != 0.0) atomicMin(&s_failCase, c0 + c);
        }
        __syncthreads();

        // phase 2: L z = p, y = D^-1 z, L^T x = y ; one lane per load case
        if (tid < nc) {
            const int c = tid;
            for (int i = 0; i < n; i++) {
                double acc = xs[i * CS + c];
                const int q1 = __ldg(T.rLp + i + 1);
                for (int q = __ldg(T.rLp + i); q < q1; q++) acc = fma(-LvRow[q], xs[__ldg(T.rLcol + q) * CS + c], acc);
                xs[i * CS + c] = acc;
            }
            for (int i = n - 1; i >= 0; i--) {
                double acc = xs[i * CS + c] * dinv[i];
                const int q1 = __ldg(T.Lp + i + 1);
                for (int q = __ldg(T.Lp + i); q < q1; q++) acc = fma(-Lv[q], xs[__ldg(T.Li + q) * CS + c], acc);
                xs[i * CS + c] = acc;
            }
        }
        __syncthreads();

        // phase 3a: displacements, model.go:446-451
        if (L.disp) {
            double *out = L.disp + ((size_t)model * L.nLC + c0) * dof;
            for (int idx = tid; idx < nc * dof; idx += SOL_THREADS) {
                const int c = idx / dof, g = idx - c * dof;
                const int xr = __ldg(T.xrow + g);
                out[idx] = xr >= 0 ? xs[xr * CS + c] : 0.0;
            }
        }
        // phase 3b: reactions, model.go:470-489 (columns ascending; eliminated columns carry d = 0)
        if (L.react) {
            double *out = L.react + ((size_t)model * L.nLC + c0) * dof;
            for (int idx = tid; idx < nc * dof; idx += SOL_THREADS) {
                const int c = idx / dof, g = idx - c * dof;
                const int si = __ldg(T.supIndex + g);
                double r = 0.0;
                if (si >= 0) {
                    r = -p[idx];
                    const int q1 = __ldg(T.rptr + si + 1);
                    for (int q = __ldg(T.rptr + si); q < q1; q++) r = fma(Kr[q], xs[__ldg(T.rcol + q) * CS + c], r);
                }
                out[idx] = r;
            }
        }
        // phase 3c: beam end forces s = k (T Zo), model.go:455-468; one thread per beam, k in registers
        if (L.forces) {
            double *out = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
            const double *sec = L.sections + (size_t)model * nB * 3;
            for (int b = tid; b < nB; b += SOL_THREADS) {
                const BeamGeom g = beam_geom(T, b);
                const BeamCoef cf = beam_coef(sec[3 * b], sec[3 * b + 1], sec[3 * b + 2], g.length);
                double k[36];
                stiff_beam2d(cf, T.pinMask[b], k);
                int rows[6];
#pragma unroll
                for (int e = 0; e < 6; e++) rows[e] = __ldg(T.beamRows + b * 6 + e);
                for (int c = 0; c < nc; c++) {
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
        __syncthreads(); // xs is reused by the next pass
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
    k_solve_recover<<<L.nModels, SOL_THREADS, smem, st>>>(T, L, cfg);
    return cudaGetLastError();
}

} // namespace hdb
