// kernels_modal.cu — batched Modal path (reference model.go:719-851), see modal.hpp for the formulation.
//
//   k_modal_loads   builds the s unit load vectors e_{S_j} sqrt(m_j) of every model (m_j = sum of the
//                   masses on DOF S_j divided by Gravity, model.go:780-789); the LinearStatic kernels
//                   then return K^-1 applied to them in displacement layout.
//   k_modal_eigen   one CTA per model: B = M^1/2 (K^-1)_SS M^1/2 in shared memory, Householder
//                   tridiagonalisation and implicit QL with accumulated vectors (the EISPACK tred2 /
//                   tql2 pair, parallelised over rows), Hz = 1/(2 pi sqrt(lambda)) ascending
//                   (model.go:830, 846), shapes = K^-1[:,S] M^1/2 y as one dense FP64 contraction,
//                   scaled to unit 2-norm over the 3*nP entries (what Dgeev returns, model.go:836-841).
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "modal.hpp"

namespace hdb {

constexpr double GRAVITY = 9.80665; // model.go:713
constexpr int EIG_THREADS = 256;
constexpr int QL_RING = 8; // sweeps of rotations in flight between the QL recurrence and their application

int modal_mass_rows(const Symbolic &S, int nMass, const int64_t *massN, std::vector<int32_t> &dofs,
                    std::vector<std::vector<int32_t>> *massOf)
{
    std::vector<std::vector<int32_t>> lists(S.dof);
    for (int q = 0; q < nMass; q++) {
        if (massN[q] < 0 || massN[q] >= S.nP) return HDB_ERR_INPUT;
        for (int dir = 0; dir < 2; dir++) lists[massN[q] * 3 + dir].push_back(q); // model.go:780-788
    }
    dofs.clear();
    if (massOf) massOf->clear();
    for (int g = 0; g < S.dof; g++)
        if (!lists[g].empty() && S.xrow[g] >= 0) { // a mass on an eliminated DOF gives a zero column of H
            dofs.push_back(g);
            if (massOf) massOf->push_back(lists[g]);
        }
    return (int)dofs.size();
}

// ------------------------------------------------------------------------------------------------
__global__ void k_modal_loads(const ModalLaunch M)
{
    const int model = blockIdx.y;
    const int j = blockIdx.x; // mass row
    double m = 0.0;
    for (int q = M.mptr[j]; q < M.mptr[j + 1]; q++) m += M.mass[(size_t)model * M.nMass + M.midx[q]] / GRAVITY;
    const double r = sqrt(m);
    if (threadIdx.x == 0) M.msqrt[(size_t)model * M.s + j] = r;
    double *row = M.unitLoads + ((size_t)model * M.s + j) * M.dof;
    const int gd = M.massDof[j];
    for (int g = threadIdx.x; g < M.dof; g += blockDim.x) row[g] = (g == gd) ? r : 0.0;
}

cudaError_t launch_modal_loads(const ModalLaunch &M, cudaStream_t st)
{
    if (M.s == 0 || M.nModels == 0) return cudaSuccess;
    dim3 grid(M.s, M.nModels);
    k_modal_loads<<<grid, 128, 0, st>>>(M);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// block-wide sum; every thread gets the result. red[] has EIG_THREADS/32 + 1 doubles.
__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads(); // red[] may still be read from the previous call
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < EIG_THREADS / 32; w++) s += red[w];
    return s;
}

// Symmetric eigenproblem resident in shared memory: Householder tridiagonalisation (tred2), accumulation of
// the transformations, implicit QL with accumulated vectors (tql2).  V[n][n] (element (i, j) at i + j * LD)
// holds the symmetric matrix on entry and the eigenvectors (columns) on exit, d the eigenvalues.  Called by
// every thread of a block of EIG_THREADS threads; returns non-zero if the QL iteration did not converge.
// First half: Householder tridiagonalisation and accumulation of the transformations (tred2).  On exit V holds the
// orthogonal Q with Q^T A Q = T, d the diagonal of T and e its subdiagonal shifted as tql2 wants it (e[i] couples i, i+1).
__device__ __forceinline__ void sym_tridiag_smem(double *V, const int LD, const int n, double *d, double *e, double *red)
{
    const int tid = threadIdx.x;
#ifdef HDB_MODAL_CLOCKS
    long long tc0 = clock64();
#define MSTAMP(name) do { __syncthreads(); if (tid == 0 && blockIdx.x == 0) { long long t_ = clock64(); printf("[modal] %-12s %lld\n", name, t_ - tc0); tc0 = t_; } } while (0)
#else
#define MSTAMP(name) do {} while (0)
#endif
    __shared__ double s_h;
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    __syncthreads();
    // ---- Householder reduction to tridiagonal form (tred2), lower triangle ----
    for (int j = tid; j < n; j += EIG_THREADS) d[j] = VV(n - 1, j);
    __syncthreads();
    for (int i = n - 1; i > 0; i--) {
        double part = 0.0;
        for (int k = tid; k < i; k += EIG_THREADS) part += fabs(d[k]);
        const double scale = block_sum(part, red);
        if (scale == 0.0) {
            if (tid == 0) {
                e[i] = d[i - 1];
                s_h = 0.0;
            }
            for (int j = tid; j < i; j += EIG_THREADS) {
                d[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
                VV(j, i) = 0.0;
            }
            __syncthreads();
        } else {
            part = 0.0;
            for (int k = tid; k < i; k += EIG_THREADS) {
                const double t = d[k] / scale;
                d[k] = t;
                part += t * t;
            }
            double h = block_sum(part, red);
            if (tid == 0) {
                const double f = d[i - 1];
                double g = sqrt(h);
                if (f > 0) g = -g;
                e[i] = scale * g;
                h = h - f * g;
                d[i - 1] = f - g;
                s_h = h;
            }
            __syncthreads();
            h = s_h;
            // e = A d (A symmetric, lower triangle stored), V(j, i) = d[j]
            // two threads per row (even / odd k), combined with one shuffle: the row loops are the serial part here
            for (int j0 = 0; j0 < i; j0 += EIG_THREADS / 2) { // uniform trip count: the shuffle needs the whole warp
                const int j = j0 + (tid >> 1), half = tid & 1;
                double g = 0.0;
                if (j < i) {
                    if (half == 0) VV(j, i) = d[j];
                    for (int k = half; k < i; k += 2) g = fma(k <= j ? VV(j, k) : VV(k, j), d[k], g);
                }
                g += __shfl_xor_sync(0xffffffffu, g, 1);
                if (j < i && half == 0) e[j] = g / h;
            }
            __syncthreads();
            part = 0.0;
            for (int j = tid; j < i; j += EIG_THREADS) part += e[j] * d[j];
            const double f = block_sum(part, red);
            const double hh = f / (h + h);
            for (int j = tid; j < i; j += EIG_THREADS) e[j] -= hh * d[j];
            __syncthreads();
            // rank-2 update of the lower triangle: thread k owns row k
            for (int k = tid >> 1; k < i; k += EIG_THREADS / 2) {
                const double ek = e[k], dk = d[k];
                for (int j = tid & 1; j <= k; j += 2) VV(k, j) -= d[j] * ek + e[j] * dk;
            }
            __syncthreads();
            for (int j = tid; j < i; j += EIG_THREADS) {
                d[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
            }
            __syncthreads();
        }
        if (tid == 0) d[i] = s_h;
        __syncthreads();
    }
    MSTAMP("tred2");
    // ---- accumulate the transformations ----
    for (int i = 0; i < n - 1; i++) {
        if (tid == 0) {
            VV(n - 1, i) = VV(i, i);
            VV(i, i) = 1.0;
            s_h = d[i + 1];
        }
        __syncthreads();
        const double h = s_h;
        if (h != 0.0) {
            for (int k = tid; k <= i; k += EIG_THREADS) d[k] = VV(k, i + 1) / h;
            __syncthreads();
            for (int j0 = 0; j0 <= i; j0 += EIG_THREADS / 2) { // two threads per column (even / odd k)
                const int j = j0 + (tid >> 1), half = tid & 1;
                double g = 0.0;
                if (j <= i)
                    for (int k = half; k <= i; k += 2) g = fma(VV(k, i + 1), VV(k, j), g);
                g += __shfl_xor_sync(0xffffffffu, g, 1);
                if (j <= i)
                    for (int k = half; k <= i; k += 2) VV(k, j) -= g * d[k];
            }
            __syncthreads();
        }
        for (int k = tid; k <= i; k += EIG_THREADS) VV(k, i + 1) = 0.0;
        __syncthreads();
    }
    for (int j = tid; j < n; j += EIG_THREADS) {
        d[j] = VV(n - 1, j);
        VV(n - 1, j) = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        VV(n - 1, n - 1) = 1.0;
        e[0] = 0.0;
        // tql2 prologue
        for (int i = 1; i < n; i++) e[i - 1] = e[i];
        e[n - 1] = 0.0;
    }
    __syncthreads();
    MSTAMP("accumulate");
#undef VV
}

// The same reduction for the split pipeline's k_modal_tridiag (a block of TRI_THREADS threads, nothing else resident),
// rearranged around what a block pays for, barriers and strided shared-memory walks:
//  * V holds the FULL symmetric matrix and the rank-2 update keeps both triangles (the two products are added without
//    contraction, so V(k, j) and V(j, k) stay bit-identical): the matrix-vector product and the update walk rows, lanes
//    on consecutive rows, instead of jumping a column stride wherever the triangle runs out;
//  * four threads per row or column (a quarter of the k range each, two shuffles), so the dependent chain of a step is
//    i / 4 long; with LD = 12 (mod 16) both the row walk V(j, k) and the column walk V(k, j) of a warp's 8 x 4 threads
//    fall into 16 distinct 8-byte bank slots, two lanes each: the two wavefronts a 256-byte warp access needs anyway;
//  * three barriers per elimination step instead of twelve: every thread derives g, h and the last component of the
//    Householder vector from the (uniform) squared norm; p = A v / h and v.p come out of one pass (the block sum that
//    follows is also the barrier behind p); the correction q = p - (v.p / 2h) v is applied on the fly in the update; the
//    owners of row i - 1 write the next step's vector (double-buffered) and its squared norm while they update it;
//  * one barrier per accumulation step: the vectors are not divided by h beforehand (g / h multiplies them), and the
//    column that tred2 zeroes after use is instead written analytically (it is the unit vector minus its update).
// The norm is taken without tred2's scaling by sum |v|: B's entries are displacements per unit mass force, far from the
// ends of the double range.  Results agree with tred2's to rounding (the tests compare frequencies with the oracle's
// non-symmetric solve at 1e-9).  On exit as sym_tridiag_smem: V = Q, d = diagonal, e = subdiagonal shifted for tql2.
constexpr int TRI_THREADS = 512;
__device__ __forceinline__ double block_sum_tri(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads(); // red[] may still be read from the previous call
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TRI_THREADS / 32; w++) s += red[w];
    return s;
}
// w: n doubles of work space (the second vector buffer, then the diagonal)
__device__ __forceinline__ void sym_tridiag_full(double *V, const int LD, const int n, double *d, double *e, double *w, double *red)
{
    const int tid = threadIdx.x, q4 = tid & 3, rl = tid >> 2;
    constexpr int RPT = TRI_THREADS / 4; // rows (columns) per trip
    __shared__ double s_sigma;
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    __syncthreads();
    double *cur = d, *nxt = w; // the vector being eliminated lives in [0, i) of one of them; d[i..n) collects the h of the steps
    double part = 0.0;
    for (int j = tid; j < n - 1; j += TRI_THREADS) {
        const double t = VV(n - 1, j);
        cur[j] = t;
        part = fma(t, t, part);
    }
    double sigma = block_sum_tri(part, red);
    for (int i = n - 1; i > 0; i--) {
        if (sigma == 0.0) { // nothing to eliminate in this row (uniform)
            if (tid == 0) {
                e[i] = cur[i - 1];
                d[i] = 0.0;
            }
            part = 0.0;
            for (int j = tid; j < i; j += TRI_THREADS) {
                const double t = VV(i - 1, j);
                nxt[j] = t;
                if (j < i - 1) part = fma(t, t, part);
                VV(i, j) = 0.0;
                VV(j, i) = 0.0;
            }
            sigma = block_sum_tri(part, red);
        } else {
            const double f = cur[i - 1];
            double g = sqrt(sigma);
            if (f > 0) g = -g;
            const double h = sigma - f * g, vlast = f - g;
            // p = A v / h over the leading i x i block, and v . p
            double fpart = 0.0;
            for (int j0 = 0; j0 < i; j0 += RPT) { // uniform trip count: the shuffles need the whole warp
                const int j = j0 + rl;
                double gs = 0.0;
                if (j < i) { // four independent partial sums, their loads issued together: a thread's own chain is the critical path
                    double g1 = 0.0, g2 = 0.0, g3 = 0.0;
                    int k = q4;
                    for (; k + 12 < i; k += 16) {
                        const double a0 = VV(j, k), a1 = VV(j, k + 4), a2 = VV(j, k + 8), a3 = VV(j, k + 12);
                        const double v0 = cur[k], v1 = cur[k + 4], v2 = cur[k + 8], v3 = k + 12 == i - 1 ? vlast : cur[k + 12];
                        gs = fma(a0, v0, gs); g1 = fma(a1, v1, g1); g2 = fma(a2, v2, g2); g3 = fma(a3, v3, g3);
                    }
                    for (; k < i; k += 4) gs = fma(VV(j, k), k == i - 1 ? vlast : cur[k], gs);
                    gs = (gs + g1) + (g2 + g3);
                }
                gs += __shfl_xor_sync(0xffffffffu, gs, 1);
                gs += __shfl_xor_sync(0xffffffffu, gs, 2);
                if (j < i && q4 == 0) {
                    const double pj = gs / h;
                    e[j] = pj;
                    fpart = fma(pj, j == i - 1 ? vlast : cur[j], fpart);
                }
            }
            const double hh = block_sum_tri(fpart, red) / (h + h); // (its barriers: p is complete)
            if (tid == 0) {
                e[i] = g;
                d[i] = h;
            }
            // A -= v q^T + q v^T with q = p - hh v, both triangles; row i - 1 of the result is the next vector
            double spart = 0.0;
            for (int k0 = 0; k0 < i; k0 += RPT) {
                const int k = k0 + rl;
                if (k < i) {
                    const double vk = k == i - 1 ? vlast : cur[k], qk = fma(-hh, vk, e[k]);
                    auto upd = [&](const int j, const double x, const double vj, const double ej) {
                        const double qj = fma(-hh, vj, ej);
                        return __dsub_rn(x, __dadd_rn(__dmul_rn(vj, qk), __dmul_rn(qj, vk)));
                    };
                    int j = q4;
                    for (; j + 12 < i; j += 16) { // four elements at a time, loads first
                        const double x0 = VV(k, j), x1 = VV(k, j + 4), x2 = VV(k, j + 8), x3 = VV(k, j + 12);
                        const double v0 = cur[j], v1 = cur[j + 4], v2 = cur[j + 8], v3 = j + 12 == i - 1 ? vlast : cur[j + 12];
                        const double e0 = e[j], e1 = e[j + 4], e2 = e[j + 8], e3 = e[j + 12];
                        const double n0 = upd(j, x0, v0, e0), n1 = upd(j + 4, x1, v1, e1), n2 = upd(j + 8, x2, v2, e2), n3 = upd(j + 12, x3, v3, e3);
                        VV(k, j) = n0; VV(k, j + 4) = n1; VV(k, j + 8) = n2; VV(k, j + 12) = n3;
                        if (k == i - 1) {
                            nxt[j] = n0; nxt[j + 4] = n1; nxt[j + 8] = n2;
                            spart = fma(n0, n0, spart); spart = fma(n1, n1, spart); spart = fma(n2, n2, spart);
                            if (j + 12 < i - 1) {
                                nxt[j + 12] = n3;
                                spart = fma(n3, n3, spart);
                            }
                        }
                    }
                    for (; j < i; j += 4) {
                        const double nv = upd(j, VV(k, j), j == i - 1 ? vlast : cur[j], e[j]);
                        VV(k, j) = nv;
                        if (k == i - 1 && j < i - 1) {
                            nxt[j] = nv;
                            spart = fma(nv, nv, spart);
                        }
                    }
                }
                if (k0 <= i - 1 && i - 1 < k0 + RPT) { // the trip that holds row i - 1: its four owners sum the next squared norm
                    spart += __shfl_xor_sync(0xffffffffu, spart, 1);
                    spart += __shfl_xor_sync(0xffffffffu, spart, 2);
                    if (k == i - 1 && q4 == 0) s_sigma = spart;
                }
            }
            for (int j = tid; j < i; j += TRI_THREADS) {
                VV(j, i) = j == i - 1 ? vlast : cur[j]; // the Householder vector, for the accumulation below
                VV(i, j) = 0.0;
            }
            __syncthreads();
            sigma = s_sigma;
        }
        double *t = cur;
        cur = nxt;
        nxt = t;
    }
    // ---- accumulate the transformations: Q = H(1) ... H(n-1) built up from the identity, leading block by leading block ----
    __syncthreads();
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] reduction done at %lld\n", clock64());
#endif
    for (int j = tid; j < n; j += TRI_THREADS) {
        w[j] = VV(j, j); // the diagonal of T (neither vector buffer is needed any more)
        VV(j, j) = 1.0;
    }
    __syncthreads();
    for (int i = 0; i < n - 1; i++) {
        const double h = d[i + 1];
        for (int j0 = 0; j0 <= i; j0 += RPT) { // four threads per column; column i is still the unit vector
            const int j = j0 + rl;
            double gs = 0.0;
            if (j < i) {
                double g1 = 0.0, g2 = 0.0, g3 = 0.0;
                int k = q4;
                for (; k + 12 <= i; k += 16) {
                    const double a0 = VV(k, i + 1), a1 = VV(k + 4, i + 1), a2 = VV(k + 8, i + 1), a3 = VV(k + 12, i + 1);
                    const double b0 = VV(k, j), b1 = VV(k + 4, j), b2 = VV(k + 8, j), b3 = VV(k + 12, j);
                    gs = fma(a0, b0, gs); g1 = fma(a1, b1, g1); g2 = fma(a2, b2, g2); g3 = fma(a3, b3, g3);
                }
                for (; k <= i; k += 4) gs = fma(VV(k, i + 1), VV(k, j), gs);
                gs = (gs + g1) + (g2 + g3);
            }
            gs += __shfl_xor_sync(0xffffffffu, gs, 1);
            gs += __shfl_xor_sync(0xffffffffu, gs, 2);
            if (j < i) {
                const double gh = h != 0.0 ? gs / h : 0.0;
                int k = q4;
                for (; k + 12 <= i; k += 16) {
                    const double a0 = VV(k, i + 1), a1 = VV(k + 4, i + 1), a2 = VV(k + 8, i + 1), a3 = VV(k + 12, i + 1);
                    const double b0 = VV(k, j), b1 = VV(k + 4, j), b2 = VV(k + 8, j), b3 = VV(k + 12, j);
                    VV(k, j) = fma(-gh, a0, b0); VV(k + 4, j) = fma(-gh, a1, b1); VV(k + 8, j) = fma(-gh, a2, b2); VV(k + 12, j) = fma(-gh, a3, b3);
                }
                for (; k <= i; k += 4) VV(k, j) = fma(-gh, VV(k, i + 1), VV(k, j));
            } else if (j == i) {
                const double gh = h != 0.0 ? VV(i, i + 1) / h : 0.0;
                for (int k = q4; k <= i; k += 4) VV(k, i) = fma(-gh, VV(k, i + 1), k == i ? 1.0 : 0.0); // (overwrites the vector of the step before)
            }
        }
        __syncthreads();
    }
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] accumulation done at %lld\n", clock64());
#endif
    for (int k = tid; k < n - 1; k += TRI_THREADS) VV(k, n - 1) = 0.0; // the last vector's column: the unit vector
    for (int j = tid; j < n; j += TRI_THREADS) d[j] = w[j];
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < n; i++) e[i - 1] = e[i]; // tql2 prologue
        e[n - 1] = 0.0;
    }
    __syncthreads();
#undef VV
}

// Second half: implicit QL with accumulated vectors (tql2) on the tridiagonal (d, e), rotating the columns of V.
__device__ __forceinline__ int sym_ql_smem(double *V, const int LD, const int n, double *d, double *e, double *cc, double *ss, double *red)
{
    const int tid = threadIdx.x;
#ifdef HDB_MODAL_CLOCKS
    long long tc0 = clock64();
#endif
    __shared__ int s_fail;
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    if (tid == 0) s_fail = 0;
    __syncthreads();
    // ---- implicit QL with accumulated vectors (tql2), producer / consumer.  The scalar recurrence of a sweep
    //      depends only on d and e, never on the vectors: thread 0 runs all sweeps back to back and leaves the
    //      rotations of each in a ring of QL_RING slots; warps 1.. apply them to the rows of V as they appear.
    //      (Before, the other 255 threads waited at a barrier for every sweep: 71 % barrier stalls at n = 130.)
    __shared__ volatile int s_prod, s_cons, s_done;
    __shared__ int s_lm[2 * QL_RING];
    if (tid == 0) { s_prod = 0; s_cons = 0; s_done = 0; }
    __syncthreads();
    const double eps = 2.220446049250313e-16;
    if (tid < 32) {
        // warp 0: lane 0 runs the recurrence; the whole warp does the O(n) shift d[i] -= h of every sweep, which
        // would otherwise sit on the chain as ~n dependent shared-memory round trips
        const int lane = tid;
        double f = 0.0, tst1 = 0.0;
        int sweep = 0;
        for (int l = 0; l < n; l++) {
            int m = l;
            if (lane == 0) {
                tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
                while (m < n) {
                    if (fabs(e[m]) <= eps * tst1) break;
                    m++;
                }
                if (m >= n) m = n - 1;
            }
            m = __shfl_sync(0xffffffffu, m, 0);
            if (m > l) {
                int iter = 0, more = 0;
                do {
                    iter++;
                    double g = 0.0, p = 0.0, r = 0.0, dl1 = 0.0, h = 0.0;
                    double *rc = cc + (size_t)(sweep % QL_RING) * 2 * n, *rs = rc + n;
                    if (lane == 0) {
                        while (sweep - s_cons >= QL_RING) __nanosleep(100); // ring full: wait for the consumers
                        g = d[l];
                        p = (d[l + 1] - g) / (2.0 * e[l]);
                        r = hypot(p, 1.0);
                        if (p < 0) r = -r;
                        d[l] = e[l] / (p + r);
                        d[l + 1] = e[l] * (p + r);
                        dl1 = d[l + 1];
                        h = g - d[l];
                    }
                    h = __shfl_sync(0xffffffffu, h, 0);
                    for (int i = l + 2 + lane; i < n; i += 32) d[i] -= h;
                    __syncwarp();
                    if (lane == 0) {
                        f += h;
                        p = d[m];
                        double c = 1.0, c2 = c, c3 = c;
                        const double el1 = e[l + 1];
                        double s = 0.0, s2 = 0.0;
                        // explicit shared-window addresses: generic pointers into dynamic shared memory make the compiler
                        // re-derive the window base inside the loop (S2UR on the XU pipe), on the critical chain here
                        const unsigned dA = (unsigned)__cvta_generic_to_shared(d), eA = (unsigned)__cvta_generic_to_shared(e);
                        const unsigned cA = (unsigned)__cvta_generic_to_shared(rc), sA = (unsigned)__cvta_generic_to_shared(rs);
                        double eiN, diN; // e[i], d[i] of the next step, loaded one step ahead
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(eiN) : "r"(eA + 8u * (unsigned)(m - 1)));
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(diN) : "r"(dA + 8u * (unsigned)(m - 1)));
                        for (int i = m - 1; i >= l; i--) {
                            c3 = c2;
                            c2 = c;
                            s2 = s;
                            const double ei = eiN, di = diN;
                            if (i > l) {
                                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(eiN) : "r"(eA + 8u * (unsigned)(i - 1)));
                                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(diN) : "r"(dA + 8u * (unsigned)(i - 1)));
                            }
                            g = c * ei;
                            h = c * p;
                            // the tridiagonal entries are O(lambda): no overflow / underflow guard needed.  One
                            // reciprocal square root instead of a square root followed by a division (each a
                            // ~100-cycle dependent sequence on this chain), and the next p computed off that chain.
                            const double q2 = fma(p, p, ei * ei);
                            const double tp = fma(p, di, -ei * g); // = r * (c d[i] - s g)
                            const double rinv = rsqrt(q2);
                            r = q2 * rinv;
                            asm volatile("st.shared.f64 [%0], %1;" ::"r"(eA + 8u * (unsigned)(i + 1)), "d"(s * r) : "memory");
                            s = ei * rinv;
                            c = p * rinv;
                            p = tp * rinv;
                            asm volatile("st.shared.f64 [%0], %1;" ::"r"(dA + 8u * (unsigned)(i + 1)), "d"(h + s * (c * g + s * di)) : "memory");
                            asm volatile("st.shared.f64 [%0], %1;" ::"r"(cA + 8u * (unsigned)i), "d"(c) : "memory");
                            asm volatile("st.shared.f64 [%0], %1;" ::"r"(sA + 8u * (unsigned)i), "d"(s) : "memory");
                        }
                        p = -s * s2 * c3 * el1 * e[l] / dl1;
                        e[l] = s * p;
                        d[l] = c * p;
                        more = fabs(e[l]) > eps * tst1;
                        if (iter > 60) {
                            more = 0;
                            s_fail = 1;
                        }
                        s_lm[2 * (sweep % QL_RING)] = l;
                        s_lm[2 * (sweep % QL_RING) + 1] = m;
                        __threadfence_block();
                        s_prod = sweep + 1;
                    }
                    sweep++;
                    more = __shfl_sync(0xffffffffu, more, 0);
                } while (more);
            }
            if (lane == 0) {
                d[l] = d[l] + f;
                e[l] = 0.0;
            }
            __syncwarp();
        }
        if (lane == 0) {
            __threadfence_block();
            s_done = 1;
        }
    } else if ((tid >> 5) & 3) { // the producer keeps SM sub-partition 0 (warps 0, 4, ...) to itself
        const int w_ = tid >> 5, ct = (w_ - (w_ >> 2) - 1) * 32 + (tid & 31), NC = EIG_THREADS / 32 / 4 * 3 * 32;
        for (int k = 0;; k++) {
            while (s_prod <= k && !s_done) __nanosleep(200); // back off: the spinning reads share the shared-memory pipe with the producer
            if (s_prod <= k) break; // producer finished and nothing left (s_prod is final once s_done is visible)
            const double *rc = cc + (size_t)(k % QL_RING) * 2 * n, *rs = rc + n;
            const int l = s_lm[2 * (k % QL_RING)], m = s_lm[2 * (k % QL_RING) + 1];
            for (int row = ct; row < n; row += NC) {
                double hk = VV(row, m);
                for (int i = m - 1; i >= l; i--) {
                    const double c = rc[i], sn = rs[i], vi = VV(row, i);
                    VV(row, i + 1) = fma(sn, vi, c * hk);
                    hk = fma(c, vi, -sn * hk);
                }
                VV(row, l) = hk;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory"); // every consumer is done with this slot
            if (ct == 0) s_cons = k + 1;
        }
    }

    __syncthreads();
    MSTAMP("ql");
    return s_fail;
#undef VV
}

// Symmetric eigenproblem resident in shared memory: V[n][n] (element (i, j) at i + j * LD) holds the symmetric matrix on
// entry and the eigenvectors (columns) on exit, d the eigenvalues.  Called by every thread of a block of EIG_THREADS
// threads; returns non-zero if the QL iteration did not converge.
__device__ __forceinline__ int sym_eig_smem(double *V, const int LD, const int n, double *d, double *e, double *cc, double *ss, double *red)
{
    sym_tridiag_smem(V, LD, n, d, e, red);
    return sym_ql_smem(V, LD, n, d, e, cc, ss, red);
}

__global__ void __launch_bounds__(EIG_THREADS, 1) k_modal_eigen(const ModalLaunch M, const int LD)
{
    extern __shared__ __align__(16) double sm[];
    const int n = M.s, tid = threadIdx.x, model = blockIdx.x;
    double *V = sm;                 // [n][n] element (i, j) at i + j * LD
    double *d = V + (size_t)LD * n; // [n]
    double *e = d + n;              // [n]
    double *cc = e + n;             // [QL_RING][2n] rotations of the QL sweeps in flight
    double *ss = cc + n;            // (cc is the ring of QL rotations: QL_RING slots of [cosines n | sines n])
    double *red = cc + (size_t)2 * QL_RING * n; // [EIG_THREADS/32 + 8]
    int *order = reinterpret_cast<int *>(red + EIG_THREADS / 32 + 8); // [n]
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    const double *sol = M.sol + (size_t)model * n * M.dof;
    const double *ms = M.msqrt + (size_t)model * n;

    // ---- B = M^1/2 (K^-1)_SS M^1/2, symmetrised ----
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
        const int i = idx % n, j = idx / n;
        const double bij = ms[i] * sol[(size_t)j * M.dof + M.massDof[i]];
        const double bji = ms[j] * sol[(size_t)i * M.dof + M.massDof[j]];
        VV(i, j) = 0.5 * (bij + bji);
    }
    __syncthreads();

#ifdef HDB_MODAL_CLOCKS
    long long tk0 = clock64();
    if (tid == 0 && blockIdx.x == 0) printf("[modal] build B\n");
#endif
    const int failed = sym_eig_smem(V, LD, n, d, e, cc, ss, red);
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[modal] eig total %lld\n", clock64() - tk0);
    tk0 = clock64();
#endif

    // ---- order: ascending Hz = descending lambda (stable on ties) ----
    for (int j = tid; j < n; j += EIG_THREADS) {
        const double lj = fabs(d[j]);
        int rank = 0;
        for (int i = 0; i < n; i++) {
            const double li = fabs(d[i]);
            if (li > lj || (li == lj && i < j)) rank++;
        }
        order[rank] = j;
    }
    __syncthreads();
    // A zero eigenvalue is not a mode: the reference skips it (model.go:818-834).  A mass row whose masses sum to exactly 0 N
    // is a zero column of K^-1 M there (an exact zero after Dgeev's balancing); here it is a zero row and column of B whose
    // eigenvalue comes out as round-off (the reflectors mix it in), one of the smallest in magnitude: B is positive
    // semi-definite, real modes have lambda = 1 / omega^2 far above round-off.  So as many of the last (smallest |lambda|)
    // entries as there are such rows are dropped, and any exact zero after them.
    __shared__ int s_nz;
    if (tid == 0) {
        int nz = n;
        const double *msq = M.msqrt + (size_t)model * n;
        for (int i = 0; i < n; i++)
            if (msq[i] == 0.0) nz--;
        while (nz > 0 && d[order[nz - 1]] == 0.0) nz--;
        s_nz = nz;
    }
    __syncthreads();
    const int nOut = min(s_nz, M.maxModes);
    for (int r = tid; r < nOut; r += EIG_THREADS) {
        const double val = fabs(d[order[r]]);
        M.hz[(size_t)model * M.maxModes + r] = 1. / (sqrt(val) * 2.0 * M_PI); // model.go:830
    }
    if (tid == 0) {
        M.nModes[model] = nOut;
        if (failed && M.status[model] == HDB_STATUS_OK) M.status[model] = HDB_STATUS_EIGEN;
    }
    if (!M.shapes) return;

    // ---- shapes: raw[r][g] = sum_j sol[j][g] * y_r[j]; unit 2-norm over all 3*nP entries ----
    constexpr int MT = 8;
    double *out = M.shapes + (size_t)model * M.maxModes * M.dof;
    for (int r0 = 0; r0 < nOut; r0 += MT) {
        int col[MT];
#pragma unroll
        for (int q = 0; q < MT; q++) col[q] = order[min(r0 + q, nOut - 1)];
        for (int g = tid; g < M.dof; g += EIG_THREADS) {
            double acc[MT];
#pragma unroll
            for (int q = 0; q < MT; q++) acc[q] = 0.0;
            for (int j = 0; j < n; j++) {
                const double a = sol[(size_t)j * M.dof + g];
#pragma unroll
                for (int q = 0; q < MT; q++) acc[q] = fma(a, VV(j, col[q]), acc[q]);
            }
#pragma unroll
            for (int q = 0; q < MT; q++)
                if (r0 + q < nOut) out[(size_t)(r0 + q) * M.dof + g] = acc[q];
        }
    }
    __syncthreads();
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[modal] shapes gemm %lld\n", clock64() - tk0);
    tk0 = clock64();
#endif
    // unit 2-norm: a warp per mode sums the squares in a fixed order (bit-reproducible run to run), then scales
    for (int r = tid >> 5; r < nOut; r += EIG_THREADS / 32) {
        double *row = out + (size_t)r * M.dof;
        double v = 0.0;
        for (int g = tid & 31; g < M.dof; g += 32) v = fma(row[g], row[g], v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const double nr = sqrt(v);
        for (int g = tid & 31; g < M.dof; g += 32) row[g] = row[g] / nr;
    }
#ifdef HDB_MODAL_CLOCKS
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) printf("[modal] normalise %lld\n", clock64() - tk0);
#endif
#undef VV
}

static size_t eigen_smem(int n, int &LD)
{
    LD = n | 1;
    if ((LD & 3) == 1) LD += 2; // LD = 3 mod 4 keeps column strides off the same banks
    return ((size_t)LD * n + 2 * (size_t)n + 2 * (size_t)QL_RING * n + EIG_THREADS / 32 + 8) * sizeof(double) + (size_t)n * sizeof(int) + 64;
}

int modal_max_s(int smemOptin)
{
    int n = 1, LD;
    while (eigen_smem(n + 1, LD) + 1024 <= (size_t)smemOptin) n++;
    return n;
}

cudaError_t launch_modal_eigen(const ModalLaunch &M, int smemOptin, cudaStream_t st)
{
    if (M.nModels == 0) return cudaSuccess;
    int LD;
    const size_t smem = eigen_smem(M.s > 0 ? M.s : 1, LD);
    if (smem + 1024 > (size_t)smemOptin) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_modal_eigen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_modal_eigen<<<M.nModels, EIG_THREADS, smem, st>>>(M, LD);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The split Modal pipeline (what hdb_modal_batch runs when its scratch fits): three kernels instead of k_modal_eigen.
//
//   k_modal_tridiag    CTA per model: B = M^1/2 (K^-1)_SS M^1/2, tred2, Q accumulated; Q, d, e -> global scratch.
//   k_modal_ql_chain   the scalar QL recurrence of every model, ONE WARP per model and all models of the batch at the
//                      same time.  The recurrence is a chain of dependent square roots (about 220 cycles a rotation,
//                      3 M cycles a model at n = 130) that nothing inside a model can shorten; inside k_modal_eigen it
//                      kept a whole SM (the shared-memory resident Q) waiting.  Here the chains of seven models share an
//                      SM, and their rotations are logged to global memory: (cos, sin) pairs, one header per sweep.
//   k_modal_finish     CTA per model: Q back into shared memory, the logged rotations applied row by row (a row is
//                      independent of every other row: no barrier at all), frequencies, then the shapes as one
//                      FP64 tensor-core contraction (mma.sync m8n8k4, SASS DMMA): out[r][g] = sum_j y_r[j] sol[j][g]
//                      with the solution rows streamed through shared memory exactly once (cp.async, double buffered),
//                      squared norms accumulated on the way, and a last in-place scaling pass over the L2-hot result.
// ------------------------------------------------------------------------------------------------
constexpr int CH_WARPS = 4; // models per block of the chain kernel

__global__ void __launch_bounds__(TRI_THREADS, 1) k_modal_tridiag(const ModalLaunch M, const int LD, const int LDT)
{
    extern __shared__ __align__(16) double sm[];
    const int n = M.s, tid = threadIdx.x, model = blockIdx.x;
    double *V = sm; // [n][n], element (i, j) at i + j * LDT
    double *d = V + (size_t)LDT * n, *e = d + n, *w = e + n, *red = w + n;
#define VV(i, j) V[(i) + (size_t)(j) * LDT]
    const double *sol = M.sol + (size_t)model * n * M.dof;
    const double *ms = M.msqrt + (size_t)model * n;
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] start at %lld\n", clock64());
#endif
    for (int idx = tid; idx < n * n; idx += TRI_THREADS) { // B, symmetrised (both triangles: bij + bji commutes)
        const int i = idx % n, j = idx / n;
        const double bij = ms[i] * sol[(size_t)j * M.dof + M.massDof[i]];
        const double bji = ms[j] * sol[(size_t)i * M.dof + M.massDof[j]];
        VV(i, j) = i <= j ? 0.5 * (bij + bji) : 0.5 * (bji + bij);
    }
    __syncthreads();
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] B built at %lld\n", clock64());
#endif
    sym_tridiag_full(V, LDT, n, d, e, w, red);
#ifdef HDB_MODAL_CLOCKS
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] tridiagonal form at %lld\n", clock64());
#endif
    double *q = M.qScratch + (size_t)model * n * LD; // the finish kernel's layout: column stride LD
    for (int idx = tid; idx < n * n; idx += TRI_THREADS) {
        const int i = idx % n, j = idx / n;
        q[i + (size_t)j * LD] = VV(i, j);
    }
#ifdef HDB_MODAL_CLOCKS
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) printf("[tridiag] Q stored at %lld\n", clock64());
#endif
    double *de = M.deScratch + (size_t)model * 2 * n;
    for (int j = tid; j < n; j += TRI_THREADS) {
        de[j] = d[j];
        de[n + j] = e[j];
    }
#undef VV
}

__global__ void __launch_bounds__(CH_WARPS * 32) k_modal_ql_chain(const ModalLaunch M)
{
    extern __shared__ __align__(16) double sm[]; // [CH_WARPS][2n]
    const int n = M.s, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int model = blockIdx.x * CH_WARPS + w;
    if (model >= M.nModels) return;
    double *d = sm + (size_t)w * 2 * n, *e = d + n;
    const double *de = M.deScratch + (size_t)model * 2 * n;
    for (int j = lane; j < 2 * n; j += 32) d[j] = de[j];
    __syncwarp();
    double2 *rot = M.rotLog + (size_t)model * M.rotCap;
    int4 *hdr = M.rotHdr + (size_t)model * M.rotMaxSweeps;
    const double eps = 2.220446049250313e-16;
    double f = 0.0, tst1 = 0.0;
    int sweep = 0, off = 0, fail = 0;
    for (int l = 0; l < n && !fail; l++) {
        int m = l;
        if (lane == 0) {
            tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
            while (m < n) {
                if (fabs(e[m]) <= eps * tst1) break;
                m++;
            }
            if (m >= n) m = n - 1;
        }
        m = __shfl_sync(0xffffffffu, m, 0);
        if (m > l) {
            int iter = 0, more = 0;
            do {
                iter++;
                if (sweep >= M.rotMaxSweeps || off + (m - l) > M.rotCap) { fail = 1; break; } // uniform: the log is full
                double g = 0.0, p = 0.0, r = 0.0, dl1 = 0.0, h = 0.0;
                if (lane == 0) {
                    g = d[l];
                    p = (d[l + 1] - g) / (2.0 * e[l]);
                    r = hypot(p, 1.0);
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    dl1 = d[l + 1];
                    h = g - d[l];
                }
                h = __shfl_sync(0xffffffffu, h, 0);
                for (int i = l + 2 + lane; i < n; i += 32) d[i] -= h;
                __syncwarp();
                if (lane == 0) {
                    f += h;
                    p = d[m];
                    double c = 1.0, c2 = c, c3 = c;
                    const double el1 = e[l + 1];
                    double s = 0.0, s2 = 0.0;
                    const unsigned dA = (unsigned)__cvta_generic_to_shared(d), eA = (unsigned)__cvta_generic_to_shared(e);
                    double2 *rs = rot + off + (m - 1); // rotation of step i at rs[-i]: the log is written in the order it is applied
                    double eiN, diN; // e[i], d[i] of the next step, loaded one step ahead
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(eiN) : "r"(eA + 8u * (unsigned)(m - 1)));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(diN) : "r"(dA + 8u * (unsigned)(m - 1)));
                    for (int i = m - 1; i >= l; i--) {
                        c3 = c2;
                        c2 = c;
                        s2 = s;
                        const double ei = eiN, di = diN;
                        if (i > l) {
                            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(eiN) : "r"(eA + 8u * (unsigned)(i - 1)));
                            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(diN) : "r"(dA + 8u * (unsigned)(i - 1)));
                        }
                        g = c * ei;
                        h = c * p;
                        // one reciprocal square root instead of a square root followed by a division, and the next p
                        // computed off that chain (the tridiagonal entries are O(lambda): no overflow guard needed)
                        const double q2 = fma(p, p, ei * ei);
                        const double tp = fma(p, di, -ei * g); // = r * (c d[i] - s g)
                        const double rinv = rsqrt(q2);
                        r = q2 * rinv;
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(eA + 8u * (unsigned)(i + 1)), "d"(s * r) : "memory");
                        s = ei * rinv;
                        c = p * rinv;
                        p = tp * rinv;
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(dA + 8u * (unsigned)(i + 1)), "d"(h + s * (c * g + s * di)) : "memory");
                        __stcg(rs - i, make_double2(c, s));
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                    more = fabs(e[l]) > eps * tst1;
                    if (iter > 60) {
                        more = 0;
                        fail = 1;
                    }
                    hdr[sweep] = make_int4(l, m, off, 0);
                }
                off += m - l;
                sweep++;
                more = __shfl_sync(0xffffffffu, more, 0);
                fail = __shfl_sync(0xffffffffu, fail, 0);
            } while (more);
        }
        if (lane == 0) {
            d[l] = d[l] + f;
            e[l] = 0.0;
        }
        __syncwarp();
    }
    double *dOut = M.deScratch + (size_t)model * 2 * n; // eigenvalues replace the diagonal
    for (int j = lane; j < n; j += 32) dOut[j] = d[j];
    if (lane == 0) M.rotSweeps[model] = fail ? -1 : sweep;
}

__device__ __forceinline__ void cp_async16(unsigned smem, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int SH_NT = 32;     // columns (DOFs) of one streamed tile of the solution rows
constexpr int SH_LDB = 36;    // its leading dimension in shared memory: 4 mod 16, fragment loads hit distinct banks
constexpr int SH_WARPS = EIG_THREADS / 32;

__global__ void __launch_bounds__(EIG_THREADS, 1) k_modal_finish(const ModalLaunch M, const int LD)
{
    extern __shared__ __align__(16) double sm[];
    const int n = M.s, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, model = blockIdx.x;
#ifdef HDB_MODAL_CLOCKS
    long long tf0 = clock64();
#define FSTAMP(name) do { __syncthreads(); if (tid == 0 && blockIdx.x == 0) { long long t_ = clock64(); printf("[finish] %-12s %lld\n", name, t_ - tf0); tf0 = t_; } } while (0)
#else
#define FSTAMP(name) do {} while (0)
#endif
    const int KP = (n + 3) & ~3; // contraction length padded to the k = 4 of the tensor-core shape
    double *V = sm;                                  // [KP rows used][n] element (i, j) at i + j * LD; rows n..KP-1 are zero
    double *d = V + (size_t)LD * n;                  // [n] eigenvalues
    double *nrm = d + n;                             // [n] squared norms of the raw shapes
    double *tile = nrm + n + (((size_t)LD * n) & 1);  // [2][KP][SH_LDB] streamed solution rows, 16-byte aligned for cp.async (kept as
                                                     // pointer arithmetic on sm: a cast through an integer would turn every access generic)
    int *order = reinterpret_cast<int *>(tile + (size_t)2 * KP * SH_LDB); // [n]
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    const double *sol = M.sol + (size_t)model * n * M.dof;
    {
        const double *q = M.qScratch + (size_t)model * n * LD;
        for (int idx = tid; idx < n * LD; idx += EIG_THREADS) V[idx] = q[idx];
        const double *de = M.deScratch + (size_t)model * 2 * n;
        for (int j = tid; j < n; j += EIG_THREADS) {
            d[j] = de[j];
            nrm[j] = 0.0;
        }
    }
    __syncthreads();
    FSTAMP("load Q");
    // ---- the logged QL rotations, applied to the rows of Q: thread = row.  The log is consumed front to back; it streams
    //      through shared memory in chunks (cp.async by all threads, double buffered, in the region the shape tiles use
    //      later), the sweep headers sit in shared memory too.  One barrier per chunk, none per sweep. ----
    const int nSweeps = M.rotSweeps[model];
    {
        const double2 *rot = M.rotLog + (size_t)model * M.rotCap;
        const int4 *hdrG = M.rotHdr + (size_t)model * M.rotMaxSweeps;
        int4 *hdr = reinterpret_cast<int4 *>(tile);                                  // [nSweeps] (at most rotMaxSweeps)
        const int hdrDoubles = 2 * M.rotMaxSweeps;
        double2 *ring = reinterpret_cast<double2 *>(tile + hdrDoubles);              // [2][RC]
        const int RC = (int)((((size_t)2 * KP * SH_LDB - hdrDoubles) / 4) & ~(size_t)3); // rotations per chunk
        const unsigned ringA = (unsigned)__cvta_generic_to_shared(ring);
        const int ns = nSweeps > 0 ? nSweeps : 0;
        for (int k = tid; k < ns; k += EIG_THREADS) hdr[k] = hdrG[k];
        __syncthreads();
        const int total = ns > 0 ? hdr[ns - 1].z + (hdr[ns - 1].y - hdr[ns - 1].x) : 0;
        const int nChunks = (total + RC - 1) / RC;
        auto fetch = [&](int c) {
            const int p0 = c * RC, cnt = min(RC, total - p0);
            for (int q = tid; q < cnt; q += EIG_THREADS) cp_async16(ringA + (unsigned)(((c & 1) * RC + q) * 16), rot + p0 + q);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (nChunks > 0) fetch(0);
        // walk state of this thread's row
        const bool mine = tid < n;
        double *vr = V + (mine ? tid : 0);
        int k = 0, l = 0, i = -1;
        double hk = 0.0;
        if (ns > 0) {
            const int4 h = hdr[0];
            l = h.x;
            i = h.y - 1;
            hk = vr[(size_t)h.y * LD];
        }
        for (int c = 0; c < nChunks; c++) {
            if (c + 1 < nChunks) {
                fetch(c + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads(); // chunk c has landed for everybody
            if (mine) {
                const double2 *rc = ring + (size_t)(c & 1) * RC;
                int left = min(RC, total - c * RC);
                while (left > 0) {
                    int run = min(left, i - l + 1);
                    left -= run;
                    for (; run >= 4; run -= 4, i -= 4, rc += 4) { // four steps at a time: their loads are issued before the first dependent store
                        const double2 r0 = rc[0], r1 = rc[1], r2 = rc[2], r3 = rc[3];
                        const double v0 = vr[(size_t)i * LD], v1 = vr[(size_t)(i - 1) * LD], v2 = vr[(size_t)(i - 2) * LD], v3 = vr[(size_t)(i - 3) * LD];
                        // one dependent multiply-add per step: the products with the row's old entries do not wait for hk
                        const double sv0 = r0.y * v0, cv0 = r0.x * v0, sv1 = r1.y * v1, cv1 = r1.x * v1;
                        const double sv2 = r2.y * v2, cv2 = r2.x * v2, sv3 = r3.y * v3, cv3 = r3.x * v3;
                        const double h1 = fma(-r0.y, hk, cv0);
                        vr[(size_t)(i + 1) * LD] = fma(r0.x, hk, sv0);
                        const double h2 = fma(-r1.y, h1, cv1);
                        vr[(size_t)i * LD] = fma(r1.x, h1, sv1);
                        const double h3 = fma(-r2.y, h2, cv2);
                        vr[(size_t)(i - 1) * LD] = fma(r2.x, h2, sv2);
                        vr[(size_t)(i - 2) * LD] = fma(r3.x, h3, sv3);
                        hk = fma(-r3.y, h3, cv3);
                    }
                    for (; run > 0; run--, i--, rc++) {
                        const double2 r0 = rc[0];
                        const double v0 = vr[(size_t)i * LD];
                        vr[(size_t)(i + 1) * LD] = fma(r0.x, hk, r0.y * v0);
                        hk = fma(-r0.y, hk, r0.x * v0);
                    }
                    if (i < l) { // sweep finished: next one
                        vr[(size_t)l * LD] = hk;
                        if (++k < ns) {
                            const int4 h = hdr[k];
                            l = h.x;
                            i = h.y - 1;
                            hk = vr[(size_t)h.y * LD];
                        }
                    }
                }
            }
            __syncthreads(); // everybody is done with this buffer before chunk c + 2 overwrites it
        }
    }
    __syncthreads();
    FSTAMP("rotations");
    // ---- order: ascending Hz = descending lambda (stable on ties) ----
    for (int j = tid; j < n; j += EIG_THREADS) {
        const double lj = fabs(d[j]);
        int rank = 0;
        for (int i = 0; i < n; i++) {
            const double li = fabs(d[i]);
            if (li > lj || (li == lj && i < j)) rank++;
        }
        order[rank] = j;
    }
    __syncthreads();
    // A zero eigenvalue is not a mode: the reference skips it (model.go:818-834).  A mass row whose masses sum to exactly 0 N
    // is a zero column of K^-1 M there (an exact zero after Dgeev's balancing); here it is a zero row and column of B whose
    // eigenvalue comes out as round-off (the reflectors mix it in), one of the smallest in magnitude: B is positive
    // semi-definite, real modes have lambda = 1 / omega^2 far above round-off.  So as many of the last (smallest |lambda|)
    // entries as there are such rows are dropped, and any exact zero after them.
    __shared__ int s_nz;
    if (tid == 0) {
        int nz = n;
        const double *msq = M.msqrt + (size_t)model * n;
        for (int i = 0; i < n; i++)
            if (msq[i] == 0.0) nz--;
        while (nz > 0 && d[order[nz - 1]] == 0.0) nz--;
        s_nz = nz;
    }
    __syncthreads();
    const int nOut = min(s_nz, M.maxModes);
    for (int r = tid; r < nOut; r += EIG_THREADS) {
        const double val = fabs(d[order[r]]);
        M.hz[(size_t)model * M.maxModes + r] = 1. / (sqrt(val) * 2.0 * M_PI); // model.go:830
    }
    if (tid == 0) {
        M.nModes[model] = nOut;
        if (nSweeps < 0 && M.status[model] == HDB_STATUS_OK) M.status[model] = HDB_STATUS_EIGEN;
    }
    if (!M.shapes) return;
    FSTAMP("order, hz");

    // ---- shapes: out[r][g] = sum_j V(j, order[r]) sol[j][g], on the FP64 tensor cores ----
    // A fragment (8 modes x 4 j): lane holds A[lane / 4][lane % 4]; B fragment (4 j x 8 DOFs): lane holds B[lane % 4][lane / 4];
    // C fragment (8 x 8): lane holds C[lane / 4][2 (lane % 4) + {0, 1}].  A warp owns mode tiles warp, warp + 8, ... and all
    // four 8-DOF tiles of the streamed 32-DOF tile.
    const int nMT = (nOut + 7) / 8, nGT = (M.dof + SH_NT - 1) / SH_NT;
    const unsigned tileA = (unsigned)__cvta_generic_to_shared(tile);
    auto stage = [&](int gt, int buf) { // solution rows j = 0..n-1, DOFs [gt * 32, gt * 32 + 32) -> tile[buf]; out-of-range = 0
        const int g0 = gt * SH_NT;
        const bool vec = ((M.dof & 1) == 0) && ((reinterpret_cast<uintptr_t>(sol) & 15) == 0);
        for (int idx = tid; idx < KP * (SH_NT / 2); idx += EIG_THREADS) {
            const int j = idx / (SH_NT / 2), c2 = (idx - j * (SH_NT / 2)) * 2, g = g0 + c2;
            double *dst = tile + ((size_t)buf * KP + j) * SH_LDB + c2;
            if (j < n && g + 1 < M.dof && vec) cp_async16(tileA + (unsigned)(((size_t)buf * KP + j) * SH_LDB + c2) * 8u, sol + (size_t)j * M.dof + g);
            else {
                dst[0] = (j < n && g < M.dof) ? sol[(size_t)j * M.dof + g] : 0.0;
                dst[1] = (j < n && g + 1 < M.dof) ? sol[(size_t)j * M.dof + g + 1] : 0.0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double *out = M.shapes + (size_t)model * M.maxModes * M.dof;
    stage(0, 0);
    for (int gt = 0; gt < nGT; gt++) {
        if (gt + 1 < nGT) {
            stage(gt + 1, (gt + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double *tb = tile + (size_t)(gt & 1) * KP * SH_LDB;
        const int g0 = gt * SH_NT;
        for (int mt = warp; mt < nMT; mt += SH_WARPS) {
            const int rA = mt * 8 + (lane >> 2);                          // this lane's mode row of the A fragment
            const double *ya = V + (size_t)order[rA < nOut ? rA : nOut - 1] * LD + (lane & 3);
            double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            const double *bb = tb + (size_t)(lane & 3) * SH_LDB + (lane >> 2);
#pragma unroll 2
            for (int k0 = 0; k0 < KP; k0 += 4) {
                const double a = (k0 + (lane & 3) < n) ? ya[k0] : 0.0;
                const double *bk = bb + (size_t)k0 * SH_LDB;
#pragma unroll
                for (int t = 0; t < 4; t++) dmma8x8x4(c[t][0], c[t][1], a, bk[8 * t]);
            }
            const int r = mt * 8 + (lane >> 2);
            double sq = 0.0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int g = g0 + 8 * t + 2 * (lane & 3);
                if (r < nOut) {
                    if (g < M.dof) { out[(size_t)r * M.dof + g] = c[t][0]; sq = fma(c[t][0], c[t][0], sq); }
                    if (g + 1 < M.dof) { out[(size_t)r * M.dof + g + 1] = c[t][1]; sq = fma(c[t][1], c[t][1], sq); }
                }
            }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            if ((lane & 3) == 0 && r < nOut) nrm[r] += sq; // this warp alone owns mode tile mt: fixed order of additions
        }
        __syncthreads(); // the tile may be overwritten
    }
    FSTAMP("shapes");
    // ---- unit 2-norm over all 3*nP entries (what Dgeev returns, model.go:836-841): the raw rows are still in L2 ----
    for (int r = tid; r < nOut; r += EIG_THREADS) nrm[r] = 1.0 / sqrt(nrm[r]);
    __syncthreads();
    for (int r = warp; r < nOut; r += SH_WARPS) { // a warp per mode row: no division per element
        const double sc = nrm[r];
        double *row = out + (size_t)r * M.dof;
        for (int g = lane; g < M.dof; g += 32) row[g] *= sc;
    }
    FSTAMP("normalise");
#undef VV
}

static size_t finish_smem(int n, int &LD)
{
    LD = n | 1;
    if ((LD & 3) == 1) LD += 2;
    const int KP = (n + 3) & ~3;
    if (LD < KP) LD = KP | 1;
    return ((size_t)LD * n + 2 * (size_t)n + (size_t)2 * KP * SH_LDB) * sizeof(double) + (size_t)n * sizeof(int) + 80;
}

static size_t tridiag_smem(int n, int &LDT)
{
    LDT = (n + 15) / 16 * 16 + 12; // = 12 (mod 16): see sym_tridiag_full
    if (LDT - 16 >= n) LDT -= 16;
    return ((size_t)LDT * n + 3 * (size_t)n + TRI_THREADS / 32 + 8) * sizeof(double) + 64;
}

bool modal_split_fits(int n, int smemOptin)
{
    int LD, LDT;
    return n >= 1 && finish_smem(n, LD) + 1024 <= (size_t)smemOptin && tridiag_smem(n, LDT) + 1024 <= (size_t)smemOptin;
}

void modal_split_scratch(int n, size_t &qDoubles, size_t &rotCap, size_t &maxSweeps)
{
    int LD;
    finish_smem(n, LD);
    qDoubles = (size_t)n * LD;
    rotCap = (size_t)3 * n * n + 64;     // (cos, sin) pairs: QL needs about 1.6 sweeps of average length n / 2 per eigenvalue
    maxSweeps = (size_t)8 * n + 8;
}

cudaError_t launch_modal_split(const ModalLaunch &M, int smemOptin, cudaStream_t st)
{
    if (M.nModels == 0) return cudaSuccess;
    int LD;
    const size_t smemF = finish_smem(M.s, LD);
    int LDT;
    const size_t smemT = tridiag_smem(M.s, LDT);
    if (smemF + 1024 > (size_t)smemOptin || smemT + 1024 > (size_t)smemOptin) return cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (smemT > 48 * 1024 && (e = cudaFuncSetAttribute(k_modal_tridiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemT)) != cudaSuccess) return e;
    if (smemF > 48 * 1024 && (e = cudaFuncSetAttribute(k_modal_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemF)) != cudaSuccess) return e;
    k_modal_tridiag<<<M.nModels, TRI_THREADS, smemT, st>>>(M, LD, LDT);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const size_t smemC = (size_t)CH_WARPS * 2 * M.s * sizeof(double);
    if (smemC > 48 * 1024 && (e = cudaFuncSetAttribute(k_modal_ql_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC)) != cudaSuccess) return e;
    k_modal_ql_chain<<<(M.nModels + CH_WARPS - 1) / CH_WARPS, CH_WARPS * 32, smemC, st>>>(M);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_modal_finish<<<M.nModels, EIG_THREADS, smemF, st>>>(M, LD);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// linear buckling (model.go:491-601).  The reference forms the dense H = K^-1 G and takes the real eigenvalues of
// the non-symmetric product.  Only free rows and columns matter (eliminated rows of H are zero), and with
// S = (K^-1)_ff = C C^T (Cholesky; K is SPD where the model is stable) the eigenvalues of S G are those of the
// symmetric B = C^T G C, eigenvectors x = C y: the shared-memory symmetric solver above does the rest.
// One block per (model, load case).  G(i, j) = sum over beams of axialForce(beam) * gUnit (feBeam2d.go:175-275).
// ------------------------------------------------------------------------------------------------
__global__ void k_unit_loads(double *loads, int nModels, int s, int dof, const int32_t *freeDof)
{
    const size_t n = (size_t)nModels * s * dof;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % dof), c = (int)((i / dof) % s);
        loads[i] = (freeDof[c] == g) ? 1.0 : 0.0;
    }
}

cudaError_t launch_unit_loads(double *loads, int nModels, int s, int dof, const int32_t *freeDof, cudaStream_t st)
{
    k_unit_loads<<<1184, 256, 0, st>>>(loads, nModels, s, dof, freeDof);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(EIG_THREADS, 1) k_buckling_eigen(const BucklingLaunch P, const int LD, double *gscr)
{
    extern __shared__ __align__(16) double sm[];
    const int n = P.s, tid = threadIdx.x;
    const int model = blockIdx.x / P.nLC, lc = blockIdx.x - model * P.nLC;
    // the three n x n work matrices live in shared memory when they fit (gscr == NULL) and otherwise in a global scratch
    // record of this block (models with more than ~95 free DOFs: slower, L2 / HBM resident, same code)
    double *A = gscr ? gscr + (size_t)blockIdx.x * 3 * LD * n : sm; // S, then its Cholesky factor C (lower), element (i, j) at i + j * LD
    double *Gm = A + (size_t)LD * n; // G, then B = C^T G C, then the eigenvectors Y
    double *Tm = Gm + (size_t)LD * n; // G C, then X = C Y
    double *d = gscr ? sm : Tm + (size_t)LD * n;
    double *e = d + n, *cc = e + n, *ss = cc + n; // cc: ring of QL rotations, QL_RING slots of 2n
    double *red = cc + (size_t)2 * QL_RING * n;   // [EIG_THREADS/32 + 8]
    double *ax = red + EIG_THREADS / 32 + 8; // [nB] axial force of every beam
    int *order = reinterpret_cast<int *>(ax + P.nB);
    __shared__ int s_bad, s_cnt;
#define AA(i, j) A[(i) + (size_t)(j) * LD]
#define GG(i, j) Gm[(i) + (size_t)(j) * LD]
#define TT(i, j) Tm[(i) + (size_t)(j) * LD]
    const double *sol = P.sol + (size_t)model * n * P.dof;
    const double *bf = P.forces + ((size_t)model * P.nLC + lc) * P.nB * 6;
    if (tid == 0) { s_bad = 0; s_cnt = 0; }
    for (int b = tid; b < P.nB; b += EIG_THREADS) { // feBeam2d.go:180-188
        double a = 0.0;
        if (bf[6 * b] > 0) a = bf[6 * b];
        if (-bf[6 * b + 3] > a) a = -bf[6 * b + 3];
        ax[b] = a;
    }
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
        const int i = idx % n, j = idx / n;
        AA(i, j) = 0.5 * (sol[(size_t)j * P.dof + P.freeDof[i]] + sol[(size_t)i * P.dof + P.freeDof[j]]);
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
        const int i = idx % n, j = idx / n;
        double sum = 0.0;
        for (int q = P.gPtr[idx]; q < P.gPtr[idx + 1]; q++) {
            const int code = P.gIdx[q];
            sum = fma(ax[code / 36], P.gUnit[code], sum);
        }
        GG(i, j) = sum;
    }
    // ---- Cholesky S = C C^T, in place, right-looking ----
    for (int k = 0; k < n; k++) {
        __syncthreads();
        const double piv = AA(k, k);
        if (!(piv > 0.0)) { // not positive definite: unstable or failed model
            if (tid == 0) s_bad = 1;
            break; // uniform: every thread reads the same value
        }
        const double r = sqrt(piv);
        __syncthreads();
        for (int i = k + tid; i < n; i += EIG_THREADS) AA(i, k) = (i == k) ? r : AA(i, k) / r;
        __syncthreads();
        const int m = n - k - 1;
        for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
            const int i = k + 1 + idx % m, j = k + 1 + idx / m;
            if (j <= i) AA(i, j) = fma(-AA(i, k), AA(j, k), AA(i, j));
        }
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) {
            P.nOut[(size_t)model * P.nLC + lc] = 0;
            if (P.status[model] == HDB_STATUS_OK) P.status[model] = HDB_STATUS_EIGEN;
        }
        return;
    }
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) { // upper triangle of C = 0
        const int i = idx % n, j = idx / n;
        if (j > i) AA(i, j) = 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) { // T = G C
        const int i = idx % n, j = idx / n;
        double sum = 0.0;
        for (int k = j; k < n; k++) sum = fma(GG(i, k), AA(k, j), sum);
        TT(i, j) = sum;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) { // B = C^T T (lower triangle computed, mirrored)
        const int i = idx % n, j = idx / n;
        if (j > i) continue;
        double sum = 0.0;
        for (int k = i; k < n; k++) sum = fma(AA(k, i), TT(k, j), sum);
        GG(i, j) = sum;
        GG(j, i) = sum;
    }
    __syncthreads();
    const int failed = sym_eig_smem(Gm, LD, n, d, e, cc, ss, red);
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) { // X = C Y
        const int i = idx % n, r = idx / n;
        double sum = 0.0;
        for (int k = 0; k <= i; k++) sum = fma(AA(i, k), GG(k, r), sum);
        TT(i, r) = sum;
    }
    __syncthreads();
    // ---- model.go:560-600: non-zero eigenvalues, factor 1/|val|, ascending factor, first `amount` ----
    if (tid == 0) { // n is small: a serial scan is cheaper than a reduction
        double mx = 0.0;
        for (int j = 0; j < n; j++) mx = fmax(mx, fabs(d[j]));
        red[0] = mx;
    }
    __syncthreads();
    const double vmax = red[0];
    const double tiny = vmax * 1e-13; // an exact zero of the reference's H shows up as round-off here
    for (int j = tid; j < n; j += EIG_THREADS) {
        const double lj = fabs(d[j]);
        if (!(lj > tiny)) continue;
        int rank = 0;
        for (int i = 0; i < n; i++) {
            const double li = fabs(d[i]);
            if (li > tiny && (li > lj || (li == lj && i < j))) rank++;
        }
        order[rank] = j;
        atomicAdd(&s_cnt, 1);
    }
    __syncthreads();
    const int cnt = min(s_cnt, P.amount), nOut = min(cnt, P.maxOut);
    double *fo = P.factors + ((size_t)model * P.nLC + lc) * P.maxOut;
    for (int r = tid; r < nOut; r += EIG_THREADS) fo[r] = 1. / fabs(d[order[r]]);
    if (tid == 0) {
        P.nOut[(size_t)model * P.nLC + lc] = cnt;
        if (failed && P.status[model] == HDB_STATUS_OK) P.status[model] = HDB_STATUS_EIGEN;
    }
    if (!P.shapes) return;
    double *out = P.shapes + ((size_t)model * P.nLC + lc) * P.maxOut * P.dof;
    for (int idx = tid; idx < nOut * P.dof; idx += EIG_THREADS) out[idx] = 0.0;
    __syncthreads();
    for (int r = 0; r < nOut; r++) { // unit 2-norm over all DOFs (eliminated DOFs stay zero)
        const int col = order[r];
        double p2 = 0.0;
        for (int i = tid; i < n; i += EIG_THREADS) p2 = fma(TT(i, col), TT(i, col), p2);
        const double nrm = sqrt(block_sum(p2, red));
        for (int i = tid; i < n; i += EIG_THREADS) out[(size_t)r * P.dof + P.freeDof[i]] = TT(i, col) / nrm;
    }
#undef AA
#undef GG
#undef TT
}

static size_t buckling_smem(int n, int nB, int &LD)
{
    LD = n | 1;
    return ((size_t)3 * LD * n + 2 * (size_t)n + 2 * (size_t)QL_RING * n + EIG_THREADS / 32 + 8 + nB) * sizeof(double) + (size_t)n * sizeof(int) + 64;
}

int buckling_max_s(int smemOptin)
{
    int n = 0, LD;
    while (buckling_smem(n + 1, 0, LD) + 16384 <= (size_t)smemOptin) n++;
    return n;
}

// shared memory of a block without the three work matrices (global-scratch path)
static size_t buckling_smem_small(int n, int nB)
{
    return (2 * (size_t)n + 2 * (size_t)QL_RING * n + EIG_THREADS / 32 + 8 + nB) * sizeof(double) + (size_t)n * sizeof(int) + 64;
}

size_t buckling_scratch_doubles(int n, int nB, int smemOptin)
{ // per (model, case) block; 0 = the matrices fit shared memory
    int LD;
    if (buckling_smem(n, nB, LD) + 1024 <= (size_t)smemOptin) return 0;
    return (size_t)3 * LD * n;
}

cudaError_t launch_buckling_eigen(const BucklingLaunch &B, int smemOptin, cudaStream_t st, double *gscr)
{
    int LD;
    size_t smem = buckling_smem(B.s, B.nB, LD);
    if (gscr) smem = buckling_smem_small(B.s, B.nB);
    if (smem + 1024 > (size_t)smemOptin) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_buckling_eigen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_buckling_eigen<<<B.nModels * B.nLC, EIG_THREADS, smem, st>>>(B, LD, gscr);
    return cudaGetLastError();
}

} // namespace hdb
