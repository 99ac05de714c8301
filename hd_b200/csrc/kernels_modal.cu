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
#include <cmath>

#include "modal.hpp"

namespace hdb {

constexpr double GRAVITY = 9.80665; // model.go:713
constexpr int EIG_THREADS = 256;

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

__global__ void __launch_bounds__(EIG_THREADS, 1) k_modal_eigen(const ModalLaunch M, const int LD)
{
    extern __shared__ __align__(16) double sm[];
    const int n = M.s, tid = threadIdx.x, model = blockIdx.x;
    double *V = sm;                 // [n][n] element (i, j) at i + j * LD
    double *d = V + (size_t)LD * n; // [n]
    double *e = d + n;              // [n]
    double *cc = e + n;             // [n] rotation cosines of one QL sweep
    double *ss = cc + n;            // [n] rotation sines
    double *red = ss + n;           // [EIG_THREADS/32 + 8]
    int *order = reinterpret_cast<int *>(red + EIG_THREADS / 32 + 8); // [n]
    __shared__ double s_h, s_hh;
    __shared__ int s_m, s_more, s_fail;
#define VV(i, j) V[(i) + (size_t)(j) * LD]
    const double *sol = M.sol + (size_t)model * n * M.dof;
    const double *ms = M.msqrt + (size_t)model * n;

    if (tid == 0) s_fail = 0;
    // ---- B = M^1/2 (K^-1)_SS M^1/2, symmetrised ----
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
        const int i = idx % n, j = idx / n;
        const double bij = ms[i] * sol[(size_t)j * M.dof + M.massDof[i]];
        const double bji = ms[j] * sol[(size_t)i * M.dof + M.massDof[j]];
        VV(i, j) = 0.5 * (bij + bji);
    }
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
            for (int j = tid; j < i; j += EIG_THREADS) {
                VV(j, i) = d[j];
                double g = 0.0;
                for (int k = 0; k <= j; k++) g = fma(VV(j, k), d[k], g);
                for (int k = j + 1; k < i; k++) g = fma(VV(k, j), d[k], g);
                e[j] = g / h;
            }
            __syncthreads();
            part = 0.0;
            for (int j = tid; j < i; j += EIG_THREADS) part += e[j] * d[j];
            const double f = block_sum(part, red);
            const double hh = f / (h + h);
            for (int j = tid; j < i; j += EIG_THREADS) e[j] -= hh * d[j];
            __syncthreads();
            // rank-2 update of the lower triangle: thread k owns row k
            for (int k = tid; k < i; k += EIG_THREADS) {
                const double ek = e[k], dk = d[k];
                for (int j = 0; j <= k; j++) VV(k, j) -= d[j] * ek + e[j] * dk;
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
            for (int j = tid; j <= i; j += EIG_THREADS) { // thread j owns column j
                double g = 0.0;
                for (int k = 0; k <= i; k++) g = fma(VV(k, i + 1), VV(k, j), g);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
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

    // ---- implicit QL with accumulated vectors (tql2): thread 0 runs the scalar recurrence of a sweep
    //      and leaves its rotations in shared memory, then every thread applies them to its own row ----
    double f = 0.0, tst1 = 0.0; // live in thread 0 only
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        if (tid == 0) {
            tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
            int m = l;
            while (m < n) {
                if (fabs(e[m]) <= eps * tst1) break;
                m++;
            }
            if (m >= n) m = n - 1;
            s_m = m;
        }
        __syncthreads();
        const int m = s_m;
        if (m > l) {
            int iter = 0;
            do {
                if (tid == 0) {
                    iter++;
                    double g = d[l];
                    double p = (d[l + 1] - g) / (2.0 * e[l]);
                    double r = hypot(p, 1.0);
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    const double dl1 = d[l + 1];
                    double h = g - d[l];
                    for (int i = l + 2; i < n; i++) d[i] -= h;
                    f += h;
                    p = d[m];
                    double c = 1.0, c2 = c, c3 = c;
                    const double el1 = e[l + 1];
                    double s = 0.0, s2 = 0.0;
                    for (int i = m - 1; i >= l; i--) {
                        c3 = c2;
                        c2 = c;
                        s2 = s;
                        const double ei = e[i];
                        g = c * ei;
                        h = c * p;
                        // the tridiagonal entries are O(lambda): no overflow / underflow guard needed, and the
                        // scalar recurrence is the critical path of the whole kernel
                        r = sqrt(fma(p, p, ei * ei));
                        const double rinv = 1.0 / r;
                        e[i + 1] = s * r;
                        s = ei * rinv;
                        c = p * rinv;
                        p = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        cc[i] = c;
                        ss[i] = s;
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                    s_more = (fabs(e[l]) > eps * tst1) ? 1 : 0;
                    if (iter > 60) {
                        s_more = 0;
                        s_fail = 1;
                    }
                }
                __syncthreads();
                for (int k = tid; k < n; k += EIG_THREADS) {
                    double hk = VV(k, m);
                    for (int i = m - 1; i >= l; i--) {
                        const double c = cc[i], s = ss[i], vi = VV(k, i);
                        VV(k, i + 1) = fma(s, vi, c * hk);
                        hk = fma(c, vi, -s * hk);
                    }
                    VV(k, l) = hk;
                }
                __syncthreads();
            } while (s_more);
        }
        if (tid == 0) {
            d[l] = d[l] + f;
            e[l] = 0.0;
        }
        __syncthreads();
    }

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
    const int nOut = min(n, M.maxModes);
    for (int r = tid; r < nOut; r += EIG_THREADS) {
        const double val = fabs(d[order[r]]);
        M.hz[(size_t)model * M.maxModes + r] = 1. / (sqrt(val) * 2.0 * M_PI); // model.go:830
    }
    if (tid == 0) {
        M.nModes[model] = nOut;
        if (s_fail && M.status[model] == HDB_STATUS_OK) M.status[model] = HDB_STATUS_EIGEN;
    }
    if (!M.shapes) return;

    // ---- shapes: raw[r][g] = sum_j sol[j][g] * y_r[j]; unit 2-norm over all 3*nP entries ----
    double *nrm = e; // reuse: squared norms per output mode
    for (int r = tid; r < n; r += EIG_THREADS) nrm[r] = 0.0;
    __syncthreads();
    constexpr int MT = 8;
    double *out = M.shapes + (size_t)model * M.maxModes * M.dof;
    for (int r0 = 0; r0 < nOut; r0 += MT) {
        int col[MT];
#pragma unroll
        for (int q = 0; q < MT; q++) col[q] = order[min(r0 + q, nOut - 1)];
        double sq[MT];
#pragma unroll
        for (int q = 0; q < MT; q++) sq[q] = 0.0;
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
                if (r0 + q < nOut) {
                    out[(size_t)(r0 + q) * M.dof + g] = acc[q];
                    sq[q] = fma(acc[q], acc[q], sq[q]);
                }
        }
#pragma unroll
        for (int q = 0; q < MT; q++) {
            double v = sq[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && r0 + q < nOut) atomicAdd(&nrm[r0 + q], v);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < nOut * M.dof; idx += EIG_THREADS) {
        const int r = idx / M.dof;
        out[idx] = out[idx] / sqrt(nrm[r]);
    }
#undef VV
}

static size_t eigen_smem(int n, int &LD)
{
    LD = n | 1;
    if ((LD & 3) == 1) LD += 2; // LD = 3 mod 4 keeps column strides off the same banks
    return ((size_t)LD * n + 4 * (size_t)n + EIG_THREADS / 32 + 8) * sizeof(double) + (size_t)n * sizeof(int) + 64;
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

} // namespace hdb
