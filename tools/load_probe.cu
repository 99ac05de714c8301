// load_probe.cu — what bounds the READ side of a persistent one-block-per-SM streaming kernel?
// Each block reads its models' contiguous 124 kB load blocks and writes them into shared memory (and optionally
// back to global).  Variants: threads, loads in flight, 8- vs 16-byte loads, shared-memory size (L1 carve-out),
// cp.async.cg (L1 bypass) into a shared-memory ring.
#include <cstdio>
#include <cuda_runtime.h>

template <int U, int VEC>
__global__ void k_read(const double *in, double *out, int nModels, int nIn, int wr, long long *cyc)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, NT = blockDim.x;
    long long t0 = clock64();
    double acc = 0;
    for (int m = blockIdx.x; m < nModels; m += gridDim.x) {
        const double *p = in + (size_t)m * nIn;
        double *o = out + (size_t)m * nIn;
        if (VEC == 1) {
            for (int i0 = 0; i0 < nIn; i0 += NT * U) {
                double v[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * NT + tid;
                    v[u] = i < nIn ? __ldcs(p + i) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * NT + tid;
                    if (i < nIn) {
                        sm[i & 4095] = v[u];
                        if (wr) o[i] = -v[u];
                    }
                    acc += v[u];
                }
            }
        } else {
            const double2 *p2 = reinterpret_cast<const double2 *>(p);
            double2 *o2 = reinterpret_cast<double2 *>(o);
            const int n2 = nIn / 2;
            for (int i0 = 0; i0 < n2; i0 += NT * U) {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * NT + tid;
                    v[u] = i < n2 ? __ldcs(p2 + i) : make_double2(0, 0);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * NT + tid;
                    if (i < n2) {
                        sm[(2 * i) & 4095] = v[u].x;
                        sm[(2 * i + 1) & 4095] = v[u].y;
                        if (wr) o2[i] = make_double2(-v[u].x, -v[u].y);
                    }
                    acc += v[u].x + v[u].y;
                }
            }
        }
        __syncthreads();
    }
    if (acc == 123.456) out[0] = acc;
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

// cp.async.cg 16 B per thread into a per-thread ring of D slots in shared memory (L1 bypass)
template <int D>
__global__ void k_cpasync(const double *in, double *out, int nModels, int nIn, int wr, long long *cyc)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, NT = blockDim.x;
    long long t0 = clock64();
    double acc = 0;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(sm) + tid * 16u; // slot d at ring + d * NT * 16
    for (int m = blockIdx.x; m < nModels; m += gridDim.x) {
        const double2 *p2 = reinterpret_cast<const double2 *>(in + (size_t)m * nIn);
        double2 *o2 = reinterpret_cast<double2 *>(out + (size_t)m * nIn);
        const int n2 = nIn / 2;
        const int rounds = (n2 + NT - 1) / NT;
        for (int r = 0; r < D - 1; r++) {
            const int i = r * NT + tid;
            if (r < rounds && i < n2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (r % D) * NT * 16u), "l"(p2 + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int r = 0; r < rounds; r++) {
            const int rp = r + D - 1, ip = rp * NT + tid;
            if (rp < rounds && ip < n2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (rp % D) * NT * 16u), "l"(p2 + ip) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
            const int i = r * NT + tid;
            if (i < n2) {
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ring + (r % D) * NT * 16u));
                acc += v.x + v.y;
                if (wr) o2[i] = make_double2(-v.x, -v.y);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    if (acc == 123.456) out[0] = acc;
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <typename K>
static void run(const char *name, K kern, int threads, int smemKB, const double *in, double *out, int M, int nIn, int wr, long long *cyc)
{
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemKB * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<148, threads, smemKB * 1024>>>(in, out, M, nIn, wr, cyc);
    cudaEventRecord(e0);
    kern<<<148, threads, smemKB * 1024>>>(in, out, M, nIn, wr, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; i++) c += (double)h[i];
    const cudaError_t e = cudaGetLastError();
    printf("%-28s thr %4d smem %3d kB wr %d: %.3f ms  %6.0f GB/s read  %6.0f cycles/model %s\n", name, threads, smemKB, wr, ms,
           (double)M * nIn * 8 / (ms * 1e-3) / 1e9, c / M, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const int M = 4096, nIn = 15480;
    double *in, *out;
    long long *cyc;
    cudaMalloc(&in, (size_t)M * nIn * 8);
    cudaMalloc(&out, (size_t)M * nIn * 8);
    cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0, (size_t)M * nIn * 8);
    for (int wr = 0; wr < 2; wr++)
        for (int smem : {210, 150, 90}) {
            for (int thr : {384, 768}) {
                run("ldg64 U=7", k_read<7, 1>, thr, smem, in, out, M, nIn, wr, cyc);
                run("ldg64 U=14", k_read<14, 1>, thr, smem, in, out, M, nIn, wr, cyc);
                run("ldg64 U=28", k_read<28, 1>, thr, smem, in, out, M, nIn, wr, cyc);
                run("ldg128 U=7", k_read<7, 2>, thr, smem, in, out, M, nIn, wr, cyc);
                run("ldg128 U=14", k_read<14, 2>, thr, smem, in, out, M, nIn, wr, cyc);
            }
            run("cp.async.cg D=4", k_cpasync<4>, 384, smem, in, out, M, nIn, wr, cyc);
            run("cp.async.cg D=8", k_cpasync<8>, 384, smem, in, out, M, nIn, wr, cyc);
            run("cp.async.cg D=16", k_cpasync<16>, 384, smem, in, out, M, nIn, wr, cyc);
            run("cp.async.cg D=8", k_cpasync<8>, 768, smem, in, out, M, nIn, wr, cyc);
        }
    return 0;
}
