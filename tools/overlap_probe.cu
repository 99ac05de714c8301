// overlap_probe.cu — can one SM read under its own stores?  Persistent block per SM, 384 threads, 200 kB of shared
// memory (one block per SM).  Per model: read 140 kB (16-byte loads, 7 in flight per thread), write 494 kB
// (16-byte streaming stores).  A: all warps read, then all warps write.  B: warps 0-3 only read, warps 4-11 only
// write, concurrently.  C/D: only the reads / only the writes with 12 warps (the two legs of A on their own).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void do_reads(const double2 *p, int n2, int t, int nt, double *sm, double &acc)
{
    for (int i0 = 0; i0 < n2; i0 += nt * 7) {
        double2 v[7];
#pragma unroll
        for (int u = 0; u < 7; u++) {
            const int i = i0 + u * nt + t;
            v[u] = i < n2 ? __ldcs(p + i) : make_double2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 7; u++) {
            const int i = i0 + u * nt + t;
            if (i < n2) sm[(2 * i) & 4095] = v[u].x;
            acc += v[u].y;
        }
    }
}
// reads through a per-thread cp.async.cg ring of D slots (L1 bypass, no registers held)
template <int D>
__device__ __forceinline__ void do_reads_async(const double2 *p, int n2, int t, int nt, double *ringBase, double &acc)
{
    const unsigned ring = (unsigned)__cvta_generic_to_shared(ringBase) + t * 16u;
    const int rounds = (n2 + nt - 1) / nt;
    for (int r = 0; r < D - 1; r++) {
        const int i = r * nt + t;
        if (r < rounds && i < n2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (r % D) * nt * 16u), "l"(p + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int r = 0; r < rounds; r++) {
        const int rp = r + D - 1, ip = rp * nt + t;
        if (rp < rounds && ip < n2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (rp % D) * nt * 16u), "l"(p + ip) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        if (r * nt + t < n2) {
            double x, y;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(ring + (r % D) * nt * 16u));
            acc += x + y;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void do_writes(double2 *o, int n2, int t, int nt, double a)
{
    for (int i = t; i < n2; i += nt) __stcs(o + i, make_double2(a, a + i));
}

__global__ void __launch_bounds__(384, 1) k_probe(const double *in, double *out, int nModels, int nIn, int nOut, int mode, long long *cyc)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const long long t0 = clock64();
    double acc = 0;
    for (int m = blockIdx.x; m < nModels; m += gridDim.x) {
        const double2 *p = reinterpret_cast<const double2 *>(in + (size_t)m * nIn);
        double2 *o = reinterpret_cast<double2 *>(out + (size_t)m * nOut);
        if (mode == 0) { // A
            do_reads(p, nIn / 2, tid, 384, sm, acc);
            __syncthreads();
            do_writes(o, nOut / 2, tid, 384, acc);
            __syncthreads();
        } else if (mode == 1) { // B
            if (tid < 128) do_reads(p, nIn / 2, tid, 128, sm, acc);
            else do_writes(o, nOut / 2, tid - 128, 256, acc);
            __syncthreads();
        } else if (mode == 2) { // C
            do_reads(p, nIn / 2, tid, 384, sm, acc);
            __syncthreads();
        } else if (mode == 3) { // D
            do_writes(o, nOut / 2, tid, 384, acc);
            __syncthreads();
        } else if (mode == 5) { // F: 4 warps read through a 16-deep cp.async ring while 8 warps write
            if (tid < 128) do_reads_async<16>(p, nIn / 2, tid, 128, sm + 4096, acc);
            else do_writes(o, nOut / 2, tid - 128, 256, acc);
            __syncthreads();
        } else if (mode == 6) { // G: the same reads alone
            if (tid < 128) do_reads_async<16>(p, nIn / 2, tid, 128, sm + 4096, acc);
            __syncthreads();
        } else { // E (mode 4): writes with 8 warps only
            if (tid >= 128) do_writes(o, nOut / 2, tid - 128, 256, acc);
            __syncthreads();
        }
    }
    if (acc == 123.456) out[0] = acc;
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main()
{
    const int M = 4096, nIn = 17544, nOut = 61800; // 140 kB in (loads + factor), 494 kB out per model
    double *in, *out;
    long long *cyc;
    cudaMalloc(&in, (size_t)M * nIn * 8);
    cudaMalloc(&out, (size_t)M * nOut * 8);
    cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0, (size_t)M * nIn * 8);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char *names[7] = {"A reads then writes (12 warps)", "B reads (4 warps) || writes (8 warps)", "C reads only (12 warps)",
                            "D writes only (12 warps)", "E writes only (8 warps)", "F cp.async reads (4 warps, ring 16) || writes (8 warps)",
                            "G cp.async reads only (4 warps, ring 16)"};
    for (int mode : {0, 1, 2, 3, 4, 5, 6}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_probe<<<148, 384, 200 * 1024>>>(in, out, M, nIn, nOut, mode, cyc);
        cudaEventRecord(e0);
        k_probe<<<148, 384, 200 * 1024>>>(in, out, M, nIn, nOut, mode, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; i++) c += (double)h[i];
        printf("%-40s %.3f ms per %d models, %6.0f cycles/model, %s\n", names[mode], ms, M, c / M, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
