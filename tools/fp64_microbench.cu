// fp64_microbench.cu — FP64 pipe facts the design depends on (B200 peaks for FP64 are not in
// MEASURED_PEAKS.json): DFMA dependent-issue latency, single-warp DFMA issue rate, full-chip DFMA
// throughput, and shared-memory load-to-use latency.   nvcc -O3 -arch=sm_100a fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_latency(double *out, double a, double b, int n, long long *cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_ilp(double *out, double a, double b, int n, long long *cyc)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds(double *out, int n, long long *cyc)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) p = idx[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 8);
    const int n = 4096;
    k_latency<<<1, 32>>>(out, 1.0000001, 1e-9, n, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", (double)h / (n * 16.0));
    k_ilp<8><<<1, 32>>>(out, 1.0000001, 1e-9, n, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("one warp, 8 independent chains: %.2f cycles per DFMA warp-instruction\n", (double)h / (n * 32.0));
    for (int warps = 4; warps <= 32; warps *= 2) {
        k_ilp<8><<<1, 32 * warps>>>(out, 1.0000001, 1e-9, n, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps on one SM: %.2f DFMA lanes/cycle/SM\n", warps, warps * 32.0 * n * 32.0 / (double)h);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_ilp<8><<<148 * 8, 256>>>(out, 1.0000001, 1e-9, n, cyc);
    cudaEventRecord(e0);
    k_ilp<8><<<148 * 8, 256>>>(out, 1.0000001, 1e-9, n, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("full chip: %.2f TFLOP/s FP64 (DFMA = 2 flop)\n", 148.0 * 8 * 256 * n * 32.0 * 2 / (ms * 1e-3) / 1e12);
    k_lds<<<1, 32>>>(out, n, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent load-to-use: %.2f cycles\n", (double)h / (n * 16.0));
    return 0;
}
