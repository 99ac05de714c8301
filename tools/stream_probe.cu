// stream_probe.cu — how fast can "one CTA per model" stream the LinearStatic traffic pattern
// (read 124 kB, write 494 kB per model) as a function of CTAs per SM and threads per CTA?
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void k_stream(const double *in, double *out, int nIn, int nOut, int dummySmem)
{
    extern __shared__ double sm[];
    const int model = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const double *p = in + (size_t)model * nIn;
    double *o = out + (size_t)model * nOut;
    double acc = 0;
    for (int i0 = 0; i0 < nIn; i0 += NT * UNROLL) {
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int i = i0 + u * NT + tid;
            v[u] = i < nIn ? p[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int i = i0 + u * NT + tid;
            if (i < nIn) sm[i & 4095] = v[u];
            acc += v[u];
        }
    }
    __syncthreads();
    for (int i = tid; i < nOut; i += NT) o[i] = sm[i & 4095] + acc;
}

int main()
{
    const int M = 4096, nIn = 15480 + 771, nOut = 61800;
    double *in, *out;
    cudaMalloc(&in, (size_t)M * nIn * 8);
    cudaMalloc(&out, (size_t)M * nOut * 8);
    cudaMemset(in, 0, (size_t)M * nIn * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_stream<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int threads : {256, 512, 1024})
        for (int smemKB : {200, 100, 70, 50, 33}) {
            const int smem = smemKB * 1024;
            k_stream<16><<<M, threads, smem>>>(in, out, nIn, nOut, 0);
            cudaEventRecord(e0);
            for (int r = 0; r < 3; r++) k_stream<16><<<M, threads, smem>>>(in, out, nIn, nOut, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= 3;
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<16>, threads, smem);
            printf("threads %4d smem %3d kB (CTAs/SM %2d): %.3f ms per %d models -> %.0f GB/s\n", threads, smemKB, occ, ms, M,
                   (double)M * (nIn + nOut) * 8 / (ms * 1e-3) / 1e9);
        }
    return 0;
}
