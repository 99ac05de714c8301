// pcie_probe.cu — what the host of the GPU box can move over PCIe with 1 / 2 / 4 / 8 GPUs copying at the same time.
// One host thread per GPU, pinned buffers (cudaHostAlloc), one cudaMemcpyAsync per chunk on the GPU's own stream, H2D
// alone, D2H alone and both directions at once.  The end-to-end figure of bench.py moves 4.05 GB D2H + 1.07 GB H2D per
// step and GPU; this probe says whether its N-GPU number is the library's pipeline or the host's memory system.
//   nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu && tools/pcie_probe [chunk MiB] [chunks]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

static double run(int nGpu, int mode, size_t chunk, int chunks)
{ // mode 0 = H2D, 1 = D2H, 2 = both
    std::vector<std::thread> th;
    std::vector<double> secs(nGpu, 0.0);
    std::vector<int> ready(nGpu, 0);
    volatile int go = 0;
    for (int g = 0; g < nGpu; g++)
        th.emplace_back([&, g] {
            cudaSetDevice(g);
            void *hIn = nullptr, *hOut = nullptr, *dIn = nullptr, *dOut = nullptr;
            cudaHostAlloc(&hIn, chunk, cudaHostAllocDefault);
            cudaHostAlloc(&hOut, chunk, cudaHostAllocDefault);
            cudaMalloc(&dIn, chunk);
            cudaMalloc(&dOut, chunk);
            cudaStream_t sIn, sOut;
            cudaStreamCreateWithFlags(&sIn, cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&sOut, cudaStreamNonBlocking);
            cudaMemcpyAsync(dIn, hIn, chunk, cudaMemcpyHostToDevice, sIn); // touch everything once
            cudaMemcpyAsync(hOut, dOut, chunk, cudaMemcpyDeviceToHost, sOut);
            cudaDeviceSynchronize();
            __atomic_store_n(&ready[g], 1, __ATOMIC_SEQ_CST);
            while (!go) std::this_thread::yield();
            const auto t0 = std::chrono::steady_clock::now();
            for (int c = 0; c < chunks; c++) {
                if (mode != 1) cudaMemcpyAsync(dIn, hIn, chunk, cudaMemcpyHostToDevice, sIn);
                if (mode != 0) cudaMemcpyAsync(hOut, dOut, chunk, cudaMemcpyDeviceToHost, sOut);
            }
            cudaStreamSynchronize(sIn);
            cudaStreamSynchronize(sOut);
            secs[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            cudaFreeHost(hIn); cudaFreeHost(hOut); cudaFree(dIn); cudaFree(dOut);
            cudaStreamDestroy(sIn); cudaStreamDestroy(sOut);
        });
    for (int g = 0; g < nGpu; g++)
        while (!__atomic_load_n(&ready[g], __ATOMIC_SEQ_CST)) std::this_thread::yield();
    go = 1;
    for (auto &t : th) t.join();
    double worst = 0.0;
    for (double s : secs) worst = s > worst ? s : worst;
    const double bytes = (double)chunk * chunks * nGpu * (mode == 2 ? 2 : 1);
    return bytes / worst / 1e9;
}

int main(int argc, char **argv)
{
    const size_t chunk = (size_t)(argc > 1 ? atoi(argv[1]) : 256) << 20;
    const int chunks = argc > 2 ? atoi(argv[2]) : 16;
    int n = 0;
    cudaGetDeviceCount(&n);
    printf("# pinned host <-> device, %zu MiB per copy, %d copies per GPU and direction, aggregate GB/s (slowest GPU's clock)\n", chunk >> 20, chunks);
    printf("# gpus   H2D    D2H    both directions\n");
    for (int g = 1; g <= n; g *= 2) printf("%6d %6.1f %6.1f %6.1f\n", g, run(g, 0, chunk, chunks), run(g, 1, chunk, chunks), run(g, 2, chunk, chunks));
    return 0;
}
