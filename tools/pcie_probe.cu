// Developer probe: pinned H2D and D2H bandwidth alone and concurrently (two streams).
//   nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu && ./tools/pcie_probe
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t n = (size_t)4 << 30;
    void *h1, *h2, *d1, *d2;
    cudaHostAlloc(&h1, n, cudaHostAllocDefault); cudaHostAlloc(&h2, n, cudaHostAllocDefault);
    cudaMalloc(&d1, n); cudaMalloc(&d2, n);
    cudaStream_t a, b; cudaStreamCreate(&a); cudaStreamCreate(&b);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, 0);
            cudaStreamWaitEvent(a, e0, 0); cudaStreamWaitEvent(b, e0, 0);
            if (mode != 1) cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, a);
            if (mode != 0) cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, b);
            cudaDeviceSynchronize();
            cudaEventRecord(e1, 0); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const char* name[3] = {"H2D alone", "D2H alone", "H2D + D2H concurrently (each direction)"};
        printf("%s: %.1f GB/s\n", name[mode], n / (ms * 1e-3) / 1e9);
    }
    return 0;
}
