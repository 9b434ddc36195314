// tools/pcie_both.cu -- what bounds the host-buffer (e2e) path at N GPUs: one process per GPU, each copying a 513^3 field (1.08 GB)
// host -> device and another device -> host at the same time from pinned memory, `reps` times; prints this process's GB/s per direction.
// Started for N GPUs at once by tools/pcie_concurrent.sh; the aggregate against N x the single-GPU figure shows the host-side limit.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0, reps = argc > 2 ? atoi(argv[2]) : 10;
    CK(cudaSetDevice(dev));
    const size_t bytes = 513ull * 513 * 513 * 8;
    double *hin, *hout, *d1, *d2;
    CK(cudaMallocHost(&hin, bytes)); CK(cudaMallocHost(&hout, bytes));
    CK(cudaMalloc(&d1, bytes)); CK(cudaMalloc(&d2, bytes));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warm = 0; warm < 2; ++warm) { cudaMemcpyAsync(d1, hin, bytes, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(hout, d2, bytes, cudaMemcpyDeviceToHost, s2); }
    cudaDeviceSynchronize();
    cudaEventRecord(a, s1);
    for (int r = 0; r < reps; ++r) { cudaMemcpyAsync(d1, hin, bytes, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(hout, d2, bytes, cudaMemcpyDeviceToHost, s2); }
    cudaStreamSynchronize(s2); cudaEventRecord(b, s1); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("{\"gpu\": %d, \"reps\": %d, \"ms_per_rep\": %.3f, \"gbs_each_way\": %.1f}\n", dev, reps, ms / reps, bytes / (ms / reps) / 1e6);
    return 0;
}
