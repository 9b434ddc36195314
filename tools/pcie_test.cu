#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    const size_t n0 = 513, n1 = 513, n2 = 513, pitch = 528;
    const size_t dense = n0 * n1 * n2 * 8, pitched = pitch * n1 * n2 * 8;
    double *hin, *hout, *d1, *d2;
    CK(cudaMallocHost(&hin, dense)); CK(cudaMallocHost(&hout, dense));
    CK(cudaMalloc(&d1, pitched)); CK(cudaMalloc(&d2, pitched));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    auto p3 = [&](double* dev, double* host, bool up, size_t z0, size_t z1, cudaStream_t st) {
        cudaMemcpy3DParms p = {};
        cudaPitchedPtr d = make_cudaPitchedPtr(dev + z0 * pitch * n1, pitch * 8, n0, n1);
        cudaPitchedPtr h = make_cudaPitchedPtr(host + z0 * n0 * n1, n0 * 8, n0, n1);
        p.srcPtr = up ? h : d; p.dstPtr = up ? d : h; p.extent = make_cudaExtent(n0 * 8, n1, z1 - z0);
        p.kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        return cudaMemcpy3DAsync(&p, st);
    };
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a, s1); CK(cudaMemcpyAsync(d1, hin, dense, cudaMemcpyHostToDevice, s1)); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("dense H2D            %.2f ms  %.1f GB/s\n", ms, dense / ms / 1e6);
        cudaEventRecord(a, s1); CK(cudaMemcpyAsync(hout, d2, dense, cudaMemcpyDeviceToHost, s1)); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("dense D2H            %.2f ms  %.1f GB/s\n", ms, dense / ms / 1e6);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s1); CK(cudaMemcpyAsync(d1, hin, dense, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, d2, dense, cudaMemcpyDeviceToHost, s2)); cudaStreamSynchronize(s2); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("dense both           %.2f ms  %.1f GB/s total\n", ms, 2 * dense / ms / 1e6);
        cudaEventRecord(a, s1); CK(p3(d1, hin, true, 0, n2, s1)); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("pitched H2D          %.2f ms  %.1f GB/s\n", ms, dense / ms / 1e6);
        cudaEventRecord(a, s1); CK(p3(d2, hout, false, 0, n2, s1)); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("pitched D2H          %.2f ms  %.1f GB/s\n", ms, dense / ms / 1e6);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s1); CK(p3(d1, hin, true, 0, n2, s1)); CK(p3(d2, hout, false, 0, n2, s2)); cudaStreamSynchronize(s2); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("pitched both         %.2f ms  %.1f GB/s total\n", ms, 2 * dense / ms / 1e6);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s1);
        for (int c = 0; c < 16; ++c) { CK(p3(d1, hin, true, n2 * c / 16, n2 * (c + 1) / 16, s1)); CK(p3(d2, hout, false, n2 * c / 16, n2 * (c + 1) / 16, s2)); }
        cudaStreamSynchronize(s2); cudaEventRecord(b, s1); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("pitched both chunked %.2f ms  %.1f GB/s total\n", ms, 2 * dense / ms / 1e6);
    }
    return 0;
}
