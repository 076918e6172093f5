// Measures the FP64 FMA peak (DFMA pipe) and a write-only / copy HBM stream on the box.
// Output feeds profiles/ and the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void write_kernel(double2* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = make_double2(1.0, 2.0);
}
__global__ void copy_kernel(double2* out, const double2* in, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs %d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    double* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int rep = 0; rep < 4; ++rep) {
        int blocks = p.multiProcessorCount * 4, threads = 512;
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("dfma rep %d: %.3f ms  %.2f TFLOP/s\n", rep, ms, flops / ms * 1e-9);
    }
    size_t n = (size_t)1 << 28;  // 4 GiB of double2
    double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
    cudaMemset(a, 0, n * 16);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        write_kernel<<<p.multiProcessorCount * 16, 512>>>(b, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("write-only rep %d: %.3f ms  %.1f GB/s\n", rep, ms, n * 16.0 / ms * 1e-6);
    }
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        copy_kernel<<<p.multiProcessorCount * 16, 512>>>(b, a, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy rep %d: %.3f ms  %.1f GB/s (read+write)\n", rep, ms, 2 * n * 16.0 / ms * 1e-6);
    }
    // pinned D2H bandwidth
    void* h; size_t hb = (size_t)1 << 30; cudaMallocHost(&h, hb);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(h, a, hb, cudaMemcpyDeviceToHost);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("D2H pinned rep %d: %.1f GB/s\n", rep, hb / ms * 1e-6);
        cudaEventRecord(e0);
        cudaMemcpyAsync(a, h, hb, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("H2D pinned rep %d: %.1f GB/s\n", rep, hb / ms * 1e-6);
    }
    return 0;
}
