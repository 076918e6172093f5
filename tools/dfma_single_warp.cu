// How fast can ONE warp issue DFMA on sm_100a?  (ILP = 16 independent accumulators, 1 warp per SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps) {
    double* d; long long* c; cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&c, 8);
    int iters = 2000;
    k<ILP><<<148, 32 * warps>>>(d, c, iters);
    k<ILP><<<148, 32 * warps>>>(d, c, iters);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("ILP %2d, %d warp(s)/SM: %.2f cycles per warp-level DFMA (per warp), %.2f cycles per DFMA per SM sub-partition\n", ILP, warps,
           (double)h / (iters * ILP), (double)h / (iters * ILP) / ((warps + 3) / 4));
    cudaFree(d); cudaFree(c);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1); run<16>(1); run<18>(1);
    run<16>(4); run<16>(8); run<16>(16);
    return 0;
}
