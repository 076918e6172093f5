// Measures the FP64 tensor-core (DMMA, mma.sync f64) peak of the box next to the DFMA peak (tools/fp64_peak.cu).
// Shapes: m8n8k4 (sm_80+), m16n8k4 / m16n8k8 / m16n8k16 (sm_90+).  ILP = independent accumulator chains per warp.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma884(double* out, int iters) {
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 1.0 + i; }
    double a = 1.0000001, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma1688(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int k = 0; k < 4; ++k) c[i][k] = threadIdx.x * 1e-3 + i + k;
    double a0 = 1.0000001, a1 = 0.999, a2 = 1.001, a3 = 0.5, b0 = 1e-9 * threadIdx.x, b1 = 2e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) for (int k = 0; k < 4; ++k) s += c[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma1684(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int k = 0; k < 4; ++k) c[i][k] = threadIdx.x * 1e-3 + i + k;
    double a0 = 1.0000001, a1 = 0.999, b0 = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(b0));
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) for (int k = 0; k < 4; ++k) s += c[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// single-warp dependent chain: latency of one m8n8k4
__global__ void dmma_lat(double* out, int iters, long long* cyc) {
    double c0 = threadIdx.x, c1 = 1.0, a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename K>
static void run(const char* name, K kern, double fma_per_warp_instr, int ilp, int threads, int blocks_per_sm, double* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4000, blocks = 148 * blocks_per_sm;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double warps = (double)blocks * threads / 32;
    double flops = 2.0 * fma_per_warp_instr * ilp * iters * warps;
    printf("%-10s ilp %d  %4d thr x %d/SM: %.3f ms  %.2f TFLOP/s  (%s)\n", name, ilp, threads, blocks_per_sm, best, flops / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
    for (int thr : {128, 256, 512}) {
        run("m8n8k4", dmma884<1>, 256, 1, thr, 2, d);
        run("m8n8k4", dmma884<4>, 256, 4, thr, 2, d);
        run("m8n8k4", dmma884<8>, 256, 8, thr, 2, d);
        run("m16n8k4", dmma1684<4>, 512, 4, thr, 2, d);
        run("m16n8k8", dmma1688<4>, 1024, 4, thr, 2, d);
    }
    run("m8n8k4", dmma884<8>, 256, 8, 1024, 2, d);
    long long* cyc; cudaMalloc(&cyc, 8);
    dmma_lat<<<1, 32>>>(d, 10000, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("m8n8k4 dependent-chain latency: %.1f cycles\n", h / 10000.0);
    return 0;
}
