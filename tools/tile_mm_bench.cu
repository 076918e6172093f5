// Single-warp latency of the 3x3 complex register-tile product from shared memory, in the variants the kernels could use.
#include <cstdio>
#include <cuda_runtime.h>
#define T 3
__device__ __forceinline__ void cfma(double2& c, double2 a, double2 b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x); c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
// variant 0: runtime strides + conj multipliers (what qck_kernels.cu does), unroll U
template <int U>
__device__ __forceinline__ void mm_runtime(const double2* A, bool tA, const double2* B, bool tB, int K, int ld, int r0, int c0, double2 (&acc)[T][T]) {
    for (int i = 0; i < T; ++i) for (int j = 0; j < T; ++j) acc[i][j] = make_double2(0, 0);
    const int ar = tA ? ld : 1, ak = tA ? 1 : ld, bc = tB ? 1 : ld, bk = tB ? ld : 1;
    const double sa = tA ? -1.0 : 1.0, sb = tB ? -1.0 : 1.0;
    const double2* a = A + r0 * ar; const double2* b = B + c0 * bc;
#pragma unroll U
    for (int k = 0; k < K; ++k) {
        double2 av[T], bv[T];
#pragma unroll
        for (int i = 0; i < T; ++i) { av[i] = a[i * ar + k * ak]; av[i].y *= sa; }
#pragma unroll
        for (int j = 0; j < T; ++j) { bv[j] = b[j * bc + k * bk]; bv[j].y *= sb; }
#pragma unroll
        for (int i = 0; i < T; ++i)
#pragma unroll
            for (int j = 0; j < T; ++j) cfma(acc[i][j], av[i], bv[j]);
    }
}
// variant 1: compile-time layout (no transposition), K and ld compile-time, full unroll
template <int N, int LD>
__device__ __forceinline__ void mm_static(const double2* A, const double2* B, int r0, int c0, double2 (&acc)[T][T]) {
    for (int i = 0; i < T; ++i) for (int j = 0; j < T; ++j) acc[i][j] = make_double2(0, 0);
    const double2* a = A + r0; const double2* b = B + c0 * LD;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double2 av[T], bv[T];
#pragma unroll
        for (int i = 0; i < T; ++i) av[i] = a[i + k * LD];
#pragma unroll
        for (int j = 0; j < T; ++j) bv[j] = b[j * LD + k];
#pragma unroll
        for (int i = 0; i < T; ++i)
#pragma unroll
            for (int j = 0; j < T; ++j) cfma(acc[i][j], av[i], bv[j]);
    }
}
template <int V>
__global__ void bench(double* out, long long* cyc, int reps, int N, int ld, int flagA) {
    extern __shared__ double2 sm[];
    double2* A = sm; double2* B = sm + 128; double2* C = sm + 256;
    for (int i = threadIdx.x; i < 384; i += blockDim.x) sm[i] = make_double2(0.001 * i, 0.002 * i);
    __syncthreads();
    const int lane = threadIdx.x;
    const int tile = lane % 9, r0 = (tile / 3) * 3, c0 = (tile % 3) * 3;
    double2 acc[T][T];
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (V == 0) mm_runtime<3>(A, flagA & 1, B, flagA & 2, N, ld, r0, c0, acc);
        else if (V == 1) mm_runtime<9>(A, flagA & 1, B, flagA & 2, N, ld, r0, c0, acc);
        else if (V == 2) mm_runtime<1>(A, flagA & 1, B, flagA & 2, N, ld, r0, c0, acc);
        else mm_static<9, 9>(A, B, r0, c0, acc);
        for (int i = 0; i < T; ++i) for (int j = 0; j < T; ++j) C[r0 + i + ld * (c0 + j)] = acc[i][j];
        __syncwarp();
        double2* tmp = A; A = C; C = tmp;  // make repetitions dependent
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc[0][0].x + acc[2][2].y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int V>
void run(const char* name, int flagA) {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    int reps = 200;
    bench<V><<<1, 32, 384 * 16>>>(d, c, reps, 9, 9, flagA);
    bench<V><<<1, 32, 384 * 16>>>(d, c, reps, 9, 9, flagA);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s flags %d: %.0f cycles per 9x9x9 tile pass (324 DFMA/thread -> floor ~690)\n", name, flagA, (double)h / reps);
}
int main() {
    run<0>("runtime strides, conj multipliers, unroll 3", 0);
    run<0>("runtime strides, conj multipliers, unroll 3", 1);
    run<1>("runtime strides, conj multipliers, unroll 9", 0);
    run<2>("runtime strides, conj multipliers, unroll 1", 0);
    run<3>("compile-time layout, full unroll", 0);
    return 0;
}
