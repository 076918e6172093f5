// Micro-benchmark: shared -> global write-out of a 68 KB "knot block" per iteration from a 26 KB image:
//   mode 0: 16-byte st.global loops (4 warps)         mode 1: cp.async.bulk by lane 0 of every warp
//   mode 2: cp.async.bulk by thread 0 only            mode 3: bulk, lanes 0..k of every warp issue one copy each
// 4 CTAs/SM x 128 threads like the quantum kernel; reports GB/s and issue cycles per iteration (thread 0).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(128, 4) k(double* out, int iters, int mode, int chunk_doubles, int delay, long long* cyc) {
    extern __shared__ __align__(16) double img[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = 8487 & ~1;  // doubles per iteration
    const int nchunks = (total + chunk_doubles - 1) / chunk_doubles;
    for (int i = tid; i < 3300; i += 128) img[i] = i;
    __syncthreads();
    long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        double* dst = out + ((size_t)(blockIdx.x * iters + it)) * total;
        // "compute": touch the image
        if (mode != 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        __syncthreads();
        for (int i = tid; i < 3300; i += 128) img[i] += 1.0;
        for (int d = 0; d < delay; ++d) asm volatile("nanosleep.u32 20;\n");
        if (mode != 0) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        long long t0 = clock64();
        if (mode == 0) {
            // each warp a quarter of the range, 16-byte stores, source wraps over the image
            const int per = (total / 2 + 3) / 4;
            const double2* s2 = reinterpret_cast<const double2*>(img);
            double2* d2 = reinterpret_cast<double2*>(dst);
            for (int k = warp * per + lane; k < (warp + 1) * per && k < total / 2; k += 32) d2[k] = s2[k % 1600];
        } else {
            for (int c = 0; c < nchunks; ++c) {
                const int off = c * chunk_doubles, n = min(chunk_doubles, total - off);
                bool mine = mode == 1 ? (lane == 0 && (c & 3) == warp) : (mode == 2 ? tid == 0 : ((c & 3) == warp && lane == ((c >> 2) & 31)));
                if (mine) bulk_store(dst + off, img + (off % 1600 & ~1), n * 8);
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        acc += clock64() - t0;
    }
    if (mode != 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    if (tid == 0 && blockIdx.x == 0) *cyc = acc / iters;
}
int main() {
    const int iters = 17, blocks = 148 * 4, total = 8486;
    double* out; cudaMalloc(&out, (size_t)blocks * iters * total * 8);
    long long* cyc; cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 56000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int delay : {0, 100}) for (int mode = 0; mode < 4; ++mode) for (int chunk : {162, 324, 1296, 4244}) {
        if (mode == 0 && chunk != 324) continue;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k<<<blocks, 128, 56000>>>(out, iters, mode, chunk, delay, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("delay %3d mode %d chunk %5d doubles: %.1f us  %.0f GB/s  issue %lld cyc/iter  (%s)\n", delay, mode, chunk, best * 1e3,
               (double)blocks * iters * total * 8 / best * 1e-6, h, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
