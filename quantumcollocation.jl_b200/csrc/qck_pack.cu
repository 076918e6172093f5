// Pack / unpack kernels of the host-buffer path (qck_pipe.cpp) and of the multi-GPU exchange (qck_multi.cpp):
// gather the positions of a knot-major value array that have to travel (every kron(I_N, B) block once; the entries a
// shard owns; the Hessian entries several systems share) into a contiguous buffer, and scatter them back.
#include <algorithm>

#include "qck_internal.h"

namespace {

__global__ void __launch_bounds__(256) qck_pack_kernel(const double* __restrict__ arr, double* __restrict__ out, const int* __restrict__ src,
                                                       int C, long long ostride, long long nnz, long long nk) {
    // one CTA walks whole knots: consecutive threads read (mostly) consecutive positions and write consecutive ones
    for (long long t = blockIdx.x; t < nk; t += gridDim.x) {
        const double* a = arr + t * nnz;
        double* o = out + t * ostride;
        for (int i = threadIdx.x; i < C; i += blockDim.x) o[i] = a[__ldg(src + i)];
    }
}

__global__ void __launch_bounds__(256) qck_unpack_kernel(double* __restrict__ arr, const double* __restrict__ in, const int* __restrict__ src,
                                                         int C, long long ostride, long long nnz, long long nk) {
    for (long long t = blockIdx.x; t < nk; t += gridDim.x) {
        double* a = arr + t * nnz;
        const double* o = in + t * ostride;
        for (int i = threadIdx.x; i < C; i += blockDim.x) a[__ldg(src + i)] = o[i];
    }
}

// few positions per knot (shared Hessian entries): one thread per (knot, position)
__global__ void __launch_bounds__(256) qck_pack_small_kernel(const double* __restrict__ arr, double* __restrict__ out, const int* __restrict__ src,
                                                             int C, long long ostride, long long nnz, long long nk) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nk * C; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / C;
        const int k = (int)(i - t * C);
        out[t * ostride + k] = arr[t * nnz + __ldg(src + k)];
    }
}
__global__ void __launch_bounds__(256) qck_unpack_small_kernel(double* __restrict__ arr, const double* __restrict__ in, const int* __restrict__ src,
                                                               int C, long long ostride, long long nnz, long long nk) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nk * C; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / C;
        const int k = (int)(i - t * C);
        arr[t * nnz + __ldg(src + k)] = in[t * ostride + k];
    }
}

}  // namespace

int qck_launch_pack(const double* arr, double* out, const int* src, int C, long long ostride, long long nnz, long long nk, cudaStream_t stream, int* launches) {
    if (C <= 0 || nk <= 0) return 0;
    if (C < 128) {
        const long long n = nk * C;
        qck_pack_small_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(arr, out, src, C, ostride, nnz, nk);
    } else {
        qck_pack_kernel<<<(unsigned)std::min<long long>(nk, 148 * 8), 256, 0, stream>>>(arr, out, src, C, ostride, nnz, nk);
    }
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_launch_unpack(double* arr, const double* in, const int* src, int C, long long ostride, long long nnz, long long nk, cudaStream_t stream, int* launches) {
    if (C <= 0 || nk <= 0) return 0;
    if (C < 128) {
        const long long n = nk * C;
        qck_unpack_small_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(arr, in, src, C, ostride, nnz, nk);
    } else {
        qck_unpack_kernel<<<(unsigned)std::min<long long>(nk, 148 * 8), 256, 0, stream>>>(arr, in, src, C, ostride, nnz, nk);
    }
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}
