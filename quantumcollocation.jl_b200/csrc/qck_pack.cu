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

// structure-order policies other than CSC: caller-order array from the canonical one (and, for per-integrator duplicates, from
// the partial columns of the shared Hessian positions)
__global__ void __launch_bounds__(256) qck_reorder_kernel(const double* __restrict__ arr, const double* __restrict__ partial, double* __restrict__ out,
                                                          const int* __restrict__ src, int C, long long nnz, int npart, long long nk) {
    for (long long t = blockIdx.x; t < nk; t += gridDim.x) {
        const double* a = arr + t * nnz;
        double* o = out + t * C;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            const int s = __ldg(src + i);
            o[i] = s < nnz ? a[s] : partial[t * npart + (s - nnz)];
        }
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

// Ensemble sharding, device-resident: the Hessian entries on the shared controls are sums over integrators that live on
// different GPUs.  Every GPU reads the partial columns of ALL GPUs over NVLink (peer pointers) and sums them itself -- an
// all-reduce by all-read, no collective call, one launch.  The contributor list of a position is the concatenation of the
// GPUs' lists in ascending integrator order and lane l adds entries l, l+32, ... before the butterfly: the same order as the
// single-GPU reduce kernel, so the sums are bitwise identical to a one-GPU run.
__global__ void __launch_bounds__(256) qck_peer_reduce_kernel(const QckPeerReduce r, double* __restrict__ H, long long n_knots, long long nnzH) {
    const int lane = threadIdx.x & 31;
    const long long total = n_knots * r.n_shared;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += nwarps) {
        const long long t = i / r.n_shared;
        const int s = (int)(i - t * r.n_shared);
        const int k0 = r.ptr[s], k1 = r.ptr[s + 1];
        double acc = 0.0;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int e = r.cols[k], g = e >> 24, col = e & 0xffffff;
            acc += r.partial[g][t * r.npart[g] + col];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) H[t * nnzH + r.pos[s]] = acc;
    }
}

}  // namespace

int qck_launch_peer_reduce(const QckPeerReduce& R, double* H, long long n_knots, long long nnzH, cudaStream_t stream, int* launches) {
    if (R.n_shared == 0 || n_knots <= 0) return 0;
    const long long total = n_knots * R.n_shared;
    long long grid = (total + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    qck_peer_reduce_kernel<<<(unsigned)grid, 256, 0, stream>>>(R, H, n_knots, nnzH);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_launch_reorder(const double* arr, const double* partial, double* out, const int* src, int C, long long nnz, int npart, long long nk,
                       cudaStream_t stream, int* launches) {
    if (C <= 0 || nk <= 0) return 0;
    qck_reorder_kernel<<<(unsigned)std::min<long long>(nk, 148 * 8), 256, 0, stream>>>(arr, partial, out, src, C, nnz, npart, nk);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

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
