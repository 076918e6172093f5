// Rollouts on the device: unitary_rollout / rollout / unitary_rollout_fidelity of the reference
// (used for trajectory initialisation, /root/reference/src/trajectory_initialization.jl:426,493, and in every template test's
// assertion, e.g. src/problem_templates/unitary_smooth_pulse_problem.jl:218-220; robustness sweeps over sampled systems,
// unitary_sampling_problem.jl:233-243):
//     X_1 = X_init,   X_{t+1} = exp(dt_t G_s(a_t)) X_t      for every system s of a batch sharing the controls.
// Sequential in t, so it is done as a scan:
//   1. propagators  E_{s,t} = exp(-i H_s(a_t) dt_t): one warp per (system, knot), scaling and squaring of a degree-12 Taylor
//      polynomial (||Y||_1 <= 1/4: truncation 2e-18 relative), matrices in shared memory;
//   2. chunk-local prefix products (one warp per chunk of QCK_RO_CHUNK knots, sequential inside the chunk);
//   3. sequential scan over the chunk totals (one warp per system);
//   4. X_t = (local prefix) (chunk prefix) X_init, written as iso-vectors in trajectory layout (unitary: 2N^2 x T, ket: 2N x T).
// Complex N x N arithmetic, N <= 16.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "qck_internal.h"

#define QCK_RO_CHUNK 32
#define QCK_RO_MAXN 16

namespace {

thread_local std::string g_ro_error;

__device__ __forceinline__ void cfma_(double2& c, double2 a, double2 b) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(a.y, b.x, c.y);
}

// C = A B (N x N, column-major, all in shared memory or global), one warp
__device__ __forceinline__ void warp_mm(double2* C, const double2* A, const double2* B, int N, int lane) {
    for (int e = lane; e < N * N; e += 32) {
        const int r = e % N, c = e / N;
        double2 acc = make_double2(0.0, 0.0);
        for (int k = 0; k < N; ++k) cfma_(acc, A[r + N * k], B[k + N * c]);
        C[e] = acc;
    }
}

// E[s][t] = exp(dt_t * (-i H_s(a_t))), t < T - 1.  One warp per (s, t); 3 N x N buffers per warp in shared memory.
__global__ void __launch_bounds__(128) qck_ro_expm_kernel(const double2* __restrict__ A0, const double2* __restrict__ Aj, int N, int nd, int S,
                                                          long long T, const double* __restrict__ a, const double* __restrict__ dt,
                                                          double2* __restrict__ E) {
    extern __shared__ __align__(16) double2 sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NN = N * N;
    double2* Y = sm + (size_t)w * 3 * NN;
    double2* P = Y + NN;
    double2* Q = P + NN;
    __shared__ double cs[4][QCK_RO_MAXN];
    const long long total = (long long)S * (T - 1), nw = (long long)gridDim.x * 4;
    for (long long item = (long long)blockIdx.x * 4 + w; item < total; item += nw) {
        const int s = (int)(item / (T - 1));
        const long long t = item - (long long)s * (T - 1);
        const double h = dt[t];
        for (int e = lane; e < NN; e += 32) {
            double2 v = A0[(size_t)s * NN + e];
            for (int j = 0; j < nd; ++j) {
                const double aj = a[t * nd + j];
                const double2 d = Aj[((size_t)s * nd + j) * NN + e];
                v.x = fma(aj, d.x, v.x);
                v.y = fma(aj, d.y, v.y);
            }
            Y[e] = make_double2(h * v.x, h * v.y);
        }
        __syncwarp();
        if (lane < N) {  // 1-norm
            double c = 0.0;
            for (int r = 0; r < N; ++r) c += fabs(Y[r + N * lane].x) + fabs(Y[r + N * lane].y);
            cs[w][lane] = c;
        }
        __syncwarp();
        double nrm = 0.0;
        for (int c = 0; c < N; ++c) nrm = fmax(nrm, cs[w][c]);
        int sq = 0;
        while (nrm > 0.25 && sq < 60) { nrm *= 0.5; ++sq; }
        const double sc = ldexp(1.0, -sq);
        for (int e = lane; e < NN; e += 32) {
            Y[e].x *= sc; Y[e].y *= sc;
            const int r = e % N, c = e / N;
            P[e] = make_double2((r == c ? 1.0 : 0.0) + Y[e].x * (1.0 / 12.0), Y[e].y * (1.0 / 12.0));  // Horner start: I + Y/12
        }
        __syncwarp();
        for (int m = 11; m >= 1; --m) {  // P <- I + (Y/m) P
            warp_mm(Q, Y, P, N, lane);
            __syncwarp();
            const double im = 1.0 / m;
            for (int e = lane; e < NN; e += 32) {
                const int r = e % N, c = e / N;
                P[e] = make_double2((r == c ? 1.0 : 0.0) + im * Q[e].x, im * Q[e].y);
            }
            __syncwarp();
        }
        for (int k = 0; k < sq; ++k) {  // E <- E E
            warp_mm(Q, P, P, N, lane);
            __syncwarp();
            for (int e = lane; e < NN; e += 32) P[e] = Q[e];
            __syncwarp();
        }
        double2* out = E + ((size_t)s * (T - 1) + t) * NN;
        for (int e = lane; e < NN; e += 32) out[e] = P[e];
        __syncwarp();
    }
}

// chunk c of system s: L[t] = E_t E_{t-1} ... E_{c0} for t in the chunk (in place), one warp per (s, chunk)
__global__ void __launch_bounds__(128) qck_ro_local_kernel(double2* __restrict__ E, int N, int S, long long nsteps, long long nchunks) {
    extern __shared__ __align__(16) double2 sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NN = N * N;
    double2* acc = sm + (size_t)w * 3 * NN;
    double2* cur = acc + NN;
    double2* tmp = cur + NN;
    const long long total = (long long)S * nchunks, nw = (long long)gridDim.x * 4;
    for (long long item = (long long)blockIdx.x * 4 + w; item < total; item += nw) {
        const int s = (int)(item / nchunks);
        const long long c = item - (long long)s * nchunks, t0 = c * QCK_RO_CHUNK, t1 = min(nsteps, t0 + QCK_RO_CHUNK);
        double2* Es = E + (size_t)s * nsteps * NN;
        for (int e = lane; e < NN; e += 32) acc[e] = Es[(size_t)t0 * NN + e];
        __syncwarp();
        for (long long t = t0 + 1; t < t1; ++t) {
            for (int e = lane; e < NN; e += 32) cur[e] = Es[(size_t)t * NN + e];
            __syncwarp();
            warp_mm(tmp, cur, acc, N, lane);
            __syncwarp();
            for (int e = lane; e < NN; e += 32) { acc[e] = tmp[e]; Es[(size_t)t * NN + e] = tmp[e]; }
            __syncwarp();
        }
    }
}

// prefix over the chunk totals: C[c] = (total of chunk c-1) ... (total of chunk 0), C[0] = I; one warp per system
__global__ void __launch_bounds__(32) qck_ro_scan_kernel(const double2* __restrict__ E, double2* __restrict__ Cp, int N, long long nsteps, long long nchunks) {
    extern __shared__ __align__(16) double2 sm[];
    const int lane = threadIdx.x, NN = N * N, s = blockIdx.x;
    double2* acc = sm;
    double2* cur = acc + NN;
    double2* tmp = cur + NN;
    const double2* Es = E + (size_t)s * nsteps * NN;
    double2* Cs = Cp + (size_t)s * nchunks * NN;
    for (int e = lane; e < NN; e += 32) { acc[e] = make_double2((e % N == e / N) ? 1.0 : 0.0, 0.0); Cs[e] = acc[e]; }
    __syncwarp();
    for (long long c = 1; c < nchunks; ++c) {
        const long long last = min(nsteps, c * QCK_RO_CHUNK) - 1;  // last step of chunk c-1 holds that chunk's total
        for (int e = lane; e < NN; e += 32) cur[e] = Es[(size_t)last * NN + e];
        __syncwarp();
        warp_mm(tmp, cur, acc, N, lane);
        __syncwarp();
        for (int e = lane; e < NN; e += 32) { acc[e] = tmp[e]; Cs[(size_t)c * NN + e] = tmp[e]; }
        __syncwarp();
    }
}

// X_{t+1} = L[t] C[chunk(t)] X_init -> iso-vectors in trajectory layout; X_1 = X_init.  ncols = N (unitary) or 1 (ket).
__global__ void __launch_bounds__(128) qck_ro_apply_kernel(const double2* __restrict__ E, const double2* __restrict__ Cp, const double2* __restrict__ X0,
                                                           int N, int ncols, int S, long long T, long long nchunks, double* __restrict__ out) {
    extern __shared__ __align__(16) double2 sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NN = N * N, nx = N * ncols;
    double2* Pm = sm + (size_t)w * 3 * NN;
    double2* Xc = Pm + NN;
    double2* Xo = Xc + NN;
    const long long nsteps = T - 1, total = (long long)S * T, nw = (long long)gridDim.x * 4;
    const int dim = 2 * nx;
    for (long long item = (long long)blockIdx.x * 4 + w; item < total; item += nw) {
        const int s = (int)(item / T);
        const long long k = item - (long long)s * T;  // knot
        const double2* x0 = X0 + (size_t)s * nx;
        double* o = out + ((size_t)s * T + k) * dim;
        if (k == 0) {
            for (int e = lane; e < nx; e += 32) { const int r = e % N, c = e / N; o[c * 2 * N + r] = x0[e].x; o[c * 2 * N + N + r] = x0[e].y; }
            continue;
        }
        const long long t = k - 1, c = t / QCK_RO_CHUNK;
        // Xc = C[c] X_init, then X = L[t] Xc
        for (int e = lane; e < NN; e += 32) Pm[e] = Cp[((size_t)s * nchunks + c) * NN + e];
        __syncwarp();
        for (int e = lane; e < nx; e += 32) {
            const int r = e % N, cc = e / N;
            double2 acc = make_double2(0.0, 0.0);
            for (int q = 0; q < N; ++q) cfma_(acc, Pm[r + N * q], x0[q + N * cc]);
            Xc[e] = acc;
        }
        __syncwarp();
        for (int e = lane; e < NN; e += 32) Pm[e] = E[((size_t)s * nsteps + t) * NN + e];
        __syncwarp();
        for (int e = lane; e < nx; e += 32) {
            const int r = e % N, cc = e / N;
            double2 acc = make_double2(0.0, 0.0);
            for (int q = 0; q < N; ++q) cfma_(acc, Pm[r + N * q], Xc[q + N * cc]);
            Xo[e] = acc;
        }
        __syncwarp();
        for (int e = lane; e < nx; e += 32) { const int r = e % N, cc = e / N; o[cc * 2 * N + r] = Xo[e].x; o[cc * 2 * N + N + r] = Xo[e].y; }
        __syncwarp();
    }
}

int ro_fail(int code, const std::string& msg) { g_ro_error = msg; return code; }

}  // namespace

extern "C" {

const char* qck_rollout_last_error(void) { return g_ro_error.c_str(); }

// unitary_rollout (ket = 0) / rollout (ket = 1) for n_systems systems that share the controls.
//   H_drift [n_systems][N*N], H_drives [n_systems][n_drives][N*N]: ComplexF64 column-major, interleaved (re, im)
//   a [T][n_drives] (= vec of the n_drives x T control matrix), dt [T] (dt[T-1] unused)
//   X_init [n_systems][2N*ncols] iso-vec(s) (NULL: identity / first basis ket), ncols = N (unitary) or 1 (ket)
//   X_out [n_systems][T][2N*ncols]: per system the dim x T trajectory component, column-major
int qck_rollout(int32_t device, int32_t ket, int32_t levels, int32_t n_drives, int32_t n_systems, const double* H_drift,
                const double* H_drives, int64_t T, const double* a, const double* dt, const double* X_init, double* X_out) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    const int N = levels, nd = n_drives, S = n_systems, NN = N * N, ncols = ket ? 1 : N, nx = N * ncols;
    if (N < 1 || N > QCK_RO_MAXN || nd < 0 || S < 1 || T < 1 || !a || !dt || !X_out || !H_drives) return ro_fail(QCK_EINVAL, "qck_rollout: bad arguments (levels <= 16)");
    cudaError_t ce;
    int ndev = 0;
    if ((ce = cudaGetDeviceCount(&ndev)) != cudaSuccess || device < 0 || device >= ndev) return ro_fail(QCK_ENODEVICE, "qck_rollout: no such CUDA device; libqcknot has no CPU fallback");
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return ro_fail(QCK_ECUDA, cudaGetErrorString(ce));
    // A = -i H
    std::vector<double2> A0((size_t)S * NN, make_double2(0.0, 0.0)), Aj((size_t)S * nd * NN), X0((size_t)S * nx, make_double2(0.0, 0.0));
    for (size_t e = 0; e < A0.size(); ++e)
        if (H_drift) A0[e] = make_double2(H_drift[2 * e + 1], -H_drift[2 * e]);
    for (size_t e = 0; e < Aj.size(); ++e) Aj[e] = make_double2(H_drives[2 * e + 1], -H_drives[2 * e]);
    for (int s = 0; s < S; ++s)
        for (int c = 0; c < ncols; ++c)
            for (int r = 0; r < N; ++r) {
                double2 v = make_double2(r == c ? 1.0 : 0.0, 0.0);
                if (X_init) v = make_double2(X_init[(size_t)s * 2 * nx + c * 2 * N + r], X_init[(size_t)s * 2 * nx + c * 2 * N + N + r]);
                X0[(size_t)s * nx + r + N * c] = v;
            }
    const long long nsteps = T - 1, nchunks = std::max<long long>(1, (nsteps + QCK_RO_CHUNK - 1) / QCK_RO_CHUNK);
    double2 *dA0 = nullptr, *dAj = nullptr, *dX0 = nullptr, *dE = nullptr, *dC = nullptr;
    double *da = nullptr, *ddt = nullptr, *dout = nullptr;
    std::vector<void*> allocs;
    auto alloc = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16)); if (e == cudaSuccess) allocs.push_back(*p); return e; };
    auto cleanup = [&]() { for (void* p : allocs) cudaFree(p); };
#define RO_TRY(call) do { if ((ce = (call)) != cudaSuccess) { cleanup(); return ro_fail(QCK_ECUDA, std::string(#call) + ": " + cudaGetErrorString(ce)); } } while (0)
    RO_TRY(alloc((void**)&dA0, sizeof(double2) * A0.size()));
    RO_TRY(alloc((void**)&dAj, sizeof(double2) * Aj.size()));
    RO_TRY(alloc((void**)&dX0, sizeof(double2) * X0.size()));
    RO_TRY(alloc((void**)&dE, sizeof(double2) * (size_t)S * std::max<long long>(nsteps, 1) * NN));
    RO_TRY(alloc((void**)&dC, sizeof(double2) * (size_t)S * nchunks * NN));
    RO_TRY(alloc((void**)&da, sizeof(double) * (size_t)T * std::max(nd, 1)));
    RO_TRY(alloc((void**)&ddt, sizeof(double) * (size_t)T));
    RO_TRY(alloc((void**)&dout, sizeof(double) * (size_t)S * T * 2 * nx));
    RO_TRY(cudaMemcpy(dA0, A0.data(), sizeof(double2) * A0.size(), cudaMemcpyHostToDevice));
    if (!Aj.empty()) RO_TRY(cudaMemcpy(dAj, Aj.data(), sizeof(double2) * Aj.size(), cudaMemcpyHostToDevice));
    RO_TRY(cudaMemcpy(dX0, X0.data(), sizeof(double2) * X0.size(), cudaMemcpyHostToDevice));
    if (nd) RO_TRY(cudaMemcpy(da, a, sizeof(double) * (size_t)T * nd, cudaMemcpyHostToDevice));
    RO_TRY(cudaMemcpy(ddt, dt, sizeof(double) * (size_t)T, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double2) * 3 * NN * 4;
    RO_TRY(cudaFuncSetAttribute(qck_ro_expm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    RO_TRY(cudaFuncSetAttribute(qck_ro_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    RO_TRY(cudaFuncSetAttribute(qck_ro_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    if (nsteps > 0) {
        const long long items = (long long)S * nsteps;
        qck_ro_expm_kernel<<<(unsigned)std::min<long long>((items + 3) / 4, 148 * 8), 128, smem>>>(dA0, dAj, N, nd, S, T, da, ddt, dE);
        qck_ro_local_kernel<<<(unsigned)std::min<long long>(((long long)S * nchunks + 3) / 4, 148 * 8), 128, smem>>>(dE, N, S, nsteps, nchunks);
        qck_ro_scan_kernel<<<S, 32, sizeof(double2) * 3 * NN>>>(dE, dC, N, nsteps, nchunks);
    }
    qck_ro_apply_kernel<<<(unsigned)std::min<long long>(((long long)S * T + 3) / 4, 148 * 8), 128, smem>>>(dE, dC, dX0, N, ncols, S, T, nchunks, dout);
    RO_TRY(cudaGetLastError());
    RO_TRY(cudaMemcpy(X_out, dout, sizeof(double) * (size_t)S * T * 2 * nx, cudaMemcpyDeviceToHost));
#undef RO_TRY
    cleanup();
    return QCK_OK;
}

}  // extern "C"
