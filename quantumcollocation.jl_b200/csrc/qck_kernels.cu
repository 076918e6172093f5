// Hand-written sm_100a kernels of libqcknot.so: per-knot-point dynamics residual, Jacobian values and
// Hessian-of-Lagrangian values, written straight into the solver's fixed-structure value arrays.
//
// Replaces the arithmetic of the reference's per-knot loop
//   for t = 1:T-1:  f(z_t, z_t+1), df(z_t, z_t+1), mu_d2f(z_t, z_t+1, mu_t)   (SURVEY.md section 3.2 steps 3-5,
//   exercised at /root/reference/test/scripts/integrator_test_1qubit.jl:45-52)
// for the integrators built at unitary_smooth_pulse_problem.jl:163-179.
//
// Design (DESIGN.md has the long version)
//  * one CTA per (knot, quantum integrator) work item, persistent grid-stride loop, item order knot-major so
//    concurrently running CTAs write neighbouring segments of the value arrays;
//  * all arithmetic in COMPLEX N x N form (A = -i H(a)), not the reference's real 2N x 2N iso form: half the
//    flops; the iso layout only appears in the load of z_t and in the output maps;
//  * small dense complex products run on the FP64 pipe from shared memory with 3x3 complex register tiles;
//    products with the constant drive matrices use a fixed-width sparse (ELL) form;
//  * every structural nonzero of the knot block has a precomputed (position, scratch slot, sign) map entry, so
//    the "scatter" is a gather from shared memory followed by position-ordered, coalesced 8-byte stores;
//  * Hessian entries that several integrators contribute to (shared controls) go to a partial buffer that a
//    second kernel reduces in fixed integrator order (bitwise run-to-run reproducible, no atomics).
#include <cstdio>

#include "qck_internal.h"

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(double2& c, double2 a, double2 b) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(a.y, b.x, c.y);
}

// C[r0:r0+3, c0:c0+TC] = A[r0:r0+3, 0:N] * B[0:N, c0:c0+TC]; column-major, leading dimension ld (complex elements).
template <int TC>
__device__ __forceinline__ void tile_mm(const double2* __restrict__ A, const double2* __restrict__ B,
                                        double2* __restrict__ C, int N, int ld, int r0, int c0) {
    double2 acc[QCK_TILE][TC];
#pragma unroll
    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = make_double2(0.0, 0.0);
    const double2* a = A + r0;
    const double2* b = B + (size_t)ld * c0;
    for (int k = 0; k < N; ++k) {
        double2 av[QCK_TILE], bv[TC];
#pragma unroll
        for (int i = 0; i < QCK_TILE; ++i) av[i] = a[i + ld * k];
#pragma unroll
        for (int j = 0; j < TC; ++j) bv[j] = b[k + ld * j];
#pragma unroll
        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) cfma(acc[i][j], av[i], bv[j]);
    }
#pragma unroll
    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) C[r0 + i + ld * (c0 + j)] = acc[i][j];
}

struct Prod {
    const double2* A;
    const double2* B;
    double2* C;
};

// out[r, c] = sum_w val[r][w] * X[col[r][w], c]   (fixed-width sparse row format of a constant drive matrix)
__device__ __forceinline__ double2 ell_row(const double2* __restrict__ val, const int* __restrict__ col, int W,
                                           const double2* __restrict__ X, int ld, int r, int c) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < W; ++w) {
        double2 v = __ldg(val + r * W + w);
        int k = __ldg(col + r * W + w);
        cfma(acc, v, X[k + ld * c]);
    }
    return acc;
}

__device__ __forceinline__ double re_dot(const double2* X, const double2* Y, int N, int nc, int ld, int lane) {
    // sum over real elements of Re(conj(X) Y), strided over the 32 lanes of a warp (not yet reduced)
    double s = 0.0;
    int n = N * nc;
    for (int e = lane; e < n; e += 32) {
        int r = e % N, c = e / N;
        double2 x = X[r + ld * c], y = Y[r + ld * c];
        s = fma(x.x, y.x, s);
        s = fma(x.y, y.y, s);
    }
    return s;
}
__device__ __forceinline__ double warp_sum(double s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

__device__ __forceinline__ void do_aux(const QckLaunch& p, long long t, int tid, int nthreads) {
    const QckClassDev& c = p.c;
    const double* zt = p.Z + t * c.zdim;
    for (int k = tid; k < p.n_aux; k += nthreads) {
        QckAux a = p.aux[k];
        if (!((p.mask >> a.out) & 1u)) continue;
        double dt = c.free_time ? zt[c.dt_off] : c.dt_fixed;
        double v;
        switch (a.op) {
            case QAUX_CONST: v = a.c; break;
            case QAUX_NEG_DT: v = -dt; break;
            case QAUX_NEG_Z: v = -zt[a.i0]; break;
            case QAUX_NEG_MU: v = -p.mu[t * c.dyn + a.i0]; break;
            default: v = zt[c.zdim + a.i0] - zt[a.i0] - dt * zt[a.i1]; break;
        }
        if (a.out == 0) p.F[t * c.dyn + a.pos] = v;
        else if (a.out == 1) p.J[t * p.nnzJ + a.pos] = v;
        else if (a.pos < p.nnzH) p.H[t * p.nnzH + a.pos] = v;
        else p.partial[t * p.npart + (a.pos - p.nnzH)] = v;
    }
}

// gather from scratch through the (position, slot, sign) map and store position-ordered
__device__ __forceinline__ void write_map(const double* __restrict__ sm, const uint32_t* __restrict__ pos,
                                          const uint16_t* __restrict__ src, int cnt, double* __restrict__ out,
                                          long long limit, double* __restrict__ partial, int tid, int nthreads) {
    for (int k = tid; k < cnt; k += nthreads) {
        unsigned s = __ldg(src + k);
        unsigned q = __ldg(pos + k);
        double v = sm[s & 0x7fffu];
        if (s & 0x8000u) v = -v;
        if ((long long)q < limit) out[q] = v;
        else partial[q - limit] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pade-4 integrators (UnitaryPadeIntegrator / QuantumStatePadeIntegrator, order 4).
//   F = I + h/2 A + h^2/12 A^2,  B = I - h/2 A + h^2/12 A^2,  residual R = B U1 - F U0,  A = -i H(a), h = dt
// with D = U1-U0, S = U1+U0, M = multipliers as a complex matrix (mu^T vec_iso(R) = Re <M, R>):
//   R       = D - h/2 A S + h^2/12 A (A D)
//   dR/da_j = -h/2 A_j S + h^2/12 (A_j (A D) + A (A_j D))
//   dR/dh   = -1/2 A S + h/6 A (A D)
//   (dF_j)^H M = h/2 A_j^H M + h^2/12 (A^H (A_j^H M) + A_j^H (A^H M)),  (dB_j)^H M: first term negated
//   (dF/dh)^H M = 1/2 A^H M + h/6 A^H (A^H M),                         (dB/dh)^H M: first term negated
//   d2/da_i da_j = h^2/12 Re(<A_i^H M, A_j D> + <A_j^H M, A_i D>)
//   d2/da_j dh   = Re <M, -1/2 A_j S + h/6 (A_j (A D) + A (A_j D))>,   d2/dh2 = 1/6 Re <M, A (A D)>
// 6 + 2 n_d dense N^3 products per knot (vs ~45 in the reference's real-iso formulation).
// ------------------------------------------------------------------------------------------------------------
template <int TC>
__global__ void qck_pade4_kernel(const QckLaunch p) {
    extern __shared__ double sm[];
    const QckClassDev& c = p.c;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int N = c.N, NP = c.NP, nc = c.nc, ncp = c.ncp, nd = c.nd, W = c.W;
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H, needF = p.mask & QCK_EVAL_F;
    const bool needT = needJ || needH;  // first-order drive terms
    const bool free_time = c.free_time;

    double2* SA = reinterpret_cast<double2*>(sm + c.off_A);
    double2* SS = reinterpret_cast<double2*>(sm + c.off_S);
    double* SX = sm + c.off_X;
    const int msa = NP * NP, mss = NP * ncp;  // complex elements per matrix
#define MA(i) (SA + (i) * msa)
#define MS(i) (SS + (i) * mss)
#define MD(j, k) (SS + (QS_FIXED + QD_COUNT * (j) + (k)) * mss)

    for (int i = tid; i < c.scratch_doubles; i += nthreads) sm[i] = 0.0;
    if (tid == 0) SX[QX_ONE] = 1.0;

    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    const int tilesS = (NP / QCK_TILE) * (ncp / TC);
    const int tilesA = (NP / QCK_TILE) * (NP / QCK_TILE);
    const int tcols = ncp / TC;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long t = item / nact;
        const int mi = (int)(item - t * nact);
        const int m = p.member_begin + mi;
        __syncthreads();  // previous item's write-out has finished reading scratch
        const double* zt = p.Z + t * c.zdim;
        const double* zt1 = zt + c.zdim;
        const int soff = c.state_off[m], coff = c.ctrl_off[m], roff = c.row_off[m];
        const double2* cm = c.cmat + (size_t)m * c.cmat_stride;
        const double2* A0 = cm;
        const double2* Adr = cm + N * N;
        const double2* ellv = Adr + nd * N * N;
        const int* ellc = c.ell_col + (size_t)m * c.ell_stride;
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;

        // ---- stage 0: load the knot pair, build A = -i H(a) and A^H ------------------------------------------
        {
            const int dim = 2 * N * nc;
            const double* mut = needH ? p.mu + t * c.dyn + roff : nullptr;
            for (int idx = tid; idx < dim; idx += nthreads) {
                int cc = idx / (2 * N), q = idx - cc * 2 * N;
                int im = q >= N, r = q - im * N;
                double u0 = zt[soff + idx], u1 = zt1[soff + idx];
                int o = 2 * (r + NP * cc) + im;
                reinterpret_cast<double*>(MS(QS_D))[o] = u1 - u0;
                reinterpret_cast<double*>(MS(QS_S))[o] = u1 + u0;
                if (needH) reinterpret_cast<double*>(MS(QS_M))[o] = mut[idx];
            }
            for (int e = tid; e < N * N; e += nthreads) {
                int r = e % N, k = e / N;
                double2 v = __ldg(A0 + e);
                for (int j = 0; j < nd; ++j) {
                    double aj = zt[coff + j];
                    double2 d = __ldg(Adr + j * N * N + e);
                    v.x = fma(aj, d.x, v.x);
                    v.y = fma(aj, d.y, v.y);
                }
                MA(QA_A)[r + NP * k] = v;
                MA(QA_AH)[k + NP * r] = make_double2(v.x, -v.y);
            }
        }
        __syncthreads();

        // ---- stage 1: A2, A D, A S, A^H M (dense);  A_j S, A_j D, A_j^H M (sparse) ----------------------------
        {
            const int nA = needJ ? tilesA : 0;       // A2 only feeds F and B (Jacobian state blocks)
            const int nS = (needH ? 3 : 2) * tilesS;  // AD, AS, (AhM)
            for (int w = tid; w < nA + nS; w += nthreads) {
                if (w < nA) {
                    int tr = w / (NP / QCK_TILE), tcc = w - tr * (NP / QCK_TILE);
                    tile_mm<QCK_TILE>(MA(QA_A), MA(QA_A), MA(QA_A2), N, NP, tr * QCK_TILE, tcc * QCK_TILE);
                } else {
                    int w2 = w - nA;
                    int pi = w2 / tilesS, tl = w2 - pi * tilesS;
                    int tr = tl / tcols, tcc = tl - tr * tcols;
                    const double2* Aop = pi == 2 ? MA(QA_AH) : MA(QA_A);
                    const double2* Bop = pi == 0 ? MS(QS_D) : (pi == 1 ? MS(QS_S) : MS(QS_M));
                    double2* Cop = pi == 0 ? MS(QS_AD) : (pi == 1 ? MS(QS_AS) : MS(QS_AHM));
                    tile_mm<TC>(Aop, Bop, Cop, N, NP, tr * QCK_TILE, tcc * TC);
                }
            }
            if (needT) {
                const int per = N * nc;
                const int nsp = (needH ? 3 : 2) * nd * per;  // P_j = A_j S, Q1_j = A_j D, N1_j = A_j^H M
                for (int w = nthreads - 1 - tid; w < nsp; w += nthreads) {
                    int pj = w / per, e = w - pj * per;
                    int kind = pj / nd, j = pj - kind * nd;
                    int r = e % N, cc = e / N;
                    int adj = kind == 2;
                    const double2* X = kind == 0 ? MS(QS_S) : (kind == 1 ? MS(QS_D) : MS(QS_M));
                    double2 v = ell_row(ellv + (j * 2 + adj) * N * W, ellc + (j * 2 + adj) * N * W, W, X, NP, r, cc);
                    MD(j, kind == 0 ? QD_P : (kind == 1 ? QD_Q1 : QD_N1))[r + NP * cc] = v;
                }
            }
        }
        __syncthreads();

        // ---- stage 2: A(AD), A^H(A^H M), A(A_j D), A^H(A_j^H M) (dense);  A_j(AD), A_j^H(A^H M) (sparse); F, B ----
        {
            const int nP = 1 + (needT ? nd : 0) + (needH ? 1 + nd : 0);
            for (int w = tid; w < nP * tilesS; w += nthreads) {
                int pi = w / tilesS, tl = w - pi * tilesS;
                int tr = tl / tcols, tcc = tl - tr * tcols;
                const double2 *Aop, *Bop;
                double2* Cop;
                if (pi == 0) {
                    Aop = MA(QA_A); Bop = MS(QS_AD); Cop = MS(QS_AAD);
                } else if (pi <= nd && needT) {
                    int j = pi - 1;
                    Aop = MA(QA_A); Bop = MD(j, QD_Q1); Cop = MD(j, QD_AQ1);
                } else if (pi == 1 + nd) {
                    Aop = MA(QA_AH); Bop = MS(QS_AHM); Cop = MS(QS_AHAHM);
                } else {
                    int j = pi - 2 - nd;
                    Aop = MA(QA_AH); Bop = MD(j, QD_N1); Cop = MD(j, QD_AHN1);
                }
                tile_mm<TC>(Aop, Bop, Cop, N, NP, tr * QCK_TILE, tcc * TC);
            }
            if (needT) {
                const int per = N * nc;
                const int nsp = (needH ? 2 : 1) * nd * per;  // Q2_j = A_j (A D), N2_j = A_j^H (A^H M)
                for (int w = nthreads - 1 - tid; w < nsp; w += nthreads) {
                    int pj = w / per, e = w - pj * per;
                    int kind = pj / nd, j = pj - kind * nd;
                    int r = e % N, cc = e / N;
                    const double2* X = kind == 0 ? MS(QS_AD) : MS(QS_AHM);
                    double2 v = ell_row(ellv + (j * 2 + kind) * N * W, ellc + (j * 2 + kind) * N * W, W, X, NP, r, cc);
                    MD(j, kind == 0 ? QD_Q2 : QD_N2)[r + NP * cc] = v;
                }
            }
            if (needJ) {
                const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0);
                for (int e = tid; e < N * N; e += nthreads) {
                    int r = e % N, k = e / N;
                    double2 a = MA(QA_A)[r + NP * k], a2 = MA(QA_A2)[r + NP * k];
                    double id = r == k ? 1.0 : 0.0;
                    MA(QA_F)[r + NP * k] = make_double2(id + c1h * a.x + c2h2 * a2.x, c1h * a.y + c2h2 * a2.y);
                    MA(QA_B)[r + NP * k] = make_double2(id - c1h * a.x + c2h2 * a2.x, -c1h * a.y + c2h2 * a2.y);
                }
            }
        }
        __syncthreads();

        // ---- stage 3a: Lagrangian-weighted scalar second derivatives (one warp per dot product) ---------------
        if (needH) {
            const double c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
            const int npair = nd * (nd + 1) / 2;
            const int ntask = npair + (free_time ? nd + 1 : 0);
            for (int task = warp; task < ntask; task += nwarps) {
                double s;
                int slot;
                if (task < npair) {
                    int i = 0, rem = task;
                    while (rem >= nd - i) { rem -= nd - i; ++i; }
                    int j = i + rem;
                    s = c2h2 * (re_dot(MD(i, QD_N1), MD(j, QD_Q1), N, nc, NP, lane) +
                                re_dot(MD(j, QD_N1), MD(i, QD_Q1), N, nc, NP, lane));
                    slot = qx_haa(nd, i, j);
                } else if (task < npair + nd) {
                    int j = task - npair;
                    s = -0.5 * re_dot(MS(QS_M), MD(j, QD_P), N, nc, NP, lane) +
                        c2h * (re_dot(MS(QS_M), MD(j, QD_Q2), N, nc, NP, lane) +
                               re_dot(MS(QS_M), MD(j, QD_AQ1), N, nc, NP, lane));
                    slot = QX_HAH + j;
                } else {
                    s = (1.0 / 6.0) * re_dot(MS(QS_M), MS(QS_AAD), N, nc, NP, lane);
                    slot = QX_HHH;
                }
                s = warp_sum(s);
                if (lane == 0) SX[slot] = s;
            }
            __syncthreads();
        }

        // ---- stage 3b: assemble the outputs in place ------------------------------------------------------------
        {
            const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
            const int per = N * nc;
            for (int e = tid; e < per; e += nthreads) {
                int r = e % N, cc = e / N;
                int o = r + NP * cc;
                double2 d = MS(QS_D)[o], as = MS(QS_AS)[o], aad = MS(QS_AAD)[o];
                MS(QS_D)[o] = make_double2(d.x - c1h * as.x + c2h2 * aad.x, d.y - c1h * as.y + c2h2 * aad.y);
                MS(QS_AS)[o] = make_double2(-0.5 * as.x + c2h * aad.x, -0.5 * as.y + c2h * aad.y);
                if (needH) {
                    double2 ahm = MS(QS_AHM)[o], ahahm = MS(QS_AHAHM)[o];
                    MS(QS_AHM)[o] = make_double2(-(0.5 * ahm.x + c2h * ahahm.x), -(0.5 * ahm.y + c2h * ahahm.y));
                    MS(QS_AHAHM)[o] = make_double2(-0.5 * ahm.x + c2h * ahahm.x, -0.5 * ahm.y + c2h * ahahm.y);
                }
            }
            if (needT) {
                for (int w = tid; w < nd * per; w += nthreads) {
                    int j = w / per, e = w - j * per;
                    int r = e % N, cc = e / N;
                    int o = r + NP * cc;
                    double2 pj = MD(j, QD_P)[o], q2 = MD(j, QD_Q2)[o], aq1 = MD(j, QD_AQ1)[o];
                    MD(j, QD_P)[o] = make_double2(-c1h * pj.x + c2h2 * (q2.x + aq1.x), -c1h * pj.y + c2h2 * (q2.y + aq1.y));
                    if (needH) {
                        double2 n1 = MD(j, QD_N1)[o], n2 = MD(j, QD_N2)[o], ahn1 = MD(j, QD_AHN1)[o];
                        double xr = n2.x + ahn1.x, xi = n2.y + ahn1.y;
                        MD(j, QD_N2)[o] = make_double2(-(c1h * n1.x + c2h2 * xr), -(c1h * n1.y + c2h2 * xi));
                        MD(j, QD_AHN1)[o] = make_double2(-c1h * n1.x + c2h2 * xr, -c1h * n1.y + c2h2 * xi);
                    }
                }
            }
        }
        __syncthreads();

        // ---- stage 4: write-out ------------------------------------------------------------------------------------
        if (needF) {
            const int dim = 2 * N * nc;
            double* Fo = p.F + t * c.dyn + roff;
            const double* R = reinterpret_cast<const double*>(MS(QS_D));
            for (int idx = tid; idx < dim; idx += nthreads) {
                int cc = idx / (2 * N), q = idx - cc * 2 * N;
                int im = q >= N, r = q - im * N;
                Fo[idx] = R[2 * (r + NP * cc) + im];
            }
        }
        if (needJ)
            write_map(sm, c.posJ + (size_t)m * c.cntJ, c.srcJ + (size_t)m * c.cntJ, c.cntJ, p.J + t * p.nnzJ, p.nnzJ,
                      nullptr, tid, nthreads);
        if (needH)
            write_map(sm, c.posH + (size_t)m * c.cntH, c.srcH + (size_t)m * c.cntH, c.cntH, p.H + t * p.nnzH, p.nnzH,
                      p.partial + t * p.npart, tid, nthreads);
        if (mi == 0 && p.n_aux) do_aux(p, t, tid, nthreads);
    }
#undef MA
#undef MS
#undef MD
}

__global__ void qck_aux_kernel(const QckLaunch p) {
    for (long long t = blockIdx.x; t < p.n_knots; t += gridDim.x) do_aux(p, t, threadIdx.x, blockDim.x);
}

__global__ void qck_reduce_kernel(const QckReduce r, double* __restrict__ H, const double* __restrict__ partial,
                                  long long n_knots, long long nnzH, int npart) {
    long long total = n_knots * r.n_shared;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i / r.n_shared;
        int s = (int)(i - t * r.n_shared);
        double acc = 0.0;
        for (int k = r.ptr[s]; k < r.ptr[s + 1]; ++k) acc += partial[t * npart + r.cols[k]];
        H[t * nnzH + r.pos[s]] = acc;
    }
}

}  // namespace

// shared by the host map builder: where each output lives inside the CTA scratch
void qck_scratch_layout(QckClassDev& c, int eval_hessian) {
    (void)eval_hessian;
    c.msa = 2 * c.NP * c.NP;
    c.mss = 2 * c.NP * c.ncp;
    c.off_A = 0;
    int n_s = QS_FIXED + QD_COUNT * c.nd;
    c.off_S = c.off_A + QA_COUNT * c.msa;
    c.off_X = c.off_S + n_s * c.mss;
    c.scratch_doubles = c.off_X + QX_HAH + c.nd + c.nd * c.nd;
    c.scratch_doubles = (c.scratch_doubles + 1) & ~1;
}

static int pick_threads(const QckClassDev& c, int tc) {
    int tilesS = (c.NP / QCK_TILE) * (c.ncp / tc);
    int items = (2 + 2 * c.nd) * tilesS;
    int th = ((items + 31) / 32) * 32;
    if (th < 64) th = 64;
    if (th > 256) th = 256;
    return th;
}

int qck_launch_quantum(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches) {
    const QckClassDev& c = L.c;
    size_t smem = (size_t)c.scratch_doubles * sizeof(double);
    long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    if (n_items <= 0) return 0;
    const bool unitary = c.kind == QCK_UNITARY_PADE || c.kind == QCK_UNITARY_EXP;
    void (*kern)(const QckLaunch) = nullptr;
    int tc = unitary ? QCK_TILE : 1;
    if (c.kind == QCK_UNITARY_PADE && c.order == 4) kern = qck_pade4_kernel<QCK_TILE>;
    else if (c.kind == QCK_KET_PADE && c.order == 4) kern = qck_pade4_kernel<1>;
    else return (int)cudaErrorNotSupported;
    int threads = pick_threads(c, tc);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_items) grid = n_items;
    kern<<<(unsigned)grid, threads, smem, stream>>>(L);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_launch_aux(const QckLaunch& L, cudaStream_t stream, int* launches) {
    if (L.n_aux == 0 || L.n_knots <= 0) return 0;
    long long grid = L.n_knots < 4096 ? L.n_knots : 4096;
    qck_aux_kernel<<<(unsigned)grid, 64, 0, stream>>>(L);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_launch_reduce(const QckReduce& R, double* H, const double* partial, long long n_knots, long long nnzH,
                      int npart, cudaStream_t stream, int* launches) {
    if (R.n_shared == 0 || n_knots <= 0) return 0;
    long long total = n_knots * R.n_shared;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    qck_reduce_kernel<<<(unsigned)grid, 256, 0, stream>>>(R, H, partial, n_knots, nnzH, npart);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}
