// Large-level kernel of libqcknot.so: Pade-4 unitaries and kets whose output image does not fit one CTA's shared memory
// (levels up to 32: north_star "qudit dimensions up to about 32"; also 16 levels with four drives).
//
// One CTA of 256 threads per (knot, integrator).  Only the OPERANDS live in shared memory -- A = -i H(a) and the state-shaped
// matrices D = U1 - U0, S = U1 + U0, M (multipliers), A D, V = -h/2 S + h^2/12 A D, A^H M and two scratch matrices (9 N x N
// complex matrices: 144 KB at N = 32) -- every output quantity leaves for the value arrays straight from the epilogue of the
// product that completes it, at the destination the host's placement pass found for it (qdst: first position of every
// quantity inside the knot block, the same table the column kernel uses).  The kron(I_N, .) blocks -iso(F), iso(B) are formed
// once in shared memory (2N x 2N reals) and copied N times with 16-byte stores by the whole CTA.
//
// Algebra in matrix form (same as the row-slice kernel):
//   R = D + A V,  V = -h/2 S + h^2/12 A D                 d/dh = A V',  V' = -1/2 S + h/6 A D
//   d/da_j = A_j V + h^2/12 A (A_j D)                     F / B = I +- h/2 A + h^2/12 A^2
//   state x dt: -(1/2 W1 +- h/6 A^H W1), W1 = A^H M       state x a_j: -(h/2 Z1 +- h^2/12 (A_j^H W1 + A^H Z1)), Z1 = A_j^H M
//   dt x dt = 1/6 Re<W1, A D>    a_j x dt = -1/2 Re<Z1_j, S> + h/6 (Re<Z1_j, A D> + Re<W1, A_j D>)
//   a_i x a_j = h^2/12 Re tr({A_i, A_j} G),  G = D M^H
// 6 + 2 n_d dense N x N x nc products (+ A^2 for the blocks: iso(B) follows from -iso(F) in place; + G), register tiles of
// 2 x 2 complex, drives as sparse rows; two vertically adjacent results leave as one 16-byte store per part.
#include <algorithm>

#include "qck_device.cuh"

namespace {

__device__ __forceinline__ double big_block_sum(double v, double* sh) {  // result valid on every thread
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    return s;
}

// C[r][c] = sum_k op(A)[r][k] B[k][c],  op = conjugate transpose when adj == 1; adj == 2: A is anti-Hermitian, A^H = -A, so the
// product runs on the plain (bank-conflict-free) loads and is negated (the transposed loads of a column-major A put all the
// lanes of a quarter-warp on one bank: 7.7-way conflicts in the first profile).  A: N x N, B: N x nc, column-major (ld = N).
// The epilogue gets two vertically adjacent results at a time: epi(r, c, C[r][c], C[r+1][c], second row valid).
template <class Epi>
__device__ __forceinline__ void big_mm(const double2* __restrict__ A, int adj, const double2* __restrict__ B, int N, int nc, Epi epi) {
    const int ntr = (N + 1) >> 1, ntc = (nc + 1) >> 1;
    for (int tile = threadIdx.x; tile < ntr * ntc; tile += blockDim.x) {
        const int r0 = 2 * (tile % ntr), c0 = 2 * (tile / ntr);
        const bool r1ok = r0 + 1 < N, c1ok = c0 + 1 < nc;
        double2 a00 = make_double2(0.0, 0.0), a01 = a00, a10 = a00, a11 = a00;
        const int r1 = r1ok ? r0 + 1 : r0, c1 = c1ok ? c0 + 1 : c0;
        for (int k = 0; k < N; ++k) {
            double2 x0, x1;
            if (adj == 1) { x0 = A[k + N * r0]; x0.y = -x0.y; x1 = A[k + N * r1]; x1.y = -x1.y; }
            else { x0 = A[r0 + N * k]; x1 = A[r1 + N * k]; }
            const double2 y0 = B[k + N * c0], y1 = B[k + N * c1];
            cfma(a00, x0, y0); cfma(a01, x0, y1); cfma(a10, x1, y0); cfma(a11, x1, y1);
        }
        if (adj == 2) {
            a00 = make_double2(-a00.x, -a00.y); a01 = make_double2(-a01.x, -a01.y);
            a10 = make_double2(-a10.x, -a10.y); a11 = make_double2(-a11.x, -a11.y);
        }
        epi(r0, c0, a00, a10, r1ok);  // rows r0, r0 + 1 of one column: neighbours in every iso-vector quantity
        if (c1ok) epi(r0, c1, a01, a11, r1ok);
    }
}

template <bool KET>
__global__ void __launch_bounds__(256) qck_big_kernel(const QckLaunch p) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int N = c.N, nc = KET ? 1 : N, NN = N * N, NS = N * nc, n2 = 2 * N, nd = c.nd, W = c.W;
    const int tid = threadIdx.x;
    double2* const mA = reinterpret_cast<double2*>(smem_all);
    double2* const mD = mA + NN;
    double2* const mS = mD + NS;
    double2* const mM = mS + NS;
    double2* const mX2 = mM + NS;
    double2* const mV = mX2 + NS;
    double2* const mW1 = mV + NS;
    double2* const T1 = mW1 + NS;        // N x N scratch
    double2* const T2 = T1 + NN;         // N x N scratch (T1 | T2 = one 2N x 2N real block)
    double* const red = reinterpret_cast<double*>(T2 + NN);
    const bool needF = p.mask & QCK_EVAL_F, needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const bool free_time = c.free_time;
    const int adjH = c.antiherm ? 2 : 1;  // A^H products: Hermitian Hamiltonians take the negated plain product
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    auto rdot = [](double2 x, double2 y) { return x.x * y.x + x.y * y.y; };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long t = item / nact;
        const int mi = (int)(item - t * nact), m = p.member_begin + mi;
        const int soff = c.moff[3 * m], coff = c.moff[3 * m + 1], roff = c.moff[3 * m + 2];
        const double* zt = p.Z + t * c.zdim;
        const int* qd = c.qdst + (size_t)m * QO_COUNT;
        const double2* gv = c.cmat + (size_t)m * c.cmat_stride;     // [A0 | ell values | anticommutator values | contributor values]
        const int* gc = c.ell_col + (size_t)m * c.icon_stride;      // [ell columns | kk_ptr | kk_rc | ac_ptr | ac_j]
        const double2* ellv = gv + NN;
        const double2* kkv = ellv + c.ell_stride;
        const double2* acv = kkv + c.kk_cap;
        const int* ellc = gc;
        const int* kkptr = gc + c.ell_stride;
        const int npair = nd * (nd + 1) / 2;
        const int* kkrc = kkptr + npair + 1;
        const int* acptr = kkrc + c.kk_cap;
        const int* acj = acptr + NN + 1;
        double* const oF = p.F + t * c.dyn;
        double* const oJ = p.J + t * p.nnzJ;
        double* const oH = p.H + t * p.nnzH;
        double* const oP = p.partial + t * p.npart;
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
        // iso-vector quantity q, element (r, col): real part at (col * 2N + r), imaginary part N further
        // rows r, r + 1 of one column: two consecutive reals and two consecutive imaginaries -> 16-byte stores where aligned
        auto put = [&](double* arr0, int d0, int q, int r, int col, double2 v0, double2 v1, bool two) {
            const int st = c.pl_stride[q], i = col * n2 + r;
            double* pr = arr0 + d0 + i * st;
            double* pi = arr0 + d0 + (i + N) * st;
            if (two && st == 1 && !(reinterpret_cast<uintptr_t>(pr) & 15) && !(reinterpret_cast<uintptr_t>(pi) & 15)) {
                *reinterpret_cast<double2*>(pr) = make_double2(v0.x, v1.x);
                *reinterpret_cast<double2*>(pi) = make_double2(v0.y, v1.y);
            } else {
                pr[0] = v0.x; pi[0] = v0.y;
                if (two) { pr[st] = v1.x; pi[st] = v1.y; }
            }
        };
        auto putJ = [&](int q, int r, int col, double2 v0, double2 v1, bool two) { const int d0 = qd[q]; if (d0 >= 0) put(oJ, d0, q, r, col, v0, v1, two); };
        auto putH = [&](int q, int r, int col, double2 v0, double2 v1, bool two) {
            const int d0 = qd[q];
            if (d0 < 0) return;
            if (d0 < p.nnzH) put(oH, d0, q, r, col, v0, v1, two);
            else put(oP, d0 - (int)p.nnzH, q, r, col, v0, v1, two);
        };
        auto put_scalar = [&](int q, double v) {
            const int d0 = qd[q];
            if (d0 < 0) return;
            if (d0 < p.nnzH) oH[d0] = v;
            else oP[d0 - p.nnzH] = v;
        };
        // sparse row r of A_j (adj = 0) or A_j^H (adj = 1) times column col of X
        auto drive_elem = [&](int j, int adj, const double2* X, int r, int col) {
            double2 acc = make_double2(0.0, 0.0);
            const int o = ((j * 2 + adj) * N + r) * W;
            for (int w = 0; w < W; ++w) {
                const double2 v = __ldg(ellv + o + w);
                if (v.x != 0.0 || v.y != 0.0) cfma(acc, v, X[__ldg(ellc + o + w) + N * col]);
            }
            return acc;
        };

        __syncthreads();  // the previous item's readers of the operand matrices are through
        // ---- S0: operands ---------------------------------------------------------------------------------------------------
        for (int e = tid; e < NS; e += blockDim.x) {
            const int r = e % N, col = e / N;
            const double u0r = zt[soff + col * n2 + r], u0i = zt[soff + col * n2 + N + r];
            const double u1r = zt[c.zdim + soff + col * n2 + r], u1i = zt[c.zdim + soff + col * n2 + N + r];
            mD[e] = make_double2(u1r - u0r, u1i - u0i);
            mS[e] = make_double2(u1r + u0r, u1i + u0i);
            if (needH) mM[e] = make_double2(p.mu[t * c.dyn + roff + col * n2 + r], p.mu[t * c.dyn + roff + col * n2 + N + r]);
        }
        for (int e = tid; e < NN; e += blockDim.x) {  // A = A0 + sum_j a_j A_j (per-element contributor lists), column-major
            double2 v = __ldg(gv + e);
            for (int u = __ldg(acptr + e); u < __ldg(acptr + e + 1); ++u) {
                const double aj = zt[coff + __ldg(acj + u)];
                const double2 dv = __ldg(acv + u);
                v.x = fma(aj, dv.x, v.x);
                v.y = fma(aj, dv.y, v.y);
            }
            mA[e] = v;
        }
        if (mi == 0 && p.n_aux) do_aux(p, t, tid, blockDim.x);  // derivative-integrator entries of this knot
        __syncthreads();
        // ---- S1: A D (-> V, V'), A^H M ------------------------------------------------------------------------------------------
        big_mm(mA, false, mD, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
            for (int i = 0; i < (two ? 2 : 1); ++i) {
                const int e = r + i + N * col;
                const double2 acc = i ? a1 : a0, s = mS[e];
                mX2[e] = acc;
                mV[e] = make_double2(-c1h * s.x + c2h2 * acc.x, -c1h * s.y + c2h2 * acc.y);
                T1[e] = make_double2(-0.5 * s.x + c2h * acc.x, -0.5 * s.y + c2h * acc.y);
            }
        });
        if (needH) big_mm(mA, adjH, mM, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
            mW1[r + N * col] = a0;
            if (two) mW1[r + 1 + N * col] = a1;
        });
        __syncthreads();
        // ---- S2: residual and d/dh --------------------------------------------------------------------------------------------------
        if (needF) {
            const int d0 = qd[QO_R];
            big_mm(mA, false, mV, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
                const double2 d0v = mD[r + N * col], d1v = two ? mD[r + 1 + N * col] : make_double2(0.0, 0.0);
                if (d0 >= 0) put(oF, d0, QO_R, r, col, make_double2(d0v.x + a0.x, d0v.y + a0.y), make_double2(d1v.x + a1.x, d1v.y + a1.y), two);
            });
        }
        if (needJ && free_time) big_mm(mA, false, T1, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) { putJ(QO_TH, r, col, a0, a1, two); });
        if (needH && free_time) {
            double sloc = 0.0;
            for (int e = tid; e < NS; e += blockDim.x) sloc += rdot(mW1[e], mX2[e]);
            const double tot = big_block_sum(sloc, red);
            if (tid == 0) put_scalar(QO_HHH, tot * (1.0 / 6.0));
        }
        __syncthreads();
        // ---- S3: the kron blocks -iso(F), +iso(B) (A^2 once per block: no spare N x N matrix to keep it) ----------------------------------
        if (needJ) {
            double* const blk = reinterpret_cast<double*>(T1);  // 2N x 2N reals, column-major: [variable column][row]
            const int bl = n2 * n2, ncopy = KET ? 1 : N;
            auto copy_out = [&](int q) {  // the block N times, 16-byte stores by the whole CTA
                const int d0 = qd[q];
                if (d0 < 0) return;
                double* dst0 = oJ + d0;
                if ((reinterpret_cast<uintptr_t>(dst0) & 15) == 0) {
                    const double2* s2 = reinterpret_cast<const double2*>(blk);
                    for (int cb = 0; cb < ncopy; ++cb) {
                        double2* d2 = reinterpret_cast<double2*>(dst0 + (size_t)cb * bl);
                        for (int i = tid; i < bl / 2; i += blockDim.x) d2[i] = s2[i];
                    }
                } else {
                    for (int cb = 0; cb < ncopy; ++cb)
                        for (int i = tid; i < bl; i += blockDim.x) dst0[(size_t)cb * bl + i] = blk[i];
                }
            };
            // -iso(F), F = I + h/2 A + h^2/12 A^2;  iso(X) = [Re X, -Im X; Im X, Re X]: column c holds (Re; Im), column c + N (-Im; Re)
            big_mm(mA, false, mA, N, N, [&](int r, int col, double2 a0, double2 a1, bool two) {
                for (int i = 0; i < (two ? 2 : 1); ++i) {
                    const int rr = r + i;
                    const double2 acc = i ? a1 : a0, a = mA[rr + N * col];
                    const double xr = (rr == col ? 1.0 : 0.0) + c1h * a.x + c2h2 * acc.x, xi = c1h * a.y + c2h2 * acc.y;
                    blk[rr + n2 * col] = -xr;
                    blk[rr + N + n2 * col] = -xi;
                    blk[rr + n2 * (col + N)] = xi;
                    blk[rr + N + n2 * (col + N)] = -xr;
                }
            });
            __syncthreads();
            copy_out(QO_ISOF);
            __syncthreads();
            // +iso(B), B = F - h A:  iso(B) = -(-iso(F)) - h iso(A), in place
            for (int e = tid; e < NN; e += blockDim.x) {
                const int rr = e % N, col = e / N;
                const double2 a = mA[e];
                const double hr = h * a.x, hi = h * a.y;
                blk[rr + n2 * col] = -blk[rr + n2 * col] - hr;
                blk[rr + N + n2 * col] = -blk[rr + N + n2 * col] - hi;
                blk[rr + n2 * (col + N)] = -blk[rr + n2 * (col + N)] + hi;
                blk[rr + N + n2 * (col + N)] = -blk[rr + N + n2 * (col + N)] - hr;
            }
            __syncthreads();
            copy_out(QO_ISOB);
            __syncthreads();
        }
        // ---- S4: d/da_j ---------------------------------------------------------------------------------------------------------------
        for (int j = 0; j < nd && (needJ || needH); ++j) {
            double s_ah = 0.0;
            for (int e = tid; e < NS; e += blockDim.x) {  // U_j = A_j D -> T1,  A_j V -> T2
                const int r = e % N, col = e / N;
                const double2 u = drive_elem(j, 0, mD, r, col);
                T1[e] = u;
                if (needJ) T2[e] = drive_elem(j, 0, mV, r, col);
                if (needH) s_ah += c2h * rdot(mW1[e], u);
            }
            __syncthreads();
            if (needJ) big_mm(mA, false, T1, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
                const double2 y0 = T2[r + N * col], y1 = two ? T2[r + 1 + N * col] : make_double2(0.0, 0.0);
                putJ(QO_TA + j, r, col, make_double2(y0.x + c2h2 * a0.x, y0.y + c2h2 * a0.y), make_double2(y1.x + c2h2 * a1.x, y1.y + c2h2 * a1.y), two);
            });
            __syncthreads();
            if (needH) {  // Z1_j = A_j^H M -> T1,  A_j^H W1 -> T2; state x a_j blocks; a_j x dt
                for (int e = tid; e < NS; e += blockDim.x) {
                    const int r = e % N, col = e / N;
                    const double2 z1 = drive_elem(j, 1, mM, r, col);
                    T1[e] = z1;
                    T2[e] = drive_elem(j, 1, mW1, r, col);
                    s_ah += -0.5 * rdot(z1, mS[e]) + c2h * rdot(z1, mX2[e]);
                }
                __syncthreads();
                big_mm(mA, adjH, T1, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
                    double2 k0[2], k1[2];
                    for (int i = 0; i < (two ? 2 : 1); ++i) {
                        const int e = r + i + N * col;
                        const double2 acc = i ? a1 : a0, z1 = T1[e], z2 = T2[e];
                        const double cr = c2h2 * (z2.x + acc.x), ci = c2h2 * (z2.y + acc.y);
                        k0[i] = make_double2(-c1h * z1.x - cr, -c1h * z1.y - ci);
                        k1[i] = make_double2(-c1h * z1.x + cr, -c1h * z1.y + ci);
                    }
                    putH(QO_KA0 + j, r, col, k0[0], k0[1], two);
                    putH(QO_KA1 + j, r, col, k1[0], k1[1], two);
                });
                if (free_time) {
                    const double tot = big_block_sum(s_ah, red);
                    if (tid == 0) put_scalar(QO_HAH + j, tot);
                }
                __syncthreads();
            }
        }
        // ---- S5: state x dt, a_i x a_j ------------------------------------------------------------------------------------------------------
        if (needH) {
            if (free_time) big_mm(mA, adjH, mW1, N, nc, [&](int r, int col, double2 a0, double2 a1, bool two) {
                double2 k0[2], k1[2];
                for (int i = 0; i < (two ? 2 : 1); ++i) {
                    const double2 acc = i ? a1 : a0, w1 = mW1[r + i + N * col];
                    k0[i] = make_double2(-0.5 * w1.x - c2h * acc.x, -0.5 * w1.y - c2h * acc.y);
                    k1[i] = make_double2(-0.5 * w1.x + c2h * acc.x, -0.5 * w1.y + c2h * acc.y);
                }
                putH(QO_KH0, r, col, k0[0], k0[1], two);
                putH(QO_KH1, r, col, k1[0], k1[1], two);
            });
            for (int e = tid; e < NN; e += blockDim.x) {  // G[r][q] = sum_c D[r][c] conj(M[q][c])
                const int r = e % N, q = e / N;
                double2 acc = make_double2(0.0, 0.0);
                for (int cc = 0; cc < nc; ++cc) {
                    double2 mq = mM[q + N * cc];
                    mq.y = -mq.y;
                    cfma(acc, mD[r + N * cc], mq);
                }
                T1[r * N + q] = acc;  // row-major: G[r][q] at r * N + q (the anticommutator lists address it that way)
            }
            __syncthreads();
            const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
            for (int pr = warp; pr < npair; pr += nwarps) {
                double val = 0.0;
                for (int u = __ldg(kkptr + pr) + lane, u1 = __ldg(kkptr + pr + 1); u < u1; u += 32) {
                    const int rc = __ldg(kkrc + u);
                    const double2 kv = __ldg(kkv + u);
                    const double2 g = T1[(rc & 255) * N + (rc >> 8)];
                    val = fma(kv.x, g.x, val);
                    val = fma(-kv.y, g.y, val);
                }
                val = warp_sum(val);
                if (lane == 0) {
                    int jj = 0, rem = pr;
                    while (rem > jj) { rem -= jj + 1; ++jj; }
                    put_scalar(qo_haa(rem, jj), c2h2 * val);
                }
            }
        }
    }
}

}  // namespace

size_t qck_big_smem(const QckClassDev& c) {
    const size_t NN = (size_t)c.N * c.N, NS = (size_t)c.N * (c.nc == 1 ? 1 : c.N);
    return 16 * (NN + 6 * NS + 2 * NN) + 64 * 8;
}

int qck_launch_big(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    if (!c.big || !c.qdst) return 0;
    const bool ket = c.kind == QCK_KET_PADE;
    typedef void (*kern_t)(const QckLaunch);
    kern_t kern = ket ? qck_big_kernel<true> : qck_big_kernel<false>;
    const size_t smem = qck_big_smem(c);
    if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
    int per_sm = 1;
    if (L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem) {
        per_sm = L.plan->per_sm;
    } else {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
        if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = per_sm; }
    }
    const long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    long long grid = std::min<long long>(n_items, (long long)sm_count * per_sm);
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] large-level kernel: N=%d nd=%d ket=%d smem=%zu B CTAs/SM=%d grid=%lld\n", c.N, c.nd, (int)ket, smem, per_sm, grid);
    kern<<<(unsigned)grid, 256, smem, stream>>>(L);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}
