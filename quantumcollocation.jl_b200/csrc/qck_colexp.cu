// Spectral column kernels for the exponential integrators of small systems (2..4 levels, unitaries and kets, Hermitian
// Hamiltonians, any number of members) of libqcknot.so.  Compiled as its own translation unit.
#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// UnitaryExponentialIntegrator / QuantumStateExponentialIntegrator at 2..4 levels (Hadamard, sampling and quantum-state problems in
// their exponential form).  Same mathematics as the 9-level spectral kernel (qck_expeig.cu: H = V L V^H, divided differences of
// exp in the eigenbasis), same thread mapping as the Pade column kernel (qck_column.cu): ONE LANE per column of the state, N lanes
// per (knot, integrator) item (one for a ket), 32 / N items per warp, values stored straight from registers.
//   kernel 1 (qck_coleig_kernel):  one LANE per item diagonalises H(a_t) by cyclic Jacobi, everything in registers (static
//             indices, sweeps until the off-diagonal norm is at rounding level), V and the eigenvalues go to the class scratch;
//   kernel 2 (qck_colexp_kernel):  a lane holds V and its own columns of W0 = V^H U0 and Mt = V^H M in registers; the per-item
//             matrices every lane needs in full -- B_j = V^H A_j V, Phi (first-order divided differences) and the second-order
//             divided differences -- are built column by column by the item's lanes in shared memory.
// With Gamma = W0 Mt^H = sum over columns of w0 mt^H, every scalar second derivative is a sum of per-column terms: no Gamma matrix.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int ce_tet(int n) { return n * (n + 1) * (n + 2) / 6; }
__host__ __device__ constexpr int ce_rec(int n) { return n * n + (n + 1) / 2; }          // double2 per item in the scratch: V | eigenvalues
// double2 per item in shared memory; odd, so that the items of a warp start on different banks (256 B per item put all 16 two-level
// items of a warp on the same banks: 81 % of the shared-memory wavefronts were conflicts)
__host__ __device__ constexpr int ce_item(int n, int nd) { return (nd * n * n + n * n + ce_tet(n)) | 1; }

__device__ __forceinline__ void ce_rot(double al, double ga, double2 be, double& c, double2& sg) {
    const double b2 = be.x * be.x + be.y * be.y;
    c = 1.0;
    sg = make_double2(0.0, 0.0);
    if (b2 > 0.0) {  // inner rotation J = [[c, conj(sg)], [-sg, c]] that diagonalises [[al, be], [conj(be), ga]]
        const double d = 0.5 * (ga - al);
        const double inv_r = rsqrt(fma(d, d, b2));
        const double u = fma(0.5 * fabs(d), inv_r, 0.5);
        const double inv_c = rsqrt(u);
        c = u * inv_c;
        const double f = copysign(0.5 * inv_r * inv_c, d);
        sg = make_double2(f * be.x, -f * be.y);
    }
}
__device__ __forceinline__ void ce_right(double2& y1, double2& y2, double c, double2 sg) {  // (y1, y2) <- (y1, y2) J
    const double2 a = y1, b = y2;
    y1 = make_double2(c * a.x - (sg.x * b.x - sg.y * b.y), c * a.y - (sg.x * b.y + sg.y * b.x));
    y2 = make_double2(c * b.x + (sg.x * a.x + sg.y * a.y), c * b.y + (sg.x * a.y - sg.y * a.x));
}
__device__ __forceinline__ void ce_left(double2& z1, double2& z2, double c, double2 sg) {  // (z1; z2) <- J^H (z1; z2)
    const double2 a = z1, b = z2;
    z1 = make_double2(c * a.x - (sg.x * b.x + sg.y * b.y), c * a.y - (sg.x * b.y - sg.y * b.x));
    z2 = make_double2(c * b.x + (sg.x * a.x - sg.y * a.y), c * b.y + (sg.x * a.y + sg.y * a.x));
}

// ---- kernel 1: eigen-decomposition, one lane per (knot, member) item ------------------------------------------------------------
template <int N, int ND>
__global__ void __launch_bounds__(128) qck_coleig_kernel(const QckLaunch p) {
    const QckClassDev& c = p.c;
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (long long)gridDim.x * blockDim.x) {
        const long long t = it / nact;
        const int mi = (int)(it - t * nact), m = p.member_begin + mi;
        const double* zt = p.Z + t * c.zdim;
        const int coff = p.moff_global[3 * mi + 1];
        double a[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) a[j] = zt[coff + j];
        const double2* const A0g = c.cmat + (size_t)m * c.cmat_stride;  // column-major
        const double2* const Ajg = c.dense_aj + (size_t)m * ND * N * N;  // [drive][row][column]
        double2 H[N][N], V[N][N];
        double fro2 = 0.0;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double2 v = __ldg(A0g + r + N * k);
#pragma unroll
                for (int j = 0; j < ND; ++j) {
                    const double2 w = __ldg(Ajg + (j * N + r) * N + k);
                    v.x = fma(a[j], w.x, v.x);
                    v.y = fma(a[j], w.y, v.y);
                }
                H[r][k] = make_double2(-v.y, r == k ? 0.0 : v.x);  // H = i A
                V[r][k] = make_double2(r == k ? 1.0 : 0.0, 0.0);
                fro2 = fma(v.x, v.x, fma(v.y, v.y, fro2));
            }
        for (int sweep = 0; sweep < 14; ++sweep) {
            double off2 = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int k = r + 1; k < N; ++k) off2 += 2.0 * (H[r][k].x * H[r][k].x + H[r][k].y * H[r][k].y);
            if (off2 <= 1e-30 * fro2) break;
#pragma unroll
            for (int pp = 0; pp < N - 1; ++pp)
#pragma unroll
                for (int qq = pp + 1; qq < N; ++qq) {
                    double cr;
                    double2 sg;
                    ce_rot(H[pp][pp].x, H[qq][qq].x, H[pp][qq], cr, sg);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        ce_right(H[k][pp], H[k][qq], cr, sg);
                        ce_right(V[k][pp], V[k][qq], cr, sg);
                    }
#pragma unroll
                    for (int k = 0; k < N; ++k) ce_left(H[pp][k], H[qq][k], cr, sg);
                    H[pp][qq] = H[qq][pp] = make_double2(0.0, 0.0);
                    H[pp][pp].y = 0.0;
                    H[qq][qq].y = 0.0;
                }
        }
        double2* out = c.tape + (size_t)it * ce_rec(N);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int k = 0; k < N; ++k) out[r * N + k] = V[r][k];
#pragma unroll
        for (int k = 0; k < N; ++k) reinterpret_cast<double*>(out + N * N)[k] = H[k][k].x;
    }
}

// ---- kernel 2: values -------------------------------------------------------------------------------------------------------------
// NC: columns of the state (N for unitaries, 1 for kets)
template <int N, int ND, int NC>
__global__ void __launch_bounds__(256) qck_colexp_kernel(const QckLaunch p) {
    constexpr int n2 = 2 * N, blk = n2 * n2, IPW = 32 / NC, NN = N * N, TET = ce_tet(N), ITEM = ce_item(N, ND);
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / NC, col = lane - gi * NC;  // item slot inside the warp, column
    const bool needF = p.mask & QCK_EVAL_F, needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    const long long nslots = (long long)gridDim.x * (blockDim.x >> 5) * IPW;
    const bool free_time = c.free_time;
    double2* const Bs = reinterpret_cast<double2*>(smem_all) + ((size_t)warp * IPW + (gi < IPW ? gi : 0)) * ITEM;  // [drive][p][q]
    double2* const Phi = Bs + ND * NN;                                                                               // [p][q]
    double2* const f3 = Phi + NN;                                                                                    // sorted triples

    auto store_run = [](double* dst, const double (&v)[n2]) {  // 2N consecutive doubles, 16-byte stores where aligned
        if (reinterpret_cast<uintptr_t>(dst) & 8) {
            dst[0] = v[0];
#pragma unroll
            for (int i = 0; i < N - 1; ++i) *reinterpret_cast<double2*>(dst + 1 + 2 * i) = make_double2(v[1 + 2 * i], v[2 + 2 * i]);
            dst[n2 - 1] = v[n2 - 1];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) *reinterpret_cast<double2*>(dst + 2 * i) = make_double2(v[2 * i], v[2 * i + 1]);
        }
    };
    auto gsum = [&](double v) {  // sum over the NC lanes of this lane's item
        double r = v;
#pragma unroll
        for (int o = 1; o < NC; ++o) r += __shfl_sync(0xffffffffu, v, (gi * NC + (col + o) % NC) & 31);
        return r;
    };
    auto tri = [](int a, int b, int d) {
        const int lo = min(a, min(b, d)), hi = max(a, max(b, d)), mid = a + b + d - lo - hi;
        return hi * (hi + 1) * (hi + 2) / 6 + mid * (mid + 1) / 2 + lo;
    };

    for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * IPW; base < n_items; base += nslots) {
        const long long item = base + gi;
        const bool on = gi < IPW && item < n_items;
        const long long it = on ? item : base;  // idle lanes shadow a valid item (no stores)
        const long long t = it / nact;
        const int mi = (int)(it - t * nact), m = p.member_begin + mi;
        const int soff = p.moff_global[3 * mi], roff = p.moff_global[3 * mi + 2];
        const double* zt = p.Z + t * c.zdim;
        const int* qd = c.qdst + (size_t)m * QO_COUNT;
        double* const oF = p.F + t * c.dyn;
        double* const oJ = p.J + t * p.nnzJ;
        auto put_vec = [&](double* arr0, int d0, int q, const double2 (&x)[N]) {
            const int st = c.pl_stride[q];
            if (st == 1) {
                double v[n2];
#pragma unroll
                for (int r = 0; r < N; ++r) { v[r] = x[r].x; v[N + r] = x[r].y; }
                store_run(arr0 + d0 + col * n2, v);
            } else {
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    arr0[d0 + (col * n2 + r) * st] = x[r].x;
                    arr0[d0 + (col * n2 + N + r) * st] = x[r].y;
                }
            }
        };
        auto put_J = [&](int q, const double2 (&x)[N]) {
            const int d0 = qd[q];
            if (d0 >= 0 && on) put_vec(oJ, d0, q, x);
        };
        auto put_H = [&](int q, const double2 (&x)[N]) {  // (>= nnzH: partial column of a shared position)
            const int d0 = qd[q];
            if (d0 < 0 || !on) return;
            if (d0 < p.nnzH) put_vec(p.H + t * p.nnzH, d0, q, x);
            else put_vec(p.partial + t * p.npart, d0 - (int)p.nnzH, q, x);
        };
        auto put_scalar = [&](int q, double v) {
            const int d0 = qd[q];
            if (d0 < 0 || !on || col != 0) return;
            if (d0 < p.nnzH) p.H[t * p.nnzH + d0] = v;
            else p.partial[t * p.npart + (d0 - p.nnzH)] = v;
        };
        // ---- inputs: eigenvectors / eigenvalues of the item, this lane's columns of U0, U1 and of the multipliers ------------------
        const double2* const scr = c.tape + (size_t)it * ce_rec(N);
        double2 V[N][N];
        double lam[N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int k = 0; k < N; ++k) V[r][k] = __ldg(scr + r * N + k);
#pragma unroll
        for (int k = 0; k < N; ++k) lam[k] = __ldg(reinterpret_cast<const double*>(scr + NN) + k);
        double2 u0[N], u1[N], mm[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            u0[r] = make_double2(zt[soff + col * n2 + r], zt[soff + col * n2 + N + r]);
            u1[r] = make_double2(zt[c.zdim + soff + col * n2 + r], zt[c.zdim + soff + col * n2 + N + r]);
            mm[r] = needH ? make_double2(p.mu[t * c.dyn + roff + col * n2 + r], p.mu[t * c.dyn + roff + col * n2 + N + r]) : make_double2(0.0, 0.0);
        }
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        const double2* const Ajg = c.dense_aj + (size_t)m * ND * N * N;  // [drive][row][column]
        double2 ex[N], hx[N];  // e^{x_p}, e^{x_p / 2},  x_p = -i h l_p
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double sn, cs;
            sincos(0.5 * h * lam[k], &sn, &cs);
            hx[k] = make_double2(cs, -sn);
            ex[k] = make_double2(cs * cs - sn * sn, -2.0 * sn * cs);
        }
        auto mvV = [&](double2 (&y)[N], const double2 (&x)[N]) {  // y = V x
#pragma unroll
            for (int r = 0; r < N; ++r) {
                y[r] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) cfma(y[r], V[r][k], x[k]);
            }
        };
        auto mvVH = [&](double2 (&y)[N], const double2 (&x)[N]) {  // y = V^H x
#pragma unroll
            for (int k = 0; k < N; ++k) {
                y[k] = make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < N; ++r) cfma(y[k], make_double2(V[r][k].x, -V[r][k].y), x[r]);
            }
        };
        double2 w0[N], mt[N];
        mvVH(w0, u0);
        mvVH(mt, mm);
        // ---- per-item matrices in shared memory: this lane's column(s) of B_j = V^H A_j V and of Phi, its share of exp[x_p,x_q,x_r]
        __syncwarp();  // (the previous item's readers are done)
        if (needJ || needH) {
#pragma unroll
            for (int cb0 = 0; cb0 < (NC == 1 ? N : 1); ++cb0) {
                const int cb = NC == 1 ? cb0 : col;
#pragma unroll 1
                for (int j = 0; j < ND; ++j) {
                    double2 T[N], Bc[N];
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        T[r] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            double2 vk = V[k][0];
#pragma unroll
                            for (int q = 1; q < N; ++q) vk = cb == q ? V[k][q] : vk;  // V[k][cb] without dynamic register indexing
                            cfma(T[r], __ldg(Ajg + (j * N + r) * N + k), vk);
                        }
                    }
                    mvVH(Bc, T);
#pragma unroll
                    for (int pp = 0; pp < N; ++pp) Bs[j * NN + pp * N + cb] = Bc[pp];
                }
                double lc = lam[0];
                double2 hc = hx[0];
#pragma unroll
                for (int q = 1; q < N; ++q) { lc = cb == q ? lam[q] : lc; hc = cb == q ? hx[q] : hc; }
#pragma unroll
                for (int pp = 0; pp < N; ++pp) {  // Phi_pq = e^{(x_p + x_q)/2} sinc(h (l_p - l_q) / 2)
                    const double dl = 0.5 * h * (lam[pp] - lc);
                    const double sc = dl == 0.0 ? 1.0 : sin(dl) / dl;
                    const double2 g = cmul(hx[pp], hc);
                    Phi[pp * N + cb] = make_double2(sc * g.x, sc * g.y);
                }
            }
        }
        __syncwarp();
        if (needH) {
            for (int idx = col; idx < TET; idx += NC) {  // second-order divided differences of the sorted triples (lo <= mid <= hi)
                int hi = 0;
                while ((hi + 1) * (hi + 2) * (hi + 3) / 6 <= idx) ++hi;
                const int rem = idx - hi * (hi + 1) * (hi + 2) / 6;
                int mid = 0;
                while ((mid + 1) * (mid + 2) / 2 <= rem) ++mid;
                const int lo = rem - mid * (mid + 1) / 2;
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                double2 e0 = ex[0];
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    t0 = lo == q ? h * lam[q] : t0;
                    t1 = mid == q ? h * lam[q] : t1;
                    t2 = hi == q ? h * lam[q] : t2;
                    e0 = lo == q ? ex[q] : e0;
                }
                const double g01 = fabs(t0 - t1), g02 = fabs(t0 - t2), g12 = fabs(t1 - t2);
                int u, v, w;
                double gap;
                if (g01 >= g02 && g01 >= g12) { u = lo; w = mid; v = hi; gap = t0 - t1; }
                else if (g02 >= g12) { u = lo; w = hi; v = mid; gap = t0 - t2; }
                else { u = mid; w = hi; v = lo; gap = t1 - t2; }
                double2 f;
                if (fabs(gap) >= 0.4) {  // (exp[x_u, x_v] - exp[x_v, x_w]) / (x_u - x_w),  x_u - x_w = -i gap
                    const double2 a = Phi[u * N + v], b = Phi[v * N + w];
                    const double inv = 1.0 / gap;
                    f = make_double2(-(a.y - b.y) * inv, (a.x - b.x) * inv);
                } else {  // series about x_lo: e^{x_lo} sum_k (-i)^k h_k(0, d1, d2) / (k + 2)!
                    const double d1 = t1 - t0, d2 = t2 - t0;
                    double q2 = 1.0, r3 = 1.0, re = 0.5, im = 0.0, inv_fact = 0.5;
#pragma unroll
                    for (int k = 1; k <= 15; ++k) {
                        q2 *= d1;
                        r3 = fma(r3, d2, q2);
                        inv_fact /= (double)(k + 2);
                        const double term = r3 * inv_fact;
                        if ((k & 3) == 1) im -= term;
                        else if ((k & 3) == 2) re -= term;
                        else if ((k & 3) == 3) im += term;
                        else re += term;
                    }
                    f = cmul(e0, make_double2(re, im));
                }
                f3[idx] = f;
            }
        }
        __syncwarp();
        // ---- residual, d/dh, -iso(E), identity block -----------------------------------------------------------------------------------
        {
            double2 x[N], y[N], o[N];
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = cmul(ex[k], w0[k]);
            mvV(y, x);
            if (needF && on && qd[QO_R] >= 0) {
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(u1[r].x - y[r].x, u1[r].y - y[r].y);
                put_vec(oF, qd[QO_R], QO_R, o);
            }
            if (needJ && free_time) {
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = make_double2(lam[k] * x[k].y, -lam[k] * x[k].x);  // (-i l) e^x w0
                mvV(y, x);
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-y[r].x, -y[r].y);
                put_J(QO_TH, o);
            }
        }
        if (needJ && on && qd[QO_ISOF] >= 0) {
            const int dF = qd[QO_ISOF];
            auto block_columns = [&](int bc, int copy0, int copy1) {  // columns bc, bc + N of -iso(E), E = V e^x V^H
                double2 x[N], e[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    double2 vb = V[0][k];
#pragma unroll
                    for (int q = 1; q < N; ++q) vb = bc == q ? V[q][k] : vb;  // V[bc][k]
                    x[k] = cmul(ex[k], make_double2(vb.x, -vb.y));
                }
                mvV(e, x);
                double f0[n2], f1[n2];
#pragma unroll
                for (int r = 0; r < N; ++r) { f0[r] = -e[r].x; f0[N + r] = -e[r].y; f1[r] = e[r].y; f1[N + r] = -e[r].x; }
                for (int cb = copy0; cb < copy1; ++cb) {
                    store_run(oJ + dF + cb * blk + bc * n2, f0);
                    store_run(oJ + dF + cb * blk + (bc + N) * n2, f1);
                }
            };
            if constexpr (NC == N) block_columns(col, 0, N);
            else {
#pragma unroll
                for (int bc = 0; bc < N; ++bc) block_columns(bc, 0, 1);
            }
        }
        if (needJ) {
            double2 one[N];
#pragma unroll
            for (int r = 0; r < N; ++r) one[r] = make_double2(1.0, 1.0);
            put_J(QO_ONE, one);  // identity block on state_{t+1}: this lane's 2N diagonal entries
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {  // d/da_j = -V ((h B_j o Phi) w0)
                double2 y[N], o[N];
#pragma unroll
                for (int pp = 0; pp < N; ++pp) {
                    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int q = 0; q < N; ++q) cfma(acc, cmul(Bs[j * NN + pp * N + q], Phi[pp * N + q]), w0[q]);
                    y[pp] = make_double2(h * acc.x, h * acc.y);
                }
                mvV(o, y);
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-o[r].x, -o[r].y);
                put_J(QO_TA + j, o);
            }
        }
        if (needH) {
            if (free_time) {  // state x dt = -V (conj(-i l e^x) . mt)
                double2 x[N], y[N];
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = cmul(make_double2(lam[k] * ex[k].y, lam[k] * ex[k].x), mt[k]);
                mvV(y, x);
#pragma unroll
                for (int r = 0; r < N; ++r) y[r] = make_double2(-y[r].x, -y[r].y);
                put_H(QO_KH0, y);
                double s_hh = 0.0;  // dt x dt = Re sum_p l_p^2 e^{x_p} Gamma_pp,  Gamma_pp = sum over columns of w0_p conj(mt_p)
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double2 g = cmul(w0[k], make_double2(mt[k].x, -mt[k].y));
                    s_hh += lam[k] * lam[k] * (ex[k].x * g.x - ex[k].y * g.y);
                }
                put_scalar(QO_HHH, gsum(s_hh));
            }
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {  // state x a_j = -V ((h B_j o Phi)^H mt);  a_j x dt
                double2 y[N], o[N];
                double s_ah = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) y[q] = make_double2(0.0, 0.0);
#pragma unroll
                for (int pp = 0; pp < N; ++pp) {
                    const double2 mc = make_double2(mt[pp].x, -mt[pp].y);
#pragma unroll
                    for (int q = 0; q < N; ++q) {
                        const double2 b = Bs[j * NN + pp * N + q], ph = Phi[pp * N + q];
                        const double2 l0 = cmul(b, ph);
                        const double2 l = make_double2(h * l0.x, h * l0.y);
                        cfma(y[q], make_double2(l.x, -l.y), mt[pp]);
                        // a_j x dt:  -(B_j,pq e^{x_q} - i l_p Lt_pq) Gamma_qp,  Gamma_qp = w0_q conj(mt_p) (this column)
                        const double2 be = cmul(b, ex[q]);
                        const double2 tot = make_double2(be.x + lam[pp] * l.y, be.y - lam[pp] * l.x);
                        const double2 g = cmul(w0[q], mc);
                        s_ah -= tot.x * g.x - tot.y * g.y;
                    }
                }
                mvV(o, y);
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-o[r].x, -o[r].y);
                put_H(QO_KA0 + j, o);
                if (free_time) put_scalar(QO_HAH + j, gsum(s_ah));
            }
            // a_i x a_j = -h^2 Re sum_qr (B_j,qr K^i_qr + B_i,qr K^j_qr),  K^i_qr = w0_r sum_p conj(mt_p) exp[x_p,x_q,x_r] B_i,pq  (this column)
            double s_aa[ND][ND];
#pragma unroll
            for (int i = 0; i < ND; ++i)
#pragma unroll
                for (int j = 0; j < ND; ++j) s_aa[i][j] = 0.0;
#pragma unroll 1
            for (int q = 0; q < N; ++q)
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double2 K[ND];
#pragma unroll
                    for (int i = 0; i < ND; ++i) K[i] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int pp = 0; pp < N; ++pp) {
                        const double2 gf = cmul(make_double2(mt[pp].x, -mt[pp].y), f3[tri(pp, q, r)]);
#pragma unroll
                        for (int i = 0; i < ND; ++i) cfma(K[i], gf, Bs[i * NN + pp * N + q]);
                    }
#pragma unroll
                    for (int i = 0; i < ND; ++i) K[i] = cmul(K[i], w0[r]);
#pragma unroll
                    for (int j = 0; j < ND; ++j) {
                        const double2 bj = Bs[j * NN + q * N + r];
#pragma unroll
                        for (int i = 0; i < ND; ++i) s_aa[i][j] += bj.x * K[i].x - bj.y * K[i].y;  // Re(B_j,qr K^i_qr)
                    }
                }
#pragma unroll
            for (int j = 0; j < ND; ++j)
#pragma unroll
                for (int i = 0; i <= j; ++i) put_scalar(qo_haa(i, j), -h * h * gsum(s_aa[i][j] + s_aa[j][i]));
        }
        if (mi == 0 && p.n_aux && on) do_aux(p, t, col, NC);  // derivative-integrator entries of this knot
    }
}

typedef void (*ce_kern_t)(const QckLaunch);
template <int N>
ce_kern_t ce_pick_eig(int nd) {
    return nd == 1 ? qck_coleig_kernel<N, 1> : (nd == 2 ? qck_coleig_kernel<N, 2> : (nd == 3 ? qck_coleig_kernel<N, 3> : qck_coleig_kernel<N, 4>));
}
template <int N, int NC>
ce_kern_t ce_pick_main(int nd) {
    return nd == 1 ? qck_colexp_kernel<N, 1, NC> : (nd == 2 ? qck_colexp_kernel<N, 2, NC> : (nd == 3 ? qck_colexp_kernel<N, 3, NC> : qck_colexp_kernel<N, 4, NC>));
}

}  // namespace

size_t qck_colexp_scratch_rec(int N) { return (size_t)ce_rec(N); }

// small exponential classes (2..4 levels, Hermitian Hamiltonians): eigen kernel + spectral column kernel
int qck_launch_colexp(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    if (!c.colexp) return 0;
    const bool ket = c.kind == QCK_KET_EXP;
    if ((c.kind != QCK_UNITARY_EXP && !ket) || c.N < 2 || c.N > 4 || c.nd < 1 || c.nd > 4 || !c.dense_aj || !c.qdst || !c.tape) return (int)cudaErrorInvalidConfiguration;
    const long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    if (n_items > (long long)c.max_ctas) return (int)cudaErrorInvalidConfiguration;  // eigen scratch: one record per item
    ce_kern_t eig = nullptr, kern = nullptr;
    if (c.N == 2) { eig = ce_pick_eig<2>(c.nd); kern = ket ? ce_pick_main<2, 1>(c.nd) : ce_pick_main<2, 2>(c.nd); }
    else if (c.N == 3) { eig = ce_pick_eig<3>(c.nd); kern = ket ? ce_pick_main<3, 1>(c.nd) : ce_pick_main<3, 3>(c.nd); }
    else { eig = ce_pick_eig<4>(c.nd); kern = ket ? ce_pick_main<4, 1>(c.nd) : ce_pick_main<4, 4>(c.nd); }
    const int ipw = 32 / (ket ? 1 : c.N);
    const size_t item_bytes = (size_t)ce_item(c.N, c.nd) * 16;
    int threads = 256;
    while (threads > 32 && (size_t)(threads / 32) * ipw * item_bytes > 96 * 1024) threads /= 2;
    const size_t smem = (size_t)(threads / 32) * ipw * item_bytes;
    int per_sm = 0;
    if (L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem) {
        per_sm = L.plan->per_sm;
    } else {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return (int)e;
        if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = per_sm; }
    }
    if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
    {
        long long egrid = (n_items + 127) / 128;
        if (egrid > (long long)sm_count * 16) egrid = (long long)sm_count * 16;
        eig<<<(unsigned)egrid, 128, 0, stream>>>(L);
        if (launches) ++*launches;
    }
    const int wpb = threads / 32;
    long long grid = (long long)sm_count * per_sm;
    const long long need = (n_items + (long long)wpb * ipw - 1) / ((long long)wpb * ipw);
    if (grid > need) grid = need;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] spectral column kernels: N=%d nd=%d ket=%d threads=%d smem=%zu CTAs/SM=%d grid=%lld items=%lld\n", c.N, c.nd, (int)ket, threads, smem, per_sm, grid, n_items);
    kern<<<(unsigned)grid, threads, smem, stream>>>(L);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}
