// Generic spectral kernel for the exponential integrators (5..16 levels, unitaries and kets, Hermitian Hamiltonians, any number of
// members) of libqcknot.so.  Compiled as its own translation unit.
#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// The classes the specialised spectral kernels do not take (qck_expeig.cu: 9-level unitaries with one member; qck_colexp.cu: 2..4
// levels): 5..8 and 10..16 levels, 9-level kets and ensembles.  Same mathematics (H = V L V^H by Jacobi, divided differences of
// exp in the eigenbasis, see qck_expeig.cu); ONE WARP per (knot, integrator) item, every matrix of the item in the warp's shared
// memory, run-time level count, each output element a dot product walked by one lane.  Values go straight to the arrays through
// the per-member destination tables (qdst), shared-control entries of an ensemble to the partial columns.  No tape, no squaring
// count.  Not as tight as the specialised kernels (no register-resident operands) but 3-6x the scaling-and-squaring kernel on these
// classes, and many more warps per SM.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int ge_tet(int n) { return n * (n + 1) * (n + 2) / 6; }
// double2 units of one warp's shared memory
__host__ __device__ inline size_t ge_warp_units(int N, int nc, int nd) {
    const int NL = N * (N | 1), NS = N * nc;  // square matrices: odd leading dimension (no bank conflicts on column walks)
    return (size_t)2 * NL /* S = [H; V] */ + 4 * NS /* U0|T2, M, W0, Mt */ + (size_t)nd * NL + 3 * NL /* Phi, Gamma, T3 */ + ge_tet(N) + 2 * N /* ex, hx */ +
           (N + 1) / 2 /* lam */ + 2 * ((N + 1) / 2) /* rotations */ + 2;
}

template <int ND, bool KET>
__global__ void __launch_bounds__(256) qck_genexp_kernel(const QckLaunch p, int wunits) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int N = c.N, nc = KET ? 1 : N, NN = N * N, NS = N * nc, n2 = 2 * N, blk = n2 * n2, W = c.W, P = N / 2;
    const int LD = N | 1, NL = N * LD;  // square matrices in shared memory: element (r, k) at r * LD + k, LD odd
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)(blockDim.x >> 5);
    const bool needF = p.mask & QCK_EVAL_F, needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const bool free_time = c.free_time;
    double2* const S = reinterpret_cast<double2*>(smem_all) + (size_t)warp * wunits;  // rows 0..N-1: H, rows N..2N-1: V (row-major)
    double2* const mV = S + NL;
    double2* const vU0 = mV + NL;      // columns of U0: [c * N + r]; later scratch T2
    double2* const vM = vU0 + NS;
    double2* const vW0 = vM + NS;      // V^H U0
    double2* const vMt = vW0 + NS;     // V^H M
    double2* const vB = vMt + NS;      // B_j = V^H A_j V, row-major
    double2* const Phi = vB + ND * NL;
    double2* const vG = Phi + NL;      // Gamma = W0 Mt^H, row-major
    double2* const T3 = vG + NL;       // Lt_j = h B_j o Phi
    double2* const f3 = T3 + NL;
    double2* const ex = f3 + ge_tet(N);
    double2* const hx = ex + N;
    double* const lam = reinterpret_cast<double*>(hx + N);
    double2* const rot = hx + N + (N + 1) / 2;  // per pair: (c, 0), sigma
    double2* const T1 = S;             // after the decomposition: scratch (N x N)
    double2* const T2 = vU0;
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;

    for (long long it = (long long)blockIdx.x * nwarps + warp; it < n_items; it += (long long)gridDim.x * nwarps) {
        const long long t = it / nact;
        const int mi = (int)(it - t * nact), m = p.member_begin + mi;
        const int soff = p.moff_global[3 * mi], coff = p.moff_global[3 * mi + 1], roff = p.moff_global[3 * mi + 2];
        const double* zt = p.Z + t * c.zdim;
        const int* qd = c.qdst + (size_t)m * QO_COUNT;
        const double2* gv = c.cmat + (size_t)m * c.cmat_stride;  // [A0 | ell values | anticommutator values | contributor values]
        const int* gc = c.ell_col + (size_t)m * c.icon_stride;   // [ell columns | kk_ptr | kk_rc | ac_ptr | ac_j]
        const double2* ellv = gv + NN;
        const double2* acv = ellv + c.ell_stride + c.kk_cap;
        const int* ellc = gc;
        const int* acptr = gc + c.ell_stride + ND * (ND + 1) / 2 + 1 + c.kk_cap;
        const int* acj = acptr + NN + 1;
        double* const oF = p.F + t * c.dyn;
        double* const oJ = p.J + t * p.nnzJ;
        double* const oH = p.H + t * p.nnzH;
        double* const oP = p.partial + t * p.npart;
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        auto put = [&](double* arr0, int d0, int q, int r, int col, double2 v) {  // element (r, col) of an iso-vector quantity
            const int st = c.pl_stride[q], i = col * n2 + r;
            arr0[d0 + i * st] = v.x;
            arr0[d0 + (i + N) * st] = v.y;
        };
        auto putJ = [&](int q, int r, int col, double2 v) { const int d0 = qd[q]; if (d0 >= 0) put(oJ, d0, q, r, col, v); };
        auto putH = [&](int q, int r, int col, double2 v) {
            const int d0 = qd[q];
            if (d0 < 0) return;
            if (d0 < p.nnzH) put(oH, d0, q, r, col, v);
            else put(oP, d0 - (int)p.nnzH, q, r, col, v);
        };
        auto put_scalar = [&](int q, double v) {
            const int d0 = qd[q];
            if (d0 < 0 || lane != 0) return;
            if (d0 < p.nnzH) oH[d0] = v;
            else oP[d0 - p.nnzH] = v;
        };
        __syncwarp();  // the previous item's readers are through
        // ---- operands; H = i A(a) (row-major), V = I ---------------------------------------------------------------------------------
        for (int e = lane; e < NS; e += 32) {
            const int r = e % N, col = e / N;
            vU0[e] = make_double2(zt[soff + col * n2 + r], zt[soff + col * n2 + N + r]);
            if (needH) vM[e] = make_double2(p.mu[t * c.dyn + roff + col * n2 + r], p.mu[t * c.dyn + roff + col * n2 + N + r]);
        }
        double fro2 = 0.0;
        for (int e = lane; e < NN; e += 32) {
            double2 v = __ldg(gv + e);
            for (int u = __ldg(acptr + e), u1 = __ldg(acptr + e + 1); u < u1; ++u) {
                const double aj = zt[coff + __ldg(acj + u)];
                const double2 d = __ldg(acv + u);
                v.x = fma(aj, d.x, v.x);
                v.y = fma(aj, d.y, v.y);
            }
            const int r = e % N, col = e / N;  // A0 is column-major
            S[r * LD + col] = make_double2(-v.y, r == col ? 0.0 : v.x);
            mV[r * LD + col] = make_double2(r == col ? 1.0 : 0.0, 0.0);
            fro2 = fma(v.x, v.x, fma(v.y, v.y, fro2));
        }
        fro2 = warp_sum(fro2);
        __syncwarp();
        // ---- cyclic Jacobi, round-robin ordering (odd N: N rounds, level s idle; even N: N-1 rounds, level N-1 paired with s) --------
        {
            const int M = (N & 1) ? N : N - 1;
            for (int sweep = 0; sweep < 14; ++sweep) {
                double off2 = 0.0;
                for (int e = lane; e < NN; e += 32) {
                    const int r = e / N, k = e - r * N;
                    const double2 v = S[r * LD + k];
                    if (r != k) off2 = fma(v.x, v.x, fma(v.y, v.y, off2));
                }
                off2 = warp_sum(off2);
                if (off2 <= 1e-30 * fro2) break;
                for (int s = 0; s < M; ++s) {
                    auto pair_of = [&](int l, int& pp, int& qq) {  // l = 0 .. P-1
                        if (!(N & 1) && l == P - 1) { pp = s; qq = N - 1; return; }
                        pp = s + l + 1; qq = s - l - 1;
                        pp -= pp >= M ? M : 0; qq += qq < 0 ? M : 0;
                    };
                    if (lane < P) {  // rotation of pair `lane` from the pivots: J = [[c, conj(sg)], [-sg, c]], inner rotation
                        int pp, qq;
                        pair_of(lane, pp, qq);
                        const double al = S[pp * LD + pp].x, ga = S[qq * LD + qq].x;
                        const double2 be = S[pp * LD + qq];
                        const double b2 = be.x * be.x + be.y * be.y;
                        double cr = 1.0;
                        double2 sg = make_double2(0.0, 0.0);
                        if (b2 > 0.0) {
                            const double d = 0.5 * (ga - al);
                            const double inv_r = rsqrt(fma(d, d, b2));
                            const double u = fma(0.5 * fabs(d), inv_r, 0.5);
                            const double inv_c = rsqrt(u);
                            cr = u * inv_c;
                            const double f = copysign(0.5 * inv_r * inv_c, d);
                            sg = make_double2(f * be.x, -f * be.y);
                        }
                        rot[2 * lane] = make_double2(cr, 0.0);
                        rot[2 * lane + 1] = sg;
                    }
                    __syncwarp();
                    for (int task = lane; task < 2 * N * P; task += 32) {  // pass R: S <- S J, rows of H and of V alike
                        const int l = task / (2 * N), rho = task - l * 2 * N;
                        int pp, qq;
                        pair_of(l, pp, qq);
                        const double cr = rot[2 * l].x;
                        const double2 sg = rot[2 * l + 1];
                        const double2 a = S[rho * LD + pp], b = S[rho * LD + qq];
                        S[rho * LD + pp] = make_double2(cr * a.x - (sg.x * b.x - sg.y * b.y), cr * a.y - (sg.x * b.y + sg.y * b.x));
                        S[rho * LD + qq] = make_double2(cr * b.x + (sg.x * a.x + sg.y * a.y), cr * b.y + (sg.x * a.y - sg.y * a.x));
                    }
                    __syncwarp();
                    for (int task = lane; task < N * P; task += 32) {  // pass L: H <- J^H H; the pivots take their exact values
                        const int l = task / N, gam = task - l * N;
                        int pp, qq;
                        pair_of(l, pp, qq);
                        const double cr = rot[2 * l].x;
                        const double2 sg = rot[2 * l + 1];
                        const double2 a = S[pp * LD + gam], b = S[qq * LD + gam];
                        double2 za = make_double2(cr * a.x - (sg.x * b.x + sg.y * b.y), cr * a.y - (sg.x * b.y - sg.y * b.x));
                        double2 zb = make_double2(cr * b.x + (sg.x * a.x - sg.y * a.y), cr * b.y + (sg.x * a.y + sg.y * a.x));
                        if (gam == pp) { za.y = 0.0; zb = make_double2(0.0, 0.0); }
                        if (gam == qq) { zb.y = 0.0; za = make_double2(0.0, 0.0); }
                        S[pp * LD + gam] = za;
                        S[qq * LD + gam] = zb;
                    }
                    __syncwarp();
                }
            }
        }
        // ---- spectrum ------------------------------------------------------------------------------------------------------------------
        if (lane < N) {
            const double l = S[lane * LD + lane].x;
            lam[lane] = l;
            double sn, cs;
            sincos(0.5 * h * l, &sn, &cs);
            hx[lane] = make_double2(cs, -sn);
            ex[lane] = make_double2(cs * cs - sn * sn, -2.0 * sn * cs);
        }
        __syncwarp();
        // ---- W0 = V^H U0, Mt = V^H M ----------------------------------------------------------------------------------------------------
        for (int e = lane; e < NS; e += 32) {
            const int k = e % N, col = e / N;
            double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
            for (int r = 0; r < N; ++r) {
                const double2 v = mV[r * LD + k];
                const double2 vc = make_double2(v.x, -v.y);
                cfma(a0, vc, vU0[col * N + r]);
                if (needH) cfma(a1, vc, vM[col * N + r]);
            }
            vW0[e] = a0;
            if (needH) vMt[e] = a1;
        }
        __syncwarp();
        if (needJ || needH) {
            for (int j = 0; j < ND; ++j) {  // B_j = V^H (A_j V)
                for (int e = lane; e < NN; e += 32) {  // T1[r][k] = sum_w A_j[r][.] V[.][k]
                    const int r = e / N, k = e - r * N;
                    const int o = ((j * 2) * N + r) * W;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int w = 0; w < W; ++w) {
                        const double2 v = __ldg(ellv + o + w);
                        if (v.x != 0.0 || v.y != 0.0) cfma(acc, v, mV[__ldg(ellc + o + w) * LD + k]);
                    }
                    T1[r * LD + k] = acc;
                }
                __syncwarp();
                for (int e = lane; e < NN; e += 32) {
                    const int pp = e / N, q = e - pp * N;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int r = 0; r < N; ++r) {
                        const double2 v = mV[r * LD + pp];
                        cfma(acc, make_double2(v.x, -v.y), T1[r * LD + q]);
                    }
                    vB[j * NL + pp * LD + q] = acc;
                }
                __syncwarp();
            }
            for (int e = lane; e < NN; e += 32) {  // Phi_pq = e^{(x_p + x_q)/2} sinc(h (l_p - l_q) / 2)
                const int pp = e / N, q = e - pp * N;
                const double dl = 0.5 * h * (lam[pp] - lam[q]);
                const double sc = dl == 0.0 ? 1.0 : sin(dl) / dl;
                const double2 g = cmul(hx[pp], hx[q]);
                Phi[pp * LD + q] = make_double2(sc * g.x, sc * g.y);
            }
            __syncwarp();
        }
        // ---- residual, d/dh --------------------------------------------------------------------------------------------------------------
        for (int e = lane; e < NS; e += 32) {
            const int r = e % N, col = e / N;
            double2 yE = make_double2(0.0, 0.0), yT = make_double2(0.0, 0.0);
            for (int k = 0; k < N; ++k) {
                const double2 w = cmul(ex[k], vW0[col * N + k]);
                const double2 v = mV[r * LD + k];
                cfma(yE, v, w);
                cfma(yT, v, make_double2(lam[k] * w.y, -lam[k] * w.x));  // (-i l) e^x w0
            }
            if (needF && qd[QO_R] >= 0) {
                const double u1r = zt[c.zdim + soff + col * n2 + r], u1i = zt[c.zdim + soff + col * n2 + N + r];
                put(oF, qd[QO_R], QO_R, r, col, make_double2(u1r - yE.x, u1i - yE.y));
            }
            if (needJ && free_time) putJ(QO_TH, r, col, make_double2(-yT.x, -yT.y));
        }
        if (needJ) {
            // -iso(E), E = V e^x V^H, into every one of the nc copies; identity block on state_{t+1}
            for (int e = lane; e < NN; e += 32) {
                const int r = e / N, cI = e - r * N;
                double2 acc = make_double2(0.0, 0.0);
                for (int k = 0; k < N; ++k) {
                    const double2 v = mV[cI * LD + k];
                    cfma(acc, mV[r * LD + k], cmul(ex[k], make_double2(v.x, -v.y)));
                }
                T1[r * LD + cI] = acc;  // E[r][cI]
            }
            __syncwarp();
            const int dF = qd[QO_ISOF];
            if (dF >= 0)
                for (int e = lane; e < blk; e += 32) {
                    const int col2 = e / n2, row2 = e - col2 * n2;
                    const int r = row2 < N ? row2 : row2 - N, cI = col2 < N ? col2 : col2 - N;
                    const double2 ev = T1[r * LD + cI];
                    const double val = (row2 < N) == (col2 < N) ? -ev.x : (row2 >= N ? -ev.y : ev.y);
                    for (int cb = 0; cb < nc; ++cb) oJ[dF + cb * blk + e] = val;
                }
            const int dO = qd[QO_ONE];
            if (dO >= 0) {
                const int st = c.pl_stride[QO_ONE];
                for (int i = lane; i < 2 * NS; i += 32) oJ[dO + i * st] = 1.0;
            }
            __syncwarp();
        }
        if (needH) {
            for (int e = lane; e < NN; e += 32) {  // Gamma[r][p] = sum_c W0[r][c] conj(Mt[p][c])
                const int r = e / N, pp = e - r * N;
                double2 acc = make_double2(0.0, 0.0);
                for (int col = 0; col < nc; ++col) {
                    const double2 mv = vMt[col * N + pp];
                    cfma(acc, vW0[col * N + r], make_double2(mv.x, -mv.y));
                }
                vG[r * LD + pp] = acc;
            }
            for (int idx = lane; idx < ge_tet(N); idx += 32) {  // second-order divided differences of the sorted triples
                int hi = 0;
                while ((hi + 1) * (hi + 2) * (hi + 3) / 6 <= idx) ++hi;
                const int rem = idx - hi * (hi + 1) * (hi + 2) / 6;
                int mid = 0;
                while ((mid + 1) * (mid + 2) / 2 <= rem) ++mid;
                const int lo = rem - mid * (mid + 1) / 2;
                const double t0 = h * lam[lo], t1 = h * lam[mid], t2 = h * lam[hi];
                const double g01 = fabs(t0 - t1), g02 = fabs(t0 - t2), g12 = fabs(t1 - t2);
                int u, v, w;
                double gap;
                if (g01 >= g02 && g01 >= g12) { u = lo; w = mid; v = hi; gap = t0 - t1; }
                else if (g02 >= g12) { u = lo; w = hi; v = mid; gap = t0 - t2; }
                else { u = mid; w = hi; v = lo; gap = t1 - t2; }
                double2 f;
                if (fabs(gap) >= 0.4) {
                    const double2 a = Phi[u * LD + v], b = Phi[v * LD + w];
                    const double inv = 1.0 / gap;
                    f = make_double2(-(a.y - b.y) * inv, (a.x - b.x) * inv);
                } else {
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
                    f = cmul(ex[lo], make_double2(re, im));
                }
                f3[idx] = f;
            }
            if (free_time)
                for (int e = lane; e < NS; e += 32) {  // state x dt = -V (conj(-i l e^x) . Mt)
                    const int r = e % N, col = e / N;
                    double2 y = make_double2(0.0, 0.0);
                    for (int k = 0; k < N; ++k) cfma(y, mV[r * LD + k], cmul(make_double2(lam[k] * ex[k].y, lam[k] * ex[k].x), vMt[col * N + k]));
                    putH(QO_KH0, r, col, make_double2(-y.x, -y.y));
                }
            __syncwarp();
        }
        // ---- drive terms -----------------------------------------------------------------------------------------------------------------
        if (needJ || needH) {
            for (int j = 0; j < ND; ++j) {
                const double2* Bj = vB + j * NL;
                double s_ah = 0.0;
                for (int e = lane; e < NN; e += 32) {  // Lt_j = h B_j o Phi; a_j x dt takes (B_j,pq e^{x_q} - i l_p Lt_pq) Gamma_qp from here
                    const int pp = e / N, q = e - pp * N;
                    const int ei = pp * LD + q;
                    const double2 b = Bj[ei];
                    const double2 l0 = cmul(b, Phi[ei]);
                    const double2 l = make_double2(h * l0.x, h * l0.y);
                    T3[ei] = l;
                    if (needH && free_time) {
                        const double2 be = cmul(b, ex[q]);
                        const double2 tot = make_double2(be.x + lam[pp] * l.y, be.y - lam[pp] * l.x);
                        const double2 g = vG[q * LD + pp];
                        s_ah -= tot.x * g.x - tot.y * g.y;
                    }
                }
                __syncwarp();
                if (needJ) {  // d/da_j = -V (Lt_j W0)
                    for (int e = lane; e < NS; e += 32) {
                        const int pp = e % N, col = e / N;
                        double2 acc = make_double2(0.0, 0.0);
                        for (int q = 0; q < N; ++q) cfma(acc, T3[pp * LD + q], vW0[col * N + q]);
                        T2[e] = acc;
                    }
                    __syncwarp();
                    for (int e = lane; e < NS; e += 32) {
                        const int r = e % N, col = e / N;
                        double2 y = make_double2(0.0, 0.0);
                        for (int k = 0; k < N; ++k) cfma(y, mV[r * LD + k], T2[col * N + k]);
                        putJ(QO_TA + j, r, col, make_double2(-y.x, -y.y));
                    }
                    __syncwarp();
                }
                if (needH) {  // state x a_j = -V (Lt_j^H Mt)
                    for (int e = lane; e < NS; e += 32) {
                        const int q = e % N, col = e / N;
                        double2 acc = make_double2(0.0, 0.0);
                        for (int pp = 0; pp < N; ++pp) {
                            const double2 l = T3[pp * LD + q];
                            cfma(acc, make_double2(l.x, -l.y), vMt[col * N + pp]);
                        }
                        T2[e] = acc;
                    }
                    __syncwarp();
                    for (int e = lane; e < NS; e += 32) {
                        const int r = e % N, col = e / N;
                        double2 y = make_double2(0.0, 0.0);
                        for (int k = 0; k < N; ++k) cfma(y, mV[r * LD + k], T2[col * N + k]);
                        putH(QO_KA0 + j, r, col, make_double2(-y.x, -y.y));
                    }
                    if (free_time) put_scalar(QO_HAH + j, warp_sum(s_ah));
                    __syncwarp();
                }
            }
        }
        if (needH) {
            // a_i x a_j = -h^2 Re sum_qr (B_j,qr K^i_qr + B_i,qr K^j_qr),  K^i_qr = sum_p Gamma_rp exp[x_p,x_q,x_r] B_i,pq
            double s_aa[ND][ND];
#pragma unroll
            for (int i = 0; i < ND; ++i)
#pragma unroll
                for (int j = 0; j < ND; ++j) s_aa[i][j] = 0.0;
            for (int e = lane; e < NN; e += 32) {
                const int q = e / N, r = e - q * N;
                double2 K[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) K[i] = make_double2(0.0, 0.0);
                for (int pp = 0; pp < N; ++pp) {
                    const int lo = min(pp, min(q, r)), hi = max(pp, max(q, r)), mid = pp + q + r - lo - hi;
                    const double2 gf = cmul(vG[r * LD + pp], f3[hi * (hi + 1) * (hi + 2) / 6 + mid * (mid + 1) / 2 + lo]);
#pragma unroll
                    for (int i = 0; i < ND; ++i) cfma(K[i], gf, vB[i * NL + pp * LD + q]);
                }
#pragma unroll
                for (int j = 0; j < ND; ++j) {
                    const double2 bj = vB[j * NL + q * LD + r];
#pragma unroll
                    for (int i = 0; i < ND; ++i) s_aa[i][j] += bj.x * K[i].x - bj.y * K[i].y;
                }
            }
#pragma unroll
            for (int j = 0; j < ND; ++j)
#pragma unroll
                for (int i = 0; i <= j; ++i) put_scalar(qo_haa(i, j), -h * h * warp_sum(s_aa[i][j] + s_aa[j][i]));
            if (free_time) {  // dt x dt = Re sum_p l_p^2 e^{x_p} Gamma_pp
                double v = 0.0;
                if (lane < N) {
                    const double l = lam[lane];
                    const double2 e = ex[lane], g = vG[lane * LD + lane];
                    v = l * l * (e.x * g.x - e.y * g.y);
                }
                put_scalar(QO_HHH, warp_sum(v));
            }
        }
        if (mi == 0 && p.n_aux) do_aux(p, t, lane, 32);  // derivative-integrator entries of this knot
    }
}

typedef void (*ge_kern_t)(const QckLaunch, int);
template <bool KET>
ge_kern_t ge_pick(int nd) {
    return nd == 1 ? qck_genexp_kernel<1, KET> : (nd == 2 ? qck_genexp_kernel<2, KET> : (nd == 3 ? qck_genexp_kernel<3, KET> : qck_genexp_kernel<4, KET>));
}

}  // namespace

size_t qck_genexp_warp_bytes(int N, int nc, int nd) { return ((ge_warp_units(N, nc, nd) + 1) & ~(size_t)1) * 16; }

int qck_launch_genexp(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    if (!c.genexp) return 0;
    const bool ket = c.kind == QCK_KET_EXP;
    if ((c.kind != QCK_UNITARY_EXP && !ket) || c.N < 2 || c.N > 16 || c.nd < 1 || c.nd > 4 || !c.qdst) return (int)cudaErrorInvalidConfiguration;
    const ge_kern_t kern = ket ? ge_pick<true>(c.nd) : ge_pick<false>(c.nd);
    const size_t wbytes = qck_genexp_warp_bytes(c.N, ket ? 1 : c.N, c.nd);
    int nwarps = 8;
    while (nwarps > 1 && nwarps * wbytes > 110 * 1024) --nwarps;  // two CTAs per SM where they fit
    if (nwarps * wbytes > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
    const size_t smem = nwarps * wbytes;
    int per_sm = 0;
    if (L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem) {
        per_sm = L.plan->per_sm;
    } else {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = per_sm; }
    }
    if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
    const long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    long long grid = (long long)sm_count * per_sm;
    const long long need = (n_items + nwarps - 1) / nwarps;
    if (grid > need) grid = need;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] generic spectral kernel: N=%d nd=%d ket=%d warps/CTA=%d smem=%zu CTAs/SM=%d grid=%lld items=%lld\n", c.N, c.nd, (int)ket, nwarps, smem, per_sm, grid, n_items);
    kern<<<(unsigned)grid, nwarps * 32, smem, stream>>>(L, (int)(wbytes / 16));
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}
