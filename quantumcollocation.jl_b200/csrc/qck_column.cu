// Column kernel (2..4-level Pade-4 unitaries and kets, one lane per column) of libqcknot.so.  Compiled as its own translation unit so that the kernel families build in parallel.
#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// Column kernel: Pade-4, unitaries and kets, 2..4 levels (Hadamard / sampling / quantum-state problems).  ONE LANE per column:
// N lanes per (knot, integrator) work item, 32 / N items per warp, no shared memory, no barriers.
//
// A lane holds ALL of A = -i H(a) (N x N complex) and its own columns d, s, m of D = U1 - U0, S = U1 + U0, M in registers;
// every product of the path is a local matrix-vector product (same matrix-vector form as the row-slice kernel), the scalar
// second derivatives are dot products summed over the item's N lanes with shuffles.  The constant drives A_j are read
// (dense, per member) through L1.  A lane's values of one output quantity are 2N consecutive doubles of the value arrays
// (its column of an iso-vector; its two columns of every copy of a kron(I_N, .) block), so they leave as 16-byte stores
// straight from registers: no staging image.  Destinations per member come from the host's placement pass.
// ------------------------------------------------------------------------------------------------------------
// NC: columns of the state (N for unitaries, 1 for kets: QuantumStatePadeIntegrator = the same algebra on one column)
//
// ST (block-staged write-out; single-system problems: one member, no partial columns, this launch the only writer of the knot
// blocks): a lane's run of 2N doubles is 32..64 bytes, so a warp-wide 16-byte store of 32 such runs touches 32 different lines and
// the memory pipeline (ncu: L1/TEX the busiest unit, 2.9x the algorithmic L2 sectors), not HBM, bounds the direct-store kernel.
// The IPW items of a warp are IPW CONSECUTIVE knots, i.e. one contiguous range of F, of the Jacobian values and of the Hessian
// values.  With ST the lanes store to a shared-memory copy of those ranges at the very offsets they would use in global memory
// (kron blocks N times) and the warp then copies each range out linearly, 512 contiguous bytes per store instruction.  Two phases
// share the space: F + Jacobian, copied out, then the Hessian.  The member's constant matrices sit in shared memory for the whole
// kernel.  The derivative-integrator entries inside those ranges are written by the flat stand-alone aux kernel right after this
// one.  Measured and dropped (tools/column_bench.py, profiles/r02_column_staged.txt): the same entries filled in by the lanes
// inside the staged kernel, in three forms (table walk from global memory, from shared memory, entry descriptions resident in
// registers: +15..130 us -- with 8..12 warps per SM whatever a warp does serially adds to its critical path) and the inputs staged
// through shared memory by 512-byte loads (+3..9 us: the copy-in serialises with the items' own latency instead of hiding behind it).
template <int N, int ND, int NC, bool ST>
__global__ void __launch_bounds__(256) qck_column_kernel(const QckLaunch p, const int warp_doubles) {
    constexpr int n2 = 2 * N, blk = n2 * n2, IPW = 32 / NC, NPAIR = ND * (ND + 1) / 2;
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31;
    extern __shared__ double2 col_smem[];
    // per warp: the staged output ranges;  after the warps: the member's constant matrices (one member: they never change)
    double* const wimg = ST ? reinterpret_cast<double*>(col_smem) + (size_t)(threadIdx.x >> 5) * warp_doubles : nullptr;
    double* const wF = wimg;                                    // phase 1: [IPW][dyn] | [IPW][nnzJ]
    double* const wJ = ST ? wimg + IPW * c.dyn : nullptr;
    double* const wH = wimg;                                    // phase 2: [IPW][nnzH]
    double2* const sA0 = reinterpret_cast<double2*>(reinterpret_cast<double*>(col_smem) + (size_t)(blockDim.x >> 5) * warp_doubles);
    double2* const sAj = sA0 + N * N;
    if constexpr (ST) {
        const double2* const A0m = c.cmat + (size_t)p.member_begin * c.cmat_stride;
        const double2* const Ajm = c.dense_aj + (size_t)p.member_begin * ND * N * N;
        for (int i = threadIdx.x; i < N * N; i += blockDim.x) sA0[i] = A0m[i];
        for (int i = threadIdx.x; i < ND * N * N; i += blockDim.x) sAj[i] = Ajm[i];
        __syncthreads();
    }
    // linear copy of `count` doubles shared -> global (same parity of both addresses by construction: even knot, even offsets)
    auto copy_out = [&](double* g, const double* s, long long count) {
        __syncwarp();
        if ((reinterpret_cast<uintptr_t>(g) & 8) == 0) {
            const int pairs = (int)(count >> 1);
#if QCK_STORE_HINT
            for (int k = lane; k < pairs; k += 32) __stcs(reinterpret_cast<double2*>(g) + k, reinterpret_cast<const double2*>(s)[k]);
#else
            for (int k = lane; k < pairs; k += 32) reinterpret_cast<double2*>(g)[k] = reinterpret_cast<const double2*>(s)[k];
#endif
            if ((count & 1) && lane == 0) g[count - 1] = s[count - 1];
        } else {
            for (int k = lane; k < (int)count; k += 32) g[k] = s[k];
        }
        __syncwarp();
    };
    const int gi = lane / NC, col = lane - gi * NC;  // item slot inside the warp, column
    const bool needF = p.mask & QCK_EVAL_F, needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    const long long nslots = (long long)gridDim.x * (blockDim.x >> 5) * IPW;
    const bool free_time = c.free_time;

#ifndef QCK_DIRECT_STCS
#define QCK_DIRECT_STCS 0  // (A/B knob) direct variant: st.global.cs (L2 evict-first) for the value-array stores
#endif
    auto st2 = [](double* dst, double2 v) {
        if constexpr (!ST && QCK_DIRECT_STCS) __stcs(reinterpret_cast<double2*>(dst), v);
        else *reinterpret_cast<double2*>(dst) = v;
    };
    auto store_run = [&](double* dst, const double (&v)[n2]) {  // 2N consecutive doubles, 16-byte stores where aligned
        if (reinterpret_cast<uintptr_t>(dst) & 8) {
            dst[0] = v[0];
#pragma unroll
            for (int i = 0; i < N - 1; ++i) st2(dst + 1 + 2 * i, make_double2(v[1 + 2 * i], v[2 + 2 * i]));
            dst[n2 - 1] = v[n2 - 1];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) st2(dst + 2 * i, make_double2(v[2 * i], v[2 * i + 1]));
        }
    };
    auto rdot = [](double2 x, double2 y) { return x.x * y.x + x.y * y.y; };  // Re <x, y>
    auto gsum = [&](double v) {  // sum over the N lanes of this lane's item
        double r = v;
#pragma unroll
        for (int o = 1; o < NC; ++o) r += __shfl_sync(0xffffffffu, v, (gi * NC + (col + o) % NC) & 31);
        return r;
    };

    for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * IPW; base < n_items; base += nslots) {
        const long long item = base + gi;
        const bool on = gi < IPW && item < n_items;
        const long long it = on ? item : base;  // idle lanes shadow a valid item (no stores)
        const long long t = it / nact;
        const int mi = (int)(it - t * nact), m = p.member_begin + mi;
        const int soff = __ldg(p.moff_global + 3 * mi), coff = __ldg(p.moff_global + 3 * mi + 1), roff = __ldg(p.moff_global + 3 * mi + 2);
        const int n_on = (int)((n_items - base) < (long long)IPW ? (n_items - base) : (long long)IPW);
        const double* zt = p.Z + t * c.zdim;
        const double* mut = p.mu + t * c.dyn;
        auto ldin = [](const double* q) { return __ldg(q); };
        auto ldcm = [](const double2* q) { return ST ? *q : __ldg(q); };
        const int* qd = c.qdst + (size_t)m * QO_COUNT;
        const int gs = gi < IPW ? gi : 0;  // (idle lanes never store)
        double* const oF = ST ? wF + gs * c.dyn : p.F + t * c.dyn;
        double* const oJ = ST ? wJ + gs * p.nnzJ : p.J + t * p.nnzJ;
        double* const oH = ST ? wH + gs * p.nnzH : p.H + t * p.nnzH;
        // iso-vector quantity q: this lane's column (rows 0..N-1 real, then imaginary); arr0 = start of the knot block
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
            if (ST || d0 < p.nnzH) put_vec(oH, d0, q, x);
            else put_vec(p.partial + t * p.npart, d0 - (int)p.nnzH, q, x);
        };
        auto put_scalar = [&](int q, double v) {
            const int d0 = qd[q];
            if (d0 < 0 || !on || col != 0) return;
            if (ST || d0 < p.nnzH) oH[d0] = v;
            else p.partial[t * p.npart + (d0 - p.nnzH)] = v;
        };
        // ---- inputs: this lane's column of U0, U1 and of the multipliers ------------------------------------------------------
        double2 d[N], s[N], mm[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const double u0r = ldin(zt + soff + col * n2 + r), u0i = ldin(zt + soff + col * n2 + N + r);
            const double u1r = ldin(zt + c.zdim + soff + col * n2 + r), u1i = ldin(zt + c.zdim + soff + col * n2 + N + r);
            d[r] = make_double2(u1r - u0r, u1i - u0i);
            s[r] = make_double2(u1r + u0r, u1i + u0i);
            mm[r] = needH ? make_double2(ldin(mut + roff + col * n2 + r), ldin(mut + roff + col * n2 + N + r)) : make_double2(0.0, 0.0);
        }
        const double h = free_time ? ldin(zt + c.dt_off) : c.dt_fixed;
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
        double a[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) a[j] = ldin(zt + coff + j);
        // ---- A = A0 + sum_j a_j A_j ---------------------------------------------------------------------------------------------
        const double2* const A0g = ST ? sA0 : c.cmat + (size_t)m * c.cmat_stride;  // column-major
        const double2* const Ajg = ST ? sAj : c.dense_aj + (size_t)m * ND * N * N;  // [drive][row][column]
        double2 A[N][N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double2 v = ldcm(A0g + r + N * k);
#pragma unroll
                for (int j = 0; j < ND; ++j) {
                    const double2 w = ldcm(Ajg + (j * N + r) * N + k);
                    v.x = fma(a[j], w.x, v.x);
                    v.y = fma(a[j], w.y, v.y);
                }
                A[r][k] = v;
            }
        auto mvA = [&](double2 (&y)[N], const double2 (&x)[N]) {
#pragma unroll
            for (int r = 0; r < N; ++r) {
                y[r] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) cfma(y[r], A[r][k], x[k]);
            }
        };
        auto mvAH = [&](double2 (&y)[N], const double2 (&x)[N]) {
#pragma unroll
            for (int r = 0; r < N; ++r) {
                y[r] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) cfma(y[r], make_double2(A[k][r].x, -A[k][r].y), x[k]);
            }
        };

        // ---- residual, d/dh; q = -1/2 s + h/6 A d and v = -h/2 s + h^2/12 A d for the drive terms ---------------------------------
        double2 qv[N], vv[N];
        double s_hh = 0.0;
        {
            double2 x1[N], x2[N], x3[N], o[N];
            mvA(x1, s);
            mvA(x2, d);
            mvA(x3, x2);
#pragma unroll
            for (int r = 0; r < N; ++r) {
                qv[r] = make_double2(-0.5 * s[r].x + c2h * x2[r].x, -0.5 * s[r].y + c2h * x2[r].y);
                vv[r] = make_double2(-c1h * s[r].x + c2h2 * x2[r].x, -c1h * s[r].y + c2h2 * x2[r].y);
                s_hh += rdot(mm[r], x3[r]);
            }
            if (needF && on && qd[QO_R] >= 0) {
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(d[r].x - c1h * x1[r].x + c2h2 * x3[r].x, d[r].y - c1h * x1[r].y + c2h2 * x3[r].y);
                put_vec(oF, qd[QO_R], QO_R, o);
            }
            if (needJ) {
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-0.5 * x1[r].x + c2h * x3[r].x, -0.5 * x1[r].y + c2h * x3[r].y);
                put_J(QO_TH, o);
            }
        }
        // ---- -iso(F), +iso(B): this lane's columns (col, col + N) of the 2N x 2N block, into every one of the N copies -------------
        if (needJ && on && qd[QO_ISOF] >= 0) {
            const int dF = qd[QO_ISOF], dB = qd[QO_ISOB];
            auto block_columns = [&](int bc, const double2 (&acol)[N], int copy0, int copy1) {  // columns bc, bc + N of the block
                double2 a2[N];
                mvA(a2, acol);
                double f0[n2], f1[n2], b0[n2], b1[n2];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double id = r == bc ? 1.0 : 0.0;
                    const double fr = id + c1h * acol[r].x + c2h2 * a2[r].x, fi = c1h * acol[r].y + c2h2 * a2[r].y;
                    const double br = id - c1h * acol[r].x + c2h2 * a2[r].x, bi = -c1h * acol[r].y + c2h2 * a2[r].y;
                    f0[r] = -fr; f0[N + r] = -fi; f1[r] = fi; f1[N + r] = -fr;
                    b0[r] = br;  b0[N + r] = bi;  b1[r] = -bi; b1[N + r] = br;
                }
                for (int cb = copy0; cb < copy1; ++cb) {
                    store_run(oJ + dF + cb * blk + bc * n2, f0);
                    store_run(oJ + dF + cb * blk + (bc + N) * n2, f1);
                    if (dB >= 0) {
                        store_run(oJ + dB + cb * blk + bc * n2, b0);
                        store_run(oJ + dB + cb * blk + (bc + N) * n2, b1);
                    }
                }
            };
            if constexpr (NC == N) {  // unitary: this lane's column pair, into every one of the N copies
                double2 acol[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {  // column `col` of A (a lane-dependent column: rebuilt from the constants)
                    double2 v = ldcm(A0g + k + N * col);
#pragma unroll
                    for (int j = 0; j < ND; ++j) {
                        const double2 w = ldcm(Ajg + (j * N + k) * N + col);
                        v.x = fma(a[j], w.x, v.x);
                        v.y = fma(a[j], w.y, v.y);
                    }
                    acol[k] = v;
                }
                block_columns(col, acol, 0, N);
            } else {  // ket: the single lane writes all column pairs of the one block
#pragma unroll
                for (int bc = 0; bc < N; ++bc) {
                    double2 acol[N];
#pragma unroll
                    for (int k = 0; k < N; ++k) acol[k] = A[k][bc];
                    block_columns(bc, acol, 0, 1);
                }
            }
        }
        // ---- drive terms -----------------------------------------------------------------------------------------------------------------
        double2 u[ND][N];
        if (needJ || needH) {
#pragma unroll
            for (int j = 0; j < ND; ++j) {
                double2 y[N], y3[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    y[r] = u[j][r] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double2 w = ldcm(Ajg + (j * N + r) * N + k);
                        if (needJ) cfma(y[r], w, vv[k]);
                        cfma(u[j][r], w, d[k]);
                    }
                }
                if (needJ) {  // (a Hessian-only call needs A_j d only)
                    mvA(y3, u[j]);
#pragma unroll
                    for (int r = 0; r < N; ++r) y[r] = make_double2(y[r].x + c2h2 * y3[r].x, y[r].y + c2h2 * y3[r].y);
                    put_J(QO_TA + j, y);
                }
            }
        }
        if constexpr (ST) {  // phase 1 out: the warp's n_on knots of F and of the Jacobian values are one contiguous range each
            if (needF) copy_out(p.F + base * c.dyn, wF, (long long)n_on * c.dyn);
            if (needJ) copy_out(p.J + base * p.nnzJ, wJ, (long long)n_on * p.nnzJ);
        }
        if (needH) {
            double2 w1[N];
            double s_ah[ND], pz[ND][ND];  // pz[i][j] = Re <A_i^H m, A_j d> (this lane's column)
            {
                double2 w2[N], o[N];
                mvAH(w1, mm);
                mvAH(w2, w1);
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-0.5 * w1[r].x - c2h * w2[r].x, -0.5 * w1[r].y - c2h * w2[r].y);
                put_H(QO_KH0, o);
#pragma unroll
                for (int r = 0; r < N; ++r) o[r] = make_double2(-0.5 * w1[r].x + c2h * w2[r].x, -0.5 * w1[r].y + c2h * w2[r].y);
                put_H(QO_KH1, o);
            }
#pragma unroll
            for (int j = 0; j < ND; ++j) {
                double2 z1[N], z2[N], z3[N], o[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    z1[r] = z2[r] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        double2 w = ldcm(Ajg + (j * N + k) * N + r);  // conj(A_j[k][r])
                        w.y = -w.y;
                        cfma(z1[r], w, mm[k]);
                        cfma(z2[r], w, w1[k]);
                    }
                }
                mvAH(z3, z1);
                s_ah[j] = 0.0;
#pragma unroll
                for (int i2 = 0; i2 < ND; ++i2) pz[j][i2] = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double cr = c2h2 * (z2[r].x + z3[r].x), ci = c2h2 * (z2[r].y + z3[r].y);
                    o[r] = make_double2(-c1h * z1[r].x - cr, -c1h * z1[r].y - ci);
                    z2[r] = make_double2(-c1h * z1[r].x + cr, -c1h * z1[r].y + ci);
                    s_ah[j] += rdot(z1[r], qv[r]) + c2h * rdot(w1[r], u[j][r]);
#pragma unroll
                    for (int i2 = 0; i2 < ND; ++i2) pz[j][i2] += rdot(z1[r], u[i2][r]);
                }
                put_H(QO_KA0 + j, o);
                put_H(QO_KA1 + j, z2);
            }
            double s_aa[NPAIR];
#pragma unroll
            for (int j = 0, q = 0; j < ND; ++j)
#pragma unroll
                for (int i2 = 0; i2 <= j; ++i2, ++q) s_aa[q] = pz[i2][j] + pz[j][i2];
            s_hh = gsum(s_hh);
            put_scalar(QO_HHH, s_hh * (1.0 / 6.0));
#pragma unroll
            for (int j = 0; j < ND; ++j) put_scalar(QO_HAH + j, gsum(s_ah[j]));
#pragma unroll
            for (int j = 0, q = 0; j < ND; ++j)
#pragma unroll
                for (int i2 = 0; i2 <= j; ++i2, ++q) put_scalar(qo_haa(i2, j), c2h2 * gsum(s_aa[q]));
        }
        if constexpr (ST) {
            if (needH) copy_out(p.H + base * p.nnzH, wH, (long long)n_on * p.nnzH);
        } else {
            if (mi == 0 && p.n_aux && on) do_aux(p, t, col, NC);  // derivative-integrator entries of this knot
        }
    }
}


}  // namespace

// one lane per column, everything in registers (2..4-level Pade-4 unitaries, up to four drives, any number of members)
int qck_launch_column(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    static const int enabled = getenv("QCK_COLUMN") ? atoi(getenv("QCK_COLUMN")) : 1;
    const bool ket = c.kind == QCK_KET_PADE;
    if (!enabled || (c.kind != QCK_UNITARY_PADE && !ket) || c.order != 4 || c.N < 2 || c.N > 4 || c.nd < 1 || c.nd > 4 || !c.dense_aj || !c.qdst) return 0;
    typedef void (*kern_t)(const QckLaunch, const int);
    kern_t kern = nullptr;
    // block-staged variant (single-system problems, unitaries and kets; QCK_COLUMN_STAGED=0: loads and stores straight from / to
    // global memory as in round 1, which ensembles and several kets sharing their controls keep)
    static const int staged_knob = getenv("QCK_COLUMN_STAGED") ? atoi(getenv("QCK_COLUMN_STAGED")) : 1;
    // (short calls stay on the direct variant: one launch instead of two, and nothing to gain below a few waves of warps)
    static const long long staged_min = getenv("QCK_COLUMN_STAGED_MIN") ? atoll(getenv("QCK_COLUMN_STAGED_MIN")) : 1024;
    bool staged = staged_knob && L.sole_writer && L.member_end - L.member_begin == 1 && L.npart == 0 && L.n_knots >= staged_min;
    const int ipw = 32 / (ket ? 1 : c.N);
    int warp_doubles = 0, wpc = 8;
    size_t smem = 0;
    if (staged) {
        const long long od = ((long long)ipw * std::max<long long>((long long)c.dyn + L.nnzJ, L.nnzH) + 1) & ~1LL;  // staged outputs
        warp_doubles = (int)od;
        const size_t tab = (size_t)(1 + c.nd) * c.N * c.N * sizeof(double2);
        static const int wpc_knob = getenv("QCK_COLUMN_WPC") ? atoi(getenv("QCK_COLUMN_WPC")) : 4;  // warps per CTA (1..8 measured: flat within 3 %)
        wpc = (int)std::min<long long>(std::max(1, std::min(8, wpc_knob)), (200LL * 1024) / ((long long)warp_doubles * 8));
        if ((200LL * 1024) / ((long long)warp_doubles * 8) < 4) { staged = false; wpc = 8; }  // (blocks too large to stage with enough warps per SM)
        else smem = (size_t)wpc * warp_doubles * sizeof(double) + tab;
    }
#define QCK_COL3(N_, ND_, NC_) (staged ? qck_column_kernel<N_, ND_, NC_, true> : qck_column_kernel<N_, ND_, NC_, false>)
#define QCK_COL2(N_, NC_) (c.nd == 1 ? QCK_COL3(N_, 1, NC_) : (c.nd == 2 ? QCK_COL3(N_, 2, NC_) : (c.nd == 3 ? QCK_COL3(N_, 3, NC_) : QCK_COL3(N_, 4, NC_))))
#define QCK_COL(N_) (ket ? QCK_COL2(N_, 1) : QCK_COL2(N_, N_))
    kern = c.N == 2 ? QCK_COL(2) : (c.N == 3 ? QCK_COL(3) : QCK_COL(4));
#undef QCK_COL
#undef QCK_COL2
#undef QCK_COL3
    const int threads = 32 * wpc;
    int per_sm = 0;
    if (L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem) {
        per_sm = L.plan->per_sm;
    } else {
        cudaError_t e = cudaSuccess;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return (int)e;
        if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = per_sm; }
    }
    if (per_sm < 1) return 0;
    const long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    long long grid = (long long)sm_count * per_sm;
    const long long need = (n_items + (long long)wpc * ipw - 1) / ((long long)wpc * ipw);
    if (grid > need) grid = need;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] column kernel: N=%d nd=%d staged=%d (%d doubles/warp, %d warps/CTA) CTAs/SM=%d grid=%lld items=%lld\n", c.N, c.nd, (int)staged, warp_doubles, wpc, per_sm, grid, n_items);
    if (staged && L.n_aux) {  // whole knot blocks are stored: the derivative-integrator entries follow in their own pass
        kern<<<(unsigned)grid, threads, smem, stream>>>(L, warp_doubles);
        if (launches) ++*launches;
        *done = true;
        const int e = (int)cudaGetLastError();
        return e ? e : qck_launch_aux(L, stream, launches);
    }
    kern<<<(unsigned)grid, threads, smem, stream>>>(L, warp_doubles);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}

