// Column-group kernel "rs3" (9-level Pade-4 unitaries: the two-transmon CZ problem, the headline path) of libqcknot.so.
//
// THREE WARPS per knot, one per group of three columns of the unitaries.  Every product of the path acts column-wise
// (R[:,c] = D[:,c] - h/2 A S[:,c] + ...), so the column groups never exchange vectors: lane (c, k) of warp g owns element
// (row k, column 3g + c) of every state-shaped quantity (27 lanes) and keeps ROW k of A = -i H(a) in registers for the whole
// knot.  A dense product is then ONE complex dot product per lane, y[k] = sum_j A[k][j] x[j], x = the lane's column read from
// shared memory (the nine lanes of a column read the same address: broadcast), and the result goes back to the warp's own
// slice of the vector -- __syncwarp is the only synchronisation between dependent products.  Per knot the three warps meet
// at five named barriers: A assembled; staging buffer free / image complete, once per output phase.
//
// Against the one-warp-per-knot kernel (qck_rowslice.cu) the per-knot latency drops ~3x at the same shared memory per knot
// (staging 13 KB + vectors 14 KB), so 6-8 knots = 18-24 warps are resident per SM instead of 8 and the write-out of one knot
// overlaps the arithmetic of the others.  Algebra (same as DESIGN.md 4.1; d, s, m = columns of D = U1-U0, S = U1+U0, M):
//     R = d - h/2 A s + h^2/12 A (A d)                    d/dh = -1/2 A s + h/6 A (A d)
//     d/da_j = A_j (-h/2 s + h^2/12 A d) + h^2/12 A (A_j d)
//     state x dt:  -(1/2 w1 + h/6 A^H w1),  w1 = A^H m     state x a_j:  -(h/2 z1 + h^2/12 (A_j^H w1 + A^H z1)),  z1 = A_j^H m
//     dt x dt = 1/6 sum Re<m, A A d>      a_j x dt = sum -1/2 Re<z1_j, s> + h/6 (Re<z1_j, A d> + Re<w1, A_j d>)
//     a_i x a_j = h^2/12 Re tr({A_i, A_j} G),  G = D M^H = sum over the column groups of partial products (linear: every warp
//     traces its own partial G; the partial scalars of the three warps are added after a barrier)
//
// Write-out.  The host lays the image out in the solver's structure order with every run at an image offset of the SAME
// PARITY as its destination (qck_host.cpp: place_array(parity)); the kernel shifts the image of an array by one double when
// that array's knot block starts at 8 mod 16 (odd nnzH flips the Hessian base every knot).  Every unit is then
// [scalar head] + 16-byte aligned TMA bulk copy (cp.async.bulk.global.shared::cta, SASS UBLKCP) + [scalar tail], decoded and
// issued by ONE LANE per unit, all lanes of the three warps in parallel; kron(I_N, .) blocks are copied N times from one image.
#include "qck_device.cuh"

namespace {

__device__ __forceinline__ void bar_knot(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }

// one lane = one unit: [head] + bulk body + [tail], nrep back-to-back copies
__device__ __forceinline__ void rs3_flush(const double* __restrict__ stage, const QckSeg* __restrict__ units, int u0, int u1, int lane,
                                          double* baseF, double* baseJ, double* baseH, int shF, int shJ, int shH, int hoff, unsigned mask) {
    for (int u = u0 + lane; u < u1; u += 32) {
        const QckSeg sg = units[u];
        const int arr = sg.arr & 255;
        if (!((mask >> arr) & 1u)) continue;
        double* dst = (arr == 0 ? baseF : (arr == 1 ? baseJ : baseH)) + sg.dst;
        const int off = sg.img_nrep & 0xffff, nrep = sg.img_nrep >> 16, n = sg.n;
        const double* src = stage + (arr == 2 ? off - hoff + shH : off + (arr == 0 ? shF : shJ));
        const int head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);  // == parity of src by construction
        const int body = (n - head) & ~1;
        for (int r = 0; r < nrep; ++r, dst += n) {
            if (head) dst[0] = src[0];
            if (body) bulk_store(dst + head, src + head, (unsigned)body * 8u);
            if (head + body < n) dst[n - 1] = src[n - 1];
        }
    }
}

// KPC: knots per CTA (3 warps each).  WC: compile-time width of the sparse rows of the drives (0 = run-time width c.W).
// AH: A is anti-Hermitian (Hermitian Hamiltonians): A^H x = -(A x) runs on the register-resident row.
template <int ND, int WC, bool AH, int KPC>
__global__ void __launch_bounds__(96 * KPC, 1) qck_rs3_kernel(const QckLaunch p) {
    constexpr int N = 9, NN = 81, n2 = 18;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = warp / 3, g = warp - 3 * slot;            // knot slot of the CTA, column group
    const bool act = lane < 27;
    const int cl = act ? lane / 9 : 0, k = act ? lane - 9 * (lane / 9) : 0;  // local column, row
    const int cc = 3 * g + cl;                                  // column of the unitaries
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const int W = WC > 0 ? WC : c.W, elln = c.ell_stride, kkc = c.kk_cap, acn = c.ac_cap;
    // ---- CTA-wide tables: [A0 | sparse rows of A_j, A_j^H | anticommutator lists | contributor lists of A], units ------------
    double2* const conv = reinterpret_cast<double2*>(smem_all);
    const int nconv = NN + elln + kkc + acn;
    int* const coni = reinterpret_cast<int*>(conv + nconv);
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    QckSeg* const segtab = reinterpret_cast<QckSeg*>(smem_all + (((size_t)nconv * 16 + (size_t)c.icon_stride * 4 + 15) & ~(size_t)15));
    // ---- per knot slot: staging (phase 1: F + J part of the image, phase 2: Hessian part, same space) + vectors ----------------
    const int hoff = p.hoff, stage_doubles = (hoff > c.img_doubles - hoff ? hoff : c.img_doubles - hoff) + 4;
    const int stage_bytes = ((stage_doubles + 1) & ~1) * 8;
    constexpr int NVEC = 7 + ND;  // D S M X2 V W1 | U_j / Z1_j | A
    unsigned char* const sbase = reinterpret_cast<unsigned char*>(segtab + nrec) + (size_t)slot * (stage_bytes + NVEC * NN * 16 + 3 * 16 * 8);
    double* const stage = reinterpret_cast<double*>(sbase);
    double2* const vD = reinterpret_cast<double2*>(sbase + stage_bytes);  // element [column * 9 + row]
    double2* const vS = vD + NN;
    double2* const vM = vS + NN;
    double2* const vX2 = vM + NN;   // A D
    double2* const vV = vX2 + NN;   // -h/2 S + h^2/12 A D
    double2* const vW1 = vV + NN;   // A^H M
    double2* const vU = vW1 + NN;   // A_j D, j = 0 .. ND-1   (phase 2: A_j^H M)
    double2* const mA = vU + ND * NN;  // A, row-major
    double* const red = reinterpret_cast<double*>(mA + NN);  // [3 warps][16] partial scalars
    {
        const int m0 = p.member_begin;
        const double2* gv = c.cmat + (size_t)m0 * c.cmat_stride;
        const int* gc = c.ell_col + (size_t)m0 * c.icon_stride;
        for (int e = threadIdx.x; e < nconv; e += blockDim.x) conv[e] = gv[e];
        for (int e = threadIdx.x; e < c.icon_stride; e += blockDim.x) coni[e] = gc[e];
        const QckSeg* gs = c.segs + (size_t)m0 * nrec;
        for (int i = threadIdx.x; i < nrec; i += blockDim.x) segtab[i] = gs[i];
        if (g == 0)
            for (int i = lane; i < stage_bytes / 8; i += 32) stage[i] = 0.0;
        __syncthreads();
    }
    const double2* const A0 = conv;
    const double2* const ellv = conv + NN;
    const double2* const kkv = ellv + elln;
    const double2* const acv = kkv + kkc;
    const int* const ellc = coni;
    const int* const kkptr = coni + elln;
    const int* const kkrc = kkptr + ND * (ND + 1) / 2 + 1;
    const int* const acptr = kkrc + kkc;
    const int* const acj = acptr + NN + 1;
    const int* const hdr = reinterpret_cast<const int*>(segtab);   // [phase][warp] first unit; hdr[3] / hdr[7] = ends
    const QckSeg* const units = segtab + QCK_SEG_HDR / 4;
    const int soff = p.moff_global[0], coff = p.moff_global[1], roff = p.moff_global[2];
    const int own = cc * N + k;                    // this lane's element of the vectors
    const int colb = cc * N;                       // this lane's column
    const int bar_id = 1 + slot;
    const long long tstep = (long long)gridDim.x * KPC;
    long long t = (long long)slot * gridDim.x + blockIdx.x;  // slot-major: the partial last round spreads over all SMs

    // inputs of the first knot
    double nx0r = 0.0, nx0i = 0.0, nx1r = 0.0, nx1i = 0.0, nmr = 0.0, nmi = 0.0;
    auto load_inputs = [&](long long tt) {
        if (tt < p.n_knots && act) {
            const double* z = p.Z + tt * c.zdim + soff + cc * n2 + k;
            nx0r = z[0]; nx0i = z[N];
            nx1r = z[c.zdim]; nx1i = z[c.zdim + N];
            if (needH) {
                const double* mu = p.mu + tt * c.dyn + roff + cc * n2 + k;
                nmr = mu[0]; nmi = mu[N];
            }
        }
    };
    load_inputs(t);

    for (; t < p.n_knots; t += tstep) {
        const double* zt = p.Z + t * c.zdim;
        const double h = c.free_time ? __ldg(zt + c.dt_off) : c.dt_fixed;
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
        // ---- P0: own elements, A assembled by the three warps (27 elements each) -----------------------------------------------------
        const double2 d = make_double2(nx1r - nx0r, nx1i - nx0i), s = make_double2(nx1r + nx0r, nx1i + nx0i), m = make_double2(nmr, nmi);
        if (act) {
            vD[own] = d; vS[own] = s;
            if (needH) vM[own] = m;
            const int e = 27 * g + lane;  // column-major index into A0
            double2 v = A0[e];
            for (int u = acptr[e]; u < acptr[e + 1]; ++u) {
                const double aj = __ldg(zt + coff + acj[u]);
                const double2 dv = acv[u];
                v.x = fma(aj, dv.x, v.x);
                v.y = fma(aj, dv.y, v.y);
            }
            mA[(e % N) * N + e / N] = v;
        }
        load_inputs(t + tstep);  // the next knot's inputs travel while this one computes
        if (p.n_aux) do_aux(p, t, g * 32 + lane, 96);  // derivative-integrator entries of this knot (straight to the value arrays)
        double* const baseF = p.F + t * c.dyn;
        double* const baseJ = p.J + t * p.nnzJ;
        double* const baseH = p.H + t * p.nnzH;
        const int shF = (int)((reinterpret_cast<uintptr_t>(baseF) >> 3) & 1), shJ = (int)((reinterpret_cast<uintptr_t>(baseJ) >> 3) & 1),
                  shH = (int)((reinterpret_cast<uintptr_t>(baseH) >> 3) & 1);
        double* const imgF = stage + shF;
        double* const imgJ = stage + shJ;
        double* const imgH = stage - hoff + shH;
        bar_knot(bar_id);
        // ---- P1a: row k of A into registers, first products ------------------------------------------------------------------------------
        double2 Ar[N];
#pragma unroll
        for (int j = 0; j < N; ++j) Ar[j] = mA[k * N + j];
        double2 Ahr[AH ? 1 : N];  // row k of A^H (general case)
        if constexpr (!AH) {
#pragma unroll
            for (int j = 0; j < N; ++j) { Ahr[j] = mA[j * N + k]; Ahr[j].y = -Ahr[j].y; }
        }
        auto mv = [&](const double2* x) {  // sum_j A[k][j] x[j], two accumulator chains per component
            double2 y0 = make_double2(0.0, 0.0), y1 = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (j & 1) cfma(y1, Ar[j], x[j]);
                else cfma(y0, Ar[j], x[j]);
            }
            return make_double2(y0.x + y1.x, y0.y + y1.y);
        };
        auto mvAH = [&](const double2* x) {  // sum_j conj(A[j][k]) x[j]
            if constexpr (AH) {
                const double2 y = mv(x);
                return make_double2(-y.x, -y.y);
            } else {
                double2 y0 = make_double2(0.0, 0.0), y1 = make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (j & 1) cfma(y1, Ahr[j], x[j]);
                    else cfma(y0, Ahr[j], x[j]);
                }
                return make_double2(y0.x + y1.x, y0.y + y1.y);
            }
        };
        auto rdot = [](double2 x, double2 y) { return x.x * y.x + x.y * y.y; };  // Re <x, y>
        auto put = [&](double* image, int q, double2 v) {  // element (row k, column cc) of an iso-vector quantity
            const int b = c.pl_base[q], st = c.pl_stride[q], ire = cc * n2 + k;
            if (b >= 0 && act) { image[b + ire * st] = v.x; image[b + (ire + N) * st] = v.y; }
        };
        // sparse rows of the constant drives: out += sum_w val[row][w] x[col[row][w]]  (adj = 1: rows of A_j^H)
        auto drive_row1 = [&](int j, int adj, const double2* x, double2& o0) {
            const int o = ((j * 2 + adj) * N + k) * W;
            if constexpr (WC > 0) {
#pragma unroll
                for (int w = 0; w < WC; ++w) cfma(o0, ellv[o + w], x[colb + ellc[o + w]]);
            } else {
                for (int w = 0; w < W; ++w) cfma(o0, ellv[o + w], x[colb + ellc[o + w]]);
            }
        };

        const double2 x1 = mv(vS + colb);
        double2 w1 = make_double2(0.0, 0.0);
        double s_hh = 0.0, s_ah[ND];
        {
            const double2 x2 = mv(vD + colb);
            if (needH) w1 = mvAH(vM + colb);
            if (act) {
                vX2[own] = x2;
                vV[own] = make_double2(-c1h * s.x + c2h2 * x2.x, -c1h * s.y + c2h2 * x2.y);
                if (needH) vW1[own] = w1;
            }
        }
        // u_j = A_j d  (its column is the operand of A (A_j d) after the next warp barrier)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            double2 u = make_double2(0.0, 0.0);
            drive_row1(j, 0, vD, u);
            if (act) vU[j * NN + own] = u;
            s_ah[j] = needH && act ? c2h * rdot(w1, u) : 0.0;
        }
        __syncwarp();
        const double2 x3 = mv(vX2 + colb);
        if (needH && act) s_hh = rdot(m, x3);
        // ---- staging buffer free? (the copy engine has read the previous knot's Hessian image) ----------------------------------------
        bulk_wait_read();
        bar_knot(bar_id);
        put(imgF, QO_R, make_double2(d.x - c1h * x1.x + c2h2 * x3.x, d.y - c1h * x1.y + c2h2 * x3.y));
        put(imgJ, QO_TH, make_double2(-0.5 * x1.x + c2h * x3.x, -0.5 * x1.y + c2h * x3.y));
        if (needJ) {
            {   // element (k, cc) of A^2 -> -iso(F), +iso(B) with F/B = I +- h/2 A + h^2/12 A^2
                double2 acol[N];
#pragma unroll
                for (int j = 0; j < N; ++j) acol[j] = mA[j * N + cc];
                const double2 a2 = mv(acol);
                if (act) {
                    const double2 akc = mA[k * N + cc];
                    const double id = k == cc ? 1.0 : 0.0;
                    const double fr = id + c1h * akc.x + c2h2 * a2.x, fi = c1h * akc.y + c2h2 * a2.y;
                    const double br = id - c1h * akc.x + c2h2 * a2.x, bi = -c1h * akc.y + c2h2 * a2.y;
                    const int k00 = k + n2 * cc, k01 = k + n2 * (cc + N);
                    const int bF = c.pl_base[QO_ISOF], sF = c.pl_stride[QO_ISOF], bB = c.pl_base[QO_ISOB], sB = c.pl_stride[QO_ISOB];
                    imgJ[bF + k00 * sF] = -fr; imgJ[bF + (k00 + N) * sF] = -fi; imgJ[bF + k01 * sF] = fi; imgJ[bF + (k01 + N) * sF] = -fr;
                    imgJ[bB + k00 * sB] = br;  imgJ[bB + (k00 + N) * sB] = bi;  imgJ[bB + k01 * sB] = -bi; imgJ[bB + (k01 + N) * sB] = br;
                }
            }
#pragma unroll
            for (int j = 0; j < ND; ++j) {  // d/da_j = A_j (-h/2 s + h^2/12 A d) + h^2/12 A (A_j d)
                double2 y = make_double2(0.0, 0.0);
                drive_row1(j, 0, vV, y);
                const double2 y3 = mv(vU + j * NN + colb);
                put(imgJ, QO_TA + j, make_double2(y.x + c2h2 * y3.x, y.y + c2h2 * y3.y));
            }
        }
        fence_async_smem();
        bar_knot(bar_id);
        rs3_flush(stage, units, hdr[g], hdr[g + 1], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & (QCK_EVAL_F | QCK_EVAL_J));
        bulk_commit();
        // ---- phase 2: Hessian-of-Lagrangian values ---------------------------------------------------------------------------------------------
        if (needH) {
            // z1_j = A_j^H m: own element into the scalar sums, column shared through vU_j (its phase-1 readers passed the barrier)
            {
                const double2 x2 = vX2[own];
#pragma unroll
                for (int j = 0; j < ND; ++j) {
                    double2 z1 = make_double2(0.0, 0.0);
                    drive_row1(j, 1, vM, z1);
                    if (act) {
                        vU[j * NN + own] = z1;
                        s_ah[j] += -0.5 * rdot(z1, s) + c2h * rdot(z1, x2);
                    }
                }
            }
            // a_i x a_j partials of this column group: three lanes per pair walk the constant sparse anticommutator {A_i, A_j};
            // G[r][q] = sum_{c in group} D[r][c] conj(M[q][c]) is formed on the fly (three terms per entry)
            constexpr int NPAIR = ND * (ND + 1) / 2;
            const int pr = lane / 3, sub = lane - 3 * pr;
            double val = 0.0;
            if (pr < NPAIR)
                for (int u = kkptr[pr] + sub, u1 = kkptr[pr + 1]; u < u1; u += 3) {
                    const int rc = kkrc[u];
                    const double2 kv = kkv[u];
                    const int r = rc & 255, q = rc >> 8;
                    double2 gv = make_double2(0.0, 0.0);
#pragma unroll
                    for (int c3 = 0; c3 < 3; ++c3) {
                        double2 mc = vM[(3 * g + c3) * N + q];
                        mc.y = -mc.y;
                        cfma(gv, vD[(3 * g + c3) * N + r], mc);
                    }
                    val = fma(kv.x, gv.x, val);  // Re (K[q, r] G[r, q])
                    val = fma(-kv.y, gv.y, val);
                }
            val += __shfl_down_sync(0xffffffffu, val, 1) + __shfl_down_sync(0xffffffffu, val, 2);
            s_hh = warp_sum(s_hh);
#pragma unroll
            for (int j = 0; j < ND; ++j) s_ah[j] = warp_sum(s_ah[j]);
            double* const myred = red + g * 16;
            if (pr < NPAIR && sub == 0) myred[1 + ND + pr] = c2h2 * val;
            if (lane == 0) {
                myred[0] = s_hh * (1.0 / 6.0);
#pragma unroll
                for (int j = 0; j < ND; ++j) myred[1 + j] = s_ah[j];
            }
            const double2 w2 = mvAH(vW1 + colb);
            bulk_wait_read();  // phase-1 copies have left the staging buffer
            bar_knot(bar_id);  // (also: every warp's z1 columns and partial scalars are visible)
            if (g == 0 && lane < 1 + ND + NPAIR) {  // the three partial sums of every scalar -> image
                const double v = red[lane] + red[16 + lane] + red[32 + lane];
                int q;
                if (lane == 0) q = QO_HHH;
                else if (lane <= ND) q = QO_HAH + lane - 1;
                else {
                    int j = 0, rem = lane - 1 - ND;
                    while (rem > j) { rem -= j + 1; ++j; }
                    q = qo_haa(rem, j);
                }
                if (c.pl_base[q] >= 0) imgH[c.pl_base[q]] = v;
            }
            put(imgH, QO_KH0, make_double2(-0.5 * w1.x - c2h * w2.x, -0.5 * w1.y - c2h * w2.y));
            put(imgH, QO_KH1, make_double2(-0.5 * w1.x + c2h * w2.x, -0.5 * w1.y + c2h * w2.y));
#pragma unroll
            for (int j = 0; j < ND; ++j) {
                const double2 z1 = vU[j * NN + own];
                double2 z2 = make_double2(0.0, 0.0);
                drive_row1(j, 1, vW1, z2);
                const double2 z3 = mvAH(vU + j * NN + colb);
                const double cr = c2h2 * (z2.x + z3.x), ci = c2h2 * (z2.y + z3.y);
                put(imgH, QO_KA0 + j, make_double2(-c1h * z1.x - cr, -c1h * z1.y - ci));
                put(imgH, QO_KA1 + j, make_double2(-c1h * z1.x + cr, -c1h * z1.y + ci));
            }
            fence_async_smem();
            bar_knot(bar_id);
            rs3_flush(stage, units, hdr[4 + g], hdr[4 + g + 1], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & QCK_EVAL_H);
            bulk_commit();
        }
    }
    bulk_wait_all();
}

}  // namespace

// The file is compiled once per drive count (-DQCK_RS3_ND=1..4, in parallel); every object exports the kernel selector of
// its drive count, the ND = 4 object also carries the launcher.
#ifndef QCK_RS3_ND
#define QCK_RS3_ND 4
#endif
typedef void (*qck_rs3_kern_t)(const QckLaunch);
#define QCK_RS3_CAT2(a, b) a##b
#define QCK_RS3_CAT(a, b) QCK_RS3_CAT2(a, b)
qck_rs3_kern_t QCK_RS3_CAT(qck_rs3_get_, QCK_RS3_ND)(int wc, bool ah, int kpc) {
#define QCK_RS3_K(WC_, AH_) (kpc == 7 ? qck_rs3_kernel<QCK_RS3_ND, WC_, AH_, 7> : (kpc == 6 ? qck_rs3_kernel<QCK_RS3_ND, WC_, AH_, 6> : qck_rs3_kernel<QCK_RS3_ND, WC_, AH_, 5>))
#define QCK_RS3_W(AH_) (wc == 1 ? QCK_RS3_K(1, AH_) : (wc == 2 ? QCK_RS3_K(2, AH_) : QCK_RS3_K(0, AH_)))
    return ah ? QCK_RS3_W(true) : QCK_RS3_W(false);
#undef QCK_RS3_W
#undef QCK_RS3_K
}

#if QCK_RS3_ND == 4
qck_rs3_kern_t qck_rs3_get_1(int wc, bool ah, int kpc);
qck_rs3_kern_t qck_rs3_get_2(int wc, bool ah, int kpc);
qck_rs3_kern_t qck_rs3_get_3(int wc, bool ah, int kpc);

size_t qck_rs3_smem(const QckClassDev& c, int hoff, int kpc) {
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    const int stage_doubles = (hoff > c.img_doubles - hoff ? hoff : c.img_doubles - hoff) + 4;
    const size_t per_slot = (size_t)((stage_doubles + 1) & ~1) * 8 + (size_t)(7 + c.nd) * 81 * 16 + 3 * 16 * 8;
    const size_t shared = ((((size_t)(81 + c.ell_stride + c.kk_cap + c.ac_cap) * 16 + (size_t)c.icon_stride * 4) + 15) & ~(size_t)15) + (size_t)nrec * 16;
    return shared + (size_t)kpc * per_slot;
}

int qck_rs3_hoff(const QckClassDev& c) {
    int hoff = c.img_doubles;
    for (int q = 0; q < QO_COUNT; ++q) {
        const bool hq = q == QO_KH0 || q == QO_KH1 || (q >= QO_KA0 && q < QO_ONE);
        if (hq && c.pl_base[q] >= 0 && c.pl_base[q] < hoff) hoff = c.pl_base[q];
    }
    return hoff & ~1;
}

// three warps per knot (9-level Pade-4 unitaries, one active member, up to four drives); the class must have been built with
// the parity-matched image placement and the rs3 unit table (c.rs3 = knots per CTA)
int qck_launch_rs3(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    if (c.rs3 < 5) return 0;  // (1: the one-warp-per-knot kernel of qck_rowslice.cu takes the class)
    const int wc = c.W <= 2 ? c.W : 0;
    const int kpc = c.rs3;
    if (kpc > 7) return (int)cudaErrorInvalidConfiguration;
    qck_rs3_kern_t kern = c.nd == 1 ? qck_rs3_get_1(wc, c.antiherm, kpc) : (c.nd == 2 ? qck_rs3_get_2(wc, c.antiherm, kpc) :
                          (c.nd == 3 ? qck_rs3_get_3(wc, c.antiherm, kpc) : qck_rs3_get_4(wc, c.antiherm, kpc)));
    const int hoff = qck_rs3_hoff(c);
    const size_t smem = qck_rs3_smem(c, hoff, kpc);
    if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
    if (!(L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
        if (e != cudaSuccess) return (int)e;
        if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = 1; }
    }
    long long grid = sm_count;
    if (grid * kpc > L.n_knots) grid = (L.n_knots + kpc - 1) / kpc;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] rs3 kernel: N=9 nd=%d knots/CTA=%d smem=%zu B grid=%lld units=%d\n", c.nd, kpc, smem, grid, c.nseg);
    QckLaunch L2 = L;
    L2.hoff = hoff;
    kern<<<(unsigned)grid, 96 * kpc, smem, stream>>>(L2);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}
#endif
