// Row-slice kernel (9-level Pade-4 unitaries, one warp per knot) of libqcknot.so.  Compiled as its own translation unit so that the kernel families build in parallel.
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// Row-slice kernel: Pade-4, unitaries, 9 levels (the two-transmon CZ problem).  ONE WARP per knot, no block barriers.
//
// The columns of the unitaries are independent under every product of the path (R[:,c] = D[:,c] - h/2 A S[:,c] + ...), so
// lane (c, k) of a warp owns rows 3k .. 3k+2 of column c (27 lanes) and keeps ITS THREE ROWS OF A = -i H(a) IN REGISTERS
// for the whole knot.  Every dense product of the path becomes a row-slice matrix-vector product
//     y[3k + i] = sum_j A[3k + i][j] x[j],     x = a full column read from shared memory (all lanes of a column read the
// same address: broadcast), 27 complex FMAs per lane, no operand re-load from shared memory for A.  Products with the
// constant drives A_j, with A_j^H and with A^H read their matrix from shared memory the same way.  In matrix-vector form
//     R = d - h/2 A s + h^2/12 A (A d)                    d/dh = -1/2 A s + h/6 A (A d)
//     d/da_j = A_j (-h/2 s + h^2/12 A d) + h^2/12 A (A_j d)
//     state x dt:  -(1/2 w1 + h/6 A^H w1),  w1 = A^H m     state x a_j:  -(h/2 z1 + h^2/12 (A_j^H w1 + A^H z1)),  z1 = A_j^H m
//     dt x dt = 1/6 sum Re<m, A A d>      a_j x dt = sum -1/2 Re<z1_j, s> + h/6 (Re<z1_j, A d> + Re<w1, A_j d>)
//     a_i x a_j = h^2/12 Re tr({A_i, A_j} G),  G = D M^H     (sums over rows and columns = one warp reduction each)
// 8 + 2 n_d dense row-slice products and 4 n_d products with the (sparse) drives per knot.  Values go into the warp's
// staging buffer in the solver's structure order: first the residual + Jacobian part, flushed, then the Hessian part in the
// same space.  Write-out (round 2): the host places every run at an image offset of the SAME PARITY as its destination
// (place_array(parity)) and the kernel shifts an array's image by one double when that array's knot block starts at 8 mod 16
// (odd nnzH flips the Hessian base every knot), so every unit is [scalar head] + one 16-byte aligned TMA bulk copy + [scalar
// tail], decoded and issued by one lane per unit with all lanes in parallel (flush_units_lanes): ~10x fewer instructions than
// the unit-by-unit loop of round 1 and no misaligned slow path.
// ------------------------------------------------------------------------------------------------------------
// WC: compile-time width of the sparse rows of the drives (loops fully unrolled); 0 = dense drive matrices
// AH: A is anti-Hermitian (Hermitian Hamiltonians): A^H x = -(A x) runs on the register-resident rows of A
// BW: "block warps".  The two kron(I_N, .) blocks -iso(F), iso(B) are 69 % of a knot's bytes but need only A and A^2: four extra
//     warps per CTA (a third warpgroup that hands its registers to the eight computing warps with setmaxnreg) assemble A, form the
//     blocks in their own small buffers and issue the 2 x N bulk copies, on their own schedule -- no handshake with the computing
//     warps, which drop the A^2 product, the block stores and 18 of their 24 blocking bulk-copy issues per knot.
template <int ND, int WC, bool AH, bool BW>
__global__ void __launch_bounds__(BW ? 384 : 256, 1) qck_rowslice9_kernel(const QckLaunch p) {
    constexpr int N = 9, NN = 81, n2 = 18, dim = 162;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = BW ? 8 : (int)(blockDim.x >> 5);
    const bool act = lane < 3 * N;
    // (idle lanes 27..31 shadow lane 24: their reads then broadcast with column 8 instead of adding a bank conflict against column 0)
    const int cc = act ? lane / 3 : N - 1, k3 = act ? 3 * (lane - 3 * (lane / 3)) : 0;  // column, first row of this lane
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const bool needT = needJ || needH;
    const bool free_time = c.free_time;
    const int m = p.member_begin;
    // CTA-wide tables: the member's constant block [A0 | sparse rows of A_j, A_j^H | anticommutator lists | contributor
    // lists of A] (same layout as the tiled kernel's) and the write-out units
    const int W = WC > 0 ? WC : c.W, elln = c.ell_stride, kkc = c.kk_cap, acn = c.ac_cap;
    double2* const conv = reinterpret_cast<double2*>(smem_all);
    const int nconv = NN + elln + kkc + acn;
    int* const coni = reinterpret_cast<int*>(conv + nconv);
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    QckSeg* const segtab = reinterpret_cast<QckSeg*>(smem_all + (((size_t)nconv * 16 + (size_t)c.icon_stride * 4 + 15) & ~(size_t)15));
    // output staging: the F + J part of the image first, flushed, then the Hessian part in the same space
    // staging: the F + J part of the image (image offsets < hoff) and the Hessian part (>= hoff), each + 4 doubles of slack for
    // the parity shifts.  p.db: two separate buffers -- the copy engine drains one phase's image while the warp computes the
    // next phase (the wait before re-using a buffer then concerns copies issued a whole phase earlier); else both share one.
    const int hoff = p.hoff;
    const int jbytes = ((hoff + 4 + 1) & ~1) * 8, hbytes = ((c.img_doubles - hoff + 4 + 1) & ~1) * 8;
    const int img_bytes = p.db ? jbytes + hbytes : (jbytes > hbytes ? jbytes : hbytes);
    QckAux* const auxs = reinterpret_cast<QckAux*>(segtab + nrec);   // derivative-integrator entries (CTA-wide copy)
    double2* const cAj = reinterpret_cast<double2*>(auxs + p.n_aux);  // WC == 0: dense A_j, row-major
    unsigned char* const wbase0 = reinterpret_cast<unsigned char*>(cAj + (WC > 0 ? 0 : ND * NN));
    constexpr int kWarpExtra = ND * 32 * 8;  // per-lane partial sums of the a_j x dt entries, carried from phase 1 to phase 2
    unsigned char* const wbase = wbase0 + (size_t)(warp < nwarps ? warp : 0) * (img_bytes + 7 * NN * 16 + kWarpExtra);
    double* const stage = reinterpret_cast<double*>(wbase);
    double* const stageH = p.db ? stage + jbytes / 8 : stage;
    double2* const vD = reinterpret_cast<double2*>(wbase + img_bytes);  // columns of D = U1 - U0: element [c * 9 + r]
    double2* const vS = vD + NN;       // S = U1 + U0
    double2* const vM = vS + NN;       // multipliers
    double2* const vX2 = vM + NN;      // A D
    double2* const vW1 = vX2 + NN;     // A^H M
    double2* const vU = vW1 + NN;      // A_j D (current drive)
    double2* const vZ1 = vU;           // A_j^H M (current drive; phase 2 only, when A_j D is dead)
    double2* const mA = vU + NN;       // A, row-major (for A^H products and column access)
    double* const sah = reinterpret_cast<double*>(mA + NN) + lane;  // [drive][lane]
    {
        const double2* gv = c.cmat + (size_t)m * c.cmat_stride;
        const int* gc = c.ell_col + (size_t)m * c.icon_stride;
        for (int e = threadIdx.x; e < nconv; e += blockDim.x) conv[e] = gv[e];
        for (int e = threadIdx.x; e < c.icon_stride; e += blockDim.x) coni[e] = gc[e];
        const QckSeg* gs = c.segs + (size_t)m * nrec;
        for (int i = threadIdx.x; i < nrec; i += blockDim.x) segtab[i] = gs[i];
        if (warp < nwarps)
            for (int i = lane; i < img_bytes / 8; i += 32) stage[i] = 0.0;
        for (int i = threadIdx.x; i < p.n_aux; i += blockDim.x) auxs[i] = p.aux[i];
        if (WC == 0) {
            for (int e = threadIdx.x; e < ND * NN; e += blockDim.x) cAj[e] = make_double2(0.0, 0.0);
            __syncthreads();
            for (int w = threadIdx.x; w < ND * N * W; w += blockDim.x) {  // dense A_j from the fixed-width sparse rows
                const int j = w / (N * W), rem = w - j * N * W, r = rem / W, u = rem - r * W;
                const int o = ((j * 2) * N + r) * W + u;
                const double2 v = gv[NN + o];
                if (v.x != 0.0 || v.y != 0.0) cAj[j * NN + r * N + gc[o]] = v;
            }
        }
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
    const int* seghdr = reinterpret_cast<const int*>(segtab);   // [phase][warp of three] first unit; hdr[3] / hdr[7] = ends
    const QckSeg* segs = segtab + QCK_SEG_HDR / 4;
    const int soff = p.moff_global[0], coff = p.moff_global[1], roff = p.moff_global[2];
    const int xo = cc * N;  // this lane's column inside the vector buffers

    if constexpr (BW) {
        if (warp >= 8) {
            // ---- block warps: -iso(F), +iso(B) of the knots of this CTA, N copies each, straight from their own buffers ------------
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
            if (!needJ) return;
            const int lw = warp - 8;
            double* const bufF = reinterpret_cast<double*>(wbase0 + (size_t)8 * (img_bytes + 7 * NN * 16 + kWarpExtra)) + (size_t)lw * (2 * 328 + 2 * NN);
            double* const bufB = bufF + 328;
            double2* const sA = reinterpret_cast<double2*>(bufB + 328);  // A, row-major
            const int imgF = c.pl_base[QO_ISOF];
            const int u0 = seghdr[11], u1 = seghdr[12];
            int dF0 = 0, dB0 = 0;  // first destination of either block inside the knot block (parity of the copies)
            for (int u = u0; u < u1; ++u) {
                const QckSeg sg = segs[u];
                if ((sg.img_nrep & 0xffff) == imgF) { dF0 = sg.dst; break; }
            }
            for (int u = u0; u < u1; ++u) {
                const QckSeg sg = segs[u];
                if ((sg.img_nrep & 0xffff) != imgF) { dB0 = sg.dst; break; }
            }
            for (long long t = (long long)lw * gridDim.x + blockIdx.x; t < p.n_knots; t += (long long)gridDim.x * 4) {
                const double* zt = p.Z + t * c.zdim;
                const double h = free_time ? __ldg(zt + c.dt_off) : c.dt_fixed;
                const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0);
                for (int e = lane; e < NN; e += 32) {
                    double2 v = A0[e];
                    for (int u = acptr[e]; u < acptr[e + 1]; ++u) {
                        const double aj = __ldg(zt + coff + acj[u]);
                        const double2 d = acv[u];
                        v.x = fma(aj, d.x, v.x);
                        v.y = fma(aj, d.y, v.y);
                    }
                    sA[(e % N) * N + e / N] = v;  // A0 is column-major
                }
                double* const baseJ = p.J + t * p.nnzJ;
                const int pF = (int)((reinterpret_cast<uintptr_t>(baseJ + dF0) >> 3) & 1), pB = (int)((reinterpret_cast<uintptr_t>(baseJ + dB0) >> 3) & 1);
                bulk_wait_read();  // the previous knot's copies have left the buffers
                __syncwarp();
                for (int e = lane; e < NN; e += 32) {
                    const int r = e % N, col = e / N;
                    double2 a2 = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < N; ++k) cfma(a2, sA[r * N + k], sA[k * N + col]);
                    const double2 av = sA[r * N + col];
                    const double id = r == col ? 1.0 : 0.0;
                    const double fr = id + c1h * av.x + c2h2 * a2.x, fi = c1h * av.y + c2h2 * a2.y;
                    const double br = id - c1h * av.x + c2h2 * a2.x, bi = -c1h * av.y + c2h2 * a2.y;
                    const int k00 = r + n2 * col, k01 = r + n2 * (col + N);
                    double* const iF = bufF + pF;
                    double* const iB = bufB + pB;
                    iF[k00] = -fr; iF[k00 + N] = -fi; iF[k01] = fi; iF[k01 + N] = -fr;
                    iB[k00] = br;  iB[k00 + N] = bi;  iB[k01] = -bi; iB[k01 + N] = br;
                }
                fence_async_smem();
                __syncwarp();
                for (int u = u0 + lane; u < u1; u += 32) {  // one lane, one copy of one block
                    const QckSeg sg = segs[u];
                    double* dst = baseJ + sg.dst;
                    const bool isF = (sg.img_nrep & 0xffff) == imgF;
                    const double* src = isF ? bufF + pF : bufB + pB;
                    const int n = sg.n, head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1), body = (n - head) & ~1;
                    if (head) dst[0] = src[0];
                    if (body) bulk_store(dst + head, src + head, (unsigned)body * 8u);
                    if (head + body < n) dst[n - 1] = src[n - 1];
                }
                bulk_commit();
            }
            bulk_wait_all();
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    }

    // warp-major slots: the last, partially filled round spreads over all SMs (fewer active warps per SM run faster each)
    for (long long t = (long long)warp * gridDim.x + blockIdx.x; t < p.n_knots; t += (long long)gridDim.x * nwarps) {
        const double* zt = p.Z + t * c.zdim;
        // ---- inputs: coalesced loads of the two state vectors and the multipliers, unpacked into complex columns ----------
        constexpr int NLD = (dim + 31) / 32;
        double in0[NLD], in1[NLD], inm[NLD];  // all global loads of the knot are issued before the first use
        {
            const double* mut = p.mu + t * c.dyn + roff;
#pragma unroll
            for (int q = 0; q < NLD; ++q) {
                const int idx = lane + 32 * q;
                const bool ok = idx < dim;
                in0[q] = ok ? zt[soff + idx] : 0.0;
                in1[q] = ok ? zt[c.zdim + soff + idx] : 0.0;
                inm[q] = ok && needH ? mut[idx] : 0.0;
            }
        }
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        double ctl[ND];  // the controls a_j of this knot: issued with the other global loads, used by the assembly of A
#pragma unroll
        for (int j = 0; j < ND; ++j) ctl[j] = zt[coff + j];
        double* const baseF = p.F + t * c.dyn;
        double* const baseJ = p.J + t * p.nnzJ;
        double* const baseH = p.H + t * p.nnzH;
        const int shF = (int)((reinterpret_cast<uintptr_t>(baseF) >> 3) & 1), shJ = (int)((reinterpret_cast<uintptr_t>(baseJ) >> 3) & 1),
                  shH = (int)((reinterpret_cast<uintptr_t>(baseH) >> 3) & 1);
        double* const imgF = stage + shF;
        double* const imgJ = stage + shJ;
        double* const imgH = stageH - hoff + shH;
        if (t + (long long)gridDim.x * nwarps < p.n_knots) {  // pull the next knot of this warp into L2 meanwhile
            const double* zn = zt + (long long)gridDim.x * nwarps * c.zdim;
            const double* mn = p.mu + (t + (long long)gridDim.x * nwarps) * c.dyn + roff;
            for (int b = lane * 16; b < 2 * c.zdim; b += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(zn + b));
            if (needH)
                for (int b = lane * 16; b < dim; b += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(mn + b));
        }
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int idx = lane + 32 * q;
            if (idx < dim) {
                const int col = idx / n2, qq = idx - col * n2, im = qq >= N, r = qq - im * N;
                const int o = 2 * (col * N + r) + im;
                reinterpret_cast<double*>(vD)[o] = in1[q] - in0[q];
                reinterpret_cast<double*>(vS)[o] = in1[q] + in0[q];
                if (needH) reinterpret_cast<double*>(vM)[o] = inm[q];
            }
        }
        // ---- A = A0 + sum_j a_j A_j (per-element contributor lists) into shared memory, this lane's three rows into registers
        for (int e = lane; e < NN; e += 32) {
            double2 v = A0[e];
            for (int u = acptr[e]; u < acptr[e + 1]; ++u) {
                const int jd = acj[u];
                double aj = ctl[0];
#pragma unroll
                for (int j = 1; j < ND; ++j) aj = jd == j ? ctl[j] : aj;
                const double2 d = acv[u];
                v.x = fma(aj, d.x, v.x);
                v.y = fma(aj, d.y, v.y);
            }
            mA[(e % N) * N + e / N] = v;  // A0 is column-major
        }
        __syncwarp();
        double2 Ar[3][N];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) Ar[i][j] = mA[(k3 + i) * N + j];
        // row-slice products: y[i] = sum_j Mat[3k + i][j] x[j]
        auto mv_reg = [&](double2 (&y)[3], const double2* x) {
#pragma unroll
            for (int i = 0; i < 3; ++i) y[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 xv = x[j];
#pragma unroll
                for (int i = 0; i < 3; ++i) cfma(y[i], Ar[i][j], xv);
            }
        };
        auto mvH = [&](double2 (&y)[3], const double2* Mat, const double2* x) {  // y = Mat^H x
#pragma unroll
            for (int i = 0; i < 3; ++i) y[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 xv = x[j];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double2 mv = Mat[j * N + k3 + i];
                    mv.y = -mv.y;
                    cfma(y[i], mv, xv);
                }
            }
        };
        auto put = [&](double* image, int q, int i, double2 v) {  // element (row k3 + i, column cc) of an iso-vector quantity
            const int b = c.pl_base[q], s = c.pl_stride[q], ire = cc * n2 + k3 + i;
            if (b >= 0 && act) { image[b + ire * s] = v.x; image[b + (ire + N) * s] = v.y; }
        };
        auto rdot = [](double2 x, double2 y) { return x.x * y.x + x.y * y.y; };  // Re <x, y>
        auto mvAH = [&](double2 (&y)[3], const double2* x) {  // y = A^H x
            if constexpr (AH) {
                mv_reg(y, x);
#pragma unroll
                for (int i = 0; i < 3; ++i) y[i] = make_double2(-y[i].x, -y[i].y);
            } else {
                mvH(y, mA, x);
            }
        };

        double s_hh = 0.0;
        double2 w1[3];
        {
            double2 x1[3], x2[3], x3[3];
            mv_reg(x1, vS + xo);
            mv_reg(x2, vD + xo);
            if (act) {
#pragma unroll
                for (int i = 0; i < 3; ++i) vX2[xo + k3 + i] = x2[i];
            }
            if (needH) {
                mvAH(w1, vM + xo);
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vW1[xo + k3 + i] = w1[i];
                }
            }
            __syncwarp();
            mv_reg(x3, vX2 + xo);
            if (needH && act) {
#pragma unroll
                for (int i = 0; i < 3; ++i) s_hh += rdot(vM[xo + k3 + i], x3[i]);
            }
            // the copy engine has finished reading this buffer's previous image (two buffers: the group before the last one)
            if (p.db && needH) bulk_wait_read_1();
            else bulk_wait_read();
            __syncwarp();
            // ---- phase 1: residual and Jacobian values ---------------------------------------------------------------------------
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 d = vD[xo + k3 + i];
                put(imgF, QO_R, i, make_double2(d.x - c1h * x1[i].x + c2h2 * x3[i].x, d.y - c1h * x1[i].y + c2h2 * x3[i].y));
                put(imgJ, QO_TH, i, make_double2(-0.5 * x1[i].x + c2h * x3[i].x, -0.5 * x1[i].y + c2h * x3[i].y));
            }
        }
        if (needJ && !BW) {  // column cc of A^2 -> -iso(F), +iso(B)  (BW: the block warps do this)
            double2 a2[3], acol[N];
#pragma unroll
            for (int j = 0; j < N; ++j) acol[j] = mA[j * N + cc];
            mv_reg(a2, acol);
            if (act) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int r = k3 + i;
                    const double2 av = mA[r * N + cc];
                    const double id = r == cc ? 1.0 : 0.0;
                    const double fr = id + c1h * av.x + c2h2 * a2[i].x, fi = c1h * av.y + c2h2 * a2[i].y;
                    const double br = id - c1h * av.x + c2h2 * a2[i].x, bi = -c1h * av.y + c2h2 * a2[i].y;
                    const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                    const int bF = c.pl_base[QO_ISOF], sF = c.pl_stride[QO_ISOF], bB = c.pl_base[QO_ISOB], sB = c.pl_stride[QO_ISOB];
                    imgJ[bF + k00 * sF] = -fr; imgJ[bF + (k00 + N) * sF] = -fi; imgJ[bF + k01 * sF] = fi; imgJ[bF + (k01 + N) * sF] = -fr;
                    imgJ[bB + k00 * sB] = br;  imgJ[bB + (k00 + N) * sB] = bi;  imgJ[bB + k01 * sB] = -bi; imgJ[bB + (k01 + N) * sB] = br;
                }
            }
            if (p.spread) {  // the 2 x N block copies leave now and drain behind the drive loop (shorter bursts at the copy engine)
                fence_async_smem();
                __syncwarp();
                flush_units_lanes(stage, segs, seghdr[11], seghdr[12], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & QCK_EVAL_J);
            }
        }
        if (needJ) {  // (a Hessian-only call gets its A_j d inside the Hessian loop and skips this one)
            // (the drive loops are NOT unrolled: the straight-line kernel of round 1 was 110 KB of SASS and its warps stalled on
            //  instruction fetch -- no_instruction 1.0-1.4 per issue, profiles/r02_*)
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {
                double2 y[3], u[3], y3[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) y[i] = u[i] = make_double2(0.0, 0.0);
                // y = A_j (-h/2 s + h^2/12 A d),  u = A_j d
                if constexpr (WC > 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int o0 = ((j * 2) * N + k3 + i) * WC;
#pragma unroll
                        for (int w = 0; w < WC; ++w) {
                            const double2 av = ellv[o0 + w];
                            const int col = xo + ellc[o0 + w];
                            const double2 sv = vS[col], xv = vX2[col];
                            cfma(y[i], av, make_double2(-c1h * sv.x + c2h2 * xv.x, -c1h * sv.y + c2h2 * xv.y));
                            cfma(u[i], av, vD[col]);
                        }
                    }
                } else {
                    const double2* Aj = cAj + j * NN;
#pragma unroll
                    for (int jj = 0; jj < N; ++jj) {
                        const double2 sv = vS[xo + jj], xv = vX2[xo + jj], dv = vD[xo + jj];
                        const double2 vv = make_double2(-c1h * sv.x + c2h2 * xv.x, -c1h * sv.y + c2h2 * xv.y);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double2 aij = Aj[(k3 + i) * N + jj];
                            cfma(y[i], aij, vv);
                            cfma(u[i], aij, dv);
                        }
                    }
                }
                __syncwarp();  // the previous drive's readers of vU are done
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vU[xo + k3 + i] = u[i];
                }
                __syncwarp();
                mv_reg(y3, vU + xo);
#pragma unroll
                double sp = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    put(imgJ, QO_TA + j, i, make_double2(y[i].x + c2h2 * y3[i].x, y[i].y + c2h2 * y3[i].y));
                    if (needH && act) sp += c2h * rdot(w1[i], u[i]);
                }
                if (needH) sah[j * 32] = sp;
            }
        }
        if (p.n_aux) do_aux_smem(p, auxs, t, lane, 32);  // derivative-integrator entries of this knot
        fence_async_smem();
        __syncwarp();
        if (BW || (p.spread && needJ)) flush_units_lanes(stage, segs, seghdr[12], seghdr[13], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & (QCK_EVAL_F | QCK_EVAL_J));
        else flush_units_lanes(stage, segs, seghdr[0], seghdr[3], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & (QCK_EVAL_F | QCK_EVAL_J));
        bulk_commit();
        __syncwarp();
        // ---- phase 2: Hessian-of-Lagrangian values, staged in the same buffer ------------------------------------------------------
        if (needH) {
            double2 w2[3];
            mvAH(w2, vW1 + xo);
            if (p.db) bulk_wait_read_1();  // the previous knot's Hessian image has left its buffer
            else bulk_wait_read();         // phase-1 copies have left the shared buffer
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                put(imgH, QO_KH0, i, make_double2(-0.5 * w1[i].x - c2h * w2[i].x, -0.5 * w1[i].y - c2h * w2[i].y));
                put(imgH, QO_KH1, i, make_double2(-0.5 * w1[i].x + c2h * w2[i].x, -0.5 * w1[i].y + c2h * w2[i].y));
            }
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {
                double sp = needJ ? sah[j * 32] : 0.0;
                double2 z1[3], z2[3], z3[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) z1[i] = z2[i] = make_double2(0.0, 0.0);
                // z1 = A_j^H m,  z2 = A_j^H w1
                if constexpr (WC > 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int o1 = ((j * 2 + 1) * N + k3 + i) * WC;  // (the table holds the rows of A_j^H)
#pragma unroll
                        for (int w = 0; w < WC; ++w) {
                            const double2 av = ellv[o1 + w];
                            const int col = xo + ellc[o1 + w];
                            cfma(z1[i], av, vM[col]);
                            cfma(z2[i], av, vW1[col]);
                        }
                    }
                } else {
                    const double2* Aj = cAj + j * NN;
#pragma unroll
                    for (int jj = 0; jj < N; ++jj) {
                        const double2 mv = vM[xo + jj], wv = vW1[xo + jj];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            double2 aji = Aj[jj * N + k3 + i];
                            aji.y = -aji.y;
                            cfma(z1[i], aji, mv);
                            cfma(z2[i], aji, wv);
                        }
                    }
                }
                if (!needJ) {  // Hessian-only call: the a_j x dt entry's Re <w1, A_j d> term (otherwise taken in phase 1)
                    double2 u[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) u[i] = make_double2(0.0, 0.0);
                    if constexpr (WC > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int o0 = ((j * 2) * N + k3 + i) * WC;
#pragma unroll
                            for (int w = 0; w < WC; ++w) cfma(u[i], ellv[o0 + w], vD[xo + ellc[o0 + w]]);
                        }
                    } else {
                        const double2* Aj2 = cAj + j * NN;
#pragma unroll
                        for (int jj = 0; jj < N; ++jj) {
                            const double2 dv = vD[xo + jj];
#pragma unroll
                            for (int i = 0; i < 3; ++i) cfma(u[i], Aj2[(k3 + i) * N + jj], dv);
                        }
                    }
                    if (act) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) sp += c2h * rdot(w1[i], u[i]);
                    }
                }
                __syncwarp();  // the previous drive's readers of vZ1 are done
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vZ1[xo + k3 + i] = z1[i];
                }
                __syncwarp();
                mvAH(z3, vZ1 + xo);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double cr = c2h2 * (z2[i].x + z3[i].x), ci = c2h2 * (z2[i].y + z3[i].y);
                    put(imgH, QO_KA0 + j, i, make_double2(-c1h * z1[i].x - cr, -c1h * z1[i].y - ci));
                    put(imgH, QO_KA1 + j, i, make_double2(-c1h * z1[i].x + cr, -c1h * z1[i].y + ci));
                    if (act) sp += -0.5 * rdot(z1[i], vS[xo + k3 + i]) + c2h * rdot(z1[i], vX2[xo + k3 + i]);
                }
                sp = warp_sum(sp);
                if (lane == 0 && c.pl_base[QO_HAH + j] >= 0) imgH[c.pl_base[QO_HAH + j]] = sp;
            }
            // a_i x a_j = h^2/12 Re tr({A_i, A_j} G),  G = D M^H (one more row-slice product, into the idle A_j D buffer);
            // the constant sparse anticommutators come as (row, column, value) lists, three lanes per pair
            double2 gr[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) gr[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int jj = 0; jj < N; ++jj) {
                double2 mv = vM[jj * N + cc];  // conj(M[cc][jj])
                mv.y = -mv.y;
#pragma unroll
                for (int i = 0; i < 3; ++i) cfma(gr[i], vD[jj * N + k3 + i], mv);
            }
            if (act) {
#pragma unroll
                for (int i = 0; i < 3; ++i) vX2[(k3 + i) * N + cc] = gr[i];  // G[row][column]  (vX2 = A D: its last readers were the drive loops)
            }
            __syncwarp();
            constexpr int NPAIR = ND * (ND + 1) / 2;
            const int pr = lane / 3, sub = lane - 3 * pr;
            double val = 0.0;
            if (pr < NPAIR)
                for (int u = kkptr[pr] + sub, u1 = kkptr[pr + 1]; u < u1; u += 3) {
                    const int rc = kkrc[u];
                    const double2 kv = kkv[u];
                    const double2 gv = vX2[(rc & 255) * N + (rc >> 8)];  // K[r, k] G[k, r]
                    val = fma(kv.x, gv.x, val);
                    val = fma(-kv.y, gv.y, val);
                }
            const double v1 = __shfl_down_sync(0xffffffffu, val, 1), v2 = __shfl_down_sync(0xffffffffu, val, 2);
            if (pr < NPAIR && sub == 0) {
                int j = 0, rem = pr;
                while (rem > j) { rem -= j + 1; ++j; }
                const int q = qo_haa(rem, j);
                if (c.pl_base[q] >= 0) imgH[c.pl_base[q]] = c2h2 * (val + v1 + v2);
            }
            s_hh = warp_sum(s_hh);
            if (lane == 0 && c.pl_base[QO_HHH] >= 0) imgH[c.pl_base[QO_HHH]] = s_hh * (1.0 / 6.0);
            fence_async_smem();
            __syncwarp();
            flush_units_lanes(stageH, segs, seghdr[4], seghdr[7], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & QCK_EVAL_H);
            bulk_commit();
            __syncwarp();
        }
    }
    if (QCK_BULK_STORE) bulk_wait_all();
}


}  // namespace

// one warp per knot, A rows in registers (9-level Pade-4 unitaries, one active member, up to four drives)
int qck_launch_rowslice9(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    static const int enabled = getenv("QCK_ROWSLICE") ? atoi(getenv("QCK_ROWSLICE")) : 1;
    if (!enabled || !c.rs3 || c.kind != QCK_UNITARY_PADE || c.order != 4 || c.N != 9 || L.member_end - L.member_begin != 1 || c.nd < 1 || c.nd > 4) return 0;
    typedef void (*kern_t)(const QckLaunch);
    static const int sparse_ok = getenv("QCK_ROWSLICE_DENSE") ? 0 : 1;
    const int wc = sparse_ok && c.W <= 2 ? c.W : 0;  // sparse drive rows of width 1 or 2 are unrolled; wider ones run dense
    kern_t kern;
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    // staging: F + J part and Hessian part of the output image share one buffer (the Hessian part starts at hoff)
    const int hoff = qck_rs3_hoff(c);
    const size_t jbytes = (size_t)((hoff + 4 + 1) & ~1) * 8, hbytes = (size_t)((c.img_doubles - hoff + 4 + 1) & ~1) * 8;
    const size_t shared = ((((size_t)(81 + c.ell_stride + c.kk_cap + c.ac_cap) * 16 + (size_t)c.icon_stride * 4) + 15) & ~(size_t)15) + (size_t)nrec * 16 +
                          (size_t)L.n_aux * sizeof(QckAux) + (wc > 0 ? 0 : (size_t)c.nd * 81 * 16);
    // two staging buffers per warp (6 warps fit) when the Hessian phase runs, one shared buffer (8 warps) otherwise
    // knobs (A/B measurements in profiles/): issue the kron block copies early; separate staging buffers for the two phases
    // (6 warps instead of 8: measured slower, off by default).  A write-out with plain 16-byte stores from registers instead of
    // TMA bulk copies was measured 10 % slower (profiles/r02_rowslice_variants.txt) and is not kept.
    static const int spread_knob = getenv("QCK_ROWSLICE_SPREAD") ? atoi(getenv("QCK_ROWSLICE_SPREAD")) : 1;
    static const int db_knob = getenv("QCK_ROWSLICE_DB") ? atoi(getenv("QCK_ROWSLICE_DB")) : 0;
    static const int knob = getenv("QCK_ROWSLICE_WARPS") ? atoi(getenv("QCK_ROWSLICE_WARPS")) : 0;
    int db = db_knob && (L.mask & QCK_EVAL_H) && (L.mask & (QCK_EVAL_F | QCK_EVAL_J)) ? 1 : 0;
    const size_t wextra = (size_t)c.nd * 32 * 8;
    size_t per_warp = (db ? jbytes + hbytes : std::max(jbytes, hbytes)) + 7 * 81 * 16 + wextra;
    if (db && shared + 4 * per_warp > 227 * 1024) { db = 0; per_warp = std::max(jbytes, hbytes) + 7 * 81 * 16 + wextra; }
    // block warps: only when the Jacobian is asked for, the blocks are contiguous in the structure, and the drives are sparse
    // (the dense-drive variants keep the plain kernel: fewer template instances)
    static const int bw_knob = getenv("QCK_ROWSLICE_BW") ? atoi(getenv("QCK_ROWSLICE_BW")) : 1;
    const size_t light_bytes = (size_t)4 * (2 * 328 * 8 + 2 * 81 * 8);  // block warps: two block buffers + A each
    const bool bw = bw_knob && wc > 0 && (L.mask & QCK_EVAL_J) && !db && c.pl_stride[QO_ISOF] == 1 && c.pl_stride[QO_ISOB] == 1 && c.pl_base[QO_ISOF] >= 0 &&
                    shared + 8 * per_warp + light_bytes <= 227 * 1024;
#define QCK_RS3(ND_, WC_, AH_) (bw ? qck_rowslice9_kernel<ND_, WC_, AH_, true> : qck_rowslice9_kernel<ND_, WC_, AH_, false>)
#define QCK_RS(ND_) (c.antiherm ? (wc == 1 ? QCK_RS3(ND_, 1, true) : (wc == 2 ? QCK_RS3(ND_, 2, true) : qck_rowslice9_kernel<ND_, 0, true, false>)) \
                                 : (wc == 1 ? QCK_RS3(ND_, 1, false) : (wc == 2 ? QCK_RS3(ND_, 2, false) : qck_rowslice9_kernel<ND_, 0, false, false>)))
    kern = c.nd == 1 ? QCK_RS(1) : (c.nd == 2 ? QCK_RS(2) : (c.nd == 3 ? QCK_RS(3) : QCK_RS(4)));
#undef QCK_RS
#undef QCK_RS3
    int nwarps = 8;
    if (knob >= 1 && knob <= 8 && !bw) nwarps = knob;
    const size_t light = bw ? light_bytes : 0;
    while (nwarps > 1 && shared + nwarps * per_warp + light > 227 * 1024) --nwarps;
    const size_t smem = shared + nwarps * per_warp + light;
    if (smem > 227 * 1024) return 0;
    {   // once per (device, kernel): the largest opt-in size covers every mask; masks with and without the Jacobian alternate
        // between two kernels (with / without block warps), so the per-class plan cache (one kernel) is not enough here
        static std::mutex mu;
        static std::set<std::pair<int, const void*>> ready;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (!ready.count({dev, (const void*)kern})) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
            ready.insert({dev, (const void*)kern});
        }
    }
    long long grid = sm_count;
    if (grid * nwarps > L.n_knots) grid = (L.n_knots + nwarps - 1) / nwarps;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] row-slice kernel: N=9 nd=%d warps/CTA=%d smem=%zu B grid=%lld units=%d two-buffers=%d early-blocks=%d block-warps=%d\n", c.nd, nwarps, smem, grid, c.nseg, db, spread_knob, (int)bw);
    QckLaunch L2 = L;
    L2.hoff = hoff;
    L2.db = db;
    L2.spread = spread_knob;
    kern<<<(unsigned)grid, bw ? 384 : nwarps * 32, smem, stream>>>(L2);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}

