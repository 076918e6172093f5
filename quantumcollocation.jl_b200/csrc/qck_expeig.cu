// Spectral kernel for the exponential integrators (9-level unitaries, Hermitian Hamiltonians, one warp per knot) of libqcknot.so
// (see DESIGN.md section 4).  Compiled as its own translation unit so that the kernel families build in parallel.
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// UnitaryExponentialIntegrator, residual R = U1 - exp(h A(a)) U0 with A = -i H(a), H Hermitian (the host checks the matrices).
//
// The scaling-and-squaring kernel (qck_kernels.cu, QK_EXP) carries n_d + 1 Frechet jets through ~14 dense products and a
// reverse sweep over a tape for the second derivatives: 1.5 MFLOP per 9-level knot.  Here ONE WARP diagonalises H = V L V^H
// (cyclic Jacobi with a round-robin ordering: the 4 disjoint rotations of a round are applied at once, each lane owns one 2x2
// block of H or one row pair of V; 4-6 sweeps) and every derivative of the exponential becomes a DIVIDED DIFFERENCE in the
// eigenbasis (Daleckii-Krein), with x_p = -i h l_p, W0 = V^H U0, Mt = V^H M, B_j = V^H A_j V:
//     exp(hA) = V diag(e^x) V^H                               D exp(hA)[h A_j] = V (h B_j o Phi) V^H,  Phi_pq = exp[x_p, x_q]
//     Phi_pq = e^{(x_p + x_q)/2} sinc(h (l_p - l_q) / 2)      (stable for every gap, degenerate levels included)
//     R = U1 - V (e^x . W0)        d/dh = -V ((-i l e^x) . W0)        d/da_j = -V ((h B_j o Phi) W0)
//     state x dt = -V (conj(-i l e^x) . Mt)                   state x a_j = -V ((h B_j o Phi)^H Mt)
//     with Gamma = W0 Mt^H:   dt x dt = Re sum_p l_p^2 e^{x_p} Gamma_pp
//                             a_j x dt = -Re sum_pq B_j,pq Gamma_qp (e^{x_q} - i h l_p Phi_pq)
//                             a_i x a_j = -h^2 Re sum_qr (B_j,qr K^i_qr + B_i,qr K^j_qr),  K^i_qr = sum_p Gamma_rp exp[x_p,x_q,x_r] B_i,pq
// The second-order divided differences exp[x_p, x_q, x_r] (165 distinct triples) come from the first-order ones across the
// LARGEST gap of the triple; a triple whose levels all lie within 0.4 of each other uses the Taylor series of the divided
// difference about their mean (complete homogeneous polynomials) instead -- no cancellation for any spectrum.
// About 30 product-sized steps of 9x9 complex work per knot instead of ~250; no tape, no limit on ||h A||.
// Lane (c, k) owns rows 3k .. 3k+2 of column c like the row-slice kernel (qck_rowslice.cu); products with V^H and V run on
// register-resident rows; values go through the warp's staging image in the solver's structure order (two phases) and leave
// as TMA bulk copies (flush_units_lanes).
// ------------------------------------------------------------------------------------------------------------
constexpr int EN = 9, ENN = 81, EN2 = 18, EDIM = 162, EF3 = 165;
constexpr int kEigBufs = 6;       // T1|Gamma, mV, vU0|T2, vM|Phi, vW0|T3, vMt  (+ ND buffers B_j)
constexpr int kEigScratch = 86;   // double2 per knot in the eigen scratch: V row-major (81) | l_0 .. l_8 (+ pad)

__device__ __forceinline__ size_t expeig_warp_bytes(int nd, int img_bytes) {
    return (size_t)img_bytes + (size_t)(kEigBufs + nd) * ENN * 16 + (size_t)EF3 * 16 + (size_t)EN * 16 * 2 + (size_t)EN * 8 + 8;
}

// (y1, y2) <- (y1, y2) J,  J = [[c, conj(sg)], [-sg, c]]  (column rotation of a row vector)
__device__ __forceinline__ void rot_right(double2& y1, double2& y2, double c, double2 sg) {
    const double2 a = y1, b = y2;
    y1 = make_double2(c * a.x - (sg.x * b.x - sg.y * b.y), c * a.y - (sg.x * b.y + sg.y * b.x));
    y2 = make_double2(c * b.x + (sg.x * a.x + sg.y * a.y), c * b.y + (sg.x * a.y - sg.y * a.x));
}
// (z1; z2) <- J^H (z1; z2)  (row rotation of a column vector)
__device__ __forceinline__ void rot_left(double2& z1, double2& z2, double c, double2 sg) {
    const double2 a = z1, b = z2;
    z1 = make_double2(c * a.x - (sg.x * b.x + sg.y * b.y), c * a.y - (sg.x * b.y - sg.y * b.x));
    z2 = make_double2(c * b.x + (sg.x * a.x - sg.y * a.y), c * b.y + (sg.x * a.y + sg.y * a.x));
}

// sorted triples (lo | mid << 4 | hi << 8) of the 165 second-order divided differences, index hi(hi+1)(hi+2)/6 + mid(mid+1)/2 + lo
__device__ const unsigned short kTriples[EF3] = {
    0x000, 0x100, 0x110, 0x111, 0x200, 0x210, 0x211, 0x220, 0x221, 0x222, 0x300, 0x310, 0x311, 0x320, 0x321,
    0x322, 0x330, 0x331, 0x332, 0x333, 0x400, 0x410, 0x411, 0x420, 0x421, 0x422, 0x430, 0x431, 0x432, 0x433,
    0x440, 0x441, 0x442, 0x443, 0x444, 0x500, 0x510, 0x511, 0x520, 0x521, 0x522, 0x530, 0x531, 0x532, 0x533,
    0x540, 0x541, 0x542, 0x543, 0x544, 0x550, 0x551, 0x552, 0x553, 0x554, 0x555, 0x600, 0x610, 0x611, 0x620,
    0x621, 0x622, 0x630, 0x631, 0x632, 0x633, 0x640, 0x641, 0x642, 0x643, 0x644, 0x650, 0x651, 0x652, 0x653,
    0x654, 0x655, 0x660, 0x661, 0x662, 0x663, 0x664, 0x665, 0x666, 0x700, 0x710, 0x711, 0x720, 0x721, 0x722,
    0x730, 0x731, 0x732, 0x733, 0x740, 0x741, 0x742, 0x743, 0x744, 0x750, 0x751, 0x752, 0x753, 0x754, 0x755,
    0x760, 0x761, 0x762, 0x763, 0x764, 0x765, 0x766, 0x770, 0x771, 0x772, 0x773, 0x774, 0x775, 0x776, 0x777,
    0x800, 0x810, 0x811, 0x820, 0x821, 0x822, 0x830, 0x831, 0x832, 0x833, 0x840, 0x841, 0x842, 0x843, 0x844,
    0x850, 0x851, 0x852, 0x853, 0x854, 0x855, 0x860, 0x861, 0x862, 0x863, 0x864, 0x865, 0x866, 0x870, 0x871,
    0x872, 0x873, 0x874, 0x875, 0x876, 0x877, 0x880, 0x881, 0x882, 0x883, 0x884, 0x885, 0x886, 0x887, 0x888,
};

__device__ __forceinline__ int f3_index(int a, int b, int c) {
    const int lo = min(a, min(b, c)), hi = max(a, max(b, c)), mid = a + b + c - lo - hi;
    return hi * (hi + 1) * (hi + 2) / 6 + mid * (mid + 1) / 2 + lo;
}

// ------------------------------------------------------------------------------------------------------------
// Eigen kernel: H(a_t) = V diag(l) V^H for every knot, one warp per knot, 32+ warps per SM (2.6 KB of shared memory and ~60
// registers per warp: the rotations' dependent latencies hide behind other knots, which the 7-warp main kernel cannot do).
// Cyclic Jacobi with the round-robin ordering: round s rotates the 4 disjoint pairs ((s + l) mod 9, (s - l) mod 9), l = 1 .. 4
// (level s sits out).  H and V are stacked into one 18 x 9 row-major matrix S = [H; V]; a round is
//   (1) every lane forms the rotation of pair l = lane >> 3 from the pivots,  J = [[c, conj(sg)], [-sg, c]]
//   (2) pass R:  S <- S J    72 tasks (row, pair) of two elements each, same code for the rows of H and of V
//   (3) pass L:  H <- J^H H  36 tasks (column, pair); the pivot entries are set to their exact values (0 / real)
// Results go to the class's scratch (V row-major, then the eigenvalues): 1.4 KB per knot, L2-resident for the main kernel.
// ------------------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256, 5) qck_eig9_kernel(const QckLaunch p) {
    constexpr int N = EN, NN = ENN;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)(blockDim.x >> 5);
    double2* const S = reinterpret_cast<double2*>(smem_all) + (size_t)warp * 2 * NN;
    const int m = p.member_begin;
    const double2* const gv = c.cmat + (size_t)m * c.cmat_stride;
    const int* const gc = c.ell_col + (size_t)m * c.icon_stride;
    const double2* const A0 = gv;
    const double2* const acv = gv + NN + c.ell_stride + c.kk_cap;
    const int* const acptr = gc + c.ell_stride + ND * (ND + 1) / 2 + 1 + c.kk_cap;
    const int* const acj = acptr + NN + 1;
    const int coff = p.moff_global[1];
    const int lA = lane >> 3, row0 = lane & 7;  // pair and row / column of this lane: a quarter-warp shares one pair (no bank conflicts)
    for (long long t = (long long)blockIdx.x * nwarps + warp; t < p.n_knots; t += (long long)gridDim.x * nwarps) {
        const double* zt = p.Z + t * c.zdim;
        double ctl[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) ctl[j] = zt[coff + j];
        // H = i A(a) = H_drift + sum_j a_j H_j (row-major), V = I
        double fro2 = 0.0;
        for (int e = lane; e < NN; e += 32) {
            double2 v = __ldg(A0 + e);
            for (int u = __ldg(acptr + e), u1 = __ldg(acptr + e + 1); u < u1; ++u) {
                const int jd = __ldg(acj + u);
                double aj = ctl[0];
#pragma unroll
                for (int j = 1; j < ND; ++j) aj = jd == j ? ctl[j] : aj;
                const double2 d = __ldg(acv + u);
                v.x = fma(aj, d.x, v.x);
                v.y = fma(aj, d.y, v.y);
            }
            const int r = e % N, col = e / N;  // A0 is column-major
            S[r * N + col] = make_double2(-v.y, r == col ? 0.0 : v.x);  // i (x + i y) = -y + i x; the diagonal of H is real
            S[NN + e] = make_double2(r == col ? 1.0 : 0.0, 0.0);
            fro2 = fma(v.x, v.x, fma(v.y, v.y, fro2));
        }
        fro2 = warp_sum(fro2);
        __syncwarp();
        int pl = lA + 1, ql = N - 1 - lA;  // pair l of round 0; both move up by one (mod 9) every round
        for (int sweep = 0; sweep < 12; ++sweep) {
            double off2 = 0.0;
            for (int e = lane; e < NN; e += 32) {
                const double2 v = S[e];
                if (e % (N + 1)) off2 = fma(v.x, v.x, fma(v.y, v.y, off2));
            }
            off2 = warp_sum(off2);
            if (off2 <= 1e-30 * fro2) break;
#pragma unroll 1
            for (int s = 0; s < N; ++s) {
                double cr = 1.0;
                double2 sg = make_double2(0.0, 0.0);
                {
                    const double al = S[pl * (N + 1)].x, ga = S[ql * (N + 1)].x;
                    const double2 be = S[pl * N + ql];
                    const double b2 = be.x * be.x + be.y * be.y;
                    if (b2 > 0.0) {  // inner rotation: c = cos >= 1/sqrt(2)
                        const double d = 0.5 * (ga - al);
                        const double inv_r = rsqrt(fma(d, d, b2));
                        const double u = fma(0.5 * fabs(d), inv_r, 0.5);
                        const double inv_c = rsqrt(u);
                        cr = u * inv_c;
                        const double f = copysign(0.5 * inv_r * inv_c, d);
                        sg = make_double2(f * be.x, -f * be.y);
                    }
                }
                __syncwarp();  // every lane has read its pivots
#pragma unroll
                for (int it = 0; it < 3; ++it) {
                    const int rho = row0 + 8 * it;
                    if (it < 2 || rho < 2 * N) {
                        double2 a = S[rho * N + pl], b = S[rho * N + ql];
                        rot_right(a, b, cr, sg);
                        S[rho * N + pl] = a; S[rho * N + ql] = b;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int gam = row0 + 8 * it;
                    if (gam < N) {
                        double2 a = S[pl * N + gam], b = S[ql * N + gam];
                        rot_left(a, b, cr, sg);
                        if (gam == pl) { a.y = 0.0; b = make_double2(0.0, 0.0); }
                        if (gam == ql) { b.y = 0.0; a = make_double2(0.0, 0.0); }
                        S[pl * N + gam] = a; S[ql * N + gam] = b;
                    }
                }
                __syncwarp();
                pl = pl + 1 == N ? 0 : pl + 1;
                ql = ql + 1 == N ? 0 : ql + 1;
            }
        }
        double2* const out = c.tape + (size_t)t * kEigScratch;
        for (int e = lane; e < NN; e += 32) out[e] = S[NN + e];
        if (lane < N) reinterpret_cast<double*>(out + NN)[lane] = S[lane * (N + 1)].x;
        __syncwarp();  // S is rebuilt for the warp's next knot
    }
}

template <int ND>
__global__ void __launch_bounds__(256, 1) qck_expeig9_kernel(const QckLaunch p) {
    constexpr int N = EN, NN = ENN, n2 = EN2, dim = EDIM;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)(blockDim.x >> 5);
    const bool act = lane < 3 * N;
    // (idle lanes 27..31 shadow lane 24: their reads then broadcast with column 8 instead of adding a bank conflict against column 0)
    const int cc = act ? lane / 3 : N - 1, k3 = act ? 3 * (lane - 3 * (lane / 3)) : 0;  // column, first row of this lane
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const bool needT = needJ || needH;
    const bool free_time = c.free_time;
    const int m = p.member_begin;
    const int W = c.W, elln = c.ell_stride, kkc = c.kk_cap, acn = c.ac_cap;
    double2* const conv = reinterpret_cast<double2*>(smem_all);
    const int nconv = NN + elln + kkc + acn;
    int* const coni = reinterpret_cast<int*>(conv + nconv);
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    QckSeg* const segtab = reinterpret_cast<QckSeg*>(smem_all + (((size_t)nconv * 16 + (size_t)c.icon_stride * 4 + 15) & ~(size_t)15));
    const int hoff = p.hoff;
    const int jbytes = ((hoff + 4 + 1) & ~1) * 8, hbytes = ((c.img_doubles - hoff + 4 + 1) & ~1) * 8;
    const int img_bytes = jbytes > hbytes ? jbytes : hbytes;
    QckAux* const auxs = reinterpret_cast<QckAux*>(segtab + nrec);
    unsigned char* const wbase = reinterpret_cast<unsigned char*>(auxs + p.n_aux) + (size_t)warp * expeig_warp_bytes(ND, img_bytes);
    double* const stage = reinterpret_cast<double*>(wbase);
    double2* const T1 = reinterpret_cast<double2*>(wbase + img_bytes);  // scratch matrix
    double2* const mV = T1 + NN;    // eigenvectors, row-major: V[r][k] = mV[r * N + k]
    double2* const vU0 = mV + NN;   // columns of U0: element [c * N + r]; later scratch
    double2* const vM = vU0 + NN;   // multipliers; later Phi[p * N + q]
    double2* const vW0 = vM + NN;   // V^H U0 (column-major)
    double2* const vMt = vW0 + NN;  // V^H M  (column-major)
    double2* const vB = vMt + NN;   // B_j = V^H A_j V, row-major, j = 0 .. ND-1
    double2* const f3 = vB + ND * NN;  // exp[x_a, x_b, x_c], a <= b <= c
    double2* const ex = f3 + EF3;      // e^{x_p}
    double2* const hx = ex + N;        // e^{x_p / 2}
    double* const lam = reinterpret_cast<double*>(hx + N);
    double2* const T2 = vU0;
    double2* const Phi = vM;
    double2* const vG = T1;   // Gamma = W0 Mt^H, row-major (phase 2: T1 is free)
    double2* const T3 = vW0;  // phase 2 scratch once Gamma is formed
    {
        const double2* gv = c.cmat + (size_t)m * c.cmat_stride;
        const int* gc = c.ell_col + (size_t)m * c.icon_stride;
        for (int e = threadIdx.x; e < nconv; e += blockDim.x) conv[e] = gv[e];
        for (int e = threadIdx.x; e < c.icon_stride; e += blockDim.x) coni[e] = gc[e];
        const QckSeg* gs = c.segs + (size_t)m * nrec;
        for (int i = threadIdx.x; i < nrec; i += blockDim.x) segtab[i] = gs[i];
        for (int i = lane; i < img_bytes / 8; i += 32) stage[i] = 0.0;
        for (int i = threadIdx.x; i < p.n_aux; i += blockDim.x) auxs[i] = p.aux[i];
        __syncthreads();
    }
    const double2* const A0 = conv;
    const double2* const ellv = conv + NN;
    const double2* const acv = ellv + elln + kkc;
    const int* const ellc = coni;
    const int* const acptr = coni + elln + ND * (ND + 1) / 2 + 1 + kkc;
    const int* const acj = acptr + NN + 1;
    const int* seghdr = reinterpret_cast<const int*>(segtab);
    const QckSeg* segs = segtab + QCK_SEG_HDR / 4;
    const int soff = p.moff_global[0], coff = p.moff_global[1], roff = p.moff_global[2];
    const int xo = cc * N;

    for (long long t = (long long)warp * gridDim.x + blockIdx.x; t < p.n_knots; t += (long long)gridDim.x * nwarps) {
        const double* zt = p.Z + t * c.zdim;
        constexpr int NLD = (dim + 31) / 32;
        double in0[NLD], inm[NLD];
        {
            const double* mut = p.mu + t * c.dyn + roff;
#pragma unroll
            for (int q = 0; q < NLD; ++q) {
                const int idx = lane + 32 * q;
                const bool ok = idx < dim;
                in0[q] = ok ? zt[soff + idx] : 0.0;
                inm[q] = ok && needH ? mut[idx] : 0.0;
            }
        }
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        double u1r[3], u1i[3];  // this lane's elements of U1
        {
            const double* z1 = zt + c.zdim + soff + cc * n2 + k3;
#pragma unroll
            for (int i = 0; i < 3; ++i) { u1r[i] = z1[i]; u1i[i] = z1[i + N]; }
        }
        double* const baseF = p.F + t * c.dyn;
        double* const baseJ = p.J + t * p.nnzJ;
        double* const baseH = p.H + t * p.nnzH;
        const int shF = (int)((reinterpret_cast<uintptr_t>(baseF) >> 3) & 1), shJ = (int)((reinterpret_cast<uintptr_t>(baseJ) >> 3) & 1),
                  shH = (int)((reinterpret_cast<uintptr_t>(baseH) >> 3) & 1);
        double* const imgF = stage + shF;
        double* const imgJ = stage + shJ;
        double* const imgH = stage - hoff + shH;
        if (t + (long long)gridDim.x * nwarps < p.n_knots) {  // pull the next knot of this warp into L2 meanwhile
            const double* zn = zt + (long long)gridDim.x * nwarps * c.zdim;
            const double* mn = p.mu + (t + (long long)gridDim.x * nwarps) * c.dyn + roff;
            for (int b = lane * 16; b < 2 * c.zdim; b += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(zn + b));
            if (needH)
                for (int b = lane * 16; b < dim; b += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(mn + b));
        }
        double2 vin[3];  // eigenvectors (row-major) of H(a_t) from the eigen kernel's scratch
        const double2* const scr = c.tape + (size_t)t * kEigScratch;
#pragma unroll
        for (int q = 0; q < 3; ++q) vin[q] = lane + 32 * q < NN ? scr[lane + 32 * q] : make_double2(0.0, 0.0);
        const double lam_in = lane < N ? reinterpret_cast<const double*>(scr + NN)[lane] : 0.0;
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int idx = lane + 32 * q;
            if (idx < dim) {
                const int col = idx / n2, qq = idx - col * n2, im = qq >= N, r = qq - im * N;
                const int o = 2 * (col * N + r) + im;
                reinterpret_cast<double*>(vU0)[o] = in0[q];
                if (needH) reinterpret_cast<double*>(vM)[o] = inm[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (lane + 32 * q < NN) mV[lane + 32 * q] = vin[q];
        // ---- spectrum: l_p, e^{x_p / 2}, e^{x_p} ---------------------------------------------------------------------------------------
        if (lane < N) {
            const double l = lam_in;
            lam[lane] = l;
            double sn, cs;
            sincos(0.5 * h * l, &sn, &cs);
            hx[lane] = make_double2(cs, -sn);
            ex[lane] = make_double2(cs * cs - sn * sn, -2.0 * sn * cs);
        }
        __syncwarp();
        // ---- products with V^H on register-resident rows: W0 = V^H U0, Mt = V^H M, B_j = V^H (A_j V) ---------------------------------
        double2 Vr[3][N];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double2 v = mV[k * N + k3 + i];  // V^H[k3 + i][k] = conj(V[k][k3 + i])
                Vr[i][k] = make_double2(v.x, -v.y);
            }
        auto mv_reg = [&](double2 (&y)[3], const double2* x) {
#pragma unroll
            for (int i = 0; i < 3; ++i) y[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double2 xv = x[k];
#pragma unroll
                for (int i = 0; i < 3; ++i) cfma(y[i], Vr[i][k], xv);
            }
        };
        auto put = [&](double* image, int q, int i, double2 v) {  // element (row k3 + i, column cc) of an iso-vector quantity
            const int b = c.pl_base[q], s = c.pl_stride[q], ire = cc * n2 + k3 + i;
            if (b >= 0 && act) { image[b + ire * s] = v.x; image[b + (ire + N) * s] = v.y; }
        };
        {
            double2 y[3];
            mv_reg(y, vU0 + xo);
            if (act) {
#pragma unroll
                for (int i = 0; i < 3; ++i) vW0[xo + k3 + i] = y[i];
            }
            if (needH) {
                mv_reg(y, vM + xo);
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vMt[xo + k3 + i] = y[i];
                }
            }
        }
        __syncwarp();  // U0 and M are consumed: their buffers become T2 and Phi
        if (needT) {
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {
                double2* const T = (j & 1) ? T2 : T1;
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // (A_j V)[k3 + i][cc] from the sparse rows of A_j
                        const int o0 = ((j * 2) * N + k3 + i) * W;
                        double2 u = make_double2(0.0, 0.0);
                        for (int w = 0; w < W; ++w) cfma(u, ellv[o0 + w], mV[ellc[o0 + w] * N + cc]);
                        T[xo + k3 + i] = u;
                    }
                }
                __syncwarp();
                double2 y[3];
                mv_reg(y, T + xo);
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vB[j * NN + (k3 + i) * N + cc] = y[i];
                }
            }
            // Phi_pq = e^{(x_p + x_q)/2} sinc(h (l_p - l_q) / 2)
            for (int e = lane; e < NN; e += 32) {
                const int pp = e / N, qq = e - pp * N;
                const double dl = 0.5 * h * (lam[pp] - lam[qq]);
                const double sc = dl == 0.0 ? 1.0 : sin(dl) / dl;
                const double2 g = cmul(hx[pp], hx[qq]);
                Phi[e] = make_double2(sc * g.x, sc * g.y);
            }
        }
        __syncwarp();
        // ---- products with V: register-resident rows of V ------------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < N; ++k) Vr[i][k] = mV[(k3 + i) * N + k];
        bulk_wait_read();  // the copy engine has finished reading the previous knot's image
        __syncwarp();
        {
            // E U0 = V (e^x . W0),  A E U0 = V ((-i l e^x) . W0),  E = V (e^x . V^H): one pass over k, three right-hand sides
            double2 yE[3], yT[3], yB[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) yE[i] = yT[i] = yB[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double2 e = ex[k];
                const double l = lam[k];
                const double2 w = cmul(e, vW0[xo + k]);
                const double2 wt = make_double2(l * w.y, -l * w.x);  // (-i l) w
                double2 vc = mV[cc * N + k];
                vc.y = -vc.y;
                const double2 eb = cmul(e, vc);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    cfma(yE[i], Vr[i][k], w);
                    if (needJ) { cfma(yT[i], Vr[i][k], wt); cfma(yB[i], Vr[i][k], eb); }
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                put(imgF, QO_R, i, make_double2(u1r[i] - yE[i].x, u1i[i] - yE[i].y));
                if (needJ) put(imgJ, QO_TH, i, make_double2(-yT[i].x, -yT[i].y));
            }
            if (needJ && act) {  // -iso(E), stored once, written N times (kron(I_N, .))
                const int bF = c.pl_base[QO_ISOF], sF = c.pl_stride[QO_ISOF];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int r = k3 + i;
                    const double fr = yB[i].x, fi = yB[i].y;
                    const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                    imgJ[bF + k00 * sF] = -fr; imgJ[bF + (k00 + N) * sF] = -fi; imgJ[bF + k01 * sF] = fi; imgJ[bF + (k01 + N) * sF] = -fr;
                }
            }
            if (needJ && lane == 0 && c.pl_base[QO_ONE] >= 0) imgJ[c.pl_base[QO_ONE]] = 1.0;  // identity block on state_{t+1}
        }
        if (needJ) {
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {  // d/da_j = -V (Lt_j W0),  Lt_j = h B_j o Phi
                const double2* const Bj = vB + j * NN;
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int o = (k3 + i) * N + cc;
                        const double2 l = cmul(Bj[o], Phi[o]);
                        T1[o] = make_double2(h * l.x, h * l.y);
                    }
                }
                __syncwarp();
                double2 acc[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) acc[i] = make_double2(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    const double2 w = vW0[xo + q];
#pragma unroll
                    for (int i = 0; i < 3; ++i) cfma(acc[i], T1[(k3 + i) * N + q], w);
                }
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) T2[xo + k3 + i] = acc[i];
                }
                __syncwarp();
                double2 y[3];
                mv_reg(y, T2 + xo);
#pragma unroll
                for (int i = 0; i < 3; ++i) put(imgJ, QO_TA + j, i, make_double2(-y[i].x, -y[i].y));
            }
        }
        if (p.n_aux) do_aux_smem(p, auxs, t, lane, 32);  // derivative-integrator entries of this knot
        fence_async_smem();
        __syncwarp();
        flush_units_lanes(stage, segs, seghdr[0], seghdr[3], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & (QCK_EVAL_F | QCK_EVAL_J));
        bulk_commit();
        __syncwarp();
        // ---- phase 2: Hessian-of-Lagrangian values, staged in the same buffer ------------------------------------------------------
        if (needH) {
            // second-order divided differences (needs Phi; independent of the staging buffer)
            for (int idx = lane; idx < EF3; idx += 32) {
                const int tri = kTriples[idx], lo = tri & 15, mid = (tri >> 4) & 15, hi = tri >> 8;
                const double t0 = h * lam[lo], t1 = h * lam[mid], t2 = h * lam[hi];
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
                } else {  // series about x_lo: e^{x_lo} sum_k (-i)^k h_k(0, d1, d2) / (k + 2)!,  h_k = complete homogeneous polynomial
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
            // Gamma = W0 Mt^H (row-major)
            {
                double2 g[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) g[i] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    double2 mv = vMt[k * N + cc];
                    mv.y = -mv.y;
#pragma unroll
                    for (int i = 0; i < 3; ++i) cfma(g[i], vW0[k * N + k3 + i], mv);
                }
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) vG[(k3 + i) * N + cc] = g[i];
                }
            }
            bulk_wait_read();  // phase-1 copies have left the staging buffer
            __syncwarp();
            if (free_time) {  // state x dt = -V (conj(-i l e^x) . Mt)
                double2 y[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) y[i] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double2 e = ex[k];
                    const double l = lam[k];
                    const double2 x = cmul(make_double2(l * e.y, l * e.x), vMt[xo + k]);
#pragma unroll
                    for (int i = 0; i < 3; ++i) cfma(y[i], Vr[i][k], x);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) put(imgH, QO_KH0, i, make_double2(-y[i].x, -y[i].y));
            }
#pragma unroll 1
            for (int j = 0; j < ND; ++j) {  // state x a_j = -V (Lt_j^H Mt),  Lt_j = h B_j o Phi
                const double2* const Bj = vB + j * NN;
                double sp = 0.0;
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        // element (p = k3 + i, q = cc); a_j x dt takes  B_j,pq Gamma_qp e^{x_q} - i l_p Lt_j,pq Gamma_qp  from here
                        const int pp = k3 + i, o = pp * N + cc;
                        const double2 bq = Bj[o];
                        const double2 l0 = cmul(bq, Phi[o]);
                        const double2 l = make_double2(h * l0.x, h * l0.y);
                        T2[o] = l;
                        const double2 g = vG[cc * N + pp], e = ex[cc];
                        const double lp = lam[pp];
                        const double2 fac = cmul(bq, e);  // B e^{x_q}
                        const double2 tot = make_double2(fac.x + lp * l.y, fac.y - lp * l.x);  // ... - i l_p Lt
                        sp -= tot.x * g.x - tot.y * g.y;
                    }
                }
                __syncwarp();
                double2 acc[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) acc[i] = make_double2(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    const double2 mv = vMt[xo + q];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        double2 l = T2[q * N + k3 + i];
                        l.y = -l.y;
                        cfma(acc[i], l, mv);
                    }
                }
                if (act) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) T3[xo + k3 + i] = acc[i];
                }
                __syncwarp();
                double2 y[3];
                mv_reg(y, T3 + xo);
#pragma unroll
                for (int i = 0; i < 3; ++i) put(imgH, QO_KA0 + j, i, make_double2(-y[i].x, -y[i].y));
                sp = warp_sum(sp);
                if (lane == 0 && c.pl_base[QO_HAH + j] >= 0) imgH[c.pl_base[QO_HAH + j]] = sp;
            }
            // a_i x a_j: lane (q = k3 + i, r = cc) forms K^d_qr = sum_p Gamma_rp exp[x_p, x_q, x_r] B_d,pq for every drive d
            {
                double2 K[ND][3];
#pragma unroll
                for (int d = 0; d < ND; ++d)
#pragma unroll
                    for (int i = 0; i < 3; ++i) K[d][i] = make_double2(0.0, 0.0);
#pragma unroll 1
                for (int pp = 0; pp < N; ++pp) {
                    const double2 g = vG[cc * N + pp];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double2 gf = cmul(g, f3[f3_index(pp, k3 + i, cc)]);
#pragma unroll
                        for (int d = 0; d < ND; ++d) cfma(K[d][i], gf, vB[d * NN + pp * N + k3 + i]);
                    }
                }
                double2 Bq[ND][3];
#pragma unroll
                for (int d = 0; d < ND; ++d)
#pragma unroll
                    for (int i = 0; i < 3; ++i) Bq[d][i] = vB[d * NN + (k3 + i) * N + cc];
                const double mh2 = -h * h;
#pragma unroll
                for (int j = 0; j < ND; ++j)
#pragma unroll
                    for (int i2 = 0; i2 <= j; ++i2) {
                        double v = 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            v += Bq[j][i].x * K[i2][i].x - Bq[j][i].y * K[i2][i].y;
                            v += Bq[i2][i].x * K[j][i].x - Bq[i2][i].y * K[j][i].y;
                        }
                        v = warp_sum(act ? v : 0.0);
                        const int q = qo_haa(i2, j);
                        if (lane == 0 && c.pl_base[q] >= 0) imgH[c.pl_base[q]] = mh2 * v;
                    }
            }
            if (free_time) {  // dt x dt = Re sum_p l_p^2 e^{x_p} Gamma_pp
                double v = 0.0;
                if (lane < N) {
                    const double l = lam[lane];
                    const double2 e = ex[lane], g = vG[lane * N + lane];
                    v = l * l * (e.x * g.x - e.y * g.y);
                }
                v = warp_sum(v);
                if (lane == 0 && c.pl_base[QO_HHH] >= 0) imgH[c.pl_base[QO_HHH]] = v;
            }
            fence_async_smem();
            __syncwarp();
            flush_units_lanes(stage, segs, seghdr[4], seghdr[7], lane, baseF, baseJ, baseH, shF, shJ, shH, hoff, p.mask & QCK_EVAL_H);
            bulk_commit();
            __syncwarp();
        }
    }
    bulk_wait_all();
}

}  // namespace

// one warp per knot, spectral form (9-level exponential unitaries with Hermitian Hamiltonians, one active member, <= 4 drives);
// the class must have been built with the parity-matched image placement and the per-phase unit table (c.eig)
int qck_launch_expeig(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    if (!c.eig) return 0;
    if (c.kind != QCK_UNITARY_EXP || c.N != 9 || !c.antiherm || L.member_end - L.member_begin != 1 || c.nd < 1 || c.nd > 4) return (int)cudaErrorInvalidConfiguration;
    typedef void (*kern_t)(const QckLaunch);
    const kern_t kern = c.nd == 1 ? qck_expeig9_kernel<1> : (c.nd == 2 ? qck_expeig9_kernel<2> : (c.nd == 3 ? qck_expeig9_kernel<3> : qck_expeig9_kernel<4>));
    const kern_t eig = c.nd == 1 ? qck_eig9_kernel<1> : (c.nd == 2 ? qck_eig9_kernel<2> : (c.nd == 3 ? qck_eig9_kernel<3> : qck_eig9_kernel<4>));
    if (!c.tape || (long long)c.max_ctas < L.n_knots) return (int)cudaErrorInvalidConfiguration;  // eigen scratch: one record per knot
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    const int hoff = qck_rs3_hoff(c);
    const size_t jbytes = (size_t)((hoff + 4 + 1) & ~1) * 8, hbytes = (size_t)((c.img_doubles - hoff + 4 + 1) & ~1) * 8;
    const size_t shared = ((((size_t)(81 + c.ell_stride + c.kk_cap + c.ac_cap) * 16 + (size_t)c.icon_stride * 4) + 15) & ~(size_t)15) + (size_t)nrec * 16 +
                          (size_t)L.n_aux * sizeof(QckAux);
    const size_t per_warp = std::max(jbytes, hbytes) + (size_t)(kEigBufs + c.nd) * ENN * 16 + (size_t)EF3 * 16 + (size_t)EN * 16 * 2 + (size_t)EN * 8 + 8;
    static const int knob = getenv("QCK_EXPEIG_WARPS") ? atoi(getenv("QCK_EXPEIG_WARPS")) : 0;
    int nwarps = knob >= 1 && knob <= 8 ? knob : 8;
    while (nwarps > 1 && shared + nwarps * per_warp > 227 * 1024) --nwarps;
    const size_t smem = shared + nwarps * per_warp;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
    {
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
    if (dbg) fprintf(stderr, "[qcknot] spectral exponential kernel: N=9 nd=%d warps/CTA=%d smem=%zu B grid=%lld units=%d\n", c.nd, nwarps, smem, grid, c.nseg);
    QckLaunch L2 = L;
    L2.hoff = hoff;
    {   // eigen-decompositions of all knots first: 8 warps per CTA, 2.6 KB of shared memory per warp
        long long egrid = (L.n_knots + 7) / 8;
        if (egrid > (long long)sm_count * 8) egrid = (long long)sm_count * 8;
        eig<<<(unsigned)egrid, 256, 8 * 2 * ENN * 16, stream>>>(L2);
        if (launches) ++*launches;
    }
    kern<<<(unsigned)grid, nwarps * 32, smem, stream>>>(L2);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}
