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
//  * every structural nonzero of the knot block belongs to a precomputed output segment (first position, length,
//    table of scratch slots + signs, repeat period) held in shared memory, so the "scatter" is a gather from
//    shared memory followed by position-ordered, coalesced 8-byte stores with no global loads in the loop;
//  * Hessian entries that several integrators contribute to (shared controls) go to a partial buffer that a
//    second kernel reduces in fixed integrator order (bitwise run-to-run reproducible, no atomics).
#include <cstdio>
#include <cstdlib>

#include "qck_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// Pade-4 integrators (UnitaryPadeIntegrator / QuantumStatePadeIntegrator, order 4).
//   F = I + h/2 A + h^2/12 A^2,  B = I - h/2 A + h^2/12 A^2,  residual R = B U1 - F U0,  A = -i H(a), h = dt.
// With D = U1-U0, S = U1+U0, M = multipliers as a complex matrix (mu^T vec_iso(R) = Re <M, R>), C_j = A_j A + A A_j,
// G = D M^H, G2 = S M^H (so that Re <M, K D> = Re tr(K G) for any N x N matrix K):
//   R       = D - h/2 A S + h^2/12 A^2 D                     dR/dh = -1/2 A S + h/6 A^2 D
//   dR/da_j = -h/2 A_j S + h^2/12 C_j D
//   state_t x a_j:   -(h/2 A_j^H M + h^2/12 C_j^H M)         a_j x state_t+1:  -h/2 A_j^H M + h^2/12 C_j^H M
//   state_t x dt:    -(1/2 A^H M + h/6 (A^2)^H M)            dt x state_t+1:   -1/2 A^H M + h/6 (A^2)^H M
//   a_i x a_j = h^2/12 Re tr({A_i, A_j} G)     a_j x dt = -1/2 Re tr(A_j G2) + h/6 Re tr(C_j G)     dt x dt = 1/6 Re tr(A^2 G)
// Dense N^3-type products per knot: stage 1: A^2, A S, A^H M, D M^H, S M^H;  stage 2: A^2 D, (A^2)^H M, C_j D, C_j^H M
// (5 + 2 + 2 n_d; the reference's real-iso formulation needs ~45 of twice the size).  {A_i, A_j} is a constant sparse
// matrix (host-precomputed), A_j is sparse (ELL rows): every product with them is O(nnz).
//
// Every product's epilogue writes final values straight into the CTA's OUTPUT IMAGE, a shared-memory buffer laid out
// in the solver's structure order (host-computed placement), so the write-out is a contiguous smem -> HBM copy.
// Shared memory per CTA: [matrices | image | per-member segments + constants (x2 if several members) | input staging | aux].
// The inputs (and tables, when the member changes) of the NEXT work item are prefetched with cp.async during compute.
// ------------------------------------------------------------------------------------------------------------
// optional per-stage cycle accounting (QCK_DEBUG_TIMING=1): thread 0 of every CTA accumulates clock64() deltas
#define QCK_TICK(k)                                                                  \
    if (p.timing && tid == 0) {                                                      \
        long long now_ = clock64();                                                  \
        atomicAdd((unsigned long long*)p.timing + (k), (unsigned long long)(now_ - tick_)); \
        tick_ = now_;                                                                \
    }

template <int KIND, int TC, int CN, bool MULTI, bool DMMA = false>
__global__ void __launch_bounds__(CN == 0 ? 256 : (KIND != QK_PADE4 ? 160 : 128), CN == 0 ? 1 : (KIND != QK_PADE4 ? 3 : 4))
qck_quantum_kernel(const QckLaunch p) {
    // A CTA holds `ngroups` independent groups of G threads; each group works through its own sequence of work items in
    // its own slice of shared memory.  Large problems use one group per CTA (block barrier); for small level counts a
    // group is a single warp (warp barrier only), so several items per CTA progress independently of one another.
    extern __shared__ __align__(16) unsigned char smem_all[];
    // FP64 tensor-core (DMMA) variant of the Pade-4 unitary path for 9 levels (see the DM block below)
    constexpr bool DM = DMMA && KIND == QK_PADE4 && TC == QCK_TILE && CN == 9 && !MULTI;
    constexpr int QDM_LD = 2 * (CN + 1);
    // (MULTI is a template parameter so that the single-group kernels keep absolute shared-memory addressing)
    const int G = MULTI ? 32 : (int)blockDim.x, grp = MULTI ? (int)(threadIdx.x >> 5) : 0, ngroups = MULTI ? (int)(blockDim.x >> 5) : 1;
    unsigned char* smem_raw = MULTI ? smem_all + (size_t)grp * p.group_smem : smem_all;
    const long long gid = (long long)blockIdx.x * ngroups + grp, gstride = (long long)gridDim.x * ngroups;
#define GSYNC()                     \
    do {                            \
        if (MULTI) __syncwarp();    \
        else __syncthreads();       \
    } while (0)
    double* sm = reinterpret_cast<double*>(smem_raw);
    const QckClassDev& c = p.c;
    const int tid = threadIdx.x - grp * G, nthreads = G;
    // CN > 0: levels known at compile time (index arithmetic folds, k-loops unroll); CN == 0: generic fallback
    const int N = CN > 0 ? CN : c.N;
    const int NP = CN > 0 ? ((CN + QCK_TILE - 1) / QCK_TILE) * QCK_TILE : c.NP;
    const int nc = TC == 1 ? 1 : N, ncp = TC == 1 ? 1 : NP;
    const int nd = c.nd, W = c.W;
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H, needF = p.mask & QCK_EVAL_F;
    const bool needT = needJ || needH;  // drive terms
    const bool free_time = c.free_time;
    const int dim = 2 * N * nc, n2 = 2 * N;
    const int npair = nd * (nd + 1) / 2;

    double2* SA = reinterpret_cast<double2*>(sm + c.off_A);
    double2* SS = reinterpret_cast<double2*>(sm + c.off_S);
    double* image = sm + c.off_img;
    double* stage = reinterpret_cast<double*>(smem_raw + c.sm_stage);  // [z_t state | z_t+1 state | mu | a | h]
    // per group: operands of the auxiliary entries; per CTA (behind all groups): the entries themselves and the member offsets
    double* auxv = reinterpret_cast<double*>(smem_raw + c.sm_bytes);
    unsigned char* cta_shared = smem_all + (size_t)ngroups * p.group_smem;
    QckAux* auxs = reinterpret_cast<QckAux*>(cta_shared);
    const int nact = p.member_end - p.member_begin;
    // per-member offsets (state, drive, row) of the active members: in shared memory when they fit, else global
    const int* moff = p.moff_smem ? reinterpret_cast<const int*>(cta_shared + p.n_aux * (int)sizeof(QckAux)) : p.moff_global;
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;  // 16-byte records of the per-member write-out table
    const int lane = tid & 31, warp = tid >> 5, nwarps_ = nthreads >> 5;
    const int elln = c.ell_stride, kkc = c.kk_cap, acc_n = c.ac_cap;
    const int msa = NP * NP, mss = NP * ncp;  // complex elements per matrix
#define MA(i) (SA + (i) * msa)
#define MS(i) (SS + (i) * mss)
#define SEGBUF(b) reinterpret_cast<QckSeg*>(smem_raw + c.sm_seg + (b) * c.seg_bytes)
#define CONV(b) reinterpret_cast<double2*>(smem_raw + c.sm_con + (b) * c.con_bytes)
#define CONI(b) reinterpret_cast<int*>(smem_raw + c.sm_con + (b) * c.con_bytes + (N * N + elln + kkc + acc_n) * 16)
#define PUT(q, i, v) image[c.pl_base[q] + (i) * c.pl_stride[q]] = (v)

    const long long n_items = p.n_knots * nact;
    long long tick_ = clock64();
    const int tilesS = (NP / QCK_TILE) * (ncp / TC);
    const int tilesA = (NP / QCK_TILE) * (NP / QCK_TILE);
    const int tcols = ncp / TC, tcolsA = NP / QCK_TILE;

    // prefetch of one work item: inputs into `stage`, and the member's tables into buffer b if asked
    auto prefetch = [&](long long t, int m, int b, bool tables) {
        const double* zt = p.Z + t * c.zdim;
        const int soff = moff[3 * (m - p.member_begin)], coff = moff[3 * (m - p.member_begin) + 1], roff_n = moff[3 * (m - p.member_begin) + 2];
        {
            const double* g0 = zt + soff + tid;
            const double* g1 = g0 + c.zdim;
            const double* g2 = p.mu + t * c.dyn + roff_n + tid;
            double* d0 = stage + tid;
            for (int i = tid; i < dim; i += nthreads, g0 += nthreads, g1 += nthreads, g2 += nthreads, d0 += nthreads) {
                cp_async8(d0, g0);
                cp_async8(d0 + dim, g1);
                if (needH) cp_async8(d0 + 2 * dim, g2);
            }
        }
        if (tid < nd) cp_async8(stage + 3 * dim + tid, zt + coff + tid);
        if (tid == nd && free_time) cp_async8(stage + 3 * dim + nd, zt + c.dt_off);
        if (p.n_aux && m == p.member_begin) {  // operands of the derivative-integrator entries this item also writes
            for (int k = tid; k < p.n_aux; k += nthreads) {
                const QckAux a = auxs[k];
                if (a.op == QAUX_NEG_Z || a.op == QAUX_FROW) cp_async8(auxv + 3 * k, zt + a.i0);
                if (a.op == QAUX_FROW) { cp_async8(auxv + 3 * k + 1, zt + c.zdim + a.i0); cp_async8(auxv + 3 * k + 2, zt + a.i1); }
                if (a.op == QAUX_NEG_MU && needH) cp_async8(auxv + 3 * k + 2, p.mu + t * c.dyn + a.i0);
            }
        }
        if (tables) {
            const QckSeg* gs = c.segs + (size_t)m * nrec;
            for (int i = tid; i < nrec; i += nthreads) cp_async16(SEGBUF(b) + i, gs + i);
            const double2* gv = c.cmat + (size_t)m * c.cmat_stride;  // [A0 | ell_val | kk_val]
            const int* gc = c.ell_col + (size_t)m * c.icon_stride;
            for (int i = tid; i < N * N + elln + kkc + acc_n; i += nthreads) cp_async16(CONV(b) + i, gv + i);
            for (int i = tid; i < c.icon_stride; i += nthreads) cp_async4(CONI(b) + i, gc + i);
        }
        cp_async_commit();
    };

    if (p.stagger_ns) {  // de-phase the CTAs that share an SM (they run identical sequences and would otherwise stay in lockstep)
        const unsigned k = blockIdx.x / (unsigned)p.sm_count;
        for (unsigned i = 0; i < k; ++i) __nanosleep(p.stagger_ns);
    }
    for (int i = tid; i < c.scratch_doubles; i += nthreads) sm[i] = 0.0;
    for (int i = threadIdx.x; i < p.n_aux; i += blockDim.x) auxs[i] = p.aux[i];
    if (p.moff_smem)
        for (int i = threadIdx.x; i < 3 * nact; i += blockDim.x) const_cast<int*>(moff)[i] = p.moff_global[i];
    __syncthreads();  // CTA-wide tables ready (every thread of the CTA gets here)
    if (tid == 0 && c.pl_base[QO_ONE] >= 0) image[c.pl_base[QO_ONE]] = 1.0;
    int buf = 0, buf_member = -1;
    if (gid < n_items) {
        buf_member = p.member_begin + (int)(gid % nact);
        prefetch(gid / nact, buf_member, buf, true);
    }

    // (t, mi) of the current item are advanced incrementally: a 64-bit division per item is not free
    const int step_t = (int)(gstride / nact), step_m = (int)(gstride % nact);
    long long t = gid / nact;
    int mi = (int)(gid % nact);
    for (long long item = gid; item < n_items; item += gstride) {
        const int vwarp = warp, vtid = tid;
        const int m = p.member_begin + mi;
        QCK_TICK(0);
        cp_async_wait_all();
        GSYNC();  // staged inputs visible; previous item's write-out has finished reading the image
        QCK_TICK(1);
        const double2* A0 = CONV(buf);
        const double2* ellv = CONV(buf) + N * N;
        const double2* kkv = ellv + elln;
        const int* ellc = CONI(buf);
        const int* kkptr = CONI(buf) + elln;
        const int* kkrc = kkptr + npair + 1;
        const double2* acv = kkv + kkc;        // per-element contributors of A = A0 + sum_j a_j A_j
        const int* acptr = kkrc + kkc;
        const int* acj = acptr + N * N + 1;
        const int* seghdr = reinterpret_cast<const int*>(SEGBUF(buf));
        const QckSeg* segs = SEGBUF(buf) + QCK_SEG_HDR / 4;
        const double h = free_time ? stage[3 * dim + nd] : c.dt_fixed;
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);

        // ---- stage 0: unpack the staged knot pair into D, S, M; build A = -i H(a) -----------------------------
        // (the derivative-integrator entries depend on the staged inputs only: written now, before the next prefetch
        //  reuses the staging buffers)
        if (mi == 0 && p.n_aux) do_aux_staged(p, auxs, auxv, h, t, tid, nthreads);
        if constexpr (DM) {
            // tensor-core path: every matrix is row-major complex with QDM_LD doubles per row (fragment loads are then free
            // of bank conflicts); [D; S] and [A2; C_1 .. C_nd] are stacked so that they form ONE left operand each
            double* const mA = sm;
            double* const mD = sm + (1 + (1 + nd) + 2) * N * QDM_LD;
            double* const mM = mD + 2 * N * QDM_LD;
            for (int e = tid; e < N * N; e += nthreads) {
                const int n = e % N, r = e / N;  // column index fastest: conflict-free 16-byte stores
                const double u0r = stage[n * n2 + r], u0i = stage[n * n2 + N + r];
                const double u1r = stage[dim + n * n2 + r], u1i = stage[dim + n * n2 + N + r];
                *reinterpret_cast<double2*>(mD + r * QDM_LD + 2 * n) = make_double2(u1r - u0r, u1i - u0i);
                *reinterpret_cast<double2*>(mD + (N + r) * QDM_LD + 2 * n) = make_double2(u1r + u0r, u1i + u0i);
                if (needH) *reinterpret_cast<double2*>(mM + r * QDM_LD + 2 * n) = make_double2(stage[2 * dim + n * n2 + r], stage[2 * dim + n * n2 + N + r]);
            }
            for (int e = nthreads - 1 - tid; e < N * N; e += nthreads) {
                double2 v = A0[e];
                for (int u = acptr[e]; u < acptr[e + 1]; ++u) {
                    const double aj = stage[3 * dim + acj[u]];
                    const double2 d = acv[u];
                    v.x = fma(aj, d.x, v.x);
                    v.y = fma(aj, d.y, v.y);
                }
                *reinterpret_cast<double2*>(mA + (e % N) * QDM_LD + 2 * (e / N)) = v;
            }
        } else {
            for (int idx = tid; idx < dim; idx += nthreads) {  // the last warp joins once A is built
                int cc = idx / n2, q = idx - cc * n2;
                int im = q >= N, r = q - im * N;
                double u0 = stage[idx], u1 = stage[dim + idx];
                int o = 2 * (r + NP * cc) + im;
                reinterpret_cast<double*>(MS(QS_D))[o] = KIND != QK_PADE4 ? u0 : u1 - u0;  // exp / general Pade keep U0, U1 themselves
                reinterpret_cast<double*>(MS(QS_S))[o] = KIND != QK_PADE4 ? u1 : u1 + u0;
                if (needH) reinterpret_cast<double*>(MS(QS_M))[o] = stage[2 * dim + idx];
            }
            // A = A0 + sum_j a_j A_j: every element gathers its own (host-listed) drive contributions, no ordering constraints
            for (int e = nthreads - 1 - tid; e < N * N; e += nthreads) {
                double2 v = A0[e];
                for (int u = acptr[e]; u < acptr[e + 1]; ++u) {
                    const double aj = stage[3 * dim + acj[u]];
                    const double2 d = acv[u];
                    v.x = fma(aj, d.x, v.x);
                    v.y = fma(aj, d.y, v.y);
                }
                MA(QA_A)[(e % N) + NP * (e / N)] = v;
            }
        }
        if (QCK_BULK_STORE && lane == 0) bulk_wait_read();  // the copy engine has finished reading the previous item's image
        GSYNC();
        QCK_TICK(2);
        // staging is free again: fetch the next item's inputs (and tables, if its member differs) behind the compute
        int next_buf = buf, next_member = buf_member;
        {
            const long long nitem = item + gstride;
            if (nitem < n_items) {
                int nmi = mi + step_m;
                long long nt = t + step_t;
                if (nmi >= nact) { nmi -= nact; ++nt; }
                const int nm = p.member_begin + nmi;
                const bool tables = nm != buf_member;
                if (tables) { next_buf = buf ^ 1; next_member = nm; }  // several active members => two table buffers
                prefetch(nt, nm, next_buf, tables);
            }
        }

        QCK_TICK(3);
        if constexpr (DM) {
        // ============================ Pade-4, unitaries, FP64 tensor cores (DMMA m8n8k4) ====================================
        // Complex products run as REAL tile products on the tensor cores.  For Z = X Y (X: m x N, Y: N x n complex):
        //     [Zr Zi] = [Xr Xi] [[Yr Yi]; [-Yi Yr]]     with interleaved real indices  k' = 2c + (im),  n' = 2n + (im),
        // so the left operand IS the row-major complex storage of X, the right operand is read from the complex storage of Y
        // with a per-lane sign, and lane (g, t) of an 8 x 8 accumulator tile holds the complete complex Z[8i + g][4j + t].
        // Conjugate-transposed left operands (X^H Y) read the storage of X transposed and fold the conjugation into the sign
        // pattern of the right operand; Y = M^H is read from the storage of M likewise.  K = 2N = 18 is padded to 20: the
        // two padding k' of the last step are masked to zero in both fragments.
        //   stage 1:  warp 0: A [A | S] -> A2 (stored), AS (kept in registers for stage 2)
        //             warp 1: [D; S] M^H -> G, G2              warp 2: A^H M -> AhM (registers)
        //             warps 3 (+2, or 1..3 without a Hessian): C_j = A_j A + A A_j (sparse)
        //   stage 2:  warps 0, 1: [A2; C_j] D   -> R, d/dh, d/da_j         (rows split between the two warps)
        //             warps 2, 3: [A2; C_j]^H M -> state x dt, state x a_j Hessian blocks
        //             then the scalar traces and the -iso(F) / iso(B) blocks on all warps
        constexpr int LD = QDM_LD, KS = (2 * CN + 3) / 4, NT = (2 * CN + 7) / 8, MTW = 3;
        double* const mA = sm;
        double* const mX = sm + N * LD;
        const int mrows = (1 + nd) * N;
        double* const mG = mX + mrows * LD;
        double* const mD = mG + 2 * N * LD;
        double* const mS = mD + N * LD;
        double* const mM = mD + 2 * N * LD;
        const int g = lane >> 2, tq = lane & 3;
        const int aoff = g * LD + tq;                                       // left operand, row-major storage
        const int atoff = (tq >> 1) * LD + 2 * g + (tq & 1);               // left operand = transpose of the storage
        const int boff = (tq >> 1) * LD + 2 * (g >> 1) + ((g ^ tq) & 1);   // right operand from Y[c][n] storage
        const int btoff = (g >> 1) * LD + 2 * (tq >> 1) + ((g ^ tq) & 1);  // right operand from storage indexed [n][c]
        const double sN = ((tq & 1) && !(g & 1)) ? -1.0 : 1.0;             // X Y
        const double sC = ((tq & 1) && (g & 1)) ? -1.0 : 1.0;              // X^H Y (left read transposed, unconjugated)
        const double sH = (!(tq & 1) && (g & 1)) ? -1.0 : 1.0;             // X M^H (right read from M[n][c])
        const bool kpad = 2 * CN - 4 * (KS - 1) <= tq;                     // this lane's k' of the last step is padding
        double keep[2][NT][2];  // AS (warp 0) / AhM (warp 2): rows 8i + g, columns 4j + t, for the stage-2 epilogues
        {
            if (warp == 0) {
                double acc[2][2 * NT][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2 * NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const bool z = ks == KS - 1 && kpad;
                    double a[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) a[i] = z ? 0.0 : mA[i * 8 * LD + aoff + 4 * ks];
#pragma unroll
                    for (int j = 0; j < 2 * NT; ++j) {
                        const double* Y = j < NT ? mA : mS;
                        const double b = z ? 0.0 : sN * Y[2 * ks * LD + 8 * (j % NT) + boff];
#pragma unroll
                        for (int i = 0; i < 2; ++i) dmma884(acc[i][j], a[i], b);
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int row = 8 * i + g, col = 4 * j + tq;
                        if (row < N && col < N) *reinterpret_cast<double2*>(mX + row * LD + 2 * col) = make_double2(acc[i][j][0], acc[i][j][1]);
                        keep[i][j][0] = acc[i][NT + j][0];
                        keep[i][j][1] = acc[i][NT + j][1];
                    }
            } else if (warp == 1 && needH) {
                double acc[3][NT][2];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const bool z = ks == KS - 1 && kpad;
                    double a[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) a[i] = z ? 0.0 : mD[i * 8 * LD + aoff + 4 * ks];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double b = z ? 0.0 : sH * mM[4 * j * LD + 4 * ks + btoff];
#pragma unroll
                        for (int i = 0; i < 3; ++i) dmma884(acc[i][j], a[i], b);
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int row = 8 * i + g, col = 4 * j + tq;
                        if (row < 2 * N && col < N) *reinterpret_cast<double2*>(mG + row * LD + 2 * col) = make_double2(acc[i][j][0], acc[i][j][1]);
                    }
            } else if (warp == 2 && needH) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) keep[i][j][0] = keep[i][j][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const bool z = ks == KS - 1 && kpad;
                    double a[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) a[i] = z ? 0.0 : mA[2 * ks * LD + 16 * i + atoff];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double b = z ? 0.0 : sC * mM[2 * ks * LD + 8 * j + boff];
#pragma unroll
                        for (int i = 0; i < 2; ++i) dmma884(keep[i][j], a[i], b);
                    }
                }
            }
            if (needT && warp >= (needH ? 2 : 1)) {
                // C_j[r][c] = sum_u A_j[r, k_u] A[k_u, c] + sum_u A[r, k_u] A_j[k_u, c]; the second sum walks row c of A_j^H.
                // 32-element chunks are dealt to the warps the dense products leave free (two of three to warp 3 with a Hessian).
                const int total = nd * N * N, nch = (total + 31) >> 5;
                for (int ch = 0; ch < nch; ++ch) {
                    const int owner = needH ? (ch % 3 == 2 ? 2 : 3) : 1 + ch % 3;
                    if (owner != warp) continue;
                    const int w = ch * 32 + lane;
                    if (w >= total) continue;
                    const int j = w / (N * N), e = w - j * N * N;
                    const int r = e / N, cc = e - r * N;
                    const int o0 = ((j * 2) * N + r) * W, o1 = ((j * 2 + 1) * N + cc) * W;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int u = 0; u < W; ++u) {
                        cfma(acc, ellv[o0 + u], *reinterpret_cast<const double2*>(mA + ellc[o0 + u] * LD + 2 * cc));
                        double2 ah = ellv[o1 + u];
                        ah.y = -ah.y;
                        cfma(acc, *reinterpret_cast<const double2*>(mA + r * LD + 2 * ellc[o1 + u]), ah);
                    }
                    *reinterpret_cast<double2*>(mX + (N + j * N + r) * LD + 2 * cc) = acc;
                }
            }
        }
        GSYNC();
        QCK_TICK(4);
        {
            const int mtiles = (mrows + 7) >> 3;
            int mt_half = (mtiles + 1) >> 1;
            if (mt_half < 2) mt_half = 2;  // rows 0 .. N-1 (they pair with `keep`) stay on warps 0 / 2
            const bool is_b = warp >= 2;
            const int mt_begin = (warp & 1) * mt_half;
            const int mt_end = mt_begin + mt_half < mtiles ? mt_begin + mt_half : mtiles;
            if (is_b ? needH : true)
            for (int mt0 = mt_begin; mt0 < mt_end; mt0 += MTW) {  // (one pass for up to four drives)
                const int mtn = mt_end - mt0 < MTW ? mt_end - mt0 : MTW;
                double acc[MTW][NT][2];
                int prow[MTW];  // (block p, row r) of this lane's row of m-tile i, packed p * 16 + r
                int abase[MTW];
#pragma unroll
                for (int i = 0; i < MTW; ++i) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                    const int m = (mt0 + i) * 8 + g, pb = m / N, r = m - pb * N;
                    prow[i] = pb * 16 + r;
                    abase[i] = is_b ? pb * N * LD + 2 * r + (tq >> 1) * LD + (tq & 1) : m * LD + tq;
                }
                const int astep = is_b ? 2 * LD : 4;
                const double* const Bm = is_b ? mM : mD;
                const double sB = is_b ? sC : sN;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const bool z = ks == KS - 1 && kpad;
                    double a[MTW];
#pragma unroll
                    for (int i = 0; i < MTW; ++i)
                        if (i < mtn) a[i] = z ? 0.0 : mX[abase[i] + ks * astep];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double b = z ? 0.0 : sB * Bm[2 * ks * LD + 8 * j + boff];
#pragma unroll
                        for (int i = 0; i < MTW; ++i)
                            if (i < mtn) dmma884(acc[i][j], a[i], b);
                    }
                }
                // epilogue: out1 = extra + e1 E + x1 Z,  out2 = e2 E + x2 Z   (Z = accumulator, E = AS | AhM | A_j S | A_j^H M)
#pragma unroll
                for (int i = 0; i < MTW; ++i) {
                    if (i >= mtn) continue;
                    const int pb = prow[i] >> 4, r = prow[i] & 15, j = pb - 1;
                    if (pb > nd) continue;
                    double e1, x1, e2, x2;
                    int q1, q2;
                    if (!is_b && j < 0) { e1 = -c1h; x1 = c2h2; e2 = -0.5; x2 = c2h; q1 = QO_R; q2 = QO_TH; }
                    else if (is_b && j < 0) { e1 = -0.5; x1 = -c2h; e2 = -0.5; x2 = c2h; q1 = QO_KH0; q2 = QO_KH1; }
                    else if (!is_b) { e1 = -c1h; x1 = c2h2; e2 = 0.0; x2 = 0.0; q1 = QO_TA + j; q2 = -1; }
                    else { e1 = -c1h; x1 = -c2h2; e2 = -c1h; x2 = c2h2; q1 = QO_KA0 + j; q2 = QO_KA1 + j; }
                    const bool want2 = q2 >= 0 && c.pl_base[q2] >= 0;
                    const int b1 = c.pl_base[q1], s1 = c.pl_stride[q1];
                    const int b2 = q2 >= 0 ? c.pl_base[q2] : 0, s2 = q2 >= 0 ? c.pl_stride[q2] : 0;
                    const double* const Es = is_b ? mM : mS;
                    const int eo = j < 0 ? 0 : ((j * 2 + (is_b ? 1 : 0)) * N + r) * W;
#pragma unroll
                    for (int jt = 0; jt < NT; ++jt) {
                        const int n = 4 * jt + tq;
                        if (n >= N) continue;
                        double2 E;
                        if (j < 0) {
                            E = make_double2(i < 2 ? keep[i < 2 ? i : 0][jt][0] : 0.0, i < 2 ? keep[i < 2 ? i : 0][jt][1] : 0.0);
                        } else {
                            E = make_double2(0.0, 0.0);
                            for (int u = 0; u < W; ++u) cfma(E, ellv[eo + u], *reinterpret_cast<const double2*>(Es + ellc[eo + u] * LD + 2 * n));
                        }
                        const double zr = acc[i][jt][0], zi = acc[i][jt][1];
                        double o1r = e1 * E.x + x1 * zr, o1i = e1 * E.y + x1 * zi;
                        if (q1 == QO_R) { const double2 d = *reinterpret_cast<const double2*>(mD + r * LD + 2 * n); o1r += d.x; o1i += d.y; }
                        const int ire = n * n2 + r;
                        if (b1 >= 0) {
                            image[b1 + ire * s1] = o1r;
                            image[b1 + (ire + N) * s1] = o1i;
                        }
                        if (want2) {
                            image[b2 + ire * s2] = e2 * E.x + x2 * zr;
                            image[b2 + (ire + N) * s2] = e2 * E.y + x2 * zi;
                        }
                    }
                }
            }
            if (needJ) {
                // -iso(F), +iso(B) blocks from A and A2 (odd warps; even ones carry the rows that need `keep`)
                if (warp & 1)
                    for (int e = lane + 32 * (warp >> 1); e < N * N; e += 64) {
                        const int r = e / N, cc = e - r * N;
                        const double2 a = *reinterpret_cast<const double2*>(mA + r * LD + 2 * cc);
                        const double2 a2 = *reinterpret_cast<const double2*>(mX + r * LD + 2 * cc);
                        const double id = r == cc ? 1.0 : 0.0;
                        const double fr = id + c1h * a.x + c2h2 * a2.x, fi = c1h * a.y + c2h2 * a2.y;
                        const double br = id - c1h * a.x + c2h2 * a2.x, bi = -c1h * a.y + c2h2 * a2.y;
                        const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                        PUT(QO_ISOF, k00, -fr); PUT(QO_ISOF, k00 + N, -fi); PUT(QO_ISOF, k01, fi); PUT(QO_ISOF, k01 + N, -fr);
                        PUT(QO_ISOB, k00, br);  PUT(QO_ISOB, k00 + N, bi);  PUT(QO_ISOB, k01, -bi); PUT(QO_ISOB, k01 + N, br);
                    }
            }
            if (needH) {
                // scalar second derivatives as traces against G = D M^H and G2 = S M^H:
                //   dt x dt = 1/6 Re tr(A2 G);  a_j x dt = -1/2 Re tr(A_j G2) + h/6 Re tr(C_j G);  a_i x a_j = h^2/12 Re tr({A_i, A_j} G)
                const double* const mG2 = mG + N * LD;
                if (free_time)
                    for (int task = warp; task < 1 + nd; task += nwarps_) {
                        const int j = task - 1;
                        const double* X = mX + task * N * LD;
                        double s1 = 0.0, s2 = 0.0;
                        for (int e = lane; e < N * N; e += 32) {
                            const int r = e / N, k = e - r * N;
                            const double2 xv = *reinterpret_cast<const double2*>(X + r * LD + 2 * k);
                            const double2 gv = *reinterpret_cast<const double2*>(mG + k * LD + 2 * r);
                            s2 = fma(xv.x, gv.x, s2);
                            s2 = fma(-xv.y, gv.y, s2);
                        }
                        if (j >= 0)
                            for (int e = lane; e < N * W; e += 32) {
                                const int r = e / W;
                                const double2 av = ellv[(j * 2) * N * W + e];
                                const double2 gv = *reinterpret_cast<const double2*>(mG2 + ellc[(j * 2) * N * W + e] * LD + 2 * r);
                                s1 = fma(av.x, gv.x, s1);
                                s1 = fma(-av.y, gv.y, s1);
                            }
                        s1 = warp_sum(s1);
                        s2 = warp_sum(s2);
                        const int q = j < 0 ? QO_HHH : QO_HAH + j;
                        if (lane == 0 && c.pl_base[q] >= 0) image[c.pl_base[q]] = j < 0 ? s2 * (1.0 / 6.0) : -0.5 * s1 + c2h * s2;
                    }
                if (warp == nwarps_ - 1 && lane < npair) {  // a_i x a_j; tasks in (j, i <= j) order, one lane each
                    int j = 0, rem = lane;
                    while (rem > j) { rem -= j + 1; ++j; }
                    double val = 0.0;
                    for (int u = kkptr[lane]; u < kkptr[lane + 1]; ++u) {
                        const int rc = kkrc[u];
                        const double2 kv = kkv[u];
                        const double2 gv = *reinterpret_cast<const double2*>(mG + (rc & 255) * LD + 2 * (rc >> 8));  // K[r, k] G[k, r]
                        val = fma(kv.x, gv.x, val);
                        val = fma(-kv.y, gv.y, val);
                    }
                    const int q = qo_haa(rem, j);
                    if (c.pl_base[q] >= 0) image[c.pl_base[q]] = val * c2h2;
                }
            }
        }
        if (QCK_BULK_STORE) fence_async_smem();
        GSYNC();
        } else if constexpr (KIND == QK_PADE4) {
        // ---- stage 1: A2 = A A (+ F, B blocks), AS = A S, AhM = A^H M, G = D M^H, G2 = S M^H;  C_j = A_j A + A A_j ----
        {
            const int nG = needH ? (free_time ? 2 : 1) : 0;
            const int nA = (1 + nG) * tilesA;
            const int nS = (needH && free_time ? 2 : 1) * tilesS;
            const int nDense = nA + nS;
            for (int w = vtid; w < nDense; w += nthreads) {
                if constexpr (TC == QCK_TILE) {
                    // unitaries: every stage-1 product is an N x N x K product in 3 x 3 tiles -> ONE instruction stream for the
                    // whole warp (separate A-type / state-type branches would run one after the other inside a warp)
                    const int pi = w / tilesA, tl = w - pi * tilesA;
                    const int r0 = (tl / tcolsA) * QCK_TILE, c0 = (tl - (tl / tcolsA) * tcolsA) * QCK_TILE;
                    // product list: [A A] [D M^H] [S M^H] (1 + nG of them), then [A S] [A^H M]
                    const int ps = pi - (1 + nG);  // >= 0: state-type product
                    const double2* Aop = ps >= 0 ? MA(QA_A) : (pi == 0 ? MA(QA_A) : (pi == 1 ? MS(QS_D) : MS(QS_S)));
                    const double2* Bop = ps >= 0 ? (ps == 0 ? MS(QS_S) : MS(QS_M)) : (pi == 0 ? MA(QA_A) : MS(QS_M));
                    double2* Cop = ps >= 0 ? MS(ps == 0 ? QS_AS : QS_AHM) : MA(pi == 0 ? QA_A2 : (pi == 1 ? QA_G : QA_G2));
                    double2 acc[QCK_TILE][QCK_TILE];
                    tile_mm<QCK_TILE>(Aop, ps == 1, Bop, ps < 0 && pi != 0, N, NP, r0, c0, acc);
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int j = 0; j < QCK_TILE; ++j) Cop[r0 + i + NP * (c0 + j)] = acc[i][j];
                    if (pi == 0 && needJ) {
#pragma unroll
                        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                            for (int j = 0; j < QCK_TILE; ++j) {
                                const int r = r0 + i, cc = c0 + j;
                                if (r < N && cc < N) {
                                    const double2 a = MA(QA_A)[r + NP * cc];
                                    const double id = r == cc ? 1.0 : 0.0;
                                    const double fr = id + c1h * a.x + c2h2 * acc[i][j].x, fi = c1h * a.y + c2h2 * acc[i][j].y;
                                    const double br = id - c1h * a.x + c2h2 * acc[i][j].x, bi = -c1h * a.y + c2h2 * acc[i][j].y;
                                    const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                                    PUT(QO_ISOF, k00, -fr); PUT(QO_ISOF, k00 + N, -fi); PUT(QO_ISOF, k01, fi); PUT(QO_ISOF, k01 + N, -fr);
                                    PUT(QO_ISOB, k00, br);  PUT(QO_ISOB, k00 + N, bi);  PUT(QO_ISOB, k01, -bi); PUT(QO_ISOB, k01 + N, br);
                                }
                            }
                    }
                } else if (w < nA) {
                    const int pi = w / tilesA, tl = w - pi * tilesA;
                    const int r0 = (tl / tcolsA) * QCK_TILE, c0 = (tl - (tl / tcolsA) * tcolsA) * QCK_TILE;
                    double2 acc[QCK_TILE][QCK_TILE];
                    // pi 0: A A (K = N);  pi 1: D M^H;  pi 2: S M^H  (K = nc, B operand conj-transposed)
                    const double2* Aop = pi == 0 ? MA(QA_A) : (pi == 1 ? MS(QS_D) : MS(QS_S));
                    const double2* Bop = pi == 0 ? MA(QA_A) : MS(QS_M);
                    tile_mm<QCK_TILE>(Aop, false, Bop, pi != 0, pi == 0 ? N : nc, NP, r0, c0, acc);
                    double2* Cop = MA(pi == 0 ? QA_A2 : (pi == 1 ? QA_G : QA_G2));
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int j = 0; j < QCK_TILE; ++j) {
                            const int r = r0 + i, cc = c0 + j;
                            Cop[r + NP * cc] = acc[i][j];
                            if (pi == 0 && needJ && r < N && cc < N) {
                                const double2 a = MA(QA_A)[r + NP * cc];
                                const double id = r == cc ? 1.0 : 0.0;
                                const double fr = id + c1h * a.x + c2h2 * acc[i][j].x, fi = c1h * a.y + c2h2 * acc[i][j].y;
                                const double br = id - c1h * a.x + c2h2 * acc[i][j].x, bi = -c1h * a.y + c2h2 * acc[i][j].y;
                                const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                                PUT(QO_ISOF, k00, -fr); PUT(QO_ISOF, k00 + N, -fi); PUT(QO_ISOF, k01, fi); PUT(QO_ISOF, k01 + N, -fr);
                                PUT(QO_ISOB, k00, br);  PUT(QO_ISOB, k00 + N, bi);  PUT(QO_ISOB, k01, -bi); PUT(QO_ISOB, k01 + N, br);
                            }
                        }
                } else {
                    const int w2 = w - nA;
                    const int pi = w2 / tilesS, tl = w2 - pi * tilesS;
                    const int r0 = (tl / tcols) * QCK_TILE, c0 = (tl - (tl / tcols) * tcols) * TC;
                    double2 acc[QCK_TILE][TC];
                    // pi 0: A S;  pi 1: A^H M
                    tile_mm<TC>(MA(QA_A), pi == 1, pi == 0 ? MS(QS_S) : MS(QS_M), false, N, NP, r0, c0, acc);
                    double2* Cop = MS(pi == 0 ? QS_AS : QS_AHM);
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int j = 0; j < TC; ++j) Cop[r0 + i + NP * (c0 + j)] = acc[i][j];
                }
            }
            if (needT) {
                // C_j[r, c] = sum_w A_j[r, k_w] A[k_w, c] + sum_w A[r, k_w] A_j[k_w, c]; the second sum walks row c of A_j^H
                // (A_j[k, c] = conj(A_j^H[c, k])).  Runs on the warps the dense tiles left free.
                int first = ((nDense + 31) >> 5) << 5;
                if (first >= nthreads) first = 0;
                const int nsp = nd * N * N, stride = nthreads - first;
                for (int w = vtid - first; w >= 0 && w < nsp; w += stride) {
                    const int j = w / (N * N), e = w - j * N * N;
                    const int r = e % N, cc = e / N;
                    const int o0 = ((j * 2) * N + r) * W, o1 = ((j * 2 + 1) * N + cc) * W;
                    const double2* Acol = MA(QA_A) + NP * cc;
                    const double2* Arow = MA(QA_A) + r;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int u = 0; u < W; ++u) {
                        cfma(acc, ellv[o0 + u], Acol[ellc[o0 + u]]);
                        double2 ah = ellv[o1 + u];
                        ah.y = -ah.y;
                        cfma(acc, Arow[NP * ellc[o1 + u]], ah);
                    }
                    MA(QA_C + j)[r + NP * cc] = acc;
                }
            }
        }
        GSYNC();
        QCK_TICK(4);

        // ---- stage 2: A2 D, (A2)^H M, C_j D, C_j^H M with fused epilogues into the image;  scalar traces ------------
        {
            const bool hdt = needH && free_time;  // timestep Hessian blocks exist
            const int nP = 1 + (needT ? nd : 0) + (hdt ? 1 : 0) + (needH ? nd : 0);
            const int nDense = nP * tilesS;
            for (int w = vtid; w < nDense; w += nthreads) {
                const int pi = w / tilesS, tl = w - pi * tilesS;
                const int r0 = (tl / tcols) * QCK_TILE, c0 = (tl - (tl / tcols) * tcols) * TC;
                // product list: [A2 D -> R, TH] [C_j D -> TA_j]*nd [A2^H M -> KH0, KH1] [C_j^H M -> KA0_j, KA1_j]*nd
                bool adj = false;
                int j = -1, rest = pi - 1;
                if (pi > 0) {
                    if (needT && rest < nd) j = rest;
                    else {
                        if (needT) rest -= nd;
                        adj = true;
                        if (hdt && rest == 0) j = -1;
                        else j = rest - (hdt ? 1 : 0);
                    }
                }
                const double2* Aop = j < 0 ? MA(QA_A2) : MA(QA_C + j);
                double2 acc[QCK_TILE][TC];
                tile_mm<TC>(Aop, adj, adj ? MS(QS_M) : MS(QS_D), false, N, NP, r0, c0, acc);
                // epilogue: out1 = extra + e1 E + x1 X,  out2 = e2 E + x2 X   (X = tile result, E = AS | AhM | A_j S | A_j^H M)
                double e1, x1, e2, x2;
                int q1, q2;
                const double2* Esrc;
                if (j < 0 && !adj) { e1 = -c1h; x1 = c2h2; e2 = -0.5; x2 = c2h; q1 = QO_R; q2 = QO_TH; Esrc = MS(QS_AS); }
                else if (j < 0) { e1 = -0.5; x1 = -c2h; e2 = -0.5; x2 = c2h; q1 = QO_KH0; q2 = QO_KH1; Esrc = MS(QS_AHM); }
                else if (!adj) { e1 = -c1h; x1 = c2h2; e2 = 0.0; x2 = 0.0; q1 = QO_TA + j; q2 = -1; Esrc = MS(QS_S); }
                else { e1 = -c1h; x1 = -c2h2; e2 = -c1h; x2 = c2h2; q1 = QO_KA0 + j; q2 = QO_KA1 + j; Esrc = MS(QS_M); }
                const bool want2 = q2 >= 0 && c.pl_base[q2] >= 0;
                const int b1 = c.pl_base[q1], s1 = c.pl_stride[q1];
                const int b2 = q2 >= 0 ? c.pl_base[q2] : 0, s2 = q2 >= 0 ? c.pl_stride[q2] : 0;
                // E tile: stored matrix (A2 products) or sparse product with the constant drive matrix, row data shared by the tile's columns
                double2 E[QCK_TILE][TC];
                if (j < 0) {
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int jj = 0; jj < TC; ++jj) E[i][jj] = Esrc[r0 + i + NP * (c0 + jj)];
                } else {
                    const int eo = (j * 2 + (adj ? 1 : 0)) * N * W;
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i) {
#pragma unroll
                        for (int jj = 0; jj < TC; ++jj) E[i][jj] = make_double2(0.0, 0.0);
                        if (r0 + i < N)
                            for (int u = 0; u < W; ++u) {
                                const double2 v = ellv[eo + (r0 + i) * W + u];
                                const double2* xr = Esrc + ellc[eo + (r0 + i) * W + u] + NP * c0;
#pragma unroll
                                for (int jj = 0; jj < TC; ++jj) cfma(E[i][jj], v, xr[NP * jj]);
                            }
                    }
                }
#pragma unroll
                for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                    for (int jj = 0; jj < TC; ++jj) {
                        const int r = r0 + i, cc = c0 + jj;
                        if (r < N && cc < nc) {
                            const double2 X = acc[i][jj];
                            double o1r = e1 * E[i][jj].x + x1 * X.x, o1i = e1 * E[i][jj].y + x1 * X.y;
                            if (q1 == QO_R) { const double2 d = MS(QS_D)[r + NP * cc]; o1r += d.x; o1i += d.y; }
                            const int ire = cc * n2 + r;
                            if (b1 >= 0) {
                                image[b1 + ire * s1] = o1r;
                                image[b1 + (ire + N) * s1] = o1i;
                            }
                            if (want2) {
                                image[b2 + ire * s2] = e2 * E[i][jj].x + x2 * X.x;
                                image[b2 + (ire + N) * s2] = e2 * E[i][jj].y + x2 * X.y;
                            }
                        }
                    }
            }
            if (needH) {
                // scalar second derivatives as traces against G = D M^H and G2 = S M^H, on the warps the dense tiles left
                // free: the N^2-term traces (dt x dt, a_j x dt) take 4 lanes each, the sparse a_i x a_j traces one lane each
                int first = ((nDense + 31) >> 5) << 5;
                if (first >= nthreads) first = 0;
                const int fw = first >> 5, nfw = (nthreads >> 5) - fw;
                const int n4 = free_time ? 1 + nd : 0, nv = 4 * n4 + npair;
                if (vwarp >= fw)
                    for (int vb = (vwarp - fw) * 32; vb < nv; vb += nfw * 32) {
                        const int v = vb + lane;
                        double val = 0.0;
                        int q = -1;
                        if (v < 4 * n4) {
                            const int task = v >> 2, part = v & 3, j = task - 1;
                            const double2* X = task == 0 ? MA(QA_A2) : MA(QA_C + j);
                            double s1 = 0.0, s2 = 0.0;
                            for (int e = part; e < N * N; e += 4) {  // Re tr(X G)
                                const int r = e % N, k = e / N;
                                const double2 xv = X[r + NP * k], g = MA(QA_G)[k + NP * r];
                                s2 = fma(xv.x, g.x, s2);
                                s2 = fma(-xv.y, g.y, s2);
                            }
                            if (task > 0)
                                for (int e = part; e < N * W; e += 4) {  // Re tr(A_j G2), A_j sparse
                                    const int r = e / W;
                                    const double2 av = ellv[(j * 2) * N * W + e], g = MA(QA_G2)[ellc[(j * 2) * N * W + e] + NP * r];
                                    s1 = fma(av.x, g.x, s1);
                                    s1 = fma(-av.y, g.y, s1);
                                }
                            val = task == 0 ? s2 * (1.0 / 6.0) : -0.5 * s1 + c2h * s2;  // dt x dt | a_j x dt
                            q = part == 0 ? (task == 0 ? QO_HHH : QO_HAH + j) : -1;
                        } else if (v < nv) {  // a_i x a_j = h^2/12 Re tr({A_i, A_j} G); tasks in (j, i <= j) order
                            const int task = v - 4 * n4;
                            int j = 0, rem = task;
                            while (rem > j) { rem -= j + 1; ++j; }
                            for (int u = kkptr[task]; u < kkptr[task + 1]; ++u) {
                                const int rc = kkrc[u];
                                const double2 kv = kkv[u], g = MA(QA_G)[(rc & 255) + NP * (rc >> 8)];  // K[r, k] G[k, r]
                                val = fma(kv.x, g.x, val);
                                val = fma(-kv.y, g.y, val);
                            }
                            val *= c2h2;
                            q = qo_haa(rem, j);
                        }
                        double red = val + __shfl_xor_sync(0xffffffffu, val, 1);
                        red += __shfl_xor_sync(0xffffffffu, red, 2);
                        if (v < 4 * n4) val = red;
                        if (q >= 0 && c.pl_base[q] >= 0) image[c.pl_base[q]] = val;
                    }
            }
        }
        if (QCK_BULK_STORE) fence_async_smem();  // image writes -> visible to the copy engine
        GSYNC();
        } else if constexpr (KIND == QK_EXP) {
        // ============================ exponential integrators =============================================================
        // residual U1 - exp(h A) U0.  exp and its Frechet derivatives d/da_j by scaling and squaring of a degree-8 Taylor
        // polynomial: Y = h A / 2^s with ||Y||_1 <= 1/16 (truncation: exp 4e-17, first derivatives 6e-15 relative),
        //   Horner:    P <- I + (Y/m) P,        L_j <- (Y_j P + Y L_j)/m        m = 8 .. 1,  Y_j = h A_j / 2^s
        //   squaring:  E <- E E,                L_j <- E L_j + L_j E            s times
        // Outputs: -iso(E) block, identity block, d/da_j = -L_j U0, d/dh = -A E U0.
        constexpr int TK = 8;
        // N x N products use 3 x 1 tiles here: the phases are strictly sequential (7 Horner steps + s squarings), so the
        // latency of one phase matters more than shared-memory traffic -> three times as many, three times shorter tasks
        constexpr int XC = 1;
        const int tcolsX = NP / XC, tilesX = (NP / QCK_TILE) * tcolsX;
        const int nthr_tiles = (1 + nd) * tilesX;
#define XE(b) MA(1 + (b))
#define XL(b, j) MA(3 + (b) * nd + (j))
        // scaling parameter from the 1-norm (|re| + |im| per entry bounds the modulus)
        if (tid < N) {
            double cs = 0.0;
            for (int r = 0; r < N; ++r) { const double2 a = MA(QA_A)[r + NP * tid]; cs += fabs(a.x) + fabs(a.y); }
            image[tid] = cs;
        }
        GSYNC();
        double nrm = 0.0;
        for (int cc = 0; cc < N; ++cc) nrm = fmax(nrm, image[cc]);
        nrm *= fabs(h);
        int sq = 0;
        while (nrm > 0.0625 && sq < 40) { nrm *= 0.5; ++sq; }
        if (c.tape && sq > c.tape_levels) {
            // ||h A||_1 > 2^(tape_levels - 4): the Hessian tape cannot hold that many squaring levels.  The same clamp applies
            // to every call on the handle (with or without the Hessian in the mask) and the host reports QCK_ERANGE.
            sq = c.tape_levels;
            if (tid == 0 && p.status) atomicOr(p.status, QCK_ST_EXP_RANGE);
        }
        const double y = ldexp(h, -sq);
        int cur = 0;
        {   // Horner start (m = TK): P = I + (Y/TK), L_j = Y_j/TK
            const double c0 = y / TK;
            for (int e = tid; e < N * N; e += nthreads) {
                const int r = e % N, cc = e / N;
                const double2 a = MA(QA_A)[r + NP * cc];
                XE(0)[r + NP * cc] = make_double2((r == cc ? 1.0 : 0.0) + c0 * a.x, c0 * a.y);
                for (int j = 0; j < nd; ++j) {
                    const int o = ((j * 2) * N + r) * W;
                    double2 v = make_double2(0.0, 0.0);
                    for (int u = 0; u < W; ++u)
                        if (ellc[o + u] == cc) { v.x += ellv[o + u].x; v.y += ellv[o + u].y; }
                    XL(0, j)[r + NP * cc] = make_double2(c0 * v.x, c0 * v.y);
                }
            }
        }
        GSYNC();
        // Hessian: the inputs of every Horner step / squaring level are kept on a per-CTA tape in global memory (L2)
        double2* tapeH = c.tape ? c.tape + (size_t)gid * c.tape_stride : nullptr;  // [(TK-1) steps][nd][N*N]
        double2* tapeS = tapeH ? tapeH + (size_t)(TK - 1) * nd * N * N : nullptr;        // [levels][1 + nd][N*N]
        const bool taping = needH && tapeH != nullptr;
        auto tape_put = [&](double2* dst, const double2* src) {
            for (int e = tid; e < N * N; e += nthreads) dst[e] = src[(e % N) + NP * (e / N)];
        };
        auto tape_get = [&](double2* dst, const double2* src) {
            for (int e = tid; e < N * N; e += nthreads) dst[(e % N) + NP * (e / N)] = src[e];
        };
        for (int mth = TK - 1; mth >= 1; --mth) {
            const double cm = y / mth;
            if (taping)
                for (int j = 0; j < nd; ++j) tape_put(tapeH + (size_t)((TK - 1 - mth) * nd + j) * N * N, XL(cur, j));
            for (int w = tid; w < nthr_tiles; w += nthreads) {
                const int pi = w / tilesX, tl = w - pi * tilesX;
                const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                double2 acc[QCK_TILE][XC];
                tile_mm<XC>(MA(QA_A), false, pi == 0 ? XE(cur) : XL(cur, pi - 1), false, N, NP, r0, c0, acc);
                double2* Cop = pi == 0 ? XE(cur ^ 1) : XL(cur ^ 1, pi - 1);
                const int eo = ((pi - 1) * 2) * N * W;
#pragma unroll
                for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                    for (int jj = 0; jj < XC; ++jj) {
                        const int r = r0 + i, cc = c0 + jj;
                        double2 v = acc[i][jj];
                        if (pi == 0) {
                            v = make_double2((r == cc ? 1.0 : 0.0) + cm * v.x, cm * v.y);
                        } else {
                            if (r < N && cc < N)
                                for (int u = 0; u < W; ++u) cfma(v, ellv[eo + r * W + u], XE(cur)[ellc[eo + r * W + u] + NP * cc]);
                            v = make_double2(cm * v.x, cm * v.y);
                        }
                        if (r < N && cc < N) Cop[r + NP * cc] = v;
                    }
            }
            GSYNC();
            cur ^= 1;
        }
        for (int k = 0; k < sq; ++k) {
            if (taping) {
                tape_put(tapeS + (size_t)(k * (1 + nd)) * N * N, XE(cur));
                for (int j = 0; j < nd; ++j) tape_put(tapeS + (size_t)(k * (1 + nd) + 1 + j) * N * N, XL(cur, j));
            }
            for (int w = tid; w < nthr_tiles; w += nthreads) {
                const int pi = w / tilesX, tl = w - pi * tilesX;
                const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                double2 acc[QCK_TILE][XC];
                if (pi == 0) {
                    tile_mm<XC>(XE(cur), false, XE(cur), false, N, NP, r0, c0, acc);
                } else {
                    tile_mm<XC>(XE(cur), false, XL(cur, pi - 1), false, N, NP, r0, c0, acc);
                    tile_mm<XC, false>(XL(cur, pi - 1), false, XE(cur), false, N, NP, r0, c0, acc);
                }
                double2* Cop = pi == 0 ? XE(cur ^ 1) : XL(cur ^ 1, pi - 1);
#pragma unroll
                for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                    for (int jj = 0; jj < XC; ++jj)
                        if (r0 + i < N && c0 + jj < N) Cop[r0 + i + NP * (c0 + jj)] = acc[i][jj];
            }
            GSYNC();
            cur ^= 1;
        }
        QCK_TICK(4);
        // ---- outputs 1: E U0 (-> residual), L_j U0 (-> d/da_j), -iso(E) block; Hessian: -L_j^H M, A^H M, Gamma = U0 M^H ------
        const bool hdt = needH && free_time;
        {
            const int nP = 1 + (needT ? nd : 0) + (needH ? nd : 0) + (hdt ? 1 : 0);
            const int nS = nP * tilesS, nDense = nS + (needH ? tilesX : 0);
            for (int w = tid; w < nDense; w += nthreads) {
                if (w >= nS) {  // Gamma = U0 M^H  (Re <M, K U0> = Re tr(K Gamma)) into the idle E buffer
                    const int tl = w - nS;
                    const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                    double2 acc[QCK_TILE][XC];
                    tile_mm<XC>(MS(QS_D), false, MS(QS_M), true, nc, NP, r0, c0, acc);
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int jj = 0; jj < XC; ++jj) XE(cur ^ 1)[r0 + i + NP * (c0 + jj)] = acc[i][jj];
                    continue;
                }
                const int pi = w / tilesS, tl = w - pi * tilesS;
                const int r0 = (tl / tcols) * QCK_TILE, c0 = (tl - (tl / tcols) * tcols) * TC;
                // product list: [E U0] [L_j U0]*nd (needT) [L_j^H M]*nd (needH) [A^H M] (hdt)
                int kind = 0, j = -1, rest = pi - 1;  // kind 0: E U0, 1: L_j U0, 2: L_j^H M, 3: A^H M
                if (pi > 0) {
                    if (needT && rest < nd) { kind = 1; j = rest; }
                    else {
                        if (needT) rest -= nd;
                        if (needH && rest < nd) { kind = 2; j = rest; }
                        else kind = 3;
                    }
                }
                const double2* Aop = kind == 0 ? XE(cur) : (kind == 3 ? MA(QA_A) : XL(cur, j));
                double2 acc[QCK_TILE][TC];
                tile_mm<TC>(Aop, kind >= 2, kind >= 2 ? MS(QS_M) : MS(QS_D), false, N, NP, r0, c0, acc);
                const int q1 = kind == 0 ? QO_R : (kind == 1 ? QO_TA + j : (kind == 2 ? QO_KA0 + j : -1));
                const int b1 = q1 >= 0 ? c.pl_base[q1] : -1, s1 = q1 >= 0 ? c.pl_stride[q1] : 0;
#pragma unroll
                for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                    for (int jj = 0; jj < TC; ++jj) {
                        const int r = r0 + i, cc = c0 + jj;
                        if (r < N && cc < nc) {
                            double2 v = acc[i][jj];
                            if (kind == 3) { MS(QS_AHM)[r + NP * cc] = v; continue; }
                            if (kind == 0) {
                                MS(QS_AS)[r + NP * cc] = v;  // E U0, operand of the d/dh product
                                const double2 u1 = MS(QS_S)[r + NP * cc];
                                v = make_double2(v.x - u1.x, v.y - u1.y);
                            }
                            const int ire = cc * n2 + r;
                            if (b1 >= 0) { image[b1 + ire * s1] = -v.x; image[b1 + (ire + N) * s1] = -v.y; }
                        }
                    }
            }
            if (needJ) {
                int first = ((nDense + 31) >> 5) << 5;
                if (first >= nthreads) first = 0;
                for (int e = tid - first; e >= 0 && e < N * N; e += nthreads - first) {
                    const int r = e % N, cc = e / N;
                    const double2 v = XE(cur)[r + NP * cc];
                    const int k00 = r + n2 * cc, k01 = r + n2 * (cc + N);
                    PUT(QO_ISOF, k00, -v.x); PUT(QO_ISOF, k00 + N, -v.y); PUT(QO_ISOF, k01, v.y); PUT(QO_ISOF, k01 + N, -v.x);
                }
            }
        }
        GSYNC();
        // ---- outputs 2: V = A (E U0) -> d/dh (V kept in an idle L buffer);  state_t x dt = -E^H (A^H M) ---------------------
        double2* Vbuf = XL(cur ^ 1, 0);
        if (free_time && needT) {
            const int nP = 1 + (hdt ? 1 : 0);
            for (int w = tid; w < nP * tilesS; w += nthreads) {
                const int pi = w / tilesS, tl = w - pi * tilesS;
                const int r0 = (tl / tcols) * QCK_TILE, c0 = (tl - (tl / tcols) * tcols) * TC;
                double2 acc[QCK_TILE][TC];
                tile_mm<TC>(pi == 0 ? MA(QA_A) : XE(cur), pi == 1, pi == 0 ? MS(QS_AS) : MS(QS_AHM), false, N, NP, r0, c0, acc);
                const int q1 = pi == 0 ? QO_TH : QO_KH0;
                const int b1 = c.pl_base[q1], s1 = c.pl_stride[q1];
#pragma unroll
                for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                    for (int jj = 0; jj < TC; ++jj) {
                        const int r = r0 + i, cc = c0 + jj;
                        if (r < N && cc < nc) {
                            if (pi == 0) Vbuf[r + NP * cc] = acc[i][jj];
                            const int ire = cc * n2 + r;
                            if (b1 >= 0) { image[b1 + ire * s1] = -acc[i][jj].x; image[b1 + (ire + N) * s1] = -acc[i][jj].y; }
                        }
                    }
            }
        }
        GSYNC();
        if (needH) {
            // ---- dt x dt = -Re <A^H M, V>,  a_j x dt = -Re <M, A_j (E U0)> - Re <A^H M, L_j U0>  (one warp per scalar) ----------
            if (free_time)
                for (int task = warp; task < 1 + nd; task += nwarps_) {
                    double sacc = 0.0;
                    const int j = task - 1;
                    const int bt = j >= 0 ? c.pl_base[QO_TA + j] : 0, st = j >= 0 ? c.pl_stride[QO_TA + j] : 0;
                    for (int e = lane; e < N * nc; e += 32) {
                        const int r = e % N, cc = e / N, o = r + NP * cc;
                        const double2 ahm = MS(QS_AHM)[o];
                        if (j < 0) {
                            const double2 v = Vbuf[o];
                            sacc -= ahm.x * v.x + ahm.y * v.y;
                        } else {
                            const double2 aeu = ell_row(ellv + (j * 2) * N * W, ellc + (j * 2) * N * W, W, MS(QS_AS), NP, r, cc);
                            const double2 mm = MS(QS_M)[o];
                            const int ire = cc * n2 + r;
                            const double lr = -image[bt + ire * st], li = -image[bt + (ire + N) * st];  // L_j U0
                            sacc -= mm.x * aeu.x + mm.y * aeu.y + ahm.x * lr + ahm.y * li;
                        }
                    }
                    sacc = warp_sum(sacc);
                    if (lane == 0) {
                        const int q = j < 0 ? QO_HHH : QO_HAH + j;
                        if (c.pl_base[q] >= 0) image[c.pl_base[q]] = sacc;
                    }
                }
            // ---- a_i x a_j = -Re tr(Gamma d2E/da_i da_j): reverse sweep over the tape ------------------------------------------
            //   squaring level k (E_k = E_{k-1}^2):  += Re tr(Gamma_k (L^i L^j + L^j L^i)),  Gamma_{k-1} = Gamma_k E + E Gamma_k
            //   Horner step (P' = I + Y P / m):       += (y/m) Re tr(Lam (A_i L^j + A_j L^i)),  Lam' = (y/m) Lam A
            // with E = E_{k-1}, L = L_{k-1} (resp. the step's input jets) read back from the tape.
            double hacc[QCK_MAX_DRIVES * (QCK_MAX_DRIVES + 1) / 2];
#pragma unroll
            for (int q = 0; q < QCK_MAX_DRIVES * (QCK_MAX_DRIVES + 1) / 2; ++q) hacc[q] = 0.0;
            int gb = cur ^ 1;             // Gamma lives in XE(gb), the other E buffer receives the update
            double2* XT = MA(3 + 2 * nd);  // E_{k-1} read back from the tape
            auto pair_traces = [&](double wgt) {  // hacc[(i,j)] += wgt * Re(tr(Q^i L^j) + tr(Q^j L^i)),  Q in XL(1,.), L in XL(0,.)
                for (int e = tid; e < N * N; e += nthreads) {
                    const int a = e % N, b = e / N;
                    int q = 0;
                    for (int j = 0; j < nd; ++j) {
                        const double2 lj = XL(0, j)[b + NP * a], qj = XL(1, j)[a + NP * b];
                        for (int i = 0; i <= j; ++i, ++q) {
                            const double2 li = XL(0, i)[b + NP * a], qi = XL(1, i)[a + NP * b];
                            hacc[q] += wgt * (qi.x * lj.x - qi.y * lj.y + qj.x * li.x - qj.y * li.y);
                        }
                    }
                }
            };
            if (taping) {
                for (int k = sq - 1; k >= 0; --k) {
                    GSYNC();  // previous level's traces / products are done with the buffers
                    tape_get(XT, tapeS + (size_t)(k * (1 + nd)) * N * N);
                    for (int j = 0; j < nd; ++j) tape_get(XL(0, j), tapeS + (size_t)(k * (1 + nd) + 1 + j) * N * N);
                    GSYNC();
                    for (int w = tid; w < nthr_tiles; w += nthreads) {
                        const int pi = w / tilesX, tl = w - pi * tilesX;
                        const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                        double2 acc[QCK_TILE][XC];
                        if (pi == 0) {
                            tile_mm<XC>(XE(gb), false, XT, false, N, NP, r0, c0, acc);
                            tile_mm<XC, false>(XT, false, XE(gb), false, N, NP, r0, c0, acc);
                        } else {
                            tile_mm<XC>(XE(gb), false, XL(0, pi - 1), false, N, NP, r0, c0, acc);
                        }
                        double2* Cop = pi == 0 ? XE(gb ^ 1) : XL(1, pi - 1);
#pragma unroll
                        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                            for (int jj = 0; jj < XC; ++jj)
                                if (r0 + i < N && c0 + jj < N) Cop[r0 + i + NP * (c0 + jj)] = acc[i][jj];
                    }
                    GSYNC();
                    pair_traces(1.0);
                    gb ^= 1;
                }
                for (int n = TK - 2; n >= 0; --n) {
                    const double cm = y / (TK - 1 - n);
                    GSYNC();
                    for (int j = 0; j < nd; ++j) tape_get(XL(0, j), tapeH + (size_t)(n * nd + j) * N * N);
                    GSYNC();
                    // Q^i = Lam A_i (sparse, walks row b of A_i^H: A_i[k, b] = conj(A_i^H[b, k])), Lam' = cm Lam A (dense)
                    for (int w = tid; w < tilesX; w += nthreads) {
                        const int r0 = (w / tcolsX) * QCK_TILE, c0 = (w - (w / tcolsX) * tcolsX) * XC;
                        double2 acc[QCK_TILE][XC];
                        tile_mm<XC>(XE(gb), false, MA(QA_A), false, N, NP, r0, c0, acc);
#pragma unroll
                        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                            for (int jj = 0; jj < XC; ++jj)
                                if (r0 + i < N && c0 + jj < N) XE(gb ^ 1)[r0 + i + NP * (c0 + jj)] = make_double2(cm * acc[i][jj].x, cm * acc[i][jj].y);
                    }
                    {
                        int first = ((tilesX + 31) >> 5) << 5;
                        if (first >= nthreads) first = 0;
                        for (int w = tid - first; w >= 0 && w < nd * N * N; w += nthreads - first) {
                            const int i = w / (N * N), e = w - i * N * N;
                            const int a = e % N, b = e / N;
                            const int o1 = ((i * 2 + 1) * N + b) * W;
                            double2 v = make_double2(0.0, 0.0);
                            for (int u = 0; u < W; ++u) {
                                double2 ah = ellv[o1 + u];
                                ah.y = -ah.y;
                                cfma(v, XE(gb)[a + NP * ellc[o1 + u]], ah);
                            }
                            XL(1, i)[a + NP * b] = v;
                        }
                    }
                    GSYNC();
                    pair_traces(cm);
                    gb ^= 1;
                }
            }
            // block reduction of the pair sums: shuffles inside the warps, then across warps through the V buffer
            GSYNC();
            double* red = reinterpret_cast<double*>(XL(0, 0));
            for (int q = 0; q < npair; ++q) {
                const double v = warp_sum(hacc[q]);
                if (lane == 0) red[q * nwarps_ + warp] = v;
            }
            GSYNC();
            if (tid < npair) {
                double v = 0.0;
                for (int w2 = 0; w2 < nwarps_; ++w2) v += red[tid * nwarps_ + w2];
                int j = 0, rem = tid;
                while (rem > j) { rem -= j + 1; ++j; }
                const int qq = qo_haa(rem, j);
                if (c.pl_base[qq] >= 0) image[c.pl_base[qq]] = -v;
            }
            GSYNC();
        }
#undef XE
#undef XL
        } else {
        // ============================ Pade integrators of general order 2m (6, 8, 10, 12) ====================================
        // F = p(X), B = p(-X), X = h A, p(X) = sum_k c_k X^k evaluated in ratio form P_k = I + r_k X P_{k+1} (r_k = c_{k+1}/c_k,
        // P_m = I, p = P_0), with tangents in the directions a_j (dX = h A_j, sparse) and h (dX = A):
        //     T^d_k = r_k (X_d P_{k+1} + X T^d_{k+1}),
        // and the second derivatives contracted with Gamma = V M^H (V = U0 for F, U1 for B) by a reverse sweep
        //     += r_k Re tr(Lam_k (X_de P_{k+1} + X_d T^e_{k+1} + X_e T^d_{k+1})),  Lam_{k+1} = r_k Lam_k X,  X_(a_j h) = A_j.
        // The two polynomials are processed one after the other (sg = +1: F, sg = -1: B) in the same buffers; the image
        // entries that receive both parts (residual, d/da_j, d/dh) are assigned by the F pass and updated by the B pass.
        constexpr int XC = 1;
        const int tcolsX = NP / XC, tilesX = (NP / QCK_TILE) * tcolsX;
        const int mdeg = c.pade_m;
        const int ndir = nd + (free_time ? 1 : 0);  // tangent directions: a_0..a_{nd-1}, then h
#define XP(b) MA(1 + (b))
#define XJ(b, d) MA(3 + (b) * (nd + 1) + (d))
#define XT MA(3 + 2 * (nd + 1))
#define XG(b) MA(4 + 2 * (nd + 1) + (b))
        double2* tape = c.tape ? c.tape + (size_t)gid * c.tape_stride : nullptr;  // [(m-1) levels][1 + ndir][N*N]
        const bool taping = needH && tape != nullptr;
        auto tape_put = [&](double2* dst, const double2* src) {
            for (int e = tid; e < N * N; e += nthreads) dst[e] = src[(e % N) + NP * (e / N)];
        };
        auto tape_get = [&](double2* dst, const double2* src) {
            for (int e = tid; e < N * N; e += nthreads) dst[(e % N) + NP * (e / N)] = src[e];
        };
        constexpr int NPAIR_MAX = (QCK_MAX_DRIVES + 1) * (QCK_MAX_DRIVES + 2) / 2;
        double hacc[NPAIR_MAX];
#pragma unroll
        for (int q = 0; q < NPAIR_MAX; ++q) hacc[q] = 0.0;
        // hacc[e(e+1)/2 + d] += wgt * Re tr(X Y),  tr(X Y) = sum_ab X[a,b] Y[b,a]
        auto tr_acc = [&](int d, int e2, double wgt, const double2* X, const double2* Y) {
            double s = 0.0;
            for (int el = tid; el < N * N; el += nthreads) {
                const int a = el % N, b = el / N;
                const double2 x = X[a + NP * b], yv = Y[b + NP * a];
                s += x.x * yv.x - x.y * yv.y;
            }
            hacc[e2 * (e2 + 1) / 2 + d] += wgt * s;
        };
        for (int pass = 0; pass < 2; ++pass) {
            const double sg = pass == 0 ? 1.0 : -1.0;  // F = p(X), B = p(-X)
            const double2* V = pass == 0 ? MS(QS_D) : MS(QS_S);  // U0 | U1
            int cur = 0;
            GSYNC();
            {   // level m-1: P = I + sg r X,  T^j = sg r h A_j,  T^h = sg r A
                const double r = sg * c.pade_r[mdeg - 1];
                for (int e = tid; e < N * N; e += nthreads) {
                    const int rr = e % N, cc = e / N;
                    const double2 a = MA(QA_A)[rr + NP * cc];
                    XP(0)[rr + NP * cc] = make_double2((rr == cc ? 1.0 : 0.0) + r * h * a.x, r * h * a.y);
                    for (int j = 0; j < nd; ++j) {
                        const int o = ((j * 2) * N + rr) * W;
                        double2 v = make_double2(0.0, 0.0);
                        for (int u = 0; u < W; ++u)
                            if (ellc[o + u] == cc) { v.x += ellv[o + u].x; v.y += ellv[o + u].y; }
                        XJ(0, j)[rr + NP * cc] = make_double2(r * h * v.x, r * h * v.y);
                    }
                    if (free_time) XJ(0, nd)[rr + NP * cc] = make_double2(r * a.x, r * a.y);
                }
            }
            GSYNC();
            for (int kk = mdeg - 2; kk >= 0; --kk) {  // P_kk from level kk+1 (buffers `cur`)
                const double r = sg * c.pade_r[kk];
                if (taping) {
                    double2* tp = tape + (size_t)(kk * (1 + ndir)) * N * N;
                    tape_put(tp, XP(cur));
                    for (int d = 0; d < ndir; ++d) tape_put(tp + (size_t)(1 + d) * N * N, XJ(cur, d));
                }
                for (int w = tid; w < (1 + ndir) * tilesX; w += nthreads) {
                    const int pi = w / tilesX, tl = w - pi * tilesX;
                    const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                    double2 acc[QCK_TILE][XC], acc2[QCK_TILE][XC];
                    const int d = pi - 1;  // -1: P itself
                    tile_mm<XC>(MA(QA_A), false, d < 0 ? XP(cur) : XJ(cur, d), false, N, NP, r0, c0, acc);
                    if (d == nd) tile_mm<XC>(MA(QA_A), false, XP(cur), false, N, NP, r0, c0, acc2);  // X_h P = A P
                    double2* Cop = d < 0 ? XP(cur ^ 1) : XJ(cur ^ 1, d);
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int jj = 0; jj < XC; ++jj) {
                            const int rr = r0 + i, cc = c0 + jj;
                            if (rr < N && cc < N) {
                                double2 v = make_double2(h * acc[i][jj].x, h * acc[i][jj].y);  // X (.) = h A (.)
                                if (d < 0) {
                                    v = make_double2((rr == cc ? 1.0 : 0.0) + r * v.x, r * v.y);
                                } else {
                                    if (d < nd) {  // + h A_j P
                                        double2 s2 = make_double2(0.0, 0.0);
                                        const int eo = (d * 2) * N * W + rr * W;
                                        for (int u = 0; u < W; ++u) cfma(s2, ellv[eo + u], XP(cur)[ellc[eo + u] + NP * cc]);
                                        v.x += h * s2.x; v.y += h * s2.y;
                                    } else {       // + A P
                                        v.x += acc2[i][jj].x; v.y += acc2[i][jj].y;
                                    }
                                    v = make_double2(r * v.x, r * v.y);
                                }
                                Cop[rr + NP * cc] = v;
                            }
                        }
                }
                GSYNC();
                cur ^= 1;
            }
            // ---- outputs of this polynomial: P_0 V, T^d_0 V (accumulated), (T^d_0)^H M, iso block, Gamma_0 = V M^H ---------
            {
                const int nP = 1 + ndir + (needH ? ndir : 0);
                const int nS = nP * tilesS, nDense = nS + (needH ? tilesX : 0);
                for (int w = tid; w < nDense; w += nthreads) {
                    if (w >= nS) {
                        const int tl = w - nS;
                        const int r0 = (tl / tcolsX) * QCK_TILE, c0 = (tl - (tl / tcolsX) * tcolsX) * XC;
                        double2 acc[QCK_TILE][XC];
                        tile_mm<XC>(V, false, MS(QS_M), true, nc, NP, r0, c0, acc);
#pragma unroll
                        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                            for (int jj = 0; jj < XC; ++jj) XG(0)[r0 + i + NP * (c0 + jj)] = acc[i][jj];
                        continue;
                    }
                    const int pi = w / tilesS, tl = w - pi * tilesS;
                    const int r0 = (tl / tcols) * QCK_TILE, c0 = (tl - (tl / tcols) * tcols) * TC;
                    // product list: [P_0 V] [T^d_0 V]*ndir [(T^d_0)^H M]*ndir
                    const bool adj = pi > ndir;
                    const int d = pi == 0 ? -1 : (adj ? pi - 1 - ndir : pi - 1);
                    double2 acc[QCK_TILE][TC];
                    tile_mm<TC>(d < 0 ? XP(cur) : XJ(cur, d), adj, adj ? MS(QS_M) : V, false, N, NP, r0, c0, acc);
                    int q1;
                    if (!adj) q1 = d < 0 ? QO_R : (d < nd ? QO_TA + d : QO_TH);
                    else if (pass == 0) q1 = d < nd ? QO_KA0 + d : QO_KH0;
                    else q1 = d < nd ? QO_KA1 + d : QO_KH1;
                    const int b1 = c.pl_base[q1], s1 = c.pl_stride[q1];
                    // signs: R = B U1 - F U0 -> the F pass enters with -, the B pass with +; K0 = -(dF)^H M, K1 = +(dB)^H M
                    const double sgn = pass == 0 ? -1.0 : 1.0;
                    const bool accumulate = !adj && pass == 1;
#pragma unroll
                    for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                        for (int jj = 0; jj < TC; ++jj) {
                            const int rr = r0 + i, cc = c0 + jj;
                            if (rr < N && cc < nc && b1 >= 0) {
                                const int ire = cc * n2 + rr;
                                double vr = sgn * acc[i][jj].x, vi = sgn * acc[i][jj].y;
                                if (accumulate) { vr += image[b1 + ire * s1]; vi += image[b1 + (ire + N) * s1]; }
                                image[b1 + ire * s1] = vr;
                                image[b1 + (ire + N) * s1] = vi;
                            }
                        }
                }
                if (needJ) {
                    int first = ((nDense + 31) >> 5) << 5;
                    if (first >= nthreads) first = 0;
                    const int qb = pass == 0 ? QO_ISOF : QO_ISOB;
                    for (int e = tid - first; e >= 0 && e < N * N; e += nthreads - first) {
                        const int rr = e % N, cc = e / N;
                        const double2 v = XP(cur)[rr + NP * cc];
                        const double s = pass == 0 ? -1.0 : 1.0;
                        const int k00 = rr + n2 * cc, k01 = rr + n2 * (cc + N);
                        PUT(qb, k00, s * v.x); PUT(qb, k00 + N, s * v.y); PUT(qb, k01, -s * v.y); PUT(qb, k01 + N, s * v.x);
                    }
                }
            }
            // ---- reverse sweep: scalar second derivatives of Re <M, (B U1 - F U0)> ----------------------------------------------
            if (needH) {
                const double fs = pass == 0 ? -1.0 : 1.0;
                int gb = 0;
                for (int kk = 0; kk < mdeg; ++kk) {
                    const double wk = fs * sg * c.pade_r[kk];
                    const bool trivial = kk == mdeg - 1;  // level m: P = I, tangents = 0
                    GSYNC();
                    if (!trivial && taping) {
                        const double2* tp = tape + (size_t)(kk * (1 + ndir)) * N * N;
                        tape_get(XT, tp);
                        for (int d = 0; d < ndir; ++d) tape_get(XJ(0, d), tp + (size_t)(1 + d) * N * N);
                    }
                    GSYNC();
                    // Q^j = Lam A_j (sparse: walks row b of A_j^H), Q^h = Lam A (dense) -> XJ(1, .)
                    for (int w = tid; w < tilesX; w += nthreads) {
                        const int r0 = (w / tcolsX) * QCK_TILE, c0 = (w - (w / tcolsX) * tcolsX) * XC;
                        double2 acc[QCK_TILE][XC];
                        tile_mm<XC>(XG(gb), false, MA(QA_A), false, N, NP, r0, c0, acc);
#pragma unroll
                        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
                            for (int jj = 0; jj < XC; ++jj)
                                if (r0 + i < N && c0 + jj < N) XJ(1, nd)[r0 + i + NP * (c0 + jj)] = acc[i][jj];
                    }
                    {
                        int first = ((tilesX + 31) >> 5) << 5;
                        if (first >= nthreads) first = 0;
                        for (int w = tid - first; w >= 0 && w < nd * N * N; w += nthreads - first) {
                            const int i = w / (N * N), e = w - i * N * N;
                            const int a = e % N, b = e / N;
                            const int o1 = ((i * 2 + 1) * N + b) * W;
                            double2 v = make_double2(0.0, 0.0);
                            for (int u = 0; u < W; ++u) {
                                double2 ah = ellv[o1 + u];
                                ah.y = -ah.y;
                                cfma(v, XG(gb)[a + NP * ellc[o1 + u]], ah);
                            }
                            XJ(1, i)[a + NP * b] = v;
                        }
                    }
                    GSYNC();
                    if (!trivial) {
                        for (int j = 0; j < nd; ++j) {
                            for (int i = 0; i <= j; ++i) {  // a_i x a_j: wk h (tr(Q^i T^j) + tr(Q^j T^i))
                                tr_acc(i, j, wk * h, XJ(1, i), XJ(0, j));
                                tr_acc(i, j, wk * h, XJ(1, j), XJ(0, i));
                            }
                            if (free_time) {  // a_j x h: wk (tr(Q^j P) + h tr(Q^j T^h) + tr(Q^h T^j))
                                tr_acc(j, nd, wk, XJ(1, j), XT);
                                tr_acc(j, nd, wk * h, XJ(1, j), XJ(0, nd));
                                tr_acc(j, nd, wk, XJ(1, nd), XJ(0, j));
                            }
                        }
                        if (free_time) tr_acc(nd, nd, 2.0 * wk, XJ(1, nd), XJ(0, nd));
                    } else if (free_time) {  // P = I: only tr(Q^j) remains
                        for (int j = 0; j < nd; ++j) {
                            double s = 0.0;
                            for (int a = tid; a < N; a += nthreads) s += XJ(1, j)[a + NP * a].x;
                            hacc[nd * (nd + 1) / 2 + j] += wk * s;
                        }
                    }
                    // Lam <- wk/fs * h * Lam A  (the sign fs belongs to the contraction, not to the recursion)
                    GSYNC();
                    {
                        const double lam = sg * c.pade_r[kk] * h;
                        for (int e = tid; e < N * N; e += nthreads) {
                            const int o = (e % N) + NP * (e / N);
                            const double2 q = XJ(1, nd)[o];
                            XG(gb ^ 1)[o] = make_double2(lam * q.x, lam * q.y);
                        }
                    }
                    gb ^= 1;
                }
                // hand the Gamma buffer back: the next pass writes its Gamma_0 into XG(0)
                GSYNC();
            }
        }
        if (needH) {
            // block reduction of the pair sums and placement
            GSYNC();
            double* red = reinterpret_cast<double*>(MA(1));  // every N x N work matrix is free by now
            const int npd = (ndir) * (ndir + 1) / 2;
            for (int q = 0; q < npd; ++q) {
                const double v = warp_sum(hacc[q]);
                if (lane == 0) red[q * nwarps_ + warp] = v;
            }
            GSYNC();
            if (tid < npd) {
                double v = 0.0;
                for (int w2 = 0; w2 < nwarps_; ++w2) v += red[tid * nwarps_ + w2];
                int e2 = 0, rem = tid;
                while (rem > e2) { rem -= e2 + 1; ++e2; }
                const int d = rem;
                const int qq = e2 < nd ? qo_haa(d, e2) : (d < nd ? QO_HAH + d : QO_HHH);
                if (c.pl_base[qq] >= 0) image[c.pl_base[qq]] = v;
            }
        }
        GSYNC();
#undef XP
#undef XJ
#undef XT
#undef XG
        }
        if (QCK_BULK_STORE && KIND != QK_PADE4) {
            fence_async_smem();
            GSYNC();
        }
        QCK_TICK(5);

        // ---- stage 3: write-out: contiguous copies image -> value arrays ---------------------------------------------
        write_units<(TC == QCK_TILE && CN > 0) ? 2 * CN * CN : 0>(image, segs, seghdr[vwarp], seghdr[vwarp + 1], p, t, lane, p.mask);
        if (QCK_BULK_STORE && lane == 0) bulk_commit();
        QCK_TICK(6);
        buf = next_buf;
        buf_member = next_member;
        mi += step_m;
        t += step_t;
        if (mi >= nact) { mi -= nact; ++t; }
    }
    cp_async_wait_all();
    if (QCK_BULK_STORE && lane == 0) bulk_wait_all();
#undef MA
#undef MS
#undef SEGBUF
#undef CONV
#undef CONI
#undef PUT
#undef GSYNC
}

// stand-alone pass over the derivative-integrator entries: (knot, entry) pairs over the threads, consecutive threads on the
// entries of one knot.  26 us for the 2.4 M entries of the Hadamard problem at T = 100,000 (ncu), and the form does not matter:
// one entry per lane with a warp per knot 37 us, a thread per knot over a shared-memory table 38 us (with a switch, branch-free, or
// with the operand loads of 8 entries batched: 38..42 us) -- the cost is the 8-byte stores themselves, partial writes of sectors
// the previous kernel has just streamed out (profiles/r02_column_staged.txt)
__global__ void __launch_bounds__(256) qck_aux_kernel(const QckLaunch p) {
    const long long total = p.n_knots * p.n_aux;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / p.n_aux;
        const int k = (int)(i - t * p.n_aux);
        aux_entry(p, t, k, p.c.free_time ? __ldg(p.Z + t * p.c.zdim + p.c.dt_off) : p.c.dt_fixed);
    }
}

// One warp per (knot, shared position): the lanes walk the contributors' partial columns in ascending order with stride 32,
// then a shuffle tree combines the 32 partial sums.  The order of the additions is fixed (no atomics): bitwise reproducible.
__global__ void qck_reduce_kernel(const QckReduce r, double* __restrict__ H, const double* __restrict__ partial,
                                  long long n_knots, long long nnzH, int npart) {
    const int lane = threadIdx.x & 31;
    const long long total = n_knots * r.n_shared;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += nwarps) {
        const long long t = i / r.n_shared;
        const int s = (int)(i - t * r.n_shared);
        const int k0 = r.ptr[s], k1 = r.ptr[s + 1];
        double acc = 0.0;
        for (int k = k0 + lane; k < k1; k += 32) acc += partial[t * npart + r.cols[k]];
        acc = warp_sum(acc);
        if (lane == 0) H[t * nnzH + r.pos[s]] = acc;
    }
}

}  // namespace

// phase 1: matrices (the image is sized by the host's placement pass)
void qck_scratch_layout(QckClassDev& c) {
    c.msa = 2 * c.NP * c.NP;
    c.mss = 2 * c.NP * c.ncp;
    c.off_A = 0;
    const bool is_exp = c.kind == QCK_UNITARY_EXP || c.kind == QCK_KET_EXP;
    const bool is_paden = !is_exp && c.order != 4;
    c.off_S = c.off_A + (is_exp ? 4 + 2 * c.nd : (is_paden ? 6 + 2 * (c.nd + 1) : QA_C + c.nd)) * c.msa;
    c.off_img = c.off_S + QS_COUNT * c.mss;
}

// phase 2: image + per-member tables + staging, after img_doubles / W / kk_cap / nseg are known
void qck_smem_finalize(QckClassDev& c) {
    auto al = [](int b) { return (b + 15) & ~15; };
    const int dim = 2 * c.N * c.nc;
    const int npair = c.nd * (c.nd + 1) / 2;
    c.scratch_doubles = (c.off_img + c.img_doubles + 1) & ~1;
    c.icon_stride = c.ell_stride + npair + 1 + c.kk_cap + c.N * c.N + 1 + c.ac_cap;
    c.sm_seg = al(c.scratch_doubles * 8);
    c.seg_bytes = al((QCK_SEG_HDR / 4 + c.nseg) * (int)sizeof(QckSeg));
    c.sm_con = c.sm_seg + c.n_tbuf * c.seg_bytes;
    c.con_bytes = al((c.N * c.N + c.ell_stride + c.kk_cap + c.ac_cap) * 16 + c.icon_stride * 4);
    c.sm_stage = c.sm_con + c.n_tbuf * c.con_bytes;
    c.sm_bytes = al(c.sm_stage + (3 * dim + c.nd + 1) * 8);
}

int qck_pick_threads(const QckClassDev& c) {
    if (c.kind == QCK_UNITARY_PADE && c.order == 4 && c.N == 9) return 128;  // tensor-core path: fixed roles for four warps
    if (const char* e = getenv("QCK_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) return v; }  // tuning knob
    const int tc = (c.kind == QCK_UNITARY_PADE || c.kind == QCK_UNITARY_EXP) ? QCK_TILE : 1;
    int tilesS = (c.NP / QCK_TILE) * (c.ncp / tc);
    const bool is_exp = c.kind == QCK_UNITARY_EXP || c.kind == QCK_KET_EXP || c.order != 4;  // (general-order Pade shares the exponential kernel's task shape)
    int items = is_exp ? (1 + c.nd) * (c.NP / QCK_TILE) * c.NP : (2 + 2 * c.nd) * tilesS;
    int th = ((items + 31) / 32) * 32 + (is_exp ? 0 : 32);  // Pade: one extra warp for the sparse products / scalar traces
    if (items <= 32 && !is_exp) th = 32;  // everything fits one warp: warp-sized groups, several items per CTA
    else if (th < 64) th = 64;
    if (th > 256) th = 256;
    if (c.N != 2 && c.N != 3 && c.N != 4 && c.N != 9) return th;  // generic kernel: up to 256
    return th > (is_exp ? 160 : 128) ? (is_exp ? 160 : 128) : th;
}

typedef void (*qck_kernel_t)(const QckLaunch);
template <int KIND, int TC, bool MULTI>
static qck_kernel_t kernel_for(int N) {
    switch (N) {
        case 2: return qck_quantum_kernel<KIND, TC, 2, MULTI>;
        case 3: return qck_quantum_kernel<KIND, TC, 3, MULTI>;
        case 4: return qck_quantum_kernel<KIND, TC, 4, MULTI>;
        case 9: return qck_quantum_kernel<KIND, TC, 9, MULTI>;
        default: return qck_quantum_kernel<KIND, TC, 0, MULTI>;
    }
}
template <int KIND, int TC>
static qck_kernel_t kernel_for(int N, bool multi) {
    if constexpr (KIND == QK_PADE4) {  // warp-sized groups exist for the Pade kernels only
        if (multi) return kernel_for<KIND, TC, true>(N);
    }
    return kernel_for<KIND, TC, false>(N);
}

#define QCK_MAX_FUSED_AUX 256

int qck_launch_quantum(const QckLaunch& L0, int sm_count, cudaStream_t stream, int* launches) {
    QckLaunch L = L0;
    const QckClassDev& c = L.c;
    long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    if (n_items <= 0) return 0;
    {
        bool done = false;
        int rc = qck_launch_big(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        if (c.big) return (int)cudaErrorInvalidConfiguration;  // (no other kernel can hold this class)
        rc = qck_launch_rs3(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        rc = qck_launch_expeig(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        rc = qck_launch_rowslice9(L, sm_count, stream, launches, &done);
        if (!rc && !done && c.rs3) return (int)cudaErrorInvalidConfiguration;  // the class tables were built for the row-slice kernels only
        if (rc || done) return rc;
        rc = qck_launch_colexp(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        rc = qck_launch_genexp(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        rc = qck_launch_column(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
    }
    const bool unitary = c.kind == QCK_UNITARY_PADE || c.kind == QCK_UNITARY_EXP;
    qck_kernel_t kern = nullptr;
    int tc = unitary ? QCK_TILE : 1;
    const bool is_pade = c.kind == QCK_UNITARY_PADE || c.kind == QCK_KET_PADE;
    const bool multi = is_pade && c.order == 4 && c.threads == 32;  // warp-sized groups: four independent items per CTA
    static const int use_dmma = getenv("QCK_DMMA") ? atoi(getenv("QCK_DMMA")) : 0;  // FP64 tensor-core variant of the 9-level Pade-4 path
    if (c.kind == QCK_UNITARY_PADE && c.order == 4 && c.N == 9 && !multi && use_dmma) kern = qck_quantum_kernel<QK_PADE4, QCK_TILE, 9, false, true>;
    else if (c.kind == QCK_UNITARY_PADE && c.order == 4) kern = kernel_for<QK_PADE4, QCK_TILE>(c.N, multi);
    else if (c.kind == QCK_KET_PADE && c.order == 4) kern = kernel_for<QK_PADE4, 1>(c.N, multi);
    else if (c.kind == QCK_UNITARY_PADE) kern = kernel_for<QK_PADEN, QCK_TILE>(c.N, false);
    else if (c.kind == QCK_KET_PADE) kern = kernel_for<QK_PADEN, 1>(c.N, false);
    else if (c.kind == QCK_UNITARY_EXP) kern = kernel_for<QK_EXP, QCK_TILE>(c.N, multi);
    else if (c.kind == QCK_KET_EXP) kern = kernel_for<QK_EXP, 1>(c.N, multi);
    else return (int)cudaErrorNotSupported;
    const int nact = L.member_end - L.member_begin;
    L.moff_smem = nact <= 1024 ? 1 : 0;
    L.sm_count = sm_count;
    {
        static const int stg = getenv("QCK_STAGGER_NS") ? atoi(getenv("QCK_STAGGER_NS")) : 0;
        L.stagger_ns = (unsigned)stg;
    }
    const int G = c.threads;
    const int ngroups = multi ? 4 : 1;
    // Launch geometry.  With several active members the per-member tables are double-buffered in shared memory unless the
    // group count is a multiple of the member count (then every group stays on ONE member and one buffer is enough):
    // try the single-buffer layout first, fall back to two buffers if the grid cannot be aligned.
    L.group_threads = G;
    size_t smem = 0;
    int threads = G * ngroups, per_sm = 0;
    long long grid = 0;
    (void)tc;
    for (int attempt = 0; attempt < 2; ++attempt) {
        L.c.n_tbuf = (nact > 1 && attempt == 1) ? 2 : 1;
        qck_smem_finalize(L.c);
        size_t gsmem = (size_t)L.c.sm_bytes + (size_t)L.n_aux * 3 * sizeof(double);
        gsmem = (gsmem + 15) & ~(size_t)15;
        L.group_smem = (int)gsmem;
        smem = gsmem * ngroups + (size_t)L.n_aux * sizeof(QckAux) + (L.moff_smem ? (size_t)nact * 12 + 16 : 0);
        if (L.plan && L.plan->kern == (const void*)kern && L.plan->smem == smem) {
            per_sm = L.plan->per_sm;  // found at an earlier launch of this class on this device
        } else {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QCK_MAX_DYN_SMEM);
            if (e != cudaSuccess) return (int)e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
            if (e != cudaSuccess) return (int)e;
            if (L.plan) { L.plan->kern = (const void*)kern; L.plan->smem = smem; L.plan->per_sm = per_sm; }
        }
        if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
        grid = (long long)sm_count * per_sm;
        if (c.max_ctas > 0 && grid * ngroups > c.max_ctas) grid = c.max_ctas / ngroups;  // the Hessian tape was sized for this many groups
        if (grid * ngroups > n_items) grid = (n_items + ngroups - 1) / ngroups;
        bool aligned = nact == 1;
        if (nact > 1 && grid * ngroups >= nact) {
            long long total = grid * ngroups;
            total -= total % nact;
            if (total % ngroups == 0) { grid = total / ngroups; aligned = true; }
        }
        if (aligned || attempt == 1) break;
    }
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    static const bool tim = getenv("QCK_DEBUG_TIMING") != nullptr;
    long long* d_tim = nullptr;
    if (tim) {
        cudaMalloc(&d_tim, 8 * sizeof(long long));  // (debug knob only; freed below)
        cudaMemsetAsync(d_tim, 0, 8 * sizeof(long long), stream);
        L.timing = d_tim;
    }
    if (dbg) fprintf(stderr, "[qcknot] N=%d nd=%d threads=%dx%d smem=%zu B (image %d doubles) CTAs/SM=%d grid=%lld units=%d\n", c.N, c.nd, G, ngroups, smem, c.img_doubles, per_sm, grid, c.nseg);
    kern<<<(unsigned)grid, threads, smem, stream>>>(L);
    if (tim) {
        long long h[8];
        cudaMemcpyAsync(h, d_tim, sizeof h, cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        double per = 1.0 / (double)n_items;
        fprintf(stderr, "[qcknot timing] cycles/item (thread 0): loop-top %.0f | wait+bar %.0f | stage0 %.0f | prefetch-issue %.0f | stage1 %.0f | stage2 %.0f | write-out %.0f | mask=%u\n", h[0] * per, h[1] * per, h[2] * per, h[3] * per, h[4] * per, h[5] * per, h[6] * per, L.mask);
        cudaFree(d_tim);
    }
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_fused_aux_limit(void) {
    static const int lim = getenv("QCK_FUSED_AUX_LIMIT") ? atoi(getenv("QCK_FUSED_AUX_LIMIT")) : QCK_MAX_FUSED_AUX;  // (-1: never fused)
    return lim;
}

int qck_launch_aux(const QckLaunch& L, cudaStream_t stream, int* launches) {
    if (L.n_aux == 0 || L.n_knots <= 0) return 0;
    const long long total = L.n_knots * L.n_aux;
    const long long grid = std::min<long long>((total + 255) / 256, 148 * 8);
    qck_aux_kernel<<<(unsigned)grid, 256, 0, stream>>>(L);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}

int qck_launch_reduce(const QckReduce& R, double* H, const double* partial, long long n_knots, long long nnzH,
                      int npart, cudaStream_t stream, int* launches) {
    if (R.n_shared == 0 || n_knots <= 0) return 0;
    long long total = n_knots * R.n_shared;  // one warp each
    long long grid = (total + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    qck_reduce_kernel<<<(unsigned)grid, 256, 0, stream>>>(R, H, partial, n_knots, nnzH, npart);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}
