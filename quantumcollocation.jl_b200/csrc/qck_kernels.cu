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
#pragma unroll 9
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
        double2 v = val[r * W + w];
        int k = col[r * W + w];
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

// same entries, operands already staged in shared memory by the prefetch (fused path: no global load latency)
__device__ __forceinline__ void do_aux_staged(const QckLaunch& p, const QckAux* auxs, const double* auxv, double dt,
                                              long long t, int tid, int nthreads) {
    const QckClassDev& c = p.c;
    for (int k = tid; k < p.n_aux; k += nthreads) {
        const QckAux a = auxs[k];
        if (!((p.mask >> a.out) & 1u)) continue;
        double v;
        switch (a.op) {
            case QAUX_CONST: v = a.c; break;
            case QAUX_NEG_DT: v = -dt; break;
            case QAUX_NEG_Z: v = -auxv[3 * k]; break;
            case QAUX_NEG_MU: v = -auxv[3 * k + 2]; break;
            default: v = auxv[3 * k + 1] - auxv[3 * k] - dt * auxv[3 * k + 2]; break;
        }
        if (a.out == 0) p.F[t * c.dyn + a.pos] = v;
        else if (a.out == 1) p.J[t * p.nnzJ + a.pos] = v;
        else if (a.pos < p.nnzH) p.H[t * p.nnzH + a.pos] = v;
        else p.partial[t * p.npart + (a.pos - p.nnzH)] = v;
    }
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ double slot_val(const double* __restrict__ sm, unsigned s) {
    double v = sm[s & 0x7fffu];
    return (s & 0x8000u) ? -v : v;
}

// Output segments: out[dst + k] = +-scratch[tab[src_off + k % period]].  Everything the loop touches except the
// destination is in shared memory; consecutive threads store consecutive positions (coalesced).  Stores are 16-byte
// (two values) whenever the destination is 16-byte aligned, 8-byte otherwise.
// A periodic segment (kron(I_N, .) block) gathers each value once and stores it len/period times.
__device__ __forceinline__ void write_segments(const double* __restrict__ sm, const uint16_t* __restrict__ tab,
                                               const QckSeg* __restrict__ segs, int nseg, double* __restrict__ out,
                                               long long limit, double* __restrict__ partial, int tid, int nthreads) {
    for (int s = 0; s < nseg; ++s) {
        const QckSeg sg = segs[s];
        double* dst = (long long)sg.dst < limit ? out + sg.dst : partial + (sg.dst - limit);
        const uint16_t* tb = tab + sg.src_off;
        const bool odd = (reinterpret_cast<uintptr_t>(dst) & 15) != 0;
        if (sg.period == sg.len) {
            // head (one value if misaligned), 16-byte body, tail
            const int head = odd ? 1 : 0;
            const int pairs = (sg.len - head) >> 1;
            if (tid == 0 && head) dst[0] = slot_val(sm, tb[0]);
            double2* d2 = reinterpret_cast<double2*>(dst + head);
            const uint16_t* t2 = tb + head;
            int k = tid;
            for (; k + nthreads < pairs; k += 2 * nthreads) {
                double a0 = slot_val(sm, t2[2 * k]), a1 = slot_val(sm, t2[2 * k + 1]);
                double b0 = slot_val(sm, t2[2 * (k + nthreads)]), b1 = slot_val(sm, t2[2 * (k + nthreads) + 1]);
                d2[k] = make_double2(a0, a1);
                d2[k + nthreads] = make_double2(b0, b1);
            }
            for (; k < pairs; k += nthreads) d2[k] = make_double2(slot_val(sm, t2[2 * k]), slot_val(sm, t2[2 * k + 1]));
            if (tid == nthreads - 1 && head + 2 * pairs < sg.len) dst[sg.len - 1] = slot_val(sm, tb[sg.len - 1]);
        } else {
            const int nrep = sg.len / sg.period;
            if (!odd && !(sg.period & 1)) {
                const int hp = sg.period >> 1;
                for (int k = tid; k < hp; k += nthreads) {
                    const double2 v = make_double2(slot_val(sm, tb[2 * k]), slot_val(sm, tb[2 * k + 1]));
                    double2* d = reinterpret_cast<double2*>(dst) + k;
                    for (int r = 0; r < nrep; ++r) d[(size_t)r * hp] = v;
                }
            } else {
                for (int k = tid; k < sg.period; k += nthreads) {
                    const double v = slot_val(sm, tb[k]);
                    double* d = dst + k;
                    for (int r = 0; r < nrep; ++r) d[(size_t)r * sg.period] = v;
                }
            }
        }
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
// 6 + 2 n_d dense N^3 products per knot (vs ~45 in the reference's real-iso formulation); products with the
// constant drive matrices A_j are sparse (ELL rows) and the cheap ones are recomputed on the fly instead of stored.
//
// Shared memory per CTA: [scratch matrices | slot table | per-member segments x2 | per-member ELL x2 | input staging].
// The inputs (and, when the member changes, the tables) of the NEXT work item are prefetched with cp.async while
// the current item computes, so no global-load latency sits on the per-item critical path.
// ------------------------------------------------------------------------------------------------------------
template <int TC, int CN>
__global__ void __launch_bounds__(CN == 0 ? 256 : 128, CN == 0 ? 1 : 4) qck_pade4_kernel(const QckLaunch p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const QckClassDev& c = p.c;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    // CN > 0: levels known at compile time (index arithmetic folds, k-loops unroll); CN == 0: generic fallback
    const int N = CN > 0 ? CN : c.N;
    const int NP = CN > 0 ? ((CN + QCK_TILE - 1) / QCK_TILE) * QCK_TILE : c.NP;
    const int nc = TC == 1 ? 1 : N, ncp = TC == 1 ? 1 : NP;
    const int nd = c.nd, W = c.W;
    const bool needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H, needF = p.mask & QCK_EVAL_F;
    const bool needT = needJ || needH;  // first-order drive terms
    const bool free_time = c.free_time;
    const int dim = 2 * N * nc;
    // auxiliary (derivative-integrator) entries live behind the class's shared-memory carve-up
    QckAux* auxs = reinterpret_cast<QckAux*>(smem_raw + c.sm_bytes);
    double* auxv = reinterpret_cast<double*>(smem_raw + c.sm_bytes + p.n_aux * (int)sizeof(QckAux));

    double2* SA = reinterpret_cast<double2*>(sm + c.off_A);
    double2* SS = reinterpret_cast<double2*>(sm + c.off_S);
    double* SX = sm + c.off_X;
    uint16_t* tab = reinterpret_cast<uint16_t*>(smem_raw + c.sm_tab);
    double* stage = reinterpret_cast<double*>(smem_raw + c.sm_stage);  // [z_t state | z_t+1 state | mu | a | h]
    const int nseg = c.nsegJ + c.nsegH;
    const int elln = c.ell_stride;
    const int msa = NP * NP, mss = NP * ncp;  // complex elements per matrix
#define MA(i) (SA + (i) * msa)
#define MS(i) (SS + (i) * mss)
#define MD(j, k) (SS + (QS_FIXED + QD_COUNT * (j) + (k)) * mss)
#define SEGBUF(b) reinterpret_cast<QckSeg*>(smem_raw + c.sm_seg + (b) * c.seg_bytes)
#define ELLV(b) reinterpret_cast<double2*>(smem_raw + c.sm_ell + (b) * c.ell_bytes)
#define ELLC(b) reinterpret_cast<int*>(smem_raw + c.sm_ell + (b) * c.ell_bytes + elln * 16)

    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    const int tilesS = (NP / QCK_TILE) * (ncp / TC);
    const int tilesA = (NP / QCK_TILE) * (NP / QCK_TILE);
    const int tcols = ncp / TC;

    // prefetch of one work item: inputs into `stage`, and the member's segment + ELL tables into buffer b if asked
    auto prefetch = [&](long long item, int b, bool tables) {
        const long long t = item / nact;
        const int m = p.member_begin + (int)(item - t * nact);
        const double* zt = p.Z + t * c.zdim;
        const int soff = __ldg(c.state_off + m), coff = __ldg(c.ctrl_off + m), roff_n = __ldg(c.row_off + m);
        for (int i = tid; i < dim; i += nthreads) {
            cp_async8(stage + i, zt + soff + i);
            cp_async8(stage + dim + i, zt + c.zdim + soff + i);
            if (needH) cp_async8(stage + 2 * dim + i, p.mu + t * c.dyn + roff_n + i);
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
            const QckSeg* gs = c.segs + (size_t)m * nseg;
            for (int i = tid; i < nseg; i += nthreads) cp_async16(SEGBUF(b) + i, gs + i);
            const double2* gv = c.cmat + (size_t)m * c.cmat_stride + N * N * (1 + nd);
            const int* gc = c.ell_col + (size_t)m * c.ell_stride;
            for (int i = tid; i < elln; i += nthreads) { cp_async16(ELLV(b) + i, gv + i); cp_async4(ELLC(b) + i, gc + i); }
        }
        cp_async_commit();
    };

    for (int i = tid; i < c.scratch_doubles; i += nthreads) sm[i] = 0.0;
    for (int i = tid; i < c.tab_len; i += nthreads) tab[i] = c.tab[i];
    for (int i = tid; i < p.n_aux; i += nthreads) auxs[i] = p.aux[i];
    __syncthreads();
    int buf = 0, buf_member = -1;
    if ((long long)blockIdx.x < n_items) {
        buf_member = p.member_begin + (int)(blockIdx.x % nact);
        prefetch(blockIdx.x, buf, true);
    }

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long t = item / nact;
        const int mi = (int)(item - t * nact);
        const int m = p.member_begin + mi;
        cp_async_wait_all();
        __syncthreads();  // staged inputs visible; previous item's write-out has finished reading scratch
        if (tid == 0) SX[QX_ONE] = 1.0;
        const int roff = __ldg(c.row_off + m);
        const double2* cm = c.cmat + (size_t)m * c.cmat_stride;
        const double2* A0 = cm;
        const double2* Adr = cm + N * N;
        const double2* ellv = ELLV(buf);
        const int* ellc = ELLC(buf);
        const QckSeg* segs = SEGBUF(buf);
        const double h = free_time ? stage[3 * dim + nd] : c.dt_fixed;

        // ---- stage 0: unpack the staged knot pair, build A = -i H(a) and A^H ---------------------------------
        {
            for (int idx = tid; idx < dim; idx += nthreads) {
                int cc = idx / (2 * N), q = idx - cc * 2 * N;
                int im = q >= N, r = q - im * N;
                double u0 = stage[idx], u1 = stage[dim + idx];
                int o = 2 * (r + NP * cc) + im;
                reinterpret_cast<double*>(MS(QS_D))[o] = u1 - u0;
                reinterpret_cast<double*>(MS(QS_S))[o] = u1 + u0;
                if (needH) reinterpret_cast<double*>(MS(QS_M))[o] = stage[2 * dim + idx];
            }
            for (int e = tid; e < N * N; e += nthreads) {
                int r = e % N, k = e / N;
                double2 v = __ldg(A0 + e);
                for (int j = 0; j < nd; ++j) {
                    double aj = stage[3 * dim + j];
                    double2 d = __ldg(Adr + j * N * N + e);
                    v.x = fma(aj, d.x, v.x);
                    v.y = fma(aj, d.y, v.y);
                }
                MA(QA_A)[r + NP * k] = v;
                MA(QA_AH)[k + NP * r] = make_double2(v.x, -v.y);
            }
        }
        __syncthreads();
        // staging is free again: fetch the next item's inputs (and tables, if its member differs) behind the compute
        int next_buf = buf, next_member = buf_member;
        {
            const long long nitem = item + gridDim.x;
            if (nitem < n_items) {
                const int nm = p.member_begin + (int)(nitem % nact);
                const bool tables = nm != buf_member;
                if (tables) { next_buf = buf ^ 1; next_member = nm; }  // several active members => two table buffers
                prefetch(nitem, next_buf, tables);
            }
        }

        // ---- stage 1: A2, A D, A S, A^H M (dense);  Q1_j = A_j D, N1_j = A_j^H M (sparse) -----------------------
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
                const int nsp = (needH ? 2 : 1) * nd * per;
                for (int w = nthreads - 1 - tid; w < nsp; w += nthreads) {
                    int pj = w / per, e = w - pj * per;
                    int adj = pj / nd, j = pj - adj * nd;
                    int r = e % N, cc = e / N;
                    const double2* X = adj ? MS(QS_M) : MS(QS_D);
                    double2 v = ell_row(ellv + (j * 2 + adj) * N * W, ellc + (j * 2 + adj) * N * W, W, X, NP, r, cc);
                    MD(j, adj ? QD_N1 : QD_Q1)[r + NP * cc] = v;
                }
            }
        }
        __syncthreads();

        // ---- stage 2: A(AD), A^H(A^H M), A(A_j D), A^H(A_j^H M) (dense) --------------------------------------------
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
                    // Re <M, -1/2 A_j S + h/6 (A_j (A D) + A (A_j D))>, the two sparse products recomputed on the fly
                    int j = task - npair;
                    s = 0.0;
                    for (int e = lane; e < N * nc; e += 32) {
                        int r = e % N, cc = e / N;
                        double2 mm = MS(QS_M)[r + NP * cc];
                        double2 pj = ell_row(ellv + (j * 2) * N * W, ellc + (j * 2) * N * W, W, MS(QS_S), NP, r, cc);
                        double2 q2 = ell_row(ellv + (j * 2) * N * W, ellc + (j * 2) * N * W, W, MS(QS_AD), NP, r, cc);
                        double2 aq = MD(j, QD_AQ1)[r + NP * cc];
                        double vr = -0.5 * pj.x + c2h * (q2.x + aq.x), vi = -0.5 * pj.y + c2h * (q2.y + aq.y);
                        s = fma(mm.x, vr, s);
                        s = fma(mm.y, vi, s);
                    }
                    slot = QX_HAH + j;
                } else {
                    s = (1.0 / 6.0) * re_dot(MS(QS_M), MS(QS_AAD), N, nc, NP, lane);
                    slot = qx_hhh(nd);
                }
                s = warp_sum(s);
                if (lane == 0) SX[slot] = s;
            }
            __syncthreads();
        }

        // ---- stage 3b-1: per-drive outputs (read S, AD, AhM; overwrite only their own element) and F, B ----------
        {
            const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0);
            const int per = N * nc;
            if (needT) {
                for (int w = tid; w < nd * per; w += nthreads) {
                    int j = w / per, e = w - j * per;
                    int r = e % N, cc = e / N;
                    int o = r + NP * cc;
                    const double2* ev = ellv + (j * 2) * N * W;
                    const int* ec = ellc + (j * 2) * N * W;
                    double2 pj = ell_row(ev, ec, W, MS(QS_S), NP, r, cc);
                    double2 q2 = ell_row(ev, ec, W, MS(QS_AD), NP, r, cc);
                    double2 aq1 = MD(j, QD_AQ1)[o];
                    MD(j, QD_TA)[o] = make_double2(-c1h * pj.x + c2h2 * (q2.x + aq1.x), -c1h * pj.y + c2h2 * (q2.y + aq1.y));
                    if (needH) {
                        double2 n1 = MD(j, QD_N1)[o], ahn1 = MD(j, QD_AHN1)[o];
                        double2 n2 = ell_row(ev + N * W, ec + N * W, W, MS(QS_AHM), NP, r, cc);
                        double xr = n2.x + ahn1.x, xi = n2.y + ahn1.y;
                        MD(j, QD_KA0)[o] = make_double2(-(c1h * n1.x + c2h2 * xr), -(c1h * n1.y + c2h2 * xi));
                        MD(j, QD_KA1)[o] = make_double2(-c1h * n1.x + c2h2 * xr, -c1h * n1.y + c2h2 * xi);
                    }
                }
            }
            if (needJ) {
                for (int e = nthreads - 1 - tid; e < N * N; e += nthreads) {
                    int r = e % N, k = e / N;
                    double2 a = MA(QA_A)[r + NP * k], a2 = MA(QA_A2)[r + NP * k];
                    double id = r == k ? 1.0 : 0.0;
                    MA(QA_F)[r + NP * k] = make_double2(id + c1h * a.x + c2h2 * a2.x, c1h * a.y + c2h2 * a2.y);
                    MA(QA_B)[r + NP * k] = make_double2(id - c1h * a.x + c2h2 * a2.x, -c1h * a.y + c2h2 * a2.y);
                }
            }
        }
        __syncthreads();

        // ---- stage 3b-2: fixed outputs in place: R, dR/dh, and the dt Hessian blocks -----------------------------------
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
        }
        __syncthreads();

        // ---- stage 4: write-out ------------------------------------------------------------------------------------
        if (needF) {
            double* Fo = p.F + t * c.dyn + roff;
            const double* R = reinterpret_cast<const double*>(MS(QS_D));
            for (int idx = tid; idx < dim; idx += nthreads) {
                int cc = idx / (2 * N), q = idx - cc * 2 * N;
                int im = q >= N, r = q - im * N;
                Fo[idx] = R[2 * (r + NP * cc) + im];
            }
        }
        if (needJ) write_segments(sm, tab, segs, c.nsegJ, p.J + t * p.nnzJ, (long long)1 << 60, nullptr, tid, nthreads);
        if (needH)
            write_segments(sm, tab, segs + c.nsegJ, c.nsegH, p.H + t * p.nnzH, p.nnzH, p.partial + t * p.npart, tid, nthreads);
        if (mi == 0 && p.n_aux) do_aux_staged(p, auxs, auxv, h, t, tid, nthreads);
        buf = next_buf;
        buf_member = next_member;
    }
    cp_async_wait_all();
#undef MA
#undef MS
#undef MD
#undef SEGBUF
#undef ELLV
#undef ELLC
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
void qck_scratch_layout(QckClassDev& c) {
    c.msa = 2 * c.NP * c.NP;
    c.mss = 2 * c.NP * c.ncp;
    c.off_A = 0;
    int n_s = QS_FIXED + QD_COUNT * c.nd;
    c.off_S = c.off_A + QA_COUNT * c.msa;
    c.off_X = c.off_S + n_s * c.mss;
    c.scratch_doubles = c.off_X + qx_haa(c.nd, 0, 0) + c.nd * c.nd;
    c.scratch_doubles = (c.scratch_doubles + 1) & ~1;
}

void qck_smem_finalize(QckClassDev& c) {
    auto al = [](int b) { return (b + 15) & ~15; };
    const int dim = 2 * c.N * c.nc;
    c.sm_tab = al(c.scratch_doubles * 8);
    c.sm_seg = al(c.sm_tab + c.tab_len * 2);
    c.seg_bytes = al((c.nsegJ + c.nsegH) * (int)sizeof(QckSeg));
    c.sm_ell = c.sm_seg + 2 * c.seg_bytes;
    c.ell_bytes = al(c.ell_stride * 16 + c.ell_stride * 4);
    c.sm_stage = c.sm_ell + 2 * c.ell_bytes;
    c.sm_bytes = al(c.sm_stage + (3 * dim + c.nd + 1) * 8);
}

static int pick_threads(const QckClassDev& c, int tc) {
    int tilesS = (c.NP / QCK_TILE) * (c.ncp / tc);
    int items = (2 + 2 * c.nd) * tilesS;
    int th = ((items + 31) / 32) * 32;
    if (th < 64) th = 64;
    if (th > 256) th = 256;
    if (c.N >= 8 && th < 128) th = 128;  // the write-out and assembly phases want the extra warp
    return th;
}

typedef void (*qck_kernel_t)(const QckLaunch);
template <int TC>
static qck_kernel_t pade4_for(int N) {
    switch (N) {
        case 2: return qck_pade4_kernel<TC, 2>;
        case 3: return qck_pade4_kernel<TC, 3>;
        case 4: return qck_pade4_kernel<TC, 4>;
        case 5: return qck_pade4_kernel<TC, 5>;
        case 6: return qck_pade4_kernel<TC, 6>;
        case 8: return qck_pade4_kernel<TC, 8>;
        case 9: return qck_pade4_kernel<TC, 9>;
        default: return qck_pade4_kernel<TC, 0>;
    }
}

#define QCK_MAX_FUSED_AUX 256

int qck_launch_quantum(const QckLaunch& L0, int sm_count, cudaStream_t stream, int* launches) {
    QckLaunch L = L0;
    const QckClassDev& c = L.c;
    long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    if (n_items <= 0) return 0;
    const bool unitary = c.kind == QCK_UNITARY_PADE || c.kind == QCK_UNITARY_EXP;
    qck_kernel_t kern = nullptr;
    int tc = unitary ? QCK_TILE : 1;
    if (c.kind == QCK_UNITARY_PADE && c.order == 4) kern = pade4_for<QCK_TILE>(c.N);
    else if (c.kind == QCK_KET_PADE && c.order == 4) kern = pade4_for<1>(c.N);
    else return (int)cudaErrorNotSupported;
    size_t smem = (size_t)c.sm_bytes + (size_t)L.n_aux * (sizeof(QckAux) + 3 * sizeof(double));
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

int qck_fused_aux_limit(void) { return QCK_MAX_FUSED_AUX; }

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
