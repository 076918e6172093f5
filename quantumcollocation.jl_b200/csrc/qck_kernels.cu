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

// acc[i][j] = sum_k opA(A)[r0+i, k] * opB(B)[k, c0+j],  k < K, 3 x TC complex register tile.
// Operands are column-major with leading dimension ld.  opX = conj-transpose when tX is set, expressed through
// runtime strides + a sign on the imaginary part so that every product of a stage runs the same instruction stream.
template <int TC, bool ZERO = true>
__device__ __forceinline__ void tile_mm(const double2* __restrict__ A, bool tA, const double2* __restrict__ B, bool tB,
                                        int K, int ld, int r0, int c0, double2 (&acc)[QCK_TILE][TC]) {
    if (ZERO) {
#pragma unroll
        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = make_double2(0.0, 0.0);
    }
    const int ar = tA ? ld : 1, ak = tA ? 1 : ld;  // A[(r0+i)*ar + k*ak]
    const int bc = tB ? 1 : ld, bk = tB ? ld : 1;  // B[(c0+j)*bc + k*bk]
    const double sa = tA ? -1.0 : 1.0, sb = tB ? -1.0 : 1.0;  // conjugation = sign of the imaginary part
    const double2* a = A + r0 * ar;
    const double2* b = B + c0 * bc;
#pragma unroll 3
    for (int k = 0; k < K; ++k) {
        double2 av[QCK_TILE], bv[TC];
#pragma unroll
        for (int i = 0; i < QCK_TILE; ++i) {
            av[i] = a[i * ar + k * ak];
            av[i].y *= sa;
        }
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            bv[j] = b[j * bc + k * bk];
            bv[j].y *= sb;
        }
#pragma unroll
        for (int i = 0; i < QCK_TILE; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) cfma(acc[i][j], av[i], bv[j]);
    }
}

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

// FP64 tensor-core tile product: D(8x8) += A(8x4, row-major fragment) * B(4x8, column fragment).  Lane (g = lane/4,
// t = lane%4) holds A[g][t], B[t][g] and C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
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
// TMA bulk copy shared -> global (one thread issues; the copy engine drains the image while the CTA computes on)
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the bulk copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// Write-out: every unit is nrep back-to-back contiguous copies image -> value array, owned by ONE warp (the host
// balanced the units over the warps).  Consecutive lanes store consecutive positions with 16-byte stores; a destination
// that sits at 8 mod 16 takes a scalar head/tail and pairs shifted by one double.  Repeated (kron(I_N, .)) blocks of
// the compile-time size 2*HPC doubles are read ONCE into registers and stored nrep times (no loads, no index wrap in
// the store loop); all loops are kept free of integer division and of per-element address arithmetic.
#define QCK_BULK_STORE 1
template <int HPC>
__device__ __forceinline__ void write_units(const double* __restrict__ image, const QckSeg* __restrict__ segs, int s0, int s1,
                                            const QckLaunch& p, long long t, int lane, unsigned mask) {
    double* const baseF = p.F + t * p.c.dyn;
    double* const baseJ = p.J + t * p.nnzJ;
    double* const baseH = p.H + t * p.nnzH;
    double* const baseP = p.partial + t * p.npart - p.nnzH;
    for (int s = s0; s < s1; ++s) {
        const QckSeg sg = segs[s];
        const int arr = sg.arr & 255;
        if (!((mask >> arr) & 1u)) continue;
        double* dst = (arr == 0 ? baseF : (arr == 1 ? baseJ : ((long long)sg.dst < p.nnzH ? baseH : baseP))) + sg.dst;
        const double* src = image + (sg.img_nrep & 0xffff);
        const int nrep = sg.img_nrep >> 16, n = sg.n;
        const bool odd = (reinterpret_cast<uintptr_t>(dst) & 15) != 0;
        if (nrep == 1 && !odd && !(n & 1) && QCK_BULK_STORE) {
            if (lane == 0) bulk_store(dst, src, (unsigned)n * 8u);
        } else if (nrep == 1) {
            // plain run: scalar head (misaligned destination) / tail, 16-byte body
            const int head = odd ? 1 : 0;
            const int pairs = (n - head) >> 1;
            if (lane == 31) {
                if (head) dst[0] = src[0];
                if (head + 2 * pairs < n) dst[n - 1] = src[n - 1];
            }
            double2* d2 = reinterpret_cast<double2*>(dst + head) + lane;
            int k = lane;
            if (!head) {
                const double2* s2 = reinterpret_cast<const double2*>(src) + lane;
                for (; k + 96 < pairs; k += 128, s2 += 128, d2 += 128) {
                    const double2 v0 = s2[0], v1 = s2[32], v2 = s2[64], v3 = s2[96];
                    d2[0] = v0; d2[32] = v1; d2[64] = v2; d2[96] = v3;
                }
                for (; k < pairs; k += 32, s2 += 32, d2 += 32) *d2 = *s2;
            } else {
                const double* sh = src + 1 + 2 * lane;
                for (; k + 32 < pairs; k += 64, sh += 128, d2 += 64) {
                    const double a0 = sh[0], a1 = sh[1], b0 = sh[64], b1 = sh[65];
                    d2[0] = make_double2(a0, a1); d2[32] = make_double2(b0, b1);
                }
                for (; k < pairs; k += 32, sh += 64, d2 += 32) *d2 = make_double2(sh[0], sh[1]);
            }
        } else if (!odd && !(n & 1) && QCK_BULK_STORE) {
            if (lane == 0)
                for (int r = 0; r < nrep; ++r) bulk_store(dst + (size_t)r * n, src, (unsigned)n * 8u);
        } else if (!odd && !(n & 1)) {
            const int hp = n >> 1;
            const double2* s2 = reinterpret_cast<const double2*>(src) + lane;
            double2* d2 = reinterpret_cast<double2*>(dst) + lane;
            if (HPC > 0 && hp == HPC) {
                constexpr int NV = HPC > 0 ? (HPC + 31) / 32 : 1;
                double2 v[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (32 * (i + 1) <= HPC || lane + 32 * i < HPC) v[i] = s2[32 * i];
                for (int r = 0; r < nrep; ++r, d2 += HPC) {
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        if (32 * (i + 1) <= HPC || lane + 32 * i < HPC) d2[32 * i] = v[i];
                }
            } else {
                for (int r = 0; r < nrep; ++r, d2 += hp) {
#pragma unroll 2
                    for (int k = lane; k < hp; k += 32) d2[k - lane] = s2[k - lane];
                }
            }
        } else {
            const int total = n * nrep, step = 32 % n;  // rare path (odd period or misaligned repeated block)
            int k = lane % n;
            for (int idx = lane; idx < total; idx += 32) {
                dst[idx] = src[k];
                k += step;
                if (k >= n) k -= n;
            }
        }
    }
}

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
        if (taping && sq > c.tape_levels) sq = c.tape_levels;  // (never for ||h A||_1 <= 2^(tape_levels-4))
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
//     a_i x a_j = h^2/12 sum (Re<z1_i, A_j d> + Re<z1_j, A_i d>)          (sums over rows and columns = one warp reduction)
// 13 + 4 n_d row-slice products per knot.  Values go into the warp's own output image (same host placement and write-out
// units as the tiled kernel), which the warp then copies out.
// ------------------------------------------------------------------------------------------------------------
// WC: compile-time width of the sparse rows of the drives (loops fully unrolled); 0 = dense drive matrices
// AH: A is anti-Hermitian (Hermitian Hamiltonians): A^H x = -(A x) runs on the register-resident rows of A
template <int ND, int WC, bool AH>
__global__ void __launch_bounds__(256, 1) qck_rowslice9_kernel(const QckLaunch p) {
    constexpr int N = 9, NN = 81, n2 = 18, dim = 162;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool act = lane < 3 * N;
    const int cc = act ? lane / 3 : 0, k3 = act ? 3 * (lane - 3 * (lane / 3)) : 0;  // column, first row of this lane
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
    const int hoff = p.hoff, stage_doubles = hoff > c.img_doubles - hoff ? hoff : c.img_doubles - hoff;
    const int img_bytes = ((stage_doubles + 1) & ~1) * 8;
    double2* const cAj = reinterpret_cast<double2*>(segtab + nrec);  // WC == 0: dense A_j, row-major
    unsigned char* const wbase = reinterpret_cast<unsigned char*>(cAj + (WC > 0 ? 0 : ND * NN)) + (size_t)warp * (img_bytes + 8 * NN * 16);
    double* const imgJ = reinterpret_cast<double*>(wbase);
    double* const imgH = imgJ - hoff;
    double2* const vD = reinterpret_cast<double2*>(wbase + img_bytes);  // columns of D = U1 - U0: element [c * 9 + r]
    double2* const vS = vD + NN;       // S = U1 + U0
    double2* const vM = vS + NN;       // multipliers
    double2* const vX2 = vM + NN;      // A D
    double2* const vW1 = vX2 + NN;     // A^H M
    double2* const vU = vW1 + NN;      // A_j D (current drive)
    double2* const vZ1 = vU + NN;      // A_j^H M (current drive)
    double2* const mA = vZ1 + NN;      // A, row-major (for A^H products and column access)
    {
        const double2* gv = c.cmat + (size_t)m * c.cmat_stride;
        const int* gc = c.ell_col + (size_t)m * c.icon_stride;
        for (int e = threadIdx.x; e < nconv; e += blockDim.x) conv[e] = gv[e];
        for (int e = threadIdx.x; e < c.icon_stride; e += blockDim.x) coni[e] = gc[e];
        const QckSeg* gs = c.segs + (size_t)m * nrec;
        for (int i = threadIdx.x; i < nrec; i += blockDim.x) segtab[i] = gs[i];
        for (int i = lane; i < img_bytes / 8; i += 32) imgJ[i] = 0.0;
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
    const int* seghdr = reinterpret_cast<const int*>(segtab);
    const QckSeg* segs = segtab + QCK_SEG_HDR / 4;
    const int soff = p.moff_global[0], coff = p.moff_global[1], roff = p.moff_global[2];
    const int xo = cc * N;  // this lane's column inside the vector buffers

    for (long long t = (long long)blockIdx.x * nwarps + warp; t < p.n_knots; t += (long long)gridDim.x * nwarps) {
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
                const double aj = __ldg(zt + coff + acj[u]);  // (L1 hit: the controls were just loaded)
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

        double s_hh = 0.0, s_ah[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) s_ah[j] = 0.0;
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
            if (QCK_BULK_STORE) {
                if (lane == 0) bulk_wait_read();  // the copy engine has finished reading the previous knot's staging buffer
                __syncwarp();
            }
            // ---- phase 1: residual and Jacobian values ---------------------------------------------------------------------------
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 d = vD[xo + k3 + i];
                put(imgJ, QO_R, i, make_double2(d.x - c1h * x1[i].x + c2h2 * x3[i].x, d.y - c1h * x1[i].y + c2h2 * x3[i].y));
                put(imgJ, QO_TH, i, make_double2(-0.5 * x1[i].x + c2h * x3[i].x, -0.5 * x1[i].y + c2h * x3[i].y));
            }
        }
        if (needJ) {  // column cc of A^2 -> -iso(F), +iso(B)
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
        }
        if (needT) {
#pragma unroll
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
                for (int i = 0; i < 3; ++i) {
                    put(imgJ, QO_TA + j, i, make_double2(y[i].x + c2h2 * y3[i].x, y[i].y + c2h2 * y3[i].y));
                    if (needH && act) s_ah[j] += c2h * rdot(w1[i], u[i]);
                }
            }
        }
        if (p.n_aux) do_aux(p, t, lane, 32);  // derivative-integrator entries of this knot
        if (QCK_BULK_STORE) fence_async_smem();
        __syncwarp();
        write_units<2 * NN>(imgJ, segs, seghdr[0], seghdr[QCK_SEG_HDR - 1], p, t, lane, p.mask & (QCK_EVAL_F | QCK_EVAL_J));
        if (QCK_BULK_STORE && lane == 0) bulk_commit();
        __syncwarp();
        // ---- phase 2: Hessian-of-Lagrangian values, staged in the same buffer ------------------------------------------------------
        if (needH) {
            double2 w2[3];
            mvAH(w2, vW1 + xo);
            if (QCK_BULK_STORE) {
                if (lane == 0) bulk_wait_read();  // phase-1 copies have left the buffer
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                put(imgH, QO_KH0, i, make_double2(-0.5 * w1[i].x - c2h * w2[i].x, -0.5 * w1[i].y - c2h * w2[i].y));
                put(imgH, QO_KH1, i, make_double2(-0.5 * w1[i].x + c2h * w2[i].x, -0.5 * w1[i].y + c2h * w2[i].y));
            }
#pragma unroll
            for (int j = 0; j < ND; ++j) {
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
                    if (act) s_ah[j] += -0.5 * rdot(z1[i], vS[xo + k3 + i]) + c2h * rdot(z1[i], vX2[xo + k3 + i]);
                }
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
                for (int i = 0; i < 3; ++i) vU[(k3 + i) * N + cc] = gr[i];  // G[row][column]  (vU: its last reader ran before phase 1's flush)
            }
            __syncwarp();
            constexpr int NPAIR = ND * (ND + 1) / 2;
            const int pr = lane / 3, sub = lane - 3 * pr;
            double val = 0.0;
            if (pr < NPAIR)
                for (int u = kkptr[pr] + sub, u1 = kkptr[pr + 1]; u < u1; u += 3) {
                    const int rc = kkrc[u];
                    const double2 kv = kkv[u];
                    const double2 gv = vU[(rc & 255) * N + (rc >> 8)];  // K[r, k] G[k, r]
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
#pragma unroll
            for (int j = 0; j < ND; ++j) s_ah[j] = warp_sum(s_ah[j]);
            if (lane == 0) {
                if (c.pl_base[QO_HHH] >= 0) imgH[c.pl_base[QO_HHH]] = s_hh * (1.0 / 6.0);
#pragma unroll
                for (int j = 0; j < ND; ++j)
                    if (c.pl_base[QO_HAH + j] >= 0) imgH[c.pl_base[QO_HAH + j]] = s_ah[j];
            }
            if (QCK_BULK_STORE) fence_async_smem();
            __syncwarp();
            write_units<2 * NN>(imgH, segs, seghdr[0], seghdr[QCK_SEG_HDR - 1], p, t, lane, p.mask & QCK_EVAL_H);
            if (QCK_BULK_STORE && lane == 0) bulk_commit();
            __syncwarp();
        }
    }
    if (QCK_BULK_STORE && lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------------------
// Column kernel: Pade-4, unitaries, 2..4 levels (Hadamard / sampling problems).  ONE LANE per column of the unitaries:
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
template <int N, int ND, int NC>
__global__ void __launch_bounds__(256) qck_column_kernel(const QckLaunch p) {
    constexpr int n2 = 2 * N, blk = n2 * n2, IPW = 32 / NC, NPAIR = ND * (ND + 1) / 2;
    const QckClassDev& c = p.c;
    const int lane = threadIdx.x & 31;
    const int gi = lane / NC, col = lane - gi * NC;  // item slot inside the warp, column
    const bool needF = p.mask & QCK_EVAL_F, needJ = p.mask & QCK_EVAL_J, needH = p.mask & QCK_EVAL_H;
    const int nact = p.member_end - p.member_begin;
    const long long n_items = p.n_knots * nact;
    const long long nslots = (long long)gridDim.x * (blockDim.x >> 5) * IPW;
    const bool free_time = c.free_time;

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
        const int soff = p.moff_global[3 * mi], coff = p.moff_global[3 * mi + 1], roff = p.moff_global[3 * mi + 2];
        const double* zt = p.Z + t * c.zdim;
        const int* qd = c.qdst + (size_t)m * QO_COUNT;
        double* const oF = p.F + t * c.dyn;
        double* const oJ = p.J + t * p.nnzJ;
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
            if (d0 < p.nnzH) put_vec(p.H + t * p.nnzH, d0, q, x);
            else put_vec(p.partial + t * p.npart, d0 - (int)p.nnzH, q, x);
        };
        auto put_scalar = [&](int q, double v) {
            const int d0 = qd[q];
            if (d0 < 0 || !on || col != 0) return;
            if (d0 < p.nnzH) p.H[t * p.nnzH + d0] = v;
            else p.partial[t * p.npart + (d0 - p.nnzH)] = v;
        };
        // ---- inputs: this lane's column of U0, U1 and of the multipliers ------------------------------------------------------
        double2 d[N], s[N], mm[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const double u0r = zt[soff + col * n2 + r], u0i = zt[soff + col * n2 + N + r];
            const double u1r = zt[c.zdim + soff + col * n2 + r], u1i = zt[c.zdim + soff + col * n2 + N + r];
            d[r] = make_double2(u1r - u0r, u1i - u0i);
            s[r] = make_double2(u1r + u0r, u1i + u0i);
            mm[r] = needH ? make_double2(p.mu[t * c.dyn + roff + col * n2 + r], p.mu[t * c.dyn + roff + col * n2 + N + r]) : make_double2(0.0, 0.0);
        }
        const double h = free_time ? zt[c.dt_off] : c.dt_fixed;
        const double c1h = 0.5 * h, c2h2 = h * h * (1.0 / 12.0), c2h = h * (1.0 / 6.0);
        double a[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) a[j] = zt[coff + j];
        // ---- A = A0 + sum_j a_j A_j ---------------------------------------------------------------------------------------------
        const double2* const A0g = c.cmat + (size_t)m * c.cmat_stride;  // column-major
        const double2* const Ajg = c.dense_aj + (size_t)m * ND * N * N;  // [drive][row][column]
        double2 A[N][N];
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
                    double2 v = __ldg(A0g + k + N * col);
#pragma unroll
                    for (int j = 0; j < ND; ++j) {
                        const double2 w = __ldg(Ajg + (j * N + k) * N + col);
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
                        const double2 w = __ldg(Ajg + (j * N + r) * N + k);
                        cfma(y[r], w, vv[k]);
                        cfma(u[j][r], w, d[k]);
                    }
                }
                mvA(y3, u[j]);
                if (needJ) {
#pragma unroll
                    for (int r = 0; r < N; ++r) y[r] = make_double2(y[r].x + c2h2 * y3[r].x, y[r].y + c2h2 * y3[r].y);
                    put_J(QO_TA + j, y);
                }
            }
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
                        double2 w = __ldg(Ajg + (j * N + k) * N + r);  // conj(A_j[k][r])
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
        if (mi == 0 && p.n_aux && on) do_aux(p, t, col, NC);  // derivative-integrator entries of this knot
    }
}

__global__ void qck_aux_kernel(const QckLaunch p) {
    for (long long t = blockIdx.x; t < p.n_knots; t += gridDim.x) do_aux(p, t, threadIdx.x, blockDim.x);
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

// one warp per knot, A rows in registers (9-level Pade-4 unitaries, one active member, up to four drives)
static int launch_rowslice9(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    static const int enabled = getenv("QCK_ROWSLICE") ? atoi(getenv("QCK_ROWSLICE")) : 1;
    if (!enabled || c.kind != QCK_UNITARY_PADE || c.order != 4 || c.N != 9 || L.member_end - L.member_begin != 1 || c.nd < 1 || c.nd > 4) return 0;
    typedef void (*kern_t)(const QckLaunch);
    static const int sparse_ok = getenv("QCK_ROWSLICE_DENSE") ? 0 : 1;
    const int wc = sparse_ok && c.W <= 2 ? c.W : 0;  // sparse drive rows of width 1 or 2 are unrolled; wider ones run dense
    kern_t kern;
#define QCK_RS(ND_) (c.antiherm ? (wc == 1 ? qck_rowslice9_kernel<ND_, 1, true> : (wc == 2 ? qck_rowslice9_kernel<ND_, 2, true> : qck_rowslice9_kernel<ND_, 0, true>)) \
                                 : (wc == 1 ? qck_rowslice9_kernel<ND_, 1, false> : (wc == 2 ? qck_rowslice9_kernel<ND_, 2, false> : qck_rowslice9_kernel<ND_, 0, false>)))
    kern = c.nd == 1 ? QCK_RS(1) : (c.nd == 2 ? QCK_RS(2) : (c.nd == 3 ? QCK_RS(3) : QCK_RS(4)));
#undef QCK_RS
    const int nrec = QCK_SEG_HDR / 4 + c.nseg;
    // staging: F + J part and Hessian part of the output image share one buffer (the Hessian part starts at hoff)
    int hoff = c.img_doubles;
    for (int q = 0; q < QO_COUNT; ++q) {
        const bool hq = q == QO_KH0 || q == QO_KH1 || (q >= QO_KA0 && q < QO_ONE);
        if (hq && c.pl_base[q] >= 0 && c.pl_base[q] < hoff) hoff = c.pl_base[q];
    }
    hoff &= ~1;
    const int stage_doubles = hoff > c.img_doubles - hoff ? hoff : c.img_doubles - hoff;
    const size_t per_warp = (size_t)((stage_doubles + 1) & ~1) * 8 + 8 * 81 * 16;
    const size_t shared = ((((size_t)(81 + c.ell_stride + c.kk_cap + c.ac_cap) * 16 + (size_t)c.icon_stride * 4) + 15) & ~(size_t)15) + (size_t)nrec * 16 +
                          (wc > 0 ? 0 : (size_t)c.nd * 81 * 16);
    int nwarps = 8;
    static const int knob = getenv("QCK_ROWSLICE_WARPS") ? atoi(getenv("QCK_ROWSLICE_WARPS")) : 0;
    if (knob >= 1 && knob <= 8) nwarps = knob;
    while (nwarps > 1 && shared + nwarps * per_warp > 227 * 1024) --nwarps;
    const size_t smem = shared + nwarps * per_warp;
    if (smem > 227 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    long long grid = sm_count;
    if (grid * nwarps > L.n_knots) grid = (L.n_knots + nwarps - 1) / nwarps;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] row-slice kernel: N=9 nd=%d warps/CTA=%d smem=%zu B grid=%lld units=%d\n", c.nd, nwarps, smem, grid, c.nseg);
    QckLaunch L2 = L;
    L2.hoff = hoff;
    kern<<<(unsigned)grid, nwarps * 32, smem, stream>>>(L2);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}

// one lane per column, everything in registers (2..4-level Pade-4 unitaries, up to four drives, any number of members)
static int launch_column(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done) {
    const QckClassDev& c = L.c;
    *done = false;
    static const int enabled = getenv("QCK_COLUMN") ? atoi(getenv("QCK_COLUMN")) : 1;
    const bool ket = c.kind == QCK_KET_PADE;
    if (!enabled || (c.kind != QCK_UNITARY_PADE && !ket) || c.order != 4 || c.N < 2 || c.N > 4 || c.nd < 1 || c.nd > 4 || !c.dense_aj || !c.qdst) return 0;
    typedef void (*kern_t)(const QckLaunch);
    kern_t kern = nullptr;
#define QCK_COL2(N_, NC_) (c.nd == 1 ? qck_column_kernel<N_, 1, NC_> : (c.nd == 2 ? qck_column_kernel<N_, 2, NC_> : (c.nd == 3 ? qck_column_kernel<N_, 3, NC_> : qck_column_kernel<N_, 4, NC_>)))
#define QCK_COL(N_) (ket ? QCK_COL2(N_, 1) : QCK_COL2(N_, N_))
    kern = c.N == 2 ? QCK_COL(2) : (c.N == 3 ? QCK_COL(3) : QCK_COL(4));
#undef QCK_COL
#undef QCK_COL2
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return 0;
    const long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    const int ipw = 32 / (ket ? 1 : c.N);
    long long grid = (long long)sm_count * per_sm;
    const long long need = (n_items + 8LL * ipw - 1) / (8LL * ipw);
    if (grid > need) grid = need;
    static const bool dbg = getenv("QCK_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[qcknot] column kernel: N=%d nd=%d CTAs/SM=%d grid=%lld items=%lld\n", c.N, c.nd, per_sm, grid, n_items);
    kern<<<(unsigned)grid, 256, 0, stream>>>(L);
    if (launches) ++*launches;
    *done = true;
    return (int)cudaGetLastError();
}

int qck_launch_quantum(const QckLaunch& L0, int sm_count, cudaStream_t stream, int* launches) {
    QckLaunch L = L0;
    const QckClassDev& c = L.c;
    long long n_items = L.n_knots * (long long)(L.member_end - L.member_begin);
    if (n_items <= 0) return 0;
    {
        bool done = false;
        int rc = launch_rowslice9(L, sm_count, stream, launches, &done);
        if (rc || done) return rc;
        rc = launch_column(L, sm_count, stream, launches, &done);
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
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return (int)e;
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
    static long long* d_tim = nullptr;
    if (tim) {
        if (!d_tim) cudaMalloc(&d_tim, 8 * sizeof(long long));
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
    }
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
    long long total = n_knots * R.n_shared;  // one warp each
    long long grid = (total + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    qck_reduce_kernel<<<(unsigned)grid, 256, 0, stream>>>(R, H, partial, n_knots, nnzH, npart);
    if (launches) ++*launches;
    return (int)cudaGetLastError();
}
