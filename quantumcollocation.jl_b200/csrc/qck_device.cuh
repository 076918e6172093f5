// Device helpers shared by the kernel translation units of libqcknot.so (complex FMA, small tile products, cp.async / TMA
// bulk-copy wrappers, derivative-integrator entries, the write-out of the output image).  Internal: not part of the C-ABI.
#pragma once
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

// one derivative-integrator entry k of knot t (read-only loads, ld.global.nc: independent entries do not wait for each other's stores)
__device__ __forceinline__ void aux_entry(const QckLaunch& p, long long t, int k, double dt) {
    const QckClassDev& c = p.c;
    const double* zt = p.Z + t * c.zdim;
    const int4 a0 = __ldg(reinterpret_cast<const int4*>(p.aux + k));  // out, op, pos, i0
    const int out = a0.x, op = a0.y, pos = a0.z, i0 = a0.w;
    if (!((p.mask >> out) & 1u)) return;
    double v;
    switch (op) {
        case QAUX_CONST: v = __ldg(&p.aux[k].c); break;
        case QAUX_NEG_DT: v = -dt; break;
        case QAUX_NEG_Z: v = -__ldg(zt + i0); break;
        case QAUX_NEG_MU: v = -__ldg(p.mu + t * c.dyn + i0); break;
        default: v = __ldg(zt + c.zdim + i0) - __ldg(zt + i0) - dt * __ldg(zt + __ldg(&p.aux[k].i1)); break;
    }
    if (out == 0) p.F[t * c.dyn + pos] = v;
    else if (out == 1) p.J[t * p.nnzJ + pos] = v;
    else if (pos < p.nnzH) p.H[t * p.nnzH + pos] = v;
    else p.partial[t * p.npart + (pos - p.nnzH)] = v;
}

__device__ __forceinline__ void do_aux(const QckLaunch& p, long long t, int tid, int nthreads) {
    const QckClassDev& c = p.c;
    const double dt = c.free_time ? __ldg(p.Z + t * c.zdim + c.dt_off) : c.dt_fixed;
#pragma unroll 4
    for (int k = tid; k < p.n_aux; k += nthreads) aux_entry(p, t, k, dt);
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
// QCK_STORE_HINT (compile-time A/B knob): L2 eviction policy of the value-array stores; 0 = none, 1 = evict-first (default),
// 2 = evict-last, 3 = evict-unchanged.  The value arrays are written once and never read back by the kernels, and at T = 10,000 a
// pass writes 5x the L2: with evict-first the written lines leave L2 in the order they arrive instead of competing with the inputs
// and the staging traffic.  Measured on the CZ problem (row-slice kernel, profiles/r02_store_hint.txt): 199.2 -> 188.3 us at
// T = 10,000, 1,892 -> 1,778 us at T = 100,000; evict-last 202.8 us, evict-unchanged 199.4 us.
#ifndef QCK_STORE_HINT
#define QCK_STORE_HINT 1
#endif
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
#if QCK_STORE_HINT
    unsigned long long pol;
#if QCK_STORE_HINT == 1
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
#elif QCK_STORE_HINT == 2
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;\n" : "=l"(pol));
#endif
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(gdst), "r"(s), "r"(bytes), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the bulk copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// same entries from a shared-memory copy of the table (row-slice kernel): the knot's inputs are L1 / L2 hits by now
__device__ __forceinline__ void do_aux_smem(const QckLaunch& p, const QckAux* auxs, long long t, int tid, int nthreads) {
    const QckClassDev& c = p.c;
    const double* zt = p.Z + t * c.zdim;
    const double dt = c.free_time ? __ldg(zt + c.dt_off) : c.dt_fixed;
    for (int k = tid; k < p.n_aux; k += nthreads) {
        const QckAux a = auxs[k];
        if (!((p.mask >> a.out) & 1u)) continue;
        double v;
        switch (a.op) {
            case QAUX_CONST: v = a.c; break;
            case QAUX_NEG_DT: v = -dt; break;
            case QAUX_NEG_Z: v = -__ldg(zt + a.i0); break;
            case QAUX_NEG_MU: v = -__ldg(p.mu + t * c.dyn + a.i0); break;
            default: v = __ldg(zt + c.zdim + a.i0) - __ldg(zt + a.i0) - dt * __ldg(zt + a.i1); break;
        }
        if (a.out == 0) p.F[t * c.dyn + a.pos] = v;
        else if (a.out == 1) p.J[t * p.nnzJ + a.pos] = v;
        else if (a.pos < p.nnzH) p.H[t * p.nnzH + a.pos] = v;
        else p.partial[t * p.npart + (a.pos - p.nnzH)] = v;
    }
}

// Lane-parallel write-out for images placed with destination parity (qck_host.cpp: place_array(parity), rs3_units): one lane
// = one unit = nrep back-to-back copies of [scalar head] + 16-byte aligned TMA bulk copy + [scalar tail].  `stage` holds the
// F + J part of the image at offset 0 (shifted by shF / shJ doubles) or the Hessian part (image offsets >= hoff, shifted by
// shH); a shift is the parity of the array's knot-block base address, so source and destination of a unit always agree mod 16.
__device__ __forceinline__ void flush_units_lanes(const double* __restrict__ stage, const QckSeg* __restrict__ units, int u0, int u1, int lane,
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

// Lane-parallel variant for a warp that writes a whole knot on its own (row-slice kernel): lane u decodes unit u; units whose
// copies are 16-byte aligned leave through the copy engine right from that lane (every issuing lane commits / waits for its
// own bulk group), the others are broadcast one by one and copied by the whole warp.
__device__ __forceinline__ void write_units_lanes(const double* __restrict__ image, const QckSeg* __restrict__ segs, int s0, int s1,
                                                  const QckLaunch& p, long long t, int lane, unsigned mask) {
    double* const baseF = p.F + t * p.c.dyn;
    double* const baseJ = p.J + t * p.nnzJ;
    double* const baseH = p.H + t * p.nnzH;
    double* const baseP = p.partial + t * p.npart - p.nnzH;
    for (int sb = s0; sb < s1; sb += 32) {
        const int s = sb + lane;
        bool todo = false;
        double* dst = nullptr;
        int img = 0, n = 0, nrep = 0;
        if (s < s1) {
            const QckSeg sg = segs[s];
            const int arr = sg.arr & 255;
            if ((mask >> arr) & 1u) {
                dst = (arr == 0 ? baseF : (arr == 1 ? baseJ : ((long long)sg.dst < p.nnzH ? baseH : baseP))) + sg.dst;
                img = sg.img_nrep & 0xffff; nrep = sg.img_nrep >> 16; n = sg.n;
                if (!(reinterpret_cast<uintptr_t>(dst) & 15) && !(n & 1)) {
                    for (int r = 0; r < nrep; ++r) bulk_store(dst + (size_t)r * n, image + img, (unsigned)n * 8u);
                } else {
                    todo = true;
                }
            }
        }
        unsigned rem = __ballot_sync(0xffffffffu, todo);
        while (rem) {
            const int l = __ffs(rem) - 1;
            rem &= rem - 1;
            double* const d = reinterpret_cast<double*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), l));
            const double* const src = image + __shfl_sync(0xffffffffu, img, l);
            const int nn = __shfl_sync(0xffffffffu, n, l), nr = __shfl_sync(0xffffffffu, nrep, l);
            if (nr == 1) {  // plain run: scalar head (misaligned destination) / tail, 16-byte body
                const int head = (reinterpret_cast<uintptr_t>(d) & 15) ? 1 : 0;
                const int pairs = (nn - head) >> 1;
                if (lane == 31) {
                    if (head) d[0] = src[0];
                    if (head + 2 * pairs < nn) d[nn - 1] = src[nn - 1];
                }
                double2* d2 = reinterpret_cast<double2*>(d + head) + lane;
                int k = lane;
                if (!head) {
                    const double2* s2 = reinterpret_cast<const double2*>(src) + lane;
                    for (; k < pairs; k += 32, s2 += 32, d2 += 32) *d2 = *s2;
                } else {
                    const double* sh = src + 1 + 2 * lane;
                    for (; k + 32 < pairs; k += 64, sh += 128, d2 += 64) {
                        const double a0 = sh[0], a1 = sh[1], b0 = sh[64], b1 = sh[65];
                        d2[0] = make_double2(a0, a1); d2[32] = make_double2(b0, b1);
                    }
                    for (; k < pairs; k += 32, sh += 64, d2 += 32) *d2 = make_double2(sh[0], sh[1]);
                }
            } else {  // rare: odd period or misaligned repeated block
                const int total = nn * nr, step = 32 % nn;
                int k = lane % nn;
                for (int idx = lane; idx < total; idx += 32) {
                    d[idx] = src[k];
                    k += step;
                    if (k >= nn) k -= nn;
                }
            }
        }
    }
}

}  // namespace
