// Internal structures shared by the host side (qck_host.cpp) and the kernels (qck_kernels.cu) of libqcknot.so.
// Nothing here is part of the C-ABI (include/qcknot.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Every C-ABI entry point that switches the CUDA device puts one of these on its stack: a library behind somebody else's
// runtime (CUDA.jl, torch) must not leave the calling thread on another device.
struct QckDeviceScope {
    int prev = -1;
    QckDeviceScope() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            (void)cudaGetLastError();
        }
    }
    ~QckDeviceScope() {
        if (prev >= 0) (void)cudaSetDevice(prev);
    }
    QckDeviceScope(const QckDeviceScope&) = delete;
    QckDeviceScope& operator=(const QckDeviceScope&) = delete;
};

#include "../../include/qcknot.h"

#define QCK_TILE 3          // register tile edge of the small complex products (3x3 complex per thread)
#define QCK_MAX_DRIVES 6
#define QCK_MAX_OBJ_TERMS 12
#define QCK_MAX_PADE_M 5    // Pade order <= 10
enum { QK_PADE4 = 0, QK_EXP = 1, QK_PADEN = 2 };  // kernel families

// ---- shared-memory scratch of the quantum kernels -----------------------------------------------------------------
// "A-type" matrices are NP x NP complex, "state-type" are NP x ncp complex (ncp = NP for unitaries, 1 for kets),
// both column-major with leading dimension NP (NP = N rounded up to the register tile).
// Pade-4 kernel: A, A2, G = D M^H, G2 = S M^H, C_j = {A_j, A}  |  D, S, M, AS, AhM
enum { QA_A = 0, QA_A2 = 1, QA_G = 2, QA_G2 = 3, QA_C = 4 /* + j */ };
enum { QS_D = 0, QS_S = 1, QS_M = 2, QS_AS = 3, QS_AHM = 4, QS_COUNT = 5 };

// ---- output quantities ---------------------------------------------------------------------------------------------------
// Every value a quantum integrator writes belongs to one of these quantities.  The host decides where each quantity
// lives inside the CTA's shared-memory OUTPUT IMAGE (base + element * stride) such that the image is laid out in the
// solver's structure order; the kernel's epilogues write values there and the write-out is a plain contiguous copy.
//   iso blocks  (2N x 2N real, column-major; stored once, written N times: kron(I_N, .)):  ISOF = -iso(F|E), ISOB = +iso(B)
//   iso-vectors (element i of vec_iso(X), i = cc*2N + q):  R residual, TH d/ddt, TA_j d/da_j,
//                KH0 (state_t x dt), KH1 (dt x state_t+1), KA0_j (state_t x a_j), KA1_j (a_j x state_t+1)
//   scalars:     HHH (dt x dt), HAH_j (a_j x dt), HAA_ij (a_i x a_j, i <= j), ONE (constant 1: identity Jacobian block)
enum {
    QO_ISOF = 0, QO_ISOB = 1, QO_R = 2, QO_TH = 3, QO_KH0 = 4, QO_KH1 = 5,
    QO_TA = 6, QO_KA0 = QO_TA + QCK_MAX_DRIVES, QO_KA1 = QO_KA0 + QCK_MAX_DRIVES,
    QO_HHH = QO_KA1 + QCK_MAX_DRIVES, QO_HAH = QO_HHH + 1, QO_HAA = QO_HAH + QCK_MAX_DRIVES,
    QO_ONE = QO_HAA + QCK_MAX_DRIVES * QCK_MAX_DRIVES, QO_COUNT = QO_ONE + 1
};
static inline __host__ __device__ int qo_haa(int i, int j) { return QO_HAA + j * QCK_MAX_DRIVES + i; }

// One unit of write-out work, owned by one warp: nrep back-to-back copies of n doubles,
//   out_arr[dst + r*n + k] = image[img + k],  0 <= k < n, 0 <= r < nrep.
// nrep > 1 marks a kron(I_N, .) block: the same 2N x 2N values are stored N times.  The host cuts the per-integrator
// runs into units and balances them over the CTA's warps (per-member table: int hdr[QCK_SEG_HDR] = first unit of each warp).
struct QckSeg {
    int dst;       // first position inside the knot block (Hessian: >= nnzH means partial column dst - nnzH)
    int n;
    int img_nrep;  // image offset (doubles, always even) | nrep << 16
    int arr;       // (0 F, 1 J, 2 H) | (32 mod max(n/2, 1)) << 8
};
#define QCK_SEG_HDR 16  // ints (warp w owns units [hdr[w], hdr[w+1]))

// One auxiliary (derivative-integrator) entry of a knot block; evaluated by a single thread.
enum { QAUX_CONST = 0, QAUX_NEG_DT = 1, QAUX_NEG_Z = 2, QAUX_NEG_MU = 3, QAUX_FROW = 4 };
struct QckAux {
    int32_t out;  // 0 F, 1 J, 2 H
    int32_t op;
    int32_t pos;  // position inside the knot block of the output (H: >= nnzH means partial column pos-nnzH)
    int32_t i0;
    int32_t i1;
    int32_t pad;
    double c;
};

// One class of quantum integrators (same kind / N / columns / drives / order); members differ by offsets,
// Hamiltonians and output positions.
struct QckClassDev {
    int kind, N, NP, nc, ncp, nd, order, W;
    int free_time, dt_off, zdim, dyn;
    int antiherm;  // every member's Hamiltonians are Hermitian: A(a) = -i H(a) is anti-Hermitian
    int big;       // the class runs on the large-level kernel (qck_big.cu): operands in shared memory, outputs straight to the arrays
    int genexp;    // exponential class that runs on the generic spectral kernel (qck_genexp.cu): 5..16 levels, 9-level kets / ensembles
    int colexp;    // exponential class of 2..4 levels that runs on the spectral column kernels (qck_colexp.cu)
    int eig;       // exponential class that runs on the spectral kernel (qck_expeig.cu); built with the rs3 placement (rs3 = 1)
    int rs3;       // > 0: built for the three-warps-per-knot kernel (qck_rs3.cu) with this many knots per CTA: parity-matched
                   // image placement, unit table [phase][warp] (see qck_host.cpp)
    double dt_fixed;
    int n_members;
    // scratch (offsets in doubles)
    int off_A, msa, off_S, mss, off_img, img_doubles, scratch_doubles;
    // placement of every output quantity inside the image: element i lives at off_img + pl_base[q] + i * pl_stride[q]
    short pl_base[QO_COUNT];
    short pl_stride[QO_COUNT];
    // per member
    const int* moff;      // [member][state_off, ctrl_off, row_off]
    const double2* cmat;  // [member][A0: N*N | ell_val: nd*2*N*W | kk_val: kk_cap | ac_val: ac_cap]
    int cmat_stride;
    const int* ell_col;   // [member][ell_col: nd*2*N*W | kk_ptr: npair+1 | kk_rc: kk_cap | ac_ptr: N*N+1 | ac_j: ac_cap]
    int ell_stride;       // nd*2*N*W
    int icon_stride;      // ints per member in ell_col
    int kk_cap;
    int ac_cap;           // total nonzeros of the drives (per-element contributor lists of A)
    // write-out units per member: [member][QCK_SEG_HDR ints | nseg units], balanced over `threads`/32 warps
    int nseg, threads;
    const QckSeg* segs;  // (QCK_SEG_HDR/4 + nseg) QckSeg-sized records per member
    // shared-memory carve-up (byte offsets from the dynamic smem base; all 16-byte aligned)
    int sm_seg, sm_con, sm_stage, sm_bytes;
    int seg_bytes, con_bytes;  // size of ONE buffer of the double-buffered per-member tables
    int n_tbuf;                // 1 when a single member is active (tables never change), else 2
    // exponential integrators with a Hessian: per-CTA tape in global memory for the reverse sweep
    double2* tape;
    long long tape_stride;     // double2 elements per CTA
    int tape_levels;           // squaring levels the tape can hold
    int max_ctas;              // CTAs the tape was sized for (0 = no limit)
    // column kernel (levels <= 4): dense A_j per member [member][drive][N*N] row-major; first destination of every
    // output quantity per member [member][QO_COUNT] (-1 = absent; Hessian: >= nnzH means partial column)
    const double2* dense_aj;
    const int* qdst;
    // general-order Pade: degree m = order/2 and coefficient ratios r_k = c_{k+1}/c_k
    int pade_m;
    double pade_r[8];
};

// Opt-in cap on dynamic shared memory, always the device maximum: the attribute belongs to the kernel FUNCTION, so a launcher that set
// it to its own class's size would lower it under another live handle that runs the same instantiation with a larger size.
#define QCK_MAX_DYN_SMEM (227 * 1024)

// Launch geometry of one class, filled by the launcher at the first launch on a device (attribute + occupancy calls
// happen once per handle, not on every callback).
struct QckPlanCache {
    const void* kern = nullptr;
    size_t smem = 0;
    int per_sm = 0;
};

#define QCK_ST_EXP_RANGE 1  // exponential integrator with a Hessian: ||h A||_1 needs more squarings than the tape holds

struct QckLaunch {
    QckClassDev c;
    const double* Z;
    const double* mu;
    double* F;
    double* J;
    double* H;
    double* partial;
    long long n_knots;
    long long nnzJ, nnzH;
    int npart;
    unsigned mask;
    int member_begin, member_end;
    const QckAux* aux;  // processed by the first active member's CTA of class 0 (or by the aux kernel)
    int n_aux;
    const int* moff_global;  // [active member][state_off, ctrl_off, row_off]
    int moff_smem;           // copy them to shared memory at kernel start (set by the launcher)
    int sm_count;
    int group_threads;       // threads cooperating on one work item (set by the launcher)
    int group_smem;          // bytes of shared memory per group
    long long* timing;       // optional per-stage cycle counters (debug)
    unsigned stagger_ns;     // start-up delay step between the CTAs of one SM
    int hoff;                // row-slice kernel: image offset where the Hessian part starts
    int db;                  // row-slice kernel: separate staging buffers for the F + J image and the Hessian image
    int spread;              // row-slice kernel: the kron block copies are issued right after A^2, ahead of the rest of phase 1
    int* status;             // device-side error word (QCK_ST_* bits)
    int sole_writer;         // this launch (one class, one member, its fused aux entries) writes every position of the knot blocks
    QckPlanCache* plan;      // host-side: per-class launch plan cache (may be NULL)
};

struct QckReduce {  // fixed-order reduction of shared Hessian positions
    const int* pos;   // [n_shared] position inside the knot block
    const int* ptr;   // [n_shared+1] CSR into cols
    const int* cols;  // partial columns, ascending integrator order
    int n_shared;
};

#define QCK_MAX_GPUS 16
struct QckPeerReduce {  // shared Hessian positions summed over the partial columns of several GPUs (peer memory)
    const int* pos;   // [n_shared] position inside the knot block
    const int* ptr;   // [n_shared+1] CSR into cols
    const int* cols;  // (gpu << 24) | partial column, ascending integrator order
    int n_shared;
    const double* partial[QCK_MAX_GPUS];  // every GPU's partial-column buffer (peer pointers)
    int npart[QCK_MAX_GPUS];
};
int qck_launch_peer_reduce(const QckPeerReduce& R, double* H, long long n_knots, long long nnzH, cudaStream_t stream, int* launches);

// kernel launchers (qck_kernels.cu).  Return cudaError_t as int.
int qck_launch_quantum(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches);
int qck_launch_aux(const QckLaunch& L, cudaStream_t stream, int* launches);
// specialised kernels (own translation units); *done tells whether the class was taken
int qck_launch_rowslice9(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
int qck_launch_column(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
int qck_launch_rs3(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
int qck_launch_big(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
size_t qck_big_smem(const QckClassDev& c);
size_t qck_rs3_smem(const QckClassDev& c, int hoff, int kpc);
int qck_rs3_hoff(const QckClassDev& c);
int qck_launch_expeig(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
int qck_launch_colexp(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
size_t qck_colexp_scratch_rec(int N);
int qck_launch_genexp(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches, bool* done);
size_t qck_genexp_warp_bytes(int N, int nc, int nd);
int qck_fused_aux_limit(void);
int qck_launch_reorder(const double* arr, const double* partial, double* out, const int* src, int C, long long nnz, int npart, long long nk,
                       cudaStream_t stream, int* launches);  // out[t*C + i] = src[i] < nnz ? arr[t*nnz + src[i]] : partial[t*npart + src[i] - nnz]
int qck_pick_threads(const QckClassDev& c);  // CTA size of the quantum kernel for this class  // more aux entries than this go through the stand-alone aux kernel
int qck_launch_reduce(const QckReduce& R, double* H, const double* partial, long long n_knots, long long nnzH,
                      int npart, cudaStream_t stream, int* launches);
// gathers the non-redundant positions of a value array into the compact D2H buffer: out[t*ostride + i] = arr[t*nnz + src[i]], i < C
int qck_launch_pack(const double* arr, double* out, const int* src, int C, long long ostride, long long nnz, long long nk, cudaStream_t stream, int* launches);
// out[t*C + i] = arr[t*nnz + src[i]] the other way round (scatter back): arr[t*nnz + src[i]] = in[t*C + i]
int qck_launch_unpack(double* arr, const double* in, const int* src, int C, long long ostride, long long nnz, long long nk, cudaStream_t stream, int* launches);
// scratch sizing shared by host map builder and kernels
void qck_scratch_layout(QckClassDev& c);   // phase 1: matrices + scalars (needed to compute slots)
void qck_smem_finalize(QckClassDev& c);    // phase 2: tables + staging, after W / tab_len / nseg are known
