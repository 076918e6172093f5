// Internal structures shared by the host side (qck_host.cpp) and the kernels (qck_kernels.cu) of libqcknot.so.
// Nothing here is part of the C-ABI (include/qcknot.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/qcknot.h"

#define QCK_TILE 3          // register tile edge of the small complex products (3x3 complex per thread)
#define QCK_MAX_DRIVES 8
#define QCK_MAX_PADE_M 5    // Pade order <= 10

// ---- scratch layout of the Pade-4 kernel (indices of matrices inside the CTA's shared-memory scratch) --------
// "A-type" matrices are NP x NP complex, "state-type" are NP x ncp complex (ncp = NP for unitaries, 1 for kets).
// Several slots are overwritten in place by the final outputs during assembly (the "->" notes).
enum { QA_A = 0, QA_AH = 1 /* -> B */, QA_A2 = 2 /* -> F */, QA_COUNT = 3, QA_B = QA_AH, QA_F = QA_A2 };
// state-type, fixed part:  D -> R (residual), AS -> Th (d/ddt), AhM -> Kh0 (state_t x dt), AhAhM -> Kh1 (dt x state_t+1)
enum { QS_D = 0, QS_S = 1, QS_M = 2, QS_AD = 3, QS_AS = 4, QS_AHM = 5, QS_AAD = 6, QS_AHAHM = 7, QS_FIXED = 8 };
// state-type, per drive j (index QS_FIXED + QD_COUNT*j + k):
//   Q1 = A_j D,  N1 = A_j^H M -> Ka0 (state_t x a_j),  AQ1 = A (A_j D) -> Ta (d/da_j),  AhN1 = A^H (A_j^H M) -> Ka1 (a_j x state_t+1)
enum { QD_Q1 = 0, QD_N1 = 1, QD_AQ1 = 2, QD_AHN1 = 3, QD_COUNT = 4, QD_KA0 = QD_N1, QD_TA = QD_AQ1, QD_KA1 = QD_AHN1 };
// scalar slots (doubles) after the matrices: [ONE | Hah[0..nd) | Hhh | Haa column-wise (j*nd + i, i <= j)]
enum { QX_ONE = 0, QX_HAH = 1 /* + j */ };
static inline __host__ __device__ int qx_hhh(int nd) { return QX_HAH + nd; }
static inline __host__ __device__ int qx_haa(int nd, int i, int j) { return QX_HAH + nd + 1 + j * nd + i; }

// One contiguous run of output positions of one integrator: out[dst + k] = +-scratch[tab[src_off + k % period]].
// period < len marks the kron(I_N, .) blocks: the same 2N x 2N values are stored len/period times.
struct QckSeg {
    int dst;      // first position inside the knot block (Hessian: >= nnzH means partial column dst - nnzH)
    int len;
    int src_off;  // into the class table of scratch slots (uint16, bit 15 = negate)
    int period;
};

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
    double dt_fixed;
    int n_members;
    // scratch (offsets in doubles)
    int off_A, msa, off_S, mss, off_X, scratch_doubles;
    // per member
    const int* state_off;
    const int* ctrl_off;
    const int* row_off;
    const double2* cmat;  // [member][A0: N*N | Adr: nd*N*N | ell_val: nd*2*N*W]
    int cmat_stride;
    const int* ell_col;  // [member][nd*2*N*W]
    int ell_stride;
    // output maps: class-level table of scratch slots + per-member segments [member][nsegJ + nsegH]
    const uint16_t* tab;
    int tab_len;
    int nsegJ, nsegH;
    const QckSeg* segs;
    // shared-memory carve-up (byte offsets from the dynamic smem base; all 16-byte aligned)
    int sm_tab, sm_seg, sm_ell, sm_stage, sm_bytes;
    int seg_bytes, ell_bytes;  // size of ONE buffer of the double-buffered per-member tables
};

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
};

struct QckReduce {  // fixed-order reduction of shared Hessian positions
    const int* pos;   // [n_shared] position inside the knot block
    const int* ptr;   // [n_shared+1] CSR into cols
    const int* cols;  // partial columns, ascending integrator order
    int n_shared;
};

// kernel launchers (qck_kernels.cu).  Return cudaError_t as int.
int qck_launch_quantum(const QckLaunch& L, int sm_count, cudaStream_t stream, int* launches);
int qck_launch_aux(const QckLaunch& L, cudaStream_t stream, int* launches);
int qck_fused_aux_limit(void);  // more aux entries than this go through the stand-alone aux kernel
int qck_launch_reduce(const QckReduce& R, double* H, const double* partial, long long n_knots, long long nnzH,
                      int npart, cudaStream_t stream, int* launches);
// scratch sizing shared by host map builder and kernels
void qck_scratch_layout(QckClassDev& c);   // phase 1: matrices + scalars (needed to compute slots)
void qck_smem_finalize(QckClassDev& c);    // phase 2: tables + staging, after W / tab_len / nseg are known
