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
enum { QA_A = 0, QA_AH = 1, QA_A2 = 2, QA_F = 3, QA_B = 4, QA_COUNT = 5 };
// state-type, fixed part.  Several slots are overwritten by the final outputs during assembly:
//   D -> R (residual), AS -> Th (d/ddt), AhM -> Kh0 (state_t x dt), AhAhM -> Kh1 (dt x state_t+1)
enum { QS_D = 0, QS_S = 1, QS_M = 2, QS_AD = 3, QS_AS = 4, QS_AHM = 5, QS_AAD = 6, QS_AHAHM = 7, QS_FIXED = 8 };
// state-type, per drive j (index QS_FIXED + 7*j + k):
//   P -> Ta (d/da_j), N2 -> Ka0 (state_t x a_j), AhN1 -> Ka1 (a_j x state_t+1)
enum { QD_P = 0, QD_Q1 = 1, QD_N1 = 2, QD_Q2 = 3, QD_N2 = 4, QD_AQ1 = 5, QD_AHN1 = 6, QD_COUNT = 7 };
// scalar slots (doubles) after the matrices
enum { QX_ONE = 0, QX_HHH = 1, QX_HAH = 2 /* + j */ };
static inline __host__ __device__ int qx_haa(int nd, int i, int j) { return QX_HAH + nd + i * nd + j; }

// ---- scratch layout of the exponential kernel -----------------------------------------------------------------
// All N x N ("A-type") unless noted; see qck_kernels.cu for the algorithm.
struct QckExpLayout {
    int n_a;  // number of A-type matrices
    int n_s;  // number of state-type matrices
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
    // output maps, per member, cnt entries each
    int cntJ, cntH;
    const uint32_t* posJ;
    const uint16_t* srcJ;  // bit 15 = negate
    const uint32_t* posH;
    const uint16_t* srcH;
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
int qck_launch_reduce(const QckReduce& R, double* H, const double* partial, long long n_knots, long long nnzH,
                      int npart, cudaStream_t stream, int* launches);
// scratch sizing shared by host map builder and kernels
void qck_scratch_layout(QckClassDev& c, int eval_hessian);
