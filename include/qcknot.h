/*
 * qcknot.h -- C-ABI of libqcknot.so: the B200-native knot-point dynamics evaluator.
 *
 * This is the drop-in boundary for ONE path of QuantumCollocation.jl: the five-field `QuantumDynamics` object
 * (F, dF, dF_structure, mu_d2F, mu_d2F_structure) that the problem templates hand to Ipopt, i.e.
 *
 *   dynamics.F(Z.datavec)                    /root/reference/test/scripts/integrator_test_1qubit.jl:45
 *   dynamics.dF(Z.datavec), .dF_structure    /root/reference/test/scripts/integrator_test_1qubit.jl:46
 *   dynamics.mu_d2F(Z.datavec, mu), ..._structure   integrator_test_1qubit.jl:50-52
 *
 * built from the integrator vector the templates construct:
 *   UnitaryPadeIntegrator / UnitaryExponentialIntegrator   unitary_smooth_pulse_problem.jl:163-173
 *   QuantumStatePadeIntegrator / ...ExponentialIntegrator   quantum_state_smooth_pulse_problem.jl:142-196
 *   DerivativeIntegrator                                    unitary_smooth_pulse_problem.jl:177-178
 *   one unitary integrator per sampled system               unitary_sampling_problem.jl:134-155
 *
 * Conventions
 *  - every function returns 0 on success and a nonzero QCK_E* code otherwise; the message is available from
 *    qck_last_error().  No C++ exception crosses this boundary and nothing aborts.
 *  - there is NO CPU fallback: qck_create() fails when no sm_100 device is present.
 *  - the caller owns every array it passes; the library copies the Hamiltonians at create time and never keeps a
 *    caller pointer past the call.  A handle is not re-entrant (Ipopt's callbacks are serial).
 *  - data layout is the reference's: Z = vec(data), data is zdim x T column-major, z_t = Z[t*zdim : (t+1)*zdim]
 *    (test/test_utils.jl:52-118); unitary iso-vec = vec(vcat(real(U), imag(U))) (trajectory_initialization.jl:137);
 *    ket iso = [Re psi; Im psi] (trajectory_initialization.jl:469-470); all values are Float64, indices Int64.
 *  - a handle describes `T` consecutive knot points = T-1 constraint blocks.  Knot-range sharding over GPUs is
 *    pointer arithmetic on the caller's side: shard g creates a handle with T = t1-t0+1 (one-knot halo) and
 *    passes Z + t0*zdim, F + t0*dyn, J + t0*nnzJ, H + t0*nnzH (values are knot-major, so shards are contiguous).
 */
#ifndef QCKNOT_H
#define QCKNOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* integrator kinds (a2, a3, a4, a5) */
#define QCK_UNITARY_PADE 0
#define QCK_UNITARY_EXP 1
#define QCK_KET_PADE 2
#define QCK_KET_EXP 3
#define QCK_DERIVATIVE 4

/* output selection mask for the fused entry points */
#define QCK_EVAL_F 1u
#define QCK_EVAL_J 2u
#define QCK_EVAL_H 4u

/* error codes */
#define QCK_OK 0
#define QCK_EINVAL 1    /* bad argument / unsupported configuration */
#define QCK_ENODEVICE 2 /* no usable sm_100 CUDA device */
#define QCK_ECUDA 3     /* a CUDA runtime call or kernel failed */
#define QCK_ENOMEM 4

/* One entry of the reference's `integrators` vector.
 * Quantum kinds: state component = [state_off, state_off+state_len), drive component = [ctrl_off, ctrl_off+n_drives);
 *                H_drift (may be NULL = 0) and H_drives are N x N ComplexF64, column-major, interleaved (re, im),
 *                exactly Julia's reinterpret(Float64, H).  G(a) = iso(-i (H_drift + sum_j a_j H_drives[j])).
 * Derivative:    x component = [state_off, +state_len), dx component = [ctrl_off, +state_len); H_* ignored. */
typedef struct qck_integrator_desc {
    int32_t kind;
    int32_t order; /* Pade order (4, 6, 8, 10, 12); ignored for the other kinds */
    int32_t levels; /* N */
    int32_t n_drives;
    int32_t state_off;
    int32_t state_len;
    int32_t ctrl_off;
    int32_t reserved;
    const double* H_drift;
    const double* H_drives;
} qck_integrator_desc;

typedef struct qck_problem_desc {
    int64_t T;        /* knot points in this shard (>= 2) */
    int32_t zdim;     /* variables per knot (NamedTrajectory.dim) */
    int32_t dt_off;   /* offset of the timestep component in z_t, or -1 for a fixed timestep */
    double dt_fixed;  /* used when dt_off < 0 */
    int32_t n_integrators;
    int32_t eval_hessian; /* 0: no Hessian structure/values (PiccoloOptions.eval_hessian=false) */
    int32_t device;       /* CUDA device ordinal; -1 = structure-only handle (sizes + structures, every eval fails) */
    int32_t integ_begin;  /* ensemble sharding: this handle evaluates integrators [integ_begin, integ_end);  */
    int32_t integ_end;    /* 0,0 = all.  Structures always describe the whole problem.                        */
    int32_t reserved;
    const qck_integrator_desc* integrators;
} qck_problem_desc;

typedef struct qck_handle qck_handle;

/* QuantumDynamics(integrators, traj): builds structures once, uploads constants, allocates device buffers. */
int qck_create(const qck_problem_desc* desc, qck_handle** out);
void qck_destroy(qck_handle* h);
/* message of the last failure on this handle; pass NULL for the last qck_create failure of this thread */
const char* qck_last_error(const qck_handle* h);

/* per-knot sizes: dyn = rows per knot block (Z.dims.states), nnzJ / nnzH = values per knot block.
 * Totals are (T-1) times these (integrator_test_1qubit.jl:44,48,50). */
int qck_sizes(const qck_handle* h, int64_t* dyn, int64_t* nnzJ, int64_t* nnzH);

/* dF_structure / mu_d2F_structure: (T-1)*nnz pairs, 1-based like the reference's Vector{Tuple{Int,Int}};
 * knot_offset shifts the block index (0 for an unsharded problem, t0 for a shard that starts at knot t0). */
int qck_jacobian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols);
int qck_hessian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols);

/* Host-buffer entry points: what the MOI callbacks bind (eval_constraint, eval_constraint_jacobian,
 * eval_hessian_lagrangian).  Z has T*zdim doubles, mu and F have (T-1)*dyn, J (T-1)*nnzJ, H (T-1)*nnzH.
 * H2D copy of the inputs, one kernel pass, D2H copy of the value array; synchronous on return. */
int qck_eval_residual(qck_handle* h, const double* Z, double* F);
int qck_eval_jacobian(qck_handle* h, const double* Z, double* J);
int qck_eval_hessian(qck_handle* h, const double* Z, const double* mu, double* H);
/* fused: any of F/J/H may be NULL; mu is required only when H is requested */
int qck_eval_all(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H);

/* Device-resident entry point: all pointers are device pointers on the handle's device, `stream` is a
 * cudaStream_t (NULL = the handle's own stream).  Asynchronous: returns after the launches are enqueued. */
int qck_eval_device(qck_handle* h, uint32_t mask, const double* dZ, const double* dmu, double* dF, double* dJ,
                    double* dH, void* stream);
/* the handle's own device buffers (what the host-buffer entry points stage through) */
int qck_device_buffers(qck_handle* h, double** dZ, double** dmu, double** dF, double** dJ, double** dH);
int qck_synchronize(qck_handle* h);

/* Hessian positions (0-based, within one knot block) that receive contributions from more than one integrator
 * (shared controls in the sampling problem).  With ensemble sharding each handle leaves its local partial sum
 * there and the caller all-reduces exactly these positions.  pos may be NULL to query the count. */
int qck_shared_hessian_positions(const qck_handle* h, int64_t* count, int64_t* pos);

/* page-lock / unlock a caller array (e.g. Ipopt's value buffers) so the D2H copy runs at full PCIe speed */
int qck_host_register(void* p, size_t bytes);
int qck_host_unregister(void* p);

/* kernels launched on this handle since creation (bench.py's gpu_launches) */
int qck_launch_count(const qck_handle* h, int64_t* launches);
/* version string of the library */
const char* qck_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QCKNOT_H */
