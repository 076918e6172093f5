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
 *  - a handle describes `T` consecutive knot points = T-1 constraint blocks.
 *  - multi-GPU: n_gpus > 1 makes ONE handle drive n_gpus devices from the one serial
 *    caller (Ipopt).  shard_mode KNOT splits the knot blocks into contiguous ranges (one-knot halo; values are
 *    knot-major, so every GPU owns a contiguous segment of F, J and H), shard_mode ENSEMBLE splits the quantum
 *    integrators (the sampled systems of unitary_sampling_problem.jl:134-155).  The host-buffer entry points take and
 *    fill the caller's single arrays exactly as with one GPU: every GPU copies its own part over its own PCIe link.
 *    NCCL (dlopen'ed libnccl.so.2, single-process communicator) is used only by qck_gather_device (all-gather of the
 *    device-resident segments) and for the all-reduce of the Hessian entries that several systems share.
 *  - host-buffer calls stage through library-owned page-locked memory, in knot chunks that overlap H2D, kernels, D2H
 *    and the host-side expansion; the caller's arrays can be ordinary pageable memory.  kron(I_N, B) blocks cross
 *    PCIe once and are written N times into the caller's array by the library's host threads.
 *  - the same Z presented to qck_eval_residual / _jacobian / _hessian in succession (what Ipopt does within one
 *    iteration) is uploaded once: the handle compares Z with its staged copy and reuses device-resident results.
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

/* multi-GPU partitioning */
#define QCK_SHARD_KNOT 0
#define QCK_SHARD_ENSEMBLE 1

/* intra-knot order of the structure entries (and of the value arrays).  The reference's Core is not part of the repository, so
 * its order cannot be read off; the candidates are policies of qck_create:
 *   QCK_ORDER_CSC             union pattern of all integrators, column-major (default; the order every kernel writes in)
 *   QCK_ORDER_ROW_MAJOR       union pattern, row-major
 *   QCK_ORDER_PER_INTEGRATOR  the integrators' own entry lists one after the other (column-major inside an integrator); a
 *                             Hessian position that several integrators touch appears once per integrator, each entry holding
 *                             that integrator's contribution -- the consumer sums duplicates (test/test_utils.jl:14-27)
 * Orders other than CSC are served by one extra gather pass on the device (single-GPU handles). */
#define QCK_ORDER_CSC 0
#define QCK_ORDER_ROW_MAJOR 1
#define QCK_ORDER_PER_INTEGRATOR 2

/* error codes */
#define QCK_OK 0
#define QCK_EINVAL 1    /* bad argument / unsupported configuration */
#define QCK_ENODEVICE 2 /* no usable sm_100 CUDA device */
#define QCK_ECUDA 3     /* a CUDA runtime call or kernel failed */
#define QCK_ENOMEM 4
#define QCK_ERANGE 5    /* an input is outside the range the kernels support (reported by the device after the pass) */
#define QCK_ENCCL 6     /* libnccl could not be loaded or a collective failed */

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
    int32_t device;       /* (first) CUDA device ordinal; -1 = structure-only handle (sizes + structures, every eval fails) */
    int32_t integ_begin;  /* manual ensemble sharding: this handle evaluates integrators [integ_begin, integ_end);      */
    int32_t integ_end;    /* integ_end < 0 = all integrators; begin == end = none.  Structures describe the whole problem. */
    int32_t n_gpus;       /* 0 or 1: one GPU.  N > 1: this handle drives devices device .. device+N-1 (or `devices`)  */
    const qck_integrator_desc* integrators;
    int32_t shard_mode;   /* QCK_SHARD_KNOT | QCK_SHARD_ENSEMBLE (n_gpus > 1) */
    int32_t host_threads; /* host threads for staging / expansion of the value arrays; 0 = all hardware threads (max 32) */
    const int32_t* devices; /* optional: explicit device ordinals, n_gpus entries (NULL = consecutive from `device`) */
    int32_t structure_order; /* QCK_ORDER_CSC (0, default) | QCK_ORDER_ROW_MAJOR | QCK_ORDER_PER_INTEGRATOR */
    int32_t reserved;
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
 * eval_hessian_lagrangian).  Z has T*zdim doubles, mu and F have (T-1)*dyn, J (T-1)*nnzJ, H (T-1)*nnzH; plain (pageable)
 * host memory is fine.  Chunked H2D / kernels / compact D2H / host-side expansion, overlapped; synchronous on return.
 * An unchanged Z (compared with the staged copy) is not uploaded again and its device-resident results are reused. */
int qck_eval_residual(qck_handle* h, const double* Z, double* F);
int qck_eval_jacobian(qck_handle* h, const double* Z, double* J);
int qck_eval_hessian(qck_handle* h, const double* Z, const double* mu, double* H);
/* fused: any of F/J/H may be NULL; mu is required only when H is requested */
int qck_eval_all(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H);

/* Device-resident entry point: all pointers are device pointers on the handle's device, `stream` is a
 * cudaStream_t (NULL = the handle's own stream).  Asynchronous: returns after the launches are enqueued. */
int qck_eval_device(qck_handle* h, uint32_t mask, const double* dZ, const double* dmu, double* dF, double* dJ,
                    double* dH, void* stream);
/* the handle's own device buffers (what the host-buffer entry points stage through); under a structure_order other than CSC
 * dJ / dH are the caller-order arrays that qck_eval_resident fills (the canonical arrays stay internal) */
int qck_device_buffers(qck_handle* h, double** dZ, double** dmu, double** dF, double** dJ, double** dH);
int qck_synchronize(qck_handle* h);

/* Hessian positions (0-based, within one knot block, in the handle's structure order) that receive contributions from more than one
 * integrator (shared controls in the sampling problem; none under QCK_ORDER_PER_INTEGRATOR, where every contribution has its own entry).  With ensemble sharding each handle leaves its local partial sum
 * there and the caller all-reduces exactly these positions.  pos may be NULL to query the count. */
int qck_shared_hessian_positions(const qck_handle* h, int64_t* count, int64_t* pos);

/* page-lock / unlock a caller array (kept for callers that run their own copies from qck_device_buffers; the host-buffer
 * entry points stage through library-owned page-locked memory and do not need it) */
int qck_host_register(void* p, size_t bytes);
int qck_host_unregister(void* p);

/* ---- multi-GPU handles (n_gpus > 1); every function also accepts a single-GPU handle (one shard) ------------------------- */
/* number of shards (GPUs) behind the handle, and what shard g covers: its device, its block range [block_begin, block_end)
 * (KNOT: a contiguous part of 0 .. T-1; ENSEMBLE: everything) and its integrator range (ENSEMBLE: a part; KNOT: everything) */
int qck_shard_count(const qck_handle* h, int32_t* n);
int qck_shard_info(const qck_handle* h, int32_t g, int32_t* device, int64_t* block_begin, int64_t* block_end,
                   int32_t* integ_begin, int32_t* integ_end);
/* shard g's own device buffers (its knots of Z / mu, its segments of F / J / H) */
int qck_shard_device_buffers(qck_handle* h, int32_t g, double** dZ, double** dmu, double** dF, double** dJ, double** dH);
/* device-resident evaluation on every GPU: qck_upload copies the caller's Z (and mu, may be NULL) into the shards' buffers
 * (synchronous), qck_eval_resident enqueues one pass per GPU on the shards' streams (asynchronous; qck_synchronize waits).
 * ENSEMBLE: the Hessian entries on the shared controls are summed over the GPUs with ncclAllReduce. */
int qck_upload(qck_handle* h, const double* Z, const double* mu);
int qck_eval_resident(qck_handle* h, uint32_t mask);
/* KNOT: assemble the device-resident value arrays on EVERY GPU (NCCL over NVLink, all-gather of the segments; asynchronous
 * on the shards' streams); qck_gathered_buffers returns GPU g's assembled arrays (allocated by the first gather). */
int qck_gather_device(qck_handle* h, uint32_t mask);
int qck_gathered_buffers(qck_handle* h, int32_t g, double** dF, double** dJ, double** dH);
/* NCCL version the library loaded (dlopen) and the size of its communicator; initialises NCCL for a multi-GPU handle */
int qck_nccl_version(qck_handle* h, int32_t* version, int32_t* nranks);

/* bytes the last host-buffer call moved over PCIe (all GPUs) and the number of calls served from device-resident results
 * of an unchanged Z since creation */
int qck_transfer_stats(const qck_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* cache_hits);
/* Introspection of the host-buffer path.  qck_compact_map: the runs of value array `arr` (0 F, 1 J, 2 H) that this handle
 * writes, per knot block, as (offset in the block, offset in the compact layout, length, repeats) quadruples -- repeats > 1 is
 * a kron(I_N, B) block that crosses PCIe once.  segs may be NULL to query the count.  qck_expand_host runs the host half of the
 * path alone (compact layout -> structure-order array; no device involved, works on structure-only handles). */
int qck_compact_map(const qck_handle* h, int32_t arr, int64_t* count, int32_t* segs);
int qck_expand_host(const qck_handle* h, int32_t arr, const double* compact, double* out, int64_t nk);
/* forget the staged inputs / device-resident results (e.g. after writing into the handle's device buffers directly) */
int qck_invalidate(qck_handle* h);

/* ---- objective and terminal-constraint terms on the device ----------------------------------
 * The terms the problem templates add next to the dynamics, evaluated from the same device-resident Z as the dynamics (one
 * upload serves every callback of an Ipopt iteration):
 *   QCK_OBJ_QUADRATIC_REGULARIZER  QuadraticRegularizer(name, traj, R; timestep_name)   unitary_smooth_pulse_problem.jl:151-153
 *                                  J = sum_t 1/2 weight * dt_t^2 * sum_i R_i v_ti^2
 *   QCK_OBJ_UNITARY_INFIDELITY     UnitaryInfidelityObjective(state_name, traj, Q; subspace)   unitary_smooth_pulse_problem.jl:132-137
 *                                  J = weight * (1 - |tr(U_goal' U_T)|^2 / n_sub^2) on the final knot
 *   QCK_OBJ_MINIMUM_TIME           MinimumTimeObjective(traj; D)   unitary_minimum_time_problem.jl:67-69:  J = weight * sum_{t<T} dt_t
 * and FinalUnitaryFidelityConstraint(state_name, val, traj)  (unitary_minimum_time_problem.jl:80-84): g = F(U_T) - val >= 0. */
#define QCK_OBJ_QUADRATIC_REGULARIZER 0
#define QCK_OBJ_UNITARY_INFIDELITY 1
#define QCK_OBJ_MINIMUM_TIME 2
typedef struct qck_objective_term {
    int32_t kind;
    int32_t comp_off;   /* component inside z_t: the regularised variable / the unitary's iso-vec (ignored for MINIMUM_TIME) */
    int32_t comp_len;
    int32_t levels;     /* infidelity: N (comp_len = 2 N^2) */
    double weight;      /* Q / D / a factor on R */
    const double* R;    /* regularizer: comp_len weights, NULL = ones */
    const double* goal; /* infidelity: iso-vec of the goal operator (zeros outside the subspace of an EmbeddedOperator) */
    int32_t n_sub;      /* infidelity: dimension of the subspace the fidelity is normalised with; 0 = levels */
    int32_t reserved;
} qck_objective_term;
/* J = sum of the terms.  Works on single-GPU and knot-sharded multi-GPU handles (per-GPU partial sums, added in GPU order). */
int qck_objective_attach(qck_handle* h, const qck_objective_term* terms, int32_t n_terms);
/* n_vars = T * zdim (gradient length); nnz_hess values in the structure below */
int qck_objective_sizes(const qck_handle* h, int64_t* n_vars, int64_t* nnz_hess);
/* 1-based (row <= col), knot-major: for every knot the regularizers' [v_i x v_i | v_i x dt | dt x dt] entries in term order,
 * then the infidelity terms' dense upper triangles (by column) on the final knot; duplicates are summed by the consumer */
int qck_objective_hessian_structure(const qck_handle* h, int64_t* rows, int64_t* cols);
int qck_eval_objective(qck_handle* h, const double* Z, double* value);
int qck_eval_objective_gradient(qck_handle* h, const double* Z, double* grad);
int qck_eval_objective_hessian(qck_handle* h, const double* Z, double sigma, double* vals);
/* terminal fidelity constraint (term->kind = QCK_OBJ_UNITARY_INFIDELITY describes state component, goal, subspace).
 * qck_eval_fidelity_constraint: g (may be NULL), its Jacobian row over the final knot's state component (comp_len values,
 * 1-based columns (T-1)*zdim + comp_off + 1 ..., may be NULL) and mu * the upper triangle (by column) of its Hessian
 * (comp_len (comp_len + 1) / 2 values, may be NULL). */
int qck_fidelity_constraint_attach(qck_handle* h, const qck_objective_term* term, double min_fidelity);
int qck_eval_fidelity_constraint(qck_handle* h, const double* Z, double mu, double* g, double* jac, double* hess);

/* ---- rollouts on the device ---------------------------------------------------------------
 * unitary_rollout (ket = 0) / rollout (ket = 1) of the reference (src/trajectory_initialization.jl:426,493; the fidelity
 * assertions of every template test, e.g. unitary_smooth_pulse_problem.jl:218-220; robustness sweeps over sampled systems,
 * unitary_sampling_problem.jl:233-243):  X_1 = X_init, X_{t+1} = exp(dt_t G_s(a_t)) X_t for n_systems systems sharing the
 * controls.  H_drift [n_systems][N*N] (may be NULL), H_drives [n_systems][n_drives][N*N]: ComplexF64 column-major, interleaved;
 * a [T][n_drives] = vec of the n_drives x T control matrix; dt [T]; X_init [n_systems][dim] iso-vec(s), NULL = identity / first
 * basis ket; X_out [n_systems][T][dim], dim = 2 N^2 (unitary) or 2 N (ket): per system the dim x T component, column-major.
 * levels <= 16.  Independent of any handle; errors through qck_rollout_last_error(). */
int qck_rollout(int32_t device, int32_t ket, int32_t levels, int32_t n_drives, int32_t n_systems, const double* H_drift,
                const double* H_drives, int64_t T, const double* a, const double* dt, const double* X_init, double* X_out);
const char* qck_rollout_last_error(void);

/* kernels launched on this handle since creation (bench.py's gpu_launches) */
int qck_launch_count(const qck_handle* h, int64_t* launches);
/* version string of the library */
const char* qck_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QCKNOT_H */
