// Objective and terminal-constraint terms on the device: the terms the problem templates add
// next to the dynamics,
//   QuadraticRegularizer(name, traj, R; timestep_name)          unitary_smooth_pulse_problem.jl:151-153
//   UnitaryInfidelityObjective(state_name, traj, Q; subspace)    unitary_smooth_pulse_problem.jl:132-137
//   MinimumTimeObjective(traj; D)                                unitary_minimum_time_problem.jl:67-69
//   FinalUnitaryFidelityConstraint(state_name, val, traj)        unitary_minimum_time_problem.jl:80-84
// evaluated from the SAME device-resident Z as the dynamics (one upload serves Ipopt's eval_f, eval_grad_f, eval_g, eval_jac_g
// and eval_h of an iteration): value (fixed-order reduction, bitwise reproducible), dense gradient, and Hessian values in a
// knot-major structure (for every knot, for every term, that term's upper-triangular entries).
#include <algorithm>
#include <cmath>

#include "qck_handle.h"

namespace {

struct ObjTermDev {
    int kind, comp_off, comp_len, n_sub;
    double weight;
    const double* R;     // regularizer weights (device), comp_len
    const double* g;     // fidelity: goal iso-vec (device)
    const double* w;     // fidelity: rotated goal (tr(G' U) = g.u + i w.u)
    int hess_per_knot;   // entries per knot (regularizer) / at the last knot (infidelity)
};

struct ObjDev {
    int n_terms;
    ObjTermDev t[QCK_MAX_OBJ_TERMS];
    long long T;        // knots in this shard
    int zdim, dt_off;
    double dt_fixed;
    int last;           // this shard holds the final knot of the trajectory (terminal terms live here)
    long long own_knots;  // knots whose per-knot terms belong to this shard (a shard's halo knot belongs to the next one)
    int hess_knot;      // Hessian entries per ordinary knot
    int hess_last;      // extra entries at the final knot
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    __syncthreads();
    return s;  // valid on thread 0
}

// fid[0] = a = g.u, fid[1] = b = w.u for every fidelity term (one CTA per term), u = state component of the final knot
__global__ void __launch_bounds__(256) qck_obj_fid_kernel(const ObjDev o, const double* __restrict__ Z, double* __restrict__ fid) {
    __shared__ double sh[8];
    const ObjTermDev& tm = o.t[blockIdx.x];
    if (tm.kind != QCK_OBJ_UNITARY_INFIDELITY || !o.last) return;
    const double* u = Z + (o.T - 1) * o.zdim + tm.comp_off;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < tm.comp_len; i += blockDim.x) { a = fma(tm.g[i], u[i], a); b = fma(tm.w[i], u[i], b); }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { fid[2 * blockIdx.x] = a; fid[2 * blockIdx.x + 1] = b; }
}

// per-CTA partial sums of the per-knot terms (fixed assignment of knots to CTAs; summed in order by the second kernel)
__global__ void __launch_bounds__(256) qck_obj_value_kernel(const ObjDev o, const double* __restrict__ Z, double* __restrict__ part) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < o.own_knots; t += (long long)gridDim.x * blockDim.x) {
        const double* z = Z + t * o.zdim;
        const double dt = o.dt_off >= 0 ? z[o.dt_off] : o.dt_fixed;
        for (int k = 0; k < o.n_terms; ++k) {
            const ObjTermDev& tm = o.t[k];
            if (tm.kind == QCK_OBJ_QUADRATIC_REGULARIZER) {
                double s = 0.0;
                for (int i = 0; i < tm.comp_len; ++i) { const double v = z[tm.comp_off + i]; s = fma(tm.R[i] * v, v, s); }
                acc += 0.5 * tm.weight * dt * dt * s;
            } else if (tm.kind == QCK_OBJ_MINIMUM_TIME) {
                if (!(o.last && t == o.T - 1)) acc += tm.weight * z[tm.comp_off];  // sum over t < T of the whole trajectory
            }
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void qck_obj_value_final_kernel(const ObjDev o, const double* __restrict__ part, int nparts, const double* __restrict__ fid,
                                           double* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    if (o.last)
        for (int k = 0; k < o.n_terms; ++k)
            if (o.t[k].kind == QCK_OBJ_UNITARY_INFIDELITY) {
                const double a = fid[2 * k], b = fid[2 * k + 1], n = (double)o.t[k].n_sub;
                s += o.t[k].weight * (1.0 - (a * a + b * b) / (n * n));
            }
    out[0] = s;
}

// dense gradient over this shard's own knots: one thread per (knot, variable)
__global__ void __launch_bounds__(256) qck_obj_grad_kernel(const ObjDev o, const double* __restrict__ Z, const double* __restrict__ fid,
                                                           double* __restrict__ grad) {
    const long long total = o.own_knots * o.zdim;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / o.zdim;
        const int i = (int)(e - t * o.zdim);
        const double* z = Z + t * o.zdim;
        const double dt = o.dt_off >= 0 ? z[o.dt_off] : o.dt_fixed;
        const bool final_knot = o.last && t == o.T - 1;
        double gv = 0.0;
        for (int k = 0; k < o.n_terms; ++k) {
            const ObjTermDev& tm = o.t[k];
            if (tm.kind == QCK_OBJ_QUADRATIC_REGULARIZER) {
                if (i >= tm.comp_off && i < tm.comp_off + tm.comp_len) gv += tm.weight * dt * dt * tm.R[i - tm.comp_off] * z[i];
                if (i == o.dt_off) {
                    double s = 0.0;
                    for (int j = 0; j < tm.comp_len; ++j) { const double v = z[tm.comp_off + j]; s = fma(tm.R[j] * v, v, s); }
                    gv += tm.weight * dt * s;
                }
            } else if (tm.kind == QCK_OBJ_MINIMUM_TIME) {
                if (i == tm.comp_off && !final_knot) gv += tm.weight;
            } else if (tm.kind == QCK_OBJ_UNITARY_INFIDELITY) {
                if (final_knot && i >= tm.comp_off && i < tm.comp_off + tm.comp_len) {
                    const double a = fid[2 * k], b = fid[2 * k + 1], n = (double)tm.n_sub;
                    gv -= tm.weight * 2.0 * (a * tm.g[i - tm.comp_off] + b * tm.w[i - tm.comp_off]) / (n * n);
                }
            }
        }
        grad[e] = gv;
    }
}

// Hessian values, knot-major: per own knot hess_knot entries (+ hess_last at the final knot), scaled by sigma
__global__ void __launch_bounds__(256) qck_obj_hess_kernel(const ObjDev o, const double* __restrict__ Z, double sigma, double* __restrict__ vals) {
    for (long long t = blockIdx.x; t < o.own_knots; t += gridDim.x) {
        const double* z = Z + t * o.zdim;
        const double dt = o.dt_off >= 0 ? z[o.dt_off] : o.dt_fixed;
        double* out = vals + t * o.hess_knot;
        int base = 0;
        for (int k = 0; k < o.n_terms; ++k) {
            const ObjTermDev& tm = o.t[k];
            if (tm.kind == QCK_OBJ_QUADRATIC_REGULARIZER) {
                const int n = tm.comp_len;
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    out[base + i] = sigma * tm.weight * dt * dt * tm.R[i];
                    if (o.dt_off >= 0) out[base + n + i] = sigma * tm.weight * 2.0 * dt * tm.R[i] * z[tm.comp_off + i];
                }
                if (o.dt_off >= 0 && threadIdx.x == 0) {
                    double s = 0.0;
                    for (int j = 0; j < n; ++j) { const double v = z[tm.comp_off + j]; s = fma(tm.R[j] * v, v, s); }
                    out[base + 2 * n] = sigma * tm.weight * s;
                }
                base += tm.hess_per_knot;
            }
        }
        if (o.last && t == o.T - 1) {  // terminal terms: upper triangle (by column) of -Q * 2 (g g' + w w') / n^2
            double* lo = out + o.hess_knot;
            int lb = 0;
            for (int k = 0; k < o.n_terms; ++k) {
                const ObjTermDev& tm = o.t[k];
                if (tm.kind != QCK_OBJ_UNITARY_INFIDELITY) continue;
                const double n = (double)tm.n_sub, sc = -sigma * tm.weight * 2.0 / (n * n);
                for (int e = threadIdx.x; e < tm.hess_per_knot; e += blockDim.x) {
                    // e -> (i <= j), e = j (j + 1) / 2 + i
                    int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                    while (j * (j + 1) / 2 > e) --j;
                    while ((j + 1) * (j + 2) / 2 <= e) ++j;
                    const int i = e - j * (j + 1) / 2;
                    lo[lb + e] = sc * (tm.g[i] * tm.g[j] + tm.w[i] * tm.w[j]);
                }
                lb += tm.hess_per_knot;
            }
        }
    }
}

// terminal fidelity constraint g = F(U_T) - val, its Jacobian row and mu * Hessian (same algebra as the infidelity term)
__global__ void __launch_bounds__(256) qck_fid_constraint_kernel(const ObjTermDev tm, const double* __restrict__ u, double val, double mu,
                                                                 double* __restrict__ out_g, double* __restrict__ out_jac, double* __restrict__ out_hess) {
    __shared__ double sh[8];
    __shared__ double ab[2];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < tm.comp_len; i += blockDim.x) { a = fma(tm.g[i], u[i], a); b = fma(tm.w[i], u[i], b); }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { ab[0] = a; ab[1] = b; }
    __syncthreads();
    a = ab[0]; b = ab[1];
    const double n = (double)tm.n_sub, n2 = n * n;
    if (threadIdx.x == 0 && out_g) out_g[0] = (a * a + b * b) / n2 - val;
    if (out_jac)
        for (int i = threadIdx.x; i < tm.comp_len; i += blockDim.x) out_jac[i] = 2.0 * (a * tm.g[i] + b * tm.w[i]) / n2;
    if (out_hess)
        for (int e = threadIdx.x; e < tm.comp_len * (tm.comp_len + 1) / 2; e += blockDim.x) {
            int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (j * (j + 1) / 2 > e) --j;
            while ((j + 1) * (j + 2) / 2 <= e) ++j;
            const int i = e - j * (j + 1) / 2;
            out_hess[e] = mu * 2.0 * (tm.g[i] * tm.g[j] + tm.w[i] * tm.w[j]) / n2;
        }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------------------------------
struct QckObjective {
    ObjDev dev{};
    std::vector<qck_objective_term> terms;
    std::vector<void*> allocs;
    double *d_fid = nullptr, *d_part = nullptr, *d_out = nullptr, *d_grad = nullptr, *d_hess = nullptr;
    double *pin = nullptr;      // page-locked staging of the results
    long long pin_doubles = 0;
    int nparts = 0;
    long long nnz_hess = 0;     // this shard
    // terminal fidelity constraint
    bool has_con = false;
    ObjTermDev con{};
    double con_val = 0.0;
    double *d_con = nullptr;    // [g | jac (len) | hess (len (len + 1) / 2)]
};

namespace {

int obj_upload(qck_handle* h, QckObjective* O, const double* src, int n, const double** out) {
    void* d = nullptr;
    QCK_CUDA_TRY(h, cudaMalloc(&d, sizeof(double) * std::max(n, 1)));
    O->allocs.push_back(d);
    QCK_CUDA_TRY(h, cudaMemcpy(d, src, sizeof(double) * n, cudaMemcpyHostToDevice));
    *out = static_cast<const double*>(d);
    return QCK_OK;
}

void rotate_goal(const double* g, int N, std::vector<double>& w) {  // tr(G' U) = g.u + i w.u
    w.resize((size_t)2 * N * N);
    for (int c = 0; c < N; ++c)
        for (int r = 0; r < N; ++r) {
            w[(size_t)c * 2 * N + r] = -g[(size_t)c * 2 * N + N + r];
            w[(size_t)c * 2 * N + N + r] = g[(size_t)c * 2 * N + r];
        }
}

int term_to_dev(qck_handle* h, QckObjective* O, const qck_objective_term& s, ObjTermDev& d, bool free_time) {
    d = ObjTermDev{};
    d.kind = s.kind; d.comp_off = s.comp_off; d.comp_len = s.comp_len; d.weight = s.weight; d.n_sub = s.n_sub;
    if (s.comp_off < 0 || s.comp_len <= 0 || s.comp_off + s.comp_len > h->zdim) return qck_fail(h, QCK_EINVAL, "objective term: component out of range");
    if (s.kind == QCK_OBJ_QUADRATIC_REGULARIZER) {
        std::vector<double> R(s.comp_len, 1.0);
        if (s.R) R.assign(s.R, s.R + s.comp_len);
        if (h->device >= 0) { int rc = obj_upload(h, O, R.data(), s.comp_len, &d.R); if (rc) return rc; }
        d.hess_per_knot = free_time ? 2 * s.comp_len + 1 : s.comp_len;
    } else if (s.kind == QCK_OBJ_UNITARY_INFIDELITY) {
        const int N = s.levels;
        if (N < 1 || s.comp_len != 2 * N * N || !s.goal) return qck_fail(h, QCK_EINVAL, "infidelity term: needs levels and a goal iso-vec of 2 N^2 entries");
        if (d.n_sub <= 0) d.n_sub = N;
        std::vector<double> w;
        rotate_goal(s.goal, N, w);
        if (h->device >= 0) {
            int rc = obj_upload(h, O, s.goal, s.comp_len, &d.g);
            if (rc) return rc;
            rc = obj_upload(h, O, w.data(), s.comp_len, &d.w);
            if (rc) return rc;
        }
        d.hess_per_knot = s.comp_len * (s.comp_len + 1) / 2;
    } else if (s.kind == QCK_OBJ_MINIMUM_TIME) {
        if (!free_time) return qck_fail(h, QCK_EINVAL, "MinimumTimeObjective needs a free timestep");
        d.comp_off = h->dt_off; d.comp_len = 1;
        d.hess_per_knot = 0;
    } else {
        return qck_fail(h, QCK_EINVAL, "unknown objective term kind %d", s.kind);
    }
    return QCK_OK;
}

// shard geometry: child g of a knot-sharded handle owns knots [t0_g, t0_g+1) and, if it is the last one, the final knot
struct Shard { qck_handle* h; long long k0; bool last; };
std::vector<Shard> shards_of(qck_handle* h) {
    std::vector<Shard> v;
    if (h->children.empty()) { v.push_back({h, 0, true}); return v; }
    for (size_t g = 0; g < h->children.size(); ++g) v.push_back({h->children[g], h->child_t0[g], g + 1 == h->children.size()});
    return v;
}

int attach_one(qck_handle* h, const qck_objective_term* terms, int n, bool last) {
    if (h->objective) { qck_objective_free(h); }
    QckObjective* O = new QckObjective();
    h->objective = O;
    O->terms.assign(terms, terms + n);
    ObjDev& D = O->dev;
    D.n_terms = n; D.T = h->T; D.zdim = h->zdim; D.dt_off = h->dt_off; D.dt_fixed = h->dt_fixed; D.last = last ? 1 : 0;
    D.own_knots = last ? h->T : h->T - 1;
    if (h->device >= 0) QCK_CUDA_TRY(h, cudaSetDevice(h->device));
    for (int k = 0; k < n; ++k) {
        int rc = term_to_dev(h, O, terms[k], D.t[k], h->dt_off >= 0);
        if (rc) return rc;
        if (terms[k].kind == QCK_OBJ_QUADRATIC_REGULARIZER) D.hess_knot += D.t[k].hess_per_knot;
        if (terms[k].kind == QCK_OBJ_UNITARY_INFIDELITY) D.hess_last += D.t[k].hess_per_knot;
    }
    O->nnz_hess = D.own_knots * D.hess_knot + (last ? D.hess_last : 0);
    if (h->device < 0) return QCK_OK;
    O->nparts = (int)std::min<long long>(296, (D.own_knots + 255) / 256);
    void* p = nullptr;
    QCK_CUDA_TRY(h, cudaMalloc(&p, sizeof(double) * (2 * QCK_MAX_OBJ_TERMS + O->nparts + 8)));
    O->allocs.push_back(p);
    O->d_fid = static_cast<double*>(p);
    O->d_part = O->d_fid + 2 * QCK_MAX_OBJ_TERMS;
    O->d_out = O->d_part + O->nparts;
    QCK_CUDA_TRY(h, cudaMalloc(&p, sizeof(double) * std::max<long long>(D.own_knots * D.zdim, 1)));
    O->allocs.push_back(p);
    O->d_grad = static_cast<double*>(p);
    QCK_CUDA_TRY(h, cudaMalloc(&p, sizeof(double) * std::max<long long>(O->nnz_hess, 1)));
    O->allocs.push_back(p);
    O->d_hess = static_cast<double*>(p);
    O->pin_doubles = std::max<long long>(std::max<long long>(D.own_knots * D.zdim, O->nnz_hess), 8);
    QCK_CUDA_TRY(h, cudaHostAlloc((void**)&O->pin, sizeof(double) * O->pin_doubles, cudaHostAllocPortable));
    return QCK_OK;
}

}  // namespace

void qck_objective_free(qck_handle* h) {
    QckObjective* O = static_cast<QckObjective*>(h->objective);
    if (!O) return;
    if (h->device >= 0) {
        cudaSetDevice(h->device);
        for (void* p : O->allocs) cudaFree(p);
        if (O->pin) cudaFreeHost(O->pin);
    }
    delete O;
    h->objective = nullptr;
}

extern "C" {

int qck_objective_attach(qck_handle* h, const qck_objective_term* terms, int32_t n_terms) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !terms || n_terms <= 0) return QCK_EINVAL;
    if (n_terms > QCK_MAX_OBJ_TERMS) return qck_fail(h, QCK_EINVAL, "at most %d objective terms", QCK_MAX_OBJ_TERMS);
    if (!h->children.empty() && h->shard_mode != QCK_SHARD_KNOT) return qck_fail(h, QCK_EINVAL, "objective terms on a multi-GPU handle need knot sharding");
    for (auto& s : shards_of(h)) {
        int rc = attach_one(s.h, terms, n_terms, s.last);
        if (rc) { if (s.h != h) h->err = s.h->err; return rc; }
    }
    if (!h->children.empty()) {  // the parent carries the term list for the structure
        int rc = attach_one(h, terms, n_terms, true);
        if (rc) return rc;
    }
    return QCK_OK;
}

int qck_objective_sizes(const qck_handle* h, int64_t* n_vars, int64_t* nnz_hess) {
    if (!h || !h->objective) return QCK_EINVAL;
    const QckObjective* O = static_cast<const QckObjective*>(h->objective);
    if (n_vars) *n_vars = h->T * h->zdim;
    if (nnz_hess) *nnz_hess = h->T * O->dev.hess_knot + O->dev.hess_last;
    return QCK_OK;
}

// 1-based (row, col), row <= col, knot-major: per knot the regularizers' entries [v_i x v_i | v_i x dt | dt x dt] in term
// order, after the final knot the infidelity terms' dense upper triangles by column
int qck_objective_hessian_structure(const qck_handle* h, int64_t* rows, int64_t* cols) {
    if (!h || !h->objective || !rows || !cols) return QCK_EINVAL;
    const QckObjective* O = static_cast<const QckObjective*>(h->objective);
    const ObjDev& D = O->dev;
    long long e = 0;
    for (long long t = 0; t < h->T; ++t) {
        const long long o = t * h->zdim;
        for (int k = 0; k < D.n_terms; ++k) {
            const ObjTermDev& tm = D.t[k];
            if (tm.kind != QCK_OBJ_QUADRATIC_REGULARIZER) continue;
            for (int i = 0; i < tm.comp_len; ++i) { rows[e] = cols[e] = o + tm.comp_off + i + 1; ++e; }
            if (D.dt_off >= 0) {
                for (int i = 0; i < tm.comp_len; ++i) {
                    const long long r = o + tm.comp_off + i + 1, c = o + D.dt_off + 1;
                    rows[e] = std::min(r, c); cols[e] = std::max(r, c); ++e;
                }
                rows[e] = cols[e] = o + D.dt_off + 1; ++e;
            }
        }
    }
    const long long o = (h->T - 1) * h->zdim;
    for (int k = 0; k < D.n_terms; ++k) {
        const ObjTermDev& tm = D.t[k];
        if (tm.kind != QCK_OBJ_UNITARY_INFIDELITY) continue;
        for (int j = 0; j < tm.comp_len; ++j)
            for (int i = 0; i <= j; ++i) { rows[e] = o + tm.comp_off + i + 1; cols[e] = o + tm.comp_off + j + 1; ++e; }
    }
    return QCK_OK;
}

int qck_eval_objective(qck_handle* h, const double* Z, double* value) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !Z || !value || !h->objective) return QCK_EINVAL;
    double total = 0.0;
    for (auto& s : shards_of(h)) {
        qck_handle* c = s.h;
        QckObjective* O = static_cast<QckObjective*>(c->objective);
        if (c->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle: libqcknot has no CPU evaluation path");
        int rc = qck_pipe_ensure_z(c, Z + s.k0 * h->zdim);
        if (rc) { if (c != h) h->err = c->err; return rc; }
        cudaStream_t st = c->stream;
        qck_obj_fid_kernel<<<O->dev.n_terms, 256, 0, st>>>(O->dev, c->dZ, O->d_fid);
        qck_obj_value_kernel<<<O->nparts, 256, 0, st>>>(O->dev, c->dZ, O->d_part);
        qck_obj_value_final_kernel<<<1, 32, 0, st>>>(O->dev, O->d_part, O->nparts, O->d_fid, O->d_out);
        c->launches += 3;
        QCK_CUDA_TRY(h, cudaMemcpyAsync(O->pin, O->d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    for (auto& s : shards_of(h)) {  // partial sums of the shards, added in shard order
        QCK_CUDA_TRY(h, cudaSetDevice(s.h->device));
        QCK_CUDA_TRY(h, cudaStreamSynchronize(s.h->stream));
        total += static_cast<QckObjective*>(s.h->objective)->pin[0];
    }
    *value = total;
    return QCK_OK;
}

int qck_eval_objective_gradient(qck_handle* h, const double* Z, double* grad) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !Z || !grad || !h->objective) return QCK_EINVAL;
    for (auto& s : shards_of(h)) {
        qck_handle* c = s.h;
        QckObjective* O = static_cast<QckObjective*>(c->objective);
        if (c->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle: libqcknot has no CPU evaluation path");
        int rc = qck_pipe_ensure_z(c, Z + s.k0 * h->zdim);
        if (rc) { if (c != h) h->err = c->err; return rc; }
        cudaStream_t st = c->stream;
        const long long n = O->dev.own_knots * O->dev.zdim;
        qck_obj_fid_kernel<<<O->dev.n_terms, 256, 0, st>>>(O->dev, c->dZ, O->d_fid);
        qck_obj_grad_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(O->dev, c->dZ, O->d_fid, O->d_grad);
        c->launches += 2;
        QCK_CUDA_TRY(h, cudaMemcpyAsync(O->pin, O->d_grad, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    for (auto& s : shards_of(h)) {
        QckObjective* O = static_cast<QckObjective*>(s.h->objective);
        QCK_CUDA_TRY(h, cudaSetDevice(s.h->device));
        QCK_CUDA_TRY(h, cudaStreamSynchronize(s.h->stream));
        memcpy(grad + s.k0 * h->zdim, O->pin, sizeof(double) * O->dev.own_knots * O->dev.zdim);
    }
    return QCK_OK;
}

int qck_eval_objective_hessian(qck_handle* h, const double* Z, double sigma, double* vals) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !Z || !vals || !h->objective) return QCK_EINVAL;
    for (auto& s : shards_of(h)) {
        qck_handle* c = s.h;
        QckObjective* O = static_cast<QckObjective*>(c->objective);
        if (c->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle: libqcknot has no CPU evaluation path");
        if (!O->nnz_hess) continue;
        int rc = qck_pipe_ensure_z(c, Z + s.k0 * h->zdim);
        if (rc) { if (c != h) h->err = c->err; return rc; }
        cudaStream_t st = c->stream;
        qck_obj_hess_kernel<<<(unsigned)std::min<long long>(O->dev.own_knots, 148 * 8), 256, 0, st>>>(O->dev, c->dZ, sigma, O->d_hess);
        c->launches += 1;
        QCK_CUDA_TRY(h, cudaMemcpyAsync(O->pin, O->d_hess, sizeof(double) * O->nnz_hess, cudaMemcpyDeviceToHost, st));
    }
    for (auto& s : shards_of(h)) {
        QckObjective* O = static_cast<QckObjective*>(s.h->objective);
        if (!O->nnz_hess) continue;
        QCK_CUDA_TRY(h, cudaSetDevice(s.h->device));
        QCK_CUDA_TRY(h, cudaStreamSynchronize(s.h->stream));
        memcpy(vals + s.k0 * O->dev.hess_knot, O->pin, sizeof(double) * O->nnz_hess);
    }
    return QCK_OK;
}

// ---- FinalUnitaryFidelityConstraint: g(Z) = F(U_T) - min_fidelity >= 0 -----------------------------------------------------
int qck_fidelity_constraint_attach(qck_handle* h, const qck_objective_term* term, double min_fidelity) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !term) return QCK_EINVAL;
    if (term->kind != QCK_OBJ_UNITARY_INFIDELITY) return qck_fail(h, QCK_EINVAL, "the fidelity constraint takes a term of kind QCK_OBJ_UNITARY_INFIDELITY (goal, levels, n_sub)");
    if (!h->children.empty() && h->shard_mode != QCK_SHARD_KNOT) return qck_fail(h, QCK_EINVAL, "the fidelity constraint on a multi-GPU handle needs knot sharding");
    qck_handle* c = shards_of(h).back().h;  // the shard that holds the final knot
    if (!c->objective) {  // constraint without objective terms on this handle: an empty carrier
        c->objective = new QckObjective();
        QckObjective* O0 = static_cast<QckObjective*>(c->objective);
        O0->dev.T = c->T; O0->dev.zdim = c->zdim; O0->dev.dt_off = c->dt_off; O0->dev.dt_fixed = c->dt_fixed; O0->dev.last = 1;
    }
    QckObjective* O = static_cast<QckObjective*>(c->objective);
    if (c->device >= 0) QCK_CUDA_TRY(h, cudaSetDevice(c->device));
    int rc = term_to_dev(c, O, *term, O->con, c->dt_off >= 0);
    if (rc) { if (c != h) h->err = c->err; return rc; }
    O->has_con = true;
    O->con_val = min_fidelity;
    if (c->device >= 0) {
        const long long n = 1 + term->comp_len + (long long)term->comp_len * (term->comp_len + 1) / 2;
        void* p = nullptr;
        QCK_CUDA_TRY(h, cudaMalloc(&p, sizeof(double) * n));
        O->allocs.push_back(p);
        O->d_con = static_cast<double*>(p);
        if (O->pin_doubles < n) {
            if (O->pin) cudaFreeHost(O->pin);
            O->pin_doubles = n;
            QCK_CUDA_TRY(h, cudaHostAlloc((void**)&O->pin, sizeof(double) * n, cudaHostAllocPortable));
        }
    }
    h->con_len = term->comp_len;
    h->con_off = term->comp_off;
    return QCK_OK;
}

// g (1 value) and, if jac != NULL, the Jacobian row over the final knot's state component (comp_len values; their 1-based
// columns are (T-1)*zdim + comp_off + 1 ...); hess != NULL: mu * upper triangle (by column) of the constraint's Hessian
int qck_eval_fidelity_constraint(qck_handle* h, const double* Z, double mu, double* g, double* jac, double* hess) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !Z) return QCK_EINVAL;
    auto sh = shards_of(h);
    qck_handle* c = sh.back().h;
    QckObjective* O = static_cast<QckObjective*>(c->objective);
    if (!O || !O->has_con) return qck_fail(h, QCK_EINVAL, "no fidelity constraint attached");
    if (c->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle: libqcknot has no CPU evaluation path");
    int rc = qck_pipe_ensure_z(c, Z + sh.back().k0 * h->zdim);
    if (rc) { if (c != h) h->err = c->err; return rc; }
    const int n = O->con.comp_len;
    const long long nh = (long long)n * (n + 1) / 2;
    double* dg = O->d_con; double* dj = dg + 1; double* dh = dj + n;
    qck_fid_constraint_kernel<<<1, 256, 0, c->stream>>>(O->con, c->dZ + (c->T - 1) * c->zdim + O->con.comp_off, O->con_val, mu,
                                                        g ? dg : nullptr, jac ? dj : nullptr, hess ? dh : nullptr);
    c->launches += 1;
    QCK_CUDA_TRY(h, cudaMemcpyAsync(O->pin, O->d_con, sizeof(double) * (1 + n + (hess ? nh : 0)), cudaMemcpyDeviceToHost, c->stream));
    QCK_CUDA_TRY(h, cudaStreamSynchronize(c->stream));
    if (g) *g = O->pin[0];
    if (jac) memcpy(jac, O->pin + 1, sizeof(double) * n);
    if (hess) memcpy(hess, O->pin + 1 + n, sizeof(double) * nh);
    return QCK_OK;
}

}  // extern "C"
