/* CPU restatement (plain C + pthreads) of the reference's per-knot dynamics evaluator.  TEST INFRASTRUCTURE ONLY:
 * used by tests/ as a second checker and by bench.py as the timed host-core baseline ("cpu_baseline", kind "port").
 * Never linked into or called from libqcknot.so.
 *
 * PARITY UNPINNED (see oracle/knot_oracle.py header): the reference's arithmetic lives in un-vendored Julia
 * packages (QuantumCollocationCore 0.3, /root/reference/Project.toml:31).  This file restates that algorithm in
 * the reference's own formulation:
 *   - real isomorphic arithmetic, G(a) = G_drift + sum_j a_j G_j as a dense 2N x 2N real matrix  
 *   - 4th-order Pade residual (I (x) B) x_{t+1} - (I (x) F) x_t, B/F = I -/+ dt/2 G + dt^2/12 G^2   (SURVEY 8a2;
 *     equation at /root/reference/src/problem_templates/unitary_smooth_pulse_problem.jl:14,30)
 *   - Jacobian blocks I (x) B, -I (x) F, d/da_j via G_j G + G G_j, d/ddt; Hessian of mu^T P with precomputed
 *     anticommutators {G_i, G_j}                                                                      
 *   - DerivativeIntegrator x_{t+1} - x_t - dt dx_t                      (unitary_smooth_pulse_problem.jl:15-16)
 *   - a threaded loop over knots t = 1..T-1 that fills a per-knot dense block and copies its structural entries
 *     one by one into the flat value vector in structure order              (SURVEY 8a6, 3.2 steps 3-5;
 *     surface at /root/reference/test/scripts/integrator_test_1qubit.jl:45-52)
 *
 * Two timings of the same arithmetic (bench.py reports both, the speed-up ratios use the faster one):
 *   literal (tuned = 0)  every product is formed where the formula names it, per knot
 *   tuned   (tuned = 1)  what SURVEY 8a2 says the reference does: the constant anticommutators {G_i, G_j} are computed ONCE
 *                        per call, and G_j G + G G_j is formed once per knot and reused by the Jacobian and the Hessian terms
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define KO_UNITARY_PADE 0
#define KO_KET_PADE 2
#define KO_DERIVATIVE 4

typedef struct {
    int kind, N, n_drives, state_off, ctrl_off, row_off, dim;
    const double* G_drift;  /* 2N x 2N column-major */
    const double* G_drives; /* n_drives of them */
} ko_integ;

/* C (m x n) = alpha * op(A) (m x k) * B (k x n) + beta * C, column-major, op = transpose when ta != 0 */
static void gemm(int ta, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                 double beta, double* C, int ldc) {
    for (int j = 0; j < n; ++j) {
        double* c = C + (size_t)ldc * j;
        if (beta == 0.0) for (int i = 0; i < m; ++i) c[i] = 0.0;
        else if (beta != 1.0) for (int i = 0; i < m; ++i) c[i] *= beta;
        for (int l = 0; l < k; ++l) {
            double b = alpha * B[l + (size_t)ldb * j];
            if (!ta) {
                const double* a = A + (size_t)lda * l;
                for (int i = 0; i < m; ++i) c[i] += a[i] * b;
            } else {
                for (int i = 0; i < m; ++i) c[i] += A[l + (size_t)lda * i] * b;
            }
        }
    }
}

static double dot(const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

typedef struct {
    int n_integ; const ko_integ* integ; long T; int zdim, dt_off; double dt_fixed; const double* Z; const double* mu;
    int dyn; long nnzJ; const int *Jr, *Jc; long nnzH; const int *Hr, *Hc; double *F, *J, *H;
    int maxn2, maxnd; long t_begin, t_end; int failed;
    int tuned; const double* const* AC; /* tuned: per integrator, the anticommutators {G_i, G_j}, pairs (j, i <= j) */
} ko_job;

/* one host thread: knot blocks [t_begin, t_end)  (the reference's Threads.@threads loop body) */
static void* ko_worker(void* arg) {
    ko_job* jb_ = (ko_job*)arg;
    const int n_integ = jb_->n_integ; const ko_integ* integ = jb_->integ; const int zdim = jb_->zdim, dt_off = jb_->dt_off;
    const double dt_fixed = jb_->dt_fixed; const double* Z = jb_->Z; const double* mu = jb_->mu; const int dyn = jb_->dyn;
    const long nnzJ = jb_->nnzJ, nnzH = jb_->nnzH; const int *Jr = jb_->Jr, *Jc = jb_->Jc, *Hr = jb_->Hr, *Hc = jb_->Hc;
    double *F = jb_->F, *J = jb_->J, *H = jb_->H; const int maxn2 = jb_->maxn2, maxnd = jb_->maxnd;
    const int free_time = dt_off >= 0;
    const size_t jb = (size_t)dyn * 2 * zdim, hb = (size_t)4 * zdim * zdim;
    int failed = 0;
    {
        const int n2m = maxn2, mm = n2m * n2m;
        /* per-thread scratch: dense per-knot blocks + small matrices */
        double* Jb = J ? (double*)calloc(jb, sizeof(double)) : NULL;
        double* Hb = H ? (double*)calloc(hb, sizeof(double)) : NULL;
        const int tuned = jb_->tuned;
        double* w = (double*)malloc(sizeof(double) * (size_t)mm * (8 + 3 * maxnd + 4));
        if ((J && !Jb) || (H && !Hb) || !w) failed = 1;
        if (!failed) {
        double *G = w, *G2 = G + mm, *Fm = G2 + mm, *Bm = Fm + mm, *F1 = Bm + mm, *B1 = F1 + mm, *tmp = B1 + mm,
               *tmp2 = tmp + mm, *dFj = tmp2 + mm /* maxnd */, *dBj = dFj + (size_t)mm * maxnd /* maxnd */,
               *v1 = dBj + (size_t)mm * maxnd, *v2 = v1 + mm, *v3 = v2 + mm, *v4 = v3 + mm,
               *CG = v4 + mm /* maxnd: G_j G + G G_j of this knot (tuned) */;
        for (long t = jb_->t_begin; t < jb_->t_end; ++t) {
            const double* zt = Z + t * zdim;
            const double* zt1 = zt + zdim;
            const double dt = free_time ? zt[dt_off] : dt_fixed;
            if (H) /* control x control / control x dt / dt x dt entries are sums over the integrators that share them */
                for (int q = 0; q < n_integ; ++q) {
                    if (integ[q].kind == KO_DERIVATIVE) continue;
                    const int co = integ[q].ctrl_off, nd = integ[q].n_drives, ld = 2 * zdim;
                    for (int j = 0; j < nd; ++j) {
                        for (int i2 = 0; i2 <= j; ++i2) Hb[co + i2 + (size_t)ld * (co + j)] = 0.0;
                        if (free_time) Hb[(co + j < dt_off ? co + j : dt_off) + (size_t)ld * (co + j < dt_off ? dt_off : co + j)] = 0.0;
                    }
                    if (free_time) Hb[dt_off + (size_t)ld * dt_off] = 0.0;
                }
            for (int q = 0; q < n_integ; ++q) {
                const ko_integ* I = &integ[q];
                const int R0 = I->row_off;
                if (I->kind == KO_DERIVATIVE) {
                    for (int i = 0; i < I->dim; ++i) {
                        const int x = I->state_off + i, dx = I->ctrl_off + i;
                        if (F) F[t * dyn + R0 + i] = zt1[x] - zt[x] - dt * zt[dx];
                        if (J) {
                            Jb[R0 + i + (size_t)dyn * x] = -1.0;
                            Jb[R0 + i + (size_t)dyn * (zdim + x)] = 1.0;
                            Jb[R0 + i + (size_t)dyn * dx] = -dt;
                            if (free_time) Jb[R0 + i + (size_t)dyn * dt_off] = -zt[dx];
                        }
                        if (H && free_time) {
                            int r = dx < dt_off ? dx : dt_off, c = dx < dt_off ? dt_off : dx;
                            Hb[r + (size_t)2 * zdim * c] = -mu[t * dyn + R0 + i];
                        }
                    }
                    continue;
                }
                const int N = I->N, n2 = 2 * N, nc = I->kind == KO_UNITARY_PADE ? N : 1, nd = I->n_drives;
                const int nn = n2 * n2, so = I->state_off, co = I->ctrl_off;
                const double* W0 = zt + so;   /* 2N x nc, column-major == the iso-vec itself */
                const double* W1 = zt1 + so;
                const double* Mu = mu ? mu + t * dyn + R0 : NULL;
                /* G(a), G^2, F, B and their dt-derivatives */
                memcpy(G, I->G_drift, sizeof(double) * nn);
                for (int j = 0; j < nd; ++j) {
                    const double a = zt[co + j];
                    const double* Gj = I->G_drives + (size_t)nn * j;
                    for (int e = 0; e < nn; ++e) G[e] += a * Gj[e];
                }
                gemm(0, n2, n2, n2, 1.0, G, n2, G, n2, 0.0, G2, n2);
                for (int e = 0; e < nn; ++e) {
                    const double id = (e % n2 == e / n2) ? 1.0 : 0.0;
                    Fm[e] = id + 0.5 * dt * G[e] + dt * dt / 12.0 * G2[e];
                    Bm[e] = id - 0.5 * dt * G[e] + dt * dt / 12.0 * G2[e];
                    F1[e] = 0.5 * G[e] + dt / 6.0 * G2[e];
                    B1[e] = -0.5 * G[e] + dt / 6.0 * G2[e];
                }
                /* residual */
                gemm(0, n2, nc, n2, 1.0, Bm, n2, W1, n2, 0.0, v1, n2);
                gemm(0, n2, nc, n2, -1.0, Fm, n2, W0, n2, 1.0, v1, n2);
                if (F) memcpy(F + t * dyn + R0, v1, sizeof(double) * n2 * nc);
                if (!J && !H) continue;
                /* dF_j = dt/2 G_j + dt^2/12 (G_j G + G G_j) */
                for (int j = 0; j < nd; ++j) {
                    const double* Gj = I->G_drives + (size_t)nn * j;
                    double* cg = tuned ? CG + (size_t)nn * j : tmp;
                    gemm(0, n2, n2, n2, 1.0, Gj, n2, G, n2, 0.0, cg, n2);
                    gemm(0, n2, n2, n2, 1.0, G, n2, Gj, n2, 1.0, cg, n2);
                    for (int e = 0; e < nn; ++e) {
                        dFj[(size_t)nn * j + e] = 0.5 * dt * Gj[e] + dt * dt / 12.0 * cg[e];
                        dBj[(size_t)nn * j + e] = -0.5 * dt * Gj[e] + dt * dt / 12.0 * cg[e];
                    }
                }
                if (J) {
                    for (int cb = 0; cb < nc; ++cb)
                        for (int r = 0; r < n2; ++r)
                            for (int qq = 0; qq < n2; ++qq) {
                                Jb[R0 + cb * n2 + qq + (size_t)dyn * (so + cb * n2 + r)] = -Fm[qq + n2 * r];
                                Jb[R0 + cb * n2 + qq + (size_t)dyn * (zdim + so + cb * n2 + r)] = Bm[qq + n2 * r];
                            }
                    for (int j = 0; j < nd; ++j) {
                        gemm(0, n2, nc, n2, 1.0, dBj + (size_t)nn * j, n2, W1, n2, 0.0, v2, n2);
                        gemm(0, n2, nc, n2, -1.0, dFj + (size_t)nn * j, n2, W0, n2, 1.0, v2, n2);
                        memcpy(Jb + R0 + (size_t)dyn * (co + j), v2, sizeof(double) * n2 * nc);
                    }
                    if (free_time) {
                        gemm(0, n2, nc, n2, 1.0, B1, n2, W1, n2, 0.0, v2, n2);
                        gemm(0, n2, nc, n2, -1.0, F1, n2, W0, n2, 1.0, v2, n2);
                        memcpy(Jb + R0 + (size_t)dyn * dt_off, v2, sizeof(double) * n2 * nc);
                    }
                }
                if (H) {
                    const int dim = n2 * nc, ld = 2 * zdim;
                    for (int i = 0; i < dim; ++i) v4[i] = W1[i] - W0[i];
                    for (int j = 0; j < nd; ++j) {
                        gemm(1, n2, nc, n2, -1.0, dFj + (size_t)nn * j, n2, Mu, n2, 0.0, v2, n2); /* -dF_j^T Mu */
                        gemm(1, n2, nc, n2, 1.0, dBj + (size_t)nn * j, n2, Mu, n2, 0.0, v3, n2);  /* dB_j^T Mu */
                        for (int i = 0; i < dim; ++i) {
                            int r = so + i, c = co + j;
                            if (r <= c) Hb[r + (size_t)ld * c] = v2[i]; else Hb[c + (size_t)ld * r] = v2[i];
                            Hb[co + j + (size_t)ld * (zdim + so + i)] = v3[i];
                        }
                        for (int i2 = 0; i2 <= j; ++i2) { /* {G_i, G_j} (W1 - W0), weight dt^2/12 */
                            const double* Gi = I->G_drives + (size_t)nn * i2;
                            const double* Gj = I->G_drives + (size_t)nn * j;
                            const double* ac = tmp;
                            if (tuned) ac = jb_->AC[q] + (size_t)nn * (j * (j + 1) / 2 + i2);
                            else {
                                gemm(0, n2, n2, n2, 1.0, Gi, n2, Gj, n2, 0.0, tmp, n2);
                                gemm(0, n2, n2, n2, 1.0, Gj, n2, Gi, n2, 1.0, tmp, n2);
                            }
                            gemm(0, n2, nc, n2, dt * dt / 12.0, ac, n2, v4, n2, 0.0, v2, n2);
                            Hb[co + i2 + (size_t)ld * (co + j)] += dot(Mu, v2, dim);
                        }
                        if (free_time) { /* d/ddt of dB_j W1 - dF_j W0 */
                            const double* Gj = I->G_drives + (size_t)nn * j;
                            if (tuned) memcpy(tmp, CG + (size_t)nn * j, sizeof(double) * nn);
                            else {
                                gemm(0, n2, n2, n2, 1.0, Gj, n2, G, n2, 0.0, tmp, n2);
                                gemm(0, n2, n2, n2, 1.0, G, n2, Gj, n2, 1.0, tmp, n2);
                            }
                            for (int e = 0; e < nn; ++e) { tmp2[e] = 0.5 * Gj[e] + dt / 6.0 * tmp[e]; tmp[e] = -0.5 * Gj[e] + dt / 6.0 * tmp[e]; }
                            gemm(0, n2, nc, n2, 1.0, tmp, n2, W1, n2, 0.0, v2, n2);
                            gemm(0, n2, nc, n2, -1.0, tmp2, n2, W0, n2, 1.0, v2, n2);
                            int r = co + j, c = dt_off;
                            double val = dot(Mu, v2, dim);
                            if (r <= c) Hb[r + (size_t)ld * c] += val; else Hb[c + (size_t)ld * r] += val;
                        }
                    }
                    if (free_time) {
                        gemm(1, n2, nc, n2, -1.0, F1, n2, Mu, n2, 0.0, v2, n2);
                        gemm(1, n2, nc, n2, 1.0, B1, n2, Mu, n2, 0.0, v3, n2);
                        for (int i = 0; i < dim; ++i) {
                            int r = so + i, c = dt_off;
                            if (r <= c) Hb[r + (size_t)ld * c] = v2[i]; else Hb[c + (size_t)ld * r] = v2[i];
                            Hb[dt_off + (size_t)ld * (zdim + so + i)] = v3[i];
                        }
                        gemm(0, n2, nc, n2, 1.0 / 6.0, G2, n2, v4, n2, 0.0, v2, n2); /* d2/ddt2 = 1/6 G^2 (W1 - W0) */
                        Hb[dt_off + (size_t)ld * dt_off] += dot(Mu, v2, dim);
                    }
                }
            }
            if (J) for (long k = 0; k < nnzJ; ++k) J[t * nnzJ + k] = Jb[Jr[k] + (size_t)dyn * Jc[k]];
            if (H) for (long k = 0; k < nnzH; ++k) H[t * nnzH + k] = Hb[Hr[k] + (size_t)2 * zdim * Hc[k]];
        }
        }
        free(Jb); free(Hb); free(w);
    }
    jb_->failed = failed;
    return NULL;
}

/* evaluates every integrator on knot blocks [0, T-1); F/J/H may be NULL.  Jr/Jc/Hr/Hc: per-knot structure
 * (0-based row, col inside the dyn x 2zdim / 2zdim x 2zdim block).  nthreads <= 0: all online cores.
 * tuned: see the header.  Returns 0, or -1 on allocation failure. */
int ko_eval2(int n_integ, const ko_integ* integ, long T, int zdim, int dt_off, double dt_fixed, const double* Z,
             const double* mu, int dyn, long nnzJ, const int* Jr, const int* Jc, long nnzH, const int* Hr,
             const int* Hc, double* F, double* J, double* H, int nthreads, int tuned) {
    int maxn2 = 2, maxnd = 1;
    for (int q = 0; q < n_integ; ++q)
        if (integ[q].kind != KO_DERIVATIVE) {
            if (2 * integ[q].N > maxn2) maxn2 = 2 * integ[q].N;
            if (integ[q].n_drives > maxnd) maxnd = integ[q].n_drives;
        }
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    long nb = T - 1;
    if (nthreads > nb) nthreads = nb > 0 ? (int)nb : 1;
    /* tuned: the constant anticommutators, once per call */
    double** AC = NULL;
    if (tuned && H) {
        AC = (double**)calloc((size_t)n_integ, sizeof(double*));
        if (!AC) return -1;
        for (int q = 0; q < n_integ; ++q) {
            if (integ[q].kind == KO_DERIVATIVE) continue;
            const int n2 = 2 * integ[q].N, nn = n2 * n2, nd = integ[q].n_drives;
            AC[q] = (double*)malloc(sizeof(double) * (size_t)nn * (nd * (nd + 1) / 2));
            if (!AC[q]) return -1;
            for (int j = 0; j < nd; ++j)
                for (int i2 = 0; i2 <= j; ++i2) {
                    double* ac = AC[q] + (size_t)nn * (j * (j + 1) / 2 + i2);
                    const double* Gi = integ[q].G_drives + (size_t)nn * i2;
                    const double* Gj = integ[q].G_drives + (size_t)nn * j;
                    gemm(0, n2, n2, n2, 1.0, Gi, n2, Gj, n2, 0.0, ac, n2);
                    gemm(0, n2, n2, n2, 1.0, Gj, n2, Gi, n2, 1.0, ac, n2);
                }
        }
    }
    ko_job jobs[256];
    pthread_t tid[256];
    for (int i = 0; i < nthreads; ++i) {
        ko_job j = {n_integ, integ, T, zdim, dt_off, dt_fixed, Z, mu, dyn, nnzJ, Jr, Jc, nnzH, Hr, Hc, F, J, H,
                    maxn2, maxnd, nb * i / nthreads, nb * (i + 1) / nthreads, 0, tuned && AC, (const double* const*)AC};
        jobs[i] = j;
    }
    for (int i = 1; i < nthreads; ++i) pthread_create(&tid[i], NULL, ko_worker, &jobs[i]);
    ko_worker(&jobs[0]);
    int failed = jobs[0].failed;
    for (int i = 1; i < nthreads; ++i) { pthread_join(tid[i], NULL); failed |= jobs[i].failed; }
    if (AC) {
        for (int q = 0; q < n_integ; ++q) free(AC[q]);
        free(AC);
    }
    return failed ? -1 : 0;
}

int ko_eval(int n_integ, const ko_integ* integ, long T, int zdim, int dt_off, double dt_fixed, const double* Z,
            const double* mu, int dyn, long nnzJ, const int* Jr, const int* Jc, long nnzH, const int* Hr,
            const int* Hc, double* F, double* J, double* H, int nthreads) {
    return ko_eval2(n_integ, integ, T, zdim, dt_off, dt_fixed, Z, mu, dyn, nnzJ, Jr, Jc, nnzH, Hr, Hc, F, J, H, nthreads, 0);
}

int ko_num_cores(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
