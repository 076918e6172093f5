// The handle behind the C-ABI (include/qcknot.h) and the host-side machinery it owns.  Internal to libqcknot.so:
//   qck_host.cpp   create / destroy / structures / kernel launches on device pointers
//   qck_pipe.cpp   host-buffer evaluation: page-locked staging, knot chunks, compact D2H, threaded expansion, Z cache
//   qck_multi.cpp  one handle driving several GPUs (knot / ensemble sharding), NCCL through dlopen
#pragma once
#include <atomic>
#include <complex>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "qck_internal.h"

struct QckInteg {
    int kind = 0, order = 0, N = 0, nd = 0, state_off = 0, state_len = 0, ctrl_off = 0;
    int row_off = 0, dim = 0, nc = 0;
    std::vector<std::complex<double>> Hdrift, Hdrives;
    bool quantum() const { return kind != QCK_DERIVATIVE; }
    bool unitary() const { return kind == QCK_UNITARY_PADE || kind == QCK_UNITARY_EXP; }
    bool pade() const { return kind == QCK_UNITARY_PADE || kind == QCK_KET_PADE; }
};

struct QckClassHost {
    QckClassDev dev{};
    std::vector<int> members;  // integrator indices, ascending
    int member_begin = 0, member_end = 0;
    std::vector<void*> allocs;  // device allocations owned by the class
    QckPlanCache plan;          // launch geometry found at the first launch (no CUDA API calls on later launches)
};

// One run of positions of a value array that this handle writes: out[full + r*len + k] = compact[comp + k],
// 0 <= k < len, 0 <= r < nrep.  nrep > 1 marks a kron(I_N, B) block: it crosses PCIe once.
struct QckOwnSeg {
    int full, comp, len, nrep;
};

// ---- host thread pool (process-wide): staging copies and the expansion of the value arrays ---------------------------
// Two ways in.  A BATCH (submit / wait / parallel_for) is a bag of independent items behind a mutex + condition variable:
// fine for one-off work.  A JOB is one host-buffer call: an ordered list of pieces (staging copies first, then the transfer
// pieces as they arrive), each a bag of items; pieces become ready in order, the caller flips `ready` after the D2H event
// and the workers -- which SPIN while any job is active, no futex on the hot path -- pick the items up within a microsecond.
struct QckBatch {
    std::function<void(int)> fn;
    int n = 0;
    std::atomic<int> next{0}, left{0};
    std::mutex m;
    std::condition_variable cv;
    int users = 0;  // pool workers currently holding this batch (guarded by the pool's mutex)
};
struct QckJob {
    int npieces = 0;
    std::vector<int> nitems;
    std::unique_ptr<std::atomic<int>[]> ready, next, done;
    std::atomic<int> cursor{0};  // first piece that may still have untaken items
    std::function<void(int, int)> fn;  // (piece, item)
    std::atomic<int> users{0};
    void init(const std::vector<int>& items) {
        npieces = (int)items.size();
        nitems = items;
        ready.reset(new std::atomic<int>[npieces]);
        next.reset(new std::atomic<int>[npieces]);
        done.reset(new std::atomic<int>[npieces]);
        for (int i = 0; i < npieces; ++i) { ready[i].store(0); next[i].store(0); done[i].store(0); }
    }
    bool piece_done(int q) const { return done[q].load(std::memory_order_acquire) >= nitems[q]; }
};
class QckPool {
public:
    static QckPool& get(int want_threads);
    void submit(QckBatch* b);  // items run on the pool threads; wait() may be called from any thread
    void wait(QckBatch* b);    // the waiting thread helps with the batch's items
    void parallel_for(int n, const std::function<void(int)>& fn);  // submit + wait
    void begin(QckJob* j);     // workers start spinning on the job's pieces
    void end(QckJob* j);       // every piece must be done; returns when no worker holds the job any more
    static bool try_work(QckJob* j);  // run one item of the first ready piece that has any left (callers help while they wait)
    int size() const { return (int)threads_.size(); }
    ~QckPool();
private:
    explicit QckPool(int n);
    void worker();
    bool run_one(QckBatch* b);
    bool try_batch();
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<QckBatch*> q_;
    std::vector<QckJob*> jobs_;
    std::atomic<unsigned> generation_{0};
    std::atomic<int> n_batches_{0};
    bool stop_ = false;
};

// ---- host-buffer pipeline state of one single-GPU handle -------------------------------------------------------------
struct QckPipe {
    bool ready = false;
    int chunk_knots = 0, piece_knots = 0, n_slots = 0, per_knot = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev;       // one per ring slot
    double* pinZ = nullptr;            // staged copy of Z (T * zdim), page-locked
    double* pinMu = nullptr;           // staged copy of mu
    double* pinC = nullptr;            // D2H ring: n_slots pieces of slot_doubles
    double* dC = nullptr;              // device compact buffer: one compute chunk per stream
    double* pinShared = nullptr;       // ensemble sharding: (T-1) * n_shared partial sums of the shared Hessian entries
    double* dShared = nullptr;
    long long slot_doubles = 0, chunk_doubles = 0;
    int cC[3] = {0, 0, 0};             // compact doubles per knot of F, J, H
    const int* d_src[3] = {nullptr, nullptr, nullptr};  // compact index -> position inside the knot block
    // variant for page-locked caller arrays: long unrepeated runs leave by direct (1-D / 2-D) copies, the compact row holds the rest
    std::vector<QckOwnSeg> own_staged[3], own_direct[3];
    int cS[3] = {0, 0, 0};
    const int* d_src_staged[3] = {nullptr, nullptr, nullptr};
    // cache of the last inputs ("the same Z is presented to F, dF, mu d2F in succession")
    bool z_staged = false, z_on_device = false, mu_on_device = false;
    unsigned valid_mask = 0;           // value arrays on the device that belong to the staged Z (and mu for H)
    long long h2d_bytes = 0, d2h_bytes = 0, cache_hits = 0;  // statistics of the last host-buffer call
};

// Structure-order policy other than CSC (qck_problem_desc.structure_order): the caller-visible structure and the gather tables
// from the canonical (CSC) value arrays the kernels write.
struct QckPublicOrder {
    int order = 0;                         // QCK_ORDER_*
    long long nnzJ = 0, nnzH = 0;          // caller-visible values per knot block
    std::vector<int32_t> Jr, Jc, Hr, Hc;   // caller-visible per-knot structure, 0-based
    std::vector<int> srcJ, srcH;           // caller-visible position -> canonical position (H: >= canonical nnzH = partial column)
    std::vector<long long> shared;         // caller-visible positions with several contributors (empty for PER_INTEGRATOR)
    const int *d_srcJ = nullptr, *d_srcH = nullptr;
    double *dJ = nullptr, *dH = nullptr;   // caller-order device arrays of the host-buffer path
};

struct qck_handle {
    std::string err;
    QckPublicOrder pub;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    long long T = 0;
    int zdim = 0, dt_off = -1, eval_hessian = 1, ib = 0, ie = 0;
    double dt_fixed = 0.0;
    std::vector<QckInteg> integ;
    int dyn = 0;
    long long nnzJ = 0, nnzH = 0;
    std::vector<int32_t> Jr, Jc, Hr, Hc;  // per-knot structure, 0-based, CSC order
    std::vector<QckClassHost> classes;
    std::vector<QckAux> aux;
    const QckAux* d_aux = nullptr;
    std::vector<int> sh_pos, sh_ptr, sh_cols;      // active contributors only (reduce kernel)
    std::vector<long long> shared_positions;       // globally shared positions
    QckReduce red{};
    int npart = 0;
    std::vector<void*> allocs;
    double *dZ = nullptr, *dmu = nullptr, *dF = nullptr, *dJ = nullptr, *dH = nullptr, *dpartial = nullptr;
    int* d_status = nullptr;   // device-side error word (QCK_ST_* bits), checked after every synchronising call
    bool uses_status = false;  // some kernel of this handle can raise a status bit (exponential integrators with a Hessian tape)
    long long launches = 0;
    // positions of F / J / H this handle writes (everything for an unsharded handle), kron blocks marked
    std::vector<QckOwnSeg> own[3];
    bool exclude_shared = false;  // child of an ensemble-sharded handle: shared Hessian entries travel separately
    bool skip_local_reduce = false;  // (set around a launch) the parent sums the partial columns of all GPUs itself
    int host_threads = 0;
    QckPipe pipe;
    void* objective = nullptr;    // QckObjective* (qck_objective.cu): objective / terminal-constraint terms attached to this handle
    int con_off = 0, con_len = 0;
    // multi-GPU parent (n_gpus > 1): structure-only itself, the children do the work
    std::vector<qck_handle*> children;
    std::vector<long long> child_t0;   // KNOT: first block of every child (+ end sentinel)
    int shard_mode = QCK_SHARD_KNOT;
    void* nccl = nullptr;              // QckNccl*, created on first use
    // ENSEMBLE, device-resident: all-reduce by all-read over peer memory
    bool peer_ready = false, peer_ok = false;
    std::vector<QckPeerReduce> peer;   // per child (tables on that child's device)
    std::vector<cudaEvent_t> ev_kernel, ev_reduce;  // per child: partial columns written / read by everybody
    std::vector<double*> gF, gJ, gH;   // per child: gathered (full) value arrays, allocated by qck_gather_device
};

// qck_host.cpp
int qck_fail(qck_handle* h, int code, const char* fmt, ...);
int qck_run(qck_handle* h, uint32_t mask, long long k0, long long nk, const double* dZ, const double* dmu, double* dF, double* dJ,
            double* dH, cudaStream_t st, int slot);  // pointers are array bases; blocks [k0, k0 + nk) are evaluated;
                                                      // slot < QCK_TAPE_SLOTS: scratch set of the stream (launches on different slots may overlap)
#define QCK_TAPE_SLOTS 3
int qck_reorder(qck_handle* h, uint32_t mask, double* dJ_out, double* dH_out, cudaStream_t st);  // canonical h->dJ / h->dH (+ partial columns) -> caller order
int qck_check_status(qck_handle* h);  // after a synchronisation: turns device-side error bits into QCK_ERANGE
int qck_create_single(const qck_problem_desc* d, qck_handle** out, bool exclude_shared);
// qck_pipe.cpp
int qck_pipe_eval(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H);
int qck_pipe_ensure_z(qck_handle* h, const double* Z);  // Z staged + on the device (uploaded on h->stream unless it is already there)
void qck_objective_free(qck_handle* h);                 // qck_objective.cu
void qck_pipe_destroy(qck_handle* h);
void qck_stream_copy(double* dst, const double* src, size_t n);  // non-temporal stores
// qck_multi.cpp
int qck_multi_create(const qck_problem_desc* d, qck_handle* parent);
int qck_multi_eval(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H);
void qck_multi_destroy(qck_handle* h);
void qck_split_knots(long long n_blocks, int n, std::vector<long long>& t0);
void qck_split_integrators(const std::vector<QckInteg>& integ, int n, std::vector<int>& q0);

#define QCK_CUDA_TRY(h, call)                                                                          \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return qck_fail(h, QCK_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
