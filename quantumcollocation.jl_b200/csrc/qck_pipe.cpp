// Host-buffer evaluation path of libqcknot.so: what the MOI callbacks eval_constraint / eval_constraint_jacobian /
// eval_hessian_lagrangian reach through qck_eval_residual / _jacobian / _hessian / _all.
//
// The caller's arrays are ordinary host memory (Ipopt owns them).  Per call:
//   1. Z (and mu) are compared with / copied into the handle's page-locked staging copy by the host threads.  An unchanged
//      Z is not uploaded again and device-resident results that belong to it are reused ("the same Z is presented to F,
//      dF, mu d2F in succession").
//   2. The knot range is cut into chunks.  Per chunk, on one of two streams: H2D of the chunk's knots, the quantum kernels
//      on that sub-range, a pack kernel that gathers the NON-REDUNDANT positions of the value arrays (every kron(I_N, B)
//      block once instead of N times: 3,303 instead of 8,487 doubles per knot of the two-transmon CZ problem), D2H of the
//      compact chunk into a ring of page-locked slots.  H2D, kernels and D2H of neighbouring chunks overlap.
//   3. As chunks arrive, the host threads expand them into the caller's arrays in the solver's structure order
//      (non-temporal stores; kron blocks written N times from one copy).  Results are bit-identical to a plain D2H copy of
//      the device value arrays.
#include <time.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

#include <immintrin.h>

#include "qck_handle.h"

// ---- thread pool ---------------------------------------------------------------------------------------------------
QckPool::QckPool(int n) {
    for (int i = 0; i < n; ++i) threads_.emplace_back([this] { worker(); });
}
QckPool::~QckPool() {
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
}
QckPool& QckPool::get(int want) {
    static std::mutex gm;
    static std::unique_ptr<QckPool> pool;
    std::lock_guard<std::mutex> lk(gm);
    if (!pool) {
        int n = want;
        if (const char* e = getenv("QCK_HOST_THREADS")) n = atoi(e);
        if (n <= 0) n = (int)std::thread::hardware_concurrency();
        n = std::max(1, std::min(n, 32));
        pool.reset(new QckPool(n));
    }
    return *pool;
}
bool QckPool::run_one(QckBatch* b) {
    const int i = b->next.fetch_add(1, std::memory_order_relaxed);
    if (i >= b->n) return false;
    b->fn(i);
    if (b->left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(b->m);
        b->cv.notify_all();
    }
    return true;
}
bool QckPool::try_work(QckJob* j) {
    int q = j->cursor.load(std::memory_order_acquire);
    while (q < j->npieces) {
        if (!j->ready[q].load(std::memory_order_acquire)) return false;  // pieces become ready in order
        const int i = j->next[q].fetch_add(1, std::memory_order_relaxed);
        if (i < j->nitems[q]) {
            j->fn(q, i);
            j->done[q].fetch_add(1, std::memory_order_release);
            return true;
        }
        int expected = q;
        j->cursor.compare_exchange_strong(expected, q + 1, std::memory_order_acq_rel);
        q = j->cursor.load(std::memory_order_acquire);
    }
    return false;
}
// one item of the front batch, if any (workers keep serving batches while jobs are active)
bool QckPool::try_batch() {
    if (n_batches_.load(std::memory_order_acquire) == 0) return false;
    QckBatch* b = nullptr;
    {
        std::lock_guard<std::mutex> lk(m_);
        while (!q_.empty() && q_.front()->next.load(std::memory_order_relaxed) >= q_.front()->n) { q_.pop_front(); n_batches_.fetch_sub(1); }
        if (q_.empty()) return false;
        b = q_.front();
        ++b->users;  // keeps the batch alive while this worker may still touch it (see wait())
    }
    bool any = false;
    while (run_one(b)) any = true;
    std::lock_guard<std::mutex> lk(m_);
    --b->users;
    return any;
}
void QckPool::worker() {
    std::vector<QckJob*> mine;
    for (;;) {
        unsigned gen;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [this] { return stop_ || !q_.empty() || !jobs_.empty(); });
            if (stop_) return;
            mine = jobs_;
            for (QckJob* j : mine) j->users.fetch_add(1);
            gen = generation_.load();
        }
        // spin while the set of active jobs is unchanged: no sleeping between the pieces of a host-buffer call
        int idle = 0;
        do {
            bool any = try_batch();
            for (QckJob* j : mine) any = try_work(j) || any;
            if (any) idle = 0;
            else if (++idle < 64) _mm_pause();
            else if (mine.empty()) break;  // batches only: back to the condition variable
            else if (idle > 20000) { std::this_thread::yield(); idle = 64; }
        } while (generation_.load(std::memory_order_acquire) == gen);
        for (QckJob* j : mine) j->users.fetch_sub(1);
    }
}
void QckPool::submit(QckBatch* b) {
    b->next.store(0);
    b->left.store(b->n);
    if (b->n <= 0) return;
    {
        std::lock_guard<std::mutex> lk(m_);
        q_.push_back(b);
        n_batches_.fetch_add(1);
    }
    cv_.notify_all();
}
void QckPool::wait(QckBatch* b) {
    while (run_one(b)) {}
    {
        std::unique_lock<std::mutex> lk(b->m);
        b->cv.wait(lk, [b] { return b->left.load(std::memory_order_acquire) <= 0; });
    }
    // the batch may still sit in the queue (all items taken) and workers may still hold it: take it out, then let them leave
    std::unique_lock<std::mutex> lk(m_);
    for (auto it = q_.begin(); it != q_.end(); ++it)
        if (*it == b) { q_.erase(it); n_batches_.fetch_sub(1); break; }
    while (b->users > 0) {
        lk.unlock();
        std::this_thread::yield();
        lk.lock();
    }
}
void QckPool::parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (n == 1) { fn(0); return; }
    QckBatch b;
    b.fn = fn;
    b.n = n;
    submit(&b);
    wait(&b);
}
void QckPool::begin(QckJob* j) {
    {
        std::lock_guard<std::mutex> lk(m_);
        jobs_.push_back(j);
        generation_.fetch_add(1);
    }
    cv_.notify_all();
}
void QckPool::end(QckJob* j) {
    {
        std::lock_guard<std::mutex> lk(m_);
        jobs_.erase(std::remove(jobs_.begin(), jobs_.end(), j), jobs_.end());
        generation_.fetch_add(1);
    }
    while (j->users.load(std::memory_order_acquire) > 0) _mm_pause();
}

// ---- streaming copy: the destination (hundreds of MB of value arrays) is written once and not read back by us ---------
// Full cache lines leave as one non-temporal store where the CPU has AVX-512 (a line never sits half-filled in a
// write-combining buffer), 32-byte stores with AVX2, 16-byte stores otherwise.
namespace {
void stream_copy_sse2(double* dst, const double* src, size_t n) {
    size_t i = 0;
    if (n && (reinterpret_cast<uintptr_t>(dst) & 8)) {
        _mm_stream_si64(reinterpret_cast<long long*>(dst), reinterpret_cast<const long long*>(src)[0]);
        i = 1;
    }
    for (; i + 4 <= n; i += 4) {
        const __m128d a = _mm_loadu_pd(src + i), b = _mm_loadu_pd(src + i + 2);
        _mm_stream_pd(dst + i, a);
        _mm_stream_pd(dst + i + 2, b);
    }
    for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    if (i < n) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
}
__attribute__((target("avx2"))) void stream_copy_avx2(double* dst, const double* src, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) {
        _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
        ++i;
    }
    for (; i + 8 <= n; i += 8) {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        _mm256_stream_pd(dst + i, a);
        _mm256_stream_pd(dst + i + 4, b);
    }
    for (; i + 4 <= n; i += 4) _mm256_stream_pd(dst + i, _mm256_loadu_pd(src + i));
    for (; i < n; ++i) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
}
__attribute__((target("avx512f"))) void stream_copy_avx512(double* dst, const double* src, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) {
        _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
        ++i;
    }
    for (; i + 16 <= n; i += 16) {
        const __m512d a = _mm512_loadu_pd(src + i), b = _mm512_loadu_pd(src + i + 8);
        _mm512_stream_pd(dst + i, a);
        _mm512_stream_pd(dst + i + 8, b);
    }
    for (; i + 8 <= n; i += 8) _mm512_stream_pd(dst + i, _mm512_loadu_pd(src + i));
    for (; i < n; ++i) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
}
void stream_copy_plain(double* dst, const double* src, size_t n) { memcpy(dst, src, n * sizeof(double)); }  // cacheable stores (A/B knob)
typedef void (*stream_copy_fn)(double*, const double*, size_t);
stream_copy_fn pick_stream_copy() {
    const char* e = getenv("QCK_STREAM_ISA");  // "sse2" | "avx2" | "avx512" | "plain" (development knob)
    if (e && !strcmp(e, "plain")) return stream_copy_plain;
    __builtin_cpu_init();
    const bool a512 = __builtin_cpu_supports("avx512f"), a2 = __builtin_cpu_supports("avx2");
    if (e && !strcmp(e, "sse2")) return stream_copy_sse2;
    if (e && !strcmp(e, "avx2") && a2) return stream_copy_avx2;
    if (a512) return stream_copy_avx512;
    if (a2) return stream_copy_avx2;
    return stream_copy_sse2;
}
const stream_copy_fn g_stream_copy = pick_stream_copy();
}  // namespace

void qck_stream_copy(double* dst, const double* src, size_t n) {
    if (n < 8) {
        for (size_t i = 0; i < n; ++i) dst[i] = src[i];
        return;
    }
    g_stream_copy(dst, src, n);
}

namespace {

// Two granularities.  A COMPUTE CHUNK is what one H2D + kernel pass + pack covers (large enough to fill the GPU); its compact
// values then leave in TRANSFER PIECES, each one D2H copy + one event + one batch of expansion work, small enough that the
// host threads read a piece while it is still in the last-level cache the DMA engine wrote it into.
constexpr long long kChunkBytes = 24ll << 20;  // compact bytes per compute chunk
constexpr long long kPieceBytes = 2ll << 20;   // compact bytes per transfer piece
constexpr int kRing = 24;                      // page-locked ring depth in pieces

long long env_ll(const char* name, long long dflt) {
    const char* e = getenv(name);
    return e && *e ? atoll(e) : dflt;
}

int pipe_init(qck_handle* h) {
    QckPipe& P = h->pipe;
    if (P.ready) return QCK_OK;
    const long long nk = h->T - 1;
    long long per_knot = 0;
    for (int a = 0; a < 3; ++a) {
        int c = 0;
        for (auto& s : h->own[a]) c += s.len;
        P.cC[a] = c;
        per_knot += c;
    }
    P.per_knot = (int)per_knot;
    long long ck = env_ll("QCK_CHUNK_KNOTS", per_knot > 0 ? kChunkBytes / (8 * per_knot) : nk);
    ck = std::max<long long>(ck, 64);
    ck = (ck + 1) & ~1ll;  // even: every chunk starts at an even block (16-byte aligned bases for even strides)
    if (ck >= nk) ck = nk;
    long long pk = env_ll("QCK_PIECE_KNOTS", per_knot > 0 ? kPieceBytes / (8 * per_knot) : ck);
    pk = std::min(std::max<long long>(pk, 8), ck);
    P.chunk_knots = (int)ck;
    P.piece_knots = (int)pk;
    const long long pieces_per_chunk = (ck + pk - 1) / pk;
    const long long total_pieces = ((nk + ck - 1) / ck) * pieces_per_chunk;
    P.n_slots = (int)std::min<long long>(env_ll("QCK_RING", kRing), total_pieces);
    P.slot_doubles = (pk * per_knot + 1) & ~1ll;
    P.chunk_doubles = (ck * per_knot + 1) & ~1ll;
    QCK_CUDA_TRY(h, cudaSetDevice(h->device));
    for (int i = 0; i < 2; ++i) QCK_CUDA_TRY(h, cudaStreamCreateWithFlags(&P.st[i], cudaStreamNonBlocking));
    P.ev.resize(P.n_slots);
    for (auto& e : P.ev) QCK_CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    QCK_CUDA_TRY(h, cudaHostAlloc((void**)&P.pinZ, sizeof(double) * h->T * h->zdim, cudaHostAllocPortable));
    QCK_CUDA_TRY(h, cudaHostAlloc((void**)&P.pinMu, sizeof(double) * std::max<long long>(nk * h->dyn, 1), cudaHostAllocPortable));
    QCK_CUDA_TRY(h, cudaHostAlloc((void**)&P.pinC, sizeof(double) * std::max<long long>(P.slot_doubles * P.n_slots, 2), cudaHostAllocPortable));
    QCK_CUDA_TRY(h, cudaMalloc((void**)&P.dC, sizeof(double) * std::max<long long>(P.chunk_doubles * 2, 2)));  // one region per stream
    for (int a = 0; a < 3; ++a) {
        if (P.cC[a] == 0) continue;
        std::vector<int> src;
        src.reserve(P.cC[a]);
        for (auto& s : h->own[a])
            for (int k = 0; k < s.len; ++k) src.push_back(s.full + k);
        int* d = nullptr;
        QCK_CUDA_TRY(h, cudaMalloc((void**)&d, sizeof(int) * src.size()));
        QCK_CUDA_TRY(h, cudaMemcpy(d, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice));
        P.d_src[a] = d;
    }
    for (int a = 0; a < 3; ++a) {  // split for page-locked caller arrays: runs of >= 4 KB without repetition go direct
        int c = 0;
        for (auto& g : h->own[a]) {
            if (g.nrep == 1 && g.len >= 512) P.own_direct[a].push_back(g);
            else { P.own_staged[a].push_back({g.full, c, g.len, g.nrep}); c += g.len; }
        }
        P.cS[a] = c;
        if (P.own_direct[a].empty() || P.own_direct[a].size() > 4) {  // nothing to gain (or hundreds of copies per chunk): one variant only
            P.own_direct[a].clear();
            P.own_staged[a] = h->own[a];
            P.cS[a] = P.cC[a];
            P.d_src_staged[a] = P.d_src[a];
            continue;
        }
        if (c == 0) continue;
        std::vector<int> src;
        for (auto& g : P.own_staged[a])
            for (int k = 0; k < g.len; ++k) src.push_back(g.full + k);
        int* d = nullptr;
        QCK_CUDA_TRY(h, cudaMalloc((void**)&d, sizeof(int) * src.size()));
        QCK_CUDA_TRY(h, cudaMemcpy(d, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice));
        P.d_src_staged[a] = d;
    }
    if (h->exclude_shared && !h->sh_pos.empty()) {
        const long long n = nk * (long long)h->sh_pos.size();
        QCK_CUDA_TRY(h, cudaHostAlloc((void**)&P.pinShared, sizeof(double) * n, cudaHostAllocPortable));
        QCK_CUDA_TRY(h, cudaMalloc((void**)&P.dShared, sizeof(double) * n));
    }
    P.ready = true;
    return QCK_OK;
}

// is `src` bitwise the staged copy?  (read-only, all host threads; every thread stops at its first difference)
bool same_as_staged(QckPool& pool, const double* staged, const double* src, long long n) {
    const long long piece = 1ll << 16;  // 512 KB
    std::atomic<int> differs{0};
    pool.parallel_for((int)((n + piece - 1) / piece), [&](int i) {
        if (differs.load(std::memory_order_relaxed)) return;
        const long long o = (long long)i * piece, m = std::min(piece, n - o);
        if (memcmp(staged + o, src + o, sizeof(double) * m) != 0) differs.store(1, std::memory_order_relaxed);
    });
    return differs.load() == 0;
}

// knots [k0, k1) of one value array: compact rows (`stride` doubles per knot, this array's part at the row start `comp`) ->
// the caller's array
void expand_knots(const std::vector<QckOwnSeg>& segs, const double* comp, long long stride, double* out, long long nnz, long long k0, long long k1) {
    for (long long t = k0; t < k1; ++t) {
        const double* s = comp + t * stride;
        double* d = out + t * nnz;
        for (const QckOwnSeg& g : segs)
            for (int r = 0; r < g.nrep; ++r) qck_stream_copy(d + g.full + (long long)r * g.len, s + g.comp, (size_t)g.len);
    }
    _mm_sfence();  // the non-temporal stores are globally visible before this work item is reported done
}

// page-locked host memory the copy engine can write directly (qck_host_register / cudaHostRegister / cudaHostAlloc)?
bool is_page_locked(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

}  // namespace

int qck_pipe_eval(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) {
    QCK_CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = pipe_init(h);
    if (rc) return rc;
    QckPipe& P = h->pipe;
    QckPool& pool = QckPool::get(h->host_threads);
    static const bool timing = getenv("QCK_PIPE_TIMING") != nullptr;
    const double t_begin = timing ? now_ms() : 0.0;
    const long long nk = h->T - 1;
    const long long nnz[3] = {h->dyn, h->nnzJ, h->nnzH};
    double* outs[3] = {F, J, (H && h->eval_hessian) ? H : nullptr};
    double* dfull[3] = {h->dF, h->dJ, h->dH};
    unsigned need = (outs[0] ? QCK_EVAL_F : 0) | (outs[1] ? QCK_EVAL_J : 0) | (outs[2] ? QCK_EVAL_H : 0);
    P.h2d_bytes = P.d2h_bytes = 0;
    if (!need) return QCK_OK;
    {   // SMALL PROBLEMS (README scale: Hadamard, T = 50): the whole call is one piece of less than 1 MB -- one stream, no host threads,
        // no events; what is left is the launch and PCIe latency itself
        static const bool no_small = getenv("QCK_NO_SMALL_PATH") != nullptr;
        const long long out_doubles = nk * (nnz[0] + nnz[1] + nnz[2]);
        if (!no_small && !P.dShared && nk <= P.piece_knots && out_doubles <= (128ll << 10)) {
            static const bool no_cache = getenv("QCK_NO_CACHE") != nullptr;
            static const bool speculate = getenv("QCK_NO_SPECULATE") == nullptr;
            const size_t zb = sizeof(double) * (size_t)h->T * h->zdim, mb = sizeof(double) * (size_t)nk * h->dyn;
            const bool z_same = !no_cache && P.z_staged && P.z_on_device && memcmp(P.pinZ, Z, zb) == 0;
            if (!z_same) { P.z_on_device = false; P.valid_mask = 0; }
            bool mu_same = true;
            if (need & QCK_EVAL_H) {
                mu_same = !no_cache && P.mu_on_device && memcmp(P.pinMu, mu, mb) == 0;
                if (!mu_same) { P.mu_on_device = false; P.valid_mask &= ~QCK_EVAL_H; }
            }
            unsigned compute = need & ~P.valid_mask;
            if (speculate && (compute & QCK_EVAL_F) && !(P.valid_mask & QCK_EVAL_J) && !(need & QCK_EVAL_H)) compute |= QCK_EVAL_J;
            if (!compute) ++P.cache_hits;
            cudaStream_t st = P.st[0];
            P.z_staged = false;
            if (!z_same) {
                memcpy(P.pinZ, Z, zb);
                QCK_CUDA_TRY(h, cudaMemcpyAsync(h->dZ, P.pinZ, zb, cudaMemcpyHostToDevice, st));
                P.h2d_bytes += (long long)zb;
            }
            if ((need & QCK_EVAL_H) && !mu_same) {
                memcpy(P.pinMu, mu, mb);
                QCK_CUDA_TRY(h, cudaMemcpyAsync(h->dmu, P.pinMu, mb, cudaMemcpyHostToDevice, st));
                P.h2d_bytes += (long long)mb;
            }
            int launches = 0, stride = 0, coff[3] = {0, 0, 0};
            for (int a = 0; a < 3; ++a)
                if (outs[a]) { coff[a] = stride; stride += P.cC[a]; }
            int err = compute ? qck_run(h, compute, 0, nk, h->dZ, h->dmu, h->dF, h->dJ, h->dH, st, 1) : QCK_OK;
            for (int a = 0; a < 3 && err == QCK_OK; ++a) {
                if (!outs[a] || P.cC[a] == 0) continue;
                const int e = qck_launch_pack(dfull[a], P.dC + coff[a], P.d_src[a], P.cC[a], stride, nnz[a], nk, st, &launches);
                if (e) err = qck_fail(h, QCK_ECUDA, "pack kernel launch: %s", cudaGetErrorString((cudaError_t)e));
            }
            if (err == QCK_OK && stride > 0) {
                cudaError_t ce = cudaMemcpyAsync(P.pinC, P.dC, sizeof(double) * (size_t)nk * stride, cudaMemcpyDeviceToHost, st);
                if (ce != cudaSuccess) err = qck_fail(h, QCK_ECUDA, "D2H copy: %s", cudaGetErrorString(ce));
                P.d2h_bytes += (long long)sizeof(double) * nk * stride;
            }
            const cudaError_t cs = cudaStreamSynchronize(st);
            if (err == QCK_OK && cs != cudaSuccess) err = qck_fail(h, QCK_ECUDA, "stream: %s", cudaGetErrorString(cs));
            h->launches += launches;
            if (err != QCK_OK) {
                P.valid_mask = 0; P.z_on_device = false; P.mu_on_device = false;
                return err;
            }
            for (int a = 0; a < 3; ++a)
                if (outs[a] && P.cC[a] > 0) expand_knots(h->own[a], P.pinC + coff[a], stride, outs[a], nnz[a], 0, nk);
            P.z_staged = true;
            P.z_on_device = true;
            if ((need & QCK_EVAL_H) && !mu_same) P.mu_on_device = true;
            P.valid_mask |= compute;
            return h->uses_status ? qck_check_status(h) : QCK_OK;
        }
    }

    // ---- 1. what is already on the device?  ("the same Z is presented to F, dF, mu d2F in succession") -----------
    static const bool no_cache = getenv("QCK_NO_CACHE") != nullptr;
    const bool z_same = !no_cache && P.z_staged && P.z_on_device && same_as_staged(pool, P.pinZ, Z, h->T * h->zdim);
    if (!z_same) { P.z_on_device = false; P.valid_mask = 0; }
    bool mu_same = true;
    if (need & QCK_EVAL_H) {
        mu_same = !no_cache && P.mu_on_device && same_as_staged(pool, P.pinMu, mu, nk * h->dyn);
        if (!mu_same) { P.mu_on_device = false; P.valid_mask &= ~QCK_EVAL_H; }
    }
    unsigned compute = need & ~P.valid_mask;
    // a residual call on a fresh Z is followed by the Jacobian call on the same Z within the Ipopt iteration: the fused pass
    // costs little more than the residual alone and saves the second launch (the values stay on the device until asked for)
    static const bool speculate = getenv("QCK_NO_SPECULATE") == nullptr;
    if (speculate && (compute & QCK_EVAL_F) && !(P.valid_mask & QCK_EVAL_J) && !(need & QCK_EVAL_H)) compute |= QCK_EVAL_J;
    if (!compute) ++P.cache_hits;
    const bool up_z = !z_same, up_mu = (need & QCK_EVAL_H) && !mu_same;
    P.z_staged = false;  // (set again when every chunk has been staged)

    // ---- 2./3. chunk pipeline ----------------------------------------------------------------------------------------------------
    // DIRECT RUNS: when a requested array's caller buffer is page-locked (qck_host_register, cudaHostRegister, cudaHostAlloc), its
    // long runs without repetition (all of F and of the Hessian values, the middle 834 doubles of a CZ knot's Jacobian) need no host
    // work at all: the copy engine writes them straight to their final place, chunk by chunk (1-D copies, or 2-D copies of >= 4 KB
    // rows), and they leave the compact row -- no pack, no staging slot, no expansion; the host memory system sees those bytes once
    // instead of three times.  (Measured and NOT done for the kron blocks: strided 2-D copies of 2.6 KB rows in pieces reach 24-37
    // GB/s of a 57 GB/s link, profiles/r02_e2e_pipeline.txt.)
    static const bool no_direct = getenv("QCK_NO_DIRECT") != nullptr;
    bool dirA[3] = {false, false, false};  // this call uses the (staged | direct) split of the array
    for (int a = 0; a < 3; ++a)
        dirA[a] = outs[a] && !no_direct && !P.own_direct[a].empty() && is_page_locked(outs[a]) && is_page_locked(outs[a] + nk * nnz[a] - 1);
    const std::vector<QckOwnSeg>* segsA[3];
    int cA[3];
    const int* srcA[3];
    for (int a = 0; a < 3; ++a) {
        segsA[a] = dirA[a] ? &P.own_staged[a] : &h->own[a];
        cA[a] = dirA[a] ? P.cS[a] : P.cC[a];
        srcA[a] = dirA[a] ? P.d_src_staged[a] : P.d_src[a];
    }
    int stride = 0, coff[3] = {0, 0, 0};  // compact row of this call: the staged runs of the requested arrays side by side
    for (int a = 0; a < 3; ++a)
        if (outs[a]) { coff[a] = stride; stride += cA[a]; }
    // transfer pieces: as many knots as fit a ring slot with THIS call's compact row (a single-array callback moves half the bytes
    // per knot of the fused call: same piece size in bytes, half the per-piece overheads)
    const long long ck = P.chunk_knots;
    const long long pk = stride > 0 ? std::min(ck, std::max<long long>(P.piece_knots, P.slot_doubles / stride)) : ck;
    const int nchunks = (int)((nk + ck - 1) / ck);
    struct Piece { int chunk; long long k0, kn; };
    std::vector<Piece> pieces;
    for (int c = 0; c < nchunks; ++c) {
        const long long k0 = c * ck, kn = std::min(ck, nk - k0);
        for (long long o = 0; o < kn; o += pk) pieces.push_back({c, k0 + o, std::min(pk, kn - o)});
    }
    const int npieces = (int)pieces.size();
    // the job: pieces [0, nchunks) stage chunk c's knots of Z / mu into page-locked memory (ready at once), pieces
    // [nchunks, nchunks + npieces) expand transfer piece q into the caller's arrays (ready when its D2H copy has landed)
    const long long zdim = h->zdim, dyn = h->dyn;
    constexpr long long kStageItem = 1ll << 15;  // doubles (256 KB)
    auto z_range = [&](int c, long long& o, long long& e) {  // knots k0 .. k0+kn; the halo knot k0 belongs to the previous chunk
        const long long k0 = c * ck, kn = std::min(ck, nk - k0);
        o = (c == 0 ? 0 : k0 + 1) * zdim;
        e = (k0 + kn + 1) * zdim;
    };
    long long per[3];
    for (int a = 0; a < 3; ++a) per[a] = std::max<long long>(1, (128ll << 10) / (8 * std::max<long long>(nnz[a], 1)));  // ~128 KB of output per item
    std::vector<int> nitems(nchunks + npieces, 0);
    for (int c = 0; c < nchunks; ++c) {
        long long o, e;
        z_range(c, o, e);
        const long long kn = std::min(ck, nk - c * ck);
        nitems[c] = (up_z ? (int)((e - o + kStageItem - 1) / kStageItem) : 0) + (up_mu ? (int)((kn * dyn + kStageItem - 1) / kStageItem) : 0);
    }
    for (int q = 0; q < npieces; ++q)
        for (int a = 0; a < 3; ++a)
            if (outs[a] && cA[a] > 0) nitems[nchunks + q] += (int)((pieces[q].kn + per[a] - 1) / per[a]);
    QckJob job;
    job.init(nitems);
    double* const pinZ = P.pinZ; double* const pinMu = P.pinMu; double* const pinC = P.pinC;
    const long long slot_doubles = P.slot_doubles;
    const int n_slots = P.n_slots;
    job.fn = [&, pinZ, pinMu, pinC, slot_doubles, n_slots](int piece, int item) {
        if (piece < nchunks) {  // staging
            long long o, e;
            z_range(piece, o, e);
            const int nz = up_z ? (int)((e - o + kStageItem - 1) / kStageItem) : 0;
            if (item < nz) {
                const long long b0 = o + item * kStageItem;
                memcpy(pinZ + b0, Z + b0, sizeof(double) * std::min(kStageItem, e - b0));
            } else {
                const long long k0 = piece * ck, kn = std::min(ck, nk - k0);
                const long long b0 = k0 * dyn + (item - nz) * kStageItem;
                memcpy(pinMu + b0, mu + b0, sizeof(double) * std::min(kStageItem, (k0 + kn) * dyn - b0));
            }
            return;
        }
        const int q = piece - nchunks;
        static const bool skip_expand = getenv("QCK_DIAG_SKIP_EXPAND") != nullptr;  // diagnosis only: transport without the host half
        if (skip_expand) return;
        const Piece& pc = pieces[q];
        const double* slot_h = pinC + (long long)(q % n_slots) * slot_doubles;
        int i = item;
        for (int a = 0; a < 3; ++a) {
            if (!outs[a] || cA[a] == 0) continue;
            const int na = (int)((pc.kn + per[a] - 1) / per[a]);
            if (i < na) {
                expand_knots(*segsA[a], slot_h + coff[a], stride, outs[a] + pc.k0 * nnz[a], nnz[a], i * per[a], std::min(pc.kn, (i + 1) * per[a]));
                return;
            }
            i -= na;
        }
    };
    for (int c = 0; c < nchunks; ++c) job.ready[c].store(1, std::memory_order_release);
    pool.begin(&job);
    auto wait_piece = [&](int p) {  // the caller helps while it waits
        while (!job.piece_done(p))
            if (!QckPool::try_work(&job)) _mm_pause();
    };

    int launches = 0, chunk_issued = -1;
    auto issue_chunk = [&](int c) -> int {  // H2D + kernels + pack of compute chunk c
        const long long k0 = c * ck, kn = std::min(ck, nk - k0);
        cudaStream_t st = P.st[c & 1];
        wait_piece(c);  // this chunk's knots are in page-locked memory
        if (up_z) {
            QCK_CUDA_TRY(h, cudaMemcpyAsync(h->dZ + k0 * zdim, P.pinZ + k0 * zdim, sizeof(double) * (kn + 1) * zdim, cudaMemcpyHostToDevice, st));
            P.h2d_bytes += sizeof(double) * (kn + 1) * zdim;
        }
        if (up_mu && (compute & QCK_EVAL_H)) {
            QCK_CUDA_TRY(h, cudaMemcpyAsync(h->dmu + k0 * dyn, P.pinMu + k0 * dyn, sizeof(double) * kn * dyn, cudaMemcpyHostToDevice, st));
            P.h2d_bytes += sizeof(double) * kn * dyn;
        }
        if (compute) {
            int r = qck_run(h, compute, k0, kn, h->dZ, h->dmu, h->dF, h->dJ, h->dH, st, 1 + (c & 1));
            if (r) return r;
        }
        double* region = P.dC + (long long)(c & 1) * P.chunk_doubles;
        for (int a = 0; a < 3; ++a) {
            if (!outs[a] || cA[a] == 0) continue;
            int e = qck_launch_pack(dfull[a] + k0 * nnz[a], region + coff[a], srcA[a], cA[a], stride, nnz[a], kn, st, &launches);
            if (e) return qck_fail(h, QCK_ECUDA, "pack kernel launch: %s", cudaGetErrorString((cudaError_t)e));
        }
        for (int a = 0; a < 3; ++a) {  // direct runs: this chunk's rows straight into the caller's array
            if (!dirA[a]) continue;
            for (auto& g : P.own_direct[a]) {
                double* dst = outs[a] + k0 * nnz[a] + g.full;
                const double* src = dfull[a] + k0 * nnz[a] + g.full;
                if (g.len == nnz[a]) QCK_CUDA_TRY(h, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)g.len * kn, cudaMemcpyDeviceToHost, st));
                else QCK_CUDA_TRY(h, cudaMemcpy2DAsync(dst, sizeof(double) * nnz[a], src, sizeof(double) * nnz[a], sizeof(double) * g.len, (size_t)kn, cudaMemcpyDeviceToHost, st));
                P.d2h_bytes += sizeof(double) * (long long)g.len * kn;
            }
        }
        if (outs[2] && P.dShared) {  // ensemble child: local partial sums of the shared Hessian entries travel separately
            const int ns = (int)h->sh_pos.size();
            int e = qck_launch_pack(h->dH + k0 * h->nnzH, P.dShared + k0 * ns, h->red.pos, ns, ns, h->nnzH, kn, st, &launches);
            if (e) return qck_fail(h, QCK_ECUDA, "pack kernel launch: %s", cudaGetErrorString((cudaError_t)e));
            QCK_CUDA_TRY(h, cudaMemcpyAsync(P.pinShared + k0 * ns, P.dShared + k0 * ns, sizeof(double) * kn * ns, cudaMemcpyDeviceToHost, st));
            P.d2h_bytes += sizeof(double) * kn * ns;
        }
        return QCK_OK;
    };
    auto issue_piece = [&](int q) -> int {
        const Piece& pc = pieces[q];
        if (pc.chunk != chunk_issued) {
            int r = issue_chunk(pc.chunk);
            if (r) return r;
            chunk_issued = pc.chunk;
        }
        cudaStream_t st = P.st[pc.chunk & 1];
        const double* region = P.dC + (long long)(pc.chunk & 1) * P.chunk_doubles;
        double* slot_h = P.pinC + (long long)(q % P.n_slots) * P.slot_doubles;
        const long long rel = pc.k0 - (long long)pc.chunk * ck;
        if (stride > 0) QCK_CUDA_TRY(h, cudaMemcpyAsync(slot_h, region + rel * stride, sizeof(double) * pc.kn * stride, cudaMemcpyDeviceToHost, st));
        P.d2h_bytes += sizeof(double) * pc.kn * stride;
        QCK_CUDA_TRY(h, cudaEventRecord(P.ev[q % P.n_slots], st));
        return QCK_OK;
    };

    int issued = 0, err = QCK_OK;
    double t_first = 0.0;
    for (; issued < npieces && issued < P.n_slots; ++issued)
        if ((err = issue_piece(issued)) != QCK_OK) break;
    int released = 0;
    for (int q = 0; q < npieces && err == QCK_OK; ++q) {
        cudaError_t ce = cudaEventSynchronize(P.ev[q % P.n_slots]);
        if (ce != cudaSuccess) { err = qck_fail(h, QCK_ECUDA, "piece %d: %s", q, cudaGetErrorString(ce)); break; }
        if (timing && q == 0) t_first = now_ms();
        job.ready[nchunks + q].store(1, std::memory_order_release);  // the host threads expand it
        released = q + 1;
        if (issued < npieces) {
            // the next piece reuses the ring slot of piece issued - n_slots: its expansion has to be through first
            wait_piece(nchunks + issued - P.n_slots);
            if ((err = issue_piece(issued)) != QCK_OK) break;
            ++issued;
        }
    }
    const double t_loop = timing ? now_ms() : 0.0;
    if (err != QCK_OK) {
        // let the job drain: copies that were issued may still land; pieces that were never released carry no work
        cudaStreamSynchronize(P.st[0]);
        cudaStreamSynchronize(P.st[1]);
        for (int q = released; q < npieces; ++q) { job.nitems[nchunks + q] = 0; job.ready[nchunks + q].store(1, std::memory_order_release); }
    }
    for (int p = 0; p < nchunks + npieces; ++p) wait_piece(p);
    pool.end(&job);
    h->launches += launches;
    if (err != QCK_OK) {
        P.valid_mask = 0; P.z_on_device = false; P.mu_on_device = false; P.z_staged = false;
        return err;
    }
    _mm_sfence();
    P.z_staged = true;
    P.z_on_device = true;
    if (up_mu && (compute & QCK_EVAL_H)) P.mu_on_device = true;
    P.valid_mask |= compute;
    if (timing) {
        const double t_end = now_ms();
        fprintf(stderr, "[qcknot pipe] dev %d need %u compute %u: first piece after %.2f ms, last event after %.2f ms, done after %.2f ms (%d chunks, %d pieces of %d knots, ring %d)\n",
                h->device, need, compute, t_first - t_begin, t_loop - t_begin, t_end - t_begin, nchunks, npieces, (int)pk, P.n_slots);
    }
    return h->uses_status ? qck_check_status(h) : QCK_OK;
}

// The objective / constraint terms (qck_objective.cu) read the same device-resident Z as the dynamics: if the caller's Z is the
// staged one nothing moves, otherwise it is staged and uploaded whole (on the handle's own stream) and the cached results of
// the previous Z are dropped.
int qck_pipe_ensure_z(qck_handle* h, const double* Z) {
    QCK_CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = pipe_init(h);
    if (rc) return rc;
    QckPipe& P = h->pipe;
    QckPool& pool = QckPool::get(h->host_threads);
    static const bool no_cache = getenv("QCK_NO_CACHE") != nullptr;
    const long long n = h->T * (long long)h->zdim;
    if (!no_cache && P.z_staged && P.z_on_device && same_as_staged(pool, P.pinZ, Z, n)) { ++P.cache_hits; P.h2d_bytes = 0; return QCK_OK; }
    P.valid_mask = 0;
    P.z_staged = P.z_on_device = false;
    const long long piece = 1ll << 15;
    pool.parallel_for((int)((n + piece - 1) / piece), [&](int i) {
        const long long o = (long long)i * piece;
        memcpy(P.pinZ + o, Z + o, sizeof(double) * std::min(piece, n - o));
    });
    QCK_CUDA_TRY(h, cudaMemcpyAsync(h->dZ, P.pinZ, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    QCK_CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // (the pipeline's own streams may read dZ next)
    P.h2d_bytes = sizeof(double) * n;
    P.z_staged = P.z_on_device = true;
    return QCK_OK;
}

extern "C" {

// what crosses PCIe for value array `arr` (0 F, 1 J, 2 H): runs (full offset, compact offset, length, repeats) per knot block
int qck_compact_map(const qck_handle* h, int32_t arr, int64_t* count, int32_t* segs) {
    if (!h || arr < 0 || arr > 2 || !count) return QCK_EINVAL;
    if (h->pub.order != QCK_ORDER_CSC) return QCK_EINVAL;  // the compact transport exists for the CSC order only
    const qck_handle* src = h;
    *count = (int64_t)src->own[arr].size();
    if (segs)
        for (size_t i = 0; i < src->own[arr].size(); ++i) {
            const QckOwnSeg& g = src->own[arr][i];
            segs[4 * i] = g.full; segs[4 * i + 1] = g.comp; segs[4 * i + 2] = g.len; segs[4 * i + 3] = g.nrep;
        }
    return QCK_OK;
}

// the host half of the host-buffer path on its own (works on structure-only handles: no device involved): expands `nk` knot
// blocks of the compact layout into `out` with the library's host threads, exactly as qck_eval_* do after the D2H copy
int qck_expand_host(const qck_handle* h, int32_t arr, const double* compact, double* out, int64_t nk) {
    if (!h || arr < 0 || arr > 2 || !compact || !out || nk < 0 || h->pub.order != QCK_ORDER_CSC) return QCK_EINVAL;
    long long C = 0;
    for (auto& g : h->own[arr]) C += g.len;
    const long long nnz = arr == 0 ? h->dyn : (arr == 1 ? h->nnzJ : h->nnzH);
    const long long per = std::max<long long>(1, (128ll << 10) / (8 * std::max<long long>(nnz, 1)));
    // same mechanics as a host-buffer call: a job whose pieces (runs of knots) are released one after the other
    const long long piece = std::max<long long>(per, env_ll("QCK_PIECE_KNOTS", 256));
    const int npieces = (int)((nk + piece - 1) / piece);
    std::vector<int> nitems(npieces);
    for (int q = 0; q < npieces; ++q) nitems[q] = (int)((std::min(piece, nk - q * piece) + per - 1) / per);
    QckJob job;
    job.init(nitems);
    job.fn = [&](int q, int i) {
        const long long k0 = q * piece, kn = std::min(piece, nk - k0);
        expand_knots(h->own[arr], compact + k0 * C, C, out + k0 * nnz, nnz, i * per, std::min(kn, (i + 1) * per));
    };
    QckPool& pool = QckPool::get(h->host_threads);
    pool.begin(&job);
    for (int q = 0; q < npieces; ++q) job.ready[q].store(1, std::memory_order_release);
    for (int q = 0; q < npieces; ++q)
        while (!job.piece_done(q))
            if (!QckPool::try_work(&job)) _mm_pause();
    pool.end(&job);
    _mm_sfence();
    return QCK_OK;
}

}  // extern "C"

void qck_pipe_destroy(qck_handle* h) {
    QckPipe& P = h->pipe;
    for (int i = 0; i < 2; ++i)
        if (P.st[i]) cudaStreamDestroy(P.st[i]);
    for (auto& e : P.ev) cudaEventDestroy(e);
    if (P.pinZ) cudaFreeHost(P.pinZ);
    if (P.pinMu) cudaFreeHost(P.pinMu);
    if (P.pinC) cudaFreeHost(P.pinC);
    if (P.pinShared) cudaFreeHost(P.pinShared);
    if (P.dC) cudaFree(P.dC);
    if (P.dShared) cudaFree(P.dShared);
    for (int a = 0; a < 3; ++a) {
        if (P.d_src_staged[a] && P.d_src_staged[a] != P.d_src[a]) cudaFree(const_cast<int*>(P.d_src_staged[a]));
        if (P.d_src[a]) cudaFree(const_cast<int*>(P.d_src[a]));
    }
    P = QckPipe{};
}
