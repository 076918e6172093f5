// One handle, several GPUs (SURVEY.md section 8b "n_gpus, shard_mode", 8e): the serial caller (Ipopt's callbacks) drives
// every GPU of the box through one qck_handle.
//
//   shard_mode KNOT      GPU g owns the contiguous block range [t0_g, t0_g+1) and reads knots t0_g .. t0_g+1 (one-knot halo).
//                        Values are knot-major, so its F / J / H outputs are contiguous segments of the caller's arrays; no
//                        collective sits on the data path.
//   shard_mode ENSEMBLE  GPU g owns a contiguous range of the quantum integrators (the sampled systems of
//                        unitary_sampling_problem.jl:134-155) for every knot.  Rows / Jacobian entries of different systems are
//                        disjoint; Hessian entries on the shared controls are summed over the GPUs (host path: fixed order on
//                        the host, bitwise reproducible; device-resident path: every GPU sums the partial columns of all GPUs
//                        over NVLink peer memory in one launch -- same order, same bits as a one-GPU run; ncclAllReduce of
//                        the packed entries where peer access is not available or with QCK_ENSEMBLE_NCCL=1).
//
// Host-buffer calls run the single-GPU pipeline (qck_pipe.cpp) of every child concurrently, one host thread per GPU, each
// GPU copying over its own PCIe link into its own page-locked ring and the shared pool expanding into the caller's arrays.
// NCCL is loaded with dlopen (no link-time dependency; the process may already hold a libnccl) and used only for the
// device-resident gather of the segments and the all-reduce of the shared entries.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include <nccl.h>

#include "qck_handle.h"

namespace {

struct QckNccl {
    void* lib = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::vector<ncclComm_t> comms;
    int version = 0;
};

int nccl_init(qck_handle* h) {
    if (h->nccl) return QCK_OK;
    QckNccl* N = new QckNccl();
    const char* names[] = {getenv("QCK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        N->lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (N->lib) break;
    }
    if (!N->lib) { delete N; return qck_fail(h, QCK_ENCCL, "cannot load libnccl.so.2 (%s); set QCK_NCCL_LIB", dlerror()); }
#define QCK_SYM(field, name)                                                                                  \
    *(void**)(&N->field) = dlsym(N->lib, name);                                                               \
    if (!N->field) { delete N; return qck_fail(h, QCK_ENCCL, "libnccl lacks %s", name); }
    QCK_SYM(GetErrorString, "ncclGetErrorString")
    QCK_SYM(CommInitAll, "ncclCommInitAll")
    QCK_SYM(CommDestroy, "ncclCommDestroy")
    QCK_SYM(AllReduce, "ncclAllReduce")
    QCK_SYM(Broadcast, "ncclBroadcast")
    QCK_SYM(AllGather, "ncclAllGather")
    QCK_SYM(GroupStart, "ncclGroupStart")
    QCK_SYM(GroupEnd, "ncclGroupEnd")
    QCK_SYM(GetVersion, "ncclGetVersion")
#undef QCK_SYM
    N->GetVersion(&N->version);
    std::vector<int> devs;
    for (qck_handle* c : h->children) devs.push_back(c->device);
    N->comms.resize(devs.size());
    ncclResult_t r = N->CommInitAll(N->comms.data(), (int)devs.size(), devs.data());
    if (r != ncclSuccess) {
        const char* msg = N->GetErrorString(r);
        delete N;
        return qck_fail(h, QCK_ENCCL, "ncclCommInitAll over %d devices: %s", (int)devs.size(), msg);
    }
    h->nccl = N;
    return QCK_OK;
}

// ENSEMBLE, device-resident: peer access between all GPUs of the handle + the global contributor tables on every GPU
int peer_init(qck_handle* h) {
    if (h->peer_ready) return QCK_OK;
    h->peer_ready = true;
    h->peer_ok = false;
    const int n = (int)h->children.size();
    if (n > QCK_MAX_GPUS || getenv("QCK_ENSEMBLE_NCCL")) return QCK_OK;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, h->children[a]->device, h->children[b]->device) != cudaSuccess || !can) return QCK_OK;
        }
    for (int a = 0; a < n; ++a) {
        QCK_CUDA_TRY(h, cudaSetDevice(h->children[a]->device));
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(h->children[b]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); return QCK_OK; }
        }
    }
    // global contributor list of every shared position: the children's lists one after the other (ascending integrators)
    const int ns = (int)h->sh_pos.size();
    std::vector<int> ptr(ns + 1, 0), cols;
    for (int s = 0; s < ns; ++s) {
        for (int g = 0; g < n; ++g) {
            const qck_handle* c = h->children[g];
            for (int k = c->sh_ptr[s]; k < c->sh_ptr[s + 1]; ++k) cols.push_back((g << 24) | c->sh_cols[k]);
        }
        ptr[s + 1] = (int)cols.size();
    }
    h->peer.assign(n, QckPeerReduce{});
    h->ev_kernel.assign(n, nullptr);
    h->ev_reduce.assign(n, nullptr);
    for (int g = 0; g < n; ++g) {
        qck_handle* c = h->children[g];
        QCK_CUDA_TRY(h, cudaSetDevice(c->device));
        QckPeerReduce& R = h->peer[g];
        R.n_shared = ns;
        R.pos = c->red.pos;
        int *dptr = nullptr, *dcols = nullptr;
        QCK_CUDA_TRY(h, cudaMalloc((void**)&dptr, sizeof(int) * ptr.size()));
        QCK_CUDA_TRY(h, cudaMalloc((void**)&dcols, sizeof(int) * std::max<size_t>(cols.size(), 1)));
        c->allocs.push_back(dptr);
        c->allocs.push_back(dcols);
        QCK_CUDA_TRY(h, cudaMemcpy(dptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice));
        if (!cols.empty()) QCK_CUDA_TRY(h, cudaMemcpy(dcols, cols.data(), sizeof(int) * cols.size(), cudaMemcpyHostToDevice));
        R.ptr = dptr;
        R.cols = dcols;
        for (int b = 0; b < n; ++b) { R.partial[b] = h->children[b]->dpartial; R.npart[b] = h->children[b]->npart; }
        QCK_CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_kernel[g], cudaEventDisableTiming));
        QCK_CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_reduce[g], cudaEventDisableTiming));
    }
    h->peer_ok = true;
    return QCK_OK;
}

#define NCCL_TRY(h, N, call)                                                                         \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) return qck_fail(h, QCK_ENCCL, "%s: %s", #call, (N)->GetErrorString(r_)); \
    } while (0)

}  // namespace

// balanced contiguous block ranges; boundaries on even blocks where possible (aligned segment bases)
void qck_split_knots(long long n_blocks, int n, std::vector<long long>& t0) {
    t0.assign(n + 1, 0);
    for (int g = 0; g <= n; ++g) {
        long long b = n_blocks * g / n;
        if (g > 0 && g < n && (b & 1) && b + 1 < n_blocks * (g + 1) / n) ++b;
        t0[g] = b;
    }
    t0[n] = n_blocks;
}

// contiguous integrator ranges with an equal share of the quantum integrators each; whatever sits between two quantum
// integrators (derivative integrators) goes with the range it falls into, the tail to the last range
void qck_split_integrators(const std::vector<QckInteg>& integ, int n, std::vector<int>& q0) {
    std::vector<int> qidx;
    for (int q = 0; q < (int)integ.size(); ++q)
        if (integ[q].quantum()) qidx.push_back(q);
    q0.assign(n + 1, 0);
    const long long nq = (long long)qidx.size();
    for (int g = 1; g < n; ++g) q0[g] = nq ? qidx[(size_t)(nq * g / n)] : 0;
    q0[n] = (int)integ.size();
}

int qck_multi_create(const qck_problem_desc* d, qck_handle* parent) {
    const int n = d->n_gpus;
    parent->shard_mode = d->shard_mode;
    const bool structure_only = d->device < 0;
    std::vector<int> devs(n);
    for (int g = 0; g < n; ++g) devs[g] = structure_only ? -1 : (d->devices ? d->devices[g] : d->device + g);
    const long long nb = d->T - 1;
    std::vector<int> q0;
    if (d->shard_mode == QCK_SHARD_KNOT) {
        if (nb < n) return qck_fail(parent, QCK_EINVAL, "%lld knot blocks cannot be sharded over %d GPUs", nb, n);
        qck_split_knots(nb, n, parent->child_t0);
    } else {
        int nq = 0;
        for (auto& I : parent->integ) nq += I.quantum();
        if (nq < n) return qck_fail(parent, QCK_EINVAL, "%d quantum integrators cannot be sharded over %d GPUs", nq, n);
        qck_split_integrators(parent->integ, n, q0);
        parent->child_t0.assign(n + 1, 0);
        parent->child_t0[n] = nb;
    }
    for (int g = 0; g < n; ++g) {
        qck_problem_desc dc = *d;
        dc.n_gpus = 1; dc.devices = nullptr; dc.device = devs[g];
        if (d->shard_mode == QCK_SHARD_KNOT) {
            dc.T = parent->child_t0[g + 1] - parent->child_t0[g] + 1;
            dc.integ_begin = 0; dc.integ_end = -1;
        } else {
            dc.integ_begin = q0[g]; dc.integ_end = q0[g + 1];
        }
        qck_handle* c = nullptr;
        int rc = qck_create_single(&dc, &c, d->shard_mode == QCK_SHARD_ENSEMBLE);
        if (rc) {
            parent->err = std::string("GPU ") + std::to_string(g) + ": " + qck_last_error(nullptr);
            return rc;
        }
        parent->children.push_back(c);
    }
    parent->gF.assign(n, nullptr); parent->gJ.assign(n, nullptr); parent->gH.assign(n, nullptr);
    return QCK_OK;
}

void qck_multi_destroy(qck_handle* h) {
    if (h->nccl) {
        QckNccl* N = static_cast<QckNccl*>(h->nccl);
        for (auto c : N->comms) N->CommDestroy(c);
        delete N;  // (the library stays loaded: other handles may use it)
        h->nccl = nullptr;
    }
    for (size_t g = 0; g < h->children.size(); ++g) {
        if (h->children[g]->device >= 0) {
            cudaSetDevice(h->children[g]->device);
            if (g < h->ev_kernel.size() && h->ev_kernel[g]) cudaEventDestroy(h->ev_kernel[g]);
            if (g < h->ev_reduce.size() && h->ev_reduce[g]) cudaEventDestroy(h->ev_reduce[g]);
            for (double* p : {g < h->gF.size() ? h->gF[g] : nullptr, g < h->gJ.size() ? h->gJ[g] : nullptr, g < h->gH.size() ? h->gH[g] : nullptr})
                if (p) cudaFree(p);
        }
        qck_destroy(h->children[g]);
    }
    h->children.clear();
}

int qck_multi_eval(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) {
    const int n = (int)h->children.size();
    if (h->children[0]->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle (device=-1): libqcknot has no CPU evaluation path");
    if (H && h->eval_hessian && !mu) return qck_fail(h, QCK_EINVAL, "the Hessian needs the multipliers mu");
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    const bool knot = h->shard_mode == QCK_SHARD_KNOT;
    for (int g = 0; g < n; ++g) {
        th.emplace_back([&, g] {
            qck_handle* c = h->children[g];
            const long long t0 = knot ? h->child_t0[g] : 0;
            rcs[g] = qck_pipe_eval(c, Z + t0 * h->zdim, mu ? mu + t0 * h->dyn : nullptr, F ? F + t0 * h->dyn : nullptr,
                                   J ? J + t0 * h->nnzJ : nullptr, H ? H + t0 * h->nnzH : nullptr);
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < n; ++g)
        if (rcs[g]) { h->err = "GPU " + std::to_string(g) + ": " + h->children[g]->err; return rcs[g]; }
    if (!knot && H && h->eval_hessian && !h->sh_pos.empty()) {
        // Hessian entries on the shared controls: per-GPU partial sums, added in ascending GPU order (bitwise reproducible)
        const long long nk = h->T - 1;
        const int ns = (int)h->sh_pos.size();
        QckPool::get(h->host_threads).parallel_for((int)((nk + 4095) / 4096), [&](int i) {
            for (long long t = i * 4096ll, te = std::min(nk, t + 4096); t < te; ++t)
                for (int s = 0; s < ns; ++s) {
                    double v = 0.0;
                    for (int g = 0; g < n; ++g) {
                        const double* ps = h->children[g]->pipe.pinShared;
                        if (ps) v += ps[t * ns + s];
                    }
                    H[t * h->nnzH + h->sh_pos[s]] = v;
                }
        });
    }
    return QCK_OK;
}

extern "C" {

int qck_shard_count(const qck_handle* h, int32_t* n) {
    if (!h || !n) return QCK_EINVAL;
    *n = h->children.empty() ? 1 : (int)h->children.size();
    return QCK_OK;
}

int qck_shard_info(const qck_handle* h, int32_t g, int32_t* device, int64_t* block_begin, int64_t* block_end, int32_t* integ_begin, int32_t* integ_end) {
    if (!h) return QCK_EINVAL;
    const int n = h->children.empty() ? 1 : (int)h->children.size();
    if (g < 0 || g >= n) return QCK_EINVAL;
    const qck_handle* c = h->children.empty() ? h : h->children[g];
    const bool knot = h->children.empty() || h->shard_mode == QCK_SHARD_KNOT;
    if (device) *device = c->device;
    if (block_begin) *block_begin = h->children.empty() ? 0 : (knot ? h->child_t0[g] : 0);
    if (block_end) *block_end = h->children.empty() ? h->T - 1 : (knot ? h->child_t0[g + 1] : h->T - 1);
    if (integ_begin) *integ_begin = c->ib;
    if (integ_end) *integ_end = c->ie;
    return QCK_OK;
}

int qck_shard_device_buffers(qck_handle* h, int32_t g, double** dZ, double** dmu, double** dF, double** dJ, double** dH) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (h->children.empty()) return g == 0 ? qck_device_buffers(h, dZ, dmu, dF, dJ, dH) : QCK_EINVAL;
    if (g < 0 || g >= (int)h->children.size()) return QCK_EINVAL;
    return qck_device_buffers(h->children[g], dZ, dmu, dF, dJ, dH);
}

// H2D of the inputs into every shard's own device buffers (each GPU gets its knots + halo / all knots)
int qck_upload(qck_handle* h, const double* Z, const double* mu) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h || !Z) return QCK_EINVAL;
    std::vector<qck_handle*> one{h};
    const std::vector<qck_handle*>& cs = h->children.empty() ? one : h->children;
    const bool knot = h->children.empty() || h->shard_mode == QCK_SHARD_KNOT;
    for (size_t g = 0; g < cs.size(); ++g) {
        qck_handle* c = cs[g];
        if (c->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle");
        const long long t0 = (h->children.empty() || !knot) ? 0 : h->child_t0[g];
        QCK_CUDA_TRY(h, cudaSetDevice(c->device));
        QCK_CUDA_TRY(h, cudaMemcpyAsync(c->dZ, Z + t0 * h->zdim, sizeof(double) * c->T * c->zdim, cudaMemcpyHostToDevice, c->stream));
        if (mu) QCK_CUDA_TRY(h, cudaMemcpyAsync(c->dmu, mu + t0 * h->dyn, sizeof(double) * (c->T - 1) * c->dyn, cudaMemcpyHostToDevice, c->stream));
        c->pipe.valid_mask = 0; c->pipe.z_on_device = false; c->pipe.mu_on_device = false; c->pipe.z_staged = false;
    }
    for (qck_handle* c : cs) {
        QCK_CUDA_TRY(h, cudaSetDevice(c->device));
        QCK_CUDA_TRY(h, cudaStreamSynchronize(c->stream));
    }
    return QCK_OK;
}

// one pass on every GPU over its shard, inputs and outputs in the shards' own device buffers; asynchronous.
// ENSEMBLE: the Hessian entries on the shared controls are all-reduced (NCCL) so that every GPU holds the sums.
int qck_eval_resident(qck_handle* h, uint32_t mask) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (h->children.empty()) {
        if (h->device < 0) return qck_fail(h, QCK_ENODEVICE, "structure-only handle");
        QCK_CUDA_TRY(h, cudaSetDevice(h->device));
        int rc = qck_run(h, mask, 0, h->T - 1, h->dZ, h->dmu, (mask & QCK_EVAL_F) ? h->dF : nullptr, (mask & QCK_EVAL_J) ? h->dJ : nullptr,
                         (mask & QCK_EVAL_H) ? h->dH : nullptr, h->stream, 0);
        if (!rc && h->pub.order != QCK_ORDER_CSC) rc = qck_reorder(h, mask, h->pub.dJ, h->pub.dH, h->stream);  // qck_device_buffers returns these
        return rc;
    }
    const bool shared = h->shard_mode == QCK_SHARD_ENSEMBLE && (mask & QCK_EVAL_H) && h->eval_hessian && !h->sh_pos.empty();
    if (shared) {
        int rc = peer_init(h);
        if (rc) return rc;
    }
    const bool peer = shared && h->peer_ok;
    const int n = (int)h->children.size();
    for (int g = 0; g < n; ++g) {
        qck_handle* c = h->children[g];
        QCK_CUDA_TRY(h, cudaSetDevice(c->device));
        if (peer)  // the partial columns of the previous pass have been read by every GPU
            for (int b = 0; b < n; ++b)
                if (b != g) QCK_CUDA_TRY(h, cudaStreamWaitEvent(c->stream, h->ev_reduce[b], 0));
        c->skip_local_reduce = peer;
        int rc = qck_run(c, mask, 0, c->T - 1, c->dZ, c->dmu, (mask & QCK_EVAL_F) ? c->dF : nullptr, (mask & QCK_EVAL_J) ? c->dJ : nullptr,
                         (mask & QCK_EVAL_H) ? c->dH : nullptr, c->stream, 0);
        c->skip_local_reduce = false;
        if (rc) { h->err = c->err; return rc; }
        c->pipe.valid_mask = 0;
        if (peer) QCK_CUDA_TRY(h, cudaEventRecord(h->ev_kernel[g], c->stream));
    }
    if (peer) {
        // all-reduce by all-read: every GPU sums the partial columns of all GPUs over NVLink (one launch per GPU, fixed order)
        for (int g = 0; g < n; ++g) {
            qck_handle* c = h->children[g];
            QCK_CUDA_TRY(h, cudaSetDevice(c->device));
            for (int b = 0; b < n; ++b)
                if (b != g) QCK_CUDA_TRY(h, cudaStreamWaitEvent(c->stream, h->ev_kernel[b], 0));
            int launches = 0;
            int e = qck_launch_peer_reduce(h->peer[g], c->dH, c->T - 1, c->nnzH, c->stream, &launches);
            if (e) return qck_fail(h, QCK_ECUDA, "peer reduce launch: %s", cudaGetErrorString((cudaError_t)e));
            c->launches += launches;
            QCK_CUDA_TRY(h, cudaEventRecord(h->ev_reduce[g], c->stream));
        }
    } else if (shared) {
        int rc = nccl_init(h);
        if (rc) return rc;
        QckNccl* N = static_cast<QckNccl*>(h->nccl);
        const long long nk = h->T - 1;
        const int ns = (int)h->sh_pos.size();
        for (qck_handle* c : h->children) {
            QCK_CUDA_TRY(h, cudaSetDevice(c->device));
            if (!c->pipe.dShared) QCK_CUDA_TRY(h, cudaMalloc((void**)&c->pipe.dShared, sizeof(double) * nk * ns));
            int launches = 0;
            int e = qck_launch_pack(c->dH, c->pipe.dShared, c->red.pos, ns, ns, c->nnzH, nk, c->stream, &launches);
            if (e) return qck_fail(h, QCK_ECUDA, "pack kernel launch: %s", cudaGetErrorString((cudaError_t)e));
            c->launches += launches;
        }
        NCCL_TRY(h, N, N->GroupStart());
        for (size_t g = 0; g < h->children.size(); ++g) {
            qck_handle* c = h->children[g];
            NCCL_TRY(h, N, N->AllReduce(c->pipe.dShared, c->pipe.dShared, (size_t)(nk * ns), ncclDouble, ncclSum, N->comms[g], c->stream));
        }
        NCCL_TRY(h, N, N->GroupEnd());
        for (qck_handle* c : h->children) {
            QCK_CUDA_TRY(h, cudaSetDevice(c->device));
            int launches = 0;
            int e = qck_launch_unpack(c->dH, c->pipe.dShared, c->red.pos, ns, ns, c->nnzH, nk, c->stream, &launches);
            if (e) return qck_fail(h, QCK_ECUDA, "unpack kernel launch: %s", cudaGetErrorString((cudaError_t)e));
            c->launches += launches;
        }
    }
    return QCK_OK;
}

// KNOT sharding: all-gather of the device-resident segments so that every GPU holds the assembled value arrays
// (ncclBroadcast per segment inside one group: the segments may differ in length).  Asynchronous on the shards' streams.
int qck_gather_device(qck_handle* h, uint32_t mask) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (h->children.empty()) return QCK_OK;  // one GPU: the shard buffers are the assembled arrays
    if (h->shard_mode != QCK_SHARD_KNOT) return qck_fail(h, QCK_EINVAL, "qck_gather_device assembles knot-sharded arrays; ensemble shards interleave inside every knot block");
    int rc = nccl_init(h);
    if (rc) return rc;
    QckNccl* N = static_cast<QckNccl*>(h->nccl);
    const int n = (int)h->children.size();
    const long long nk = h->T - 1;
    const long long nnz[3] = {h->dyn, h->nnzJ, h->eval_hessian ? h->nnzH : 0};
    std::vector<double*>* gs[3] = {&h->gF, &h->gJ, &h->gH};
    for (int a = 0; a < 3; ++a) {
        if (!((mask >> a) & 1u) || nnz[a] == 0) continue;
        for (int g = 0; g < n; ++g)
            if (!(*gs[a])[g]) {
                QCK_CUDA_TRY(h, cudaSetDevice(h->children[g]->device));
                QCK_CUDA_TRY(h, cudaMalloc((void**)&(*gs[a])[g], sizeof(double) * nk * nnz[a]));
            }
    }
    NCCL_TRY(h, N, N->GroupStart());
    for (int a = 0; a < 3; ++a) {
        if (!((mask >> a) & 1u) || nnz[a] == 0) continue;
        for (int s = 0; s < n; ++s) {  // segment s lives on GPU s
            const long long off = h->child_t0[s] * nnz[a], cnt = (h->child_t0[s + 1] - h->child_t0[s]) * nnz[a];
            for (int g = 0; g < n; ++g) {
                qck_handle* c = h->children[g];
                const double* src = a == 0 ? h->children[s]->dF : (a == 1 ? h->children[s]->dJ : h->children[s]->dH);
                NCCL_TRY(h, N, N->Broadcast(g == s ? src : (const void*)((*gs[a])[g] + off), (*gs[a])[g] + off, (size_t)cnt, ncclDouble, s, N->comms[g], c->stream));
            }
        }
    }
    NCCL_TRY(h, N, N->GroupEnd());
    return QCK_OK;
}

int qck_gathered_buffers(qck_handle* h, int32_t g, double** dF, double** dJ, double** dH) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (h->children.empty()) return g == 0 ? qck_device_buffers(h, nullptr, nullptr, dF, dJ, dH) : QCK_EINVAL;
    if (g < 0 || g >= (int)h->children.size()) return QCK_EINVAL;
    if (dF) *dF = h->gF[g];
    if (dJ) *dJ = h->gJ[g];
    if (dH) *dH = h->gH[g];
    return QCK_OK;
}

int qck_nccl_version(qck_handle* h, int32_t* version, int32_t* nranks) {
    if (!h) return QCK_EINVAL;
    if (h->children.empty()) { if (version) *version = 0; if (nranks) *nranks = 1; return QCK_OK; }
    int rc = nccl_init(h);
    if (rc) return rc;
    QckNccl* N = static_cast<QckNccl*>(h->nccl);
    if (version) *version = N->version;
    if (nranks) *nranks = (int)N->comms.size();
    return QCK_OK;
}

}  // extern "C"
