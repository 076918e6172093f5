// Host side of libqcknot.so: the C-ABI of include/qcknot.h.
//
// qck_create() is the B200 counterpart of the reference's `QuantumDynamics(integrators, traj)` constructor
// (/root/reference/test/scripts/integrator_test_1qubit.jl:41): it builds the per-knot sparsity structures once,
// turns every structural nonzero into a (position, scratch slot, sign) map entry for the kernels, uploads the
// Hamiltonians, and allocates the device-resident value buffers.  The eval entry points are what the MOI
// callbacks eval_constraint / eval_constraint_jacobian / eval_hessian_lagrangian bind.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "qck_internal.h"

namespace {

thread_local std::string g_create_error;

struct Integ {
    int kind = 0, order = 0, N = 0, nd = 0, state_off = 0, state_len = 0, ctrl_off = 0;
    int row_off = 0, dim = 0, nc = 0;
    std::vector<std::complex<double>> Hdrift, Hdrives;
    bool quantum() const { return kind != QCK_DERIVATIVE; }
    bool unitary() const { return kind == QCK_UNITARY_PADE || kind == QCK_UNITARY_EXP; }
    bool pade() const { return kind == QCK_UNITARY_PADE || kind == QCK_KET_PADE; }
};

struct ClassHost {
    QckClassDev dev{};
    std::vector<int> members;  // integrator indices, ascending
    int member_begin = 0, member_end = 0;
    // device allocations owned by the class
    std::vector<void*> allocs;
};

bool g_structure_only = false;  // set while building a device-less (structure-only) handle

template <class T>
cudaError_t upload(const std::vector<T>& v, const T** out, std::vector<void*>& allocs) {
    *out = nullptr;
    if (v.empty() || g_structure_only) return cudaSuccess;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    allocs.push_back(d);
    e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = static_cast<const T*>(d);
    return e;
}

}  // namespace

struct qck_handle {
    std::string err;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    long long T = 0;
    int zdim = 0, dt_off = -1, eval_hessian = 1, ib = 0, ie = 0;
    double dt_fixed = 0.0;
    std::vector<Integ> integ;
    int dyn = 0;
    long long nnzJ = 0, nnzH = 0;
    std::vector<int32_t> Jr, Jc, Hr, Hc;  // per-knot structure, 0-based, CSC order
    std::vector<ClassHost> classes;
    std::vector<QckAux> aux;
    const QckAux* d_aux = nullptr;
    std::vector<int> sh_pos, sh_ptr, sh_cols;      // active contributors only (reduce kernel)
    std::vector<long long> shared_positions;       // globally shared positions
    QckReduce red{};
    int npart = 0;
    std::vector<void*> allocs;
    double *dZ = nullptr, *dmu = nullptr, *dF = nullptr, *dJ = nullptr, *dH = nullptr, *dpartial = nullptr;
    long long launches = 0;
};

namespace {

int fail(qck_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, call)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(h, QCK_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct HEntry {
    long long key;  // col * (2*zdim) + row  (CSC order)
    int contrib;    // integrator index
    int cls, member;  // quantum class/member or -1
    uint16_t src;     // scratch slot (+sign) for quantum contributors
    int aux_op, aux_i0;  // for derivative contributors
    int grp, period;     // output group of the integrator (segments never span groups) and its repeat period (0 = none)
};
struct JEntry {
    long long key;
    int cls, member;
    uint16_t src;
    int aux_op, aux_i0;
    double aux_c;
    int grp, period;
};
struct MapEntry {
    long long dst;
    uint16_t src;
    int grp, period;
    bool operator<(const MapEntry& o) const { return dst < o.dst; }
};

// Cut one member's (destination-sorted) entries into segments and append the slot table.  Members of one class
// must produce identical tables and segment shapes (only `dst` differs); returns false otherwise.
bool make_segments(const std::vector<MapEntry>& e, long long split, std::vector<QckSeg>& segs, std::vector<uint16_t>& tab) {
    size_t k = 0;
    while (k < e.size()) {
        size_t k2 = k + 1;
        while (k2 < e.size() && e[k2].dst == e[k2 - 1].dst + 1 && e[k2].grp == e[k].grp && ((e[k2].dst < split) == (e[k].dst < split))) ++k2;
        int len = (int)(k2 - k), period = len;
        int p = e[k].period;
        if (p > 0 && len % p == 0 && len / p >= 2) {
            bool ok = true;
            for (size_t u = k + p; u < k2 && ok; ++u) ok = e[u].src == e[u - p].src;
            if (ok) period = p;
        }
        QckSeg sg{(int)e[k].dst, len, (int)tab.size(), period};
        for (int u = 0; u < period; ++u) tab.push_back(e[k + u].src);
        segs.push_back(sg);
        k = k2;
    }
    return true;
}

// slot of iso-vec element i of state-type matrix `sidx`
inline int slot_state(const QckClassDev& c, int sidx, int i) {
    int cc = i / (2 * c.N), q = i - cc * 2 * c.N;
    int im = q >= c.N, r = q - im * c.N;
    return c.off_S + sidx * c.mss + 2 * (r + c.NP * cc) + im;
}
// slot and sign of iso(X)[q, r] for A-type matrix `aidx`:  iso(X) = [Re X, -Im X; Im X, Re X]
inline uint16_t slot_iso(const QckClassDev& c, int aidx, int q, int r, bool negate) {
    int N = c.N;
    int qi = q >= N, ri = r >= N;
    int qq = q - qi * N, rr = r - ri * N;
    int im = qi != ri;                 // off-diagonal quadrants hold Im
    bool neg = (!qi && ri) != negate;  // upper-right quadrant is -Im
    int s = c.off_A + aidx * c.msa + 2 * (qq + c.NP * rr) + im;
    return (uint16_t)(s | (neg ? 0x8000 : 0));
}

int build(qck_handle* h) {
    const int zdim = h->zdim;
    const bool free_time = h->dt_off >= 0;
    const int nI = (int)h->integ.size();
    // ---- classes ---------------------------------------------------------------------------------------------
    std::map<std::vector<int>, int> class_of;
    std::vector<int> cls_idx(nI, -1), mem_idx(nI, -1);
    for (int q = 0; q < nI; ++q) {
        Integ& I = h->integ[q];
        if (!I.quantum()) continue;
        std::vector<int> key = {I.kind, I.N, I.nd, I.order};
        auto it = class_of.find(key);
        int ci;
        if (it == class_of.end()) {
            ci = (int)h->classes.size();
            class_of[key] = ci;
            h->classes.emplace_back();
            QckClassDev& c = h->classes.back().dev;
            c.kind = I.kind; c.N = I.N; c.nd = I.nd; c.order = I.order;
            c.nc = I.nc;
            c.NP = ((I.N + QCK_TILE - 1) / QCK_TILE) * QCK_TILE;
            c.ncp = I.unitary() ? c.NP : 1;
            c.free_time = free_time; c.dt_off = h->dt_off; c.zdim = zdim; c.dyn = h->dyn; c.dt_fixed = h->dt_fixed;
            qck_scratch_layout(c);
            if (c.scratch_doubles > 32768 || (size_t)c.scratch_doubles * 8 > 227 * 1024)
                return fail(h, QCK_EINVAL, "levels=%d with %d drives needs %d scratch doubles per knot; this build supports at most %d",
                            I.N, I.nd, c.scratch_doubles, 227 * 1024 / 8);
        } else ci = it->second;
        cls_idx[q] = ci;
        mem_idx[q] = (int)h->classes[ci].members.size();
        h->classes[ci].members.push_back(q);
    }
    for (auto& C : h->classes) {
        C.dev.n_members = (int)C.members.size();
        int b = 0, e = 0;
        bool any = false;
        for (int k = 0; k < (int)C.members.size(); ++k) {
            bool act = C.members[k] >= h->ib && C.members[k] < h->ie;
            if (act && !any) { b = k; any = true; }
            if (act) e = k + 1;
        }
        C.member_begin = b; C.member_end = any ? e : b;
    }

    // ---- entries -----------------------------------------------------------------------------------------------
    std::vector<JEntry> JE;
    std::vector<HEntry> HE;
    auto jkey = [&](int row, int col) { return (long long)col * h->dyn + row; };
    auto hkey = [&](int r, int c) { if (r > c) std::swap(r, c); return (long long)c * (2 * zdim) + r; };
    for (int q = 0; q < nI; ++q) {
        const Integ& I = h->integ[q];
        const int R0 = I.row_off;
        if (I.quantum()) {
            const int ci = cls_idx[q], mi = mem_idx[q];
            const QckClassDev& c = h->classes[ci].dev;
            const int N = I.N, n2 = 2 * N;
            // Jacobian: state_t block (-F or -E), state_t+1 block (+B or identity), controls, timestep
            const int blk = n2 * n2;  // one kron(I_N, .) block repeats every 2N columns of 2N rows
            const int DRV = QS_FIXED;
            for (int cb = 0; cb < I.nc; ++cb)
                for (int r = 0; r < n2; ++r)
                    for (int qq = 0; qq < n2; ++qq) {
                        JE.push_back({jkey(R0 + cb * n2 + qq, I.state_off + cb * n2 + r), ci, mi, slot_iso(c, QA_F, qq, r, true), 0, 0, 0, 0, blk});
                        if (I.pade())
                            JE.push_back({jkey(R0 + cb * n2 + qq, zdim + I.state_off + cb * n2 + r), ci, mi, slot_iso(c, QA_B, qq, r, false), 0, 0, 0, 1, blk});
                    }
            if (!I.pade())
                for (int i = 0; i < I.dim; ++i)
                    JE.push_back({jkey(R0 + i, zdim + I.state_off + i), ci, mi, (uint16_t)(c.off_X + QX_ONE), 0, 0, 0, 1, 1});
            for (int j = 0; j < I.nd; ++j)
                for (int i = 0; i < I.dim; ++i)
                    JE.push_back({jkey(R0 + i, I.ctrl_off + j), ci, mi, (uint16_t)slot_state(c, DRV + QD_COUNT * j + QD_TA, i), 0, 0, 0, 2 + j, 0});
            if (free_time)
                for (int i = 0; i < I.dim; ++i)
                    JE.push_back({jkey(R0 + i, h->dt_off), ci, mi, (uint16_t)slot_state(c, QS_AS, i), 0, 0, 0, 2 + I.nd, 0});
            if (h->eval_hessian) {
                const int nd = I.nd;
                for (int j = 0; j < nd; ++j)
                    for (int i = 0; i < I.dim; ++i) {
                        HE.push_back({hkey(I.state_off + i, I.ctrl_off + j), q, ci, mi, (uint16_t)slot_state(c, DRV + QD_COUNT * j + QD_KA0, i), 0, 0, j, 0});
                        if (I.pade())
                            HE.push_back({hkey(I.ctrl_off + j, zdim + I.state_off + i), q, ci, mi, (uint16_t)slot_state(c, DRV + QD_COUNT * j + QD_KA1, i), 0, 0, 2 * nd + 4, 0});
                    }
                for (int i = 0; i < nd; ++i)
                    for (int j = i; j < nd; ++j)
                        HE.push_back({hkey(I.ctrl_off + i, I.ctrl_off + j), q, ci, mi, (uint16_t)(c.off_X + qx_haa(nd, i, j)), 0, 0, nd + j, 0});
                if (free_time) {
                    for (int i = 0; i < I.dim; ++i) {
                        HE.push_back({hkey(I.state_off + i, h->dt_off), q, ci, mi, (uint16_t)slot_state(c, QS_AHM, i), 0, 0, 2 * nd + 1, 0});
                        if (I.pade())
                            HE.push_back({hkey(h->dt_off, zdim + I.state_off + i), q, ci, mi, (uint16_t)slot_state(c, QS_AHAHM, i), 0, 0, 2 * nd + 4, 0});
                    }
                    for (int j = 0; j < nd; ++j)
                        HE.push_back({hkey(I.ctrl_off + j, h->dt_off), q, ci, mi, (uint16_t)(c.off_X + QX_HAH + j), 0, 0, 2 * nd + 2, 0});
                    HE.push_back({hkey(h->dt_off, h->dt_off), q, ci, mi, (uint16_t)(c.off_X + qx_hhh(nd)), 0, 0, 2 * nd + 2, 0});
                }
            }
        } else {
            for (int i = 0; i < I.dim; ++i) {
                JE.push_back({jkey(R0 + i, I.state_off + i), -1, q, 0, QAUX_CONST, 0, -1.0, 0, 0});
                JE.push_back({jkey(R0 + i, zdim + I.state_off + i), -1, q, 0, QAUX_CONST, 0, 1.0, 0, 0});
                JE.push_back({jkey(R0 + i, I.ctrl_off + i), -1, q, 0, QAUX_NEG_DT, 0, 0.0, 0, 0});
                if (free_time) JE.push_back({jkey(R0 + i, h->dt_off), -1, q, 0, QAUX_NEG_Z, I.ctrl_off + i, 0.0, 0, 0});
                if (free_time && h->eval_hessian)
                    HE.push_back({hkey(I.ctrl_off + i, h->dt_off), q, -1, -1, 0, QAUX_NEG_MU, R0 + i, 0, 0});
            }
        }
    }
    // ---- Jacobian structure (CSC order) and maps ---------------------------------------------------------------------
    std::sort(JE.begin(), JE.end(), [](const JEntry& a, const JEntry& b) { return a.key < b.key; });
    for (size_t k = 1; k < JE.size(); ++k)
        if (JE[k].key == JE[k - 1].key) return fail(h, QCK_EINVAL, "two integrators write the same Jacobian entry (overlapping rows?)");
    h->nnzJ = (long long)JE.size();
    h->Jr.resize(JE.size()); h->Jc.resize(JE.size());
    for (size_t k = 0; k < JE.size(); ++k) { h->Jc[k] = (int32_t)(JE[k].key / h->dyn); h->Jr[k] = (int32_t)(JE[k].key % h->dyn); }

    // ---- Hessian structure: unique keys, contributor lists ----------------------------------------------------------------
    std::stable_sort(HE.begin(), HE.end(), [](const HEntry& a, const HEntry& b) { return a.key != b.key ? a.key < b.key : a.contrib < b.contrib; });
    std::vector<int> hpos(HE.size());
    {
        long long last = -1; int pos = -1;
        for (size_t k = 0; k < HE.size(); ++k) {
            if (HE[k].key != last) { ++pos; last = HE[k].key; h->Hc.push_back((int32_t)(HE[k].key / (2 * zdim))); h->Hr.push_back((int32_t)(HE[k].key % (2 * zdim))); }
            hpos[k] = pos;
        }
    }
    h->nnzH = (long long)h->Hr.size();
    std::vector<int> ncontrib((size_t)h->nnzH, 0);
    for (size_t k = 0; k < HE.size(); ++k) ++ncontrib[hpos[k]];
    // shared positions get partial columns: one run of consecutive columns per active contributor (ascending
    // integrator order), so that a contributor's shared entries stay contiguous; the reduce kernel sums the columns of
    // a position in ascending integrator order.
    std::vector<long long> hdst(HE.size(), -1);  // destination: < nnzH direct, >= nnzH partial column
    h->npart = 0;
    {
        std::vector<size_t> shared_entries;
        for (size_t k = 0; k < HE.size(); ++k) {
            if (ncontrib[hpos[k]] > 1) {
                if (k == 0 || hpos[k] != hpos[k - 1]) { h->shared_positions.push_back(hpos[k]); h->sh_pos.push_back(hpos[k]); }
                if (HE[k].contrib >= h->ib && HE[k].contrib < h->ie) shared_entries.push_back(k);
            } else hdst[k] = hpos[k];
        }
        std::stable_sort(shared_entries.begin(), shared_entries.end(), [&](size_t a, size_t b) { return HE[a].contrib < HE[b].contrib; });
        for (size_t k : shared_entries) hdst[k] = h->nnzH + h->npart++;
        std::map<int, int> slot_of_pos;
        for (size_t i = 0; i < h->sh_pos.size(); ++i) slot_of_pos[h->sh_pos[i]] = (int)i;
        std::vector<std::vector<int>> cols(h->sh_pos.size());
        for (size_t k : shared_entries) cols[slot_of_pos[hpos[k]]].push_back((int)(hdst[k] - h->nnzH));
        h->sh_ptr.push_back(0);
        for (auto& v : cols) { for (int cidx : v) h->sh_cols.push_back(cidx); h->sh_ptr.push_back((int)h->sh_cols.size()); }
    }

    // ---- per-class maps ---------------------------------------------------------------------------------------------------------
    for (size_t ci = 0; ci < h->classes.size(); ++ci) {
        ClassHost& C = h->classes[ci];
        QckClassDev& c = C.dev;
        const int nm = c.n_members;
        std::vector<std::vector<MapEntry>> mj(nm), mh(nm);
        for (size_t k = 0; k < JE.size(); ++k)
            if (JE[k].cls == (int)ci) mj[JE[k].member].push_back({(long long)k, JE[k].src, JE[k].grp, JE[k].period});
        for (size_t k = 0; k < HE.size(); ++k)
            if (HE[k].cls == (int)ci && hdst[k] >= 0) mh[HE[k].member].push_back({hdst[k], HE[k].src, HE[k].grp, HE[k].period});
        std::vector<uint16_t> tab;
        std::vector<QckSeg> segs;
        c.nsegJ = c.nsegH = 0;
        {
            std::vector<std::vector<QckSeg>> per_member(nm);
            bool first = true;
            for (int m2 = C.member_begin; m2 < C.member_end; ++m2) {  // only active members are ever launched
                std::vector<uint16_t> t2;
                std::vector<QckSeg> sj, sh;
                std::sort(mj[m2].begin(), mj[m2].end());
                std::sort(mh[m2].begin(), mh[m2].end());
                make_segments(mj[m2], (long long)1 << 60, sj, t2);
                make_segments(mh[m2], h->nnzH, sh, t2);
                if (first) { tab = t2; c.nsegJ = (int)sj.size(); c.nsegH = (int)sh.size(); first = false; }
                if (t2 != tab || (int)sj.size() != c.nsegJ || (int)sh.size() != c.nsegH)
                    return fail(h, QCK_EINVAL, "integrators of one kind must see the same relative component order (state/control/timestep layout differs between members)");
                per_member[m2] = sj;
                per_member[m2].insert(per_member[m2].end(), sh.begin(), sh.end());
            }
            const int ns = c.nsegJ + c.nsegH;
            segs.assign((size_t)nm * std::max(ns, 1), QckSeg{0, 0, 0, 1});
            for (int m2 = C.member_begin; m2 < C.member_end; ++m2)
                for (int u = 0; u < ns; ++u) segs[(size_t)m2 * ns + u] = per_member[m2][u];
        }
        c.tab_len = (int)tab.size();
        // constants: A0 = -i H_drift, A_j = -i H_j dense, and ELL forms of A_j and A_j^H
        const int N = c.N, nd = c.nd;
        int W = 1;
        for (int q : C.members) {
            const Integ& I = h->integ[q];
            for (int j = 0; j < nd; ++j) {
                for (int r = 0; r < N; ++r) {
                    int cnt = 0, cnta = 0;
                    for (int k = 0; k < N; ++k) {
                        if (I.Hdrives[(size_t)j * N * N + r + (size_t)N * k] != 0.0) ++cnt;   // A_j[r,k]
                        if (I.Hdrives[(size_t)j * N * N + k + (size_t)N * r] != 0.0) ++cnta;  // A_j^H[r,k] = conj(A_j[k,r])
                    }
                    W = std::max(W, std::max(cnt, cnta));
                }
            }
        }
        c.W = W;
        c.cmat_stride = N * N * (1 + nd) + nd * 2 * N * W;
        c.ell_stride = nd * 2 * N * W;
        std::vector<double2> cmat((size_t)nm * c.cmat_stride, make_double2(0.0, 0.0));
        std::vector<int> ellc((size_t)nm * std::max(c.ell_stride, 1), 0);
        std::vector<int> soff(nm), coff(nm), roff(nm);
        for (int m2 = 0; m2 < nm; ++m2) {
            const Integ& I = h->integ[C.members[m2]];
            soff[m2] = I.state_off; coff[m2] = I.ctrl_off; roff[m2] = I.row_off;
            double2* base = cmat.data() + (size_t)m2 * c.cmat_stride;
            auto minus_i = [](std::complex<double> z) { return make_double2(z.imag(), -z.real()); };
            for (int e = 0; e < N * N; ++e) base[e] = minus_i(I.Hdrift[e]);
            for (int j = 0; j < nd; ++j)
                for (int e = 0; e < N * N; ++e) base[N * N + j * N * N + e] = minus_i(I.Hdrives[(size_t)j * N * N + e]);
            double2* ev = base + N * N * (1 + nd);
            int* ec = ellc.data() + (size_t)m2 * c.ell_stride;
            for (int j = 0; j < nd; ++j)
                for (int adj = 0; adj < 2; ++adj)
                    for (int r = 0; r < N; ++r) {
                        int w = 0;
                        for (int k = 0; k < N; ++k) {
                            std::complex<double> hv = adj ? std::conj(I.Hdrives[(size_t)j * N * N + k + (size_t)N * r])
                                                          : I.Hdrives[(size_t)j * N * N + r + (size_t)N * k];
                            if (hv == 0.0) continue;
                            // A_j = -i H_j ; A_j^H = +i H_j^H
                            std::complex<double> av = adj ? std::complex<double>(0, 1) * hv : std::complex<double>(0, -1) * hv;
                            size_t o = ((size_t)(j * 2 + adj) * N + r) * W + w;
                            ev[o] = make_double2(av.real(), av.imag());
                            ec[o] = k;
                            ++w;
                        }
                    }
        }
        cudaError_t e;
        qck_smem_finalize(c);
        if ((size_t)c.sm_bytes > 227 * 1024)
            return fail(h, QCK_EINVAL, "levels=%d with %d drives needs %d bytes of shared memory per knot; this build supports at most %d", c.N, c.nd, c.sm_bytes, 227 * 1024);
        if ((e = upload(tab, &c.tab, C.allocs)) != cudaSuccess || (e = upload(segs, &c.segs, C.allocs)) != cudaSuccess ||
            (e = upload(cmat, &c.cmat, C.allocs)) != cudaSuccess || (e = upload(ellc, &c.ell_col, C.allocs)) != cudaSuccess ||
            (e = upload(soff, &c.state_off, C.allocs)) != cudaSuccess || (e = upload(coff, &c.ctrl_off, C.allocs)) != cudaSuccess ||
            (e = upload(roff, &c.row_off, C.allocs)) != cudaSuccess)
            return fail(h, QCK_ECUDA, "uploading class constants: %s", cudaGetErrorString(e));
    }

    // ---- aux entries (derivative integrators inside the active range) ---------------------------------------------------------------
    for (int q = 0; q < nI; ++q) {
        const Integ& I = h->integ[q];
        if (I.quantum() || q < h->ib || q >= h->ie) continue;
        for (int i = 0; i < I.dim; ++i)
            h->aux.push_back({0, QAUX_FROW, I.row_off + i, I.state_off + i, I.ctrl_off + i, 0, 0.0});
    }
    for (size_t k = 0; k < JE.size(); ++k)
        if (JE[k].cls < 0 && JE[k].member >= h->ib && JE[k].member < h->ie)
            h->aux.push_back({1, JE[k].aux_op, (int32_t)k, JE[k].aux_i0, 0, 0, JE[k].aux_c});
    for (size_t k = 0; k < HE.size(); ++k)
        if (HE[k].cls < 0 && hdst[k] >= 0) h->aux.push_back({2, HE[k].aux_op, (int32_t)hdst[k], HE[k].aux_i0, 0, 0, 0.0});
    cudaError_t e;
    if ((e = upload(h->aux, &h->d_aux, h->allocs)) != cudaSuccess) return fail(h, QCK_ECUDA, "uploading aux entries: %s", cudaGetErrorString(e));
    h->red.n_shared = (int)h->sh_pos.size();
    if ((e = upload(h->sh_pos, &h->red.pos, h->allocs)) != cudaSuccess || (e = upload(h->sh_ptr, &h->red.ptr, h->allocs)) != cudaSuccess ||
        (e = upload(h->sh_cols, &h->red.cols, h->allocs)) != cudaSuccess)
        return fail(h, QCK_ECUDA, "uploading reduction tables: %s", cudaGetErrorString(e));
    return QCK_OK;
}

int run(qck_handle* h, uint32_t mask, const double* dZ, const double* dmu, double* dF, double* dJ, double* dH, cudaStream_t st) {
    if (!h->eval_hessian) mask &= ~QCK_EVAL_H;
    if (!dF) mask &= ~QCK_EVAL_F;
    if (!dJ) mask &= ~QCK_EVAL_J;
    if (!dH) mask &= ~QCK_EVAL_H;
    if (!mask) return QCK_OK;
    if ((mask & QCK_EVAL_H) && !dmu) return fail(h, QCK_EINVAL, "the Hessian needs the multipliers mu");
    QckLaunch L{};
    L.Z = dZ; L.mu = dmu; L.F = dF; L.J = dJ; L.H = dH; L.partial = h->dpartial;
    L.n_knots = h->T - 1; L.nnzJ = h->nnzJ; L.nnzH = h->nnzH; L.npart = h->npart; L.mask = mask;
    bool aux_done = h->aux.empty();
    const bool fuse_aux = (int)h->aux.size() <= qck_fused_aux_limit();
    int launches = 0;
    for (auto& C : h->classes) {
        if (C.member_end <= C.member_begin) continue;
        L.c = C.dev; L.member_begin = C.member_begin; L.member_end = C.member_end;
        const bool take_aux = !aux_done && fuse_aux;
        L.aux = take_aux ? h->d_aux : nullptr; L.n_aux = take_aux ? (int)h->aux.size() : 0;
        if (take_aux) aux_done = true;
        int rc = qck_launch_quantum(L, h->sm_count, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "quantum kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    if (!aux_done) {
        L.c = QckClassDev{};
        L.c.free_time = h->dt_off >= 0; L.c.dt_off = h->dt_off; L.c.zdim = h->zdim; L.c.dyn = h->dyn; L.c.dt_fixed = h->dt_fixed;
        L.aux = h->d_aux; L.n_aux = (int)h->aux.size();
        int rc = qck_launch_aux(L, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "aux kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    if ((mask & QCK_EVAL_H) && h->red.n_shared) {
        int rc = qck_launch_reduce(h->red, dH, h->dpartial, h->T - 1, h->nnzH, h->npart, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "reduce kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    h->launches += launches;
    return QCK_OK;
}

int eval_host(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) {
    if (!h) return QCK_EINVAL;
    if (!Z) return fail(h, QCK_EINVAL, "Z is NULL");
    if (h->device < 0) return fail(h, QCK_ENODEVICE, "structure-only handle (device=-1): libqcknot has no CPU evaluation path");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long nk = h->T - 1;
    uint32_t mask = (F ? QCK_EVAL_F : 0) | (J ? QCK_EVAL_J : 0) | ((H && h->eval_hessian) ? QCK_EVAL_H : 0);
    if ((mask & QCK_EVAL_H) && !mu) return fail(h, QCK_EINVAL, "the Hessian needs the multipliers mu");
    CUDA_TRY(h, cudaMemcpyAsync(h->dZ, Z, sizeof(double) * h->T * h->zdim, cudaMemcpyHostToDevice, h->stream));
    if (mask & QCK_EVAL_H) CUDA_TRY(h, cudaMemcpyAsync(h->dmu, mu, sizeof(double) * nk * h->dyn, cudaMemcpyHostToDevice, h->stream));
    int rc = run(h, mask, h->dZ, h->dmu, h->dF, h->dJ, h->dH, h->stream);
    if (rc) return rc;
    if (mask & QCK_EVAL_F) CUDA_TRY(h, cudaMemcpyAsync(F, h->dF, sizeof(double) * nk * h->dyn, cudaMemcpyDeviceToHost, h->stream));
    if (mask & QCK_EVAL_J) CUDA_TRY(h, cudaMemcpyAsync(J, h->dJ, sizeof(double) * nk * h->nnzJ, cudaMemcpyDeviceToHost, h->stream));
    if (mask & QCK_EVAL_H) CUDA_TRY(h, cudaMemcpyAsync(H, h->dH, sizeof(double) * nk * h->nnzH, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QCK_OK;
}

}  // namespace

extern "C" {

const char* qck_version(void) { return "qcknot 0.1 (sm_100a)"; }

const char* qck_last_error(const qck_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int qck_create(const qck_problem_desc* d, qck_handle** out) {
    if (!out) return fail(nullptr, QCK_EINVAL, "out is NULL");
    *out = nullptr;
    if (!d || !d->integrators || d->n_integrators <= 0) return fail(nullptr, QCK_EINVAL, "empty problem description");
    if (d->T < 2) return fail(nullptr, QCK_EINVAL, "T must be >= 2 (got %lld)", (long long)d->T);
    if (d->zdim <= 0 || d->dt_off >= d->zdim) return fail(nullptr, QCK_EINVAL, "bad zdim/dt_off");
    const bool structure_only = d->device == -1;  // sizes + structures only (what Ipopt asks for at set-up); every eval fails
    int ndev = 0;
    cudaError_t ce = cudaSuccess;
    cudaDeviceProp prop{};
    prop.multiProcessorCount = 148;
    if (!structure_only) {
        ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0)
            return fail(nullptr, QCK_ENODEVICE, "no CUDA device (%s); libqcknot has no CPU fallback", ce == cudaSuccess ? "count=0" : cudaGetErrorString(ce));
        if (d->device < 0 || d->device >= ndev) return fail(nullptr, QCK_ENODEVICE, "device %d out of range (%d devices)", d->device, ndev);
        if ((ce = cudaGetDeviceProperties(&prop, d->device)) != cudaSuccess) return fail(nullptr, QCK_ECUDA, "%s", cudaGetErrorString(ce));
        if (prop.major != 10) return fail(nullptr, QCK_ENODEVICE, "device %d is sm_%d%d; libqcknot is built for sm_100a only", d->device, prop.major, prop.minor);
    }

    qck_handle* h = new (std::nothrow) qck_handle();
    if (!h) return fail(nullptr, QCK_ENOMEM, "out of host memory");
    auto bail = [&](int code) { g_create_error = h->err; qck_destroy(h); return code; };
    h->device = d->device; h->sm_count = prop.multiProcessorCount;
    h->T = d->T; h->zdim = d->zdim; h->dt_off = d->dt_off < 0 ? -1 : d->dt_off; h->dt_fixed = d->dt_fixed;
    h->eval_hessian = d->eval_hessian ? 1 : 0;
    h->ib = d->integ_begin; h->ie = d->integ_end;
    if (h->ib == 0 && h->ie == 0) h->ie = d->n_integrators;
    if (h->ib < 0 || h->ie > d->n_integrators || h->ib > h->ie) { fail(h, QCK_EINVAL, "bad integrator range [%d,%d)", h->ib, h->ie); return bail(QCK_EINVAL); }
    int row = 0;
    for (int q = 0; q < d->n_integrators; ++q) {
        const qck_integrator_desc& s = d->integrators[q];
        Integ I;
        I.kind = s.kind; I.order = s.order; I.N = s.levels; I.nd = s.n_drives;
        I.state_off = s.state_off; I.state_len = s.state_len; I.ctrl_off = s.ctrl_off;
        if (I.kind < 0 || I.kind > QCK_DERIVATIVE) { fail(h, QCK_EINVAL, "integrator %d: unknown kind %d", q, I.kind); return bail(QCK_EINVAL); }
        if (I.quantum()) {
            if (I.N < 1 || I.nd < 0 || I.nd > QCK_MAX_DRIVES) { fail(h, QCK_EINVAL, "integrator %d: levels=%d n_drives=%d unsupported (max %d drives)", q, I.N, I.nd, QCK_MAX_DRIVES); return bail(QCK_EINVAL); }
            I.nc = I.unitary() ? I.N : 1;
            I.dim = 2 * I.N * I.nc;
            if (I.state_len != I.dim) { fail(h, QCK_EINVAL, "integrator %d: state_len %d != %d", q, I.state_len, I.dim); return bail(QCK_EINVAL); }
            if (I.pade() && I.order != 4) { fail(h, QCK_EINVAL, "integrator %d: Pade order %d not supported by this build (4 only)", q, I.order); return bail(QCK_EINVAL); }
            if (!I.pade()) { fail(h, QCK_EINVAL, "integrator %d: exponential integrators not supported by this build yet", q); return bail(QCK_EINVAL); }
            if (I.nd > 0 && !s.H_drives) { fail(h, QCK_EINVAL, "integrator %d: H_drives is NULL", q); return bail(QCK_EINVAL); }
            if (I.ctrl_off < 0 || I.ctrl_off + I.nd > h->zdim) { fail(h, QCK_EINVAL, "integrator %d: drive component out of range", q); return bail(QCK_EINVAL); }
            size_t nn = (size_t)I.N * I.N;
            I.Hdrift.assign(nn, 0.0);
            if (s.H_drift) for (size_t e = 0; e < nn; ++e) I.Hdrift[e] = {s.H_drift[2 * e], s.H_drift[2 * e + 1]};
            I.Hdrives.resize(nn * I.nd);
            for (size_t e = 0; e < nn * I.nd; ++e) I.Hdrives[e] = {s.H_drives[2 * e], s.H_drives[2 * e + 1]};
        } else {
            I.dim = I.state_len;
            if (I.dim <= 0 || I.ctrl_off < 0 || I.ctrl_off + I.dim > h->zdim) { fail(h, QCK_EINVAL, "integrator %d: derivative components out of range", q); return bail(QCK_EINVAL); }
        }
        if (I.state_off < 0 || I.state_off + I.state_len > h->zdim) { fail(h, QCK_EINVAL, "integrator %d: state component out of range", q); return bail(QCK_EINVAL); }
        I.row_off = row;
        row += I.dim;
        h->integ.push_back(std::move(I));
    }
    h->dyn = row;
    if (!structure_only && ((ce = cudaSetDevice(h->device)) != cudaSuccess || (ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)) {
        fail(h, QCK_ECUDA, "%s", cudaGetErrorString(ce)); return bail(QCK_ECUDA);
    }
    g_structure_only = structure_only;
    int rc = build(h);
    g_structure_only = false;
    if (rc) return bail(rc);
    if (structure_only) { *out = h; return QCK_OK; }
    const long long nk = h->T - 1;
    struct { double** p; long long n; } bufs[] = {{&h->dZ, h->T * h->zdim}, {&h->dmu, nk * h->dyn}, {&h->dF, nk * h->dyn},
                                                  {&h->dJ, nk * h->nnzJ}, {&h->dH, nk * std::max<long long>(h->nnzH, 1)},
                                                  {&h->dpartial, nk * std::max(h->npart, 1)}};
    for (auto& b : bufs) {
        void* p = nullptr;
        if ((ce = cudaMalloc(&p, sizeof(double) * (size_t)b.n)) != cudaSuccess) { fail(h, QCK_ENOMEM, "device allocation of %lld doubles failed: %s", b.n, cudaGetErrorString(ce)); return bail(QCK_ENOMEM); }
        h->allocs.push_back(p);
        cudaMemset(p, 0, sizeof(double) * (size_t)b.n);
        *b.p = static_cast<double*>(p);
    }
    if ((ce = cudaDeviceSynchronize()) != cudaSuccess) { fail(h, QCK_ECUDA, "%s", cudaGetErrorString(ce)); return bail(QCK_ECUDA); }
    *out = h;
    return QCK_OK;
}

void qck_destroy(qck_handle* h) {
    if (!h) return;
    if (h->device < 0) { delete h; return; }
    cudaSetDevice(h->device);
    for (auto& C : h->classes) for (void* p : C.allocs) cudaFree(p);
    for (void* p : h->allocs) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int qck_sizes(const qck_handle* h, int64_t* dyn, int64_t* nnzJ, int64_t* nnzH) {
    if (!h) return QCK_EINVAL;
    if (dyn) *dyn = h->dyn;
    if (nnzJ) *nnzJ = h->nnzJ;
    if (nnzH) *nnzH = h->nnzH;
    return QCK_OK;
}

int qck_jacobian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols) {
    if (!h || !rows || !cols) return QCK_EINVAL;
    for (long long t = 0; t < h->T - 1; ++t)
        for (long long k = 0; k < h->nnzJ; ++k) {
            rows[t * h->nnzJ + k] = h->Jr[k] + (t + knot_offset) * h->dyn + 1;
            cols[t * h->nnzJ + k] = h->Jc[k] + (t + knot_offset) * h->zdim + 1;
        }
    return QCK_OK;
}

int qck_hessian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols) {
    if (!h || !rows || !cols) return QCK_EINVAL;
    for (long long t = 0; t < h->T - 1; ++t)
        for (long long k = 0; k < h->nnzH; ++k) {
            rows[t * h->nnzH + k] = h->Hr[k] + (t + knot_offset) * h->zdim + 1;
            cols[t * h->nnzH + k] = h->Hc[k] + (t + knot_offset) * h->zdim + 1;
        }
    return QCK_OK;
}

int qck_eval_residual(qck_handle* h, const double* Z, double* F) { return eval_host(h, Z, nullptr, F, nullptr, nullptr); }
int qck_eval_jacobian(qck_handle* h, const double* Z, double* J) { return eval_host(h, Z, nullptr, nullptr, J, nullptr); }
int qck_eval_hessian(qck_handle* h, const double* Z, const double* mu, double* H) {
    if (h && !h->eval_hessian) return fail(h, QCK_EINVAL, "handle was created with eval_hessian=0");
    return eval_host(h, Z, mu, nullptr, nullptr, H);
}
int qck_eval_all(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) { return eval_host(h, Z, mu, F, J, H); }

int qck_eval_device(qck_handle* h, uint32_t mask, const double* dZ, const double* dmu, double* dF, double* dJ, double* dH, void* stream) {
    if (!h) return QCK_EINVAL;
    if (!dZ) return fail(h, QCK_EINVAL, "dZ is NULL");
    if (h->device < 0) return fail(h, QCK_ENODEVICE, "structure-only handle (device=-1): libqcknot has no CPU evaluation path");
    CUDA_TRY(h, cudaSetDevice(h->device));
    uint32_t m = mask;
    if (!(m & QCK_EVAL_F)) dF = nullptr;
    if (!(m & QCK_EVAL_J)) dJ = nullptr;
    if (!(m & QCK_EVAL_H)) dH = nullptr;
    return run(h, m, dZ, dmu, dF, dJ, dH, stream ? static_cast<cudaStream_t>(stream) : h->stream);
}

int qck_device_buffers(qck_handle* h, double** dZ, double** dmu, double** dF, double** dJ, double** dH) {
    if (!h) return QCK_EINVAL;
    if (dZ) *dZ = h->dZ;
    if (dmu) *dmu = h->dmu;
    if (dF) *dF = h->dF;
    if (dJ) *dJ = h->dJ;
    if (dH) *dH = h->dH;
    return QCK_OK;
}

int qck_synchronize(qck_handle* h) {
    if (!h) return QCK_EINVAL;
    if (h->device < 0) return QCK_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QCK_OK;
}

int qck_shared_hessian_positions(const qck_handle* h, int64_t* count, int64_t* pos) {
    if (!h || !count) return QCK_EINVAL;
    *count = (int64_t)h->shared_positions.size();
    if (pos) for (size_t k = 0; k < h->shared_positions.size(); ++k) pos[k] = h->shared_positions[k];
    return QCK_OK;
}

int qck_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return QCK_EINVAL;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return QCK_ECUDA; }
    return QCK_OK;
}
int qck_host_unregister(void* p) {
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return QCK_ECUDA; }
    return QCK_OK;
}

int qck_launch_count(const qck_handle* h, int64_t* launches) {
    if (!h || !launches) return QCK_EINVAL;
    *launches = h->launches;
    return QCK_OK;
}

}  // extern "C"
