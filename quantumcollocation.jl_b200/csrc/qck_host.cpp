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

#include "qck_handle.h"

namespace {

thread_local std::string g_create_error;

typedef QckInteg Integ;
typedef QckClassHost ClassHost;

// structure-only handles (device < 0) never touch the device: their tables stay on the host
template <class T>
cudaError_t upload(const qck_handle* h, const std::vector<T>& v, const T** out, std::vector<void*>& allocs) {
    *out = nullptr;
    if (v.empty() || h->device < 0) return cudaSuccess;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    allocs.push_back(d);
    e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = static_cast<const T*>(d);
    return e;
}

}  // namespace

int qck_fail(qck_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

namespace {

#define fail qck_fail
#define CUDA_TRY QCK_CUDA_TRY

// One structural nonzero of the per-knot Jacobian / Hessian block.
struct Ent {
    long long key;     // col * nrows + row  (CSC order)
    int contrib;       // integrator index
    int cls, member;   // quantum class / member, or cls = -1 for a derivative integrator (member = integrator index)
    int qid, idx;      // output quantity and element index inside it (quantum)
    int period;        // repeat period of the quantity (kron(I_N, .) blocks), 0 = none
    int aux_op, aux_i0;
    double aux_c;
};
struct MapEnt {
    long long dst;
    int qid, idx, period;
    bool operator<(const MapEnt& o) const { return dst < o.dst; }
};

// Image placement of one member's output quantities of one value array + its segments.
// Groups (one per quantity) must be affine in the element index: dst(i) = d0 + i * stride.  Stride-1 groups get
// their own contiguous image range, allocated in destination order so that destination-adjacent groups are also
// image-adjacent (they merge into one segment); groups with a common stride > 1 that interleave share one range.
struct Run { int dst, len, img, period; };
// parity: every run gets an image offset of the same parity as its destination (rs3 kernel: the write-out of a run is then
// [scalar head] + 16-byte aligned bulk copy + [scalar tail] whatever the parity of the knot block's base address).
bool place_array(std::vector<MapEnt> e, long long split, int& cursor, short* base, short* stride, std::vector<Run>& segs,
                 std::string& why, int* dst0 = nullptr, bool parity = false, bool no_image = false) {
    if (e.empty()) return true;
    struct Group { int qid; long long d0; int s, n, period; };
    std::map<int, std::vector<MapEnt>> by_q;
    for (auto& x : e) by_q[x.qid].push_back(x);
    std::vector<Group> groups;
    for (auto& kv : by_q) {
        auto& v = kv.second;
        std::sort(v.begin(), v.end(), [](const MapEnt& a, const MapEnt& b) { return a.idx < b.idx; });
        Group g{kv.first, v[0].dst, 1, (int)v.size(), v[0].period};
        if (v.size() > 1) g.s = (int)(v[1].dst - v[0].dst);
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].idx != (int)i || v[i].dst != g.d0 + (long long)i * g.s || g.s < 1) { why = "an output quantity is not laid out affinely in the structure"; return false; }
        if (g.period > 0 && (g.s != 1 || g.n % g.period)) { why = "a repeated block is not contiguous in the structure"; return false; }
        groups.push_back(g);
    }
    std::sort(groups.begin(), groups.end(), [](const Group& a, const Group& b) { return a.d0 < b.d0; });
    struct Family { long long d0; int s; int base; };
    std::vector<Family> fams;
    long long run_end = -2;      // destination just past the previous stride-1, non-periodic group
    for (auto& g : groups) {
        int b;
        if (g.s == 1) {
            const bool periodic = g.period > 0 && g.period < g.n;
            const bool glue = !periodic && run_end == g.d0 && ((g.d0 < split) == (g.d0 - 1 < split));
            b = glue ? cursor : (parity ? cursor + (int)((cursor ^ g.d0) & 1) : ((cursor + 1) & ~1));
            cursor = b + (periodic ? g.period : g.n);
            run_end = periodic ? -2 : g.d0 + g.n;
        } else {
            b = -1;
            for (auto& f : fams)
                if (f.s == g.s && g.d0 >= f.d0 && g.d0 < f.d0 + f.s) b = f.base + (int)(g.d0 - f.d0);
            if (b < 0) {
                b = parity ? cursor + (int)((cursor ^ g.d0) & 1) : ((cursor + 1) & ~1);
                fams.push_back({g.d0, g.s, b});
                cursor = b + g.n * g.s;
            }
            run_end = -2;
        }
        if (cursor > 32000 && !no_image) { why = "output image too large"; return false; }
        if (cursor > 32000) { b = 0; cursor = 32000; }  // (large-level classes never stage an image: only destinations and strides matter)
        base[g.qid] = (short)b;
        stride[g.qid] = (short)g.s;
        if (dst0) dst0[g.qid] = (int)g.d0;  // destination of element 0 (column kernel: values go straight to the arrays)
    }
    // segments: maximal runs contiguous in both destination and image
    std::sort(e.begin(), e.end());
    auto img_of = [&](const MapEnt& x) { return base[x.qid] + (x.period > 0 ? x.idx % x.period : x.idx) * stride[x.qid]; };
    size_t k = 0;
    while (k < e.size()) {
        const bool periodic = e[k].period > 0;
        size_t k2 = k + 1;
        if (periodic) {
            while (k2 < e.size() && e[k2].qid == e[k].qid && e[k2].dst == e[k2 - 1].dst + 1) ++k2;
            segs.push_back(Run{(int)e[k].dst, (int)(k2 - k), img_of(e[k]), e[k].period});
        } else {
            while (k2 < e.size() && e[k2].period == 0 && e[k2].dst == e[k2 - 1].dst + 1 && img_of(e[k2]) == img_of(e[k2 - 1]) + 1 &&
                   ((e[k2].dst < split) == (e[k].dst < split)))
                ++k2;
            segs.push_back(Run{(int)e[k].dst, (int)(k2 - k), img_of(e[k]), (int)(k2 - k)});
        }
        if (!parity && !no_image && (segs.back().img & 1)) { why = "internal: odd segment image offset"; return false; }
        if (parity && ((segs.back().img ^ segs.back().dst) & 1)) { why = "internal: image / destination parity mismatch"; return false; }
        k = k2;
    }
    return true;
}

// Cut runs into write-out units and balance them over the CTA's warps (longest-processing-time first).
std::vector<QckSeg> balance_units(const std::vector<Run> (&runs)[3], int nwarps, int* hdr /* QCK_SEG_HDR */) {
    struct Unit { QckSeg s; long long cost; };
    std::vector<Unit> units;
    long long total = 0;
    for (int arr = 0; arr < 3; ++arr) for (auto& r : runs[arr]) total += r.len;
    const long long target = std::max<long long>(64, total / (2 * nwarps));
    for (int arr = 0; arr < 3; ++arr)
        for (auto& r : runs[arr]) {
            if (r.period < r.len) {  // repeated block: split by whole repetitions
                const int nrep = r.len / r.period;
                int per = (int)std::max<long long>(1, target / r.period);
                for (int r0 = 0; r0 < nrep; r0 += per) {
                    const int k = std::min(per, nrep - r0);
                    units.push_back({{r.dst + r0 * r.period, r.period, r.img | (k << 16), arr}, (long long)k * r.period});
                }
            } else {  // plain run: split into even-length pieces (keeps the 16-byte alignment parity of every piece)
                int piece = (int)((target + 1) & ~1LL);
                for (int o = 0; o < r.len; o += piece) {
                    const int k = std::min(piece, r.len - o);
                    units.push_back({{r.dst + o, k, (r.img + o) | (1 << 16), arr}, (long long)k});
                }
            }
        }
    std::vector<size_t> order(units.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return units[a].cost > units[b].cost; });
    std::vector<long long> load(nwarps, 0);
    std::vector<std::vector<size_t>> mine(nwarps);
    for (size_t i : order) {
        int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[w] += units[i].cost + 24;  // per-unit set-up cost
        mine[w].push_back(i);
    }
    std::vector<QckSeg> out;
    for (int w = 0; w < nwarps; ++w) {
        hdr[w] = (int)out.size();
        std::sort(mine[w].begin(), mine[w].end());
        for (size_t i : mine[w]) out.push_back(units[i].s);
    }
    for (int w = nwarps; w < QCK_SEG_HDR; ++w) hdr[w] = (int)out.size();
    for (auto& sgm : out) sgm.arr |= (32 % std::max(sgm.n / 2, 1)) << 8;
    return out;
}

// rs3 kernel: units of the F + J part (phase 1) and of the Hessian part (phase 2), each balanced over the knot's three warps;
// hdr = [phase 1: first unit of warp 0, 1, 2, end | phase 2: likewise].  One lane issues one unit, so units are kept small
// enough that all lanes of the three warps take part: kron blocks go two repetitions per unit, plain runs <= 512 doubles.
std::vector<QckSeg> rs3_units(const std::vector<Run> (&runs)[3], long long nnzH, int* hdr /* QCK_SEG_HDR */, std::string& why) {
    std::vector<QckSeg> out;
    for (int phase = 0; phase < 2; ++phase) {
        struct Unit { QckSeg s; long long cost; };
        std::vector<Unit> units;
        for (int arr = phase == 0 ? 0 : 2; arr < (phase == 0 ? 2 : 3); ++arr)
            for (auto& r : runs[arr]) {
                if (arr == 2 && r.dst >= nnzH) { why = "shared Hessian entries (partial columns) are not supported by the rs3 kernel"; return {}; }
                if (r.period < r.len) {
                    const int nrep = r.len / r.period;
                    for (int r0 = 0; r0 < nrep; r0 += 2) {
                        const int k = std::min(2, nrep - r0);
                        units.push_back({{r.dst + r0 * r.period, r.period, r.img | (k << 16), arr}, (long long)k * r.period});
                    }
                } else {
                    for (int o = 0; o < r.len; o += 512) {
                        const int k = std::min(512, r.len - o);
                        units.push_back({{r.dst + o, k, (r.img + o) | (1 << 16), arr}, (long long)k});
                    }
                }
            }
        std::vector<size_t> order(units.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return units[a].cost > units[b].cost; });
        long long load[3] = {0, 0, 0};
        std::vector<size_t> mine[3];
        for (size_t i : order) {
            const int w = (int)(std::min_element(load, load + 3) - load);
            load[w] += units[i].cost + 64;
            mine[w].push_back(i);
        }
        for (int w = 0; w < 3; ++w) {
            hdr[4 * phase + w] = (int)out.size();
            for (size_t i : mine[w]) out.push_back(units[i].s);
        }
        hdr[4 * phase + 3] = (int)out.size();
    }
    hdr[8] = hdr[9] = hdr[10] = (int)out.size();
    // third list: phase 1 split by the moment the values are complete -- the kron blocks (one repetition per unit: one lane, one
    // bulk copy) right after A^2, the rest at the end of the phase: hdr[11] .. hdr[12] blocks, hdr[12] .. hdr[13] rest
    hdr[11] = (int)out.size();
    for (int arr = 0; arr < 2; ++arr)
        for (auto& r : runs[arr])
            if (r.period < r.len)
                for (int r0 = 0; r0 < r.len / r.period; ++r0) out.push_back({r.dst + r0 * r.period, r.period, r.img | (1 << 16), arr});
    hdr[12] = (int)out.size();
    for (int arr = 0; arr < 2; ++arr)
        for (auto& r : runs[arr])
            if (!(r.period < r.len))
                for (int o = 0; o < r.len; o += 512) out.push_back({r.dst + o, std::min(512, r.len - o), (r.img + o) | (1 << 16), arr});
    hdr[13] = (int)out.size();
    for (int w = 14; w < QCK_SEG_HDR; ++w) hdr[w] = (int)out.size();
    return out;
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
            c.pade_m = 0;
            if (I.pade()) {  // c_k = (2m-k)! m! / ((2m)! k! (m-k)!)  ->  r_k = c_{k+1}/c_k = (m-k) / ((2m-k)(k+1))
                const int m = I.order / 2;
                c.pade_m = m;
                for (int k2 = 0; k2 < m; ++k2) c.pade_r[k2] = (double)(m - k2) / ((double)(2 * m - k2) * (k2 + 1));
            }
            qck_scratch_layout(c);
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
    std::vector<Ent> JE, HE;
    auto jkey = [&](int row, int col) { return (long long)col * h->dyn + row; };
    auto hkey = [&](int r, int c) { if (r > c) std::swap(r, c); return (long long)c * (2 * zdim) + r; };
    for (int q = 0; q < nI; ++q) {
        const Integ& I = h->integ[q];
        const int R0 = I.row_off;
        if (I.quantum()) {
            const int ci = cls_idx[q], mi = mem_idx[q];
            const int N = I.N, n2 = 2 * N, blk = n2 * n2, nd = I.nd;
            auto J = [&](int row, int col, int qid, int idx, int period) { JE.push_back({jkey(row, col), q, ci, mi, qid, idx, period, 0, 0, 0.0}); };
            auto Hh = [&](int r, int c2, int qid, int idx) { HE.push_back({hkey(r, c2), q, ci, mi, qid, idx, 0, 0, 0, 0.0}); };
            // Jacobian: state_t block (-F or -E), state_t+1 block (+B or identity), controls, timestep
            for (int cb = 0; cb < I.nc; ++cb)
                for (int r = 0; r < n2; ++r)
                    for (int qq = 0; qq < n2; ++qq) {
                        J(R0 + cb * n2 + qq, I.state_off + cb * n2 + r, QO_ISOF, cb * blk + r * n2 + qq, blk);
                        if (I.pade()) J(R0 + cb * n2 + qq, zdim + I.state_off + cb * n2 + r, QO_ISOB, cb * blk + r * n2 + qq, blk);
                    }
            if (!I.pade())
                for (int i = 0; i < I.dim; ++i) J(R0 + i, zdim + I.state_off + i, QO_ONE, i, 1);
            for (int j = 0; j < nd; ++j)
                for (int i = 0; i < I.dim; ++i) J(R0 + i, I.ctrl_off + j, QO_TA + j, i, 0);
            if (free_time)
                for (int i = 0; i < I.dim; ++i) J(R0 + i, h->dt_off, QO_TH, i, 0);
            if (h->eval_hessian) {
                for (int j = 0; j < nd; ++j)
                    for (int i = 0; i < I.dim; ++i) {
                        Hh(I.state_off + i, I.ctrl_off + j, QO_KA0 + j, i);
                        if (I.pade()) Hh(I.ctrl_off + j, zdim + I.state_off + i, QO_KA1 + j, i);
                    }
                for (int i = 0; i < nd; ++i)
                    for (int j = i; j < nd; ++j) Hh(I.ctrl_off + i, I.ctrl_off + j, qo_haa(i, j), 0);
                if (free_time) {
                    for (int i = 0; i < I.dim; ++i) {
                        Hh(I.state_off + i, h->dt_off, QO_KH0, i);
                        if (I.pade()) Hh(h->dt_off, zdim + I.state_off + i, QO_KH1, i);
                    }
                    for (int j = 0; j < nd; ++j) Hh(I.ctrl_off + j, h->dt_off, QO_HAH + j, 0);
                    Hh(h->dt_off, h->dt_off, QO_HHH, 0);
                }
            }
        } else {
            for (int i = 0; i < I.dim; ++i) {
                JE.push_back({jkey(R0 + i, I.state_off + i), q, -1, q, 0, 0, 0, QAUX_CONST, 0, -1.0});
                JE.push_back({jkey(R0 + i, zdim + I.state_off + i), q, -1, q, 0, 0, 0, QAUX_CONST, 0, 1.0});
                JE.push_back({jkey(R0 + i, I.ctrl_off + i), q, -1, q, 0, 0, 0, QAUX_NEG_DT, 0, 0.0});
                if (free_time) JE.push_back({jkey(R0 + i, h->dt_off), q, -1, q, 0, 0, 0, QAUX_NEG_Z, I.ctrl_off + i, 0.0});
                if (free_time && h->eval_hessian)
                    HE.push_back({hkey(I.ctrl_off + i, h->dt_off), q, -1, q, 0, 0, 0, QAUX_NEG_MU, R0 + i, 0.0});
            }
        }
    }
    // ---- Jacobian structure (CSC order) ---------------------------------------------------------------------------------------
    std::sort(JE.begin(), JE.end(), [](const Ent& a, const Ent& b) { return a.key < b.key; });
    for (size_t k = 1; k < JE.size(); ++k)
        if (JE[k].key == JE[k - 1].key) return fail(h, QCK_EINVAL, "two integrators write the same Jacobian entry (overlapping rows?)");
    h->nnzJ = (long long)JE.size();
    h->Jr.resize(JE.size()); h->Jc.resize(JE.size());
    for (size_t k = 0; k < JE.size(); ++k) { h->Jc[k] = (int32_t)(JE[k].key / h->dyn); h->Jr[k] = (int32_t)(JE[k].key % h->dyn); }

    // ---- Hessian structure: unique keys, contributor lists ----------------------------------------------------------------
    std::stable_sort(HE.begin(), HE.end(), [](const Ent& a, const Ent& b) { return a.key != b.key ? a.key < b.key : a.contrib < b.contrib; });
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

    // ---- structure-order policy (qck_problem_desc.structure_order): caller-visible order + gather tables from the CSC arrays ---------
    if (h->pub.order != QCK_ORDER_CSC) {
        QckPublicOrder& P = h->pub;
        if (h->ib != 0 || h->ie != nI || h->exclude_shared)
            return fail(h, QCK_EINVAL, "structure_order other than CSC needs an unsharded handle (all integrators, one GPU)");
        const long long zz = 2LL * zdim;
        std::vector<size_t> ord(JE.size());
        for (size_t k = 0; k < ord.size(); ++k) ord[k] = k;
        auto jrow = [&](size_t k) { return JE[k].key % h->dyn; };
        auto jcol = [&](size_t k) { return JE[k].key / h->dyn; };
        if (P.order == QCK_ORDER_ROW_MAJOR)
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return jrow(a) != jrow(b) ? jrow(a) < jrow(b) : jcol(a) < jcol(b); });
        else  // per integrator, column-major inside (JE is in CSC order already)
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return JE[a].contrib < JE[b].contrib; });
        for (size_t k : ord) { P.srcJ.push_back((int)k); P.Jr.push_back((int32_t)jrow(k)); P.Jc.push_back((int32_t)jcol(k)); }
        P.nnzJ = (long long)P.srcJ.size();
        if (P.order == QCK_ORDER_ROW_MAJOR) {
            std::vector<int> pos((size_t)h->nnzH);
            for (size_t k = 0; k < pos.size(); ++k) pos[k] = (int)k;
            std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return h->Hr[a] != h->Hr[b] ? h->Hr[a] < h->Hr[b] : h->Hc[a] < h->Hc[b]; });
            std::vector<int> where((size_t)h->nnzH);
            for (size_t k = 0; k < pos.size(); ++k) { P.srcH.push_back(pos[k]); P.Hr.push_back(h->Hr[pos[k]]); P.Hc.push_back(h->Hc[pos[k]]); where[pos[k]] = (int)k; }
            for (long long sp : h->shared_positions) P.shared.push_back(where[(size_t)sp]);
            std::sort(P.shared.begin(), P.shared.end());
        } else {
            std::vector<size_t> ho(HE.size());
            for (size_t k = 0; k < ho.size(); ++k) ho[k] = k;
            std::stable_sort(ho.begin(), ho.end(), [&](size_t a, size_t b) { return HE[a].contrib != HE[b].contrib ? HE[a].contrib < HE[b].contrib : HE[a].key < HE[b].key; });
            for (size_t k : ho) {
                if (hdst[k] < 0) return fail(h, QCK_EINVAL, "internal: Hessian entry without a destination");
                P.srcH.push_back((int)hdst[k]);  // canonical position, or nnzH + partial column (this integrator's own contribution)
                P.Hr.push_back((int32_t)(HE[k].key % zz)); P.Hc.push_back((int32_t)(HE[k].key / zz));
            }
        }
        P.nnzH = (long long)P.srcH.size();
    }

    // ---- per-class image placement, segments and constants ---------------------------------------------------------------------
    for (size_t ci = 0; ci < h->classes.size(); ++ci) {
        ClassHost& C = h->classes[ci];
        QckClassDev& c = C.dev;
        const int nm = c.n_members;
        std::vector<std::vector<MapEnt>> mj(nm), mh(nm);
        for (size_t k = 0; k < JE.size(); ++k)
            if (JE[k].cls == (int)ci) mj[JE[k].member].push_back({(long long)k, JE[k].qid, JE[k].idx, JE[k].period});
        for (size_t k = 0; k < HE.size(); ++k)
            if (HE[k].cls == (int)ci && hdst[k] >= 0) mh[HE[k].member].push_back({hdst[k], HE[k].qid, HE[k].idx, 0});
        std::vector<QckSeg> segs;
        c.nseg = 0;
        c.n_tbuf = C.member_end - C.member_begin > 1 ? 2 : 1;
        c.threads = qck_pick_threads(c);
        const int nwarps = c.threads / 32;
        // Pade-4 classes whose image / matrices do not fit the tiled kernel's shared memory (up to 32 levels) run on the large-level
        // kernel (qck_big.cu): operands in shared memory, every output straight to its destination (QCK_BIG=1 forces it from
        // 5 levels on, for A/B runs)
        static const int big_knob = getenv("QCK_BIG") ? atoi(getenv("QCK_BIG")) : 0;
        const bool big_ok = (c.kind == QCK_UNITARY_PADE || c.kind == QCK_KET_PADE) && c.order == 4 && c.N >= 5 && c.N <= 32 &&
                            qck_big_smem(c) <= 227 * 1024;
        // 9-level Pade-4 unitaries, one active member, nothing shared with other integrators: parity-matched image placement and
        // per-phase unit tables for the row-slice kernel (QCK_RS3=1, default: one warp per knot) or its three-warps-per-knot
        // variant (QCK_RS3=5..7 knots per CTA; parity-green, measured slower: profiles/); QCK_RS3=0: the tiled kernel
        static const int rs3_knob = getenv("QCK_RS3") ? atoi(getenv("QCK_RS3")) : 1;
        c.antiherm = 1;  // every member's Hamiltonians are Hermitian  ->  A(a) = -i H(a) is anti-Hermitian
        for (int q : C.members) {
            const Integ& I = h->integ[q];
            for (int j = -1; j < c.nd && c.antiherm; ++j) {
                const std::complex<double>* Hm = j < 0 ? I.Hdrift.data() : I.Hdrives.data() + (size_t)j * c.N * c.N;
                for (int r = 0; r < c.N && c.antiherm; ++r)
                    for (int k = 0; k <= r; ++k)
                        if (Hm[r + (size_t)c.N * k] != std::conj(Hm[k + (size_t)c.N * r])) { c.antiherm = 0; break; }
            }
        }
        c.rs3 = ((rs3_knob == 1 || (rs3_knob >= 5 && rs3_knob <= 7)) && c.kind == QCK_UNITARY_PADE && c.order == 4 && c.N == 9 && c.nd >= 1 && c.nd <= 4 &&
                 C.member_end - C.member_begin == 1 && h->npart == 0) ? rs3_knob : 0;
        // exponential unitaries, 9 levels, Hermitian Hamiltonians, one active member: the spectral kernel (qck_expeig.cu) with the same
        // placement and unit tables as the row-slice kernel (QCK_EXPEIG=0: the scaling-and-squaring kernel)
        static const int eig_knob = getenv("QCK_EXPEIG") ? atoi(getenv("QCK_EXPEIG")) : 1;
        c.eig = (eig_knob && c.kind == QCK_UNITARY_EXP && c.N == 9 && c.nd >= 1 && c.nd <= 4 && c.antiherm && C.member_end - C.member_begin == 1 &&
                 h->npart == 0) ? 1 : 0;
        if (c.eig) c.rs3 = 1;
        // exponential unitaries and kets of 2..4 levels with Hermitian Hamiltonians: the spectral column kernels (qck_colexp.cu;
        // QCK_COLEXP=0: the scaling-and-squaring kernel)
        static const int colexp_knob = getenv("QCK_COLEXP") ? atoi(getenv("QCK_COLEXP")) : 1;
        c.colexp = (colexp_knob && (c.kind == QCK_UNITARY_EXP || c.kind == QCK_KET_EXP) && c.N >= 2 && c.N <= 4 && c.nd >= 1 && c.nd <= 4 && c.antiherm) ? 1 : 0;
        // every other exponential class with Hermitian Hamiltonians up to 16 levels (5..8, 10..16 levels, 9-level kets and ensembles):
        // the generic spectral kernel (qck_genexp.cu; QCK_GENEXP=0: the scaling-and-squaring kernel)
        static const int genexp_knob = getenv("QCK_GENEXP") ? atoi(getenv("QCK_GENEXP")) : 1;
        c.genexp = (genexp_knob && !c.eig && !c.colexp && (c.kind == QCK_UNITARY_EXP || c.kind == QCK_KET_EXP) && c.N >= 5 && c.N <= 16 && c.nd >= 1 &&
                    c.nd <= 4 && c.antiherm && qck_genexp_warp_bytes(c.N, c.kind == QCK_KET_EXP ? 1 : c.N, c.nd) <= 227 * 1024) ? 1 : 0;
        for (int q2 = 0; q2 < QO_COUNT; ++q2) { c.pl_base[q2] = -1; c.pl_stride[q2] = 0; }
        std::vector<int> qdst((size_t)nm * QO_COUNT, -1);  // per member: first destination of every output quantity
        {
            std::vector<std::vector<QckSeg>> per_member(nm);
            std::vector<std::vector<int>> hdrs(nm, std::vector<int>(QCK_SEG_HDR, 0));
            bool first = true;
            for (int m2 = C.member_begin; m2 < C.member_end; ++m2) {  // only active members are ever launched
                const Integ& I = h->integ[C.members[m2]];
                short base[QO_COUNT], stride[QO_COUNT];
                for (int q2 = 0; q2 < QO_COUNT; ++q2) { base[q2] = -1; stride[q2] = 0; }
                std::vector<Run> runs[3];
                std::vector<MapEnt> mf;
                int* d0 = &qdst[(size_t)m2 * QO_COUNT];
                for (int i = 0; i < I.dim; ++i) mf.push_back({(long long)I.row_off + i, QO_R, i, 0});
                int cursor = 0;
                std::string why;
                const bool par = c.rs3 != 0;
                bool ok = place_array(mf, (long long)1 << 60, cursor, base, stride, runs[0], why, d0, par, big_ok);
                if (par) cursor += 2;  // (the image of every array may be shifted by one double on its own)
                ok = ok && place_array(mj[m2], (long long)1 << 60, cursor, base, stride, runs[1], why, d0, par, big_ok);
                if (par) cursor = (cursor + 3) & ~1;
                ok = ok && place_array(mh[m2], h->nnzH, cursor, base, stride, runs[2], why, d0, par, big_ok);
                if (!ok) return fail(h, QCK_EINVAL, "unsupported trajectory layout: %s", why.c_str());
                for (int arr = 0; arr < 3; ++arr)
                    for (auto& r : runs[arr]) {
                        if (arr == 2 && r.dst >= h->nnzH) continue;  // partial column: the reduce kernel writes the shared position
                        if (r.period < r.len) h->own[arr].push_back({r.dst, 0, r.period, r.len / r.period});
                        else h->own[arr].push_back({r.dst, 0, r.len, 1});
                    }
                if (c.rs3) {
                    per_member[m2] = rs3_units(runs, h->nnzH, hdrs[m2].data(), why);
                    if (per_member[m2].empty()) return fail(h, QCK_EINVAL, "rs3 unit table: %s", why.c_str());
                } else {
                    per_member[m2] = balance_units(runs, nwarps, hdrs[m2].data());
                }
                if (first) {
                    memcpy(c.pl_base, base, sizeof base); memcpy(c.pl_stride, stride, sizeof stride);
                    c.img_doubles = cursor; c.nseg = (int)per_member[m2].size();
                    first = false;
                }
                bool same = !memcmp(c.pl_base, base, sizeof base) && !memcmp(c.pl_stride, stride, sizeof stride) && c.img_doubles == cursor &&
                            (int)per_member[m2].size() == c.nseg && hdrs[m2] == hdrs[C.member_begin];
                if (same && m2 > C.member_begin)
                    for (size_t u = 0; u < per_member[m2].size(); ++u) {
                        const QckSeg &a = per_member[m2][u], &b = per_member[C.member_begin][u];
                        same = same && a.n == b.n && a.img_nrep == b.img_nrep && a.arr == b.arr;  // arr includes the wrap step
                    }
                if (!same)
                    return fail(h, QCK_EINVAL, "integrators of one kind must see the same relative component order (state/control/timestep layout differs between members)");
            }
            const int rec = QCK_SEG_HDR / 4 + c.nseg;  // records of 16 bytes per member
            segs.assign((size_t)nm * rec, QckSeg{0, 0, 0, 0});
            for (int m2 = C.member_begin; m2 < C.member_end; ++m2) {
                memcpy(&segs[(size_t)m2 * rec], hdrs[m2].data(), QCK_SEG_HDR * sizeof(int));
                for (int u = 0; u < c.nseg; ++u) segs[(size_t)m2 * rec + QCK_SEG_HDR / 4 + u] = per_member[m2][u];
            }
        }
        // constants: A0 = -i H_drift, A_j = -i H_j dense, ELL forms of A_j and A_j^H, sparse anticommutators {A_i, A_j}
        const int N = c.N, nd = c.nd, npair = nd * (nd + 1) / 2;
        int W = 1;
        for (int q : C.members) {
            const Integ& I = h->integ[q];
            for (int j = 0; j < nd; ++j)
                for (int r = 0; r < N; ++r) {
                    int cnt = 0, cnta = 0;
                    for (int k = 0; k < N; ++k) {
                        if (I.Hdrives[(size_t)j * N * N + r + (size_t)N * k] != 0.0) ++cnt;   // A_j[r,k]
                        if (I.Hdrives[(size_t)j * N * N + k + (size_t)N * r] != 0.0) ++cnta;  // A_j^H[r,k] = conj(A_j[k,r])
                    }
                    W = std::max(W, std::max(cnt, cnta));
                }
        }
        c.W = W;
        c.ell_stride = nd * 2 * N * W;
        // {A_i, A_j} = -(H_i H_j + H_j H_i) for A = -iH, pairs ordered by (j, i <= j)
        std::vector<std::vector<std::vector<std::pair<int, std::complex<double>>>>> kk(nm);
        int kk_cap = 0;
        for (int m2 = 0; m2 < nm; ++m2) {
            const Integ& I = h->integ[C.members[m2]];
            kk[m2].resize(npair);
            int tot = 0;
            for (int j = 0, pidx = 0; j < nd; ++j)
                for (int i = 0; i <= j; ++i, ++pidx) {
                    const std::complex<double>* Hi = I.Hdrives.data() + (size_t)i * N * N;
                    const std::complex<double>* Hj = I.Hdrives.data() + (size_t)j * N * N;
                    for (int r = 0; r < N; ++r)
                        for (int k = 0; k < N; ++k) {
                            std::complex<double> v = 0.0;
                            for (int u = 0; u < N; ++u) v += Hi[r + (size_t)N * u] * Hj[u + (size_t)N * k] + Hj[r + (size_t)N * u] * Hi[u + (size_t)N * k];
                            if (v != 0.0) { kk[m2][pidx].push_back({(r << 8) | k, -v}); ++tot; }
                        }
                }
            kk_cap = std::max(kk_cap, tot);
        }
        if (N > 255) return fail(h, QCK_EINVAL, "levels > 255 unsupported");
        c.kk_cap = kk_cap;
        int ac_cap = 0;
        for (int q : C.members) {
            const Integ& I = h->integ[q];
            int tot = 0;
            for (size_t e2 = 0; e2 < I.Hdrives.size(); ++e2) tot += I.Hdrives[e2] != 0.0;
            ac_cap = std::max(ac_cap, tot);
        }
        c.ac_cap = ac_cap;
        qck_smem_finalize(c);
        if (c.rs3 >= 5) {  // three-warp variant: as many knots per CTA as the shared memory holds
            const int hoff = qck_rs3_hoff(c);
            while (c.rs3 >= 5 && qck_rs3_smem(c, hoff, c.rs3) > 227 * 1024) --c.rs3;
            if (c.rs3 < 5) c.rs3 = 1;
        }
        c.big = big_ok && !c.rs3 && (big_knob || (size_t)c.sm_bytes > 227 * 1024 - 4096 || c.img_doubles >= 32000) ? 1 : 0;
        if (!c.rs3 && !c.big && !c.genexp && (size_t)c.sm_bytes > 227 * 1024 - 4096)
            return fail(h, QCK_EINVAL, "levels=%d with %d drives needs %d bytes of shared memory per knot; this build supports at most %d", c.N, c.nd, c.sm_bytes, 227 * 1024 - 4096);
        c.cmat_stride = N * N + c.ell_stride + kk_cap + ac_cap;
        std::vector<double2> cmat((size_t)nm * c.cmat_stride, make_double2(0.0, 0.0));
        std::vector<int> icon((size_t)nm * c.icon_stride, 0);
        std::vector<int> moff(3 * (size_t)nm);
        for (int m2 = 0; m2 < nm; ++m2) {
            const Integ& I = h->integ[C.members[m2]];
            moff[3 * m2] = I.state_off; moff[3 * m2 + 1] = I.ctrl_off; moff[3 * m2 + 2] = I.row_off;
            double2* base = cmat.data() + (size_t)m2 * c.cmat_stride;
            auto minus_i = [](std::complex<double> z) { return make_double2(z.imag(), -z.real()); };
            for (int e = 0; e < N * N; ++e) base[e] = minus_i(I.Hdrift[e]);
            double2* ev = base + N * N;
            int* ec = icon.data() + (size_t)m2 * c.icon_stride;
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
            double2* kv = ev + c.ell_stride;
            int* kptr = ec + c.ell_stride;
            int* krc = kptr + npair + 1;
            int u = 0;
            for (int pidx = 0; pidx < npair; ++pidx) {
                kptr[pidx] = u;
                for (auto& t : kk[m2][pidx]) { krc[u] = t.first; kv[u] = make_double2(t.second.real(), t.second.imag()); ++u; }
            }
            kptr[npair] = u;
            // per-element contributor lists: A[e] = A0[e] + sum over (j, -i H_j[e]) with H_j[e] != 0
            double2* av = kv + kk_cap;
            int* aptr = krc + kk_cap;
            int* aj = aptr + N * N + 1;
            int au = 0;
            for (int e2 = 0; e2 < N * N; ++e2) {
                aptr[e2] = au;
                for (int j = 0; j < nd; ++j) {
                    const std::complex<double> hv = I.Hdrives[(size_t)j * N * N + e2];
                    if (hv != 0.0) { av[au] = minus_i(hv); aj[au] = j; ++au; }
                }
            }
            aptr[N * N] = au;
        }
        // column kernel (levels <= 4): dense drive matrices A_j = -i H_j, row-major, and the per-member destinations
        c.dense_aj = nullptr; c.qdst = nullptr;
        if (((c.kind == QCK_UNITARY_PADE || c.kind == QCK_KET_PADE) && c.order == 4 && (N <= 4 || c.big)) || c.colexp || c.genexp) {
            std::vector<double2> daj((size_t)nm * nd * N * N);
            for (int m2 = 0; m2 < nm; ++m2) {
                const Integ& I = h->integ[C.members[m2]];
                for (int j = 0; j < nd; ++j)
                    for (int r = 0; r < N; ++r)
                        for (int k = 0; k < N; ++k) {
                            const std::complex<double> a = std::complex<double>(0, -1) * I.Hdrives[(size_t)j * N * N + r + (size_t)N * k];
                            daj[((size_t)(m2 * nd + j) * N + r) * N + k] = make_double2(a.real(), a.imag());
                        }
            }
            cudaError_t e2;
            if ((e2 = upload(h, daj, &c.dense_aj, C.allocs)) != cudaSuccess || (e2 = upload(h, qdst, &c.qdst, C.allocs)) != cudaSuccess)
                return fail(h, QCK_ECUDA, "uploading column-kernel tables: %s", cudaGetErrorString(e2));
        }
        cudaError_t e;
        if ((e = upload(h, segs, &c.segs, C.allocs)) != cudaSuccess ||
            (e = upload(h, cmat, &c.cmat, C.allocs)) != cudaSuccess || (e = upload(h, icon, &c.ell_col, C.allocs)) != cudaSuccess ||
            (e = upload(h, moff, &c.moff, C.allocs)) != cudaSuccess)
            return fail(h, QCK_ECUDA, "uploading class constants: %s", cudaGetErrorString(e));
        c.tape = nullptr; c.tape_stride = 0; c.tape_levels = 0; c.max_ctas = 0;
        if ((c.kind == QCK_UNITARY_PADE || c.kind == QCK_KET_PADE) && c.order != 4 && h->eval_hessian && h->device >= 0) {
            // reverse-sweep tape of the general-order Pade Hessian: (m-1) Horner levels x (P + nd + 1 tangents) per group
            c.tape_stride = (long long)(c.pade_m - 1) * (2 + nd) * N * N;
            c.max_ctas = h->sm_count * 4;
            void* tp = nullptr;
            // (QCK_TAPE_SLOTS copies: launches on the handle's own stream and on the two pipeline streams may overlap)
            if ((e = cudaMalloc(&tp, sizeof(double2) * (size_t)c.tape_stride * c.max_ctas * QCK_TAPE_SLOTS)) != cudaSuccess)
                return fail(h, QCK_ENOMEM, "tape allocation failed: %s", cudaGetErrorString(e));
            C.allocs.push_back(tp);
            c.tape = static_cast<double2*>(tp);
        }
        if ((c.kind == QCK_UNITARY_EXP || c.kind == QCK_KET_EXP) && h->eval_hessian && h->device >= 0 && !c.eig && !c.colexp && !c.genexp) {  // (the spectral kernels have no tape)
            // reverse-sweep tape of the exponential Hessian: 7 Horner steps x nd jets + 16 squaring levels x (1 + nd) matrices per CTA
            c.tape_levels = 16;
            c.tape_stride = (long long)(7 * nd + c.tape_levels * (1 + nd)) * N * N;
            c.max_ctas = h->sm_count * 4;
            void* tp = nullptr;
            // (QCK_TAPE_SLOTS copies: launches on the handle's own stream and on the two pipeline streams may overlap)
            if ((e = cudaMalloc(&tp, sizeof(double2) * (size_t)c.tape_stride * c.max_ctas * QCK_TAPE_SLOTS)) != cudaSuccess)
                return fail(h, QCK_ENOMEM, "tape allocation failed: %s", cudaGetErrorString(e));
            C.allocs.push_back(tp);
            c.tape = static_cast<double2*>(tp);
        }
        if (c.colexp && h->device >= 0) {
            // spectral column kernels: eigenvectors + eigenvalues of every (knot, active member) item, one region per stream slot
            c.tape_levels = 0;
            c.tape_stride = (long long)qck_colexp_scratch_rec(N);
            c.max_ctas = (int)std::max<long long>(1, (h->T - 1) * (long long)std::max(1, C.member_end - C.member_begin));
            void* tp = nullptr;
            if ((e = cudaMalloc(&tp, sizeof(double2) * (size_t)c.tape_stride * c.max_ctas * QCK_TAPE_SLOTS)) != cudaSuccess)
                return fail(h, QCK_ENOMEM, "eigen scratch allocation failed: %s", cudaGetErrorString(e));
            C.allocs.push_back(tp);
            c.tape = static_cast<double2*>(tp);
        }
        if (c.eig && h->device >= 0) {
            // spectral kernel: eigenvectors + eigenvalues of every knot (86 double2 each), written by the eigen kernel and read by the
            // main kernel of the same launch; one region per stream slot like the tapes (slot offset = tape_stride * max_ctas)
            c.tape_levels = 0;
            c.tape_stride = 86;
            c.max_ctas = (int)std::max<long long>(1, h->T - 1);
            void* tp = nullptr;
            if ((e = cudaMalloc(&tp, sizeof(double2) * (size_t)c.tape_stride * c.max_ctas * QCK_TAPE_SLOTS)) != cudaSuccess)
                return fail(h, QCK_ENOMEM, "eigen scratch allocation failed: %s", cudaGetErrorString(e));
            C.allocs.push_back(tp);
            c.tape = static_cast<double2*>(tp);
        }
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
        if (HE[k].cls < 0 && hdst[k] >= 0 && HE[k].contrib >= h->ib && HE[k].contrib < h->ie)
            h->aux.push_back({2, HE[k].aux_op, (int32_t)hdst[k], HE[k].aux_i0, 0, 0, 0.0});
    // ---- positions this handle writes (host-buffer path: what crosses PCIe, and how it expands into the caller's arrays) ------
    for (auto& a : h->aux)
        if (a.out != 2 || a.pos < h->nnzH) h->own[a.out].push_back({a.pos, 0, 1, 1});
    if (!h->exclude_shared)
        for (size_t i = 0; i < h->sh_pos.size(); ++i)
            if (h->sh_ptr[i + 1] > h->sh_ptr[i]) h->own[2].push_back({h->sh_pos[i], 0, 1, 1});
    for (int arr = 0; arr < 3; ++arr) {
        auto& v = h->own[arr];
        std::sort(v.begin(), v.end(), [](const QckOwnSeg& a, const QckOwnSeg& b) { return a.full < b.full; });
        std::vector<QckOwnSeg> merged;
        for (auto& sgm : v) {
            if (!merged.empty()) {
                QckOwnSeg& b = merged.back();
                const int bend = b.full + b.len * b.nrep;
                if (sgm.full < bend) return fail(h, QCK_EINVAL, "internal: overlapping output runs");
                if (sgm.full == bend && b.nrep == 1 && sgm.nrep == 1) { b.len += sgm.len; continue; }
            }
            merged.push_back(sgm);
        }
        int comp = 0;
        for (auto& sgm : merged) { sgm.comp = comp; comp += sgm.len; }
        v.swap(merged);
    }
    cudaError_t e;
    if ((e = upload(h, h->aux, &h->d_aux, h->allocs)) != cudaSuccess) return fail(h, QCK_ECUDA, "uploading aux entries: %s", cudaGetErrorString(e));
    h->red.n_shared = (int)h->sh_pos.size();
    if ((e = upload(h, h->sh_pos, &h->red.pos, h->allocs)) != cudaSuccess || (e = upload(h, h->sh_ptr, &h->red.ptr, h->allocs)) != cudaSuccess ||
        (e = upload(h, h->sh_cols, &h->red.cols, h->allocs)) != cudaSuccess)
        return fail(h, QCK_ECUDA, "uploading reduction tables: %s", cudaGetErrorString(e));
    return QCK_OK;
}

}  // namespace

int qck_run(qck_handle* h, uint32_t mask, long long k0, long long nk, const double* dZ, const double* dmu, double* dF, double* dJ,
            double* dH, cudaStream_t st, int slot) {
    if (!h->eval_hessian) mask &= ~QCK_EVAL_H;
    if (!dF) mask &= ~QCK_EVAL_F;
    if (!dJ) mask &= ~QCK_EVAL_J;
    if (!dH) mask &= ~QCK_EVAL_H;
    if (!mask || nk <= 0) return QCK_OK;
    if ((mask & QCK_EVAL_H) && !dmu) return fail(h, QCK_EINVAL, "the Hessian needs the multipliers mu");
    QckLaunch L{};
    L.Z = dZ + k0 * h->zdim;
    L.mu = dmu ? dmu + k0 * h->dyn : nullptr;
    L.F = dF ? dF + k0 * h->dyn : nullptr;
    L.J = dJ ? dJ + k0 * h->nnzJ : nullptr;
    L.H = dH ? dH + k0 * h->nnzH : nullptr;
    L.partial = h->dpartial + k0 * h->npart;
    L.n_knots = nk; L.nnzJ = h->nnzJ; L.nnzH = h->nnzH; L.npart = h->npart; L.mask = mask;
    L.status = h->d_status;
    bool aux_done = h->aux.empty();
    const bool fuse_aux = (int)h->aux.size() <= qck_fused_aux_limit();
    int launches = 0;
    for (auto& C : h->classes) {
        if (C.member_end <= C.member_begin) continue;
        L.c = C.dev; L.member_begin = C.member_begin; L.member_end = C.member_end;
        if (L.c.tape) L.c.tape += (size_t)slot * (size_t)L.c.tape_stride * (size_t)L.c.max_ctas;  // this stream's scratch tape
        L.moff_global = C.dev.moff ? C.dev.moff + 3 * C.member_begin : nullptr;
        L.plan = &C.plan;
        const bool take_aux = !aux_done && fuse_aux;
        L.aux = take_aux ? h->d_aux : nullptr; L.n_aux = take_aux ? (int)h->aux.size() : 0;
        if (take_aux) aux_done = true;
        // one class with one member and no partial columns: whatever the launch does not write itself is written after it (the
        // stand-alone aux kernel), so a kernel may store whole knot blocks (column kernel: block-staged write-out)
        L.sole_writer = h->classes.size() == 1 && C.dev.n_members == 1 && h->npart == 0 && (take_aux || h->aux.empty() || !fuse_aux) ? 1 : 0;
        int rc = qck_launch_quantum(L, h->sm_count, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "quantum kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    L.plan = nullptr;
    if (!aux_done) {
        L.c = QckClassDev{};
        L.c.free_time = h->dt_off >= 0; L.c.dt_off = h->dt_off; L.c.zdim = h->zdim; L.c.dyn = h->dyn; L.c.dt_fixed = h->dt_fixed;
        L.aux = h->d_aux; L.n_aux = (int)h->aux.size();
        int rc = qck_launch_aux(L, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "aux kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    if ((mask & QCK_EVAL_H) && h->red.n_shared && h->npart > 0 && !h->skip_local_reduce) {
        int rc = qck_launch_reduce(h->red, L.H, L.partial, nk, h->nnzH, h->npart, st, &launches);
        if (rc) return fail(h, QCK_ECUDA, "reduce kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    }
    h->launches += launches;
    return QCK_OK;
}

int qck_check_status(qck_handle* h) {
    if (!h->d_status) return QCK_OK;
    int st = 0;
    QCK_CUDA_TRY(h, cudaMemcpy(&st, h->d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (!st) return QCK_OK;
    cudaMemset(h->d_status, 0, sizeof st);
    if (st & QCK_ST_EXP_RANGE)
        return fail(h, QCK_ERANGE, "exponential integrator: ||dt*G(a)||_1 exceeds 4096 at some knot (more squaring levels than the Hessian tape holds); results of this call are not exp(dt*G)");
    return fail(h, QCK_ERANGE, "device reported status 0x%x", st);
}

// canonical h->dJ / h->dH (+ the partial columns) -> caller-order arrays, on stream st
int qck_reorder(qck_handle* h, uint32_t mask, double* dJ_out, double* dH_out, cudaStream_t st) {
    const QckPublicOrder& P = h->pub;
    const long long nk = h->T - 1;
    int launches = 0, rc = 0;
    if ((mask & QCK_EVAL_J) && dJ_out) rc = qck_launch_reorder(h->dJ, nullptr, dJ_out, P.d_srcJ, (int)P.nnzJ, h->nnzJ, 0, nk, st, &launches);
    if (!rc && (mask & QCK_EVAL_H) && dH_out && h->eval_hessian)
        rc = qck_launch_reorder(h->dH, h->dpartial, dH_out, P.d_srcH, (int)P.nnzH, h->nnzH, h->npart, nk, st, &launches);
    h->launches += launches;
    return rc ? fail(h, QCK_ECUDA, "reorder kernel launch: %s", cudaGetErrorString((cudaError_t)rc)) : QCK_OK;
}

namespace {

// Host-buffer path under a structure-order policy other than CSC: one pass into the handle's canonical device arrays, a gather
// into caller order on the device, plain copies out (no compact transport: the kron blocks are not contiguous in these orders).
int eval_host_reordered(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) {
    const long long nk = h->T - 1;
    CUDA_TRY(h, cudaSetDevice(h->device));
    h->pipe.valid_mask = 0; h->pipe.z_staged = false; h->pipe.z_on_device = false; h->pipe.mu_on_device = false;
    if (!h->eval_hessian) H = nullptr;
    uint32_t mask = (F ? QCK_EVAL_F : 0u) | (J ? QCK_EVAL_J : 0u) | (H ? QCK_EVAL_H : 0u);
    if (!mask) return QCK_OK;
    cudaStream_t st = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->dZ, Z, sizeof(double) * h->T * h->zdim, cudaMemcpyHostToDevice, st));
    if (H) CUDA_TRY(h, cudaMemcpyAsync(h->dmu, mu, sizeof(double) * nk * h->dyn, cudaMemcpyHostToDevice, st));
    int rc = qck_run(h, mask, 0, nk, h->dZ, h->dmu, F ? h->dF : nullptr, J ? h->dJ : nullptr, H ? h->dH : nullptr, st, 0);
    if (rc) return rc;
    if ((rc = qck_reorder(h, mask, h->pub.dJ, h->pub.dH, st))) return rc;
    if (F) CUDA_TRY(h, cudaMemcpyAsync(F, h->dF, sizeof(double) * nk * h->dyn, cudaMemcpyDeviceToHost, st));
    if (J) CUDA_TRY(h, cudaMemcpyAsync(J, h->pub.dJ, sizeof(double) * nk * h->pub.nnzJ, cudaMemcpyDeviceToHost, st));
    if (H) CUDA_TRY(h, cudaMemcpyAsync(H, h->pub.dH, sizeof(double) * nk * h->pub.nnzH, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    h->pipe.h2d_bytes = (long long)sizeof(double) * (h->T * h->zdim + (H ? nk * h->dyn : 0));
    h->pipe.d2h_bytes = (long long)sizeof(double) * ((F ? nk * h->dyn : 0) + (J ? nk * h->pub.nnzJ : 0) + (H ? nk * h->pub.nnzH : 0));
    return h->uses_status ? qck_check_status(h) : QCK_OK;
}

int eval_host(qck_handle* h, const double* Z, const double* mu, double* F, double* J, double* H) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (!Z) return fail(h, QCK_EINVAL, "Z is NULL");
    if (!h->children.empty()) return qck_multi_eval(h, Z, mu, F, J, H);
    if (h->device < 0) return fail(h, QCK_ENODEVICE, "structure-only handle (device=-1): libqcknot has no CPU evaluation path");
    if (H && h->eval_hessian && !mu) return fail(h, QCK_EINVAL, "the Hessian needs the multipliers mu");
    if (h->pub.order != QCK_ORDER_CSC) return eval_host_reordered(h, Z, mu, F, J, H);
    return qck_pipe_eval(h, Z, mu, F, J, H);
}

}  // namespace

extern "C" {

const char* qck_version(void) { return "qcknot 0.2 (sm_100a)"; }

const char* qck_last_error(const qck_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

// One GPU (or structure-only): parses the description, builds structures and device tables, allocates the value buffers.
int qck_create_single(const qck_problem_desc* d, qck_handle** out, bool exclude_shared) {
    *out = nullptr;
    const bool structure_only = d->device < 0;  // sizes + structures only (what Ipopt asks for at set-up); every eval fails
    int ndev = 0;
    cudaError_t ce = cudaSuccess;
    cudaDeviceProp prop{};
    prop.multiProcessorCount = 148;
    if (!structure_only) {
        ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0)
            return fail(nullptr, QCK_ENODEVICE, "no CUDA device (%s); libqcknot has no CPU fallback", ce == cudaSuccess ? "count=0" : cudaGetErrorString(ce));
        if (d->device >= ndev) return fail(nullptr, QCK_ENODEVICE, "device %d out of range (%d devices)", d->device, ndev);
        if ((ce = cudaGetDeviceProperties(&prop, d->device)) != cudaSuccess) return fail(nullptr, QCK_ECUDA, "%s", cudaGetErrorString(ce));
        if (prop.major != 10) return fail(nullptr, QCK_ENODEVICE, "device %d is sm_%d%d; libqcknot is built for sm_100a only", d->device, prop.major, prop.minor);
    }

    qck_handle* h = new (std::nothrow) qck_handle();
    if (!h) return fail(nullptr, QCK_ENOMEM, "out of host memory");
    auto bail = [&](int code) { g_create_error = h->err; qck_destroy(h); return code; };
    h->device = structure_only ? -1 : d->device; h->sm_count = prop.multiProcessorCount;
    h->T = d->T; h->zdim = d->zdim; h->dt_off = d->dt_off < 0 ? -1 : d->dt_off; h->dt_fixed = d->dt_fixed;
    h->eval_hessian = d->eval_hessian ? 1 : 0;
    h->exclude_shared = exclude_shared;
    h->host_threads = d->host_threads;
    h->pub.order = d->structure_order;
    if (h->pub.order < QCK_ORDER_CSC || h->pub.order > QCK_ORDER_PER_INTEGRATOR) { fail(h, QCK_EINVAL, "unknown structure_order %d", d->structure_order); return bail(QCK_EINVAL); }
    h->ib = d->integ_begin; h->ie = d->integ_end;
    if (h->ie < 0) { h->ie = d->n_integrators; }  // integ_end < 0: every integrator; begin == end: none (an empty shard launches nothing)
    if (h->ib < 0 || h->ie > d->n_integrators || h->ib > h->ie) { fail(h, QCK_EINVAL, "bad integrator range [%d,%d)", h->ib, h->ie); return bail(QCK_EINVAL); }
    int row = 0;
    for (int q = 0; q < d->n_integrators; ++q) {
        const qck_integrator_desc& s = d->integrators[q];
        Integ I;
        I.kind = s.kind; I.order = s.order; I.N = s.levels; I.nd = s.n_drives;
        I.state_off = s.state_off; I.state_len = s.state_len; I.ctrl_off = s.ctrl_off;
        if (I.kind < 0 || I.kind > QCK_DERIVATIVE) { fail(h, QCK_EINVAL, "integrator %d: unknown kind %d", q, I.kind); return bail(QCK_EINVAL); }
        if (I.quantum()) {
            if (I.N < 1 || I.nd < 1 || I.nd > QCK_MAX_DRIVES) { fail(h, QCK_EINVAL, "integrator %d: levels=%d n_drives=%d unsupported (max %d drives)", q, I.N, I.nd, QCK_MAX_DRIVES); return bail(QCK_EINVAL); }
            I.nc = I.unitary() ? I.N : 1;
            I.dim = 2 * I.N * I.nc;
            if (I.state_len != I.dim) { fail(h, QCK_EINVAL, "integrator %d: state_len %d != %d", q, I.state_len, I.dim); return bail(QCK_EINVAL); }
            if (I.pade() && (I.order < 4 || I.order > 12 || (I.order & 1))) { fail(h, QCK_EINVAL, "integrator %d: Pade order %d not supported (even orders 4..12)", q, I.order); return bail(QCK_EINVAL); }
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
    int rc = build(h);
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
    {
        void* p = nullptr;
        if ((ce = cudaMalloc(&p, 16)) != cudaSuccess) { fail(h, QCK_ENOMEM, "%s", cudaGetErrorString(ce)); return bail(QCK_ENOMEM); }
        h->allocs.push_back(p);
        cudaMemset(p, 0, 16);
        h->d_status = static_cast<int*>(p);
    }
    if (h->pub.order != QCK_ORDER_CSC) {  // caller-order arrays + gather tables of the structure-order policy
        QckPublicOrder& P = h->pub;
        if ((ce = upload(h, P.srcJ, &P.d_srcJ, h->allocs)) != cudaSuccess || (ce = upload(h, P.srcH, &P.d_srcH, h->allocs)) != cudaSuccess) {
            fail(h, QCK_ECUDA, "uploading the structure-order tables: %s", cudaGetErrorString(ce)); return bail(QCK_ECUDA);
        }
        struct { double** p; long long n; } pb[] = {{&P.dJ, nk * std::max<long long>(P.nnzJ, 1)}, {&P.dH, nk * std::max<long long>(P.nnzH, 1)}};
        for (auto& b : pb) {
            void* q = nullptr;
            if ((ce = cudaMalloc(&q, sizeof(double) * (size_t)b.n)) != cudaSuccess) { fail(h, QCK_ENOMEM, "device allocation of %lld doubles failed: %s", b.n, cudaGetErrorString(ce)); return bail(QCK_ENOMEM); }
            h->allocs.push_back(q);
            *b.p = static_cast<double*>(q);
        }
    }
    for (auto& C : h->classes) h->uses_status = h->uses_status || (C.dev.tape != nullptr && C.dev.tape_levels > 0);
    if ((ce = cudaDeviceSynchronize()) != cudaSuccess) { fail(h, QCK_ECUDA, "%s", cudaGetErrorString(ce)); return bail(QCK_ECUDA); }
    *out = h;
    return QCK_OK;
}

extern "C" {

int qck_create(const qck_problem_desc* d, qck_handle** out) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!out) return fail(nullptr, QCK_EINVAL, "out is NULL");
    *out = nullptr;
    if (!d || !d->integrators || d->n_integrators <= 0) return fail(nullptr, QCK_EINVAL, "empty problem description");
    if (d->T < 2) return fail(nullptr, QCK_EINVAL, "T must be >= 2 (got %lld)", (long long)d->T);
    if (d->zdim <= 0 || d->dt_off >= d->zdim) return fail(nullptr, QCK_EINVAL, "bad zdim/dt_off");
    if (d->n_gpus < 0 || d->n_gpus > 64) return fail(nullptr, QCK_EINVAL, "bad n_gpus %d", d->n_gpus);
    if (d->n_gpus <= 1) {
        qck_problem_desc d1 = *d;
        if (d->devices && d->n_gpus == 1 && d->device >= 0) d1.device = d->devices[0];
        return qck_create_single(&d1, out, false);
    }
    if (d->shard_mode != QCK_SHARD_KNOT && d->shard_mode != QCK_SHARD_ENSEMBLE) return fail(nullptr, QCK_EINVAL, "unknown shard_mode %d", d->shard_mode);
    if (d->structure_order != QCK_ORDER_CSC) return fail(nullptr, QCK_EINVAL, "structure_order other than CSC is served on single-GPU handles only (n_gpus = %d)", d->n_gpus);
    // the parent holds sizes + structures of the whole problem; one child per GPU does the work
    qck_problem_desc dp = *d;
    dp.device = -1; dp.n_gpus = 1; dp.integ_begin = 0; dp.integ_end = -1;
    qck_handle* parent = nullptr;
    int rc = qck_create_single(&dp, &parent, false);
    if (rc) return rc;
    rc = qck_multi_create(d, parent);
    if (rc) { g_create_error = parent->err; qck_destroy(parent); return rc; }
    *out = parent;
    return QCK_OK;
}

void qck_destroy(qck_handle* h) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return;
    if (!h->children.empty() || h->nccl) qck_multi_destroy(h);
    qck_objective_free(h);
    if (h->device < 0) { delete h; return; }
    cudaSetDevice(h->device);
    qck_pipe_destroy(h);
    for (auto& C : h->classes) for (void* p : C.allocs) cudaFree(p);
    for (void* p : h->allocs) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int qck_sizes(const qck_handle* h, int64_t* dyn, int64_t* nnzJ, int64_t* nnzH) {
    if (!h) return QCK_EINVAL;
    const bool re = h->pub.order != QCK_ORDER_CSC;
    if (dyn) *dyn = h->dyn;
    if (nnzJ) *nnzJ = re ? h->pub.nnzJ : h->nnzJ;
    if (nnzH) *nnzH = re ? h->pub.nnzH : h->nnzH;
    return QCK_OK;
}

int qck_jacobian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols) {
    if (!h || !rows || !cols) return QCK_EINVAL;
    const bool re = h->pub.order != QCK_ORDER_CSC;
    const std::vector<int32_t>&Jr = re ? h->pub.Jr : h->Jr, &Jc = re ? h->pub.Jc : h->Jc;
    const long long nnz = (long long)Jr.size();
    for (long long t = 0; t < h->T - 1; ++t)
        for (long long k = 0; k < nnz; ++k) {
            rows[t * nnz + k] = Jr[k] + (t + knot_offset) * h->dyn + 1;
            cols[t * nnz + k] = Jc[k] + (t + knot_offset) * h->zdim + 1;
        }
    return QCK_OK;
}

int qck_hessian_structure(const qck_handle* h, int64_t knot_offset, int64_t* rows, int64_t* cols) {
    if (!h || !rows || !cols) return QCK_EINVAL;
    const bool re = h->pub.order != QCK_ORDER_CSC;
    const std::vector<int32_t>&Hr = re ? h->pub.Hr : h->Hr, &Hc = re ? h->pub.Hc : h->Hc;
    const long long nnz = (long long)Hr.size();
    for (long long t = 0; t < h->T - 1; ++t)
        for (long long k = 0; k < nnz; ++k) {
            rows[t * nnz + k] = Hr[k] + (t + knot_offset) * h->zdim + 1;
            cols[t * nnz + k] = Hc[k] + (t + knot_offset) * h->zdim + 1;
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
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (!h->children.empty()) return fail(h, QCK_EINVAL, "multi-GPU handle: use qck_upload + qck_eval_resident (device pointers belong to one GPU)");
    if (!dZ) return fail(h, QCK_EINVAL, "dZ is NULL");
    if (h->device < 0) return fail(h, QCK_ENODEVICE, "structure-only handle (device=-1): libqcknot has no CPU evaluation path");
    CUDA_TRY(h, cudaSetDevice(h->device));
    uint32_t m = mask;
    if (!(m & QCK_EVAL_F)) dF = nullptr;
    if (!(m & QCK_EVAL_J)) dJ = nullptr;
    if (!(m & QCK_EVAL_H)) dH = nullptr;
    if (dZ == h->dZ || dmu == h->dmu || dF == h->dF || dJ == h->dJ || dH == h->dH || h->pub.order != QCK_ORDER_CSC) { h->pipe.valid_mask = 0; h->pipe.z_on_device = false; h->pipe.mu_on_device = false; }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    if (h->pub.order != QCK_ORDER_CSC) {  // kernels write the handle's canonical arrays; one gather pass into the caller's
        if (dJ == h->dJ || dH == h->dH) return fail(h, QCK_EINVAL, "structure_order: dJ / dH must not be the handle's canonical buffers");
        int rc = qck_run(h, m, 0, h->T - 1, dZ, dmu, dF, dJ ? h->dJ : nullptr, dH ? h->dH : nullptr, st, 0);
        return rc ? rc : qck_reorder(h, m, dJ, dH, st);
    }
    return qck_run(h, m, 0, h->T - 1, dZ, dmu, dF, dJ, dH, st, 0);
}

int qck_device_buffers(qck_handle* h, double** dZ, double** dmu, double** dF, double** dJ, double** dH) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (!h->children.empty()) return fail(h, QCK_EINVAL, "multi-GPU handle: use qck_shard_device_buffers");
    if (dZ) *dZ = h->dZ;
    if (dmu) *dmu = h->dmu;
    if (dF) *dF = h->dF;
    if (dJ) *dJ = h->pub.order != QCK_ORDER_CSC ? h->pub.dJ : h->dJ;  // (caller-order arrays under a structure-order policy)
    if (dH) *dH = h->pub.order != QCK_ORDER_CSC ? h->pub.dH : h->dH;
    h->pipe.valid_mask = 0; h->pipe.z_on_device = false; h->pipe.mu_on_device = false;  // the caller may overwrite them
    return QCK_OK;
}

int qck_synchronize(qck_handle* h) {
    QckDeviceScope device_scope;  // the caller's current CUDA device is restored on return
    if (!h) return QCK_EINVAL;
    if (!h->children.empty()) {
        for (qck_handle* c : h->children) {
            int rc = qck_synchronize(c);
            if (rc) { h->err = c->err; return rc; }
        }
        return QCK_OK;
    }
    if (h->device < 0) return QCK_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return h->uses_status ? qck_check_status(h) : QCK_OK;
}

int qck_shared_hessian_positions(const qck_handle* h, int64_t* count, int64_t* pos) {
    if (!h || !count) return QCK_EINVAL;
    const std::vector<long long>& sp = h->pub.order != QCK_ORDER_CSC ? h->pub.shared : h->shared_positions;
    *count = (int64_t)sp.size();
    if (pos) for (size_t k = 0; k < sp.size(); ++k) pos[k] = sp[k];
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
    long long n = h->launches;
    for (const qck_handle* c : h->children) n += c->launches;
    *launches = n;
    return QCK_OK;
}

int qck_transfer_stats(const qck_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* cache_hits) {
    if (!h) return QCK_EINVAL;
    long long a = h->pipe.h2d_bytes, b = h->pipe.d2h_bytes, c = h->pipe.cache_hits;
    for (const qck_handle* ch : h->children) { a += ch->pipe.h2d_bytes; b += ch->pipe.d2h_bytes; c += ch->pipe.cache_hits; }
    if (h2d_bytes) *h2d_bytes = a;
    if (d2h_bytes) *d2h_bytes = b;
    if (cache_hits) *cache_hits = c;
    return QCK_OK;
}

int qck_invalidate(qck_handle* h) {
    if (!h) return QCK_EINVAL;
    h->pipe.valid_mask = 0; h->pipe.z_staged = false; h->pipe.z_on_device = false; h->pipe.mu_on_device = false;
    for (qck_handle* c : h->children) qck_invalidate(c);
    return QCK_OK;
}

}  // extern "C"
