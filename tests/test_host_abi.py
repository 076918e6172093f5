"""CPU tests of the host logic and the C-ABI: the library loads, exports every symbol include/qcknot.h declares,
builds the reference's sparsity structures bit-exactly (structure-only handles need no GPU), and fails loudly --
never falls back -- when asked to evaluate without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import qcknot
from qcknot import workloads as wl
from oracle.bridge import oracle_dynamics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qcknot.h")).read()
    return sorted(set(re.findall(r"\b(qck_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(qcknot.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"libqcknot.so does not export {s}"
    assert sorted(qcknot._lib.EXPORTS) == syms
    assert b"sm_100a" in qcknot._lib.load().qck_version()


def test_no_reference_to_oracle_in_product():
    """The product path must never import or link the oracle."""
    pkg = os.path.join(ROOT, "quantumcollocation.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), f"{f} mentions the oracle"


@pytest.mark.parametrize("name,kw", [
    ("hadamard", {}), ("hadamard", {"free_time": False}), ("cz", {"T": 4}), ("ket", {"T": 7}),
    ("ket", {"T": 4, "free_time": False}), ("sampling", {"T": 3, "n_systems": 5}),
])
def test_structures_bit_exact_vs_oracle(name, kw):
    systems, traj, integrators = wl.config(name, **kw)
    for eval_hessian in (True, False):
        D = qcknot.QuantumDynamics(integrators, traj, eval_hessian=eval_hessian, device=-1)
        O = oracle_dynamics(integrators, traj, eval_hessian=eval_hessian)
        assert (D.dyn, D.nnzJ, D.nnzH) == (O.dyn, O.nnzJ, O.nnzH)
        assert np.array_equal(D.dF_structure, np.array(O.dF_structure, dtype=np.int64).reshape(-1, 2))
        assert np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure, dtype=np.int64).reshape(-1, 2))
        D.close()


@pytest.mark.parametrize("order", ["row_major", "per_integrator"])
@pytest.mark.parametrize("name,kw", [("hadamard", {}), ("cz", {"T": 4}), ("ket", {"T": 5, "free_time": False}),
                                     ("sampling", {"T": 3, "n_systems": 5})])
def test_structure_order_policies_vs_oracle(name, kw, order):
    """qck_problem_desc.structure_order: the caller-visible structures of the non-default policies, bit-exact against the oracle's
    restatement of the same policy; every policy describes the same sparse matrices (the Core's own order is [DEP-RECALL])."""
    systems, traj, integrators = wl.config(name, **kw)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1, structure_order=order)
    O = oracle_dynamics(integrators, traj, structure_order=order)
    C0 = qcknot.QuantumDynamics(integrators, traj, device=-1)
    assert (D.dyn, D.nnzJ, D.nnzH) == (O.dyn, O.nnzJ, O.nnzH)
    assert np.array_equal(D.dF_structure, np.array(O.dF_structure, dtype=np.int64).reshape(-1, 2))
    assert np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure, dtype=np.int64).reshape(-1, 2))
    assert sorted(map(tuple, D.dF_structure)) == sorted(map(tuple, C0.dF_structure))
    assert set(map(tuple, D.mu_d2F_structure)) == set(map(tuple, C0.mu_d2F_structure))
    if order == "row_major":
        assert D.nnzH == C0.nnzH
    else:  # duplicates exactly where several integrators share a Hessian position (shared controls / timestep)
        assert D.nnzH >= C0.nnzH and (D.nnzH > C0.nnzH) == (len(C0.shared_hessian_positions()) > 0)
    for x in (D, C0):
        x.close()


def test_structure_order_rejected_on_multi_gpu_and_unknown():
    systems, traj, integrators = wl.config("hadamard", T=9)
    with pytest.raises(ValueError):
        qcknot.QuantumDynamics(integrators, traj, device=-1, structure_order="diagonal")
    with pytest.raises(qcknot.QcknotError, match="single-GPU"):
        qcknot.QuantumDynamics(integrators, traj, device=0, n_gpus=2, structure_order="row_major")


def test_survey_size_table():
    # SURVEY.md section 8 size table (pade column): C1 104/58, C2 6674/1643 (zdim 175, dyn 170), C5 (S=256) 155,664/49,162
    for name, kw, want in (("hadamard", {}, (12, 104, 58, 15)), ("cz", {"T": 3}, (170, 6674, 1643, 175)),
                           ("sampling", {"T": 3, "n_systems": 256}, (8196, 155664, 49162, 8199))):
        systems, traj, integrators = wl.config(name, **kw)
        D = qcknot.QuantumDynamics(integrators, traj, device=-1)
        assert (D.dyn, D.nnzJ, D.nnzH, D.zdim) == want
        D.close()


def test_knot_shard_structures_concatenate_to_global():
    systems, traj, integrators = wl.config("hadamard", T=11)
    full = qcknot.QuantumDynamics(integrators, traj, device=-1)
    Js, Hs = [], []
    for t0, t1 in qcknot.sharding.knot_shards(traj.T - 1, 3):
        S = qcknot.QuantumDynamics(integrators, traj, device=-1, knot_range=(t0, t1))
        assert S.n_blocks == t1 - t0
        Js.append(S.dF_structure)
        Hs.append(S.mu_d2F_structure)
    assert np.array_equal(np.concatenate(Js), full.dF_structure)
    assert np.array_equal(np.concatenate(Hs), full.mu_d2F_structure)


def test_shared_hessian_positions_of_sampling_problem():
    systems, traj, integrators = wl.config("sampling", T=3, n_systems=4)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1)
    pos = D.shared_hessian_positions()
    nd = systems[0].n_drives
    assert len(pos) == nd * (nd + 1) // 2 + nd + 1  # a x a, a x dt, dt x dt
    s = D.mu_d2F_structure[pos]  # first knot block
    a = traj.components["a"]
    dt = traj.components["Δt"].start
    for r, c in s - 1:
        assert (r in a or r == dt) and (c in a or c == dt)


def test_evaluation_without_device_fails_loudly():
    systems, traj, integrators = wl.config("hadamard", T=5)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1)
    with pytest.raises(qcknot.QcknotError, match="no CPU evaluation path"):
        D.F(traj.datavec)
    with pytest.raises(qcknot.QcknotError, match="no CPU evaluation path"):
        D.dF(traj.datavec)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(qcknot.QcknotError, match="no CPU fallback"):
            qcknot.QuantumDynamics(integrators, traj, device=0)


def test_bad_arguments_are_errors_not_crashes():
    systems, traj, integrators = wl.config("hadamard", T=5)
    with pytest.raises(ValueError):
        qcknot.QuantumDynamics(integrators, traj, device=-1, knot_range=(3, 3))
    with pytest.raises(KeyError):
        qcknot.UnitaryPadeIntegrator("nope", "a", systems[0], traj)
    with pytest.raises(ValueError):  # state length inconsistent with the system's levels
        qcknot.UnitaryPadeIntegrator("a", "a", systems[0], traj)
    bad = qcknot.UnitaryPadeIntegrator("Ũ⃗", "a", systems[0], traj, order=14)
    with pytest.raises(qcknot.QcknotError, match="order"):
        qcknot.QuantumDynamics([bad], traj, device=-1)
    # raw ABI: NULL description
    lib = qcknot._lib.load()
    h = C.c_void_p()
    assert lib.qck_create(None, C.byref(h)) != 0 and not h.value
    assert b"empty" in lib.qck_last_error(None)


def test_trajectory_layout_matches_reference_fixture():
    # component order [state, a, da, dda, dt] and datavec = vec(data) column-major (test_utils.jl:52-118)
    systems, traj, _ = wl.config("hadamard", T=5)
    assert list(traj.names) == ["Ũ⃗", "a", "da", "dda", "Δt"]
    assert traj.dim == 15 and traj.dims.states == 12  # integrator_test_1qubit.jl:44 uses Z.dims.states
    z = traj.datavec
    assert np.array_equal(z[15:30], traj.data[:, 1])
    assert np.array_equal(qcknot.operator_to_iso_vec(np.eye(2)), [1, 0, 0, 0, 0, 1, 0, 0])


# ---- round 2: host-buffer path (compact layout + threaded expansion) and multi-GPU handles, host side ------------------------
def _expected_expand(D, arr, nnz, full, nk):
    """What the host half must produce from the compact form of `full`: owned positions, kron copies = first copy."""
    segs = D.compact_map(arr)
    C_ = int(segs[:, 2].sum())
    comp = np.empty(nk * C_)
    want = np.full(nk * nnz, np.nan)
    f2, c2, w2 = full.reshape(nk, nnz), comp.reshape(nk, C_), want.reshape(nk, nnz)
    for fo, co, ln, rep in segs:
        c2[:, co:co + ln] = f2[:, fo:fo + ln]
        for r in range(rep):
            w2[:, fo + r * ln:fo + (r + 1) * ln] = f2[:, fo:fo + ln]
    return segs, comp, want


@pytest.mark.parametrize("name,kw", [("cz", {"T": 9}), ("hadamard", {"T": 40}), ("ket", {"T": 12}),
                                     ("sampling", {"T": 5, "n_systems": 7}), ("cz", {"T": 5, "integrator": "exponential"})])
def test_compact_map_covers_every_position_once_and_expands(name, kw):
    systems, traj, integrators = wl.config(name, **kw)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1)
    nk = D.n_blocks
    rng = np.random.default_rng(5)
    for arr, nnz in ((0, D.dyn), (1, D.nnzJ), (2, D.nnzH)):
        full = rng.standard_normal(nk * nnz)
        segs, comp, want = _expected_expand(D, arr, nnz, full, nk)
        cover = np.zeros(nnz, dtype=int)
        for fo, co, ln, rep in segs:
            cover[fo:fo + ln * rep] += 1
        assert np.all(cover == 1), "an unsharded handle writes every position of the knot block exactly once"
        assert np.array_equal(np.cumsum(np.r_[0, segs[:-1, 2]]), segs[:, 1])
        out = np.full(nk * nnz, np.nan)
        D.expand_host(arr, comp, out, nk)
        assert np.array_equal(out, want)
    if name == "cz" and "integrator" not in kw:
        # 5,832 of the 6,674 Jacobian values per knot are 9-fold replicas of two 18x18 blocks
        J = D.compact_map(1)
        assert int(J[:, 2].sum()) == 6674 - 5832 + 2 * 324 and sorted(J[J[:, 3] > 1][:, 3].tolist()) == [9, 9]
        assert int(D.compact_map(0)[:, 2].sum()) + int(J[:, 2].sum()) + int(D.compact_map(2)[:, 2].sum()) == 3303
    D.close()


def test_multi_gpu_handle_partitions_host_side():
    """n_gpus > 1 on a structure-only handle: the partition the library would use, no device needed."""
    systems, traj, integrators = wl.config("cz", T=1001)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1, n_gpus=8, shard_mode="knot")
    sh = D.shards()
    assert len(sh) == 8 and sh[0][1] == 0 and sh[-1][2] == 1000
    assert all(a[2] == b[1] for a, b in zip(sh, sh[1:])) and all(110 <= s[2] - s[1] <= 140 for s in sh)
    assert all(s[3:] == (0, len(integrators)) for s in sh)
    assert (D.dyn, D.nnzJ, D.nnzH) == (170, 6674, 1643)
    with pytest.raises(qcknot.QcknotError):
        D.F(traj.datavec)  # no CPU evaluation path, multi-GPU or not
    D.close()
    systems, traj, integrators = wl.config("sampling", T=4, n_systems=10)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1, n_gpus=4, shard_mode="ensemble")
    sh = D.shards()
    assert [s[3] for s in sh] == [0, 2, 5, 7] and sh[-1][4] == len(integrators) and all(s[1:3] == (0, 3) for s in sh)
    D.close()
    with pytest.raises(qcknot.QcknotError):
        qcknot.QuantumDynamics(integrators, traj, device=-1, n_gpus=16, shard_mode="ensemble")  # 10 systems over 16 GPUs
    systems, traj, integrators = wl.config("hadamard", T=4)
    with pytest.raises(qcknot.QcknotError):
        qcknot.QuantumDynamics(integrators, traj, device=-1, n_gpus=8, shard_mode="knot")  # 3 blocks over 8 GPUs


def test_ensemble_children_exclude_shared_entries_from_their_segments():
    systems, traj, integrators = wl.config("sampling", T=3, n_systems=4)
    full = qcknot.QuantumDynamics(integrators, traj, device=-1)
    shared = set(full.shared_hessian_positions().tolist())
    cover = np.zeros(full.nnzH, dtype=int)
    for q0, q1 in ((0, 2), (2, len(integrators))):
        S = qcknot.QuantumDynamics(integrators, traj, device=-1, integrator_range=(q0, q1))
        for fo, co, ln, rep in S.compact_map(2):
            cover[fo:fo + ln * rep] += 1
        S.close()
    # manual integrator ranges keep their local partial sums at the shared positions (both shards write them) ...
    assert all(cover[p] == 2 for p in shared) and all(cover[p] == 1 for p in range(full.nnzH) if p not in shared)
    full.close()
