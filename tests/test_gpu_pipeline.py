"""GPU tests of the host-buffer path (qck_pipe.cpp) and of the multi-GPU handle (qck_multi.cpp).

The host arrays the reference-facing calls fill must be BIT-IDENTICAL to a plain D2H copy of the device value arrays
(the compact kron transfer + host-side expansion + chunk overlap are transport, not arithmetic), for every chunking; an
unchanged Z must not be uploaded or evaluated again; one handle driving several GPUs must fill the caller's single arrays
exactly like one GPU does."""
import os

import numpy as np
import pytest
import torch

import qcknot
from qcknot import workloads as wl
from oracle.bridge import oracle_dynamics, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _d2h(ptr, count, device):
    """cudaMemcpy of `count` doubles from a raw device pointer (the library's own buffers) into a numpy array."""
    import ctypes
    host = np.empty(count)
    with torch.cuda.device(device):
        rc = ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(ptr),
                                                       ctypes.c_size_t(count * 8), 2)
    assert rc == 0
    return host


def _device_reference(D, Z, mu):
    """Device-resident pass on torch buffers + plain torch D2H: the arrays 'as round 1 shipped them'."""
    dev = torch.device(f"cuda:{D.device}")
    nb = D.n_blocks
    Zd, mud = torch.from_numpy(Z).to(dev), torch.from_numpy(mu).to(dev)
    F = torch.zeros(nb * D.dyn, dtype=torch.float64, device=dev)
    J = torch.zeros(nb * D.nnzJ, dtype=torch.float64, device=dev)
    H = torch.zeros(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        D.eval_device(7, Zd.data_ptr(), mud.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st.cuda_stream)
    st.synchronize()
    return F.cpu().numpy(), J.cpu().numpy(), H.cpu().numpy()[: nb * D.nnzH]


@pytest.mark.parametrize("name,kw,chunk", [
    ("cz", {"T": 301}, None), ("cz", {"T": 301}, "64"), ("cz", {"T": 1000}, "70"), ("hadamard", {"T": 777}, "100"),
    ("sampling", {"T": 40, "n_systems": 6}, "64"), ("ket", {"T": 500}, "128"), ("cz", {"T": 130, "integrator": "exponential"}, "64"),
])
def test_host_arrays_bit_identical_to_plain_copy(name, kw, chunk, monkeypatch):
    if chunk:
        monkeypatch.setenv("QCK_CHUNK_KNOTS", chunk)  # several chunks, ring reuse, ragged last chunk
    systems, traj, integrators = wl.config(name, **kw)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec[: traj.T * D.zdim].copy()
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    Fr, Jr, Hr = _device_reference(D, Z, mu)
    F, J, H = D.eval_all(Z, mu)
    assert np.array_equal(F, Fr) and np.array_equal(J, Jr) and np.array_equal(H, Hr)
    # separate callbacks, NaN-prefilled outputs: every position is written
    F2, J2, H2 = np.full_like(F, np.nan), np.full_like(J, np.nan), np.full_like(H, np.nan)
    D.F(Z, out=F2), D.dF(Z, out=J2), D.mu_d2F(Z, mu, out=H2)
    assert np.array_equal(F2, Fr) and np.array_equal(J2, Jr) and np.array_equal(H2, Hr)
    st = D.transfer_stats()
    n_comp = sum(int(D.compact_map(a)[:, 2].sum()) for a in (2,))
    assert st["d2h_bytes"] == 8 * D.n_blocks * n_comp  # last call: the Hessian in compact form
    D.close()


@pytest.mark.parametrize("name,kw,chunk", [("cz", {"T": 301}, None), ("cz", {"T": 1000}, "70"), ("hadamard", {"T": 777}, "100"),
                                           ("cz", {"T": 130, "integrator": "exponential"}, "64"), ("ket", {"T": 500}, "128"),
                                           ("sampling", {"T": 40, "n_systems": 6}, "64")])
def test_page_locked_outputs_take_the_direct_path(name, kw, chunk, monkeypatch):
    """Caller arrays registered with qck_host_register: arrays without repeated blocks (F, the Hessian values) are written by the
    copy engine straight to their final place, chunk by chunk (no pack, no staging, no host copy); the Jacobian keeps the packed
    path.  Same bytes over the link, bit-identical arrays, single-array callbacks included."""
    if chunk:
        monkeypatch.setenv("QCK_CHUNK_KNOTS", chunk)
    systems, traj, integrators = wl.config(name, **kw)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec[: traj.T * D.zdim].copy()
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    Fr, Jr, Hr = _device_reference(D, Z, mu)
    nb = D.n_blocks
    F, J, H = np.full(nb * D.dyn, np.nan), np.full(nb * D.nnzJ, np.nan), np.full(nb * D.nnzH, np.nan)
    for a in (F, J, H):
        qcknot.host_register(a)
    try:
        D.eval_all(Z, mu, F, J, H)
        assert np.array_equal(F, Fr) and np.array_equal(J, Jr) and np.array_equal(H, Hr)
        n_comp = sum(int(D.compact_map(a)[:, 2].sum()) for a in (0, 1, 2))
        assert D.transfer_stats()["d2h_bytes"] == 8 * nb * n_comp
        J[:] = np.nan
        H[:] = np.nan
        D.dF(Z, out=J)  # cached on the device: copies only
        assert np.array_equal(J, Jr) and D.transfer_stats()["h2d_bytes"] == 0
        D.mu_d2F(Z, mu, out=H)  # a call whose only array goes the direct way (no transfer pieces at all)
        assert np.array_equal(H, Hr) and D.transfer_stats()["d2h_bytes"] == 8 * nb * int(D.compact_map(2)[:, 2].sum())
        Z2 = Z + 1e-7
        F2, J2, H2 = D.eval_all(Z2, mu)  # pageable outputs on the same handle: the packed path
        D.eval_all(Z2, mu, F, J, H)
        assert np.array_equal(F, F2) and np.array_equal(J, J2) and np.array_equal(H, H2)
    finally:
        for a in (F, J, H):
            qcknot.host_unregister(a)
    D.close()


@pytest.mark.parametrize("name,T", [("cz", 200), ("hadamard", 60)])
def test_unchanged_z_is_uploaded_and_evaluated_once(name, T):
    """'the same Z is presented to F, dF, mu d2F in succession'.  (Hadamard, T = 60: the single-piece path of small
    problems, same cache rules.)"""
    systems, traj, integrators = wl.config(name, T=T)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec.copy()
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    l0 = D.launch_count
    F = D.F(Z)
    l1 = D.launch_count
    assert D.transfer_stats()["h2d_bytes"] >= Z.nbytes
    J = D.dF(Z)                      # same Z: no upload, no quantum kernel (the fused F+J pass already ran), only pack + D2H
    s1 = D.transfer_stats()
    l2 = D.launch_count
    assert s1["h2d_bytes"] == 0 and s1["cache_hits"] >= 1
    H = D.mu_d2F(Z, mu)              # same Z, new mu: mu goes up, Z does not
    s2 = D.transfer_stats()
    assert s2["h2d_bytes"] == mu.nbytes
    H_again = D.mu_d2F(Z, mu)        # everything cached
    assert D.transfer_stats()["h2d_bytes"] == 0 and np.array_equal(H, H_again)
    assert (l1 - l0) >= 1 and (l2 - l1) <= 1  # at most the pack kernel
    Z2 = Z.copy()
    Z2[5 * D.zdim + 3] += 1e-3        # a changed Z must be noticed wherever the change is
    F2 = D.F(Z2)
    assert D.transfer_stats()["h2d_bytes"] >= Z.nbytes and not np.array_equal(F, F2)
    O = oracle_dynamics(integrators, traj)
    assert rel_err(F2, O.F(Z2)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, O.mu_d2F(Z, mu)) < TOL
    J2 = D.dF(Z2)
    assert rel_err(J2, O.dF(Z2)) < TOL
    D.close()


def test_full_size_cz_through_pipeline_matches_plain_copy():
    """BASELINE size (T = 10,000): default chunking, the whole 679 MB of values bit-identical to the device arrays."""
    systems, traj, integrators = wl.config("cz", T=10000)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec.copy()
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    Fr, Jr, Hr = _device_reference(D, Z, mu)
    F, J, H = D.eval_all(Z, mu)
    assert np.array_equal(F, Fr) and np.array_equal(J, Jr) and np.array_equal(H, Hr)
    st = D.transfer_stats()
    assert st["d2h_bytes"] == 8 * D.n_blocks * 3303  # 3,303 of 8,487 doubles per knot carry information
    D.close()


def test_exponential_out_of_range_is_reported():
    """ADVICE: ||dt*G||_1 > 4096 needs more squarings than the Hessian tape of the scaling-and-squaring kernel holds: an error,
    not a silent truncation.  (A non-Hermitian generator: the class runs on that kernel; the spectral kernels of the Hermitian
    classes have no such limit.)"""
    rng = np.random.default_rng(3)
    mk = lambda: 0.3 * (rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5)))
    sys_ = qcknot.QuantumSystem(mk(), [mk(), mk()])
    traj = wl.random_pulse_trajectory([sys_], 4, 0.2, seed=11)
    D = qcknot.QuantumDynamics(wl.build_integrators([sys_], traj, integrator="exponential"), traj)
    Z = traj.datavec.copy()
    Z.reshape(traj.T, -1)[1, traj.components["a"]] = 1e6
    with pytest.raises(qcknot.QcknotError, match="4096"):
        D.F(Z)
    D.F(traj.datavec)  # the handle stays usable
    D.close()


# ---- one handle, several GPUs ---------------------------------------------------------------------------------------------
def _ngpu():
    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_knot_sharded_handle_fills_the_callers_arrays(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    systems, traj, integrators = wl.config("cz", T=403)
    D1 = qcknot.QuantumDynamics(integrators, traj)
    torch.cuda.set_device(0)
    Dn = qcknot.QuantumDynamics(integrators, traj, n_gpus=n, shard_mode="knot")
    assert len(Dn.shards()) == n and sorted(s[0] for s in Dn.shards()) == list(range(n))
    Z = traj.datavec.copy()
    mu = wl.random_multipliers(D1.n_blocks * D1.dyn)
    F1, J1, H1 = D1.eval_all(Z, mu)
    Fn, Jn, Hn = Dn.eval_all(Z, mu)
    assert torch.cuda.current_device() == 0  # the library restores the caller's CUDA device
    assert np.array_equal(F1, Fn) and np.array_equal(J1, Jn) and np.array_equal(H1, Hn)
    assert np.array_equal(Dn.dF_structure, D1.dF_structure) and np.array_equal(Dn.mu_d2F_structure, D1.mu_d2F_structure)
    assert np.array_equal(Dn.dF(Z), J1) and np.array_equal(Dn.F(Z), F1) and np.array_equal(Dn.mu_d2F(Z, mu), H1)
    # device-resident: evaluate on every GPU, all-gather the segments over NCCL, every GPU holds the assembled arrays
    Dn.upload(Z, mu)
    Dn.eval_resident(7)
    Dn.gather_device(7)
    Dn.synchronize()
    assert torch.cuda.current_device() == 0
    ver, nranks = Dn.nccl_version()
    assert nranks == n and ver >= 21800
    nb = D1.n_blocks
    for g in (0, n - 1):
        pF, pJ, pH = Dn.gathered_buffers(g)
        dev = Dn.shards()[g][0]
        for ptr, cnt, ref in ((pF, nb * D1.dyn, F1), (pJ, nb * D1.nnzJ, J1), (pH, nb * D1.nnzH, H1)):
            assert np.array_equal(_d2h(ptr, cnt, dev), ref)
    D1.close(), Dn.close()


@pytest.mark.parametrize("n", [2, 4])
def test_ensemble_sharded_handle(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    systems, traj, integrators = wl.config("sampling", T=30, n_systems=9)
    D1 = qcknot.QuantumDynamics(integrators, traj)
    Dn = qcknot.QuantumDynamics(integrators, traj, n_gpus=n, shard_mode="ensemble")
    Z = traj.datavec.copy()
    mu = wl.random_multipliers(D1.n_blocks * D1.dyn)
    F1, J1, H1 = D1.eval_all(Z, mu)
    Fn, Jn, Hn = np.full_like(F1, np.nan), np.full_like(J1, np.nan), np.full_like(H1, np.nan)
    Dn.eval_all(Z, mu, Fn, Jn, Hn)
    assert np.array_equal(F1, Fn) and np.array_equal(J1, Jn)
    shared = D1.shared_hessian_positions()
    idx = (np.arange(D1.n_blocks)[:, None] * D1.nnzH + shared[None, :]).reshape(-1)
    rest = np.setdiff1d(np.arange(H1.size), idx)
    assert np.array_equal(H1[rest], Hn[rest])                       # disjoint entries: bitwise
    assert rel_err(Hn[idx], H1[idx]) < 1e-13                        # shared entries: another (fixed) summation order
    O = oracle_dynamics(integrators, traj)
    assert rel_err(Hn, O.mu_d2F(Z, mu)) < TOL
    Hn2 = Dn.mu_d2F(Z, mu)
    assert np.array_equal(Hn, Hn2)                                  # run-to-run bitwise reproducible
    # device-resident: ncclAllReduce of the shared entries leaves the sums on every GPU
    Dn.invalidate()
    Dn.upload(Z, mu)
    Dn.eval_resident(7)
    Dn.synchronize()
    for g in range(n):
        host = _d2h(Dn.shard_device_buffers(g)[4], H1.size, Dn.shards()[g][0])
        assert rel_err(host[idx], H1[idx]) < 1e-13
    D1.close(), Dn.close()
