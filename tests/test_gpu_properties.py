"""GPU tests beyond small-case oracle parity: committed golden fixtures, size-independent properties at BASELINE sizes,
sharding through the C-ABI, the device-pointer entry point, edge cases."""
import os

import numpy as np
import pytest

import qcknot
from qcknot import workloads as wl
from oracle.bridge import oracle_dynamics, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hadamard_type1.npz"))
TOL = 1e-10


def _type1_problem(kind, free_time):
    """The reference's literal trajectory (test/test_utils.jl:52-118) with QuantumSystem(0.1 Z, [X, Y])."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden as mg
    data = mg.TYPE1 if free_time else mg.TYPE1[:14]
    comps = {"Ũ⃗": data[0:8], "a": data[8:10], "da": data[10:12], "dda": data[12:14]}
    if free_time:
        comps["Δt"] = data[14:15]
    traj = qcknot.NamedTrajectory(comps, controls=("dda", "Δt") if free_time else ("dda",), timestep="Δt" if free_time else 0.2)
    sys_ = qcknot.QuantumSystem(0.1 * mg.Zp, [mg.X, mg.Y])
    return sys_, traj, wl.build_integrators([sys_], traj, integrator="pade" if kind == "pade" else "exponential")


@pytest.mark.parametrize("kind", ["pade", "exp"])
@pytest.mark.parametrize("free_time", [True, False])
def test_golden_fixture(kind, free_time):
    sys_, traj, integrators = _type1_problem(kind, free_time)
    tag = f"{kind}_{'free' if free_time else 'fixed'}"
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec
    assert np.array_equal(Z, GOLD[f"{tag}_Z"])
    assert np.array_equal(D.dF_structure, GOLD[f"{tag}_Js"])
    assert rel_err(D.F(Z), GOLD[f"{tag}_F"]) < TOL
    assert rel_err(D.dF(Z), GOLD[f"{tag}_J"]) < TOL
    assert np.array_equal(D.mu_d2F_structure, GOLD[f"{tag}_Hs"])
    assert rel_err(D.mu_d2F(Z, GOLD[f"{tag}_mu"]), GOLD[f"{tag}_H"]) < TOL


def test_full_size_properties_cz():
    """BASELINE size (T = 10,000, N = 9): properties that need no oracle pass over the whole array.
    (1) the Hessian is linear in mu; (2) mu^T dF equals the directional derivative of mu^T F (central differences);
    (3) kron(I_N, B) structure: the N copies of each 2N x 2N Jacobian block are bit-identical;
    (4) a sample of 16 knot blocks spread over the range matches the oracle; (5) run-to-run bitwise reproducible."""
    systems, traj, integrators = wl.config("cz", T=10000)
    D = qcknot.QuantumDynamics(integrators, traj)
    nb = D.n_blocks
    Z = traj.datavec
    rng = np.random.default_rng(5)
    mu1, mu2 = rng.normal(size=nb * D.dyn), rng.normal(size=nb * D.dyn)
    F, J, H1 = D.eval_all(Z, mu1)
    H2 = D.mu_d2F(Z, mu2)
    H12 = D.mu_d2F(Z, 2.0 * mu1 - 3.0 * mu2)
    assert rel_err(H12, 2.0 * H1 - 3.0 * H2) < 1e-12
    F_b, J_b, H_b = D.eval_all(Z, mu1)
    assert np.array_equal(F, F_b) and np.array_equal(J, J_b) and np.array_equal(H1, H_b)
    # (3) block structure of the state_t Jacobian block: first 4 N^3 values of every knot = N copies of 4 N^2 values
    N = 9
    Jk = J.reshape(nb, D.nnzJ)
    blk = Jk[:, : 4 * N**3].reshape(nb, N, 4 * N * N)
    assert np.array_equal(blk, np.broadcast_to(blk[:, :1], blk.shape))
    # (2) directional derivative
    d = rng.normal(size=Z.size) * 1e-2
    eps = 1e-6
    dd = (mu1 @ D.F(Z + eps * d) - mu1 @ D.F(Z - eps * d)) / (2 * eps)
    s = D.dF_structure
    g = np.zeros(Z.size)
    np.add.at(g, s[:, 1] - 1, J * mu1[s[:, 0] - 1])
    assert abs(g @ d - dd) < 1e-6 * max(1.0, abs(dd))
    # (4) sampled oracle parity
    idx = np.linspace(0, nb - 1, 16).astype(int)
    for t in idx:
        sub = qcknot.NamedTrajectory({n: traj[n][:, t:t + 2] for n in traj.names}, controls=("dda", "Δt"), timestep="Δt")
        O = oracle_dynamics(wl.build_integrators(systems, sub), sub)
        zz, mm = sub.datavec, mu1[t * D.dyn:(t + 1) * D.dyn]
        assert rel_err(F[t * D.dyn:(t + 1) * D.dyn], O.F(zz)) < TOL
        assert rel_err(Jk[t], O.dF(zz)) < TOL
        assert rel_err(H1[t * D.nnzH:(t + 1) * D.nnzH], O.mu_d2F(zz, mm)) < TOL


def test_config4_upper_end_T100000_device_resident():
    """BASELINE config 4 at its upper end (T = 100,000, N = 9; 7.2 GB of values): device-resident pass, a sample of 24 knot blocks
    spread over the whole range (first, last, chunk borders of the persistent loop) against the oracle, the kron(I_N, B)
    replicas bit-identical over the whole Jacobian, run-to-run bitwise reproducible (checked on the device)."""
    import torch
    systems, traj, integrators = wl.config("cz", T=100000)
    D = qcknot.QuantumDynamics(integrators, traj)
    nb = D.n_blocks
    dev = torch.device("cuda:0")
    Z = traj.datavec
    mu = wl.random_multipliers(nb * D.dyn)
    dZ, dmu = torch.from_numpy(Z).to(dev), torch.from_numpy(mu).to(dev)
    out = [[torch.empty(nb * n, dtype=torch.float64, device=dev) for n in (D.dyn, D.nnzJ, D.nnzH)] for _ in range(2)]
    torch.cuda.synchronize()
    for F, J, H in out:
        D.eval_device(7, dZ.data_ptr(), dmu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), 0)
    D.synchronize()
    for a, b in zip(*out):
        assert torch.equal(a, b)
    F, J, H = out[0]
    N = 9
    blk = J.view(nb, D.nnzJ)[:, : 4 * N**3].view(nb, N, 4 * N * N)
    assert bool((blk == blk[:, :1]).all())
    idx = sorted(set(np.linspace(0, nb - 1, 20).astype(int).tolist() + [1, 147, 148, nb - 2]))
    for t in idx:
        sub = qcknot.NamedTrajectory({n: traj[n][:, t:t + 2] for n in traj.names}, controls=("dda", "Δt"), timestep="Δt")
        O = oracle_dynamics(wl.build_integrators(systems, sub), sub)
        zz, mm = sub.datavec, mu[t * D.dyn:(t + 1) * D.dyn]
        assert rel_err(F[t * D.dyn:(t + 1) * D.dyn].cpu().numpy(), O.F(zz)) < TOL
        assert rel_err(J[t * D.nnzJ:(t + 1) * D.nnzJ].cpu().numpy(), O.dF(zz)) < TOL
        assert rel_err(H[t * D.nnzH:(t + 1) * D.nnzH].cpu().numpy(), O.mu_d2F(zz, mm)) < TOL
    D.close()


def test_two_live_handles_share_a_kernel_with_different_shared_memory():
    """Two problems alive at once that run the SAME kernel instantiation with different dynamic shared-memory sizes (the generic
    spectral kernel takes the level count at run time): the opt-in cap is a property of the function, so the small problem's
    launcher must not lower it under the large one."""
    def problem(levels, seed):
        sys_ = wl.random_hermitian_system(levels, 2, seed=seed, scale=0.3)
        traj = wl.random_pulse_trajectory([sys_], 6, 0.2, seed=seed)
        return traj, wl.build_integrators([sys_], traj, integrator="exponential")
    (trA, iA), (trB, iB) = problem(12, 3), problem(6, 4)
    A, B = qcknot.QuantumDynamics(iA, trA), qcknot.QuantumDynamics(iB, trB)
    muA, muB = wl.random_multipliers(A.n_blocks * A.dyn), wl.random_multipliers(B.n_blocks * B.dyn)
    first = A.eval_all(trA.datavec, muA)
    B.eval_all(trB.datavec, muB)
    again = A.eval_all(trA.datavec + 0.0, muA)
    for x, y in zip(first, again):
        assert np.array_equal(x, y)
    OB = oracle_dynamics(iB, trB)
    assert rel_err(B.dF(trB.datavec), OB.dF(trB.datavec)) < TOL
    A.close(); B.close()


def test_exponential_residual_vanishes_on_exact_propagation():
    """U_{t+1} = exp(-i H(a_t) dt_t) U_t  =>  the exponential residual is zero to rounding, at T = 2,000."""
    import scipy.linalg as sla
    sys_ = wl.two_transmon_cz_system()
    T = 2000
    traj = wl.random_pulse_trajectory([sys_], T, 1.0, seed=3)
    U = np.eye(9, dtype=complex)
    Us = traj["Ũ⃗"]
    for t in range(T):
        Us[:, t] = qcknot.operator_to_iso_vec(U)
        U = sla.expm(-1j * sys_.H(traj["a"][:, t]) * traj["Δt"][0, t]) @ U
    D = qcknot.QuantumDynamics(wl.build_integrators([sys_], traj, integrator="exponential"), traj, eval_hessian=False)
    F = D.F(traj.datavec).reshape(T - 1, D.dyn)
    assert np.abs(F[:, :162]).max() < 5e-13


@pytest.mark.parametrize("world", [2, 3])
def test_knot_shards_equal_full(world):
    systems, traj, integrators = wl.config("cz", T=41)
    full = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers(full.n_blocks * full.dyn)
    F, J, H = full.eval_all(Z, mu)
    Fs, Js, Hs, Ss = [], [], [], []
    for r in range(world):
        sh = qcknot.QuantumDynamics(integrators, traj, knot_range=qcknot.sharding.knot_shard(traj.T - 1, r, world))
        f, j, h = sh.eval_all(Z, mu)  # full arrays in, the shard's contiguous segment out
        Fs.append(f); Js.append(j); Hs.append(h); Ss.append(sh.dF_structure)
    assert np.array_equal(np.concatenate(Fs), F) and np.array_equal(np.concatenate(Js), J) and np.array_equal(np.concatenate(Hs), H)
    assert np.array_equal(np.concatenate(Ss), full.dF_structure)


def test_ensemble_shards_sum_to_full():
    systems, traj, integrators = wl.config("sampling", T=6, n_systems=6)
    full = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers(full.n_blocks * full.dyn)
    F, J, H = full.eval_all(Z, mu)
    world = 3
    Fsum, Jsum, Hsum = np.zeros_like(F), np.zeros_like(J), np.zeros_like(H)
    for r in range(world):
        q0, q1 = qcknot.sharding.integrator_shard(len(systems), len(integrators), r, world)
        sh = qcknot.QuantumDynamics(integrators, traj, integrator_range=(q0, q1))
        f, j, h = np.zeros_like(F), np.zeros_like(J), np.zeros_like(H)  # untouched entries must stay zero
        sh.eval_all(Z, mu, f, j, h)
        Fsum += f; Jsum += j; Hsum += h
    assert np.array_equal(Fsum, F) and np.array_equal(Jsum, J)
    assert rel_err(Hsum, H) < 1e-14  # shared control entries: sum of per-rank partial sums (all-reduce)
    pos = full.shared_hessian_positions()
    mask = np.ones(full.nnzH, bool); mask[pos] = False
    assert np.array_equal(Hsum.reshape(-1, full.nnzH)[:, mask], H.reshape(-1, full.nnzH)[:, mask])


def test_device_pointer_entry_point():
    import torch
    systems, traj, integrators = wl.config("hadamard", T=20)
    D = qcknot.QuantumDynamics(integrators, traj)
    nb = D.n_blocks
    mu_h = wl.random_multipliers(nb * D.dyn)
    Fh, Jh, Hh = D.eval_all(traj.datavec, mu_h)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        Z, mu = torch.from_numpy(traj.datavec).to(dev), torch.from_numpy(mu_h).to(dev)
        F = torch.zeros(nb * D.dyn, dtype=torch.float64, device=dev)
        J = torch.zeros(nb * D.nnzJ, dtype=torch.float64, device=dev)
        H = torch.zeros(nb * D.nnzH, dtype=torch.float64, device=dev)
        n0 = D.launch_count
        D.eval_device(7, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), stream.cuda_stream)
        stream.synchronize()
    assert D.launch_count == n0 + 1  # one fused launch
    assert np.array_equal(F.cpu().numpy(), Fh) and np.array_equal(J.cpu().numpy(), Jh) and np.array_equal(H.cpu().numpy(), Hh)


@pytest.mark.parametrize("levels,nd,ket", [(7, 2, False), (10, 1, False), (12, 2, True), (2, 6, False), (16, 2, False)])
def test_generic_levels_and_many_drives(levels, nd, ket):
    """Levels without a compile-time specialisation (generic kernel), the drive-count limit, kets with N = 12."""
    sys_ = wl.random_hermitian_system(levels, nd, seed=levels + nd, scale=0.5)
    traj = wl.random_pulse_trajectory([sys_], 3, 0.2, seed=5, ket=ket)
    integrators = wl.build_integrators([sys_], traj, ket=ket)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, O.mu_d2F(Z, mu)) < TOL


def test_edge_cases():
    # T = 2: a single knot block; several kets sharing the controls (QuantumStateSmoothPulseProblem with 2 states)
    sys_ = wl.pauli_system(0.1)
    traj = wl.random_pulse_trajectory([sys_], 2, 0.2, ket=True, n_states=2, seed=9)
    integrators = wl.build_integrators([sys_], traj, ket=True)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers(D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, O.mu_d2F(Z, mu)) < TOL
    assert len(D.shared_hessian_positions()) == 6
    # global (free-phase) variables appended to datavec never enter the dynamics
    traj.global_data = {"ϕ": np.array([0.3, -0.2])}
    traj.global_dim = 2
    assert np.array_equal(D.F(traj.datavec), F)
    # wrong sizes are errors, not crashes
    with pytest.raises(ValueError):
        D.F(Z[:-3])
    with pytest.raises(ValueError):
        D.mu_d2F(Z, mu[:-1])
    # more levels than any kernel holds (> 32): reported at create
    big = wl.random_hermitian_system(40, 2, seed=1)
    tb = wl.random_pulse_trajectory([big], 2, 0.1)
    with pytest.raises(qcknot.QcknotError, match="shared memory|image"):
        qcknot.QuantumDynamics(wl.build_integrators([big], tb), tb)


@pytest.mark.parametrize("levels,nd,ket", [(16, 4, False), (24, 2, False), (32, 4, False), (32, 3, True), (20, 1, False)])
def test_large_levels(levels, nd, ket):
    """north_star: 'qudit dimensions up to about 32'.  16 levels with four drives, 24 and 32 levels: the large-level kernel
    (qck_big.cu: operands in shared memory, outputs straight to the value arrays)."""
    sys_ = wl.random_hermitian_system(levels, nd, seed=levels + nd, scale=0.3)
    traj = wl.random_pulse_trajectory([sys_], 4, 0.1, seed=2, ket=ket)
    integrators = wl.build_integrators([sys_], traj, ket=ket)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    assert np.array_equal(D.dF_structure, np.array(O.dF_structure)) and np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure).reshape(-1, 2))
    Z, mu = traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, O.mu_d2F(Z, mu)) < TOL
    assert np.array_equal(D.dF(Z), J) and np.array_equal(D.F(Z), F) and np.array_equal(D.mu_d2F(Z, mu), H)
    D.close()


def test_large_levels_non_hermitian():
    """Non-Hermitian generators (an effective Hamiltonian with loss): A^H products must take the conjugate-transpose path, not the
    anti-Hermitian shortcut."""
    rng = np.random.default_rng(21)
    mk = lambda: 0.15 * (rng.normal(size=(18, 18)) + 1j * rng.normal(size=(18, 18)))
    sys_ = qcknot.QuantumSystem(mk(), [mk(), mk()])
    traj = wl.random_pulse_trajectory([sys_], 4, 0.1, seed=4)
    integrators = wl.build_integrators([sys_], traj)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, O.mu_d2F(Z, mu)) < TOL
    D.close()


_BIG_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import qcknot
from qcknot import workloads as wl
from oracle.bridge import oracle_dynamics, rel_err
for name, kw in (("cz", {{"T": 5}}), ("cz", {{"T": 4, "free_time": False}}), ("sampling", {{"T": 3, "n_systems": 3}})):
    systems, traj, integrators = wl.config(name, **kw)
    if name == "sampling":  # 5-level systems so that the large-level kernel is eligible; shared controls -> partial columns
        systems = wl.sampling_systems(3, levels=5)
        traj = wl.random_pulse_trajectory(systems, 3, 0.2, a_bound=0.1)
        integrators = wl.build_integrators(systems, traj)
    D = qcknot.QuantumDynamics(integrators, traj); O = oracle_dynamics(integrators, traj)
    Z = traj.datavec; mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < 1e-10 and rel_err(J, O.dF(Z)) < 1e-10 and rel_err(H, O.mu_d2F(Z, mu)) < 1e-10, name
print("big ok")
"""


def test_large_level_kernel_on_small_problems():
    """QCK_BIG=1 forces the large-level kernel from 5 levels on: the CZ problem (9 levels) and a 5-level sampling problem whose
    shared-control Hessian entries go through the partial columns."""
    import subprocess, sys
    root = os.path.dirname(HERE)
    e = dict(os.environ)
    e.update({"QCK_BIG": "1", "QCK_RS3": "0"})
    out = subprocess.run([sys.executable, "-c", _BIG_SCRIPT.format(root=root)], env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "big ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("levels,nd,ket,T", [(2, 1, False, 4999), (3, 2, False, 3001), (3, 4, False, 1203), (4, 2, False, 2502), (4, 4, False, 1300),
                                             (2, 2, True, 5003), (3, 3, True, 4001), (4, 2, True, 3333)])
def test_column_kernel_every_block_single_systems(levels, nd, ket, T):
    """Block-staged column kernel (single-system problems, 2..4 levels, unitaries and kets): more knots than resident warps, knot
    counts that leave a partial last group; every block against the C port of the oracle, staged = direct variant bit for bit."""
    import subprocess, sys, tempfile
    from oracle.c_port import CPort
    root = os.path.dirname(HERE)
    sys_ = wl.random_hermitian_system(levels, nd, seed=levels * 10 + nd, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], T, 0.2, seed=3, ket=ket)
    integrators = wl.build_integrators([sys_], traj, ket=ket)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    Fo, Jo, Ho = CPort(oracle_dynamics(integrators, traj)).eval(Z, mu)
    assert rel_err(F, Fo) < TOL and rel_err(J, Jo) < TOL and rel_err(H, Ho) < TOL
    assert np.array_equal(J, D.dF(Z)) and np.array_equal(H, D.mu_d2F(Z, mu)) and np.array_equal(F, D.F(Z))
    D.close()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import qcknot; from qcknot import workloads as wl\n"
            "s = wl.random_hermitian_system(%d, %d, seed=%d, scale=0.4); tr = wl.random_pulse_trajectory([s], %d, 0.2, seed=3, ket=%r)\n"
            "D = qcknot.QuantumDynamics(wl.build_integrators([s], tr, ket=%r), tr)\n"
            "F, J, H = D.eval_all(tr.datavec, wl.random_multipliers(D.n_blocks * D.dyn)); np.savez(sys.argv[1], F=F, J=J, H=H)\n"
            % (root, levels, nd, levels * 10 + nd, T, ket, ket))
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "direct.npz")
        out = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, QCK_COLUMN_STAGED="0"), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        ref = np.load(f)
        assert np.array_equal(F, ref["F"]) and np.array_equal(J, ref["J"]) and np.array_equal(H, ref["H"])


@pytest.mark.parametrize("name,T,integ", [("cz", 2500, "pade"), ("hadamard", 5000, "pade"), ("ket", 3000, "pade"), ("cz", 1300, "exponential"),
                                          ("hadamard", 4000, "exponential"), ("cz", 1300, "pade8")])
def test_every_block_when_ctas_loop_over_many_items(name, T, integ):
    """More work items than resident CTAs: the persistent loop + prefetch path.  Every block is compared with the C port
    of the oracle (Pade) or the numpy oracle on a strided sample (exponential)."""
    from oracle.c_port import CPort
    order = int(integ[4:]) if integ.startswith("pade") and len(integ) > 4 else 4
    kind = "pade" if integ.startswith("pade") else integ
    systems, traj, integrators = wl.config(name, T=T, integrator=kind)
    if order != 4:
        integrators = wl.build_integrators(systems, traj, order=order)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    if integ == "pade":
        Fo, Jo, Ho = CPort(oracle_dynamics(integrators, traj)).eval(Z, mu)
        assert rel_err(F, Fo) < TOL and rel_err(J, Jo) < TOL and rel_err(H, Ho) < TOL
    else:
        for t in range(0, D.n_blocks, 97):
            sub = qcknot.NamedTrajectory({n: traj[n][:, t:t + 2] for n in traj.names}, controls=("dda", "Δt"), timestep="Δt")
            O = oracle_dynamics(wl.build_integrators(systems, sub, integrator=kind, order=order), sub)
            assert rel_err(F[t * D.dyn:(t + 1) * D.dyn], O.F(sub.datavec)) < TOL
            assert rel_err(J[t * D.nnzJ:(t + 1) * D.nnzJ], O.dF(sub.datavec)) < TOL
            assert rel_err(H[t * D.nnzH:(t + 1) * D.nnzH], O.mu_d2F(sub.datavec, mu[t * D.dyn:(t + 1) * D.dyn])) < TOL
