"""The remaining integrator families of the reference's templates through the same kernels:
QuantumStateSamplingProblem (quantum_state_sampling_problem.jl:99-122), UnitaryDirectSumProblem's suffix-rebuilt integrators
(unitary_direct_sum_problem.jl:125-128), UnitaryBangBangProblem (Pade order 12 + slack components,
unitary_bang_bang_problem.jl:162-175,208) and DensityOperatorSmoothPulseProblem's DensityOperatorExponentialIntegrator
(density_operator_smooth_pulse_problem.jl:104-106); plus config 5 at its real ensemble size."""
import numpy as np
import pytest

import qcknot
from qcknot import workloads as wl
from oracle.bridge import entry_err, oracle_dynamics, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def check(integrators, traj, eval_hessian=True):
    D = qcknot.QuantumDynamics(integrators, traj, eval_hessian=eval_hessian)
    O = oracle_dynamics(integrators, traj, eval_hessian=eval_hessian)
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    assert np.array_equal(D.dF_structure, np.array(O.dF_structure))
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and entry_err(J, O.dF(Z)) < 1e-9
    if eval_hessian:
        assert np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure).reshape(-1, 2))
        assert rel_err(H, O.mu_d2F(Z, mu)) < TOL and entry_err(H, O.mu_d2F(Z, mu)) < 1e-9
    D.close()


@pytest.mark.parametrize("integrator", ["pade", "exponential"])
def test_quantum_state_sampling(integrator):
    """Kets of several sampled systems sharing the controls: one QuantumState*Integrator per system."""
    systems = wl.sampling_systems(5, levels=3)
    traj = wl.random_pulse_trajectory(systems, 7, 0.2, ket=True, a_bound=0.2)
    check(wl.build_integrators(systems, traj, integrator=integrator, ket=True), traj)


def test_unitary_direct_sum():
    """Two unitary integrators with their own (suffixed) state and control components in one trajectory."""
    rng = np.random.default_rng(0)
    s1, s2 = wl.pauli_system(0.1), wl.random_hermitian_system(2, 2, seed=5)
    T = 9
    U = lambda: np.stack([qcknot.operator_to_iso_vec(wl._random_unitary(rng, 2)) + 1e-3 * rng.normal(size=8) for _ in range(T)], axis=1)
    comps = {"Ũ⃗1": U(), "Ũ⃗2": U()}
    for k in ("a", "da", "dda"):
        for sfx in ("1", "2"):
            comps[k + sfx] = rng.normal(0, 0.3, size=(2, T))
    comps["Δt"] = rng.uniform(0.1, 0.3, size=(1, T))
    traj = qcknot.NamedTrajectory(comps, controls=("dda1", "dda2", "Δt"), timestep="Δt")
    integ = [qcknot.UnitaryPadeIntegrator("Ũ⃗1", "a1", s1, traj), qcknot.UnitaryPadeIntegrator("Ũ⃗2", "a2", s2, traj)]
    integ += [qcknot.DerivativeIntegrator(x + s, dx + s, traj) for x, dx in (("a", "da"), ("da", "dda")) for s in ("1", "2")]
    check(integ, traj)


def test_bang_bang_shape_pade_order_12():
    """UnitaryBangBangProblem: pade_order = 12 (unitary_bang_bang_problem.jl:208) and slack components next to the controls."""
    rng = np.random.default_rng(1)
    sys_ = wl.pauli_system(1.0)
    T = 7
    comps = {"Ũ⃗": np.stack([qcknot.operator_to_iso_vec(wl._random_unitary(rng, 2)) for _ in range(T)], axis=1),
             "a": rng.uniform(-1, 1, size=(2, T)), "da": rng.normal(0, 0.1, size=(2, T)), "dda": rng.normal(0, 0.1, size=(2, T)),
             "s1": rng.uniform(0, 1, size=(2, T)), "s2": rng.uniform(0, 1, size=(2, T)), "Δt": rng.uniform(0.1, 0.3, size=(1, T))}
    traj = qcknot.NamedTrajectory(comps, controls=("dda", "s1", "s2", "Δt"), timestep="Δt")
    integ = [qcknot.UnitaryPadeIntegrator("Ũ⃗", "a", sys_, traj, order=12), qcknot.DerivativeIntegrator("a", "da", traj),
             qcknot.DerivativeIntegrator("da", "dda", traj)]
    check(integ, traj)


@pytest.mark.parametrize("levels", [2, 3])
def test_density_operator_exponential(levels):
    """Lindblad dynamics of a driven, decaying qudit: the ket exponential integrator on N^2 levels with the Lindbladian."""
    rng = np.random.default_rng(2)
    a = wl.annihilate(levels)
    sys_ = qcknot.OpenQuantumSystem(0.3 * a.conj().T @ a, [a + a.conj().T, 1j * (a - a.conj().T)], [0.4 * a, 0.2 * a.conj().T @ a])
    T, n2 = 6, levels * levels
    rho = []
    for _ in range(T):
        M = rng.normal(size=(levels, levels)) + 1j * rng.normal(size=(levels, levels))
        r = M @ M.conj().T
        rho.append(qcknot.ket_to_iso((r / np.trace(r)).reshape(-1, order="F")))
    comps = {"ρ⃗̃": np.stack(rho, axis=1), "a": rng.uniform(-0.5, 0.5, size=(2, T)), "da": rng.normal(0, 0.1, size=(2, T)),
             "dda": rng.normal(0, 0.1, size=(2, T)), "Δt": rng.uniform(0.1, 0.3, size=(1, T))}
    traj = qcknot.NamedTrajectory(comps, controls=("dda", "Δt"), timestep="Δt")
    assert sys_.levels == n2 and len(traj.components["ρ⃗̃"]) == 2 * n2
    integ = [qcknot.DensityOperatorExponentialIntegrator("ρ⃗̃", "a", sys_, traj), qcknot.DerivativeIntegrator("a", "da", traj),
             qcknot.DerivativeIntegrator("da", "dda", traj)]
    check(integ, traj)
    # the integrator is exact for constant controls: trace preservation of the propagated density operator
    import scipy.linalg as sl
    z = traj.data[:, 0]
    G = sys_.G(z[traj.components["a"].start:traj.components["a"].stop])
    v = sl.expm(z[-1] * G) @ z[: 2 * n2]
    assert abs(np.trace((v[:n2] + 1j * v[n2:]).reshape(levels, levels, order="F")) - 1.0) < 1e-12


def test_config5_full_ensemble_256_systems():
    """BASELINE configs[4] at its real ensemble size (256 systems, 4 levels), small T: parity of the whole problem."""
    systems, traj, integrators = wl.config("sampling", T=4, n_systems=256)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    assert (D.dyn, D.nnzJ, D.nnzH) == (8196, 155664, 49162)
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    H_ref = O.mu_d2F(Z, mu)
    assert rel_err(F, O.F(Z)) < TOL and rel_err(J, O.dF(Z)) < TOL and rel_err(H, H_ref) < TOL
    shared = D.shared_hessian_positions()
    idx = (np.arange(D.n_blocks)[:, None] * D.nnzH + shared[None, :]).reshape(-1)
    assert entry_err(H[idx], H_ref[idx]) < 1e-9  # the entries 256 systems add up in
    D.close()
