"""A whole solve through the evaluator, the way the reference's template tests exercise theirs: build the problem of
UnitarySmoothPulseProblem (README.md:108-115: Hadamard gate, drives X and Y; objective = UnitaryInfidelityObjective + three
QuadraticRegularizers, unitary_smooth_pulse_problem.jl:132-153; constraints = the dynamics + bounds + initial / final values),
hand value / gradient / Hessian of the objective and residual / Jacobian / Hessian-of-Lagrangian of the dynamics -- every one
evaluated by libqcknot.so on the GPU -- to an interior-point-style NLP solver, and assert what the reference's tests assert:
the rollout fidelity after the solve beats the one before (unitary_smooth_pulse_problem.jl:218-221).  Ipopt is not available here
(no Julia, no Ipopt.jl); scipy's trust-constr plays the solver: same callback contract (sparse Jacobian values in a fixed
structure, Hessian of the Lagrangian from (x, multipliers))."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import Bounds, NonlinearConstraint, minimize

import qcknot
from qcknot import workloads as wl
from qcknot.isomorphisms import operator_to_iso_vec

pytestmark = pytest.mark.gpu

HADAMARD = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)


def _problem(T=20, dt=0.2, integrator="pade", seed=3):
    sys_ = wl.pauli_system(0.0)
    rng = np.random.default_rng(seed)
    a = rng.uniform(-0.2, 0.2, size=(2, T))
    a[:, 0] = a[:, -1] = 0.0
    dts = np.full(T, dt)
    U = qcknot.unitary_rollout(None, a, dts, sys_)  # states start on the dynamics (trajectory_initialization.jl:426)
    da = np.zeros((2, T))
    da[:, :-1] = np.diff(a, axis=1) / dt
    dda = np.zeros((2, T))
    dda[:, :-1] = np.diff(da, axis=1) / dt
    traj = qcknot.NamedTrajectory({"Ũ⃗": U, "a": a, "da": da, "dda": dda, "Δt": dts[None, :]}, controls=("dda", "Δt"), timestep="Δt",
                                  goal={"Ũ⃗": operator_to_iso_vec(HADAMARD)})
    return sys_, traj, wl.build_integrators([sys_], traj, integrator=integrator)


def _sym(vals, structure, n):
    """Upper-triangular (values, structure) -> full symmetric sparse matrix (duplicates add, test_utils.jl:14-27)."""
    r, c = structure[:, 0] - 1, structure[:, 1] - 1
    off = r != c
    return sp.coo_matrix((np.concatenate([vals, vals[off]]), (np.concatenate([r, c[off]]), np.concatenate([c, r[off]]))), shape=(n, n)).tocsr()


@pytest.mark.parametrize("integrator", ["pade", "exponential"])
def test_smooth_pulse_solve_improves_fidelity(integrator):
    sys_, traj, integrators = _problem(integrator=integrator)
    T, zdim = traj.T, traj.dim
    n = T * zdim
    D = qcknot.QuantumDynamics(integrators, traj)
    J = (qcknot.UnitaryInfidelityObjective("Ũ⃗", traj, Q=100.0) + qcknot.QuadraticRegularizer("a", traj, 1e-2)
         + qcknot.QuadraticRegularizer("da", traj, 1e-2) + qcknot.QuadraticRegularizer("dda", traj, 1e-2))
    D.attach_objective(J)
    m = D.n_blocks * D.dyn
    Js, Hs, Os = D.dF_structure, D.mu_d2F_structure, D.objective_hessian_structure
    calls = {"F": 0, "J": 0, "H": 0}

    def con(z):
        calls["F"] += 1
        return D.F(z)

    def con_jac(z):
        calls["J"] += 1
        return sp.coo_matrix((D.dF(z), (Js[:, 0] - 1, Js[:, 1] - 1)), shape=(m, n)).tocsr()

    def con_hess(z, v):
        calls["H"] += 1
        return _sym(D.mu_d2F(z, v), Hs, n)

    # bounds, initial and final values as the template sets them (unitary_smooth_pulse_problem.jl:84-117)
    lb, ub = np.full((T, zdim), -np.inf), np.full((T, zdim), np.inf)
    c = traj.components
    lb[:, c["a"]], ub[:, c["a"]] = -1.0, 1.0
    lb[:, c["dda"]], ub[:, c["dda"]] = -5.0, 5.0
    lb[:, c["Δt"]], ub[:, c["Δt"]] = 0.1, 0.3
    lb[:, c["Ũ⃗"]], ub[:, c["Ũ⃗"]] = -1.0, 1.0          # entries of a unitary (keeps the infidelity bounded off the dynamics)
    z0 = traj.datavec.copy()
    Z0 = z0.reshape(T, zdim)
    lb[0, c["Ũ⃗"]] = ub[0, c["Ũ⃗"]] = Z0[0, c["Ũ⃗"]]      # U_1 = I
    for t in (0, T - 1):
        lb[t, c["a"]] = ub[t, c["a"]] = 0.0             # a_1 = a_T = 0

    goal = traj.goal["Ũ⃗"]
    before = qcknot.unitary_rollout_fidelity(goal, traj["a"], traj["Δt"].ravel(), sys_)
    res = minimize(D.objective, z0, jac=D.objective_gradient, hess=lambda z: _sym(D.objective_hessian(z), Os, n), method="trust-constr",
                   constraints=[NonlinearConstraint(con, 0.0, 0.0, jac=con_jac, hess=con_hess)],
                   bounds=Bounds(lb.ravel(), ub.ravel(), keep_feasible=False),
                   options={"maxiter": 300, "gtol": 1e-6, "xtol": 1e-10, "initial_constr_penalty": 10.0})
    Zs = res.x.reshape(T, zdim)
    after = qcknot.unitary_rollout_fidelity(goal, Zs[:, c["a"]].T, Zs[:, c["Δt"]].ravel(), sys_)
    assert np.abs(D.F(res.x)).max() < 1e-3, "the solver left the dynamics infeasible"
    assert after > before and after > 0.99, (before, after, res.status, res.nit)
    assert calls["J"] > 0 and calls["H"] > 0  # the second-order path was the one the solver drove
    D.close()
