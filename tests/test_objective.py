"""Objective / terminal-constraint terms: the oracle restatement against finite differences (CPU), and
the device evaluation through the C-ABI against the oracle (GPU)."""
import numpy as np
import pytest

import qcknot
from qcknot import workloads as wl
from oracle import objective_oracle as oo
from oracle.bridge import rel_err


def _problem(name="hadamard", T=9, free_time=True):
    systems, traj, integrators = wl.config(name, T=T, free_time=free_time)
    N = systems[0].levels
    rng = np.random.default_rng(3)
    goal = np.linalg.qr(rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)))[0]
    state = [n for n in traj.names if n not in ("a", "da", "dda", "Δt")][0]
    return systems, traj, integrators, state, goal, N


def _oracle_objective(traj, state, goal, N, free_time, subspace=None):
    dt_off = traj.components["Δt"].start if free_time else -1
    dtf = 0.0 if free_time else traj.timestep
    G = goal.copy()
    if subspace is not None:
        m = np.zeros((N, N)); m[np.ix_(subspace, subspace)] = 1.0
        G = G * m
    terms = [oo.UnitaryInfidelityObjective(traj.components[state], qcknot.operator_to_iso_vec(G), N, 100.0, len(subspace) if subspace else N),
             oo.QuadraticRegularizer(traj.components["a"], 1e-2, dt_off, dtf),
             oo.QuadraticRegularizer(traj.components["da"], [0.3, 0.7][: len(traj.components["da"])] if len(traj.components["da"]) == 2 else 0.5, dt_off, dtf),
             oo.QuadraticRegularizer(traj.components["dda"], 1e-2, dt_off, dtf)]
    if free_time:
        terms.append(oo.MinimumTimeObjective(dt_off, 3.0))
    return oo.Objective(terms, traj.T, traj.dim)


def _product_objective(traj, state, goal, free_time, subspace=None):
    J = qcknot.UnitaryInfidelityObjective(state, traj, 100.0, subspace=subspace, goal=goal)
    J += qcknot.QuadraticRegularizer("a", traj, 1e-2)
    nda = len(traj.components["da"])
    J += qcknot.QuadraticRegularizer("da", traj, [0.3, 0.7] if nda == 2 else 0.5)
    J += qcknot.QuadraticRegularizer("dda", traj, 1e-2)
    if free_time:
        J += qcknot.MinimumTimeObjective(traj, D=3.0)
    return J


@pytest.mark.parametrize("free_time", [True, False])
def test_oracle_objective_against_finite_differences(free_time):
    systems, traj, integrators, state, goal, N = _problem(free_time=free_time)
    O = _oracle_objective(traj, state, goal, N, free_time)
    Z = traj.datavec[: traj.T * traj.dim].copy()
    g = O.gradient(Z)
    rng = np.random.default_rng(0)
    for _ in range(5):
        d = rng.normal(size=Z.size)
        eps = 1e-6
        fd = (O.value(Z + eps * d) - O.value(Z - eps * d)) / (2 * eps)
        assert abs(fd - g @ d) < 1e-6 * max(1.0, abs(fd))
    S, vals = O.hessian_structure(), O.hessian(Z, 1.0)
    Hd = qcknot.dense(vals, S, (Z.size, Z.size))
    d = rng.normal(size=Z.size)
    fdh = (O.gradient(Z + 1e-6 * d) - O.gradient(Z - 1e-6 * d)) / 2e-6
    assert np.max(np.abs(fdh - Hd @ d)) < 1e-5 * max(1.0, np.max(np.abs(fdh)))
    assert np.all(S[:, 0] <= S[:, 1])
    # fidelity of the goal itself is one; constraint value and Jacobian against finite differences
    con = oo.FinalUnitaryFidelityConstraint(traj.components[state], qcknot.operator_to_iso_vec(goal), N, 0.99)
    Zg = Z.copy()
    Zg[(traj.T - 1) * traj.dim + traj.components[state].start:(traj.T - 1) * traj.dim + traj.components[state].stop] = qcknot.operator_to_iso_vec(goal)
    assert abs(con.value(Zg, traj.T, traj.dim) - 0.01) < 1e-12
    jac = con.jacobian(Z, traj.T, traj.dim)
    cols = con.jacobian_columns(traj.T, traj.dim) - 1
    e = np.zeros(Z.size); e[cols] = rng.normal(size=cols.size)
    fd = (con.value(Z + 1e-6 * e, traj.T, traj.dim) - con.value(Z - 1e-6 * e, traj.T, traj.dim)) / 2e-6
    assert abs(fd - jac @ e[cols]) < 1e-8


def test_objective_structure_without_a_device():
    systems, traj, integrators, state, goal, N = _problem()
    D = qcknot.QuantumDynamics(integrators, traj, device=-1)
    D.attach_objective(_product_objective(traj, state, goal, True))
    O = _oracle_objective(traj, state, goal, N, True)
    assert D.n_vars == traj.T * traj.dim
    assert np.array_equal(D.objective_hessian_structure, O.hessian_structure())
    with pytest.raises(qcknot.QcknotError):
        D.objective(traj.datavec)
    D.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,free_time,subspace", [("hadamard", True, None), ("hadamard", False, None), ("cz", True, None), ("cz", True, [0, 1, 3, 4])])
def test_device_objective_matches_oracle(name, free_time, subspace):
    systems, traj, integrators, state, goal, N = _problem(name, T=40, free_time=free_time)
    D = qcknot.QuantumDynamics(integrators, traj)
    D.attach_objective(_product_objective(traj, state, goal, free_time, subspace))
    O = _oracle_objective(traj, state, goal, N, free_time, subspace)
    Z = traj.datavec[: traj.T * traj.dim].copy()
    assert abs(D.objective(Z) - O.value(Z)) < 1e-12 * max(1.0, abs(O.value(Z)))
    assert rel_err(D.objective_gradient(Z), O.gradient(Z)) < 1e-13
    assert np.array_equal(D.objective_hessian_structure, O.hessian_structure())
    assert rel_err(D.objective_hessian(Z, 0.7), O.hessian(Z, 0.7)) < 1e-13
    assert D.objective(Z) == D.objective(Z)  # fixed-order reduction: bitwise reproducible
    # one upload serves the dynamics and the objective of the same Z
    D.F(Z)
    D.objective(Z)
    assert D.transfer_stats()["h2d_bytes"] == 0
    con = qcknot.FinalUnitaryFidelityConstraint(state, 0.9999, traj, subspace=subspace, goal=goal)
    D.attach_fidelity_constraint(con)
    G = goal.copy()
    if subspace is not None:
        m = np.zeros((N, N)); m[np.ix_(subspace, subspace)] = 1.0
        G = G * m
    oc = oo.FinalUnitaryFidelityConstraint(traj.components[state], qcknot.operator_to_iso_vec(G), N, 0.9999, len(subspace) if subspace else N)
    g, jac, hess = D.fidelity_constraint(Z, mu=-1.3)
    assert abs(g - oc.value(Z, traj.T, traj.dim)) < 1e-13
    assert rel_err(jac, oc.jacobian(Z, traj.T, traj.dim)) < 1e-13 and rel_err(hess, oc.hessian(-1.3)) < 1e-13
    D.close()


@pytest.mark.gpu
def test_objective_on_a_knot_sharded_handle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    systems, traj, integrators, state, goal, N = _problem("cz", T=61)
    O = _oracle_objective(traj, state, goal, N, True)
    Z = traj.datavec[: traj.T * traj.dim].copy()
    D = qcknot.QuantumDynamics(integrators, traj, n_gpus=2, shard_mode="knot")
    D.attach_objective(_product_objective(traj, state, goal, True))
    assert abs(D.objective(Z) - O.value(Z)) < 1e-11 * max(1.0, abs(O.value(Z)))
    assert rel_err(D.objective_gradient(Z), O.gradient(Z)) < 1e-13
    assert rel_err(D.objective_hessian(Z, 1.0), O.hessian(Z, 1.0)) < 1e-13
    D.close()
