"""Rollouts on the device against a scipy expm loop (the reference's unitary_rollout / rollout,
src/trajectory_initialization.jl:426,493) and the fidelity the template tests assert (unitary_smooth_pulse_problem.jl:218-220)."""
import numpy as np
import pytest
import scipy.linalg as sl

import qcknot
from qcknot import workloads as wl

pytestmark = pytest.mark.gpu


def _host_rollout(x0_iso, controls, dt, system, ket):
    N = system.levels
    X = qcknot.iso_to_ket(x0_iso) if ket else qcknot.iso_vec_to_operator(x0_iso)
    cols = [x0_iso.copy()]
    for t in range(controls.shape[1] - 1):
        X = sl.expm(-1j * system.H(controls[:, t]) * dt[t]) @ X
        cols.append(qcknot.ket_to_iso(X) if ket else qcknot.operator_to_iso_vec(X))
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("name,T", [("hadamard", 50), ("cz", 300), ("cz", 33)])
def test_unitary_rollout_matches_expm_loop(name, T):
    systems, traj, _ = wl.config(name, T=T)
    sys_ = systems[0]
    a, dt = traj["a"], traj["Δt"].ravel()
    N = sys_.levels
    U0 = qcknot.operator_to_iso_vec(np.eye(N))
    R = qcknot.unitary_rollout(U0, a, dt, sys_)
    ref = _host_rollout(U0, a, dt, sys_, False)
    assert R.shape == ref.shape and np.max(np.abs(R - ref)) < 1e-10
    # unitarity of the final propagator and the fidelity helper
    U = qcknot.iso_vec_to_operator(R[:, -1])
    assert np.max(np.abs(U.conj().T @ U - np.eye(N))) < 1e-11
    f = qcknot.unitary_rollout_fidelity(qcknot.iso_vec_to_operator(ref[:, -1]), a, dt, sys_)
    assert abs(f - 1.0) < 1e-10
    assert np.array_equal(qcknot.unitary_rollout(None, a, dt, sys_), R)  # identity start by default, bitwise reproducible


def test_ket_rollout_and_fixed_timestep():
    systems, traj, _ = wl.config("ket", T=70, free_time=False)
    sys_ = systems[0]
    a = traj["a"]
    psi0 = qcknot.ket_to_iso(np.array([0.6, 0.8j]))
    R = qcknot.rollout(psi0, a, 0.2, sys_)
    ref = _host_rollout(psi0, a, np.full(a.shape[1], 0.2), sys_, True)
    assert R.shape == (4, 70) and np.max(np.abs(R - ref)) < 1e-11


def test_batched_rollout_over_sampled_systems():
    """unitary_sampling_problem.jl:233-243: fidelity of one pulse on every sampled system."""
    systems = wl.sampling_systems(7, levels=4)
    rng = np.random.default_rng(3)
    T = 90
    a, dt = rng.uniform(-0.1, 0.1, size=(2, T)), rng.uniform(0.1, 0.3, size=T)
    U0 = qcknot.operator_to_iso_vec(np.eye(4))
    Rs = qcknot.unitary_rollout(U0, a, dt, systems)
    assert len(Rs) == 7
    goal = qcknot.iso_vec_to_operator(_host_rollout(U0, a, dt, systems[0], False)[:, -1])
    fids = qcknot.unitary_rollout_fidelity(goal, a, dt, systems, subspace=[0, 1])
    for s, (R, f) in enumerate(zip(Rs, fids)):
        ref = _host_rollout(U0, a, dt, systems[s], False)
        assert np.max(np.abs(R - ref)) < 1e-10
        assert abs(f - qcknot.iso_vec_unitary_fidelity(ref[:, -1], qcknot.operator_to_iso_vec(goal), [0, 1])) < 1e-10
    full = qcknot.unitary_rollout_fidelity(goal, a, dt, systems)  # no subspace: the system the goal came from reaches it exactly
    assert abs(full[0] - 1.0) < 1e-10 and min(full) < 1.0
