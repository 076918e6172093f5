"""Rollouts with the reference's names, propagated on the device.

  unitary_rollout(Ũ⃗_init, controls, Δt, system)      src/trajectory_initialization.jl:426
  rollout(ψ̃_init, controls, Δt, system)              src/trajectory_initialization.jl:493
  unitary_rollout_fidelity(...)                       unitary_smooth_pulse_problem.jl:218-220 (every template test's assertion)

`system` may be a list of systems sharing the controls (robustness sweeps, unitary_sampling_problem.jl:233-243): the batch is
propagated in one call.  No CPU path: the arithmetic happens in libqcknot.so (csrc/qck_rollout.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from .quantum_system import QuantumSystem


def _rollout(init, controls, dt, system, ket: bool, device: int):
    systems = list(system) if isinstance(system, (list, tuple)) else [system]
    N, nd = systems[0].levels, systems[0].n_drives
    controls = np.atleast_2d(np.asarray(controls, dtype=np.float64))
    if controls.shape[0] != nd:
        raise ValueError(f"controls have {controls.shape[0]} rows, the system has {nd} drives")
    T = controls.shape[1]
    dt = np.broadcast_to(np.asarray(dt, dtype=np.float64).ravel(), (T,)) if np.ndim(dt) else np.full(T, float(dt))
    dim = 2 * N * (1 if ket else N)
    Hd = np.ascontiguousarray(np.concatenate([s.drift_reim() for s in systems]))
    Hv = np.ascontiguousarray(np.concatenate([s.drives_reim() for s in systems]))
    a = np.ascontiguousarray(controls.reshape(-1, order="F"))
    dtc = np.ascontiguousarray(dt, dtype=np.float64)
    x0 = None
    if init is not None:
        x0 = np.asarray(init, dtype=np.float64)
        x0 = np.ascontiguousarray(np.broadcast_to(x0.reshape(-1, dim) if x0.ndim > 1 else x0, (len(systems), dim)))
    out = np.empty((len(systems), T, dim))
    lib = _lib.load()
    p = lambda v: None if v is None else v.ctypes.data_as(C.c_void_p)
    rc = lib.qck_rollout(int(device), int(ket), N, nd, len(systems), p(Hd), p(Hv), T, p(a), p(dtc), p(x0), p(out))
    if rc != 0:
        raise RuntimeError(f"qck_rollout failed ({rc}): {lib.qck_rollout_last_error().decode()}")
    res = [np.asfortranarray(out[s].T) for s in range(len(systems))]  # dim x T like a trajectory component
    return res if isinstance(system, (list, tuple)) else res[0]


def unitary_rollout(U_init_iso, controls, dt, system: Union[QuantumSystem, Sequence[QuantumSystem]], device: int = 0):
    """Iso-vec trajectory (2N^2 x T) of U_{t+1} = exp(dt_t G(a_t)) U_t; U_init_iso = None starts from the identity."""
    return _rollout(U_init_iso, controls, dt, system, False, device)


def rollout(psi_init_iso, controls, dt, system: Union[QuantumSystem, Sequence[QuantumSystem]], device: int = 0):
    """Iso-ket trajectory (2N x T)."""
    return _rollout(psi_init_iso, controls, dt, system, True, device)


def iso_vec_unitary_fidelity(U_iso, goal_iso, subspace: Optional[Sequence[int]] = None) -> float:
    """|tr(U_goal' U)|^2 / n^2 on the (sub)space (unitary_minimum_time_problem.jl:73-76)."""
    from .isomorphisms import iso_vec_to_operator
    U, G = iso_vec_to_operator(U_iso), iso_vec_to_operator(goal_iso)
    if subspace is not None:
        ix = np.ix_(list(subspace), list(subspace))
        U, G = U[ix], G[ix]
    return float(abs(np.trace(G.conj().T @ U)) ** 2 / U.shape[0] ** 2)


def unitary_rollout_fidelity(goal, controls, dt, system, subspace=None, U_init_iso=None, device: int = 0):
    """Fidelity of the rolled-out final unitary with `goal` (operator or iso-vec); a list of systems gives a list."""
    from .isomorphisms import operator_to_iso_vec
    g = np.asarray(goal)
    g = operator_to_iso_vec(g) if g.ndim == 2 else g
    R = unitary_rollout(U_init_iso, controls, dt, system, device)
    if isinstance(R, list):
        return [iso_vec_unitary_fidelity(r[:, -1], g, subspace) for r in R]
    return iso_vec_unitary_fidelity(R[:, -1], g, subspace)
