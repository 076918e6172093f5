"""Synthetic workloads of BASELINE.json's configs: benchmark Hamiltonians and random-pulse trajectories.

System definitions follow the reference's own templates (they only *generate* inputs; nothing here is on the hot path):
  pauli_system            README.md:108-110 (drift-free X/Y) and unitary_smooth_pulse_problem.jl:206-209 (Z drift)
  two_transmon_cz_system  unitary_robustness_problem.jl:184-194 (3-level transmon pair, 4 drives, 9-dim)
  transmon_system         src/quantum_system_templates/transmons.jl:32-103 (rotating-frame Duffing transmon)
  sampling_systems        unitary_sampling_problem.jl:209-222 (detuning drawn from Normal)
Trajectory synthesis follows SURVEY.md section 8d: seed 1234, a ~ U(-a_bound, a_bound) with zero end points,
da, dda ~ N(0, 0.01), dt ~ U(0.5, 1.5) dt0, states = random unitaries + N(0, 1e-3) perturbation, mu ~ N(0, 1).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .integrators import (
    DerivativeIntegrator,
    QuantumStateExponentialIntegrator,
    QuantumStatePadeIntegrator,
    UnitaryExponentialIntegrator,
    UnitaryPadeIntegrator,
)
from .isomorphisms import ket_to_iso, operator_to_iso_vec
from .quantum_system import QuantumSystem
from .trajectory import NamedTrajectory

PAULI_X = np.array([[0, 1], [1, 0]], dtype=complex)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
PAULI_Z = np.array([[1, 0], [0, -1]], dtype=complex)


def annihilate(levels: int) -> np.ndarray:
    return np.diag(np.sqrt(np.arange(1, levels)), 1).astype(complex)


def pauli_system(drift: float = 0.0) -> QuantumSystem:
    return QuantumSystem(drift * PAULI_Z, [PAULI_X, PAULI_Y])


def two_transmon_cz_system(levels: int = 3, delta: float = -0.1) -> QuantumSystem:
    a = annihilate(levels)
    eye = np.eye(levels)
    a1, a2 = np.kron(a, eye), np.kron(eye, a)
    d1, d2 = a1.conj().T, a2.conj().T
    H_drift = delta / 2 * d1 @ d1 @ a1 @ a1 + delta / 2 * d2 @ d2 @ a2 @ a2
    H_drives = [d1 @ a1, d2 @ a2, d1 @ a2 + a1 @ d2, 1j * (d1 @ a2 - a1 @ d2)]
    return QuantumSystem(H_drift, H_drives)


def transmon_system(levels: int = 3, omega: float = 4.0, delta: float = 0.2, frame_omega: Optional[float] = None,
                    detuning: float = 0.0, amp_scale: float = 1.0) -> QuantumSystem:
    frame_omega = omega if frame_omega is None else frame_omega
    a = annihilate(levels)
    ad = a.conj().T
    H_drift = (omega - frame_omega + detuning) * ad @ a - delta / 2 * ad @ ad @ a @ a
    H_drives = [amp_scale * (a + ad), amp_scale * 1j * (a - ad)]
    return QuantumSystem(2 * np.pi * H_drift, [2 * np.pi * h for h in H_drives])


def sampling_systems(n: int, levels: int = 4, sigma: float = 0.05, seed: int = 1234) -> List[QuantumSystem]:
    rng = np.random.default_rng(seed)
    return [transmon_system(levels, detuning=float(rng.normal(0, sigma)), amp_scale=float(1 + rng.normal(0, 0.02)))
            for _ in range(n)]


def random_hermitian_system(levels: int, n_drives: int, seed: int = 0, scale: float = 1.0) -> QuantumSystem:
    rng = np.random.default_rng(seed)

    def herm():
        M = rng.normal(size=(levels, levels)) + 1j * rng.normal(size=(levels, levels))
        return scale * (M + M.conj().T) / 2

    return QuantumSystem(herm(), [herm() for _ in range(n_drives)])


def _random_unitary(rng, n: int) -> np.ndarray:
    M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    Q, R = np.linalg.qr(M)
    return Q * (np.diag(R) / np.abs(np.diag(R)))


def random_pulse_trajectory(
    systems: Sequence[QuantumSystem],
    T: int,
    dt: float,
    *,
    ket: bool = False,
    n_states: int = 1,
    free_time: bool = True,
    a_bound: float = 1.0,
    seed: int = 1234,
    state_name: str = "Ũ⃗",
) -> NamedTrajectory:
    """Trajectory with component order [state(s)..., a, da, dda, (Δt)] (trajectory_initialization.jl:357-381).

    One state component per system (`<state>_system_k`, unitary_sampling_problem.jl:76-78) when several systems are
    given, or `n_states` ket components (`ψ̃1..`) for the quantum-state problem."""
    rng = np.random.default_rng(seed)
    N = systems[0].levels
    nd = systems[0].n_drives
    comps = {}
    if ket:
        names = [f"ψ̃{k + 1}" for k in range(n_states)] if len(systems) == 1 else [f"ψ̃_system_{k + 1}" for k in range(len(systems))]
    else:
        names = [state_name] if len(systems) == 1 else [f"{state_name}_system_{k + 1}" for k in range(len(systems))]
    for name in names:  # all T random unitaries of a component at once (batched QR of complex Ginibre matrices)
        M = rng.normal(size=(T, N, N)) + 1j * rng.normal(size=(T, N, N))
        Q, R = np.linalg.qr(M)
        dg = np.diagonal(R, axis1=1, axis2=2)
        U = Q * (dg / np.abs(dg))[:, None, :]
        if ket:
            v = np.concatenate([U[:, :, 0].real, U[:, :, 0].imag], axis=1)                      # [Re psi; Im psi]
        else:
            v = np.concatenate([U.real, U.imag], axis=1).transpose(0, 2, 1).reshape(T, 2 * N * N)  # vec(vcat(Re U, Im U))
        comps[name] = np.ascontiguousarray((v + 1e-3 * rng.normal(size=v.shape)).T)
    a = rng.uniform(-a_bound, a_bound, size=(nd, T))
    a[:, 0] = 0.0
    a[:, -1] = 0.0
    comps["a"] = a
    comps["da"] = rng.normal(0, 0.01, size=(nd, T))
    comps["dda"] = rng.normal(0, 0.01, size=(nd, T))
    if free_time:
        comps["Δt"] = rng.uniform(0.5 * dt, 1.5 * dt, size=(1, T))
        return NamedTrajectory(comps, controls=("dda", "Δt"), timestep="Δt")
    return NamedTrajectory(comps, controls=("dda",), timestep=dt)


def build_integrators(systems: Sequence[QuantumSystem], traj: NamedTrajectory, *, integrator: str = "pade",
                      order: int = 4, ket: bool = False):
    """The integrator vector the templates build: one quantum integrator per state component, then the two
    DerivativeIntegrators (unitary_smooth_pulse_problem.jl:163-179, unitary_sampling_problem.jl:134-155)."""
    state_names = [n for n in traj.names if n not in ("a", "da", "dda", "Δt")]
    if len(systems) == 1:
        systems = [systems[0]] * len(state_names)
    out = []
    for name, sys_ in zip(state_names, systems):
        if integrator == "pade":
            cls = QuantumStatePadeIntegrator if ket else UnitaryPadeIntegrator
            out.append(cls(name, "a", sys_, traj, order=order))
        elif integrator == "exponential":
            cls = QuantumStateExponentialIntegrator if ket else UnitaryExponentialIntegrator
            out.append(cls(name, "a", sys_, traj))
        else:
            raise ValueError("integrator must be one of ('pade', 'exponential')")
    out += [DerivativeIntegrator("a", "da", traj), DerivativeIntegrator("da", "dda", traj)]
    return out


def config(name: str, T: Optional[int] = None, integrator: str = "pade", free_time: bool = True, seed: int = 1234,
           n_systems: Optional[int] = None):
    """BASELINE.json configs -> (systems, traj, integrators).  name in {hadamard, cz, sampling, ket}."""
    if name == "hadamard":  # configs[0]: README example shape, N=2, T=50, dt=0.2
        systems, T, dt, ket = [pauli_system(0.0)], T or 50, 0.2, False
    elif name == "cz":  # configs[1..3]: two-transmon CZ, N=9, 4 drives
        systems, T, dt, ket = [two_transmon_cz_system()], T or 200, 1.0, False
    elif name == "sampling":  # configs[4]: robust X gate, 4-level transmon, S sampled systems
        systems, T, dt, ket = sampling_systems(n_systems or 256, levels=4, seed=seed), T or 50, 0.2, False
    elif name == "ket":  # QuantumStateSmoothPulseProblem shape (test_utils.jl:120-136)
        systems, T, dt, ket = [pauli_system(0.1)], T or 50, 0.2, True
    else:
        raise ValueError(name)
    traj = random_pulse_trajectory(systems, T, dt, ket=ket, free_time=free_time, seed=seed,
                                   a_bound=0.1 if name == "sampling" else 1.0)
    return systems, traj, build_integrators(systems, traj, integrator=integrator, ket=ket)


def random_multipliers(n: int, seed: int = 1234) -> np.ndarray:
    return np.random.default_rng(seed + 1).normal(size=n)
