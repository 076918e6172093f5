"""Integrator descriptors with the reference's constructor shapes.

  UnitaryPadeIntegrator(state_name, control_name, system, traj; order)        unitary_smooth_pulse_problem.jl:164-167
  UnitaryExponentialIntegrator(state_name, control_name, system, traj)        unitary_smooth_pulse_problem.jl:168-170
  QuantumStatePadeIntegrator / QuantumStateExponentialIntegrator(...)         quantum_state_smooth_pulse_problem.jl:145-189
  DerivativeIntegrator(x_name, dx_name, traj)                                 unitary_smooth_pulse_problem.jl:177-178

They carry no arithmetic: QuantumDynamics hands them to libqcknot.so, which evaluates them on the GPU."""
from __future__ import annotations

from . import _lib
from .quantum_system import QuantumSystem
from .trajectory import NamedTrajectory


class AbstractIntegrator:
    kind: int = -1
    order: int = 0


class _QuantumIntegrator(AbstractIntegrator):
    unitary = True

    def __init__(self, state_name: str, control_name: str, system: QuantumSystem, traj: NamedTrajectory):
        if state_name not in traj.components:
            raise KeyError(f"state component {state_name!r} not in trajectory")
        if control_name not in traj.components:
            raise KeyError(f"control component {control_name!r} not in trajectory")
        self.state_name, self.control_name, self.system = state_name, control_name, system
        self.state_components = traj.components[state_name]
        self.drive_components = traj.components[control_name]
        N = system.levels
        want = 2 * N * N if self.unitary else 2 * N
        if len(self.state_components) != want:
            raise ValueError(f"{state_name} has {len(self.state_components)} rows, expected {want} for {N} levels")
        if len(self.drive_components) != system.n_drives:
            raise ValueError(f"{control_name} has {len(self.drive_components)} rows, system has {system.n_drives} drives")
        self.dim = want
        self.freetime = traj.free_time


class UnitaryPadeIntegrator(_QuantumIntegrator):
    kind = _lib.QCK_UNITARY_PADE

    def __init__(self, state_name, control_name, system, traj, order: int = 4):
        super().__init__(state_name, control_name, system, traj)
        self.order = int(order)


class UnitaryExponentialIntegrator(_QuantumIntegrator):
    kind = _lib.QCK_UNITARY_EXP


class QuantumStatePadeIntegrator(_QuantumIntegrator):
    kind = _lib.QCK_KET_PADE
    unitary = False

    def __init__(self, state_name, control_name, system, traj, order: int = 4):
        super().__init__(state_name, control_name, system, traj)
        self.order = int(order)


class QuantumStateExponentialIntegrator(_QuantumIntegrator):
    kind = _lib.QCK_KET_EXP
    unitary = False


class DensityOperatorExponentialIntegrator(QuantumStateExponentialIntegrator):
    """density_operator_smooth_pulse_problem.jl:104-106: vec(rho)~_{t+1} = exp(dt G(a_t)) vec(rho)~_t with the Lindbladian of an
    `OpenQuantumSystem` -- the ket exponential integrator on N^2 levels (state = iso-vec [Re vec(rho); Im vec(rho)])."""


class DerivativeIntegrator(AbstractIntegrator):
    kind = _lib.QCK_DERIVATIVE

    def __init__(self, x_name: str, dx_name: str, traj: NamedTrajectory):
        for n in (x_name, dx_name):
            if n not in traj.components:
                raise KeyError(f"component {n!r} not in trajectory")
        self.x_name, self.dx_name = x_name, dx_name
        self.x_components = traj.components[x_name]
        self.dx_components = traj.components[dx_name]
        if len(self.x_components) != len(self.dx_components):
            raise ValueError("x and dx must have the same number of rows")
        self.dim = len(self.x_components)
        self.freetime = traj.free_time
