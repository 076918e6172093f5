"""Objective and terminal-constraint descriptors with the reference's constructor shapes.

  QuadraticRegularizer(name, traj, R; timestep_name)               unitary_smooth_pulse_problem.jl:151-153
  UnitaryInfidelityObjective(state_name, traj, Q; subspace)         unitary_smooth_pulse_problem.jl:132-137
  MinimumTimeObjective(traj; D)                                     unitary_minimum_time_problem.jl:67-69
  FinalUnitaryFidelityConstraint(state_name, val, traj; subspace)   unitary_minimum_time_problem.jl:80-84

They carry no arithmetic: `QuantumDynamics.attach_objective(J)` hands them to libqcknot.so, which evaluates value, gradient and
Hessian values on the GPU from the same device-resident Z as the dynamics.  Objectives add with `+` like the reference's."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .trajectory import NamedTrajectory


class Objective:
    def __init__(self, terms: Optional[List["_Term"]] = None):
        self.terms: List[_Term] = list(terms or [])

    def __add__(self, other: "Objective") -> "Objective":
        return Objective(self.terms + other.terms)

    def __rmul__(self, w: float) -> "Objective":
        out = Objective([t.scaled(float(w)) for t in self.terms])
        return out


class _Term(Objective):
    kind = -1

    def __init__(self):
        super().__init__([self])
        self.weight = 1.0
        self.comp = range(0)
        self.levels = 0
        self.R = None
        self.goal = None
        self.n_sub = 0

    def scaled(self, w):
        import copy
        t = copy.copy(self)
        t.terms = [t]
        t.weight = self.weight * w
        return t


class QuadraticRegularizer(_Term):
    kind = _lib.QCK_OBJ_QUADRATIC_REGULARIZER

    def __init__(self, name: str, traj: NamedTrajectory, R, timestep_name: Optional[str] = None):
        super().__init__()
        if name not in traj.components:
            raise KeyError(f"component {name!r} not in trajectory")
        self.name, self.comp = name, traj.components[name]
        self.R = np.ascontiguousarray(np.broadcast_to(np.asarray(R, dtype=np.float64), (len(self.comp),)))


class MinimumTimeObjective(_Term):
    kind = _lib.QCK_OBJ_MINIMUM_TIME

    def __init__(self, traj: NamedTrajectory, D: float = 1.0):
        super().__init__()
        if not traj.free_time:
            raise ValueError("MinimumTimeObjective needs a free timestep")
        self.comp = traj.components[traj.timestep]
        self.weight = float(D)


def _goal_iso(goal, N, subspace):
    from .isomorphisms import operator_to_iso_vec
    goal = np.asarray(goal)
    G = goal if goal.ndim == 2 else None
    if G is None:
        return np.ascontiguousarray(goal, dtype=np.float64), (len(subspace) if subspace is not None else N)
    if subspace is not None:
        mask = np.zeros((N, N))
        mask[np.ix_(list(subspace), list(subspace))] = 1.0
        G = G * mask
    return operator_to_iso_vec(G), (len(subspace) if subspace is not None else N)


class UnitaryInfidelityObjective(_Term):
    kind = _lib.QCK_OBJ_UNITARY_INFIDELITY

    def __init__(self, state_name: str, traj: NamedTrajectory, Q: float = 100.0, subspace: Optional[Sequence[int]] = None, goal=None):
        super().__init__()
        if state_name not in traj.components:
            raise KeyError(f"state component {state_name!r} not in trajectory")
        self.state_name, self.comp = state_name, traj.components[state_name]
        self.levels = int(round(np.sqrt(len(self.comp) / 2)))
        g = goal if goal is not None else traj.goal.get(state_name)
        if g is None:
            raise ValueError("no goal: pass goal=... (operator or iso-vec) or a trajectory with goal[state_name]")
        self.goal, self.n_sub = _goal_iso(g, self.levels, subspace)
        self.weight = float(Q)


class FinalUnitaryFidelityConstraint(UnitaryInfidelityObjective):
    """g(Z) = F(U_T) - val >= 0."""

    def __init__(self, state_name: str, val: float, traj: NamedTrajectory, subspace=None, goal=None):
        super().__init__(state_name, traj, 1.0, subspace, goal)
        self.val = float(val)


def term_array(J: Objective):
    """ctypes array of qck_objective_term + the arrays it points into (keep alive)."""
    arr = (_lib.ObjectiveTerm * len(J.terms))()
    keep = []
    for d, t in zip(arr, J.terms):
        d.kind, d.comp_off, d.comp_len, d.levels, d.weight, d.n_sub = t.kind, t.comp.start, len(t.comp), t.levels, t.weight, t.n_sub
        if t.R is not None:
            keep.append(t.R)
            d.R = t.R.ctypes.data_as(C.POINTER(C.c_double))
        if t.goal is not None:
            g = np.ascontiguousarray(t.goal, dtype=np.float64)
            keep.append(g)
            d.goal = g.ctypes.data_as(C.POINTER(C.c_double))
    return arr, keep
