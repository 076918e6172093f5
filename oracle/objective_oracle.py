"""CPU restatement (numpy) of the objective / terminal-constraint terms the problem templates add next to the dynamics
(SURVEY.md section 8f, row f1).  TEST INFRASTRUCTURE ONLY: imported by tests/ (and smoke()); never by the product.

PARITY UNPINNED.  The terms live in QuantumCollocationCore 0.3 (un-vendored, /root/reference/Project.toml:31); this file
restates them from their call sites and docstrings in /root/reference and from the 0.3-era definitions [DEP-RECALL]:

  QuadraticRegularizer(name, traj, R; timestep_name)      unitary_smooth_pulse_problem.jl:151-153
      J = sum_t 1/2 r_t' (R .* r_t),  r_t = dt_t * v_t  (dt_t = the knot's timestep variable, or the fixed timestep)
  UnitaryInfidelityObjective(state_name, traj, Q; subspace)    unitary_smooth_pulse_problem.jl:132-137
      J = Q * (1 - |tr(U_goal' U_T)|^2 / n^2) on the LAST knot, n = dimension of the (sub)space, entries outside the subspace
      dropped from the trace                                    (fidelity: unitary_smooth_pulse_problem.jl:218-220)
  MinimumTimeObjective(traj; D)                                 unitary_minimum_time_problem.jl:67-69
      J = D * sum_{t < T} dt_t
  FinalUnitaryFidelityConstraint(state_name, val, traj; subspace)   unitary_minimum_time_problem.jl:80-84
      g(Z) = F(U_T) - val >= 0

Each term gives value, dense gradient (length zdim*T) and upper-triangular Hessian entries; the objective's Hessian structure
is knot-major: for every knot t the per-knot terms' entries in term order, then the terminal terms' dense blocks on the final
knot (duplicates are summed by the consumer like every (values, structure) pair of the reference, test/test_utils.jl:14-27).
"""
from __future__ import annotations

import numpy as np


def _fid_vectors(goal_iso: np.ndarray, N: int):
    """a = g.u and b = w.u give tr(G' U) = a + i b for iso-vectors u (trajectory_initialization.jl:137 layout)."""
    g = np.asarray(goal_iso, dtype=float)
    w = np.empty_like(g)
    G = g.reshape(N, 2 * N)  # row c = column c of the operator: [Re (N) ; Im (N)]
    W = w.reshape(N, 2 * N)
    W[:, :N] = -G[:, N:]
    W[:, N:] = G[:, :N]
    return g, w


class QuadraticRegularizer:
    def __init__(self, comp, R, dt_off, dt_fixed=0.0):
        self.comp = comp  # range of the component inside z_t
        self.R = np.broadcast_to(np.asarray(R, dtype=float), (len(comp),)).copy()
        self.dt_off, self.dt_fixed = dt_off, dt_fixed

    def _dt(self, z):
        return z[self.dt_off] if self.dt_off >= 0 else self.dt_fixed

    def value(self, Z, T, zdim):
        Zm = Z[: T * zdim].reshape(T, zdim)
        return float(sum(0.5 * self._dt(z) ** 2 * np.dot(self.R, z[self.comp.start:self.comp.stop] ** 2) for z in Zm))

    def gradient(self, Z, T, zdim, g):
        G = g.reshape(T, zdim)
        Zm = Z[: T * zdim].reshape(T, zdim)
        for t in range(T):
            v, dt = Zm[t, self.comp.start:self.comp.stop], self._dt(Zm[t])
            G[t, self.comp.start:self.comp.stop] += dt * dt * self.R * v
            if self.dt_off >= 0:
                G[t, self.dt_off] += dt * np.dot(self.R, v * v)

    def hessian_entries(self, Z, T, zdim, t):
        """[(row, col, value)] of knot t, 0-based inside the whole variable vector, row <= col."""
        z = Z[t * zdim:(t + 1) * zdim]
        v, dt = z[self.comp.start:self.comp.stop], self._dt(z)
        out = [(t * zdim + self.comp.start + i, t * zdim + self.comp.start + i, dt * dt * self.R[i]) for i in range(len(v))]
        if self.dt_off >= 0:
            for i in range(len(v)):
                r, c = t * zdim + self.comp.start + i, t * zdim + self.dt_off
                out.append((min(r, c), max(r, c), 2.0 * dt * self.R[i] * v[i]))
            out.append((t * zdim + self.dt_off, t * zdim + self.dt_off, float(np.dot(self.R, v * v))))
        return out


class MinimumTimeObjective:
    def __init__(self, dt_off, D=1.0):
        self.dt_off, self.D = dt_off, D

    def value(self, Z, T, zdim):
        return float(self.D * Z[: T * zdim].reshape(T, zdim)[: T - 1, self.dt_off].sum())

    def gradient(self, Z, T, zdim, g):
        g.reshape(T, zdim)[: T - 1, self.dt_off] += self.D

    def hessian_entries(self, Z, T, zdim, t):
        return []


class UnitaryInfidelityObjective:
    """Q * (1 - F(U_T)); `goal_iso` has zeros outside the subspace, n_sub = dimension of the subspace."""

    def __init__(self, comp, goal_iso, N, Q=100.0, n_sub=None):
        self.comp, self.N, self.Q = comp, N, Q
        self.n_sub = n_sub or N
        self.g, self.w = _fid_vectors(goal_iso, N)

    def fidelity(self, u):
        a, b = np.dot(self.g, u), np.dot(self.w, u)
        return (a * a + b * b) / self.n_sub ** 2

    def value(self, Z, T, zdim):
        u = Z[(T - 1) * zdim + self.comp.start:(T - 1) * zdim + self.comp.stop]
        return float(self.Q * (1.0 - self.fidelity(u)))

    def fid_gradient(self, u):
        a, b = np.dot(self.g, u), np.dot(self.w, u)
        return 2.0 * (a * self.g + b * self.w) / self.n_sub ** 2

    def fid_hessian(self):
        return 2.0 * (np.outer(self.g, self.g) + np.outer(self.w, self.w)) / self.n_sub ** 2

    def gradient(self, Z, T, zdim, g):
        o = (T - 1) * zdim
        g[o + self.comp.start:o + self.comp.stop] -= self.Q * self.fid_gradient(Z[o + self.comp.start:o + self.comp.stop])

    def hessian_entries(self, Z, T, zdim, t):
        if t != T - 1:
            return []
        Hm, o = -self.Q * self.fid_hessian(), (T - 1) * zdim + self.comp.start
        n = len(self.g)
        return [(o + i, o + j, Hm[i, j]) for j in range(n) for i in range(j + 1)]  # upper triangle, by column


class FinalUnitaryFidelityConstraint:
    """g(Z) = F(U_T) - val >= 0; Jacobian = one row over the last knot's state component; Hessian of mu * g."""

    def __init__(self, comp, goal_iso, N, val, n_sub=None):
        self.obj = UnitaryInfidelityObjective(comp, goal_iso, N, 1.0, n_sub)
        self.comp, self.val = comp, val

    def value(self, Z, T, zdim):
        u = Z[(T - 1) * zdim + self.comp.start:(T - 1) * zdim + self.comp.stop]
        return float(self.obj.fidelity(u) - self.val)

    def jacobian(self, Z, T, zdim):
        u = Z[(T - 1) * zdim + self.comp.start:(T - 1) * zdim + self.comp.stop]
        return self.obj.fid_gradient(u)

    def jacobian_columns(self, T, zdim):
        return np.arange((T - 1) * zdim + self.comp.start, (T - 1) * zdim + self.comp.stop) + 1  # 1-based

    def hessian(self, mu):
        Hm, n = mu * self.obj.fid_hessian(), len(self.obj.g)
        return np.array([Hm[i, j] for j in range(n) for i in range(j + 1)])


class Objective:
    """J = sum of terms (the reference's `+` on objectives)."""

    def __init__(self, terms, T, zdim):
        self.terms, self.T, self.zdim = list(terms), T, zdim

    def value(self, Z):
        return float(sum(term.value(Z, self.T, self.zdim) for term in self.terms))

    def gradient(self, Z):
        g = np.zeros(self.T * self.zdim)
        for term in self.terms:
            term.gradient(Z, self.T, self.zdim, g)
        return g

    def _entries(self, Z):
        out = []
        for t in range(self.T):  # per-knot terms, knot-major
            for term in self.terms:
                if not isinstance(term, UnitaryInfidelityObjective):
                    out += term.hessian_entries(Z, self.T, self.zdim, t)
        for term in self.terms:  # terminal terms after the final knot
            if isinstance(term, UnitaryInfidelityObjective):
                out += term.hessian_entries(Z, self.T, self.zdim, self.T - 1)
        return out

    def hessian_structure(self):
        Z0 = np.ones(self.T * self.zdim)
        return np.array([(r + 1, c + 1) for r, c, _ in self._entries(Z0)], dtype=np.int64).reshape(-1, 2)

    def hessian(self, Z, sigma=1.0):
        return sigma * np.array([v for _, _, v in self._entries(Z)])
