"""CPU oracle for the per-knot-point quantum-dynamics evaluator.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference`` legs may
import this module, and only as the checker.  Nothing under ``quantumcollocation.jl_b200/`` imports it.

PARITY UNPINNED.  The arithmetic of this path lives in un-vendored Julia dependencies that are absent from
``/root/reference`` (QuantumCollocationCore compat "0.3", PiccoloQuantumObjects "0.3", NamedTrajectories "0.2",
ExponentialAction "0.2": ``/root/reference/Project.toml:25-37``; no Manifest, ``.gitignore:24``), and the
reference's own tests hold no numeric golden vector for it.  This file restates the published
equations in the reference's *real isomorphic* arithmetic (G real 2N x 2N, kron(I_N, .) blocks), anchored on

* the equations:                 /root/reference/README.md:74-89,
                                 /root/reference/src/problem_templates/unitary_smooth_pulse_problem.jl:10-30
* the integrator call sites:     unitary_smooth_pulse_problem.jl:163-179, unitary_sampling_problem.jl:134-155,
                                 quantum_state_smooth_pulse_problem.jl:142-196
* the QuantumDynamics surface:   /root/reference/test/scripts/integrator_test_1qubit.jl:36-52
* the iso-vec layout:            /root/reference/src/trajectory_initialization.jl:137, test/test_utils.jl:103
* the knot-vector layout:        trajectory_initialization.jl:357-381, test/test_utils.jl:52-118
* the (values, structure) contract (duplicates sum, square => symmetric upper): test/test_utils.jl:14-27

and is validated in tests/ by finite differences, complex-step, scipy expm/expm_frechet and mpmath.
The CUDA path uses *complex* N x N arithmetic and different algorithms (own Pade-13 scaling-and-squaring with
Frechet jets), so agreement between the two is a genuine cross-check, not an identity.

Intra-knot entry order of both structures = CSC (column-major: by column, then row) order of the per-knot
block pattern -- SURVEY.md §8a6 tags this [DEP-RECALL]; it cannot be verified in this container.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

# ----------------------------------------------------------------------------------------------------------
# isomorphisms  (trajectory_initialization.jl:137 ; test_utils.jl:103 ; ket_to_iso trajectory_initialization.jl:469)
# ----------------------------------------------------------------------------------------------------------


def iso(M: np.ndarray) -> np.ndarray:
    """iso(M) = [Re M  -Im M; Im M  Re M]   (SURVEY §8a1)."""
    M = np.asarray(M)
    return np.block([[M.real, -M.imag], [M.imag, M.real]])


def operator_to_iso_vec(U: np.ndarray) -> np.ndarray:
    """vec(vcat(real(U), imag(U))): column i of U -> [Re U[:,i]; Im U[:,i]]  (trajectory_initialization.jl:137)."""
    U = np.asarray(U, dtype=complex)
    return np.vstack([U.real, U.imag]).reshape(-1, order="F")


def iso_vec_to_operator(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=float)
    N = int(round(math.sqrt(v.size / 2)))
    W = v.reshape(2 * N, N, order="F")
    return W[:N] + 1j * W[N:]


def ket_to_iso(psi: np.ndarray) -> np.ndarray:
    psi = np.asarray(psi, dtype=complex)
    return np.concatenate([psi.real, psi.imag])


def iso_to_ket(v: np.ndarray) -> np.ndarray:
    n = v.size // 2
    return v[:n] + 1j * v[n:]


# ----------------------------------------------------------------------------------------------------------
# QuantumSystem  (README.md:110 ; SURVEY §8a1: G(a) = iso(-i H(a)), dG/da_j = G_j)
# ----------------------------------------------------------------------------------------------------------


class QuantumSystem:
    def __init__(self, H_drift, H_drives):
        H_drives = [np.asarray(h, dtype=complex) for h in H_drives]
        if H_drift is None:
            H_drift = np.zeros_like(H_drives[0])
        self.H_drift = np.asarray(H_drift, dtype=complex)
        self.H_drives = H_drives
        self.levels = self.H_drift.shape[0]
        self.n_drives = len(H_drives)
        self.G_drift = iso(-1j * self.H_drift)
        self.G_drives = [iso(-1j * h) for h in H_drives]

    def H(self, a):
        return self.H_drift + sum(aj * Hj for aj, Hj in zip(a, self.H_drives))

    def G(self, a):
        G = self.G_drift.copy()
        for aj, Gj in zip(a, self.G_drives):
            G = G + aj * Gj
        return G


# ----------------------------------------------------------------------------------------------------------
# trajectory layout (NamedTrajectory: data is dim x T, datavec = vec(data); test_utils.jl:52-118)
# ----------------------------------------------------------------------------------------------------------


@dataclass
class Layout:
    """components: name -> (0-based offset, length) inside one knot vector z_t; dt_name None => fixed timestep."""

    components: Dict[str, Tuple[int, int]]
    T: int
    dt_name: Optional[str] = None
    dt_fixed: float = 0.0
    n_global: int = 0

    @property
    def zdim(self) -> int:
        return max(o + l for o, l in self.components.values())

    def sl(self, name: str) -> slice:
        o, l = self.components[name]
        return slice(o, o + l)


def pade_coefficients(order: int) -> List[float]:
    """c_k = (2m-k)! m! / ((2m)! k! (m-k)!), m = order/2 (order 4: 1, 1/2, 1/12)."""
    m = order // 2
    f = math.factorial
    return [f(2 * m - k) * f(m) / (f(2 * m) * f(k) * f(m - k)) for k in range(m + 1)]


# ----------------------------------------------------------------------------------------------------------
# integrators.  Each exposes, for one knot pair (z_t, z_{t+1}):
#   residual()              -> (dim,)
#   jacobian()              -> dense (dim, 2*zdim) block, columns [z_t ; z_{t+1}]
#   hessian(mu)             -> dense (2*zdim, 2*zdim) UPPER-triangular block of d^2(mu^T f)
#   jac_pattern/hess_pattern-> boolean masks of the structural nonzeros (same shapes)
# ----------------------------------------------------------------------------------------------------------


class _QuantumIntegrator:
    is_unitary = True

    def __init__(self, state_name, control_name, system: QuantumSystem, layout: Layout):
        self.sys = system
        self.layout = layout
        self.state = layout.sl(state_name)
        self.ctrl = layout.sl(control_name)
        self.N = system.levels
        self.n_drives = system.n_drives
        self.free_time = layout.dt_name is not None
        self.dt_idx = layout.components[layout.dt_name][0] if self.free_time else -1
        self.ncols = self.N if self.is_unitary else 1
        self.dim = 2 * self.N * self.ncols
        assert self.state.stop - self.state.start == self.dim, "state component has the wrong length"
        assert self.ctrl.stop - self.ctrl.start == self.n_drives

    # -- helpers -------------------------------------------------------------------------------------------
    def _unpack(self, zt, zt1):
        W0 = zt[self.state].reshape(2 * self.N, self.ncols, order="F")
        W1 = zt1[self.state].reshape(2 * self.N, self.ncols, order="F")
        a = zt[self.ctrl]
        dt = zt[self.dt_idx] if self.free_time else self.layout.dt_fixed
        return W0, W1, a, dt

    def _kron(self, B):
        return np.kron(np.eye(self.ncols), B)

    def _vec(self, W):
        return W.reshape(-1, order="F")

    def _cols(self):
        zdim = self.layout.zdim
        s0 = np.arange(self.state.start, self.state.stop)
        s1 = s0 + zdim
        a = np.arange(self.ctrl.start, self.ctrl.stop)
        return zdim, s0, s1, a

    def _state_block_pattern(self, dense: bool):
        if dense:
            return np.kron(np.eye(self.ncols), np.ones((2 * self.N, 2 * self.N))) > 0
        return np.eye(self.dim) > 0

    # patterns: which state blocks exist is decided by the subclass
    jac_next_dense = True
    hess_next = True

    def jac_pattern(self):
        zdim, s0, s1, a = self._cols()
        P = np.zeros((self.dim, 2 * zdim), dtype=bool)
        P[:, s0] = self._state_block_pattern(True)
        P[:, s1] = self._state_block_pattern(self.jac_next_dense)
        P[:, a] = True
        if self.free_time:
            P[:, self.dt_idx] = True
        return P

    def hess_pattern(self, out=None):
        """Upper-triangular pattern; with `out` the UNFOLDED entries are marked in the caller's matrix instead (the
        caller folds once -- a fold per integrator is quadratic in the ensemble size)."""
        zdim, s0, s1, a = self._cols()
        P = np.zeros((2 * zdim, 2 * zdim), dtype=bool) if out is None else out
        P[np.ix_(s0, a)] = True
        P[np.ix_(a, a)] = True
        if self.hess_next:
            P[np.ix_(a, s1)] = True
        if self.free_time:
            d = self.dt_idx
            P[s0, d] = True
            P[a, d] = True
            P[d, d] = True
            if self.hess_next:
                P[d, s1] = True
        return _to_upper_pattern(P) if out is None else P


def _to_upper_pattern(P):
    """Fold a pattern onto its upper triangle (entry (r,c) with r>c is stored at (c,r); test_utils.jl:22-24)."""
    return np.triu(P | P.T)


def _put_upper(H, rows, cols, block):
    """Add block (len(rows) x len(cols)) into upper-triangular H, mirroring entries that fall below the diagonal.
    A block that straddles the diagonal symmetrically (rows == cols) must be passed already symmetric: only its
    upper triangle is stored."""
    rows = np.atleast_1d(rows)
    cols = np.atleast_1d(cols)
    block = np.asarray(block, dtype=float).reshape(len(rows), len(cols))
    same = len(rows) == len(cols) and np.all(rows == cols)
    for i, r in enumerate(rows):
        for j, c in enumerate(cols):
            if same:
                if r <= c:
                    H[r, c] += block[i, j]
            elif r <= c:
                H[r, c] += block[i, j]
            else:
                H[c, r] += block[i, j]


class _PadeMixin:
    """Implicit Pade residual  (I (x) B) x_{t+1} - (I (x) F) x_t   (SURVEY §8a2)."""

    jac_next_dense = True
    hess_next = True

    def _init_pade(self, order):
        assert order in (4, 6, 8, 10, 12), "pade order must be one of 4, 6, 8, 10, 12"
        self.order = order
        self.c = pade_coefficients(order)
        self.m = order // 2

    def _powers(self, G):
        P = [np.eye(G.shape[0])]
        for _ in range(self.m):
            P.append(P[-1] @ G)
        return P

    def _dGk(self, Gp, Gj, k):
        """d/da_j G^k = sum_i G^i G_j G^(k-1-i)."""
        return sum(Gp[i] @ Gj @ Gp[k - 1 - i] for i in range(k))

    def _d2Gk(self, Gp, Gi, Gj, k):
        """d^2/da_i da_j G^k = sum_{al+be+ga=k-2} G^al (G_i G^be G_j + G_j G^be G_i) G^ga."""
        tot = np.zeros_like(Gi)
        for al in range(k - 1):
            for be in range(k - 1 - al):
                ga = k - 2 - al - be
                tot = tot + Gp[al] @ (Gi @ Gp[be] @ Gj + Gj @ Gp[be] @ Gi) @ Gp[ga]
        return tot

    def _FB(self, Gp, dt, deriv=0):
        """F, B (deriv=0) or their first/second dt-derivatives."""
        F = np.zeros_like(Gp[0])
        B = np.zeros_like(Gp[0])
        for k in range(self.m + 1):
            if deriv == 0:
                w = dt**k
            elif deriv == 1:
                w = k * dt ** (k - 1) if k >= 1 else 0.0
            else:
                w = k * (k - 1) * dt ** (k - 2) if k >= 2 else 0.0
            F = F + self.c[k] * w * Gp[k]
            B = B + (-1) ** k * self.c[k] * w * Gp[k]
        return F, B

    def _dFB(self, Gp, Gj, dt, deriv=0):
        dF = np.zeros_like(Gj)
        dB = np.zeros_like(Gj)
        for k in range(1, self.m + 1):
            w = dt**k if deriv == 0 else k * dt ** (k - 1)
            D = self._dGk(Gp, Gj, k)
            dF = dF + self.c[k] * w * D
            dB = dB + (-1) ** k * self.c[k] * w * D
        return dF, dB

    def _d2FB(self, Gp, Gi, Gj, dt):
        dF = np.zeros_like(Gj)
        dB = np.zeros_like(Gj)
        for k in range(2, self.m + 1):
            D = self._d2Gk(Gp, Gi, Gj, k)
            dF = dF + self.c[k] * dt**k * D
            dB = dB + (-1) ** k * self.c[k] * dt**k * D
        return dF, dB

    def residual(self, zt, zt1):
        W0, W1, a, dt = self._unpack(zt, zt1)
        Gp = self._powers(self.sys.G(a))
        F, B = self._FB(Gp, dt)
        return self._vec(B @ W1 - F @ W0)

    def jacobian(self, zt, zt1):
        W0, W1, a, dt = self._unpack(zt, zt1)
        zdim, s0, s1, ac = self._cols()
        Gp = self._powers(self.sys.G(a))
        F, B = self._FB(Gp, dt)
        J = np.zeros((self.dim, 2 * zdim))
        J[:, s0] = -self._kron(F)
        J[:, s1] = self._kron(B)
        for j, Gj in enumerate(self.sys.G_drives):
            dF, dB = self._dFB(Gp, Gj, dt)
            J[:, ac[j]] = self._vec(dB @ W1 - dF @ W0)
        if self.free_time:
            F1, B1 = self._FB(Gp, dt, deriv=1)
            J[:, self.dt_idx] = self._vec(B1 @ W1 - F1 @ W0)
        return J

    def hessian(self, zt, zt1, mu, out=None):
        W0, W1, a, dt = self._unpack(zt, zt1)
        zdim, s0, s1, ac = self._cols()
        Mu = mu.reshape(2 * self.N, self.ncols, order="F")
        Gp = self._powers(self.sys.G(a))
        H = np.zeros((2 * zdim, 2 * zdim)) if out is None else out  # `out`: add into the caller's matrix
        Gd = self.sys.G_drives
        for j, Gj in enumerate(Gd):
            dF, dB = self._dFB(Gp, Gj, dt)
            _put_upper(H, s0, [ac[j]], -self._vec(dF.T @ Mu))
            _put_upper(H, [ac[j]], s1, self._vec(dB.T @ Mu))
            for i in range(j + 1):
                d2F, d2B = self._d2FB(Gp, Gd[i], Gj, dt)
                _put_upper(H, [ac[i]], [ac[j]], np.sum(Mu * (d2B @ W1 - d2F @ W0)))
            if self.free_time:
                dF1, dB1 = self._dFB(Gp, Gj, dt, deriv=1)
                _put_upper(H, [ac[j]], [self.dt_idx], np.sum(Mu * (dB1 @ W1 - dF1 @ W0)))
        if self.free_time:
            d = self.dt_idx
            F1, B1 = self._FB(Gp, dt, deriv=1)
            F2, B2 = self._FB(Gp, dt, deriv=2)
            _put_upper(H, s0, [d], -self._vec(F1.T @ Mu))
            _put_upper(H, [d], s1, self._vec(B1.T @ Mu))
            _put_upper(H, [d], [d], np.sum(Mu * (B2 @ W1 - F2 @ W0)))
        return H


class _ExpMixin:
    """Explicit exponential residual  x_{t+1} - (I (x) exp(dt G(a))) x_t  (README.md:79)."""

    jac_next_dense = False
    hess_next = False

    def residual(self, zt, zt1):
        W0, W1, a, dt = self._unpack(zt, zt1)
        E = sla.expm(dt * self.sys.G(a))
        return self._vec(W1 - E @ W0)

    def jacobian(self, zt, zt1):
        W0, W1, a, dt = self._unpack(zt, zt1)
        zdim, s0, s1, ac = self._cols()
        G = self.sys.G(a)
        E = sla.expm(dt * G)
        J = np.zeros((self.dim, 2 * zdim))
        J[:, s0] = -self._kron(E)
        J[:, s1] = np.eye(self.dim)
        for j, Gj in enumerate(self.sys.G_drives):
            L = sla.expm_frechet(dt * G, dt * Gj, compute_expm=False)
            J[:, ac[j]] = -self._vec(L @ W0)
        if self.free_time:
            J[:, self.dt_idx] = -self._vec(G @ E @ W0)
        return J

    @staticmethod
    def _expm_d2(X, E1, E2):
        """Second Frechet derivative d^2/ds dt exp(X + s E1 + t E2) at 0 via a 3x3 block-triangular exponential."""
        n = X.shape[0]
        Z = np.zeros((n, n))

        def top_right(A, B):
            M = np.block([[X, A, Z], [Z, X, B], [Z, Z, X]])
            return sla.expm(M)[:n, 2 * n :]

        return top_right(E1, E2) + top_right(E2, E1)

    def hessian(self, zt, zt1, mu, out=None):
        W0, W1, a, dt = self._unpack(zt, zt1)
        zdim, s0, s1, ac = self._cols()
        Mu = mu.reshape(2 * self.N, self.ncols, order="F")
        G = self.sys.G(a)
        E = sla.expm(dt * G)
        H = np.zeros((2 * zdim, 2 * zdim)) if out is None else out  # `out`: add into the caller's matrix
        Gd = self.sys.G_drives
        Ls = [sla.expm_frechet(dt * G, dt * Gj, compute_expm=False) for Gj in Gd]
        for j, Gj in enumerate(Gd):
            _put_upper(H, s0, [ac[j]], -self._vec(Ls[j].T @ Mu))
            for i in range(j + 1):
                L2 = self._expm_d2(dt * G, dt * Gd[i], dt * Gj)
                _put_upper(H, [ac[i]], [ac[j]], -np.sum(Mu * (L2 @ W0)))
            if self.free_time:
                _put_upper(H, [ac[j]], [self.dt_idx], -np.sum(Mu * ((Gj @ E + G @ Ls[j]) @ W0)))
        if self.free_time:
            d = self.dt_idx
            _put_upper(H, s0, [d], -self._vec((G @ E).T @ Mu))
            _put_upper(H, [d], [d], -np.sum(Mu * (G @ G @ E @ W0)))
        return H


class UnitaryPadeIntegrator(_PadeMixin, _QuantumIntegrator):
    """unitary_smooth_pulse_problem.jl:164-167 ; unitary_sampling_problem.jl:137-140."""

    is_unitary = True

    def __init__(self, state_name, control_name, system, layout, order=4):
        super().__init__(state_name, control_name, system, layout)
        self._init_pade(order)


class QuantumStatePadeIntegrator(_PadeMixin, _QuantumIntegrator):
    """quantum_state_smooth_pulse_problem.jl:145-166."""

    is_unitary = False

    def __init__(self, state_name, control_name, system, layout, order=4):
        super().__init__(state_name, control_name, system, layout)
        self._init_pade(order)


class UnitaryExponentialIntegrator(_ExpMixin, _QuantumIntegrator):
    """unitary_smooth_pulse_problem.jl:168-170 ; unitary_sampling_problem.jl:141-145."""

    is_unitary = True


class QuantumStateExponentialIntegrator(_ExpMixin, _QuantumIntegrator):
    """quantum_state_smooth_pulse_problem.jl:167-189."""

    is_unitary = False


class DerivativeIntegrator:
    """x_{t+1} - x_t - dt_t * dx_t   (unitary_smooth_pulse_problem.jl:15-16,177-178 )."""

    def __init__(self, x_name, dx_name, layout: Layout):
        self.layout = layout
        self.x = layout.sl(x_name)
        self.dx = layout.sl(dx_name)
        self.dim = self.x.stop - self.x.start
        assert self.dx.stop - self.dx.start == self.dim
        self.free_time = layout.dt_name is not None
        self.dt_idx = layout.components[layout.dt_name][0] if self.free_time else -1

    def _dt(self, zt):
        return zt[self.dt_idx] if self.free_time else self.layout.dt_fixed

    def residual(self, zt, zt1):
        return zt1[self.x] - zt[self.x] - self._dt(zt) * zt[self.dx]

    def jacobian(self, zt, zt1):
        zdim = self.layout.zdim
        J = np.zeros((self.dim, 2 * zdim))
        r = np.arange(self.dim)
        J[r, self.x.start + r] = -1.0
        J[r, zdim + self.x.start + r] = 1.0
        J[r, self.dx.start + r] = -self._dt(zt)
        if self.free_time:
            J[:, self.dt_idx] = -zt[self.dx]
        return J

    def jac_pattern(self):
        zdim = self.layout.zdim
        P = np.zeros((self.dim, 2 * zdim), dtype=bool)
        r = np.arange(self.dim)
        P[r, self.x.start + r] = True
        P[r, zdim + self.x.start + r] = True
        P[r, self.dx.start + r] = True
        if self.free_time:
            P[:, self.dt_idx] = True
        return P

    def hessian(self, zt, zt1, mu, out=None):
        zdim = self.layout.zdim
        H = np.zeros((2 * zdim, 2 * zdim)) if out is None else out  # `out`: add into the caller's matrix
        if self.free_time:
            for j in range(self.dim):
                _put_upper(H, [self.dx.start + j], [self.dt_idx], -mu[j])
        return H

    def hess_pattern(self, out=None):
        zdim = self.layout.zdim
        P = np.zeros((2 * zdim, 2 * zdim), dtype=bool) if out is None else out
        if self.free_time:
            P[np.arange(self.dx.start, self.dx.stop), self.dt_idx] = True
        return _to_upper_pattern(P) if out is None else P


# ----------------------------------------------------------------------------------------------------------
# QuantumDynamics  (integrator_test_1qubit.jl:41-52 )
# ----------------------------------------------------------------------------------------------------------


class QuantumDynamics:
    """Stacks integrators over the knots: F, dF (values), dF_structure, mu_d2F (values), mu_d2F_structure.

    Structures are lists of 1-based (row, col) tuples like the reference's Vector{Tuple{Int,Int}}.
    """

    def __init__(self, integrators: Sequence, layout: Layout, eval_hessian: bool = True, structure_order: str = "csc"):
        """structure_order: intra-knot order of the entries -- the Core's order is [DEP-RECALL], so every candidate is a policy:
        "csc" (union pattern, column-major), "row_major" (union pattern, row-major) or "per_integrator" (the integrators'
        own entry lists one after the other, column-major inside each; a Hessian position two integrators touch appears twice
        and the consumer sums the duplicates, test/test_utils.jl:14-27)."""
        assert structure_order in ("csc", "row_major", "per_integrator")
        self.structure_order = structure_order
        self.integrators = list(integrators)
        self.layout = layout
        self.zdim = layout.zdim
        self.T = layout.T
        self.row_off = np.cumsum([0] + [I.dim for I in self.integrators])
        self.dyn = int(self.row_off[-1])
        zdim = self.zdim

        def walk(P):  # entries of a boolean pattern in this policy's order (per_integrator: column-major inside the integrator)
            if structure_order == "row_major":
                rows, cols = np.nonzero(P)
            else:
                cols, rows = np.nonzero(P.T)
            return list(zip(rows.tolist(), cols.tolist()))

        self.eval_hessian = eval_hessian
        self.hess_owner = None  # per_integrator: which integrator's contribution an entry holds
        if structure_order == "per_integrator":
            self.jac_knot = []
            for I, r0 in zip(self.integrators, self.row_off):
                JP = np.zeros((self.dyn, 2 * zdim), dtype=bool)
                JP[r0 : r0 + I.dim] = I.jac_pattern()
                self.jac_knot += walk(JP)
            self.hess_knot, self.hess_owner = [], []
            if eval_hessian:
                for q, I in enumerate(self.integrators):
                    HP = np.zeros((2 * zdim, 2 * zdim), dtype=bool)
                    I.hess_pattern(out=HP)
                    e = walk(_to_upper_pattern(HP))
                    self.hess_knot += e
                    self.hess_owner += [q] * len(e)
        else:
            JP = np.zeros((self.dyn, 2 * zdim), dtype=bool)
            for I, r0 in zip(self.integrators, self.row_off):
                JP[r0 : r0 + I.dim] |= I.jac_pattern()
            self.jac_knot = walk(JP)
            if eval_hessian:
                HP = np.zeros((2 * zdim, 2 * zdim), dtype=bool)
                for I in self.integrators:
                    I.hess_pattern(out=HP)  # unfolded marks
                self.hess_knot = walk(_to_upper_pattern(HP))
            else:
                self.hess_knot = []
        self.nnzJ = len(self.jac_knot)
        self.nnzH = len(self.hess_knot)
        self.dF_structure = [
            (r + t * self.dyn + 1, c + t * zdim + 1) for t in range(self.T - 1) for (r, c) in self.jac_knot
        ]
        self.mu_d2F_structure = [
            (r + t * zdim + 1, c + t * zdim + 1) for t in range(self.T - 1) for (r, c) in self.hess_knot
        ]

    def _knots(self, Z):
        zdim = self.zdim
        for t in range(self.T - 1):
            yield t, Z[t * zdim : (t + 1) * zdim], Z[(t + 1) * zdim : (t + 2) * zdim]

    def F(self, Z):
        Z = np.asarray(Z, dtype=float)
        out = np.zeros(self.dyn * (self.T - 1))
        for t, zt, zt1 in self._knots(Z):
            for I, r0 in zip(self.integrators, self.row_off):
                out[t * self.dyn + r0 : t * self.dyn + r0 + I.dim] = I.residual(zt, zt1)
        return out

    def dF(self, Z):
        Z = np.asarray(Z, dtype=float)
        out = np.zeros(self.nnzJ * (self.T - 1))
        rr = np.array([r for r, _ in self.jac_knot])
        cc = np.array([c for _, c in self.jac_knot])
        for t, zt, zt1 in self._knots(Z):
            J = np.zeros((self.dyn, 2 * self.zdim))
            for I, r0 in zip(self.integrators, self.row_off):
                J[r0 : r0 + I.dim] = I.jacobian(zt, zt1)
            out[t * self.nnzJ : (t + 1) * self.nnzJ] = J[rr, cc]
        return out

    def mu_d2F(self, Z, mu):
        assert self.eval_hessian
        Z = np.asarray(Z, dtype=float)
        mu = np.asarray(mu, dtype=float)
        out = np.zeros(self.nnzH * (self.T - 1))
        rr = np.array([r for r, _ in self.hess_knot])
        cc = np.array([c for _, c in self.hess_knot])
        H = np.zeros((2 * self.zdim, 2 * self.zdim))
        owner = None if self.hess_owner is None else np.array(self.hess_owner)
        for t, zt, zt1 in self._knots(Z):
            mut = mu[t * self.dyn : (t + 1) * self.dyn]
            if owner is not None:  # per_integrator: every integrator's own contribution, never summed
                blk = out[t * self.nnzH : (t + 1) * self.nnzH]
                for q, (I, r0) in enumerate(zip(self.integrators, self.row_off)):
                    sel = owner == q
                    I.hessian(zt, zt1, mut[r0 : r0 + I.dim], out=H)
                    blk[sel] = H[rr[sel], cc[sel]]
                    H[rr[sel], cc[sel]] = 0.0
                continue
            for I, r0 in zip(self.integrators, self.row_off):
                I.hessian(zt, zt1, mut[r0 : r0 + I.dim], out=H)  # integrators add in order, like the dense sum did
            out[t * self.nnzH : (t + 1) * self.nnzH] = H[rr, cc]
            H[rr, cc] = 0.0  # every entry an integrator touches is inside the pattern
        return out


def dense(vals, structure, shape):
    """test/test_utils.jl:14-27: rebuild a matrix from (values, structure); duplicates sum; square => symmetric."""
    M = np.zeros(shape)
    for v, (k, j) in zip(vals, structure):
        M[k - 1, j - 1] += v
    if shape[0] == shape[1]:
        return np.triu(M) + np.triu(M, 1).T
    return M
