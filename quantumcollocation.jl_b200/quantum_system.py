"""QuantumSystem(H_drift, H_drives) -- the data definition the kernels must match (SURVEY.md section 8a1;
call sites /root/reference/README.md:110, src/problem_templates/unitary_smooth_pulse_problem.jl:199)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


class QuantumSystem:
    """H(a) = H_drift + sum_j a_j H_drives[j];  G(a) = iso(-i H(a)).

    `QuantumSystem(H_drives)` (drift-free, README.md:110) and `QuantumSystem(H_drift, H_drives)` are both accepted.
    """

    def __init__(self, H_drift, H_drives: Optional[Sequence] = None, params: Optional[dict] = None):
        if H_drives is None:  # QuantumSystem([X, Y])
            H_drives = H_drift
            H_drift = None
        H_drives = [np.array(h, dtype=np.complex128) for h in H_drives]
        if H_drift is None:
            if not H_drives:
                raise ValueError("a system needs a drift or at least one drive")
            H_drift = np.zeros_like(H_drives[0])
        self.H_drift = np.array(H_drift, dtype=np.complex128)
        if self.H_drift.ndim != 2 or self.H_drift.shape[0] != self.H_drift.shape[1]:
            raise ValueError("H_drift must be square")
        for h in H_drives:
            if h.shape != self.H_drift.shape:
                raise ValueError("all drives must have the drift's shape")
        self.H_drives = H_drives
        self.levels = self.H_drift.shape[0]
        self.n_drives = len(H_drives)
        self.params = params or {}

    def H(self, a) -> np.ndarray:
        out = self.H_drift.copy()
        for aj, Hj in zip(a, self.H_drives):
            out = out + aj * Hj
        return out

    def G(self, a) -> np.ndarray:
        M = -1j * self.H(a)
        return np.block([[M.real, -M.imag], [M.imag, M.real]])

    # --- flat Float64 views handed to the C-ABI (Julia: reinterpret(Float64, H), column-major) -----------------
    def drift_reim(self) -> np.ndarray:
        return np.ascontiguousarray(self.H_drift.reshape(-1, order="F")).view(np.float64).copy()

    def drives_reim(self) -> np.ndarray:
        if not self.H_drives:
            return np.zeros(0)
        flat = np.concatenate([h.reshape(-1, order="F") for h in self.H_drives])
        return np.ascontiguousarray(flat).view(np.float64).copy()


class OpenQuantumSystem(QuantumSystem):
    """Lindblad generator as a (non-Hermitian) `QuantumSystem` on the vectorised density operator (SURVEY.md 8f row f2;
    `DensityOperatorExponentialIntegrator`, density_operator_smooth_pulse_problem.jl:104-106; system template
    src/quantum_system_templates/cats.jl).

    d vec(rho)/dt = L(a) vec(rho),  L(a) = L_0 + sum_j a_j L_j with (column-major vec, vec(A rho B) = (B^T (x) A) vec rho)
        L_0 = -i (I (x) H_0 - H_0^T (x) I) + sum_k [conj(C_k) (x) C_k - 1/2 I (x) C_k' C_k - 1/2 (C_k' C_k)^T (x) I]
        L_j = -i (I (x) H_j - H_j^T (x) I)
    The kernels evaluate G(a) = iso(-i H(a)); handing them "H" := i L makes -i H = L, so the density-operator integrator is
    the ket exponential integrator on N^2 levels with this system."""

    def __init__(self, H_drift, H_drives, dissipation_operators=()):
        H_drift = np.array(H_drift, dtype=np.complex128)
        n = H_drift.shape[0]
        eye = np.eye(n)

        def comm(H):
            return -1j * (np.kron(eye, H) - np.kron(H.T, eye))

        L0 = comm(H_drift)
        for Ck in dissipation_operators:
            Ck = np.array(Ck, dtype=np.complex128)
            CdC = Ck.conj().T @ Ck
            L0 = L0 + np.kron(Ck.conj(), Ck) - 0.5 * np.kron(eye, CdC) - 0.5 * np.kron(CdC.T, eye)
        super().__init__(1j * L0, [1j * comm(np.array(h, dtype=np.complex128)) for h in H_drives])
        self.hilbert_levels = n
