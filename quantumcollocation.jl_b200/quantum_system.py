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
