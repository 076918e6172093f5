"""Isomorphism helpers with the reference's names and layouts (PiccoloQuantumObjects; call sites
/root/reference/src/trajectory_initialization.jl:137,469-470 ; fixture test/test_utils.jl:103)."""
from __future__ import annotations

import math

import numpy as np


def operator_to_iso_vec(U) -> np.ndarray:
    """vec(vcat(real(U), imag(U))): column i of U occupies [i*2N, (i+1)*2N) as [Re U[:, i]; Im U[:, i]]."""
    U = np.asarray(U, dtype=np.complex128)
    return np.ascontiguousarray(np.vstack([U.real, U.imag]).reshape(-1, order="F"))


def iso_vec_to_operator(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    N = int(round(math.sqrt(v.size / 2)))
    if 2 * N * N != v.size:
        raise ValueError("not a unitary iso-vec length")
    W = v.reshape(2 * N, N, order="F")
    return W[:N] + 1j * W[N:]


def ket_to_iso(psi) -> np.ndarray:
    psi = np.asarray(psi, dtype=np.complex128)
    return np.concatenate([psi.real, psi.imag])


def iso_to_ket(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    n = v.size // 2
    return v[:n] + 1j * v[n:]
