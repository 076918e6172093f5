"""Generates the committed golden fixtures from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

Input fixture: the reference's own literal trajectory `named_trajectory_type_1` (Hadamard geodesic, 15 x 5,
/root/reference/test/test_utils.jl:52-118) with the system of `smooth_unitary_problem`
(QuantumSystem(0.1 Z, [X, Y]), test_utils.jl:139-141).  The reference holds NO output vectors for this path
(SURVEY.md section 8c), so the outputs stored here come from oracle/knot_oracle.py ("parity unpinned" vs the Julia Core);
they pin the oracle + CUDA path against regressions and are cross-checked by the mpmath tests in tests/test_oracle.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import knot_oracle as ko  # noqa: E402

# test/test_utils.jl:55-71 (rows 1:8 U iso-vec, 9:10 a, 11:12 da, 13:14 dda, 15 dt)
TYPE1 = np.array([
    [1.0, 0.957107, 0.853553, 0.75, 0.707107],
    [0.0, 0.103553, 0.353553, 0.603553, 0.707107],
    [0.0, 0.103553, 0.146447, 0.103553, 1.38778e-17],
    [0.0, -0.25, -0.353553, -0.25, -1.52656e-16],
    [0.0, 0.103553, 0.353553, 0.603553, 0.707107],
    [1.0, 0.75, 0.146447, -0.457107, -0.707107],
    [0.0, -0.25, -0.353553, -0.25, -1.249e-16],
    [0.0, 0.603553, 0.853553, 0.603553, 4.16334e-16],
    [0.0, -0.243953, 0.959151, -0.665253, 0.0],
    [0.0, 0.0139165, 0.668917, 0.625329, 0.0],
    [0.00393491, 0.0240775, -0.00942396, 0.00329391, 0.00941354],
    [-0.00223794, -0.0105816, 0.00328457, 0.0204239, 0.0253415],
    [0.0058186, 0.00686586, -0.00422555, 0.00442631, 0.000319156],
    [-0.00134597, -0.00120682, 0.0114915, 0.00189333, -0.0251649],
    [0.2, 0.2, 0.2, 0.2, 0.2],
])
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Zp = np.diag([1.0, -1.0]).astype(complex)


def problem(kind: str, free_time: bool):
    T = TYPE1.shape[1]
    comps = {"Ũ⃗": (0, 8), "a": (8, 2), "da": (10, 2), "dda": (12, 2)}
    data = TYPE1 if free_time else TYPE1[:14]
    if free_time:
        comps["Δt"] = (14, 1)
    L = ko.Layout(comps, T, "Δt" if free_time else None, 0.2)
    sys_ = ko.QuantumSystem(0.1 * Zp, [X, Y])
    Q = ko.UnitaryPadeIntegrator("Ũ⃗", "a", sys_, L, order=4) if kind == "pade" else ko.UnitaryExponentialIntegrator("Ũ⃗", "a", sys_, L)
    dyn = ko.QuantumDynamics([Q, ko.DerivativeIntegrator("a", "da", L), ko.DerivativeIntegrator("da", "dda", L)], L)
    return dyn, data.reshape(-1, order="F").copy()


if __name__ == "__main__":
    out = {}
    rng = np.random.default_rng(1234)
    for kind in ("pade", "exp"):
        for free_time in (True, False):
            dyn, Zv = problem(kind, free_time)
            mu = rng.normal(size=dyn.dyn * (dyn.T - 1))
            tag = f"{kind}_{'free' if free_time else 'fixed'}"
            out[f"{tag}_Z"] = Zv
            out[f"{tag}_mu"] = mu
            out[f"{tag}_F"] = dyn.F(Zv)
            out[f"{tag}_J"] = dyn.dF(Zv)
            out[f"{tag}_H"] = dyn.mu_d2F(Zv, mu)
            out[f"{tag}_Js"] = np.array(dyn.dF_structure, dtype=np.int64)
            out[f"{tag}_Hs"] = np.array(dyn.mu_d2F_structure, dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hadamard_type1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("pade_free")})
