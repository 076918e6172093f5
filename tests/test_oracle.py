"""CPU tests of the oracle: against the committed golden fixture, the reference's own layout fixtures, and independent
derivations (finite differences, 40-digit mpmath on the complex-form definition, Pade-vs-exponential consistency)."""
import os

import mpmath as mp
import numpy as np
import pytest

from oracle import knot_oracle as ko
from oracle.c_port import CPort

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hadamard_type1.npz"))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def test_iso_layout_matches_reference_fixture():
    # test/test_utils.jl:103: iso-vec of I_2 is [1,0,0,0,0,1,0,0]; goal (X gate... Hadamard fixture) test_utils.jl:107
    assert np.array_equal(ko.operator_to_iso_vec(np.eye(2)), [1, 0, 0, 0, 0, 1, 0, 0])
    U = np.array([[1, 2 + 3j], [4j, 5]])
    v = ko.operator_to_iso_vec(U)
    assert np.array_equal(v, [1, 0, 0, 4, 2, 5, 3, 0])
    assert np.allclose(ko.iso_vec_to_operator(v), U)
    # the literal 15x5 trajectory's last column is the Hadamard gate in this layout (test_utils.jl:55-62)
    Hd = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    assert np.allclose(make_golden.TYPE1[:8, -1], ko.operator_to_iso_vec(Hd), atol=1e-6)
    # G = iso(-iH) is antisymmetric for Hermitian H and acts on [Re; Im] like -iH on the complex vector
    sys_ = ko.QuantumSystem(0.3 * make_golden.Zp, [make_golden.X, make_golden.Y])
    G = sys_.G([0.2, -0.7])
    assert np.allclose(G, -G.T)
    psi = np.array([0.3 + 0.1j, -0.5j])
    assert np.allclose(G @ ko.ket_to_iso(psi), ko.ket_to_iso(-1j * sys_.H([0.2, -0.7]) @ psi))


def test_pade_coefficients():
    # order 4: 1, 1/2, 1/12; 6: 1, 1/2, 1/10, 1/120; 8: ..., 3/28, 1/84, 1/1680
    assert np.allclose(ko.pade_coefficients(4), [1, 1 / 2, 1 / 12])
    assert np.allclose(ko.pade_coefficients(6), [1, 1 / 2, 1 / 10, 1 / 120])
    assert np.allclose(ko.pade_coefficients(8), [1, 1 / 2, 3 / 28, 1 / 84, 1 / 1680])
    assert np.allclose(ko.pade_coefficients(10), [1, 1 / 2, 1 / 9, 1 / 72, 1 / 1008, 1 / 30240])


@pytest.mark.parametrize("kind", ["pade", "exp"])
@pytest.mark.parametrize("free_time", [True, False])
def test_oracle_reproduces_golden(kind, free_time):
    dyn, Z = make_golden.problem(kind, free_time)
    tag = f"{kind}_{'free' if free_time else 'fixed'}"
    assert np.array_equal(Z, GOLD[f"{tag}_Z"])
    mu = GOLD[f"{tag}_mu"]
    assert np.array_equal(np.array(dyn.dF_structure), GOLD[f"{tag}_Js"])
    assert np.array_equal(np.array(dyn.mu_d2F_structure), GOLD[f"{tag}_Hs"])
    assert np.allclose(dyn.F(Z), GOLD[f"{tag}_F"], rtol=0, atol=1e-13)
    assert np.allclose(dyn.dF(Z), GOLD[f"{tag}_J"], rtol=0, atol=1e-13)
    assert np.allclose(dyn.mu_d2F(Z, mu), GOLD[f"{tag}_H"], rtol=0, atol=1e-13)


def test_sizes_match_survey_table():
    # SURVEY.md section 8 size table: C1 nnzJ 104/80, nnzH 58/34, dyn 12, zdim 15
    for kind, nj, nh in (("pade", 104, 58), ("exp", 80, 34)):
        dyn, _ = make_golden.problem(kind, True)
        assert (dyn.dyn, dyn.zdim, dyn.nnzJ, dyn.nnzH) == (12, 15, nj, nh)
    # row / column counts of integrator_test_1qubit.jl:44,48,50
    dyn, Z = make_golden.problem("pade", True)
    assert dyn.F(Z).size == dyn.dyn * (dyn.T - 1)
    s = np.array(dyn.dF_structure)
    assert s[:, 0].max() <= dyn.dyn * (dyn.T - 1) and s[:, 1].max() <= dyn.zdim * dyn.T
    h = np.array(dyn.mu_d2F_structure)
    assert np.all(h[:, 0] <= h[:, 1])  # upper triangle (test_utils.jl:22-24 Symmetric(M))


def _fd_jac_hess(dyn, Z, mu, eps=1e-6):
    n, F0 = Z.size, dyn.F(Z)
    J = np.zeros((F0.size, n))
    Hm = np.zeros((n, n))
    g = lambda z: ko.dense(dyn.dF(z), dyn.dF_structure, (F0.size, n)).T @ mu
    for i in range(n):
        e = np.zeros(n)
        e[i] = eps
        J[:, i] = (dyn.F(Z + e) - dyn.F(Z - e)) / (2 * eps)
        Hm[:, i] = (g(Z + e) - g(Z - e)) / (2 * eps)
    return J, Hm


@pytest.mark.parametrize("kind", ["pade4", "pade8", "exp"])
@pytest.mark.parametrize("ket", [False, True])
def test_oracle_vs_finite_differences(kind, ket):
    rng = np.random.default_rng(5)
    n = 4 if ket else 8
    comps = {"x": (0, n), "a": (n, 2), "da": (n + 2, 2), "dt": (n + 4, 1)}
    L = ko.Layout(comps, 3, "dt")
    sys_ = ko.QuantumSystem(0.3 * make_golden.Zp, [make_golden.X, make_golden.Y])
    if kind == "exp":
        Q = (ko.QuantumStateExponentialIntegrator if ket else ko.UnitaryExponentialIntegrator)("x", "a", sys_, L)
    else:
        Q = (ko.QuantumStatePadeIntegrator if ket else ko.UnitaryPadeIntegrator)("x", "a", sys_, L, order=int(kind[4:]))
    dyn = ko.QuantumDynamics([Q, ko.DerivativeIntegrator("a", "da", L)], L)
    Z = rng.normal(size=L.zdim * 3) * 0.5
    Z[L.zdim - 1 :: L.zdim] = 0.2 + 0.05 * rng.random(3)
    mu = rng.normal(size=dyn.dyn * 2)
    Jfd, Hfd = _fd_jac_hess(dyn, Z, mu)
    J = ko.dense(dyn.dF(Z), dyn.dF_structure, Jfd.shape)
    Hs = ko.dense(dyn.mu_d2F(Z, mu), dyn.mu_d2F_structure, Hfd.shape)
    assert np.abs(J - Jfd).max() < 5e-9
    assert np.abs(Hs - Hfd).max() < 5e-9


def _mp_residual(kind, H0, Hd, U0, U1, a, h):
    """Complex-form definition in mpmath: Pade-4  B U1 - F U0,  exponential  U1 - expm(-i H h) U0   (README.md:79)."""
    H = H0 + sum((aj * Hj for aj, Hj in zip(a, Hd)), mp.zeros(H0.rows))
    A = -1j * H
    I = mp.eye(H0.rows)
    if kind == "pade":
        F = I + h / 2 * A + h * h / 12 * A * A
        B = I - h / 2 * A + h * h / 12 * A * A
        return B * U1 - F * U0
    return U1 - mp.expm(A * h, method="taylor") * U0


@pytest.mark.parametrize("kind", ["pade", "exp"])
def test_oracle_vs_mpmath_definition(kind):
    """70-digit central differences on the complex definition: derivative truth to ~1e-30, compared at 1e-12."""
    mp.mp.dps = 70
    rng = np.random.default_rng(11)
    Xm, Ym, Zm = (mp.matrix(m.tolist()) for m in (make_golden.X, make_golden.Y, make_golden.Zp))
    H0, Hd = Zm * mp.mpf("0.3"), [Xm, Ym]
    U0n = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    U1n = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    an, hn = rng.normal(size=2), 0.23
    Mn = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    comps = {"x": (0, 8), "a": (8, 2), "dt": (10, 1)}
    L = ko.Layout(comps, 2, "dt")
    sys_ = ko.QuantumSystem(0.3 * make_golden.Zp, [make_golden.X, make_golden.Y])
    Q = ko.UnitaryPadeIntegrator("x", "a", sys_, L) if kind == "pade" else ko.UnitaryExponentialIntegrator("x", "a", sys_, L)
    zt = np.concatenate([ko.operator_to_iso_vec(U0n), an, [hn]])
    zt1 = np.concatenate([ko.operator_to_iso_vec(U1n), an * 0, [hn]])
    mu = ko.operator_to_iso_vec(Mn)

    def lagr(a0, a1, h):  # mu^T R = Re <M, R>
        R = _mp_residual(kind, H0, Hd, mp.matrix(U0n.tolist()), mp.matrix(U1n.tolist()), [a0, a1], h)
        return sum((mp.conj(mp.mpc(Mn[i, j])) * R[i, j]).real for i in range(2) for j in range(2))

    # residual
    R = _mp_residual(kind, H0, Hd, mp.matrix(U0n.tolist()), mp.matrix(U1n.tolist()), [mp.mpf(an[0]), mp.mpf(an[1])], mp.mpf(hn))
    Rn = np.array([[complex(R[i, j]) for j in range(2)] for i in range(2)])
    assert np.abs(Q.residual(zt, zt1) - ko.operator_to_iso_vec(Rn)).max() < 1e-13
    # gradient of mu^T R wrt (a0, a1, h) == J^T mu, and its Hessian block == oracle Hessian
    e = mp.mpf(10) ** -20
    x0 = [mp.mpf(an[0]), mp.mpf(an[1]), mp.mpf(hn)]
    f = lambda x: lagr(x[0], x[1], x[2])
    grad, hess = np.zeros(3), np.zeros((3, 3))
    for i in range(3):
        xp, xm = list(x0), list(x0)
        xp[i] += e
        xm[i] -= e
        grad[i] = float((f(xp) - f(xm)) / (2 * e))
        for j in range(3):
            xpp, xpm, xmp, xmm = list(x0), list(x0), list(x0), list(x0)
            xpp[i] += e; xpp[j] += e
            xpm[i] += e; xpm[j] -= e
            xmp[i] -= e; xmp[j] += e
            xmm[i] -= e; xmm[j] -= e
            hess[i, j] = float((f(xpp) - f(xpm) - f(xmp) + f(xmm)) / (4 * e * e))
    J = Q.jacobian(zt, zt1)
    assert np.abs(J[:, [8, 9, 10]].T @ mu - grad).max() < 1e-12
    Hh = Q.hessian(zt, zt1, mu)
    Hsym = np.triu(Hh) + np.triu(Hh, 1).T
    assert np.abs(Hsym[np.ix_([8, 9, 10], [8, 9, 10])] - hess).max() < 1e-12


def test_pade_agrees_with_exponential_to_integrator_order():
    """Sanity link: for a solution of the exponential dynamics, the order-2m Pade residual is O(dt^(2m+1))."""
    sys_ = ko.QuantumSystem(0.3 * make_golden.Zp, [make_golden.X, make_golden.Y])
    a = np.array([0.4, -0.2])
    errs = []
    for dt in (0.2, 0.1):
        comps = {"x": (0, 8), "a": (8, 2), "dt": (10, 1)}
        L = ko.Layout(comps, 2, "dt")
        U0 = np.eye(2)
        import scipy.linalg as sla
        U1 = sla.expm(-1j * sys_.H(a) * dt) @ U0
        zt = np.concatenate([ko.operator_to_iso_vec(U0), a, [dt]])
        zt1 = np.concatenate([ko.operator_to_iso_vec(U1), a, [dt]])
        assert np.abs(ko.UnitaryExponentialIntegrator("x", "a", sys_, L).residual(zt, zt1)).max() < 1e-14
        errs.append([np.abs(ko.UnitaryPadeIntegrator("x", "a", sys_, L, order=o).residual(zt, zt1)).max() for o in (4, 6)])
    assert 25 < errs[0][0] / errs[1][0] < 40  # ~2^5
    assert 100 < errs[0][1] / errs[1][1] < 160  # ~2^7


def test_dense_contract_sums_duplicates_and_symmetrises():
    # test/test_utils.jl:14-27
    M = ko.dense([1.0, 2.0, 5.0], [(1, 2), (1, 2), (2, 2)], (2, 2))
    assert np.array_equal(M, [[0, 3], [3, 5]])
    M = ko.dense([1.0, 2.0], [(1, 3), (1, 3)], (2, 3))
    assert M[0, 2] == 3


@pytest.mark.parametrize("name,kw", [("hadamard", {"T": 5}), ("sampling", {"T": 3, "n_systems": 4}), ("ket", {"T": 4, "free_time": False}),
                                     ("hadamard", {"T": 4, "integrator": "exponential"})])
def test_structure_order_policies_describe_the_same_matrices(name, kw):
    """The three intra-knot orders (csc | row_major | per_integrator with duplicates) are different listings of the same sparse
    Jacobian and Hessian of the Lagrangian: rebuilt with the reference's dense() (duplicates sum, test/test_utils.jl:14-27) they
    agree, and per_integrator has duplicates exactly where integrators share a Hessian position."""
    import qcknot  # host objects only (no device needed)
    from oracle.bridge import oracle_dynamics
    from qcknot import workloads as wl
    systems, traj, integrators = wl.config(name, **kw)
    Z = traj.datavec
    O = {o: oracle_dynamics(integrators, traj, structure_order=o) for o in ("csc", "row_major", "per_integrator")}
    mu = wl.random_multipliers((traj.T - 1) * O["csc"].dyn)
    nZ, nF = traj.T * traj.dim, (traj.T - 1) * O["csc"].dyn
    Jd = {o: ko.dense(D.dF(Z), D.dF_structure, (nF, nZ)) for o, D in O.items()}
    Hd = {o: ko.dense(D.mu_d2F(Z, mu), D.mu_d2F_structure, (nZ, nZ)) for o, D in O.items()}
    for o in ("row_major", "per_integrator"):
        assert np.array_equal(Jd[o], Jd["csc"])
        assert np.abs(Hd[o] - Hd["csc"]).max() <= 1e-14 * max(1.0, np.abs(Hd["csc"]).max())
        assert sorted(O[o].dF_structure) == sorted(O["csc"].dF_structure)
    assert sorted(O["row_major"].mu_d2F_structure) == sorted(O["csc"].mu_d2F_structure)
    dup = len(O["per_integrator"].mu_d2F_structure) - len(set(O["per_integrator"].mu_d2F_structure))
    assert dup == len(O["per_integrator"].mu_d2F_structure) - len(O["csc"].mu_d2F_structure)
    assert (dup > 0) == (sum(1 for I in integrators if type(I).__name__ != "DerivativeIntegrator") > 1)  # shared controls / timestep


@pytest.mark.parametrize("case", ["unitary", "ket", "sampling", "fixed"])
def test_c_port_matches_numpy_oracle(case):
    import qcknot  # host objects only (no device needed)
    from oracle.bridge import oracle_dynamics, rel_err
    from qcknot import workloads as wl
    cfg = {"unitary": ("cz", {"T": 4}), "ket": ("ket", {"T": 6}), "sampling": ("sampling", {"T": 3, "n_systems": 3}),
           "fixed": ("hadamard", {"T": 6, "free_time": False})}[case]
    systems, traj, integrators = wl.config(cfg[0], **cfg[1])
    O = oracle_dynamics(integrators, traj)
    Z, mu = traj.datavec, wl.random_multipliers((traj.T - 1) * O.dyn)
    for nthreads in (1, 3):
        F, J, H = CPort(O).eval(Z, mu, nthreads=nthreads)
        assert rel_err(F, O.F(Z)) < 1e-13 and rel_err(J, O.dF(Z)) < 1e-13 and rel_err(H, O.mu_d2F(Z, mu)) < 1e-13
