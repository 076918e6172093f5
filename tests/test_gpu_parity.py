"""GPU parity: libqcknot.so (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerance: 1e-10 relative to the largest magnitude of the array (BASELINE.json north_star: "relative tolerance of
1e-10 in FP64"); structures must be bit-exact."""
import numpy as np
import pytest

import qcknot
from qcknot import workloads as wl

from helpers import entry_err, oracle_dynamics, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10        # relative to the largest magnitude of the array (north_star)
ENTRY_TOL = 1e-9   # per entry, relative to max(|entry|, 1e-6 * largest magnitude): catches a wrong small term in a large array


def check(systems, traj, integrators, eval_hessian=True):
    D = qcknot.QuantumDynamics(integrators, traj, eval_hessian=eval_hessian)
    O = oracle_dynamics(integrators, traj, eval_hessian=eval_hessian)
    assert (D.dyn, D.nnzJ, D.nnzH) == (O.dyn, O.nnzJ, O.nnzH)
    assert np.array_equal(D.dF_structure, np.array(O.dF_structure))
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F = D.F(Z)
    J = D.dF(Z)
    assert rel_err(F, O.F(Z)) < TOL and entry_err(F, O.F(Z)) < ENTRY_TOL
    assert rel_err(J, O.dF(Z)) < TOL and entry_err(J, O.dF(Z)) < ENTRY_TOL
    if eval_hessian:
        assert np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure).reshape(-1, 2))
        H = D.mu_d2F(Z, mu)
        assert rel_err(H, O.mu_d2F(Z, mu)) < TOL and entry_err(H, O.mu_d2F(Z, mu)) < ENTRY_TOL
        F2, J2, H2 = D.eval_all(Z, mu)  # fused pass must agree bitwise with the separate calls
        assert np.array_equal(F, F2) and np.array_equal(J, J2) and np.array_equal(H, H2)
    D.close()


@pytest.mark.parametrize("order", ["row_major", "per_integrator"])
@pytest.mark.parametrize("name,kw", [("hadamard", {"T": 9}), ("hadamard", {"T": 6, "integrator": "exponential"}), ("cz", {"T": 5}),
                                     ("ket", {"T": 6, "free_time": False}), ("sampling", {"T": 4, "n_systems": 5})])
def test_structure_order_policies(name, kw, order):
    """Non-default structure orders (qck_problem_desc.structure_order): values against the oracle's restatement of the same policy,
    through the host-buffer entry points and the device-resident one; the assembled sparse matrices equal the CSC handle's."""
    import torch
    systems, traj, integrators = wl.config(name, **kw)
    D = qcknot.QuantumDynamics(integrators, traj, structure_order=order)
    O = oracle_dynamics(integrators, traj, structure_order=order)
    assert np.array_equal(D.dF_structure, np.array(O.dF_structure).reshape(-1, 2))
    assert np.array_equal(D.mu_d2F_structure, np.array(O.mu_d2F_structure).reshape(-1, 2))
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    for got, want in ((F, O.F(Z)), (J, O.dF(Z)), (H, O.mu_d2F(Z, mu))):
        assert rel_err(got, want) < TOL and entry_err(got, want) < ENTRY_TOL
    assert np.array_equal(J, D.dF(Z)) and np.array_equal(H, D.mu_d2F(Z, mu)) and np.array_equal(F, D.F(Z))
    # same matrices as the default order, bit for bit (Jacobian) / after summing duplicates (Hessian)
    C0 = qcknot.QuantumDynamics(integrators, traj)
    F0, J0, H0 = C0.eval_all(Z, mu)
    assert np.array_equal(F, F0)
    key = lambda st: np.lexsort((st[:, 0], st[:, 1]))
    assert np.array_equal(J[key(D.dF_structure)], J0[key(C0.dF_structure)])
    if order == "row_major":
        assert np.array_equal(H[key(D.mu_d2F_structure)], H0[key(C0.mu_d2F_structure)])
    else:
        n = traj.T * traj.dim
        dense = lambda v, st: np.bincount(((st[:, 0] - 1) * n + st[:, 1] - 1), weights=v, minlength=0)
        a, b = dense(H, D.mu_d2F_structure), dense(H0, C0.mu_d2F_structure)
        m = max(len(a), len(b))
        a, b = np.pad(a, (0, m - len(a))), np.pad(b, (0, m - len(b)))
        assert rel_err(a, b) < 1e-13
    # device-resident entry point writes the caller's arrays in the same order
    dev = torch.device("cuda:0")
    dZ, dmu = torch.from_numpy(Z).to(dev), torch.from_numpy(mu).to(dev)
    dF = torch.empty(len(F), dtype=torch.float64, device=dev)
    dJ = torch.empty(len(J), dtype=torch.float64, device=dev)
    dH = torch.empty(max(len(H), 1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    D.eval_device(7, dZ.data_ptr(), dmu.data_ptr(), dF.data_ptr(), dJ.data_ptr(), dH.data_ptr(), 0)
    D.synchronize()
    assert np.array_equal(dJ.cpu().numpy(), J) and np.array_equal(dH.cpu().numpy()[: len(H)], H) and np.array_equal(dF.cpu().numpy(), F)
    for x in (D, C0):
        x.close()


@pytest.mark.parametrize("free_time", [True, False])
def test_hadamard_pade(free_time):
    check(*wl.config("hadamard", T=12, free_time=free_time))


def test_hadamard_no_hessian():
    check(*wl.config("hadamard", T=6), eval_hessian=False)


@pytest.mark.parametrize("free_time", [True, False])
def test_cz_pade(free_time):
    check(*wl.config("cz", T=6, free_time=free_time))


@pytest.mark.parametrize("free_time", [True, False])
def test_ket_pade(free_time):
    check(*wl.config("ket", T=9, free_time=free_time))


def test_sampling_pade():
    check(*wl.config("sampling", T=4, n_systems=5))


@pytest.mark.parametrize("levels,nd", [(3, 1), (4, 2), (5, 3), (6, 2), (2, 3), (3, 4), (4, 4), (4, 5)])
def test_random_dense_systems(levels, nd):
    sys_ = wl.random_hermitian_system(levels, nd, seed=levels * 10 + nd, scale=0.7)
    traj = wl.random_pulse_trajectory([sys_], 4, 0.3, seed=7)
    check([sys_], traj, wl.build_integrators([sys_], traj))


@pytest.mark.parametrize("name,kw", [("hadamard", {"T": 9}), ("hadamard", {"T": 6, "free_time": False}), ("cz", {"T": 5}),
                                     ("ket", {"T": 8}), ("sampling", {"T": 3, "n_systems": 4})])
def test_exponential_integrators(name, kw):
    """UnitaryExponentialIntegrator / QuantumStateExponentialIntegrator: residual + Jacobian (the reference has no
    Hessian for them: SURVEY 8a3), against scipy expm / expm_frechet in the oracle."""
    check(*wl.config(name, integrator="exponential", **kw), eval_hessian=False)


@pytest.mark.parametrize("name,kw", [("hadamard", {"T": 9}), ("hadamard", {"T": 6, "free_time": False}), ("cz", {"T": 4}),
                                     ("ket", {"T": 8}), ("sampling", {"T": 3, "n_systems": 3})])
def test_exponential_integrators_with_hessian(name, kw):
    """Hessian of mu^T (U1 - exp(h A) U0): second Frechet derivatives by a reverse sweep over the scaling-and-squaring
    tape; oracle = scipy expm / expm_frechet + block-triangular second derivative."""
    check(*wl.config(name, integrator="exponential", **kw), eval_hessian=True)


def test_exponential_large_norm():
    """Scaling-and-squaring with many squarings: ||h A|| ~ 30."""
    sys_ = wl.random_hermitian_system(5, 2, seed=3, scale=4.0)
    traj = wl.random_pulse_trajectory([sys_], 4, 1.5, seed=11)
    check([sys_], traj, wl.build_integrators([sys_], traj, integrator="exponential"), eval_hessian=False)
    check([sys_], traj, wl.build_integrators([sys_], traj, integrator="exponential"), eval_hessian=True)


# ---- 9-level exponential unitaries: the spectral kernel (qck_expeig.cu) ---------------------------------------------------------
@pytest.mark.parametrize("nd", [1, 2, 3, 4])
@pytest.mark.parametrize("free_time", [True, False])
def test_spectral_exponential_dense_drives(nd, free_time):
    """Jacobi eigen-decomposition + divided differences against scipy expm / expm_frechet / block-triangular second derivatives:
    dense random Hamiltonians (sparse-row width 9), 1..4 drives, free and fixed timestep; F, F+J and F+J+H calls."""
    sys_ = wl.random_hermitian_system(9, nd, seed=190 + nd, scale=0.5)
    traj = wl.random_pulse_trajectory([sys_], 5, 0.25, seed=13 + nd, free_time=free_time)
    integrators = wl.build_integrators([sys_], traj, integrator="exponential")
    check([sys_], traj, integrators, eval_hessian=True)
    check([sys_], traj, integrators, eval_hessian=False)


@pytest.mark.parametrize("amp", [0.0, 1e-9, 1e-4, 3e-2])
def test_spectral_exponential_degenerate_levels(amp):
    """The CZ drift has exactly degenerate levels; with controls of size amp the spectrum is exactly / nearly / mildly degenerate:
    first- and second-order divided differences must not cancel (sinc form, series about the mean for close triples)."""
    systems, traj, integrators = wl.config("cz", T=5, integrator="exponential")
    a = traj["a"]  # view into the trajectory's data (controls x T)
    a[:] = amp * np.sign(a) * (1.0 + np.arange(traj.T)[None, :])
    check(systems, traj, integrators, eval_hessian=True)


def test_spectral_exponential_large_norm():
    """||h A|| ~ 60: no squaring levels, no tape, nothing to overflow -- the phases just wrap."""
    sys_ = wl.random_hermitian_system(9, 3, seed=31, scale=4.0)
    traj = wl.random_pulse_trajectory([sys_], 4, 1.5, seed=12)
    integrators = wl.build_integrators([sys_], traj, integrator="exponential")
    check([sys_], traj, integrators, eval_hessian=True)


# ---- 2..4-level exponential unitaries and kets: the spectral column kernels (qck_colexp.cu) -----------------------------------------
@pytest.mark.parametrize("levels,nd,ket", [(2, 1, False), (2, 3, False), (3, 2, False), (3, 4, False), (4, 1, False), (4, 2, False), (4, 4, False),
                                           (2, 2, True), (3, 3, True), (4, 2, True), (4, 4, True)])
@pytest.mark.parametrize("free_time", [True, False])
def test_spectral_column_exponential(levels, nd, ket, free_time):
    """One lane per item diagonalises H(a_t) (Jacobi in registers), one lane per column evaluates the divided-difference forms:
    dense random Hamiltonians, every level count / drive count / state kind the kernels are instantiated for."""
    sys_ = wl.random_hermitian_system(levels, nd, seed=300 + 10 * levels + nd, scale=0.6)
    traj = wl.random_pulse_trajectory([sys_], 7, 0.3, seed=21 + nd, free_time=free_time, ket=ket, n_states=2 if ket else 1)
    integrators = wl.build_integrators([sys_], traj, integrator="exponential", ket=ket)
    check([sys_], traj, integrators, eval_hessian=True)
    check([sys_], traj, integrators, eval_hessian=False)


@pytest.mark.parametrize("amp", [0.0, 1e-9, 1e-3])
def test_spectral_column_exponential_degenerate_levels(amp):
    """Hadamard problem without drift: H(a) = a_x X + a_y Y has the levels +-|a| -- exactly degenerate at a = 0, nearly so for tiny
    controls; a 4-level ensemble whose members share the controls goes through the partial columns."""
    systems, traj, integrators = wl.config("hadamard", T=9, integrator="exponential")
    a = traj["a"]
    a[:] = amp * np.sign(a + 1e-30) * (1.0 + np.arange(traj.T)[None, :])
    check(systems, traj, integrators, eval_hessian=True)
    systems, traj, integrators = wl.config("sampling", T=4, n_systems=5, integrator="exponential")
    a = traj["a"]
    a[:] = amp * np.sign(a + 1e-30)
    check(systems, traj, integrators, eval_hessian=True)


def test_spectral_column_exponential_large_norm():
    sys_ = wl.random_hermitian_system(4, 2, seed=5, scale=5.0)
    traj = wl.random_pulse_trajectory([sys_], 5, 2.0, seed=6)
    check([sys_], traj, wl.build_integrators([sys_], traj, integrator="exponential"), eval_hessian=True)


# ---- every other Hermitian exponential class up to 16 levels: the generic spectral kernel (qck_genexp.cu) ---------------------------
@pytest.mark.parametrize("levels,nd,ket", [(5, 2, False), (6, 1, False), (7, 3, False), (8, 4, False), (10, 2, False), (12, 2, False), (16, 2, False),
                                           (5, 2, True), (8, 3, True), (9, 4, True), (16, 1, True)])
@pytest.mark.parametrize("free_time", [True, False])
def test_generic_spectral_exponential(levels, nd, ket, free_time):
    """One warp per item, run-time level count (odd and even: the two round-robin orderings of the Jacobi sweeps), unitaries and
    kets, 1..4 drives."""
    sys_ = wl.random_hermitian_system(levels, nd, seed=500 + 10 * levels + nd, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], 4, 0.25, seed=31 + nd, free_time=free_time, ket=ket, n_states=2 if ket else 1)
    integrators = wl.build_integrators([sys_], traj, integrator="exponential", ket=ket)
    check([sys_], traj, integrators, eval_hessian=True)
    if levels in (5, 8):
        check([sys_], traj, integrators, eval_hessian=False)


def test_generic_spectral_exponential_ensemble_and_degenerate():
    """Three 9-level systems sharing the controls (shared-control Hessian entries through the partial columns); the CZ drift with
    vanishing controls as a 9-level KET problem (exactly degenerate levels)."""
    sy = [wl.random_hermitian_system(9, 2, seed=s, scale=0.4) for s in (11, 12, 13)]
    traj = wl.random_pulse_trajectory(sy, 4, 0.2, seed=3)
    check(sy, traj, wl.build_integrators(sy, traj, integrator="exponential"), eval_hessian=True)
    cz = wl.two_transmon_cz_system()
    traj = wl.random_pulse_trajectory([cz], 4, 1.0, seed=4, ket=True, n_states=2)
    traj["a"][:] = 0.0
    check([cz], traj, wl.build_integrators([cz], traj, integrator="exponential", ket=True), eval_hessian=True)


@pytest.mark.parametrize("order", [6, 8, 10, 12])
@pytest.mark.parametrize("name,kw", [("hadamard", {"T": 5}), ("hadamard", {"T": 4, "free_time": False}), ("cz", {"T": 3}), ("ket", {"T": 5})])
def test_general_pade_orders(order, name, kw):
    """UnitaryPadeIntegrator(...; order) for the other orders the reference's PADE_COEFFICIENTS cover:
    ratio-form Horner with tangents + reverse sweep for the Hessian, against the oracle's explicit power sums."""
    systems, traj, integrators = wl.config(name, **kw)
    ket = name == "ket"
    integrators = wl.build_integrators(systems, traj, order=order, ket=ket)
    check(systems, traj, integrators)
    if name == "hadamard":
        check(systems, traj, integrators, eval_hessian=False)


# ---- 9-level Pade-4 unitaries: the three-warps-per-knot kernel (qck_rs3.cu) and its variants ------------------------------------
@pytest.mark.parametrize("nd", [1, 2, 3, 4])
def test_nine_levels_dense_drives(nd):
    """Dense drive matrices (sparse-row width 9: the run-time width path of the kernel), 1..4 drives."""
    sys_ = wl.random_hermitian_system(9, nd, seed=90 + nd, scale=0.5)
    traj = wl.random_pulse_trajectory([sys_], 5, 0.25, seed=3 + nd)
    check([sys_], traj, wl.build_integrators([sys_], traj))


def test_nine_levels_non_hermitian():
    """A(a) = -i H(a) is not anti-Hermitian here: the A^H products must take the general path."""
    rng = np.random.default_rng(5)
    mk = lambda: 0.4 * (rng.normal(size=(9, 9)) + 1j * rng.normal(size=(9, 9)))
    sys_ = qcknot.QuantumSystem(mk(), [mk(), mk()])
    traj = wl.random_pulse_trajectory([sys_], 4, 0.2, seed=8)
    check([sys_], traj, wl.build_integrators([sys_], traj))


def test_nine_levels_five_drives_fall_back_to_tiled_kernel():
    """More than four drives: the tiled shared-memory kernel handles the class."""
    sys_ = wl.random_hermitian_system(9, 5, seed=77, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], 4, 0.2, seed=9)
    check([sys_], traj, wl.build_integrators([sys_], traj))


_VARIANT_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import qcknot
from qcknot import workloads as wl
from helpers import entry_err, oracle_dynamics, rel_err
for free_time in (True, False):
    systems, traj, integrators = wl.config({name!r}, T=6, free_time=free_time, **{kw!r})
    D = qcknot.QuantumDynamics(integrators, traj); O = oracle_dynamics(integrators, traj)
    Z = traj.datavec; mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    assert rel_err(F, O.F(Z)) < 1e-10 and rel_err(J, O.dF(Z)) < 1e-10 and rel_err(H, O.mu_d2F(Z, mu)) < 1e-10
    assert np.array_equal(J, D.dF(Z)) and np.array_equal(H, D.mu_d2F(Z, mu))
    D.close()
print("variant ok")
"""


@pytest.mark.parametrize("name,kw,env", [("cz", {}, {"QCK_RS3": "0"}), ("cz", {}, {"QCK_RS3": "0", "QCK_DMMA": "1"}),
                                         ("cz", {}, {"QCK_ROWSLICE_DENSE": "1"}), ("cz", {}, {"QCK_ROWSLICE_WARPS": "3"}),
                                         ("cz", {}, {"QCK_RS3": "5"}), ("cz", {}, {"QCK_RS3": "7"}),
                                         ("cz", {}, {"QCK_ROWSLICE_BW": "0"}), ("cz", {}, {"QCK_ROWSLICE_BW": "0", "QCK_ROWSLICE_SPREAD": "0"}),
                                         ("cz", {}, {"QCK_ROWSLICE_DB": "1", "QCK_ROWSLICE_BW": "0"}),
                                         ("hadamard", {}, {"QCK_COLUMN": "0"}), ("sampling", {"n_systems": 5}, {"QCK_COLUMN": "0"}),
                                         ("ket", {}, {"QCK_COLUMN": "0"}), ("cz", {"integrator": "exponential"}, {"QCK_EXPEIG": "0", "QCK_GENEXP": "0"}),
                                         ("cz", {"integrator": "exponential"}, {"QCK_EXPEIG_WARPS": "3"}),
                                         ("hadamard", {"integrator": "exponential"}, {"QCK_COLEXP": "0"}),
                                         ("sampling", {"n_systems": 4, "integrator": "exponential"}, {"QCK_COLEXP": "0"}),
                                         ("cz", {"integrator": "exponential"}, {"QCK_EXPEIG": "0", "QCK_GENEXP": "1"})])
def test_kernel_variants(name, kw, env):
    """The launch knobs are read once per process, so each variant runs in its own interpreter: the tiled DFMA kernel
    (QCK_RS3=0 / QCK_COLUMN=0), its FP64 tensor-core (DMMA) variant, the row-slice kernel with dense drives / a smaller CTA, the
    three-warps-per-knot variant of the row-slice kernel with 5 / 7 knots per CTA, the row-slice kernel without its block warps
    (with / without early block copies, with two staging buffers), the exponential CZ problem on the scaling-and-squaring kernel
    instead of the spectral one."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT.format(root=root, tests=os.path.join(root, "tests"), name=name, kw=kw)],
                         env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout + out.stderr
