"""Small evaluations of every kernel family, for compute-sanitizer runs (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcknot
from qcknot import workloads as wl

cases = [("cz", {"T": 20}, "pade"), ("hadamard", {"T": 70}, "pade"), ("sampling", {"T": 6, "n_systems": 5}, "pade"),
         ("ket", {"T": 40}, "pade"), ("cz", {"T": 30}, "exponential"), ("hadamard", {"T": 90}, "exponential"),
         ("sampling", {"T": 6, "n_systems": 5}, "exponential"), ("ket", {"T": 40}, "exponential")]
for name, kw, integ in cases:
    systems, traj, integrators = wl.config(name, integrator=integ, **kw)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    print(name, integ, "ok", float(np.abs(F).sum()), float(np.abs(J).sum()), float(np.abs(H).sum()))
    D.close()

# generic spectral kernel (7 and 8 levels, a 9-level ensemble, 9-level kets), large-level kernel (18 levels)
for levels, nd, ket, integ, T in ((7, 2, False, "exponential", 40), (8, 3, False, "exponential", 40), (9, 2, True, "exponential", 30), (18, 2, False, "pade", 6)):
    sys_ = wl.random_hermitian_system(levels, nd, seed=levels, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], T, 0.2, seed=2, ket=ket, n_states=2 if ket else 1)
    D = qcknot.QuantumDynamics(wl.build_integrators([sys_], traj, integrator=integ, ket=ket), traj)
    F, J, H = D.eval_all(traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn))
    print(levels, nd, ket, integ, "ok", float(np.abs(F).sum()), float(np.abs(J).sum()), float(np.abs(H).sum()))
    D.close()
sy = [wl.random_hermitian_system(9, 2, seed=s, scale=0.4) for s in (1, 2, 3)]
traj = wl.random_pulse_trajectory(sy, 12, 0.2, seed=3)
D = qcknot.QuantumDynamics(wl.build_integrators(sy, traj, integrator="exponential"), traj)
F, J, H = D.eval_all(traj.datavec, wl.random_multipliers(D.n_blocks * D.dyn))
print("ensemble 3 x 9 exponential ok", float(np.abs(H).sum()))
D.close()
