"""Small evaluations of every kernel family, for compute-sanitizer runs (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcknot
from qcknot import workloads as wl

cases = [("cz", {"T": 20}, "pade"), ("hadamard", {"T": 70}, "pade"), ("sampling", {"T": 6, "n_systems": 5}, "pade"),
         ("ket", {"T": 40}, "pade"), ("cz", {"T": 5}, "exponential"), ("hadamard", {"T": 9}, "exponential")]
for name, kw, integ in cases:
    systems, traj, integrators = wl.config(name, integrator=integ, **kw)
    D = qcknot.QuantumDynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers(D.n_blocks * D.dyn)
    F, J, H = D.eval_all(Z, mu)
    print(name, integ, "ok", float(np.abs(F).sum()), float(np.abs(J).sum()), float(np.abs(H).sum()))
    D.close()
