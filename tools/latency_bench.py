"""Per-callback latency of the host-buffer path on the small BASELINE configurations (Hadamard T = 50, CZ T = 200): what an Ipopt
iteration pays for F, dF and mu_d2F on a fresh Z (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcknot
from qcknot import workloads as wl

for name, T in (("hadamard", 50), ("cz", 200), ("cz", 1000), ("sampling", 50)):
    kw = {"n_systems": 16} if name == "sampling" else {}
    systems, traj, integrators = wl.config(name, T=T, **kw)
    D = qcknot.QuantumDynamics(integrators, traj)
    nb = D.n_blocks
    Z = [traj.datavec.copy(), traj.datavec.copy() + 1e-7]
    mu = [wl.random_multipliers(nb * D.dyn, seed=s) for s in (1, 2)]
    F, J, H = np.empty(nb * D.dyn), np.empty(nb * D.nnzJ), np.empty(nb * D.nnzH)
    reps = 300
    res = {}
    for label, fn in (("eval_all", lambda i: D.eval_all(Z[i & 1], mu[i & 1], F, J, H)),
                      ("F+dF+mu_d2F on one fresh Z", lambda i: (D.F(Z[i & 1], out=F), D.dF(Z[i & 1], out=J), D.mu_d2F(Z[i & 1], mu[i & 1], out=H))),
                      ("F only", lambda i: D.F(Z[i & 1], out=F))):
        for i in range(10):
            fn(i)
        t0 = time.perf_counter()
        for i in range(reps):
            fn(i)
        res[label] = (time.perf_counter() - t0) / reps * 1e6
    print(f"{name} T={T} ({nb} blocks, {8*(D.dyn+D.nnzJ+D.nnzH)*nb/1e6:.2f} MB of values): " + ", ".join(f"{k}: {v:.0f} us" for k, v in res.items()), flush=True)
    D.close()
