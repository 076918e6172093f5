import sys; sys.path.insert(0,'/root/repo')
import numpy as np, qcknot
from qcknot import workloads as wl
systems, traj, integrators = wl.config("sampling", T=6, n_systems=6)
full = qcknot.QuantumDynamics(integrators, traj)
Z = traj.datavec; mu = wl.random_multipliers(full.n_blocks*full.dyn)
F,J,H = full.eval_all(Z, mu)
pos = full.shared_hessian_positions(); print("shared", pos)
world=3
Hsum=np.zeros_like(H); parts=[]
for r in range(world):
    q0,q1 = qcknot.sharding.integrator_shard(len(systems), len(integrators), r, world)
    sh = qcknot.QuantumDynamics(integrators, traj, integrator_range=(q0,q1))
    f,j,h = np.zeros_like(F), np.zeros_like(J), np.zeros_like(H)
    sh.eval_all(Z, mu, f,j,h); Hsum+=h; parts.append(h.reshape(-1, full.nnzH))
    print(r,(q0,q1), "shared vals knot0:", h.reshape(-1,full.nnzH)[0,pos])
Hf = H.reshape(-1, full.nnzH); Hs = Hsum.reshape(-1, full.nnzH)
bad = np.argwhere(np.abs(Hf-Hs)>1e-12)
print("full shared knot0:", Hf[0,pos]); print("n bad", len(bad), "cols", sorted(set(bad[:,1].tolist()))[:20])
s = full.mu_d2F_structure[:full.nnzH]
for c in sorted(set(bad[:,1].tolist()))[:10]: print(c, s[c], Hf[0,c], Hs[0,c], [p[0,c] for p in parts])
