import sys; sys.path.insert(0,'/root/repo')
import numpy as np, qcknot
from qcknot import workloads as wl
from oracle.bridge import oracle_dynamics, rel_err
from oracle.c_port import CPort
for name, T in (("cz", 2500), ("hadamard", 5000), ("ket", 3000)):
    systems, traj, integrators = wl.config(name, T=T)
    D = qcknot.QuantumDynamics(integrators, traj)
    O = oracle_dynamics(integrators, traj)
    Z = traj.datavec; mu = wl.random_multipliers(D.n_blocks*D.dyn)
    F, J, H = D.eval_all(Z, mu)
    Fo, Jo, Ho = CPort(O).eval(Z, mu)
    eF = np.abs(F-Fo).reshape(D.n_blocks,-1).max(1); eJ = np.abs(J-Jo).reshape(D.n_blocks,-1).max(1); eH = np.abs(H-Ho).reshape(D.n_blocks,-1).max(1)
    bad = np.nonzero((eF>1e-9)|(eJ>1e-9)|(eH>1e-9))[0]
    print(name, T, "max err", eF.max(), eJ.max(), eH.max(), "bad blocks", len(bad), bad[:20])
