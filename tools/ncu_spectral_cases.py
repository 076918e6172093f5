import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch, qcknot
from qcknot import workloads as wl
which = sys.argv[1]
if which == "hadamard":
    systems, traj, integrators = wl.config("hadamard", T=100000, integrator="exponential")
elif which == "sampling":
    systems, traj, integrators = wl.config("sampling", T=200, n_systems=256, integrator="exponential")
else:
    sys_ = wl.random_hermitian_system(8, 2, seed=1, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], 4000, 0.2, seed=1)
    integrators = wl.build_integrators([sys_], traj, integrator="exponential")
D = qcknot.QuantumDynamics(integrators, traj)
nb = D.n_blocks; dev = torch.device("cuda:0")
Z = torch.from_numpy(traj.datavec).to(dev); mu = torch.from_numpy(wl.random_multipliers(nb * D.dyn)).to(dev)
F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev); J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev); H = torch.empty(nb * D.nnzH, dtype=torch.float64, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(4): D.eval_device(7, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
