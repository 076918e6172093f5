"""Quick device-resident timing of the fused F+J+H pass (development tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qcknot
from qcknot import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "cz"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
integ = sys.argv[3] if len(sys.argv) > 3 else "pade"
nsys = int(sys.argv[4]) if len(sys.argv) > 4 else None
t0 = time.time()
systems, traj, integrators = wl.config(name, T=T, integrator=integ, n_systems=nsys)
print(f"workload built in {time.time()-t0:.1f}s")
hess = True
D = qcknot.QuantumDynamics(integrators, traj, eval_hessian=hess)
nb = D.n_blocks
print(f"dyn={D.dyn} nnzJ={D.nnzJ} nnzH={D.nnzH} blocks={nb}")
dev = torch.device("cuda:0")
Z = torch.from_numpy(traj.datavec).to(dev)
mu = torch.from_numpy(wl.random_multipliers(nb * D.dyn)).to(dev)
F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev)
J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
H = torch.empty(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)
stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "the library's own stream"
torch.cuda.set_stream(stream)
st = stream.cuda_stream
bytes_per = 8 * (2 * D.zdim + 2 * D.dyn + D.nnzJ + D.nnzH)
for mask, label in ([(7, "F+J+H"), (1, "F"), (2, "J"), (4, "H"), (3, "F+J")] if hess else [(3, "F+J"), (1, "F")]):
    for _ in range(3):
        D.eval_device(mask, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        D.eval_device(mask, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    line = f"{label:6s}: {ms*1e3:9.1f} us/pass  {nb/ms*1e-3:8.2f} M evals/s"
    if mask == 7 or (not hess and mask == 3):
        line += f"  {bytes_per*nb/ms*1e-6:8.1f} GB/s algorithmic"
    print(line)
