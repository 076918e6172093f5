"""End-to-end host-buffer timing (development tool; bench.py is the contract): qck_eval_all with pageable numpy arrays,
two alternating trajectories so that every step uploads and evaluates (no cache hits)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcknot
from qcknot import workloads as wl

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
systems, traj, integrators = wl.config("cz", T=T)
D = qcknot.QuantumDynamics(integrators, traj, n_gpus=n_gpus, shard_mode="knot")
nb = D.n_blocks
Z = [traj.datavec.copy(), traj.datavec.copy()]
Z[1] += 1e-6
mu = [wl.random_multipliers(nb * D.dyn, seed=s) for s in (1, 2)]
F, J, H = np.empty(nb * D.dyn), np.empty(nb * D.nnzJ), np.empty(nb * D.nnzH)
import torch
hb = torch.empty(32 << 20, dtype=torch.float64).pin_memory(); db = torch.empty(32 << 20, dtype=torch.float64, device="cuda:0")
best = 0.0
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); hb.copy_(db, non_blocking=True); torch.cuda.synchronize()
    best = max(best, hb.numel() * 8 / (time.perf_counter() - t0) * 1e-9)
print(f"box: {os.cpu_count()} cpus, pinned D2H {best:.1f} GB/s")
del hb, db
for i in range(3):
    D.eval_all(Z[i & 1], mu[i & 1], F, J, H)
t0 = time.perf_counter()
for i in range(steps):
    D.eval_all(Z[i & 1], mu[i & 1], F, J, H)
dt = (time.perf_counter() - t0) / steps
st = D.transfer_stats()
print(f"T={T} gpus={n_gpus}: eval_all {dt*1e3:.2f} ms/step  {nb/dt*1e-6:.3f} M evals/s  h2d {st['h2d_bytes']*1e-6:.1f} MB d2h {st['d2h_bytes']*1e-6:.1f} MB "
      f"pcie {(st['h2d_bytes']+st['d2h_bytes'])/dt*1e-9:.1f} GB/s  out {(F.nbytes+J.nbytes+H.nbytes)/dt*1e-9:.1f} GB/s host-written")
# the same with page-locked (registered) output arrays: the library's direct path (no staging, no pack kernel)
for a in (F, J, H):
    qcknot.host_register(a)
for i in range(3):
    D.eval_all(Z[i & 1], mu[i & 1], F, J, H)
t0 = time.perf_counter()
for i in range(steps):
    D.eval_all(Z[i & 1], mu[i & 1], F, J, H)
dt = (time.perf_counter() - t0) / steps
print(f"  registered outputs: eval_all {dt*1e3:.2f} ms/step  {nb/dt*1e-6:.3f} M evals/s  d2h {D.transfer_stats()['d2h_bytes']*1e-6:.1f} MB")
for a in (F, J, H):
    qcknot.host_unregister(a)
for name, fn in (("F", lambda i: D.F(Z[i & 1], out=F)), ("J", lambda i: D.dF(Z[i & 1], out=J)), ("H", lambda i: D.mu_d2F(Z[i & 1], mu[i & 1], out=H))):
    fn(0); fn(1)
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    print(f"  {name} alone: {(time.perf_counter()-t0)/steps*1e3:.2f} ms/call")
# Ipopt-like sequence on one Z: F, J, H
t0 = time.perf_counter()
for i in range(steps):
    D.F(Z[i & 1], out=F); D.dF(Z[i & 1], out=J); D.mu_d2F(Z[i & 1], mu[i & 1], out=H)
print(f"  F,J,H callbacks in succession: {(time.perf_counter()-t0)/steps*1e3:.2f} ms per triple")
