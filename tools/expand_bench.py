"""Host half of the host-buffer path alone: compact layout -> structure-order arrays (no GPU needed).
Reports GB/s of caller-array bytes written; the e2e path can never be faster than this on the same host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcknot
from qcknot import workloads as wl

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
systems, traj, integrators = wl.config("cz", T=T)
D = qcknot.QuantumDynamics(integrators, traj, device=-1)
nk = D.n_blocks
tot_out = tot_in = 0
bufs = []
for arr, nnz in ((0, D.dyn), (1, D.nnzJ), (2, D.nnzH)):
    C_ = int(D.compact_map(arr)[:, 2].sum())
    bufs.append((arr, np.random.default_rng(arr).standard_normal(nk * C_), np.empty(nk * nnz)))
    tot_out += nk * nnz * 8
    tot_in += nk * C_ * 8
for rep in range(5):
    t0 = time.perf_counter()
    for arr, comp, out in bufs:
        D.expand_host(arr, comp, out, nk)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3:7.2f} ms  out {tot_out/dt*1e-9:6.1f} GB/s  (compact in {tot_in*1e-6:.0f} MB, out {tot_out*1e-6:.0f} MB, threads env QCK_HOST_THREADS={os.environ.get('QCK_HOST_THREADS','auto')}, cpus {os.cpu_count()})")
