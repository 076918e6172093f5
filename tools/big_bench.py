"""Device-resident timing of the large-level kernel (qck_big.cu): single transmons with 16 / 24 / 32 levels (two sparse drives,
src/quantum_system_templates/transmons.jl:76-85) and a dense random 32-level system with four drives (worst case for the
sparse-row drive products)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qcknot
from qcknot import workloads as wl

cases = [("transmon", 16, 2, 2000), ("transmon", 24, 2, 1000), ("transmon", 32, 2, 600), ("dense", 16, 4, 2000), ("dense", 32, 4, 600)]
for kind, N, nd, T in cases:
    sys_ = wl.transmon_system(levels=N) if kind == "transmon" else wl.random_hermitian_system(N, nd, seed=N + nd, scale=0.3)
    traj = wl.random_pulse_trajectory([sys_], T, 0.1, seed=2, a_bound=0.05)
    D = qcknot.QuantumDynamics(wl.build_integrators([sys_], traj), traj)
    nb = D.n_blocks
    dev = torch.device("cuda:0")
    Z = torch.from_numpy(traj.datavec).to(dev); mu = torch.from_numpy(wl.random_multipliers(nb * D.dyn)).to(dev)
    F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev); J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
    H = torch.empty(nb * D.nnzH, dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    bpe = 8 * (2 * D.zdim + 2 * D.dyn + D.nnzJ + D.nnzH)
    for mask, label in ((7, "F+J+H"), (3, "F+J"), (4, "H")):
        for _ in range(2):
            D.eval_device(mask, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            D.eval_device(mask, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        extra = f"  {bpe*nb/ms*1e-6:7.0f} GB/s algorithmic ({bpe} B/eval)" if mask == 7 else ""
        print(f"{kind:8s} N={N:2d} nd={nd} T={T}: {label:6s} {ms*1e3:8.0f} us/pass {nb/ms*1e-3:7.3f} M evals/s{extra}")
    D.close()
