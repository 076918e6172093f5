"""Device-resident F+J+H timing over the integrator classes the dispatcher distinguishes (development tool): finds classes whose
algorithmic GB/s falls far behind the tuned paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qcknot
from qcknot import workloads as wl

dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def run(label, systems, traj, integrators):
    D = qcknot.QuantumDynamics(integrators, traj)
    nb = D.n_blocks
    Z = torch.from_numpy(traj.datavec).to(dev)
    mu = torch.from_numpy(wl.random_multipliers(nb * D.dyn)).to(dev)
    F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev)
    J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
    H = torch.empty(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)
    args = (Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), stream.cuda_stream)
    for _ in range(3):
        D.eval_device(7, *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        D.eval_device(7, *args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpe = 8 * (2 * D.zdim + 2 * D.dyn + D.nnzJ + D.nnzH)
    print(f"{label:58s} {ms*1e3:9.1f} us  {nb/ms*1e-3:9.3f} M evals/s  {bpe*nb/ms*1e-6:8.1f} GB/s  ({bpe} B/eval, {nb} knots)", flush=True)
    D.close()


def rnd(levels, nd, T, integrator="pade", order=4, ket=False, seed=1, n_states=1):
    sys_ = wl.random_hermitian_system(levels, nd, seed=seed, scale=0.4)
    traj = wl.random_pulse_trajectory([sys_], T, 0.2, seed=seed, ket=ket, n_states=n_states)
    return [sys_], traj, wl.build_integrators([sys_], traj, integrator=integrator, order=order, ket=ket)


def main():
    for order in (6, 8, 12):
        systems, traj, _ = wl.config("cz", T=4000)
        run(f"cz N=9 pade order {order}", systems, traj, wl.build_integrators(systems, traj, order=order))
        systems, traj, _ = wl.config("hadamard", T=50000)
        run(f"hadamard N=2 pade order {order}", systems, traj, wl.build_integrators(systems, traj, order=order))
    for N in (5, 6, 7, 8):
        run(f"dense N={N} nd=2 pade-4", *rnd(N, 2, 8000))
        run(f"dense N={N} nd=2 exponential", *rnd(N, 2, 4000, integrator="exponential"))
    run("dense N=9 nd=2 ket pade-4 (2 kets)", *rnd(9, 2, 20000, ket=True, n_states=2))
    run("dense N=9 nd=2 ket exponential (2 kets)", *rnd(9, 2, 8000, integrator="exponential", ket=True, n_states=2))
    run("dense N=9 nd=5 pade-4 (tiled: > 4 drives)", *rnd(9, 5, 4000))
    run("dense N=12 nd=2 pade-4", *rnd(12, 2, 3000))
    run("dense N=12 nd=2 exponential", *rnd(12, 2, 1000, integrator="exponential"))
    sy = [wl.random_hermitian_system(9, 2, seed=s, scale=0.4) for s in range(4)]
    tr = wl.random_pulse_trajectory(sy, 2000, 0.2, seed=3)
    run("ensemble 4 x N=9 nd=2 pade-4 (shared controls)", sy, tr, wl.build_integrators(sy, tr))
    run("ensemble 4 x N=9 nd=2 exponential (shared controls)", sy, tr, wl.build_integrators(sy, tr, integrator="exponential"))


if __name__ == "__main__":
    main()
