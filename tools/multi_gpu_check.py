"""Multi-GPU functional check over NCCL (run under torchrun, one rank per GPU):

  1. knot sharding: every rank evaluates its block range on its own GPU (device-resident), the value
     segments are all-gathered over NVLink and compared with the single-GPU arrays;
  2. ensemble sharding: every rank evaluates its slice of the sampled systems, the zero-initialised value arrays are
     summed with one all-reduce (disjoint entries assemble, shared-control Hessian entries add up) and compared.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import qcknot
from qcknot import workloads as wl
from qcknot.sharding import all_gather_segments, all_reduce_shared, integrator_shard, knot_shard, knot_shards

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
ok = True


def device_eval(D, Z, mu):
    nb = D.n_blocks
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        dZ, dmu = torch.from_numpy(Z).to(dev), torch.from_numpy(mu).to(dev)
        F = torch.zeros(nb * D.dyn, dtype=torch.float64, device=dev)
        J = torch.zeros(nb * D.nnzJ, dtype=torch.float64, device=dev)
        H = torch.zeros(nb * D.nnzH, dtype=torch.float64, device=dev)
        D.eval_device(7, dZ.data_ptr(), dmu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    return F, J, H


# ---- 1. knot sharding -----------------------------------------------------------------------------------------------
systems, traj, integrators = wl.config("cz", T=40 * world + 3)
nbt = traj.T - 1
Zfull = traj.datavec
mufull = wl.random_multipliers(nbt * 170)
t0, t1 = knot_shard(nbt, rank, world)
D = qcknot.QuantumDynamics(integrators, traj, device=local, knot_range=(t0, t1))
F, J, H = device_eval(D, D._Z(Zfull), D._mu(mufull))
shards = knot_shards(nbt, world)
Fall = all_gather_segments(F, [(b - a) * D.dyn for a, b in shards])
Jall = all_gather_segments(J, [(b - a) * D.nnzJ for a, b in shards])
Hall = all_gather_segments(H, [(b - a) * D.nnzH for a, b in shards])
ref = qcknot.QuantumDynamics(integrators, traj, device=local)
Fr, Jr, Hr = ref.eval_all(Zfull, mufull)
ok &= np.array_equal(Fall.cpu().numpy(), Fr) and np.array_equal(Jall.cpu().numpy(), Jr) and np.array_equal(Hall.cpu().numpy(), Hr)
if rank == 0:
    print(f"knot sharding over {world} GPUs (NCCL all-gather of contiguous segments): {'OK' if ok else 'MISMATCH'}")

# ---- 2. ensemble sharding ---------------------------------------------------------------------------------------------
systems, traj, integrators = wl.config("sampling", T=12, n_systems=4 * world)
Z = traj.datavec
q0, q1 = integrator_shard(len(systems), len(integrators), rank, world)
D = qcknot.QuantumDynamics(integrators, traj, device=local, integrator_range=(q0, q1))
mu = wl.random_multipliers(D.n_blocks * D.dyn)
F, J, H = device_eval(D, Z, mu)
shared = torch.from_numpy(D.shared_hessian_positions()).to(dev)
for x in (F, J):
    dist.all_reduce(x)  # disjoint ownership: a sum assembles
H = all_reduce_shared(H, shared, D.nnzH)
ref = qcknot.QuantumDynamics(integrators, traj, device=local)
Fr, Jr, Hr = ref.eval_all(Z, mu)
e = max(np.abs(F.cpu().numpy() - Fr).max(), np.abs(J.cpu().numpy() - Jr).max(), np.abs(H.cpu().numpy() - Hr).max() / max(1.0, np.abs(Hr).max()))
ok2 = e < 1e-13
if rank == 0:
    print(f"ensemble sharding over {world} GPUs ({len(systems)} systems, NCCL all-reduce of shared control entries): {'OK' if ok2 else 'MISMATCH'} (max err {e:.2e})")
ok &= ok2
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
