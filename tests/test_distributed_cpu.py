"""world_size-2 gloo test of the N>1 host logic: knot-range shards evaluated independently assemble, through the
all-gather plumbing, into exactly the single-process arrays; ensemble shards assemble through the all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qcknot
    from oracle.bridge import oracle_dynamics
    from qcknot import workloads as wl
    from qcknot.sharding import all_gather_segments, all_reduce_shared, integrator_shard, knot_shard, knot_shards

    # ---- knot sharding: each rank evaluates its block range on its own slice (one-knot halo) with the CPU oracle --------
    systems, traj, integrators = wl.config("hadamard", T=10)
    full = oracle_dynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers((traj.T - 1) * full.dyn)
    t0, t1 = knot_shard(traj.T - 1, rank, world)
    S = qcknot.QuantumDynamics(integrators, traj, device=-1, knot_range=(t0, t1))  # host-side shard logic + structure
    comps = {n: (r.start, len(r)) for n, r in traj.components.items()}
    sub = oracle_dynamics(integrators, type("T", (), dict(components=traj.components, T=t1 - t0 + 1, timestep=traj.timestep,
                                                      free_time=traj.free_time, global_dim=0))())
    Zs, mus = S._Z(Z), S._mu(mu)
    J_local = torch.from_numpy(sub.dF(Zs))
    H_local = torch.from_numpy(sub.mu_d2F(Zs, mus))
    shards = knot_shards(traj.T - 1, world)
    J_all = all_gather_segments(J_local, [(b - a) * full.nnzJ for a, b in shards])
    H_all = all_gather_segments(H_local, [(b - a) * full.nnzH for a, b in shards])
    ok = np.array_equal(J_all.numpy(), full.dF(Z)) and np.array_equal(H_all.numpy(), full.mu_d2F(Z, mu))
    # structure segments line up with the value segments
    Js = torch.from_numpy(S.dF_structure.reshape(-1).copy())
    Js_all = all_gather_segments(Js, [(b - a) * full.nnzJ * 2 for a, b in shards]).numpy().reshape(-1, 2)
    ok = ok and np.array_equal(Js_all, np.array(full.dF_structure))

    # ---- ensemble sharding: each rank owns a slice of the systems; shared-control Hessian entries are all-reduced ---------
    systems, traj, integrators = wl.config("sampling", T=3, n_systems=4)
    full = oracle_dynamics(integrators, traj)
    Z = traj.datavec
    mu = wl.random_multipliers((traj.T - 1) * full.dyn)
    q0, q1 = integrator_shard(len(systems), len(integrators), rank, world)
    D = qcknot.QuantumDynamics(integrators, traj, device=-1, integrator_range=(q0, q1))
    shared = D.shared_hessian_positions()
    # oracle evaluation restricted to this rank's integrators, in the global layout
    Hloc = np.zeros(full.nnzH * (traj.T - 1))
    rr = np.array([r for r, _ in full.hess_knot])
    cc = np.array([c for _, c in full.hess_knot])
    for t in range(traj.T - 1):
        zt, zt1 = Z[t * full.zdim:(t + 1) * full.zdim], Z[(t + 1) * full.zdim:(t + 2) * full.zdim]
        Hm = np.zeros((2 * full.zdim, 2 * full.zdim))
        for I, r0 in list(zip(full.integrators, full.row_off))[q0:q1]:
            Hm += I.hessian(zt, zt1, mu[t * full.dyn + r0: t * full.dyn + r0 + I.dim])
        Hloc[t * full.nnzH:(t + 1) * full.nnzH] = Hm[rr, cc]
    Hsum = all_reduce_shared(torch.from_numpy(Hloc.copy()), torch.from_numpy(shared), full.nnzH)
    Href = full.mu_d2F(Z, mu)
    idx = (np.arange(traj.T - 1)[:, None] * full.nnzH + shared[None, :]).reshape(-1)
    # the shared entries are complete on every rank, the others untouched (each is written by exactly one rank) ...
    ok = ok and np.allclose(Hsum.numpy()[idx], Href[idx], rtol=0, atol=1e-14) and len(shared) == 6
    rest = np.setdiff1d(np.arange(Href.size), idx)
    ok = ok and np.array_equal(Hsum.numpy()[rest], Hloc[rest])
    # ... and the disjoint parts of all ranks assemble to the full array (test plumbing: sum with the shared part counted once)
    asm = Hsum.clone()
    if rank != 0:
        asm[torch.from_numpy(idx)] = 0.0
    dist.all_reduce(asm)
    ok = ok and np.allclose(asm.numpy(), Href, rtol=0, atol=1e-14)
    # a shard with no integrator at all is legal (begin == end) and 'every integrator' is spelled integ_end < 0
    E = qcknot.QuantumDynamics(integrators, traj, device=-1, integrator_range=(2, 2))
    ok = ok and E.integrator_range == (2, 2) and E.shards()[0][3:] == (2, 2)
    try:
        integrator_shard(2, 4, 0, 4)
        ok = False
    except ValueError:
        pass
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
