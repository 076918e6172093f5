"""Multi-GPU host logic: knot-range and ensemble sharding, one process per GPU over torch.distributed.

The knot blocks are independent given z_t, z_t+1, mu_t, and the value arrays are knot-major,
so rank g owns a contiguous segment of F, dF and mu_d2F.  No collective sits on the data path; an all-gather is used
only when one rank needs the assembled arrays, and an all-reduce only for Hessian entries on shared controls when
the system ensemble (unitary_sampling_problem.jl:134-155) is sharded."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def knot_shard(n_blocks: int, rank: int, world: int) -> Tuple[int, int]:
    """Balanced contiguous block range [t0, t1) of rank `rank`; the shard reads knots t0..t1 (one-knot halo)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_blocks < world:
        raise ValueError(f"{n_blocks} knot blocks cannot be sharded over {world} ranks")
    return n_blocks * rank // world, n_blocks * (rank + 1) // world


def knot_shards(n_blocks: int, world: int) -> List[Tuple[int, int]]:
    return [knot_shard(n_blocks, r, world) for r in range(world)]


def integrator_shard(n_quantum: int, n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Ensemble sharding: quantum integrators [q0, q1) of rank `rank`; the trailing non-quantum (derivative)
    integrators go to the last rank so that every integrator is evaluated exactly once."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_quantum < world:
        raise ValueError(f"{n_quantum} quantum integrators cannot be sharded over {world} ranks")
    q0, q1 = n_quantum * rank // world, n_quantum * (rank + 1) // world
    if rank == world - 1:
        q1 = n_total
    return q0, q1


def all_gather_segments(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """Assemble the global knot-major value array from per-rank contiguous segments (lengths `counts`)."""
    world = dist.get_world_size()
    if len(counts) != world or counts[dist.get_rank()] != local.numel():
        raise ValueError("segment lengths do not match the shards")
    m = max(counts)
    buf = torch.zeros(m, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def all_reduce_shared(H_local: torch.Tensor, shared_pos: torch.Tensor, nnzH: int) -> torch.Tensor:
    """Ensemble sharding, one process per GPU: sum the per-rank partial sums of the Hessian positions that several
    integrators share (a x a, a x dt, dt x dt) -- ONLY those, (T-1) * len(shared_pos) doubles -- in place.  Every other
    position of H_local is written by exactly one rank and is left alone."""
    nb = H_local.numel() // nnzH
    idx = (torch.arange(nb, device=H_local.device)[:, None] * nnzH + shared_pos.to(H_local.device)[None, :]).reshape(-1)
    buf = H_local[idx]
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    H_local[idx] = buf
    return H_local
