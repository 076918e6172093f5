"""Minimal NamedTrajectory mirror: the input container of the hot path.

Layout facts pinned by the reference: `data` is dim x T, `datavec = vec(data)` (column-major), components occupy
contiguous row ranges in declaration order (/root/reference/test/test_utils.jl:52-118 ;
src/trajectory_initialization.jl:357-381)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence, Union

import numpy as np


class NamedTrajectory:
    def __init__(
        self,
        components: Dict[str, np.ndarray],
        controls: Sequence[str] = (),
        timestep: Union[str, float, None] = None,
        bounds: Optional[dict] = None,
        initial: Optional[dict] = None,
        final: Optional[dict] = None,
        goal: Optional[dict] = None,
        global_data: Optional[Dict[str, np.ndarray]] = None,
    ):
        if not components:
            raise ValueError("a trajectory needs at least one component")
        self.names = tuple(components.keys())
        mats = []
        self.components: Dict[str, range] = {}
        off = 0
        T = None
        for name, val in components.items():
            a = np.atleast_2d(np.asarray(val, dtype=np.float64))
            if T is None:
                T = a.shape[1]
            elif a.shape[1] != T:
                raise ValueError(f"component {name} has {a.shape[1]} knots, expected {T}")
            self.components[name] = range(off, off + a.shape[0])
            off += a.shape[0]
            mats.append(a)
        self.T = int(T)
        self.dim = off
        self.data = np.asfortranarray(np.vstack(mats))
        self.control_names = tuple(controls)
        if isinstance(timestep, str):
            if timestep not in self.components:
                raise ValueError(f"timestep component {timestep} not in trajectory")
            self.timestep = timestep
        elif timestep is None:
            raise ValueError("timestep must be a component name or a number")
        else:
            self.timestep = float(timestep)
        self.bounds = bounds or {}
        self.initial = initial or {}
        self.final = final or {}
        self.goal = goal or {}
        self.global_data = {k: np.asarray(v, dtype=np.float64).ravel() for k, v in (global_data or {}).items()}
        self.global_dim = int(sum(v.size for v in self.global_data.values()))
        n_controls = sum(len(self.components[c]) for c in self.control_names)
        self.dims = SimpleNamespace(states=self.dim - n_controls, controls=n_controls,
                                    **{n: len(r) for n, r in self.components.items()})

    @property
    def free_time(self) -> bool:
        return isinstance(self.timestep, str)

    @property
    def datavec(self) -> np.ndarray:
        """vec(data) followed by the global data (free phases sit after dim*T; they never enter the dynamics)."""
        z = self.data.reshape(-1, order="F")
        if self.global_dim:
            z = np.concatenate([z] + list(self.global_data.values()))
        return np.ascontiguousarray(z)

    def __getitem__(self, name: str) -> np.ndarray:
        r = self.components[name]
        return self.data[r.start : r.stop, :]

    def update(self, datavec) -> None:
        datavec = np.asarray(datavec, dtype=np.float64)
        self.data = np.asfortranarray(datavec[: self.dim * self.T].reshape(self.dim, self.T, order="F"))
