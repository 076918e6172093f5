"""Builds the oracle's view of a problem described with the product's host objects (same systems, same layout,
same integrator order).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu-baseline legs.  Duck-typed on class names so that the oracle never imports the product package."""
from __future__ import annotations

import numpy as np

from . import knot_oracle as ko

_CLS = {
    "UnitaryPadeIntegrator": ko.UnitaryPadeIntegrator,
    "UnitaryExponentialIntegrator": ko.UnitaryExponentialIntegrator,
    "QuantumStatePadeIntegrator": ko.QuantumStatePadeIntegrator,
    "QuantumStateExponentialIntegrator": ko.QuantumStateExponentialIntegrator,
    "DensityOperatorExponentialIntegrator": ko.QuantumStateExponentialIntegrator,  # ket integrator on N^2 levels, Lindbladian as the system
}


def oracle_dynamics(integrators, traj, eval_hessian: bool = True, structure_order: str = "csc") -> ko.QuantumDynamics:
    comps = {n: (r.start, len(r)) for n, r in traj.components.items()}
    layout = ko.Layout(comps, traj.T, traj.timestep if traj.free_time else None,
                       0.0 if traj.free_time else traj.timestep, traj.global_dim)
    out = []
    for I in integrators:
        name = type(I).__name__
        if name == "DerivativeIntegrator":
            out.append(ko.DerivativeIntegrator(I.x_name, I.dx_name, layout))
        else:
            sys_ = ko.QuantumSystem(I.system.H_drift, I.system.H_drives)
            kw = {"order": I.order} if getattr(I, "order", 0) else {}
            out.append(_CLS[name](I.state_name, I.control_name, sys_, layout, **kw))
    return ko.QuantumDynamics(out, layout, eval_hessian=eval_hessian, structure_order=structure_order)


def rel_err(a, b) -> float:
    """max |a-b| relative to max(1, max|b|): the 1e-10 criterion of BASELINE.json's north_star."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def entry_err(a, b, floor: float = 1e-6) -> float:
    """Per-entry relative error, max_i |a_i - b_i| / max(|b_i|, floor * max|b|): a bad small-magnitude term inside a large
    array shows up here even when rel_err (relative to the largest entry) hides it.  The absolute floor keeps entries that
    are rounding noise next to their neighbours (cancellation of O(max|b|) terms) from dominating."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if not a.size:
        return 0.0
    scale = float(np.max(np.abs(b)))
    if scale == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * scale)))
