"""Test helpers: build the oracle's view of a qcknot problem (same systems, same layout, same integrator order)."""
from __future__ import annotations

import numpy as np

import qcknot
from oracle import knot_oracle as ko
from qcknot import integrators as qi


def oracle_dynamics(integrators, traj, eval_hessian=True) -> ko.QuantumDynamics:
    comps = {n: (r.start, len(r)) for n, r in traj.components.items()}
    layout = ko.Layout(comps, traj.T, traj.timestep if traj.free_time else None,
                       0.0 if traj.free_time else traj.timestep, traj.global_dim)
    cls = {
        qi.UnitaryPadeIntegrator: ko.UnitaryPadeIntegrator,
        qi.UnitaryExponentialIntegrator: ko.UnitaryExponentialIntegrator,
        qi.QuantumStatePadeIntegrator: ko.QuantumStatePadeIntegrator,
        qi.QuantumStateExponentialIntegrator: ko.QuantumStateExponentialIntegrator,
    }
    out = []
    for I in integrators:
        if isinstance(I, qi.DerivativeIntegrator):
            out.append(ko.DerivativeIntegrator(I.x_name, I.dx_name, layout))
        else:
            sys_ = ko.QuantumSystem(I.system.H_drift, I.system.H_drives)
            kw = {"order": I.order} if hasattr(I, "order") and I.order else {}
            out.append(cls[type(I)](I.state_name, I.control_name, sys_, layout, **kw))
    return ko.QuantumDynamics(out, layout, eval_hessian=eval_hessian)


def rel_err(a, b) -> float:
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0
