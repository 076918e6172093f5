"""QuantumDynamics(integrators, traj): the reference's five-field dynamics object, evaluated on a B200.

Reference surface (/root/reference/test/scripts/integrator_test_1qubit.jl:41-52):
    dynamics = QuantumDynamics(f, Z)
    dynamics.F(Z.datavec)
    dynamics.∂F(Z.datavec), dynamics.∂F_structure
    dynamics.μ∂²F(Z.datavec, μ), dynamics.μ∂²F_structure
Python identifiers cannot contain ∂/²; the fields are spelled F, dF, dF_structure, mu_d2F, mu_d2F_structure.

All arithmetic happens in libqcknot.so (hand-written sm_100a kernels) through the C-ABI in include/qcknot.h.
There is no CPU path: construction fails without the library or without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .integrators import AbstractIntegrator, DerivativeIntegrator, _QuantumIntegrator
from .trajectory import NamedTrajectory


class QcknotError(RuntimeError):
    pass


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class QuantumDynamics:
    """Stacks `integrators` over the knot points of `traj` (row order = integrator order, values knot-major).

    knot_range=(t0, t1)       evaluate only the constraint blocks t0 <= t < t1 (0-based; one-knot halo is read);
                              F/dF/mu_d2F then return that shard's contiguous segment of the global arrays.
    integrator_range=(q0,q1)  manual ensemble sharding: evaluate only integrators q0 <= q < q1 (structures stay global;
                              q0 == q1 is an empty shard that launches nothing).
    n_gpus, shard_mode        n_gpus > 1: this ONE object drives GPUs device .. device+n_gpus-1 (or `devices`), the knot blocks
                              ("knot") or the quantum integrators ("ensemble") partitioned inside libqcknot; F/dF/mu_d2F fill
                              the caller's single arrays exactly as with one GPU.
    structure_order           intra-knot order of the structure entries and value arrays: "csc" (default: union pattern,
                              column-major), "row_major" (union pattern, row-major) or "per_integrator" (the integrators' own
                              lists one after the other; shared Hessian positions appear once per integrator and the consumer
                              sums the duplicates, test/test_utils.jl:14-27).  The Core's own order is not visible in the
                              reference repository, so it is a policy of qck_create rather than a constant.
    """

    def __init__(
        self,
        integrators: Sequence[AbstractIntegrator],
        traj: NamedTrajectory,
        eval_hessian: bool = True,
        device: int = 0,
        knot_range: Optional[Tuple[int, int]] = None,
        integrator_range: Optional[Tuple[int, int]] = None,
        n_gpus: int = 1,
        shard_mode: str = "knot",
        devices: Optional[Sequence[int]] = None,
        host_threads: int = 0,
        structure_order: str = "csc",
    ):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.integrators = list(integrators)
        self.zdim = traj.dim
        self.T_total = traj.T
        t0, t1 = knot_range if knot_range is not None else (0, traj.T - 1)
        if not (0 <= t0 < t1 <= traj.T - 1):
            raise ValueError(f"knot_range {knot_range} outside [0, {traj.T - 1}]")
        self.knot_range = (int(t0), int(t1))
        self.T = t1 - t0 + 1  # knot points of this shard
        self.eval_hessian = bool(eval_hessian)
        self.free_time = traj.free_time
        dt_off = traj.components[traj.timestep].start if traj.free_time else -1
        dt_fixed = 0.0 if traj.free_time else float(traj.timestep)

        descs = (_lib.IntegratorDesc * len(self.integrators))()
        self._keep = []
        for d, I in zip(descs, self.integrators):
            d.kind, d.order = I.kind, getattr(I, "order", 0)
            if isinstance(I, _QuantumIntegrator):
                if I.freetime != traj.free_time:
                    raise ValueError("integrator was built for a trajectory with a different timestep mode")
                hd, hv = I.system.drift_reim(), I.system.drives_reim()
                self._keep += [hd, hv]
                d.levels, d.n_drives = I.system.levels, I.system.n_drives
                d.state_off, d.state_len = I.state_components.start, len(I.state_components)
                d.ctrl_off = I.drive_components.start
                d.H_drift = hd.ctypes.data_as(C.POINTER(C.c_double))
                d.H_drives = hv.ctypes.data_as(C.POINTER(C.c_double)) if hv.size else None
            elif isinstance(I, DerivativeIntegrator):
                d.state_off, d.state_len = I.x_components.start, len(I.x_components)
                d.ctrl_off = I.dx_components.start
            else:
                raise TypeError(f"unsupported integrator {type(I).__name__}")
        q0, q1 = integrator_range if integrator_range is not None else (0, -1)  # integ_end < 0: every integrator
        self.integrator_range = (q0, q1) if integrator_range is not None else (0, len(self.integrators))
        if shard_mode not in ("knot", "ensemble"):
            raise ValueError("shard_mode must be 'knot' or 'ensemble'")
        self.n_gpus = max(1, int(n_gpus))
        self.shard_mode = shard_mode
        devs = None
        if devices is not None:
            if len(devices) != self.n_gpus:
                raise ValueError("devices must list n_gpus ordinals")
            devs = (C.c_int32 * self.n_gpus)(*[int(x) for x in devices])
            self._keep.append(devs)
        orders = {"csc": _lib.QCK_ORDER_CSC, "row_major": _lib.QCK_ORDER_ROW_MAJOR, "per_integrator": _lib.QCK_ORDER_PER_INTEGRATOR}
        if structure_order not in orders:
            raise ValueError(f"structure_order must be one of {sorted(orders)}")
        self.structure_order = structure_order
        pd = _lib.ProblemDesc(self.T, self.zdim, dt_off, dt_fixed, len(self.integrators), int(self.eval_hessian),
                              int(device), int(q0), int(q1), self.n_gpus, descs,
                              _lib.QCK_SHARD_ENSEMBLE if shard_mode == "ensemble" else _lib.QCK_SHARD_KNOT,
                              int(host_threads), devs, orders[structure_order], 0)
        rc = self._lib.qck_create(C.byref(pd), C.byref(self._h))
        if rc != 0:
            msg = self._lib.qck_last_error(None).decode()
            self._h = C.c_void_p()
            raise QcknotError(f"qck_create failed ({rc}): {msg}")
        dyn, nj, nh = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.qck_sizes(self._h, C.byref(dyn), C.byref(nj), C.byref(nh)))
        self.dyn, self.nnzJ, self.nnzH = dyn.value, nj.value, nh.value
        self.device = int(device)
        self._J_struct = None
        self._H_struct = None

    # ------------------------------------------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise QcknotError(f"libqcknot error {rc}: {self._lib.qck_last_error(self._h).decode()}")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.qck_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_blocks(self) -> int:
        return self.T - 1

    def _Z(self, Z) -> np.ndarray:
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        t0, t1 = self.knot_range
        if Z.size >= self.T_total * self.zdim and self.T != self.T_total:
            Z = Z[t0 * self.zdim : (t1 + 1) * self.zdim]  # full datavec given to a shard
        elif Z.size >= self.T * self.zdim:
            Z = Z[: self.T * self.zdim]  # drop global (free-phase) variables: they never enter the dynamics
        else:
            raise ValueError(f"Z has {Z.size} entries, expected at least {self.T * self.zdim}")
        return np.ascontiguousarray(Z)

    def _mu(self, mu) -> np.ndarray:
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        t0, t1 = self.knot_range
        if mu.size == (self.T_total - 1) * self.dyn and self.T != self.T_total:
            mu = mu[t0 * self.dyn : t1 * self.dyn]
        if mu.size != self.n_blocks * self.dyn:
            raise ValueError(f"mu has {mu.size} entries, expected {self.n_blocks * self.dyn}")
        return np.ascontiguousarray(mu)

    # ---- the reference's five fields -------------------------------------------------------------------------------
    def F(self, Z, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty(self.n_blocks * self.dyn) if out is None else out
        self._check(self._lib.qck_eval_residual(self._h, _ptr(self._Z(Z)), _ptr(out)))
        return out

    def dF(self, Z, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty(self.n_blocks * self.nnzJ) if out is None else out
        self._check(self._lib.qck_eval_jacobian(self._h, _ptr(self._Z(Z)), _ptr(out)))
        return out

    def mu_d2F(self, Z, mu, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty(self.n_blocks * self.nnzH) if out is None else out
        self._check(self._lib.qck_eval_hessian(self._h, _ptr(self._Z(Z)), _ptr(self._mu(mu)), _ptr(out)))
        return out

    def eval_all(self, Z, mu=None, F=None, J=None, H=None):
        """Fused residual + Jacobian + Hessian in one device pass (any output may be omitted with False)."""
        F = np.empty(self.n_blocks * self.dyn) if F is None else (None if F is False else F)
        J = np.empty(self.n_blocks * self.nnzJ) if J is None else (None if J is False else J)
        want_h = self.eval_hessian and mu is not None and H is not False
        H = (np.empty(self.n_blocks * self.nnzH) if H is None else H) if want_h else None
        mu_ = self._mu(mu) if want_h else None
        self._check(self._lib.qck_eval_all(self._h, _ptr(self._Z(Z)), _ptr(mu_), _ptr(F), _ptr(J), _ptr(H)))
        return F, J, H

    def _structure(self, fn, nnz) -> np.ndarray:
        n = self.n_blocks * nnz
        rows, cols = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        self._check(fn(self._h, self.knot_range[0], _ptr(rows), _ptr(cols)))
        return np.stack([rows, cols], axis=1)

    @property
    def dF_structure(self) -> np.ndarray:
        """(n, 2) int64 array of 1-based (row, col) pairs: the reference's Vector{Tuple{Int,Int}}."""
        if self._J_struct is None:
            self._J_struct = self._structure(self._lib.qck_jacobian_structure, self.nnzJ)
        return self._J_struct

    @property
    def mu_d2F_structure(self) -> np.ndarray:
        if self._H_struct is None:
            self._H_struct = self._structure(self._lib.qck_hessian_structure, self.nnzH)
        return self._H_struct

    # ---- device-resident path (bench, multi-GPU) ----------------------------------------------------------------------
    def eval_device(self, mask: int, dZ: int, dmu: int, dF: int, dJ: int, dH: int, stream: int = 0) -> None:
        """Enqueue one pass on device pointers (ints, e.g. torch.Tensor.data_ptr()) on CUDA stream `stream`."""
        self._check(self._lib.qck_eval_device(self._h, mask, dZ or None, dmu or None, dF or None, dJ or None,
                                              dH or None, stream or None))

    def device_buffers(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        self._check(self._lib.qck_device_buffers(self._h, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def synchronize(self) -> None:
        self._check(self._lib.qck_synchronize(self._h))

    def shared_hessian_positions(self) -> np.ndarray:
        n = C.c_int64()
        self._check(self._lib.qck_shared_hessian_positions(self._h, C.byref(n), None))
        pos = np.empty(n.value, dtype=np.int64)
        if n.value:
            self._check(self._lib.qck_shared_hessian_positions(self._h, C.byref(n), _ptr(pos)))
        return pos

    # ---- multi-GPU handles -------------------------------------------------------------------------------------------
    def shards(self):
        """[(device, block_begin, block_end, integ_begin, integ_end)] of every GPU behind this object."""
        n = C.c_int32()
        self._check(self._lib.qck_shard_count(self._h, C.byref(n)))
        out = []
        for g in range(n.value):
            dev, ib, ie = C.c_int32(), C.c_int32(), C.c_int32()
            b0, b1 = C.c_int64(), C.c_int64()
            self._check(self._lib.qck_shard_info(self._h, g, C.byref(dev), C.byref(b0), C.byref(b1), C.byref(ib), C.byref(ie)))
            out.append((dev.value, b0.value, b1.value, ib.value, ie.value))
        return out

    def upload(self, Z, mu=None) -> None:
        Z_ = self._Z(Z)
        mu_ = self._mu(mu) if mu is not None else None
        self._check(self._lib.qck_upload(self._h, _ptr(Z_), _ptr(mu_)))

    def eval_resident(self, mask: int = 7) -> None:
        self._check(self._lib.qck_eval_resident(self._h, mask))

    def gather_device(self, mask: int = 7) -> None:
        self._check(self._lib.qck_gather_device(self._h, mask))

    def shard_device_buffers(self, g: int):
        ptrs = [C.c_void_p() for _ in range(5)]
        self._check(self._lib.qck_shard_device_buffers(self._h, g, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def gathered_buffers(self, g: int):
        ptrs = [C.c_void_p() for _ in range(3)]
        self._check(self._lib.qck_gathered_buffers(self._h, g, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def nccl_version(self):
        v, n = C.c_int32(), C.c_int32()
        self._check(self._lib.qck_nccl_version(self._h, C.byref(v), C.byref(n)))
        return v.value, n.value

    def transfer_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.qck_transfer_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value, "cache_hits": c.value}

    def compact_map(self, arr: int) -> np.ndarray:
        """(n, 4) int32: (offset in the knot block, offset in the compact layout, length, repeats) of value array arr."""
        n = C.c_int64()
        self._check(self._lib.qck_compact_map(self._h, arr, C.byref(n), None))
        segs = np.empty((n.value, 4), dtype=np.int32)
        if n.value:
            self._check(self._lib.qck_compact_map(self._h, arr, C.byref(n), _ptr(segs)))
        return segs

    def expand_host(self, arr: int, compact: np.ndarray, out: np.ndarray, nk: int) -> None:
        self._check(self._lib.qck_expand_host(self._h, arr, _ptr(compact), _ptr(out), nk))

    # ---- objective / terminal-constraint terms --------------------------------------------------------------
    def attach_objective(self, J) -> None:
        """J: an objectives.Objective (sum of terms); evaluated on the device from the same Z as the dynamics."""
        from .objectives import term_array
        arr, keep = term_array(J)
        self._check(self._lib.qck_objective_attach(self._h, arr, len(J.terms)))
        n, nh = C.c_int64(), C.c_int64()
        self._check(self._lib.qck_objective_sizes(self._h, C.byref(n), C.byref(nh)))
        self.n_vars, self.nnz_obj_hess = n.value, nh.value
        self._obj_struct = None

    def objective(self, Z) -> float:
        v = C.c_double()
        self._check(self._lib.qck_eval_objective(self._h, _ptr(self._Z(Z)), C.byref(v)))
        return v.value

    def objective_gradient(self, Z, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty(self.n_vars) if out is None else out
        self._check(self._lib.qck_eval_objective_gradient(self._h, _ptr(self._Z(Z)), _ptr(out)))
        return out

    def objective_hessian(self, Z, sigma: float = 1.0, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty(self.nnz_obj_hess) if out is None else out
        self._check(self._lib.qck_eval_objective_hessian(self._h, _ptr(self._Z(Z)), float(sigma), _ptr(out)))
        return out

    @property
    def objective_hessian_structure(self) -> np.ndarray:
        if self._obj_struct is None:
            rows, cols = np.empty(self.nnz_obj_hess, dtype=np.int64), np.empty(self.nnz_obj_hess, dtype=np.int64)
            self._check(self._lib.qck_objective_hessian_structure(self._h, _ptr(rows), _ptr(cols)))
            self._obj_struct = np.stack([rows, cols], axis=1)
        return self._obj_struct

    def attach_fidelity_constraint(self, con) -> None:
        from .objectives import term_array
        arr, keep = term_array(con)
        self._check(self._lib.qck_fidelity_constraint_attach(self._h, arr, con.val))
        self._con_len = len(con.comp)

    def fidelity_constraint(self, Z, mu: Optional[float] = None):
        """(g, jacobian row[, mu * Hessian upper triangle by column])."""
        g = C.c_double()
        jac = np.empty(self._con_len)
        hess = np.empty(self._con_len * (self._con_len + 1) // 2) if mu is not None else None
        self._check(self._lib.qck_eval_fidelity_constraint(self._h, _ptr(self._Z(Z)), float(mu or 0.0), C.byref(g), _ptr(jac), _ptr(hess)))
        return (g.value, jac) if mu is None else (g.value, jac, hess)

    def invalidate(self) -> None:
        self._check(self._lib.qck_invalidate(self._h))

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.qck_launch_count(self._h, C.byref(n)))
        return n.value


def host_register(a: np.ndarray) -> None:
    """Page-lock a numpy array (e.g. the solver's value buffer) for full-speed PCIe copies."""
    rc = _lib.load().qck_host_register(_ptr(a), a.nbytes)
    if rc != 0:
        raise QcknotError(f"qck_host_register failed: {_lib.load().qck_last_error(None).decode()}")


def host_unregister(a: np.ndarray) -> None:
    _lib.load().qck_host_unregister(_ptr(a))


def dense(vals, structure, shape) -> np.ndarray:
    """test/test_utils.jl:14-27: matrix from (values, structure); duplicates sum; square => symmetric upper."""
    M = np.zeros(shape)
    s = np.asarray(structure)
    np.add.at(M, (s[:, 0] - 1, s[:, 1] - 1), np.asarray(vals))
    if shape[0] == shape[1]:
        return np.triu(M) + np.triu(M, 1).T
    return M
