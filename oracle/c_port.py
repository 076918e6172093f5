"""ctypes driver of oracle/libknot_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY -- see knot_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import knot_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libknot_oracle.so")


class _Integ(C.Structure):
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("n_drives", C.c_int), ("state_off", C.c_int),
                ("ctrl_off", C.c_int), ("row_off", C.c_int), ("dim", C.c_int),
                ("G_drift", C.POINTER(C.c_double)), ("G_drives", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "knot_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        subprocess.run(["make", "-C", HERE, "-B", "libknot_oracle.so"], check=True, capture_output=True)
    return LIB


class CPort:
    """Evaluates an oracle QuantumDynamics (Pade order 4 + derivative integrators only) with the C port."""

    def __init__(self, dyn: ko.QuantumDynamics):
        build()
        self.lib = C.CDLL(LIB)
        self.dyn = dyn
        L = dyn.layout
        self._keep = []
        arr = (_Integ * len(dyn.integrators))()
        for d, I, r0 in zip(arr, dyn.integrators, dyn.row_off):
            if isinstance(I, ko.DerivativeIntegrator):
                d.kind, d.state_off, d.ctrl_off, d.row_off, d.dim = 4, I.x.start, I.dx.start, int(r0), I.dim
                continue
            if not isinstance(I, (ko.UnitaryPadeIntegrator, ko.QuantumStatePadeIntegrator)) or I.order != 4:
                raise NotImplementedError("the C port covers Pade order 4 and DerivativeIntegrator")
            gd = np.asfortranarray(I.sys.G_drift).reshape(-1, order="F").copy()
            gj = np.concatenate([np.asfortranarray(g).reshape(-1, order="F") for g in I.sys.G_drives]).copy()
            self._keep += [gd, gj]
            d.kind = 0 if I.is_unitary else 2
            d.N, d.n_drives, d.state_off, d.ctrl_off = I.N, I.n_drives, I.state.start, I.ctrl.start
            d.row_off, d.dim = int(r0), I.dim
            d.G_drift = gd.ctypes.data_as(C.POINTER(C.c_double))
            d.G_drives = gj.ctypes.data_as(C.POINTER(C.c_double))
        self.arr = arr
        self.Jr = np.array([r for r, _ in dyn.jac_knot], dtype=np.int32)
        self.Jc = np.array([c for _, c in dyn.jac_knot], dtype=np.int32)
        self.Hr = np.array([r for r, _ in dyn.hess_knot], dtype=np.int32)
        self.Hc = np.array([c for _, c in dyn.hess_knot], dtype=np.int32)
        self.dt_off = L.components[L.dt_name][0] if L.dt_name else -1
        self.lib.ko_eval2.restype = C.c_int

    def eval(self, Z, mu=None, want=("F", "J", "H"), nthreads: int = 0, T=None, tuned: bool = False, out=None):
        """tuned: constant anticommutators hoisted out of the knot loop and G_j G + G G_j reused (see knot_oracle.c);
        out = (F, J, H) preallocated arrays (a timed loop must not pay for page faults of fresh arrays)."""
        d = self.dyn
        T = d.T if T is None else T
        nb = T - 1
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        if out is not None:
            F, J, H = out
        else:
            F = np.empty(nb * d.dyn) if "F" in want else None
            J = np.empty(nb * d.nnzJ) if "J" in want else None
            H = np.empty(nb * d.nnzH) if ("H" in want and d.eval_hessian) else None
        mu = np.ascontiguousarray(mu, dtype=np.float64) if mu is not None else None
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.ko_eval2(C.c_int(len(d.integrators)), self.arr, C.c_long(T), C.c_int(d.zdim), C.c_int(self.dt_off),
                               C.c_double(d.layout.dt_fixed), p(Z), p(mu), C.c_int(d.dyn), C.c_long(d.nnzJ), p(self.Jr),
                               p(self.Jc), C.c_long(d.nnzH), p(self.Hr), p(self.Hc), p(F), p(J), p(H), C.c_int(nthreads),
                               C.c_int(1 if tuned else 0))
        if rc != 0:
            raise MemoryError("ko_eval failed")
        return F, J, H
