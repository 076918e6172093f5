"""ctypes binding of libqcknot.so (include/qcknot.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCK_LIB") or os.path.join(HERE, "libqcknot.so")  # QCK_LIB: development override (A/B builds)

QCK_UNITARY_PADE, QCK_UNITARY_EXP, QCK_KET_PADE, QCK_KET_EXP, QCK_DERIVATIVE = range(5)
QCK_EVAL_F, QCK_EVAL_J, QCK_EVAL_H = 1, 2, 4
QCK_SHARD_KNOT, QCK_SHARD_ENSEMBLE = 0, 1
QCK_ORDER_CSC, QCK_ORDER_ROW_MAJOR, QCK_ORDER_PER_INTEGRATOR = 0, 1, 2
QCK_OBJ_QUADRATIC_REGULARIZER, QCK_OBJ_UNITARY_INFIDELITY, QCK_OBJ_MINIMUM_TIME = 0, 1, 2

EXPORTS = [
    "qck_create", "qck_destroy", "qck_last_error", "qck_sizes", "qck_jacobian_structure", "qck_hessian_structure",
    "qck_eval_residual", "qck_eval_jacobian", "qck_eval_hessian", "qck_eval_all", "qck_eval_device",
    "qck_device_buffers", "qck_synchronize", "qck_shared_hessian_positions", "qck_host_register",
    "qck_host_unregister", "qck_launch_count", "qck_version",
    "qck_shard_count", "qck_shard_info", "qck_shard_device_buffers", "qck_upload", "qck_eval_resident",
    "qck_gather_device", "qck_gathered_buffers", "qck_nccl_version", "qck_transfer_stats", "qck_invalidate",
    "qck_compact_map", "qck_expand_host",
    "qck_objective_attach", "qck_objective_sizes", "qck_objective_hessian_structure", "qck_eval_objective",
    "qck_eval_objective_gradient", "qck_eval_objective_hessian", "qck_fidelity_constraint_attach",
    "qck_eval_fidelity_constraint", "qck_rollout", "qck_rollout_last_error",
]


class IntegratorDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("order", C.c_int32), ("levels", C.c_int32), ("n_drives", C.c_int32),
        ("state_off", C.c_int32), ("state_len", C.c_int32), ("ctrl_off", C.c_int32), ("reserved", C.c_int32),
        ("H_drift", C.POINTER(C.c_double)), ("H_drives", C.POINTER(C.c_double)),
    ]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("T", C.c_int64), ("zdim", C.c_int32), ("dt_off", C.c_int32), ("dt_fixed", C.c_double),
        ("n_integrators", C.c_int32), ("eval_hessian", C.c_int32), ("device", C.c_int32),
        ("integ_begin", C.c_int32), ("integ_end", C.c_int32), ("n_gpus", C.c_int32),
        ("integrators", C.POINTER(IntegratorDesc)),
        ("shard_mode", C.c_int32), ("host_threads", C.c_int32), ("devices", C.POINTER(C.c_int32)),
        ("structure_order", C.c_int32), ("reserved", C.c_int32),
    ]


class ObjectiveTerm(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("comp_off", C.c_int32), ("comp_len", C.c_int32), ("levels", C.c_int32),
        ("weight", C.c_double), ("R", C.POINTER(C.c_double)), ("goal", C.POINTER(C.c_double)),
        ("n_sub", C.c_int32), ("reserved", C.c_int32),
    ]


_lib = None


def load() -> C.CDLL:
    """Load libqcknot.so; raises if it has not been built (python -m ... build / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU fallback for the knot-point evaluator."
        )
    lib = C.CDLL(LIB_PATH)
    dp, i64p, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
    lib.qck_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
    lib.qck_destroy.argtypes = [vp]
    lib.qck_destroy.restype = None
    lib.qck_last_error.argtypes = [vp]
    lib.qck_last_error.restype = C.c_char_p
    lib.qck_sizes.argtypes = [vp, i64p, i64p, i64p]
    lib.qck_jacobian_structure.argtypes = [vp, C.c_int64, vp, vp]
    lib.qck_hessian_structure.argtypes = [vp, C.c_int64, vp, vp]
    lib.qck_eval_residual.argtypes = [vp, vp, vp]
    lib.qck_eval_jacobian.argtypes = [vp, vp, vp]
    lib.qck_eval_hessian.argtypes = [vp, vp, vp, vp]
    lib.qck_eval_all.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.qck_eval_device.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, vp]
    lib.qck_device_buffers.argtypes = [vp] + [C.POINTER(vp)] * 5
    lib.qck_synchronize.argtypes = [vp]
    lib.qck_shared_hessian_positions.argtypes = [vp, i64p, vp]
    lib.qck_host_register.argtypes = [vp, C.c_size_t]
    lib.qck_host_unregister.argtypes = [vp]
    lib.qck_launch_count.argtypes = [vp, i64p]
    lib.qck_version.restype = C.c_char_p
    i32p = C.POINTER(C.c_int32)
    lib.qck_shard_count.argtypes = [vp, i32p]
    lib.qck_shard_info.argtypes = [vp, C.c_int32, i32p, i64p, i64p, i32p, i32p]
    lib.qck_shard_device_buffers.argtypes = [vp, C.c_int32] + [C.POINTER(vp)] * 5
    lib.qck_upload.argtypes = [vp, vp, vp]
    lib.qck_eval_resident.argtypes = [vp, C.c_uint32]
    lib.qck_gather_device.argtypes = [vp, C.c_uint32]
    lib.qck_gathered_buffers.argtypes = [vp, C.c_int32] + [C.POINTER(vp)] * 3
    lib.qck_nccl_version.argtypes = [vp, i32p, i32p]
    lib.qck_transfer_stats.argtypes = [vp, i64p, i64p, i64p]
    lib.qck_invalidate.argtypes = [vp]
    lib.qck_objective_attach.argtypes = [vp, C.POINTER(ObjectiveTerm), C.c_int32]
    lib.qck_objective_sizes.argtypes = [vp, i64p, i64p]
    lib.qck_objective_hessian_structure.argtypes = [vp, vp, vp]
    lib.qck_eval_objective.argtypes = [vp, vp, dp]
    lib.qck_eval_objective_gradient.argtypes = [vp, vp, vp]
    lib.qck_eval_objective_hessian.argtypes = [vp, vp, C.c_double, vp]
    lib.qck_fidelity_constraint_attach.argtypes = [vp, C.POINTER(ObjectiveTerm), C.c_double]
    lib.qck_eval_fidelity_constraint.argtypes = [vp, vp, C.c_double, dp, vp, vp]
    lib.qck_rollout.argtypes = [C.c_int32] * 5 + [vp, vp, C.c_int64, vp, vp, vp, vp]
    lib.qck_rollout_last_error.restype = C.c_char_p
    lib.qck_compact_map.argtypes = [vp, C.c_int32, i64p, vp]
    lib.qck_expand_host.argtypes = [vp, C.c_int32, vp, vp, C.c_int64]
    _lib = lib
    return lib
