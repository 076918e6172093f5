"""qcknot: B200-native knot-point dynamics evaluator behind QuantumCollocation.jl's integrator/dynamics API.

The directory is named quantumcollocation.jl_b200 (not importable as-is because of the dot); `import qcknot`
(the shim at the repo root) loads it under the module name `qcknot`.
"""
from ._lib import QCK_EVAL_F, QCK_EVAL_H, QCK_EVAL_J, LIB_PATH  # noqa: F401
from .build import build_library  # noqa: F401
from .dynamics import QcknotError, QuantumDynamics, dense, host_register, host_unregister  # noqa: F401
from .integrators import (  # noqa: F401
    DensityOperatorExponentialIntegrator,
    DerivativeIntegrator,
    QuantumStateExponentialIntegrator,
    QuantumStatePadeIntegrator,
    UnitaryExponentialIntegrator,
    UnitaryPadeIntegrator,
)
from .objectives import (  # noqa: F401
    FinalUnitaryFidelityConstraint,
    MinimumTimeObjective,
    Objective,
    QuadraticRegularizer,
    UnitaryInfidelityObjective,
)
from .isomorphisms import iso_to_ket, iso_vec_to_operator, ket_to_iso, operator_to_iso_vec  # noqa: F401
from .quantum_system import OpenQuantumSystem, QuantumSystem  # noqa: F401
from .rollouts import iso_vec_unitary_fidelity, rollout, unitary_rollout, unitary_rollout_fidelity  # noqa: F401
from .trajectory import NamedTrajectory  # noqa: F401
from . import sharding, workloads  # noqa: F401
