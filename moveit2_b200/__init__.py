"""moveit2_b200 -- B200-native batched state-validity engine for MoveIt 2's distance-field collision path.

Only what the hot path needs: csrc/ (CUDA kernels + the C-ABI of include/b200sv.h), engine.py (ctypes host
side mirroring the reference's CollisionEnv entry points), workloads.py (synthetic benchmark inputs),
adapter/ (the MoveIt-side C++ plugin shim), resources/ (benchmark robot descriptions).
"""
from .engine import (B200svError, CollisionEnvB200, CollisionRequest, CollisionResult,  # noqa: F401
                     MultiDeviceEngine, StateValidityEngine)
from ._lib import (CHECK_ALL, CHECK_FP64, CHECK_INTRA, CHECK_SELF, CHECK_WORLD, DEVICE_NONE, FIELD_SELF,  # noqa: F401
                   FIELD_WORLD, LATTICE_BODY, LATTICE_WORLD, PROPAGATION_EXACT, PROPAGATION_REFERENCE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_SPHERE)
