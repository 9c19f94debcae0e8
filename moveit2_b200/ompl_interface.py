"""Host-side mirror of the planner-facing interface around the batched path.

Names follow the reference / OMPL so planner code reads the same:
  ModelBasedStateSpace   moveit_planners/ompl/ompl_interface/src/parameterization/model_based_state_space.cpp:158-222
  StateValidityChecker   moveit_planners/ompl/ompl_interface/src/detail/state_validity_checker.cpp:83-133
  MotionValidator        ompl::base::DiscreteMotionValidator as configured by model_based_planning_context.cpp:294-317
plus the batched entry points that are the reason this package exists.  Nothing here computes validity: every
decision is one call into libb200sv.so (b200sv_check_batch / b200sv_check_motions).  The small numpy helpers
(distance, interpolate) are the planner's own bookkeeping (nearest neighbours, tree extension), not the hot path.
"""
import numpy as np

from .engine import CHECK_ALL


class ModelBasedStateSpace:
    """Joint space of a planning group: bounds, JointModelGroup::distance / interpolate / getMaximumExtent."""

    def __init__(self, lower, upper, continuous=None, distance_factor=None, longest_valid_segment_fraction=0.01):
        self.lower = np.asarray(lower, np.float64)
        self.upper = np.asarray(upper, np.float64)
        n = len(self.lower)
        self.continuous = np.zeros(n, np.uint8) if continuous is None else np.asarray(continuous, np.uint8)
        self.factor = np.ones(n) if distance_factor is None else np.asarray(distance_factor, np.float64)
        self.longest_valid_segment_fraction = float(longest_valid_segment_fraction)  # OMPL's default is 0.01

    def getDimension(self):
        return len(self.lower)

    def getMaximumExtent(self):  # joint_model_group.cpp:451-460
        return float(np.sum((self.upper - self.lower) * self.factor))

    def getLongestValidSegmentLength(self):
        return self.longest_valid_segment_fraction * self.getMaximumExtent()

    def satisfiesBounds(self, q):
        q = np.atleast_2d(q)
        return np.all((q >= self.lower - 1e-12) & (q <= self.upper + 1e-12), axis=1)

    def distance(self, a, b):  # joint_model_group.cpp:462-472, vectorised over rows
        d = np.abs(np.asarray(a) - np.asarray(b))
        c = self.continuous.astype(bool)
        if c.any():
            w = np.fmod(d[..., c], 2.0 * np.pi)
            d[..., c] = np.where(w > np.pi, 2.0 * np.pi - w, w)
        return np.sum(d * self.factor, axis=-1)

    def interpolate(self, a, b, t):  # joint_model_group.cpp:474-486 (non-continuous variables: from + (to - from) * t)
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        out = a + (b - a) * t
        c = self.continuous.astype(bool)
        if c.any():
            diff = (b - a)[..., c]
            wrap = np.abs(diff) > np.pi
            alt = np.where(diff > 0.0, 2.0 * np.pi - diff, -2.0 * np.pi - diff)
            v = a[..., c] - alt * t
            v = np.where(v > np.pi, v - 2.0 * np.pi, np.where(v < -np.pi, v + 2.0 * np.pi, v))
            out[..., c] = np.where(wrap, v, out[..., c])
        return out

    def sampleUniform(self, rng, n=1):
        return rng.uniform(self.lower, self.upper, size=(n, len(self.lower)))


class StateValidityChecker:
    """isValid(state) as the reference's checker decides it (bounds, then collision); batches go to the GPU."""

    def __init__(self, engine, space, flags=CHECK_ALL):
        self.engine, self.space, self.flags = engine, space, flags
        self.n_states = 0

    def isValid(self, q):
        return bool(self.isValidBatch(np.asarray(q, np.float64).reshape(1, -1))[0])

    def isValidBatch(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.space.getDimension())
        ok = self.space.satisfiesBounds(q)
        self.n_states += len(q)
        return ok & self.engine.check_batch(q, self.flags)


class MotionValidator:
    """checkMotion(s1, s2) with DiscreteMotionValidator's discretisation; checkMotions validates M edges in ONE call."""

    def __init__(self, engine, space, flags=CHECK_ALL, max_states_per_edge=0):
        self.engine, self.space, self.flags, self.cap = engine, space, flags, max_states_per_edge
        self.n_edges = self.n_states = self.n_calls = 0

    def checkMotion(self, s1, s2):
        return bool(self.checkMotions(np.asarray(s1).reshape(1, -1), np.asarray(s2).reshape(1, -1))[0][0])

    def checkMotions(self, s1, s2):
        """(valid bool [M], last_valid_fraction float [M]): the fraction is OMPL's lastValid.second, (j - 1) / nd."""
        valid, first, n = self.engine.check_motions(s1, s2, self.space.getLongestValidSegmentLength(), self.space.continuous,
                                                   self.space.factor, self.cap, self.flags)
        self.n_edges += len(valid)
        self.n_states += n
        self.n_calls += 1
        nd = np.maximum(np.ceil(self.space.distance(np.asarray(s1, np.float64).reshape(len(valid), -1),
                                                    np.asarray(s2, np.float64).reshape(len(valid), -1)) /
                                self.space.getLongestValidSegmentLength()), 1.0)
        if self.cap:
            nd = np.minimum(nd, self.cap)
        frac = np.where(valid, 1.0, (np.maximum(first, 1) - 1) / nd)
        return valid, frac
