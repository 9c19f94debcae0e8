"""Host logic of the planner-facing mirror (no GPU): ModelBasedStateSpace arithmetic against the oracle's restatement of
JointModelGroup::distance / interpolate, and the RRTConnect harness against a toy validator."""
import numpy as np

from moveit2_b200.ompl_interface import ModelBasedStateSpace
from moveit2_b200.planning import rrt_connect
from oracle import motion


def test_state_space_matches_reference_restatement():
    rng = np.random.default_rng(0)
    lo, hi = np.array([-np.pi, -2.0, -np.pi]), np.array([np.pi, 2.0, np.pi])
    cont, fac = np.array([1, 0, 0], np.uint8), np.array([1.0, 1.0, 0.5])
    space = ModelBasedStateSpace(lo, hi, cont, fac)
    assert abs(space.getMaximumExtent() - (2 * np.pi + 4.0 + np.pi)) < 1e-12
    a, b = rng.uniform(lo, hi, size=(200, 3)), rng.uniform(lo, hi, size=(200, 3))
    d = space.distance(a, b)
    for i in range(200):
        assert abs(d[i] - motion.distance(a[i], b[i], cont, fac)) < 1e-12
        for t in (0.0, 0.3, 1.0):
            assert np.allclose(space.interpolate(a[i], b[i], t), motion.interpolate(a[i], b[i], t, cont), atol=1e-12)


class _WallValidator:
    """2-D toy: a wall at x in [-0.1, 0.1] with a gap at y > 0.6; edges are sampled densely."""

    def __init__(self):
        self.n_edges = self.n_states = self.n_calls = 0

    @staticmethod
    def _free(p):
        return ~((np.abs(p[:, 0]) < 0.1) & (p[:, 1] < 0.6))

    def checkMotions(self, s1, s2):
        s1, s2 = np.atleast_2d(s1), np.atleast_2d(s2)
        valid, frac = np.ones(len(s1), bool), np.ones(len(s1))
        for e in range(len(s1)):
            ts = np.linspace(0.0, 1.0, 65)[1:]
            pts = s1[e] + (s2[e] - s1[e]) * ts[:, None]
            bad = np.nonzero(~self._free(pts))[0]
            if len(bad):
                valid[e] = False
                frac[e] = ts[bad[0] - 1] if bad[0] > 0 else 0.0
        self.n_edges += len(s1)
        return valid, frac


def test_rrt_connect_finds_a_path_around_the_wall():
    space = ModelBasedStateSpace([-1.0, -1.0], [1.0, 1.0])
    v = _WallValidator()
    path, stats = rrt_connect(space, v, np.array([-0.8, -0.5]), np.array([0.8, -0.5]), np.random.default_rng(3), batch=16)
    assert path is not None and np.allclose(path[0], [-0.8, -0.5]) and np.allclose(path[-1], [0.8, -0.5])
    ok, _ = v.checkMotions(np.stack(path[:-1]), np.stack(path[1:]))
    assert ok.all()
    assert stats["edges"] > 1
