"""A self-contained RRTConnect harness for config C5 (BASELINE.json): Panda in a cluttered scene, planning time with
batched edge validation.  OMPL is not installed in this image (SURVEY.md section 8d allows a self-contained harness
that uses the same validity API); the planner logic below is the textbook bidirectional RRT (Kuffner & LaValle) with ONE
change that makes it batch-shaped: every iteration grows the active tree towards `batch` random targets at once, so
the planner asks the validator about `batch` edges per call instead of one.  All validity decisions come from the
validator object it is given (moveit2_b200.ompl_interface.MotionValidator on the GPU, or any object with the same
checkMotions signature -- the benchmark also runs the oracle's CPU restatement for the reference arm).
"""
import time

import numpy as np


class _Tree:
    def __init__(self, root, dim):
        self.q = np.empty((1024, dim))
        self.parent = np.empty(1024, np.int64)
        self.n = 0
        self.add(root, -1)

    def add(self, q, parent):
        if self.n == len(self.q):
            self.q = np.concatenate([self.q, np.empty_like(self.q)])
            self.parent = np.concatenate([self.parent, np.empty_like(self.parent)])
        self.q[self.n] = q
        self.parent[self.n] = parent
        self.n += 1
        return self.n - 1

    def nearest(self, space, targets):
        d = space.distance(self.q[None, :self.n, :], targets[:, None, :])  # [targets, nodes]
        return np.argmin(d, axis=1)

    def path_to_root(self, i):
        out = []
        while i >= 0:
            out.append(self.q[i].copy())
            i = self.parent[i]
        return out


def rrt_connect(space, validator, start, goal, rng, batch=32, range_fraction=0.2, max_iterations=400, time_limit=10.0):
    """Returns (path or None, stats).  range: OMPL's RRTConnect default, 0.2 * maximum extent."""
    t0 = time.perf_counter()
    max_range = range_fraction * space.getMaximumExtent()
    ta, tb = _Tree(start, len(start)), _Tree(goal, len(start))
    a_is_start = True
    stats = {"iterations": 0, "edges": 0}
    # the direct edge first
    ok, _ = validator.checkMotions(start[None, :], goal[None, :])
    stats["edges"] += 1
    if ok[0]:
        stats["time_s"] = time.perf_counter() - t0
        return [start, goal], stats
    for it in range(max_iterations):
        if time.perf_counter() - t0 > time_limit:
            break
        stats["iterations"] = it + 1
        targets = space.sampleUniform(rng, batch)
        near = ta.nearest(space, targets)
        qn = ta.q[near]
        d = space.distance(qn, targets)
        t = np.minimum(1.0, max_range / np.maximum(d, 1e-12))
        new = space.interpolate(qn, targets, t[:, None])
        ok, frac = validator.checkMotions(qn, new)
        stats["edges"] += len(ok)
        added = []
        for k in range(batch):
            if ok[k]:
                added.append(ta.add(new[k], near[k]))
        if added:
            # connect step: the other tree reaches for every new node with one full-length edge each
            qa = ta.q[added]
            nb = tb.nearest(space, qa)
            ok2, frac2 = validator.checkMotions(tb.q[nb], qa)
            stats["edges"] += len(ok2)
            hit = np.nonzero(ok2)[0]
            if len(hit):
                k = hit[0]
                pa, pb = ta.path_to_root(added[k]), tb.path_to_root(nb[k])
                path = pa[::-1] + pb if a_is_start else pb[::-1] + pa
                stats["time_s"] = time.perf_counter() - t0
                return path, stats
            # keep the partial progress of blocked connect edges (the state before the first invalid one)
            for k in range(len(added)):
                if 0.0 < frac2[k] < 1.0:
                    tb.add(space.interpolate(tb.q[nb[k]], qa[k], frac2[k]), nb[k])
        ta, tb = tb, ta
        a_is_start = not a_is_start
    stats["time_s"] = time.perf_counter() - t0
    return None, stats
