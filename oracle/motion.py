"""TEST INFRASTRUCTURE ONLY -- CPU restatement of edge ("motion") validation as MoveIt configures OMPL for it.

OMPL is not vendored in /root/reference (moveit_planners/ompl/package.xml:25, rosdep, no version pin), so the
discretisation below restates the published algorithm of ompl::base::DiscreteMotionValidator::checkMotion and
StateSpace::validSegmentCount; the MoveIt-side pieces follow the reference:
  distance     JointModelGroup::distance        moveit_core/robot_model/src/joint_model_group.cpp:462-472
               RevoluteJointModel::distance     moveit_core/robot_model/src/revolute_joint_model.cpp:173-182
  interpolate  JointModelGroup::interpolate     joint_model_group.cpp:474-486
               RevoluteJointModel::interpolate  revolute_joint_model.cpp:138-171
  extent       JointModelGroup::getMaximumExtent joint_model_group.cpp:451-460
Parity for the discretisation itself is unpinned (no golden data in the reference tree).
"""
import math

import numpy as np


def distance(a, b, continuous=None, factor=None):
    d = 0.0
    for k in range(len(a)):
        w = 1.0 if factor is None else float(factor[k])
        if w == 0.0:
            continue
        dk = abs(float(a[k]) - float(b[k]))
        if continuous is not None and continuous[k]:
            dk = math.fmod(dk, 2.0 * math.pi)
            dk = 2.0 * math.pi - dk if dk > math.pi else dk
        d = d + w * dk
    return d


def interpolate(a, b, t, continuous=None):
    out = np.empty(len(a), np.float64)
    for k in range(len(a)):
        f, to = float(a[k]), float(b[k])
        diff = to - f
        if continuous is not None and continuous[k] and abs(diff) > math.pi:
            diff = 2.0 * math.pi - diff if diff > 0.0 else -2.0 * math.pi - diff
            v = f - diff * t
            if v > math.pi:
                v -= 2.0 * math.pi
            elif v < -math.pi:
                v += 2.0 * math.pi
        else:
            v = f + diff * t
        out[k] = v
    return out


def edge_states(a, b, max_segment, continuous=None, factor=None, cap=0):
    """States DiscreteMotionValidator::checkMotion visits for (a, b): j / nd for j = 1 .. nd - 1, then b."""
    nd = int(math.ceil(distance(a, b, continuous, factor) / max_segment))
    nd = max(nd, 1)
    if cap:
        nd = min(nd, cap)
    qs = [interpolate(a, b, j / nd, continuous) for j in range(1, nd)]
    qs.append(np.asarray(b, np.float64).copy())
    return np.stack(qs)


def check_motions(env, q_from, q_to, max_segment, continuous=None, factor=None, cap=0, flags=7):
    """(valid bool [M], first_invalid int32 [M]) through OracleEnv.check on every visited state."""
    states, owner = [], []
    for e in range(len(q_from)):
        s = edge_states(q_from[e], q_to[e], max_segment, continuous, factor, cap)
        states.append(s)
        owner.append(np.full(len(s), e))
    allq = np.concatenate(states)
    col, _ = env.check(allq, flags)
    col = np.asarray(col).astype(bool)
    valid = np.ones(len(q_from), bool)
    first = np.full(len(q_from), -1, np.int32)
    pos = 0
    for e, s in enumerate(states):
        c = col[pos:pos + len(s)]
        pos += len(s)
        if c.any():
            valid[e] = False
            first[e] = int(np.argmax(c)) + 1
    return valid, first, len(allq)
