"""Randomised campaign for B200SV_PROPAGATION_REFERENCE: random grids, truncation radii, signed / unsigned fields and random
sequences of add / remove / update / reset calls (clustered points, duplicates, points outside); after EVERY call both halves
of the device field are compared with the oracle's.  Prints one JSON summary.
    python tools/prop_fuzz.py [seconds] [seed]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import moveit2_b200 as mb
from oracle.df import PropagationDistanceField

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time()
summary = dict(fields=0, calls=0, voxels_compared=0, mismatching_calls=0, hazard_rounds=0, rounds=0, backward_offers=0,
               queued_for_next_call_seen=0, signed_fields=0, ops={})
case = 0
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed0 * 100003 + case)
    case += 1
    res = float(rng.choice([0.01, 0.02, 0.05]))
    n = rng.integers(3, 70, 3)
    size = tuple(float(v) * res for v in n)
    origin = tuple(rng.uniform(-0.5, 0.5, 3))
    maxd = float(rng.choice([1, 2, 5, 9, 14]) * res * rng.uniform(0.6, 1.0))
    signed = bool(rng.random() < 0.35)
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=size, origin=origin, resolution=res, max_distance=maxd,
                                           signed=signed)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
    corner = [o - s / 2 for o, s in zip(origin, size)]
    ref = PropagationDistanceField(size[0], size[1], size[2], res, corner[0], corner[1], corner[2], maxd, signed)
    summary["fields"] += 1
    summary["signed_fields"] += int(signed)
    live = []  # point sets currently in the field

    def cloud():
        kind = rng.integers(0, 4)
        m = int(rng.integers(1, 400))
        if kind == 0:
            p = rng.uniform(-0.6, 0.6, (m, 3)) * size + origin                      # uniform, some outside
        elif kind == 1:
            c = rng.uniform(-0.4, 0.4, (max(1, m // 40), 3)) * size + origin       # clusters
            p = (c[:, None, :] + rng.normal(0, 1.5 * res, (len(c), 40, 3))).reshape(-1, 3)
        elif kind == 2:
            c = rng.uniform(-0.3, 0.3, 3) * size + origin                          # a solid box
            h = rng.uniform(1, 5, 3) * res
            ax = [np.arange(c[k] - h[k], c[k] + h[k] + 1e-9, res) for k in range(3)]
            p = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
            p = p[rng.permutation(len(p))]
        else:
            p = np.repeat(rng.uniform(-0.5, 0.5, (max(1, m // 10), 3)) * size + origin, 10, axis=0)  # duplicates
        return np.ascontiguousarray(p)

    for step in range(int(rng.integers(3, 10))):
        r = rng.random()
        if r < 0.45 or not live:
            op, p = "add", cloud()
            eng.field_add_points(p); ref.add_points(p); live.append(p)
        elif r < 0.65:
            op = "remove"
            p = live.pop(int(rng.integers(0, len(live))))
            if rng.random() < 0.5:
                p = p[: max(1, len(p) // 2)]
            eng.field_remove_points(p); ref.remove_points(p)
        elif r < 0.9:
            op = "update"
            k = int(rng.integers(0, len(live)))
            old, new = live[k], live[k] + rng.integers(-3, 4, 3) * res
            eng.field_update_points(old, new); ref.update_points(old, new); live[k] = new
        else:
            op = "reset"
            eng.field_reset(); ref.reset(); live = []
        g, o = eng.field_get_d2(), ref.d2()
        bad = int((g != o).sum())
        if signed:
            bad += int((eng.field_get_neg_d2() != ref.neg_d2()).sum())
        summary["calls"] += 1
        summary["voxels_compared"] += int(g.size) * (2 if signed else 1)
        summary["ops"][op] = summary["ops"].get(op, 0) + 1
        if bad:
            summary["mismatching_calls"] += 1
            print("MISMATCH case %d step %d op %s: %d voxels (grid %s, res %g, maxd %g, signed %s)" %
                  (case - 1, step, op, bad, tuple(n), res, maxd, signed), file=sys.stderr)
            break
    st = eng.field_propagation_stats()
    summary["hazard_rounds"] += st["hazard_rounds"]
    summary["rounds"] += st["rounds"]
    summary["backward_offers"] += st["backward_offers"]
    summary["queued_for_next_call_seen"] += int(st["queued_for_next_call"] > 0)
    eng.close()
summary["seconds"] = round(time.time() - t0, 1)
print(json.dumps(summary))
