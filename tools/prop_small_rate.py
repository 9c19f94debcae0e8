"""Reference-propagation mode on SMALL work: C1 (50^3 voxels, 13-cell truncation) build and single-object moves; C2 with a
small object moving.  Run with B200SV_PROP_RESIDENT=0 / 1 to compare the host-driven rounds with the resident kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads

for name, n_move in (("C1", 100), ("C2", 2000)):
    c = workloads.CONFIGS[name]
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c["size"], origin=c["origin"], resolution=c["resolution"],
                                           max_distance=c["max_distance"])
    pts = workloads.make_cloud(name)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
    ts = []
    for r in range(3):
        eng.field_reset()
        t = time.perf_counter(); eng.field_add_points(pts); ts.append(time.perf_counter() - t)
    obj = pts[:n_move].copy()
    tu = []
    for k in range(10):
        new = obj + 0.01
        t = time.perf_counter(); eng.field_update_points(obj, new); tu.append(time.perf_counter() - t)
        obj = new
    st = eng.field_propagation_stats()
    print("%s (resident=%s): build %.2f ms; move of %d points %.2f ms median (min %.2f); rounds %d, hazard rounds %d" %
          (name, os.environ.get("B200SV_PROP_RESIDENT", "1"), 1e3 * min(ts[1:]), n_move, 1e3 * float(np.median(tu)), 1e3 * min(tu),
           st["rounds"], st["hazard_rounds"]), flush=True)
    eng.close()
