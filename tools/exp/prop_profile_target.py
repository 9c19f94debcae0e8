"""Small target for ncu: C2 field in reference-propagation mode, one warm build and one object move."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads
c = workloads.CONFIGS["C2"]
eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c["size"], origin=c["origin"], resolution=c["resolution"], max_distance=c["max_distance"])
pts = workloads.make_cloud("C2")
eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
eng.field_add_points(pts)
eng.field_update_points(pts[:2000], pts[:2000] + 0.01)
eng.close()
