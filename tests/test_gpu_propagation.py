"""GPU parity tests of B200SV_PROPAGATION_REFERENCE (run on the B200 box: pytest -m gpu): the field reproduces
PropagationDistanceField's bucket-queue propagation (distance_field/src/propagation_distance_field.cpp:177-444) voxel for
voxel -- compared with the oracle's C restatement after the same sequence of add / remove / update calls, the sizes of
distance_field/test/test_distance_field.cpp and BASELINE.json's C1, C2 and C4."""
import numpy as np
import pytest

import moveit2_b200 as mb
from moveit2_b200 import workloads
from oracle.df import PropagationDistanceField

from util import config_pair

pytestmark = pytest.mark.gpu


def field_pair(size, origin, resolution, max_distance):
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=size, origin=origin, resolution=resolution,
                                           max_distance=max_distance)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
    corner = [o - s / 2 for o, s in zip(origin, size)]
    ref = PropagationDistanceField(size[0], size[1], size[2], resolution, corner[0], corner[1], corner[2], max_distance)
    return eng, ref


def same(eng, ref, what):
    g, o = eng.field_get_d2(), ref.d2()
    bad = int((g != o).sum())
    assert bad == 0, "%s: %d of %d voxels differ from the reference's propagation (max |delta| %d)" % (
        what, bad, g.size, int(np.abs(g.astype(np.int64) - o).max()))


def test_single_add_equals_the_reference_on_c1():
    eng, env, cloud, q = config_pair("C1")
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
    eng.field_add_points(cloud)
    env.world.add_points(cloud)
    same(eng, env.world, "C1 cloud")
    exact_eng, _, _, _ = config_pair("C1")
    exact_eng.field_add_points(cloud)
    n_exact_differs = int((exact_eng.field_get_d2() != env.world.d2()).sum())
    st = eng.field_propagation_stats()
    print("C1: reference mode equal on all voxels (the exact transform differs in %d); %s" % (n_exact_differs, st))
    assert st["rounds"] > 0 and st["queue_entries"] > 0
    exact_eng.close()
    eng.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_incremental_adds_with_duplicates_and_outside_points(seed):
    """Several addPointsToField calls: each propagates from the state the one before left (closest points, update
    directions, entries left in the queue), points outside the grid are skipped, duplicates are queued twice."""
    rng = np.random.default_rng(seed)
    size, origin = (0.8, 0.6, 0.5), (0.1, -0.2, 0.3)
    eng, ref = field_pair(size, origin, 0.02, 0.2)
    for call in range(5):
        n = int(rng.integers(1, 400))
        if call == 3:
            ctr = rng.uniform(-0.3, 0.3, (8, 3)) * size + origin
            pts = (ctr[:, None, :] + rng.normal(0, 0.03, (8, 60, 3))).reshape(-1, 3)
        else:
            pts = (rng.uniform(-0.6, 0.6, (n, 3)) * size + origin)
        pts = np.concatenate([pts, pts[: len(pts) // 3]])  # duplicates
        eng.field_add_points(pts)
        ref.add_points(pts)
        same(eng, ref, "seed %d call %d" % (seed, call))
    eng.close()


@pytest.mark.parametrize("seed", [0, 1])
def test_remove_and_update_points_equal_the_reference(seed):
    """removePointsFromField (:198-219, :303-388: the depth-first flood from the removed voxels, then propagation from the
    surviving boundary) and updatePointsInField (:130-175: ordered set differences, remove, add)."""
    rng = np.random.default_rng(10 + seed)
    size, origin = (0.6, 0.6, 0.6), (0.0, 0.0, 0.3)
    eng, ref = field_pair(size, origin, 0.02, 0.16)
    box = lambda c, h, n: rng.uniform(-1, 1, (n, 3)) * h + c
    a = box((0.1, 0.0, 0.3), 0.05, 600)
    b = box((-0.15, 0.1, 0.4), 0.04, 400)
    c = rng.uniform(-0.5, 0.5, (150, 3)) * size + origin
    for name, pts in (("a", a), ("b", b), ("c", c)):
        eng.field_add_points(pts)
        ref.add_points(pts)
        same(eng, ref, "add " + name)
    eng.field_remove_points(a)
    ref.remove_points(a)
    same(eng, ref, "remove a")
    assert int((eng.field_get_d2() == 0).sum()) == int((ref.d2() == 0).sum())
    # remove points that were never added, duplicates, points outside
    junk = np.concatenate([rng.uniform(-0.7, 0.7, (50, 3)) * size + origin, c[:20], c[:20]])
    eng.field_remove_points(junk)
    ref.remove_points(junk)
    same(eng, ref, "remove junk")
    # the object moves: updatePointsInField(old, new)
    b2 = b + np.array([0.06, -0.04, 0.02])
    eng.field_update_points(b, b2)
    ref.update_points(b, b2)
    same(eng, ref, "update b -> b2")
    eng.field_update_points(b2, b2[:0])
    ref.update_points(b2, b2[:0])
    same(eng, ref, "update b2 -> nothing")
    eng.field_add_points(a)
    ref.add_points(a)
    same(eng, ref, "add a again")
    # reset() and a fresh add
    eng.field_reset()
    ref.reset()
    same(eng, ref, "reset")
    eng.field_add_points(c)
    ref.add_points(c)
    same(eng, ref, "add c after reset")
    print(eng.field_propagation_stats())
    eng.close()


def test_validity_with_each_sides_own_fields_is_identical_in_reference_mode():
    """With the exact transform a handful of states per million differ from the reference when each side checks against
    its OWN field (the EDT is never above the propagation).  In reference mode both fields are the reference's: no state
    differs, world and self."""
    eng, env, cloud, q = config_pair("C1", n_states=200_000)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_WORLD)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_SELF)
    eng.field_add_points(cloud, mb.FIELD_WORLD)   # (the self field was refilled from the context's link points)
    env.world.add_points(cloud)
    assert np.array_equal(eng.field_get_d2(mb.FIELD_WORLD), env.world.d2())
    assert np.array_equal(eng.field_get_d2(mb.FIELD_SELF), env.self_field.d2())
    ref, _ = env.check(q, mb.CHECK_ALL)
    got = eng.check_batch(q, mb.CHECK_ALL)
    assert np.array_equal(got, ref == 0)
    eng.close()


def test_entry_points_that_edit_bits_refuse_in_reference_mode():
    eng, ref = field_pair((0.4, 0.4, 0.4), (0.0, 0.0, 0.2), 0.02, 0.1)
    with pytest.raises(mb.B200svError):
        eng.field_add_sphere((0.0, 0.0, 0.2), 0.05)
    with pytest.raises(mb.B200svError):
        eng.field_set_d2(np.zeros(eng.num_cells, np.int32))
    with pytest.raises(mb.B200svError):
        eng.field_begin_update()
    # back to the exact transform: everything works again, the field starts empty
    eng.field_set_propagation(mb.PROPAGATION_EXACT)
    assert int((eng.field_get_d2() == 0).sum()) == 0
    eng.field_add_sphere((0.0, 0.0, 0.2), 0.05)
    assert int((eng.field_get_d2() == 0).sum()) > 0
    eng.close()


def test_c2_field_equals_the_reference():
    """BASELINE.json configs[1]'s field: 500 000 points -> 200^3 voxels @ 1 cm, 25 cells."""
    c2 = workloads.CONFIGS["C2"]
    eng, ref = field_pair(c2["size"], c2["origin"], c2["resolution"], c2["max_distance"])
    pts = workloads.make_cloud("C2")
    eng.field_add_points(pts)
    ms = eng.stats()["edt_ms"]
    ref.add_points(pts)
    same(eng, ref, "C2")
    print("C2 field in reference mode: %.1f ms; %s" % (ms, eng.field_propagation_stats()))
    eng.close()


def test_c4_full_size_equals_the_reference():
    """BASELINE.json configs[3] at full size: 5 M points -> 400 x 400 x 200 voxels, all 32 M voxels equal the reference's."""
    c4 = workloads.CONFIGS["C4"]
    eng, ref = field_pair(c4["size"], c4["origin"], c4["resolution"], c4["max_distance"])
    pts = workloads.make_cloud("C4")
    eng.field_add_points(pts)
    ms = eng.stats()["edt_ms"]
    ref.add_points(pts)
    same(eng, ref, "C4")
    print("C4 field in reference mode: %.1f ms; %s" % (ms, eng.field_propagation_stats()))
    eng.close()


def test_base_state_change_in_reference_mode_builds_the_self_field_like_a_fresh_cache_entry():
    """compareCacheEntryToState (:1254-1312) -> a new DistanceFieldCacheEntry with a FRESH self field (:931-953): with both
    fields in reference mode the context after b200sv_set_base_state equals a fresh oracle env at the new base state voxel
    for voxel, and so do the verdicts with nothing injected."""
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    from util import RES, read
    urdf, srdf = read(RES, "dualarm_spheres.urdf"), read(RES, "dualarm.srdf")
    geo = ((3.0, 3.0, 2.5), (0.0, 0.0, 1.0), 0.05, 0.25)
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "left_arm", *geo)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_WORLD)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_SELF)
    env0 = OracleEnv(RobotModel(urdf, srdf), "left_arm", *geo)
    assert np.array_equal(eng.field_get_d2(mb.FIELD_SELF), env0.self_field.d2())
    cloud = workloads.clutter_cloud(6, 20000, 12, (-1.3, -1.3, 0.1), (1.3, 1.3, 2.0), 0.75)
    eng.field_add_points(cloud)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(22, 20000, lo, hi)
    base = np.array(env0.base_positions)
    rng = np.random.default_rng(5)
    gv = set(int(v) for v in env0.arr["group_var_index"])
    for v in RobotModel(urdf, srdf).groups["right_arm"].variable_index_list:
        if v not in gv:
            base[v] = rng.uniform(-1.0, 1.0)
    env1 = OracleEnv(RobotModel(urdf, srdf), "left_arm", *geo, base_positions=base)
    env1.world.add_points(cloud)
    eng.set_base_state(base)
    assert np.array_equal(eng.field_get_d2(mb.FIELD_SELF), env1.self_field.d2())
    assert np.array_equal(eng.field_get_d2(mb.FIELD_WORLD), env1.world.d2())
    ref1, _ = env1.check(q, mb.CHECK_ALL)
    assert np.array_equal(eng.check_batch(q, mb.CHECK_ALL), ref1 == 0)
    eng.close()


def test_several_devices_receive_the_reference_field():
    """b200sv_multi_field_set_propagation: device 0 runs the propagation, every device of the handle checks against the
    reference's voxels."""
    torch = pytest.importorskip("torch")
    from util import RES, read
    devices = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    eng, env, cloud, q = config_pair("C1", n_states=20_000)
    eng.close()
    cfg = workloads.CONFIGS["C1"]
    me = mb.MultiDeviceEngine.from_urdf(devices, read(RES, "panda_spheres.urdf"), read(RES, "panda.srdf"), cfg["group"],
                                        cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    me.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_WORLD)
    me.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_SELF)
    me.field_add_points(cloud)
    env.world.add_points(cloud)
    for c in me.contexts:
        assert np.array_equal(c.field_get_d2(mb.FIELD_WORLD), env.world.d2())
        assert np.array_equal(c.field_get_d2(mb.FIELD_SELF), env.self_field.d2())
    ref, _ = env.check(q, mb.CHECK_ALL)
    assert np.array_equal(me.check_batch(q), ref == 0)
    me.close()


@pytest.mark.parametrize("size,res,maxd", [((0.06, 0.04, 0.02), 0.02, 0.1),      # 3 x 2 x 1 cells, R = 5 > every extent
                                           ((0.5, 0.5, 0.14), 0.02, 0.02),       # 25 x 25 x 7 cells (nz < 8), R = 1
                                           ((0.3, 0.2, 0.4), 0.01, 0.3)])        # 30 x 20 x 40 cells, R = 30: 901 buckets
def test_edge_cases_empty_calls_tiny_grids_radius_one_and_radius_beyond_the_grid(size, res, maxd):
    origin = (0.05, -0.02, 0.1)
    eng, ref = field_pair(size, origin, res, maxd)
    rng = np.random.default_rng(3)
    none = np.zeros((0, 3))
    outside = rng.uniform(2.0, 3.0, (20, 3))
    corner = np.array(origin) - np.array(size) / 2
    on_faces = np.array([corner, corner + size, corner + np.array(size) * 0.5, corner + [size[0], 0, 0],
                         corner + 1e-12, corner + np.array(size) - 1e-12])
    inside = rng.uniform(0, 1, (40, 3)) * size + corner
    steps = [("add nothing", "add", none), ("remove nothing", "remove", none), ("add outside", "add", outside),
             ("add faces", "add", on_faces), ("add inside", "add", inside), ("remove outside", "remove", outside),
             ("remove half", "remove", inside[::2]), ("remove all", "remove", np.concatenate([inside, on_faces])),
             ("add again", "add", inside[:5])]
    for what, op, pts in steps:
        getattr(eng, "field_%s_points" % op)(pts)
        getattr(ref, "%s_points" % op)(pts)
        same(eng, ref, what)
    eng.field_update_points(none, inside)
    ref.update_points(none, inside)
    same(eng, ref, "update nothing -> inside")
    eng.field_update_points(inside, none)
    ref.update_points(inside, none)
    same(eng, ref, "update inside -> nothing")
    eng.close()


def test_the_references_df_fixture_through_read_df_in_reference_mode():
    """distance_field/test/resources/test_big.df (65 126 voxels of the r = 0.5 sphere the reference's TestReadWrite writes,
    committed as occupancy bits in tests/golden/df_golden.npz): readFromStream = reset + addNewObstacleVoxels of the stream's
    voxels in x, y, z order.  The device field equals the oracle's fed the same voxels in the same order."""
    import os
    from oracle.df import write_df_occupancy
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "df_golden.npz"))
    hdr = {k: float(gold["test_big_header"][i]) for i, k in enumerate(
        ("resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z"))}
    n = [int(v) for v in gold["test_big_dims"]]
    occ = np.unpackbits(gold["test_big_occ_bits"])[: n[0] * n[1] * n[2]].reshape(n).astype(bool)
    assert int(occ.sum()) == 65126
    size = (hdr["size_x"], hdr["size_y"], hdr["size_z"])
    origin = tuple(hdr["origin_%s" % a] + hdr["size_%s" % a] / 2 for a in "xyz")
    eng, ref = field_pair(size, origin, hdr["resolution"], 0.25)
    assert eng.num_cells == tuple(n)
    eng.field_read_df(write_df_occupancy(hdr, occ))
    ref.add_voxels(np.argwhere(occ).astype(np.int32))
    same(eng, ref, "test_big.df")
    assert int((eng.field_get_d2() == 0).sum()) == int(occ.sum())
    # a second stream on top: the field is reset first, as initialize() does
    occ2 = np.zeros_like(occ)
    occ2[3, 4, 5] = occ2[10, 20, 7] = True
    eng.field_read_df(write_df_occupancy(hdr, occ2))
    ref.reset()
    ref.add_voxels(np.argwhere(occ2).astype(np.int32))
    same(eng, ref, "second stream")
    eng.close()


@pytest.mark.parametrize("seed", [0, 1])
def test_signed_fields_in_reference_mode_equal_the_reference_on_both_halves(seed):
    """propagate_negative_ (:48-58): adding obstacle voxels floods the negative half from them (:255-298) and runs
    propagateNegative (:446-504); removing them queues them in the negative bucket 0 (:321-328).  Both halves equal the
    oracle's after adds of solid bodies, a removal, an update and a reset."""
    rng = np.random.default_rng(40 + seed)
    size, origin, res, maxd = (0.6, 0.5, 0.4), (0.0, 0.1, 0.2), 0.02, 0.14
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=size, origin=origin, resolution=res,
                                           max_distance=maxd, signed=True)
    eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
    corner = [o - s_ / 2 for o, s_ in zip(origin, size)]
    ref = PropagationDistanceField(size[0], size[1], size[2], res, corner[0], corner[1], corner[2], maxd, True)

    def both(what):
        same(eng, ref, what)
        g, o = eng.field_get_neg_d2(), ref.neg_d2()
        bad = int((g != o).sum())
        assert bad == 0, "%s: %d voxels of the NEGATIVE half differ" % (what, bad)

    def solid(c, h):  # every voxel centre of a box: interior voxels get negative distances
        ax = [np.arange(c[k] - h[k], c[k] + h[k] + 1e-9, res) for k in range(3)]
        p = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
        return p[rng.permutation(len(p))]

    a = solid((0.05, 0.1, 0.2), (0.08, 0.06, 0.1))
    b = solid((-0.15, 0.2, 0.25), (0.05, 0.05, 0.05))
    eng.field_add_points(a); ref.add_points(a); both("add a")
    assert int((ref.neg_d2() > 0).sum()) > 50
    eng.field_add_points(b); ref.add_points(b); both("add b")
    part = a[: len(a) // 3]
    eng.field_remove_points(part); ref.remove_points(part); both("remove a third of a")
    b2 = b + np.array([0.04, -0.02, 0.0])
    eng.field_update_points(b, b2); ref.update_points(b, b2); both("update b -> b2")
    eng.field_reset(); ref.reset(); both("reset")
    eng.field_add_points(b2); ref.add_points(b2); both("add after reset")
    eng.close()


@pytest.mark.parametrize("resident", ["1", "0"])
def test_randomised_campaign_with_and_without_the_resident_kernel(resident):
    """tools/prop_fuzz.py for a few seconds in a process of its own: random grids, radii, signed / unsigned fields, random
    add / remove / update / reset sequences, every call compared with the oracle.  B200SV_PROP_RESIDENT=0 keeps every round
    host-driven, so both ways of draining a bucket are held to the same campaign."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200SV_PROP_RESIDENT=resident)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "prop_fuzz.py"), "8", "3"], capture_output=True, text=True,
                         env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    summary = json.loads(out.stdout.strip().splitlines()[-1])
    assert summary["calls"] > 50 and summary["mismatching_calls"] == 0, (summary, out.stderr[-2000:])
