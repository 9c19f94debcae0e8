"""GPU parity tests of the entry points added in round 2 (run on the B200 box: pytest -m gpu): the full-size C4 rebuild,
the single-launch latency path, float32 states, cache-entry updates on a live context (ACM masks, base state), field
replication, several devices behind one handle, device-resident FK."""
import os

import numpy as np
import pytest

import moveit2_b200 as mb
from moveit2_b200 import workloads
from oracle.collision_model import OracleEnv
from oracle.df import PropagationDistanceField
from oracle.robot_model import RobotModel

from util import RES, config_pair, make_pair, read

pytestmark = pytest.mark.gpu

ALL = mb.CHECK_ALL


@pytest.fixture(scope="module")
def c1():
    eng, env, cloud, q = config_pair("C1")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    yield eng, env, cloud, q
    eng.close()


def test_edt_c4_full_size_is_the_exact_transform_and_never_above_the_reference():
    """BASELINE.json configs[3] at FULL size: 5 M points -> 400 x 400 x 200 voxels @ 1 cm, 25-cell truncation (4 y chunks
    per plane).  All 32 M voxels: occupancy == the reference's (oracle port of addPointsToField), distances == the exact
    truncated EDT (scipy), GPU <= the reference's propagation everywhere; the differing voxels are counted."""
    from scipy import ndimage
    c4 = workloads.CONFIGS["C4"]
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c4["size"], origin=c4["origin"],
                                           resolution=c4["resolution"], max_distance=c4["max_distance"])
    pts = workloads.make_cloud("C4")
    eng.field_add_points(pts)
    g = eng.field_get_d2()
    assert g.shape == (400, 400, 200) and eng.max_distance_sq == 625
    exact = ndimage.distance_transform_edt(g != 0)
    exact = np.minimum(np.rint(exact * exact).astype(np.int64), eng.max_distance_sq)
    n_bad = int((g != exact).sum())
    assert n_bad == 0, "%d of %d voxels differ from the exact truncated EDT" % (n_bad, g.size)
    del exact
    sx, sy, sz = c4["size"]
    ox, oy, oz = c4["origin"]
    ref = PropagationDistanceField(sx, sy, sz, c4["resolution"], ox - sx / 2, oy - sy / 2, oz - sz / 2, c4["max_distance"])
    ref.add_points(pts)
    o = ref.d2()
    assert np.array_equal(g == 0, o == 0), "occupancy differs from the reference's addPointsToField"
    assert (g <= o).all(), "GPU above the reference's propagation somewhere"
    print("C4: %d occupied voxels; %d of %d voxels below the reference's propagation (max delta %d)" %
          (int((g == 0).sum()), int((g != o).sum()), g.size, int((o - g).max())))
    # the rebuild is a function of the occupancy set only: a second build from shuffled points gives the same field
    eng.field_reset()
    eng.field_add_points(pts[np.random.default_rng(0).permutation(len(pts))])
    assert np.array_equal(eng.field_get_d2(), g)
    eng.close()


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32])
def test_latency_path_is_bit_exact(c1, n):
    """Batches of up to 32 states take ONE kernel launch (latencyCheckKernel, exact FP64, results polled in mapped host
    memory); same booleans as the oracle and as the large-batch path."""
    eng, env, cloud, q = c1
    for start in (0, 100, 977):
        qs = q[start:start + n]
        for flags in (ALL, mb.CHECK_WORLD, mb.CHECK_SELF | mb.CHECK_INTRA):
            ref, _ = env.check(qs, flags)
            got = eng.check_batch(qs, flags)
            assert np.array_equal(got, ref == 0)
            assert eng.stats()["n_states"] == n and eng.stats()["n_valid"] == int(got.sum())
    big = eng.check_batch(q[:4096], ALL)
    for i in range(0, 64, 7):
        assert eng.check_batch(q[i:i + 1], ALL)[0] == big[i]


def test_latency_path_many_calls_and_wide_group():
    """1 000 single-state calls in a row (the buffer and the polling are reused), and a group whose states do not fit the
    kernel parameters (18 variables x 2 states > 32 doubles: read from mapped host memory)."""
    eng, env = make_pair("dualarm", "whole_body", (3.0, 3.0, 2.5), (0, 0, 1.0), 0.05, 0.25)
    cloud = workloads.clutter_cloud(6, 20000, 12, (-1.3, -1.3, 0.1), (1.3, 1.3, 2.0), 0.75)
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    lo, hi = np.maximum(lo, -3.0), np.minimum(hi, 3.0)
    q = workloads.random_states(3, 1000, lo, hi)
    ref, _ = env.check(q, ALL)
    got = np.array([eng.check_batch(q[i:i + 1], ALL)[0] for i in range(len(q))])
    assert np.array_equal(got, ref == 0) and 0 < got.sum() < len(q)
    for i in range(0, 90, 3):
        assert np.array_equal(eng.check_batch(q[i:i + 3], ALL), ref[i:i + 3] == 0)
    eng.close()


def test_float32_states(c1):
    """b200sv_check_batch_f32: every value is widened on the device, so the result is the oracle's for q = double(float)."""
    eng, env, cloud, q = c1
    qf = q.astype(np.float32)
    ref, _ = env.check(qf.astype(np.float64), ALL)
    got = eng.check_batch_f32(qf, ALL)
    assert np.array_equal(got, ref == 0)
    # ragged sizes around the chunking of the copy pipeline
    for n in (1, 33, 4097):
        assert np.array_equal(eng.check_batch_f32(qf[:n], ALL), ref[:n] == 0)


def _srdf_without(srdf, a, b):
    keep = [ln for ln in srdf.splitlines() if not ("disable_collisions" in ln and ('"%s"' % a) in ln and ('"%s"' % b) in ln)]
    assert len(keep) < len(srdf.splitlines())
    return "\n".join(keep)


def test_acm_change_on_a_live_context():
    """compareCacheEntryToAllowedCollisionMatrix (collision_env_distance_field.cpp:1314-1370): when the ACM changes the
    reference regenerates the cache entry.  b200sv_set_acm_masks re-derives the masks in place; the result must equal a
    FRESH oracle env built with the changed ACM -- and differ from the old one."""
    urdf, srdf = read(RES, "panda_spheres.urdf"), read(RES, "panda.srdf")
    geo = ((1.0, 1.0, 1.0), (0.0, 0.0, 0.5), 0.02, 0.25)
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "panda_arm", *geo)
    env0 = OracleEnv(RobotModel(urdf, srdf), "panda_arm", *geo)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(21, 20000, lo, hi)
    flags = mb.CHECK_SELF | mb.CHECK_INTRA
    ref0, _ = env0.check(q, flags)
    eng.field_set_d2(env0.self_field.d2(), mb.FIELD_SELF)
    assert np.array_equal(eng.check_batch(q, flags), ref0 == 0)
    # enable a pair the SRDF disables and disable two it enables (pairs whose outcome depends on the state; the SRDF's
    # finger-finger pair is left alone: at the default opening the finger spheres touch EXACTLY, a zero-clearance tie
    # whose boolean is decided by the last bit of the world-frame FK -- the north star's "within 1e-4 m of zero" case)
    srdf1 = _srdf_without(srdf, "panda_link2", "panda_link6").replace(
        "</robot>", '<disable_collisions link1="panda_link5" link2="panda_hand" reason="test"/>\n'
                    '<disable_collisions link1="panda_link1" link2="panda_link5" reason="test"/>\n</robot>')
    env1 = OracleEnv(RobotModel(urdf, srdf1), "panda_arm", *geo)
    _, model1 = env1.flat_arrays()
    eng.set_acm_masks(model1["self_enabled"], model1["intra_enabled"])
    ref1, _ = env1.check(q, flags)
    assert ((ref1 == 0) != (ref0 == 0)).sum() > 100
    assert np.array_equal(eng.check_batch(q, flags), ref1 == 0)
    assert np.array_equal(eng.check_batch(q[:16], flags), ref1[:16] == 0)     # the latency path sees the new tables too
    # and back
    _, model0 = env0.flat_arrays()
    eng.set_acm_masks(model0["self_enabled"], model0["intra_enabled"])
    assert np.array_equal(eng.check_batch(q, flags), ref0 == 0)
    eng.close()


def test_base_state_change_on_a_live_context():
    """compareCacheEntryToState (:1254-1312): a joint outside the group moved -> new cache entry: constant link transforms
    and the self field (non-group links' points at the new state, :907-953).  b200sv_set_base_state does it in place; the
    world field is untouched.  Compared with a FRESH oracle env at the new base state."""
    urdf, srdf = read(RES, "dualarm_spheres.urdf"), read(RES, "dualarm.srdf")
    geo = ((3.0, 3.0, 2.5), (0.0, 0.0, 1.0), 0.05, 0.25)
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "left_arm", *geo)
    model = RobotModel(urdf, srdf)
    env0 = OracleEnv(model, "left_arm", *geo)
    cloud = workloads.clutter_cloud(6, 20000, 12, (-1.3, -1.3, 0.1), (1.3, 1.3, 2.0), 0.75)
    eng.field_add_points(cloud)
    world_before = eng.field_get_d2()
    lo, hi = eng.group_bounds()
    q = workloads.random_states(22, 20000, lo, hi)
    # move every variable outside the group (right arm, torso, ...) inside its bounds
    base = np.array(env0.base_positions)
    rng = np.random.default_rng(5)
    gv = set(int(v) for v in env0.arr["group_var_index"])
    right = RobotModel(urdf, srdf).groups["right_arm"]
    for v in right.variable_index_list:
        if v not in gv:
            base[v] = rng.uniform(-1.0, 1.0)
    env1 = OracleEnv(RobotModel(urdf, srdf), "left_arm", *geo, base_positions=base)
    assert not np.array_equal(env1.self_field.d2(), env0.self_field.d2())
    eng.set_base_state(base)
    g_self, o_self = eng.field_get_d2(mb.FIELD_SELF), env1.self_field.d2()
    assert np.array_equal(g_self == 0, o_self == 0) and (g_self <= o_self).all()
    assert np.array_equal(eng.field_get_d2(), world_before)
    assert np.allclose(eng.fk_batch(q[:32]), env1.fk(q[:32]), atol=1e-12)
    # validity on identical fields
    env1.world.add_points(cloud)
    eng.field_set_d2(env1.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(o_self, mb.FIELD_SELF)
    ref1, _ = env1.check(q, ALL)
    assert np.array_equal(eng.check_batch(q, ALL), ref1 == 0)
    env0.world.add_points(cloud)
    ref0, _ = env0.check(q, ALL)
    assert ((ref0 == 0) != (ref1 == 0)).sum() > 0
    eng.close()


def test_field_export_import_between_contexts():
    torch = pytest.importorskip("torch")
    eng_a, env, cloud, q = config_pair("C1")
    eng_b, _, _, _ = config_pair("C1")
    eng_a.field_add_points(cloud)
    buf = torch.empty(eng_a.field_export_size(), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for which in (mb.FIELD_WORLD, mb.FIELD_SELF):
        eng_a.field_export_device(buf.data_ptr(), which, s)
        eng_b.field_import_device(buf.data_ptr(), which, s)
    torch.cuda.synchronize()
    assert np.array_equal(eng_b.field_get_d2(), eng_a.field_get_d2())
    assert np.array_equal(eng_b.check_batch(q), eng_a.check_batch(q))
    # the imported occupancy is live: adding points to B continues from it
    extra = cloud[:10] + 0.07
    eng_a.field_add_points(extra)
    eng_b.field_add_points(extra)
    assert np.array_equal(eng_b.field_get_d2(), eng_a.field_get_d2())
    eng_a.close()
    eng_b.close()


def test_several_devices_behind_one_handle(c1):
    """b200sv_create_multi_from_urdf / b200sv_check_batch_multi: states sharded in blocks of whole bitmap words, the field
    built on the first device and copied to the others.  With one visible GPU the two contexts share it (same code path)."""
    torch = pytest.importorskip("torch")
    eng, env, cloud, q = c1
    n_gpu = torch.cuda.device_count()
    devices = [0, 1] if n_gpu >= 2 else [0, 0]
    cfg = workloads.CONFIGS["C1"]
    me = mb.MultiDeviceEngine.from_urdf(devices, read(RES, "panda_spheres.urdf"), read(RES, "panda.srdf"), cfg["group"],
                                        cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    assert me.n_devices == 2
    me.field_add_points(cloud)
    own = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=cfg["size"], origin=cfg["origin"],
                                           resolution=cfg["resolution"], max_distance=cfg["max_distance"])
    own.field_add_points(cloud)
    for c in me.contexts:                              # every device holds the field device 0 built
        assert np.array_equal(c.field_get_d2(), own.field_get_d2())
    qq = np.concatenate([q] * 3)[:25000 + 13]           # ragged: the second block is shorter, the tail is not a whole word
    want = own.check_batch(qq)
    assert np.array_equal(me.check_batch(qq), want)
    assert me.stats()["n_states"] == len(qq) and me.stats()["n_valid"] == int(want.sum())
    for n in (1, 31, 33, 64, 4096, 8193):
        assert np.array_equal(me.check_batch(qq[:n]), want[:n])
    # validity against the oracle on the reference's own field, injected into context 0 and replicated
    for which, f in ((mb.FIELD_WORLD, env.world), (mb.FIELD_SELF, env.self_field)):
        me.contexts[0].field_set_d2(f.d2(), which)
        me.field_sync(which)
    ref, _ = env.check(q, ALL)
    assert np.array_equal(me.check_batch(q), ref == 0)
    own.close()
    me.close()


def test_fk_batch_device_resident(c1):
    torch = pytest.importorskip("torch")
    eng, env, cloud, q = c1
    n = 5000
    d_q = torch.from_numpy(q[:n]).cuda()
    out = torch.empty((n, eng.n_links, 16), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for prec, tol in ((64, 1e-12), (32, 1e-5)):
        eng.fk_batch_device(d_q.data_ptr(), n, prec, out.data_ptr(), s)
        torch.cuda.synchronize()
        t = out.cpu().numpy().reshape(n, eng.n_links, 4, 4).transpose(0, 1, 3, 2)
        assert np.abs(t - env.fk(q[:n])).max() < tol


def test_device_entry_points_are_ordered_against_the_context_stream(c1):
    """A device-resident check on the caller's stream followed at once by host-buffer calls (the context's own stream) and
    by another device-resident check on a second caller stream: all share the context's recheck list and counter; the
    library orders them with events.  Every result must be the oracle's."""
    torch = pytest.importorskip("torch")
    eng, env, cloud, q = c1
    ref, _ = env.check(q, ALL)
    d_q = torch.from_numpy(q).cuda()
    n = len(q)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    bm1 = torch.zeros(4 * ((n + 31) // 32), dtype=torch.uint8, device="cuda")
    bm2 = torch.zeros_like(bm1)
    torch.cuda.synchronize()
    for _ in range(20):
        eng.check_batch_device(d_q.data_ptr(), n, ALL, bm1.data_ptr(), s1.cuda_stream)
        host = eng.check_batch(q[:5000], ALL)
        eng.check_batch_device(d_q.data_ptr(), n, ALL, bm2.data_ptr(), s2.cuda_stream)
        one = eng.check_batch(q[:1], ALL)
        torch.cuda.synchronize()
        for bm in (bm1, bm2):
            got = np.unpackbits(bm.cpu().numpy(), bitorder="little")[:n].astype(bool)
            assert np.array_equal(got, ref == 0)
        assert np.array_equal(host, ref[:5000] == 0) and one[0] == (ref[0] == 0)


def test_c_program_drives_two_devices(tmp_path):
    """tests/c_multi_check.c, plain C99 against include/b200sv.h: two devices (or one device twice) behind one handle."""
    import subprocess
    torch = pytest.importorskip("torch")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "moveit2_b200")
    exe = str(tmp_path / "c_multi_check")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-O1", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "c_multi_check.c"), "-L", libdir, "-lb200sv", "-lm",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, os.path.join(RES, "panda_spheres.urdf"), os.path.join(RES, "panda.srdf"),
                          str(torch.cuda.device_count())], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c multi ok" in out.stdout
    print(out.stdout.strip())


def test_check_motions_refuses_planar_and_floating_groups():
    """JointModelGroup::distance / interpolate are not per-variable for planar and floating joints (Euclidean + weighted
    angle, slerp): b200sv_check_motions refuses such groups instead of interpolating their variables linearly."""
    eng, env = make_pair("dualarm", "whole_body", (3.0, 3.0, 2.5), (0, 0, 1.0), 0.05, 0.25)
    lo, hi = eng.group_bounds()
    lo, hi = np.maximum(lo, -3.0), np.minimum(hi, 3.0)
    q = workloads.random_states(3, 8, lo, hi)
    with pytest.raises(mb.B200svError) as e:
        eng.check_motions(q[:4], q[4:], 0.1)
    assert e.value.code == -3 and "planar" in str(e.value)
    eng.close()


def test_field_update_bracket_runs_one_transform():
    """b200sv_field_begin_update / end_update: several occupancy edits, ONE rebuild; the result equals the edits applied one
    by one (the reference's updateDistanceObject: remove the object's old points, add the new ones)."""
    eng_a, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
    eng_b, _ = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
    rng = np.random.default_rng(8)
    a, b, c = (rng.uniform([-0.4, -0.4, 0.1], [0.4, 0.4, 0.9], size=(300, 3)) for _ in range(3))
    for e in (eng_a, eng_b):
        e.field_add_points(a)
    before = eng_a.field_get_d2()
    eng_a.field_begin_update()
    eng_a.field_remove_points(a[:150])
    eng_a.field_add_points(b)
    eng_a.field_add_shape(mb.SHAPE_BOX, (0.1, 0.08, 0.06), np.eye(4), lattice_frame=mb.LATTICE_WORLD)
    assert np.array_equal(eng_a.field_get_d2(), before)        # reads inside the bracket see the state before it
    eng_a.field_add_points(c)
    eng_a.field_end_update()
    eng_b.field_remove_points(a[:150])
    eng_b.field_add_points(b)
    eng_b.field_add_shape(mb.SHAPE_BOX, (0.1, 0.08, 0.06), np.eye(4), lattice_frame=mb.LATTICE_WORLD)
    eng_b.field_add_points(c)
    assert np.array_equal(eng_a.field_get_d2(), eng_b.field_get_d2())
    eng_a.field_begin_update()
    eng_a.field_end_update()                                    # an empty bracket is a no-op
    assert np.array_equal(eng_a.field_get_d2(), eng_b.field_get_d2())
    eng_a.close()
    eng_b.close()


def _pose(rpy, xyz):
    cr, sr, cp, sp, cy, sy = np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1]), np.cos(rpy[2]), np.sin(rpy[2])
    T = np.eye(4)
    T[:3, :3] = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                          [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], [-sp, cp * sr, cp * cr]])
    T[:3, 3] = xyz
    return T


def _hull_planes_of_a_skewed_wedge():
    """A closed convex polytope with faces in general position: 7 unit-normal planes around (0.02, -0.01, 0.03)."""
    rng = np.random.default_rng(12)
    n = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0.6, 0.5, 0.62]], np.float64)
    n += 0.15 * rng.standard_normal(n.shape)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    centre = np.array([0.02, -0.01, 0.03])
    offs = np.array([0.11, 0.09, 0.07, 0.10, 0.06, 0.08, 0.05])
    w = -(n @ centre) - offs                       # n . p + w <= 0  <=>  n . (p - centre) <= offs
    return np.concatenate([n, w[:, None]], axis=1), np.array([centre[0], centre[1], centre[2], 0.2])


@pytest.mark.parametrize("frame", ["world", "body"])
def test_field_add_convex_equals_the_restated_containment(frame):
    """b200sv_field_add_convex (bodies::ConvexMesh::containsPoint over findInternalPointsConvex's lattice, on the device)
    against the numpy restatement, voxel for voxel; then removed again."""
    from oracle.collision_model import find_internal_points_convex_planes
    planes, bs = _hull_planes_of_a_skewed_wedge()
    pose = _pose((0.3, -0.7, 1.1), (0.12, -0.08, 0.55))
    eng, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.01, 0.25)
    lf = mb.LATTICE_WORLD if frame == "world" else mb.LATTICE_BODY
    eng.field_add_convex(planes, bs, pose, lattice_frame=lf)
    pts = find_internal_points_convex_planes(planes, bs, pose, 0.01, frame)
    assert len(pts) > 500
    env.world.add_points(pts)
    g, o = eng.field_get_d2(), env.world.d2()
    assert int((o == 0).sum()) > 300 and np.array_equal(g == 0, o == 0)
    eng.field_add_convex(planes, bs, pose, lattice_frame=lf, remove=True)
    assert int((eng.field_get_d2() == 0).sum()) == 0
    eng.close()


def test_field_add_octree_leaves_equals_the_restated_expansion():
    """DistanceField::addOcTreeToField behind octomap's leaf iteration: leaves of 1, 2, 4 and 8 cells, some sticking out of
    the field; the lattice of a large leaf is the reference's accumulated `x += resolution_` loops."""
    from oracle.collision_model import octree_leaf_points
    rng = np.random.default_rng(13)
    res = 0.02
    leaves = []
    for size_cells, count in ((1, 400), (2, 60), (4, 20), (8, 6)):
        size = res * size_cells
        # octomap leaf centres sit on the (float) grid of their depth
        c = (np.floor(rng.uniform([-0.55, -0.55, -0.05], [0.55, 0.55, 1.05], size=(count, 3)) / size) + 0.5) * size
        c = c.astype(np.float32).astype(np.float64)          # octomap::point3d holds floats
        leaves.append(np.concatenate([c, np.full((count, 1), size)], axis=1))
    leaves = np.concatenate(leaves)
    eng, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), res, 0.25)
    eng.field_add_octree_leaves(leaves)
    env.world.add_points(octree_leaf_points(leaves, res))
    g, o = eng.field_get_d2(), env.world.d2()
    assert int((o == 0).sum()) > 2000 and np.array_equal(g == 0, o == 0)
    eng.close()


@pytest.mark.parametrize("robot,group", [("panda", "panda_arm"), ("dualarm", "arms"), ("dualarm", "whole_body")])
def test_fp32_budget_is_derived_and_covers_the_fast_kernels_error(robot, group):
    """The budget every FP32 decision carries is derived from the model (b200sv_get_fp32_budget; deriveFp32Eps: chain depth,
    joint origins and ranges, field extent), not a constant.  The fast kernel's own FK + posing (b200sv_debug_fast_centres
    runs the same device function) must stay within it on random states and on the corners of the joint box."""
    eng, env = make_pair(robot, group, (3.0, 3.0, 2.5), (0, 0, 1.0), 0.02, 0.25)
    used, derived = eng.fp32_budget()
    assert used == derived and 1e-6 < derived < 2e-4
    lo, hi = eng.group_bounds()
    lo, hi = np.maximum(lo, -2.0 * np.pi), np.minimum(hi, 2.0 * np.pi)
    q = workloads.random_states(77, 3000, lo, hi)
    nc = 2 ** min(len(lo), 7)
    q[:nc] = np.array([[hi[k] if (i >> (k % 7)) & 1 else lo[k] for k in range(len(lo))] for i in range(nc)])
    fast = eng.debug_fast_centres(q)
    exact = np.stack([env.sphere_centers(q[i]) for i in range(len(q))])      # the oracle's FP64 centres
    err = np.abs(fast - exact).max()
    print("%s/%s: worst FP32 centre error %.3e m, budget %.3e m (%.2f of it)" % (robot, group, err, used, err / used))
    assert err <= used
    eng.close()


def test_device_entry_points_stay_inside_their_buffers(c1):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_unavailable.txt), so the out-of-bounds check is
    done by hand: every caller-owned device buffer is surrounded by canary words that must survive the call, for ragged
    sizes (bitmap words: whole 32-bit words are written, nothing beyond; FK output: exactly n * n_links * 16 doubles)."""
    torch = pytest.importorskip("torch")
    eng, env, cloud, q = c1
    s = torch.cuda.current_stream().cuda_stream
    guard = 64
    for n in (1, 31, 32, 33, 1000, 4097):
        d_q = torch.from_numpy(q[:n]).cuda()
        words = (n + 31) // 32
        bm = torch.full((guard + 4 * words + guard,), 0xA5, dtype=torch.uint8, device="cuda")
        eng.check_batch_device(d_q.data_ptr(), n, ALL, bm.data_ptr() + guard, s)
        out = torch.full((guard + n * eng.n_links * 16 + guard,), -7.25, dtype=torch.float64, device="cuda")
        eng.fk_batch_device(d_q.data_ptr(), n, 32, out.data_ptr() + 8 * guard, s)
        torch.cuda.synchronize()
        b = bm.cpu().numpy()
        assert (b[:guard] == 0xA5).all() and (b[guard + 4 * words:] == 0xA5).all()
        ref, _ = env.check(q[:n], ALL)
        assert np.array_equal(np.unpackbits(b[guard:guard + 4 * words], bitorder="little")[:n].astype(bool), ref == 0)
        o = out.cpu().numpy()
        assert (o[:guard] == -7.25).all() and (o[guard + n * eng.n_links * 16:] == -7.25).all()
        assert not (o[guard:guard + n * eng.n_links * 16] == -7.25).any()
    # field export: exactly field_export_size bytes
    size = eng.field_export_size()
    buf = torch.full((guard + size + guard,), 0x5A, dtype=torch.uint8, device="cuda")
    eng.field_export_device(buf.data_ptr() + guard, mb.FIELD_WORLD, s)
    torch.cuda.synchronize()
    b = buf.cpu().numpy()
    assert (b[:guard] == 0x5A).all() and (b[guard + size:] == 0x5A).all()


def test_collision_gradients_with_per_link_posed_fields(c1):
    """getSelfProximityGradients with a non-empty ACM (collision_env_distance_field.cpp:386-411; what CHOMP runs: its
    GroupStateRepresentation is created with the scene's ACM, chomp_optimizer.cpp:102): every entry also consults the
    PosedDistanceFields of the other links, through PosedDistanceField::getDistanceGradient EXACTLY as written (rotation
    transposed without the translation in, translation added to the gradient out, an out-of-bounds query ends the entry's
    sphere loop).  The reference's own link fields (oracle: its propagation) are injected; same bar as the plain test."""
    eng, env, cloud, q = c1
    n = 2000
    fields, enabled = env.build_link_fields()
    for g, f in enumerate(fields):
        if f is None:
            eng.set_link_field(g, (0.1, 0.1, 0.1), (0, 0, 0), None)
        else:
            eng.set_link_field(g, f.size, f.corner, f.d2())
            back, corner = eng.get_link_field(g)
            assert back.shape == f.num_cells and np.array_equal(back, f.d2()) and np.allclose(corner, f.corner)
    eng.set_gradient_link_mask(enabled)
    try:
        d, g, t, c, loc = eng.collision_gradients(q[:n])
        rd, rg, rt, rc = env.collision_gradients(q[:n], with_link_fields=True)
        pd, pg, pt, pc = env.collision_gradients(q[:n])
        changed = int((rd != pd).sum())
        assert changed > 100, "the posed fields should change some entries of the test scene (%d)" % changed
        both_none = (t == 0) & (rt == 0)
        close = np.isclose(d, rd, rtol=0, atol=1e-9) | both_none
        assert close.mean() > 0.9999, "%d of %d sphere distances differ" % ((~close).sum(), close.size)
        same = close & (t == rt)
        assert same.mean() > 0.9995
        assert np.allclose(g[same], rg[same], rtol=0, atol=1e-9)
        assert np.isclose(c, rc, rtol=0, atol=1e-9).mean() > 0.999
        # the entries the posed fields changed are reproduced too, not just the untouched majority
        touched = (rd != pd) & ~both_none
        assert np.isclose(d[touched], rd[touched], rtol=0, atol=1e-9).mean() > 0.99
        assert np.allclose(g[touched & same], rg[touched & same], rtol=0, atol=1e-9)
    finally:
        eng.set_gradient_link_mask(None)
    # switched off again: the plain result
    d0 = eng.collision_gradients(q[:200])[0]
    assert np.isclose(d0, env.collision_gradients(q[:200])[0], rtol=0, atol=1e-9).mean() > 0.9999


def test_link_fields_built_on_the_device_are_the_exact_transform():
    """b200sv_build_link_fields (URDF path): every link's field from its interior lattice points with the device EDT --
    same geometry as the reference's (size = bounding-sphere diameter, corner = centre - size / 2, n = int(size / res)),
    occupancy equal to the reference's, distances the exact truncated transform (<= the reference's propagation)."""
    from scipy import ndimage
    eng, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
    fields, _ = env.build_link_fields()
    eng.build_link_fields()
    seen = 0
    for g, f in enumerate(fields):
        d2, corner = eng.get_link_field(g)
        if f is None:
            assert d2 is None
            continue
        o = f.d2()
        assert d2.shape == f.num_cells and np.allclose(corner, f.corner, atol=1e-12)
        assert np.array_equal(d2 == 0, o == 0) and (d2 <= o).all()
        if (d2 == 0).any():
            exact = np.minimum(np.rint(ndimage.distance_transform_edt(d2 != 0) ** 2).astype(np.int64), eng.max_distance_sq)
            assert np.array_equal(d2, exact)
        seen += 1
    assert seen >= 8
    eng.close()


# ---- an attached body of several shapes is culled as a whole (collision_env_distance_field.cpp:455-507) -------------------
def _pole_env(with_body_ids):
    """A 4 m pole held by the hand, modelled as two shapes of ONE body: the pole itself (bounding radius 2 m, its collision
    spheres near the hand) and a small collar.  With bounding radii this large the reference's squared-distance-vs-radius-sum
    cull (types.cpp:416-417) rejects the pole's own bounding sphere although its spheres touch a link -- but the collar's
    passes, and then the reference tests ALL the body's spheres."""
    env = OracleEnv(RobotModel(read(RES, "panda_spheres.urdf"), read(RES, "panda.srdf")), "panda_arm", (1, 1, 1), (0, 0, 0.5),
                    0.02, 0.25)
    touch = {"panda_hand", "panda_leftfinger", "panda_rightfinger"}
    pole = [((0.0, 0.0, 0.25 + 0.05 * k), 0.025) for k in range(8)]
    env.attach_bodies([dict(link="panda_hand", name="pole", spheres=pole, bs_center=(0, 0, 2.2), bs_radius=2.0, touch_links=touch),
                       dict(link="panda_hand", name="pole", spheres=[((0, 0, 0.3), 0.025)], bs_center=(0, 0, 0.3),
                            bs_radius=0.1, touch_links=touch)])
    robot, model = env.flat_arrays()
    if not with_body_ids:
        env.cm.body_id = None   # every entry a body of its own: culled shape by shape
        model.pop("body_id")
    eng = mb.StateValidityEngine.from_arrays(robot, model, (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
    return eng, env


def test_a_body_of_several_shapes_is_culled_as_a_whole():
    eng, env = _pole_env(True)
    eng0, env0 = _pole_env(False)
    lo, hi = env.joint_bounds()
    q = workloads.random_states(77, 200_000, lo, hi)
    I = mb.CHECK_INTRA
    whole, _ = env.check(q, I)
    by_shape, _ = env0.check(q, I)
    # the two rules differ on this body, one-sidedly: the whole-body rule tests more spheres
    assert ((whole != 0) & (by_shape == 0)).sum() > 500 and not ((whole == 0) & (by_shape != 0)).any()
    for mode in (0, mb.CHECK_FP64):
        assert np.array_equal(eng.check_batch(q, I | mode), whole == 0), mode
        assert np.array_equal(eng0.check_batch(q, I | mode), by_shape == 0), mode
    assert np.array_equal(eng.check_batch(q[:5], I), whole[:5] == 0)   # the single-launch latency path
    differ = np.flatnonzero(whole != by_shape)[:64]
    assert np.array_equal(eng.check_batch(q[differ][:8], I), whole[differ][:8] == 0)
    col, cnt, _ = eng.check_contacts(q[differ], I, 4, 2)
    rcol, rcnt, _ = env.check_contacts(q[differ], I, 4, 2)
    assert np.array_equal(col, rcol) and np.array_equal(cnt, rcnt) and col.all()
    ok = eng.check_motions(q[differ[:-1]], q[differ[1:]], 0.05, flags=I)[0]
    eng.close()
    eng0.close()
    assert not ok.any()   # both end states of every edge collide


# ---- the interleaved {world, self} compact field stays consistent whichever field is rebuilt, in whatever order -----------
def _compact_of(eng, which, torch):
    """This field's elements of the compact field (what the check kernel gathers), read back through the export buffer."""
    buf = torch.empty(eng.field_export_size(), dtype=torch.uint8, device="cuda")
    eng.field_export_device(buf.data_ptr(), which, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    nv = int(np.prod(eng.num_cells))
    nx, ny, wz = eng.num_cells[0], eng.num_cells[1], (eng.num_cells[2] + 31) // 32
    off = ((nx * ny * wz * 4 + 255) & ~255) + ((nv * 2 + 255) & ~255)
    return buf[off:off + nv].cpu().numpy().reshape(eng.num_cells)


def _expected_compact(d2, max_d2):
    want = np.minimum(d2, 255).astype(np.uint8)   # 8-bit elements: min(d^2, 255)
    free = min(max_d2, 255)
    want[0, :, :] = want[-1, :, :] = want[:, 0, :] = want[:, -1, :] = want[:, :, 0] = want[:, :, -1] = free  # the shell
    return want


def test_compact_field_survives_rebuilds_of_either_field_in_any_order():
    """The x pass of a rebuild writes whole {world, self} words where the OTHER field's x plane is known to be free and merges
    elsewhere: after every step both fields' compact elements must still be min(d^2, 255) of their own d2."""
    torch = pytest.importorskip("torch")
    urdf, srdf = read(RES, "dualarm_spheres.urdf"), read(RES, "dualarm.srdf")
    geo = ((3.0, 3.0, 2.5), (0.0, 0.0, 1.0), 0.05, 0.25)
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "left_arm", *geo)
    assert eng.info.field_bytes_per_voxel == 1
    max_d2 = eng.max_distance_sq
    env = OracleEnv(RobotModel(urdf, srdf), "left_arm", *geo)

    def both_ok(tag):
        for which in (mb.FIELD_WORLD, mb.FIELD_SELF):
            got = _compact_of(eng, which, torch)
            assert np.array_equal(got, _expected_compact(eng.field_get_d2(which), max_d2)), (tag, which)

    both_ok("created: self built, world free")
    assert (eng.field_get_d2(mb.FIELD_SELF) == 0).any()
    # a world cloud in ONE corner: most x planes of the world field stay free
    cloud = workloads.clutter_cloud(6, 20000, 4, (0.6, 0.6, 0.1), (1.3, 1.3, 1.0), 0.0, floor=False)
    eng.field_add_points(cloud)
    both_ok("world built next to a busy self field")
    base = np.array(env.base_positions)
    gv = set(int(v) for v in env.arr["group_var_index"])
    for v in RobotModel(urdf, srdf).groups["right_arm"].variable_index_list:
        if v not in gv:
            base[v] = 0.7
    eng.set_base_state(base)
    both_ok("self rebuilt next to a partly busy world field")
    eng.field_reset()
    both_ok("world reset")
    eng.field_add_points(cloud[:100] - np.array([1.5, 1.5, 0.0]))
    both_ok("world rebuilt in the opposite corner")
    eng.field_set_d2(eng.field_get_d2(mb.FIELD_SELF), mb.FIELD_WORLD)   # injected values: every plane may be busy
    eng.set_base_state(env.base_positions)
    both_ok("self rebuilt next to an injected world field")
    eng.close()


# ---- single states back to back: the resident kernel and its mailbox ------------------------------------------------------
def test_single_states_back_to_back_through_the_resident_kernel():
    """n = 1 calls are answered by a kernel that stays resident for a short while (b200sv_check_batch -> checkLinger): every
    answer equals the oracle's, other entry points in between make the kernel step aside (field rebuilds change what it must
    read), and after an idle gap the next call simply launches a new one."""
    import time
    eng, env, cloud, q = config_pair("C1", n_states=3000)
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    ref, _ = env.check(q, ALL)
    got = np.array([eng.check_batch(q[i:i + 1], ALL)[0] for i in range(1500)])
    assert np.array_equal(got, ref[:1500] == 0)
    assert eng.stats()["n_states"] == 1
    for flags in (mb.CHECK_WORLD, mb.CHECK_SELF, mb.CHECK_INTRA):
        r, _ = env.check(q[:200], flags)
        assert np.array_equal(np.array([eng.check_batch(q[i:i + 1], flags)[0] for i in range(200)]), r == 0), flags
    # the world changes between single-state calls: the resident kernel must not answer from the old field
    extra = workloads.clutter_cloud(9, 1500, 3, (-0.4, -0.4, 0.2), (0.4, 0.4, 0.9), 0.15, floor=False)
    env.world.add_points(extra)
    for i in range(1500, 1600):
        if i == 1550:
            eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
            ref2, _ = env.check(q, ALL)
            assert (ref2[1550:1700] != ref[1550:1700]).any(), "the added cloud should change some answers"
        want = (ref if i < 1550 else ref2)[i] == 0
        assert eng.check_batch(q[i:i + 1], ALL)[0] == want, i
    # batches, edges and device-stream calls in between; idle gaps longer than the kernel's linger window
    for i in range(1600, 1700):
        if i % 10 == 0:
            assert np.array_equal(eng.check_batch(q[:300], ALL), ref2[:300] == 0)
        if i % 25 == 0:
            time.sleep(0.002)
        assert eng.check_batch(q[i:i + 1], ALL)[0] == (ref2[i] == 0), i
    eng.close()


def test_resident_kernels_of_two_contexts_from_two_threads():
    import threading
    pairs = [config_pair("C1", n_states=600) for _ in range(2)]
    out = [None, None]

    def work(k):
        eng, env, cloud, q = pairs[k]
        eng.field_add_points(cloud[: 1000 * (k + 1)])
        batch = eng.check_batch(q, ALL)
        single = np.array([eng.check_batch(q[i:i + 1], ALL)[0] for i in range(len(q))])
        out[k] = np.array_equal(batch, single)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert out == [True, True]
    for eng, _, _, _ in pairs:
        eng.close()


# ---- a few edges per call: ONE launch, the batch sized on the host --------------------------------------------------------
def test_a_few_edges_per_call_take_one_launch_and_equal_the_oracle(c1):
    """b200sv_check_motions with up to 64 edges (an OMPL MotionValidator validates one per checkMotion): the host computes the
    edges' state counts itself and launches one CTA per visited state.  Same answers as the oracle's edge restatement and as
    the general (device-sized) path, continuous joints, passive variables and the per-edge cap included."""
    from oracle import motion
    eng, env, cloud, q = c1
    lo, hi = eng.group_bounds()
    V = len(lo)
    rng = np.random.default_rng(11)
    m = 192
    a = rng.uniform(lo, hi, size=(m, V))
    b = np.where(rng.random((m, 1)) < 0.5, rng.uniform(lo, hi, size=(m, V)), np.clip(a + rng.normal(0, 0.05, size=(m, V)), lo, hi))
    b[5] = a[5]                                   # a zero-length edge: one state (its end)
    a[:, 0] = rng.uniform(-np.pi, np.pi, m)
    b[:, 0] = rng.uniform(-np.pi, np.pi, m)
    cont = np.zeros(V, np.uint8)
    cont[0] = 1
    fac = np.ones(V)
    fac[3] = 0.0
    L = 0.01 * float(np.sum(hi - lo))
    for args in ((None, None, 0), (cont, fac, 0), (cont, fac, 5)):
        rv, rf, rn = motion.check_motions(env, a, b, L, *args)
        big_v, big_f, big_n = eng.check_motions(a, b, L, *args)          # 192 edges: the general path
        assert big_n == rn and np.array_equal(big_v, rv) and np.array_equal(big_f, rf)
        for per_call in (1, 7, 64):
            got_v, got_f, got_n = [], [], 0
            for i0 in range(0, m, per_call):
                v, f, n = eng.check_motions(a[i0:i0 + per_call], b[i0:i0 + per_call], L, *args)
                got_v.append(v)
                got_f.append(f)
                got_n += n
            assert got_n == rn, (per_call, args[2])
            assert np.array_equal(np.concatenate(got_v), rv) and np.array_equal(np.concatenate(got_f), rf), (per_call, args[2])
    assert 0 < rv.sum() < m
    # an edge too long for the few-edges path (more than 1024 states) silently takes the general one
    v, f, n = eng.check_motions(a[:1], b[:1], L / 4000.0)
    rv1, rf1, rn1 = motion.check_motions(env, a[:1], b[:1], L / 4000.0)
    assert n == rn1 and n > 4096 and np.array_equal(v, rv1) and np.array_equal(f, rf1)
