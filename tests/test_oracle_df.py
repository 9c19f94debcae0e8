"""Pins the ORACLE's distance-field restatement on the reference's own golden data and test properties
(moveit_core/distance_field/test/test_distance_field.cpp, test_voxel_grid.cpp).  CPU only."""
import os

import numpy as np
import pytest

from oracle.collision_model import find_internal_points_sphere
from oracle.df import PropagationDistanceField, read_df_occupancy, write_df_occupancy

# test_distance_field.cpp:48-60
WIDTH = HEIGHT = DEPTH = 1.0
RESOLUTION = 0.1
MAX_DIST = 0.3
POINT1, POINT2, POINT3 = (0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)


def golden_occ(golden, name):
    g = np.load(os.path.join(golden, "df_golden.npz"))
    dims = tuple(int(v) for v in g[name + "_dims"])
    occ = np.unpackbits(g[name + "_occ_bits"])[: dims[0] * dims[1] * dims[2]].reshape(dims).astype(bool)
    return g[name + "_header"], dims, occ


def small():
    return PropagationDistanceField(WIDTH, HEIGHT, DEPTH, RESOLUTION, 0, 0, 0, MAX_DIST)


def check_distance_field(df, points):
    """checkDistanceField (test_distance_field.cpp:294-337): every d2==0 cell is one of `points`."""
    idx = {df.world_to_grid(*p)[0] for p in points}
    for c in np.argwhere(df.d2() == 0):
        assert tuple(int(v) for v in c) in idx


def test_voxel_grid_dims():
    # test_voxel_grid.cpp:48-57: VoxelGrid(0.02^3 @ 0.01) has 2x2x2 cells
    f = PropagationDistanceField(0.02, 0.02, 0.02, 0.01, 0, 0, 0, 0.01)
    assert f.num_cells == (2, 2, 2)
    # test_distance_field.cpp:344-351 and PERF sizes :62-70
    assert small().num_cells == (10, 10, 10) and small().max_distance_sq == 9
    big = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.25)
    assert big.num_cells == (150, 150, 200) and big.max_distance_sq == 169


def test_out_of_bounds_query():
    # test_distance_field.cpp:353-358
    df = small()
    assert abs(df.get_distance(1000.0, 1000.0, 1000.0) - MAX_DIST) < 1e-4
    d, g, inb = df.get_distance_gradient(1000.0, 1000.0, 1000.0)
    assert abs(d - MAX_DIST) < 1e-4 and not inb and not g.any()


def test_add_update_remove_points():
    # test_distance_field.cpp:360-386
    df = small()
    df.add_points([POINT1, POINT2])
    df.update_points([POINT1], [POINT2, POINT3])
    check_distance_field(df, [POINT2, POINT3])
    assert (df.d2() == 0).sum() == 2
    df.remove_points([POINT2])
    check_distance_field(df, [POINT3])
    assert (df.d2() == 0).sum() == 1
    # removal == fresh build without that point (test_distance_field.cpp:508 asserts this for the signed field)
    fresh = small()
    fresh.add_points([POINT3])
    assert np.array_equal(fresh.d2(), df.d2())


def test_gradient_points_to_obstacle():
    # test_distance_field.cpp:388-438
    df = small()
    df.add_points([POINT1])
    n = df.num_cells
    seen = False
    for z in range(1, n[2] - 1):
        for x in range(1, n[0] - 1):
            for y in range(1, n[1] - 1):
                dist = df.get_cell_distance(x, y, z)
                w = df.grid_to_world(x, y, z)
                dg, grad, inb = df.get_distance_gradient(*w)
                assert inb
                assert abs(dist - dg) < 1e-4
                if 0 < dist < MAX_DIST:
                    seen = True
                    comp = w - grad / np.linalg.norm(grad) * dist
                    assert np.all(np.abs(comp - np.array(POINT1)) <= RESOLUTION + 1e-12)
    assert seen


def test_small_df_golden(golden):
    # TestReadWrite (test_distance_field.cpp:937-955): test_small.df holds exactly the voxels of POINT1-3
    hdr, dims, occ = golden_occ(golden, "test_small")
    assert dims == (10, 10, 10) and occ.sum() == 3
    df = small()
    df.add_points([POINT1, POINT2, POINT3])
    assert np.array_equal(df.d2() == 0, occ)
    # reading the file back == building from points (areDistanceFieldsDistancesEqual, :950-951)
    df2 = small()
    df2.add_voxels(np.argwhere(occ))
    assert np.array_equal(df.d2(), df2.d2())


def test_big_df_golden(golden):
    # test_big.df = addShapeToField(sphere r=.5 @ (.5,.5,.5)) in a 150x150x200 grid (:957-963).  Pins
    # findInternalPointsConvex + Sphere::containsPoint + worldToGrid: 65 126 voxels, symmetric difference 0.
    hdr, dims, occ = golden_occ(golden, "test_big")
    assert dims == (150, 150, 200) and occ.sum() == 65126
    pts = find_internal_points_sphere((0.5, 0.5, 0.5), 0.5, 0.02)
    df = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.25)
    df.add_points(pts)
    d2 = df.d2()
    assert np.array_equal(d2 == 0, occ)
    # distances equal whether built from points or from the file's voxel list (:966-969)
    df2 = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.25)
    df2.add_voxels(np.argwhere(occ))
    assert np.array_equal(d2, df2.d2())
    # a different max distance gives a different field (:976-979)
    df3 = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.27)
    df3.add_voxels(np.argwhere(occ))
    assert not np.array_equal(d2, df3.d2())
    # sanity against the definition: min(true d2, max_d2) for a sample of voxels
    rng = np.random.default_rng(0)
    occ_idx = np.argwhere(occ)
    for c in rng.integers(0, [150, 150, 200], size=(200, 3)):
        true = int(((occ_idx - c) ** 2).sum(axis=1).min())
        assert d2[tuple(c)] >= min(true, 169)  # the reference never under-estimates
        assert d2[tuple(c)] <= 169


def test_df_stream_roundtrip(golden):
    hdr, dims, occ = golden_occ(golden, "test_small")
    keys = ("resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z")
    raw = write_df_occupancy(dict(zip(keys, hdr)), occ)
    h2, n2, idx = read_df_occupancy(raw)
    assert n2 == list(dims) and np.array_equal(idx, np.argwhere(occ))


@pytest.mark.skipif(not os.path.exists("/root/reference/moveit_core/test_big.df"), reason="reference tree absent")
def test_golden_npz_matches_reference_files(golden):
    for name in ("test_small", "test_big"):
        _, n, idx = read_df_occupancy(open("/root/reference/moveit_core/%s.df" % name, "rb").read())
        _, dims, occ = golden_occ(golden, name)
        assert list(dims) == n and np.array_equal(np.argwhere(occ), idx)


def test_propagation_matches_exact_edt_dense():
    """Dense clouds: the reference's vector propagation equals the exact truncated EDT."""
    from scipy import ndimage
    rng = np.random.default_rng(7)
    df = PropagationDistanceField(0.4, 0.4, 0.4, 0.01, 0, 0, 0, 0.13)
    pts = rng.uniform(0, 0.4, size=(10000, 3))
    df.add_points(pts)
    d2 = df.d2()
    exact = ndimage.distance_transform_edt(d2 != 0) ** 2
    exact = np.minimum(np.rint(exact).astype(np.int64), df.max_distance_sq)
    assert (d2 >= exact).all()
    assert (d2 != exact).sum() <= 2


def test_motion_restatement_discretisation():
    """oracle/motion.py against hand-computed cases of validSegmentCount / interpolate (OMPL's published algorithm)."""
    from oracle import motion
    a, b = np.array([0.0, 0.0]), np.array([1.0, 0.5])
    s = motion.edge_states(a, b, 0.5)  # distance 1.5 -> nd = 3: t = 1/3, 2/3, then b
    assert len(s) == 3 and np.allclose(s[0], [1 / 3, 0.5 / 3]) and np.array_equal(s[-1], b)
    assert len(motion.edge_states(a, a, 0.5)) == 1  # zero-length edge: only s2 is checked
    # continuous joint: 3.0 -> -3.0 goes through pi, not through 0 (revolute_joint_model.cpp:147-167)
    s = motion.edge_states(np.array([3.0]), np.array([-3.0]), 0.1, continuous=[1])
    assert len(s) == 3 and np.all(np.abs(s[:-1]) > 3.0)
    assert abs(motion.distance([3.0], [-3.0], continuous=[1]) - (2 * np.pi - 6.0)) < 1e-12
    # passive variable (factor 0) does not lengthen the edge
    assert motion.distance([0.0, 0.0], [1.0, 9.0], factor=[1.0, 0.0]) == 1.0


# ---- signed fields (propagate_negative_): TestSignedAddRemovePoints, test_distance_field.cpp:441-608 ------------------
def signed_small():
    return PropagationDistanceField(WIDTH, HEIGHT, DEPTH, RESOLUTION, 0, 0, 0, MAX_DIST, True)


def exact_d2(sources, max_d2):
    """Brute force: squared cell distance to the nearest True cell of `sources`, capped at max_d2."""
    from scipy import ndimage
    if not sources.any():
        return np.full(sources.shape, max_d2, np.int64)
    return np.minimum(np.rint(ndimage.distance_transform_edt(~sources) ** 2).astype(np.int64), max_d2)


def block_points(df):
    lo, hi = df.grid_to_world(1, 1, 1), df.grid_to_world(8, 8, 8)
    pts = []
    x = lo[0]
    while x <= hi[0]:  # the test's own accumulating loops (:463-473)
        y = lo[1]
        while y <= hi[1]:
            z = lo[2]
            while z <= hi[2]:
                pts.append((x, y, z))
                z += .1
            y += .1
        x += .1
    return pts


def test_signed_add_remove_points_equals_fresh_build():
    df = signed_small()
    pts = block_points(df)
    df.reset()
    df.add_points(pts)
    occ = df.d2() == 0
    # inside the block the negative field is the distance to the nearest free cell; free cells carry 0
    assert np.array_equal(df.neg_d2(), np.where(occ, exact_d2(~occ, df.max_distance_sq), 0))
    assert (df.neg_d2()[occ] > 0).all()  # checkDistanceField(do_negs): "Obstacle point has zero negative value"
    centre = tuple(df.grid_to_world(5, 5, 5))
    df.remove_points([centre])
    fresh = signed_small()
    fresh.add_points([p for p in pts if p != centre])
    # areDistanceFieldsDistancesEqual (:508) compares getDistance of every cell: both halves of the signed field
    assert np.array_equal(df.d2(), fresh.d2()) and np.array_equal(df.neg_d2(), fresh.neg_d2())
    assert df.d2()[5, 5, 5] == 1 and df.neg_d2()[5, 5, 5] == 0 and df.neg_d2()[5, 5, 4] == 1


def test_signed_sphere_shape_field_and_gradients():
    # :510-606: sphere r=.25 at (.5,.5,.5): cell (5,5,5) is deeper than one cell inside, interior distances are negative,
    # getDistanceGradient agrees with the cell value and points out of the obstacle
    df = signed_small()
    df.add_points(find_internal_points_sphere((0.5, 0.5, 0.5), 0.25, RESOLUTION))
    nd2, d2 = df.neg_d2(), df.d2()
    assert nd2[5, 5, 5] > 1
    assert np.array_equal(nd2 > 0, d2 == 0)
    n = df.num_cells
    inside = 0
    for x in range(1, n[0] - 1):
        for y in range(1, n[1] - 1):
            for z in range(1, n[2] - 1):
                if nd2[x, y, z] == 0:
                    continue
                inside += 1
                dist = df.get_cell_distance(x, y, z)
                assert dist < 0
                w = df.grid_to_world(x, y, z)
                dg, grad, inb = df.get_distance_gradient(*w)
                assert inb and abs(dist - dg) < 1e-4
                g2 = float(grad @ grad)
                if g2 < np.finfo(float).eps:
                    continue
                comp = w - grad / np.sqrt(g2) * dist
                c, ok = df.world_to_grid(*comp)
                assert ok and d2[c] >= 1
    assert inside > 0


def test_signed_field_vs_exact_complement_transform_on_solid_bodies():
    # two solid spheres: like propagatePositive, propagateNegative is an approximation of the exact (truncated) transform
    # of the complement that never under-estimates; the differing voxels are few (31 of 1 114 here)
    df = PropagationDistanceField(1.0, 1.0, 1.0, 0.05, 0, 0, 0, 0.3, True)
    df.add_points(find_internal_points_sphere((0.4, 0.4, 0.5), 0.3, 0.05))
    df.add_points(find_internal_points_sphere((0.7, 0.6, 0.5), 0.2, 0.05))
    occ = df.d2() == 0
    exact = np.where(occ, exact_d2(~occ, df.max_distance_sq), 0)
    n = df.neg_d2()
    assert (n >= exact).all() and np.array_equal(n > 0, occ)
    assert (n != exact).sum() <= 0.05 * occ.sum()
    assert n.max() == exact.max() == 36


# ---- bodies decomposed into interior lattice points (findInternalPointsConvex over geometric_shapes bodies) --------------------
def test_find_internal_points_body_restatement():
    from oracle.collision_model import find_internal_points_body
    T = np.eye(4)
    T[:3, 3] = (0.5, 0.5, 0.5)
    # a posed sphere in the world frame IS the construction test_big.df pins
    assert np.array_equal(find_internal_points_body(1, (0.5,), T, 0.02, "world"),
                          find_internal_points_sphere((0.5, 0.5, 0.5), 0.5, 0.02))
    # axis-aligned box: the lattice points inside are the lattice values within the half extents, per axis
    res = 0.05
    pts = find_internal_points_body(4, (0.4, 0.2, 0.3), T, res, "world")
    for k, half in enumerate((0.2, 0.1, 0.15)):
        assert np.all(np.abs(pts[:, k] - 0.5) <= half)
    assert abs(len(pts) * res ** 3 / (0.4 * 0.2 * 0.3) - 1.0) < 0.5
    # body-frame lattice + pose == world-frame lattice for a pure lattice-aligned translation (identity rotation, t = k * res)
    T2 = np.eye(4)
    T2[:3, 3] = (0.25, 0.5, 0.75)
    # (extents off the lattice: a point exactly on a face is decided by the rounding of the accumulated lattice values)
    a = find_internal_points_body(4, (0.43, 0.21, 0.33), T2, res, "world")
    b = find_internal_points_body(4, (0.43, 0.21, 0.33), T2, res, "body")
    assert len(a) == len(b) and np.allclose(np.sort(a, axis=0), np.sort(b, axis=0), atol=1e-12)
    # cylinder along z: radial and axial limits
    c = find_internal_points_body(2, (0.1, 0.5), T, 0.02, "world")
    assert np.all(np.abs(c[:, 2] - 0.5) <= 0.25) and np.all((c[:, 0] - 0.5) ** 2 + (c[:, 1] - 0.5) ** 2 <= 0.01 + 1e-15)
    assert abs(len(c) * 0.02 ** 3 / (np.pi * 0.01 * 0.5) - 1.0) < 0.15
    # rotating the cylinder by 90 degrees about x swaps the roles of y and z
    Rx = np.eye(4)
    Rx[:3, :3] = [[1, 0, 0], [0, 0, -1], [0, 1, 0]]
    Rx[:3, 3] = (0.5, 0.5, 0.5)
    cr = find_internal_points_body(2, (0.1, 0.5), Rx, 0.02, "world")
    assert np.all(np.abs(cr[:, 1] - 0.5) <= 0.25 + 1e-12) and np.abs(cr[:, 2] - 0.5).max() <= 0.1 + 1e-12
