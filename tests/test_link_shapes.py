"""Capsule and convex-polytope links (BASELINE north_star (3)).  The reference collides boxes only
(collision_detection.cpp:21-22,49-51), so these shapes have no reference behaviour to pin: the oracle's definitions are
checked here against brute-force sampling and against the box test they must reduce to, and the CUDA path is checked against
the oracle (parity for these two shapes is "against the oracle only"; DESIGN.md says so)."""
import numpy as np
import pytest

from helpers import BAND, device_robot_from_oracle, parity_report


def _quat(axis, angle):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    return [*(axis * np.sin(angle / 2)), np.cos(angle / 2)]


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _tri_points(tri, n=48):
    u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n))
    keep = u + v <= 1.0 + 1e-12
    u, v = u[keep], v[keep]
    return tri[0] + u[:, None] * (tri[1] - tri[0]) + v[:, None] * (tri[2] - tri[0])


def test_oracle_capsule_clearance_matches_sampling(oracle):
    rng = np.random.default_rng(11)
    n_hit = 0
    for _ in range(150):
        radius, length = rng.uniform(0.02, 0.3), rng.uniform(0.0, 1.2)
        tf = [*rng.uniform(-0.5, 0.5, 3), *_quat(rng.normal(size=3), rng.uniform(-3, 3))]
        tri = rng.uniform(-0.9, 0.9, (3, 3))
        hit, cl = oracle.shape_tri(1, tf, tri, radius=radius, length=length)
        R, c = _rot(tf[3:]), np.array(tf[:3])
        seg = c + np.linspace(-0.5, 0.5, 400)[:, None] * (R @ np.array([0, 0, length]))
        pts = _tri_points(tri)
        d = np.sqrt(((seg[:, None, :] - pts[None, :, :]) ** 2).sum(-1)).min()
        # the sampled distance can only overestimate, by at most the sample spacing
        spacing = max(np.linalg.norm(tri[1] - tri[0]), np.linalg.norm(tri[2] - tri[0])) / 47 + length / 399
        assert cl <= d - radius + 1e-12
        assert cl >= d - radius - spacing
        assert hit == (cl <= 0.0)
        n_hit += hit
    assert 20 < n_hit < 130


def test_oracle_convex_reduces_to_the_box_test(oracle):
    """The hull of a box's eight corners must collide exactly like the box (same axes: same normalised gap)."""
    rng = np.random.default_rng(12)
    n_hit = 0
    for _ in range(300):
        size = rng.uniform(0.05, 0.8, 3)
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * size / 2
        tf = [*rng.uniform(-0.4, 0.4, 3), *_quat(rng.normal(size=3), rng.uniform(-3, 3))]
        tri = rng.uniform(-0.8, 0.8, (3, 3))
        hit_c, gap_c = oracle.shape_tri(2, tf, tri, verts=corners)
        hit_b, cl_b = oracle.box_tri(tf, size, tri)
        assert hit_c == hit_b
        if hit_b:
            assert abs(gap_c - cl_b) < 1e-12  # both are the SAT penetration depth
        else:
            assert 0.0 < gap_c <= cl_b + 1e-12  # the SAT gap is a lower bound of the distance
        n_hit += hit_b
    assert 50 < n_hit < 250


def test_oracle_convex_against_point_sampling(oracle):
    """Random polytopes: a sampled triangle point inside the hull forces a hit; a positive gap means no sampled point is inside."""
    rng = np.random.default_rng(13)
    from scipy.spatial import ConvexHull
    n_hit = 0
    for _ in range(120):
        verts = rng.normal(size=(rng.integers(5, 13), 3)) * rng.uniform(0.1, 0.4)
        hull = ConvexHull(verts)
        tf = [*rng.uniform(-0.2, 0.2, 3), *_quat(rng.normal(size=3), rng.uniform(-3, 3))]
        tri = rng.uniform(-0.7, 0.7, (3, 3))
        hit, gap = oracle.shape_tri(2, tf, tri, verts=verts)
        local = (_tri_points(tri, 64) - np.array(tf[:3])) @ _rot(tf[3:])  # world -> shape frame
        inside = (local @ hull.equations[:, :3].T + hull.equations[:, 3] <= 1e-12).all(axis=1).any()
        if inside:
            assert hit
        if not hit:
            assert gap > 0 and not inside
        n_hit += hit
    assert 15 < n_hit < 110


def _arm_with_shapes(oracle):
    """A drone whose arm is a capsule and whose end effector is a convex wedge, plus a capsule with its own transform."""
    r = oracle.Robot(None)
    base = r.add_link("flying_base")
    r.add_box(base, (0.8, 0.8, 0.2))
    arm = r.add_link("stick")
    r.add_capsule(arm, 0.04, 0.7, (0, 0, 0, *_quat((1, 0, 0), np.pi / 2)))  # axis along the link's y, like the reference's stick
    ee = r.add_link("end_effector")
    wedge = np.array([[-0.08, -0.1, -0.05], [0.08, -0.1, -0.05], [0.08, 0.1, -0.05], [-0.08, 0.1, -0.05], [0.0, 0.12, 0.09], [0.0, -0.06, 0.07]])
    r.add_convex(ee, wedge, (0.0, 0.05, 0.0, *_quat((0, 0, 1), 0.3)))
    r.add_capsule(ee, 0.02, 0.25, (0.05, 0.1, 0.02, *_quat((0, 1, 0), 0.7)))
    r.add_joint(base, arm, (0, 0.4, 0, 0, 0, 0, 1), (0, -0.4, 0, 0, 0, 0, 1), oracle.REVOLUTE, (1, 0, 0), -1.5, 1.5)
    r.add_joint(arm, ee, (0, 0.4, 0, 0, 0, 0, 1), (0, 0, 0, 0, 0, 0, 1), oracle.REVOLUTE, (0, 0, 1), -1.0, 1.0)
    return r


@pytest.fixture(scope="module")
def shape_scene(gpu, oracle):
    tree = gpu.scenes.make_tree(seed=5, trunk_triangles=3000, leaf_triangles=9000, n_fruit=20)
    mesh = tree.collision_mesh(True)
    return gpu.Scene.from_mesh(mesh), oracle.Scene(mesh.vertices, mesh.triangles)


@pytest.mark.gpu
def test_capsule_and_convex_links_match_the_oracle(gpu, oracle, shape_scene):
    dscene, oscene = shape_scene
    orb = _arm_with_shapes(oracle)
    drb = device_robot_from_oracle(gpu, orb)
    assert drb.n_variables == orb.n_variables == 2
    rng = np.random.default_rng(21)
    n = 20000
    yaw = rng.uniform(-np.pi, np.pi, n)
    states = np.concatenate([rng.uniform([-2.5, -2.5, 0.0], [2.5, 2.5, 3.5], (n, 3)), np.zeros((n, 2)), np.sin(yaw / 2)[:, None], np.cos(yaw / 2)[:, None],
                             rng.uniform(-1.4, 1.4, (n, 1)), rng.uniform(-1.0, 1.0, (n, 1))], axis=1)
    want, cl = oscene.check_states(orb, states, clearance=True, threads=8)
    got = gpu.check_states(dscene, drb, states)
    rep = parity_report(got, want, cl)
    assert rep["n_mismatch_outside_band"] == 0, rep
    assert 0.05 < want.mean() < 0.9
    # the shapes are not their bounding boxes: the same robot with boxes of the bounding sizes hits strictly more often
    a, b = states[:4000], states[4000:8000]
    wm = oscene.check_motions(orb, a, b, band=True, threads=8)
    gm = gpu.check_motions(dscene, drb, a, b)
    assert np.array_equal(gm["n_steps"], wm["n_steps"])
    ok = wm["min_abs_clearance"] > BAND
    assert np.array_equal(gm["hit"][ok], wm["hit"][ok])
    both = gm["hit"] & wm["hit"] & ok
    assert np.array_equal(gm["toi"][both], wm["toi"][both])
    assert 0.1 < wm["hit"].mean() < 0.95


@pytest.mark.gpu
def test_shapes_are_tighter_than_their_bounding_boxes(gpu, oracle, shape_scene):
    """A capsule link must not be answered with its bounding box: compare with a box robot of the bounding sizes."""
    dscene, oscene = shape_scene
    cap = oracle.Robot(None)
    base = cap.add_link("flying_base")
    cap.add_capsule(base, 0.25, 0.5)
    box = oracle.Robot(None)
    base = box.add_link("flying_base")
    box.add_box(base, (0.5, 0.5, 1.0))
    rng = np.random.default_rng(22)
    n = 20000
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    states = np.concatenate([rng.uniform([-2, -2, 0.2], [2, 2, 3.2], (n, 3)), q], axis=1)
    g_cap = gpu.check_states(dscene, device_robot_from_oracle(gpu, cap), states)
    g_box = gpu.check_states(dscene, device_robot_from_oracle(gpu, box), states)
    w_cap, cl = oscene.check_states(cap, states, clearance=True, threads=8)
    assert parity_report(g_cap, w_cap, cl)["n_mismatch_outside_band"] == 0
    assert not (g_cap & ~g_box).any()          # inside the bounding box, always
    assert (g_box & ~g_cap).sum() > 50          # and strictly tighter


@pytest.mark.gpu
def test_shape_argument_errors(gpu):
    from mgodpl_b200 import capi
    ident = (0, 0, 0, 0, 0, 0, 1)
    flat = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0.3, 0.3, 0]], float)
    with pytest.raises(capi.MgodplError) as e:
        capi.Robot(1, [(0, (0, 0, 0), ident, capi.SHAPE_CONVEX_MESH, flat)], [], 0)
    assert e.value.code == 2  # UNSUPPORTED_SHAPE: no volume
    with pytest.raises(capi.MgodplError) as e:
        capi.Robot(1, [(0, (0, 0, 0), ident, capi.SHAPE_CONVEX_MESH)], [], 0)  # no vertices
    assert e.value.code == 1
    with pytest.raises(capi.MgodplError) as e:
        capi.Robot(1, [(0, (-1.0, 0, 1.0), ident, capi.SHAPE_CAPSULE)], [], 0)
    assert e.value.code == 1
    with pytest.raises(capi.MgodplError) as e:
        capi.Robot(1, [(0, (1, 1, 1), ident, capi.SHAPE_MESH)], [], 0)  # the reference's own Mesh variant stays unsupported
    assert e.value.code == 2
