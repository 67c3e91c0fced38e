"""GPU edge cases and size-independent properties, all through the C ABI."""
import json
import os
import threading

import numpy as np
import pytest

from helpers import BAND, device_robot_from_oracle, parity_report
from test_oracle import GOLDEN, unhex, yaw_q

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(gpu, oracle):
    tree = gpu.scenes.make_tree(seed=4, trunk_triangles=3000, leaf_triangles=6000, n_fruit=12)
    mesh = tree.collision_mesh(True)
    orb = oracle.Robot(0.75, (0,))
    return dict(tree=tree, mesh=mesh, dscene=gpu.Scene.from_mesh(mesh, counters=True), oscene=oracle.Scene(mesh.vertices, mesh.triangles),
                orb=orb, drb=device_robot_from_oracle(gpu, orb))


def test_golden_reference_vectors_on_gpu(gpu, oracle):
    """The committed vectors were produced by the reference's own check_* loops (tests/golden/make_golden.py)."""
    with open(GOLDEN) as f:
        g = json.load(f)["scene"]
    tree = gpu.scenes.make_tree(**g["tree"])
    mesh = tree.trunk_mesh
    assert float(mesh.vertices.sum()).hex() == g["mesh_checksum"]
    orb = oracle.Robot(g["robot"][0], tuple(g["robot"][1]))
    drb = device_robot_from_oracle(gpu, orb)
    scene = gpu.Scene.from_mesh(mesh)
    st = oracle.gen_uniform_states(orb, g["n_states"], seed=g["states_seed"], h_range=g["h_range"], v_range=g["v_range"])
    _, clearance = oracle.Scene(mesh.vertices, mesh.triangles).check_states(orb, st, clearance=True)
    got = gpu.check_states(scene, drb, st)
    want = np.array([c == "1" for c in g["state_hits"]])
    assert parity_report(got, want, clearance)["n_mismatch_outside_band"] == 0
    res = gpu.check_motions(scene, drb, st[:100], st[100:200])
    want_m = np.array([c == "1" for c in g["motion_hits"]])
    assert np.array_equal(res["hit"], want_m)
    assert np.array_equal(np.nan_to_num(res["toi"], nan=-1.0), unhex(g["motion_toi"]))
    assert gpu.check_path(scene, drb, unhex(g["path"], g["path_shape"])) == g["path_first_colliding_segment"]


def test_empty_scene_and_empty_batches(gpu, env):
    empty = gpu.Scene(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    st = np.zeros((70, 9))
    st[:, 6] = 1
    assert not gpu.check_states(empty, env["drb"], st).any()
    assert not gpu.check_motions(empty, env["drb"], st[:35], st[35:] + 1.0)["hit"].any()
    assert gpu.check_states(env["dscene"], env["drb"], np.zeros((0, 9))).shape == (0,)
    assert gpu.check_motions(env["dscene"], env["drb"], np.zeros((0, 9)), np.zeros((0, 9)))["hit"].shape == (0,)
    assert gpu.check_path(env["dscene"], env["drb"], st[:1]) == -1
    assert gpu.sample_goals(env["dscene"], env["drb"], np.zeros((0, 3)), 16)["hit"].shape == (0, 16)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 257, 1000])
def test_ragged_batch_sizes(gpu, oracle, env, n):
    st = oracle.gen_uniform_states(env["orb"], n, seed=n, h_range=2.0, v_range=3.0)
    want, cl = env["oscene"].check_states(env["orb"], st, clearance=True)
    words = gpu.check_states(env["dscene"], env["drb"], st, packed=True)
    got = gpu.capi.unpack_bits(words, n)
    assert parity_report(got, want, cl)["n_mismatch_outside_band"] == 0
    assert len(words) == (n + 31) // 32
    if n % 32:
        assert words[-1] >> (n % 32) == 0  # tail bits are zero


@pytest.mark.parametrize("gap", [1e-3, 1e-5, 1e-7, 0.0, -1e-7, -1e-5, -1e-3])
def test_single_triangle_band_cases(gpu, oracle, gap):
    """Base box (0.8 x 0.8 x 0.2 at the state's translation) against one triangle parallel to its +x face."""
    x = 0.4 + gap
    mesh = gpu.scenes.single_triangle([x, -0.2, -0.05], [x, 0.3, -0.02], [x, 0.0, 0.08])
    scene = gpu.Scene.from_mesh(mesh)
    orb = oracle.Robot(0.75, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    state = np.array([[0, 0, 0, 0, 0, 0, 1, -1.2, 0.0]])  # arm folded away from the triangle
    got = gpu.check_states(scene, drb, state)[0]
    want, cl = oracle.Scene(mesh.vertices, mesh.triangles).check_states(orb, state, clearance=True)
    assert abs(cl[0] - gap) < 1e-12
    assert got == want[0] == (gap <= 0)
    boxes = np.array([[0, 0, 0, 0, 0, 0, 1, 0.8, 0.8, 0.2]])
    assert gpu.check_boxes(scene, boxes)[0] == (gap <= 0)


def test_step_count_boundary_matches_host_libm(gpu, oracle, env):
    """d / 0.1 == 77.0 exactly; plus a sweep of distances sitting on integer boundaries."""
    a = np.array([[0, 0, 5, 0, 0, 0, 1, 0.0, 0.0]])
    b = np.array([[6, 0, 5, *yaw_q(1.0), 0.5, 0.2]])
    assert gpu.check_motions(env["dscene"], env["drb"], a, b)["n_steps"][0] == 77
    k = np.arange(1, 400)
    a = np.zeros((len(k), 9))
    a[:, 6] = 1
    a[:, 2] = 6.0
    b = a.copy()
    b[:, 0] = k * 0.1  # distances k * 0.1: ceil(d / 0.1) is rounding-sensitive
    got = gpu.check_motions(env["dscene"], env["drb"], a, b)["n_steps"]
    want = np.array([oracle.motion_steps(oracle.distance(x, y)) for x, y in zip(a, b)])
    assert np.array_equal(got, want)


def test_degenerate_and_nan_inputs(gpu, env):
    inside = np.array([[0, 0, 1, 0, 0, 0, 1, 0.0, 0.0]])
    res = gpu.check_motions(env["dscene"], env["drb"], inside, inside)
    assert res["n_steps"][0] == 0 and res["hit"][0] and res["toi"][0] == 0.0  # SURVEY Q8 deviation: start state once
    bad = np.array([[np.nan, 0, 1, 0, 0, 0, 1, 0.0, 0.0], [0, 0, 1, np.nan, 0, 0, 1, 0.0, 0.0], [0, 0, 1, 0, 0, 0, 1, np.nan, 0.0],
                    [np.inf, 0, 1, 0, 0, 0, 1, 0.0, 0.0]])
    got = gpu.check_states(env["dscene"], env["drb"], bad)
    # non-finite base poses are reported free (and never hang); a NaN *joint* leaves the base box well defined, and the
    # base at (0,0,1) sits in the trunk — the reference's link loop also stops at that first colliding link.
    assert list(got) == [False, False, True, False]


def test_u64_index_entry_point(gpu, oracle, env):
    s64 = gpu.Scene(env["mesh"].vertices, env["mesh"].triangles.astype(np.uint64))
    st = oracle.gen_uniform_states(env["orb"], 500, seed=8, h_range=2.0, v_range=3.0)
    assert np.array_equal(gpu.check_states(s64, env["drb"], st), gpu.check_states(env["dscene"], env["drb"], st))


def test_device_variants_equal_host_variants(gpu, oracle, env):
    import torch
    st = oracle.gen_uniform_states(env["orb"], 4000, seed=21, h_range=2.5, v_range=3.5)
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    d = torch.from_numpy(st).to(dev)
    bits = torch.zeros((4000 + 31) // 32, dtype=torch.int32, device=dev)
    gpu.capi.check_states_device(env["dscene"], env["drb"], d, 4000, 9, 2, bits, stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(gpu.capi.unpack_bits(bits.cpu().numpy().view(np.uint32), 4000), gpu.check_states(env["dscene"], env["drb"], st))
    a, b = d[:2000].contiguous(), d[2000:].contiguous()
    mbits = torch.zeros((2000 + 31) // 32, dtype=torch.int32, device=dev)
    toi = torch.zeros(2000, dtype=torch.float64, device=dev)
    steps = torch.zeros(2000, dtype=torch.int32, device=dev)
    gpu.capi.check_motions_device(env["dscene"], env["drb"], a, b, 2000, 9, 2, mbits, 0.1, toi, steps, stream=stream)
    torch.cuda.synchronize()
    host = gpu.check_motions(env["dscene"], env["drb"], st[:2000], st[2000:])
    assert np.array_equal(gpu.capi.unpack_bits(mbits.cpu().numpy().view(np.uint32), 2000), host["hit"])
    assert np.array_equal(np.nan_to_num(toi.cpu().numpy(), nan=-1), np.nan_to_num(host["toi"], nan=-1))
    assert np.array_equal(steps.cpu().numpy(), host["n_steps"])


def test_sharded_batches_give_identical_bitmask(gpu, oracle, env):
    """SURVEY §8e determinism check: the same batch split 1/2/4/8 ways concatenates to the same words."""
    n = 20_011
    st = oracle.gen_uniform_states(env["orb"], n, seed=5, h_range=2.5, v_range=3.5)
    whole = gpu.check_states(env["dscene"], env["drb"], st, packed=True)
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = gpu.sharding.shard_range(n, r, world)
            parts.append(gpu.check_states(env["dscene"], env["drb"], st[lo:hi], packed=True))
        assert np.array_equal(gpu.sharding.concat_words(parts, n), whole)


def test_leaf_size_and_refinement_do_not_change_results(gpu, oracle, env):
    st = oracle.gen_uniform_states(env["orb"], 6000, seed=6, h_range=2.0, v_range=3.2)
    base = gpu.check_states(env["dscene"], env["drb"], st)
    for leaf in (1, 2, 4, 8, 16):
        s = gpu.Scene.from_mesh(env["mesh"], max_leaf_size=leaf)
        assert s.info().max_leaf_size <= leaf
        assert np.array_equal(gpu.check_states(s, env["drb"], st), base)
    s = gpu.Scene.from_mesh(env["mesh"], refine=False)
    assert np.array_equal(gpu.check_states(s, env["drb"], st), base)


def test_concurrent_host_threads_share_one_scene(gpu, oracle, env):
    """The reference calls this path from TBB workers on one shared tree (approach_planning.cpp:479-513)."""
    batches = [oracle.gen_uniform_states(env["orb"], 3000, seed=100 + i, h_range=2.0, v_range=3.0) for i in range(6)]
    want = [gpu.check_states(env["dscene"], env["drb"], b) for b in batches]
    got = [None] * len(batches)

    def work(i):
        for _ in range(5):
            got[i] = gpu.check_states(env["dscene"], env["drb"], batches[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    [t.start() for t in th]
    [t.join() for t in th]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_counters_and_scene_info(gpu, oracle, env):
    info = env["dscene"].info()
    assert info.n_triangles == env["mesh"].n_triangles and info.n_nodes == info.n_leaves - 1 > 0 and info.sah_cost > 0
    env["dscene"].counters(reset=True)
    st = oracle.gen_uniform_states(env["orb"], 2048, seed=1, h_range=2.0, v_range=3.0)
    gpu.check_states(env["dscene"], env["drb"], st)
    c = env["dscene"].counters()
    assert c["states_checked"] == 2048 and c["nodes_visited"] > 0 and c["boxes_traversed"] <= 2 * 2048
    gpu.check_motions(env["dscene"], env["drb"], st[:512], st[512:1024])
    assert env["dscene"].counters()["motions_checked"] == 512


def test_motion_symmetry_property_at_scale(gpu, oracle, env):
    """check(a, b) == check(b, a) on 20 000 motions (src_old/test/collision_checking_test.cpp:30-57); the two directions
    sample mirrored parameters, so disagreements need a grazing contact and must stay a vanishing fraction."""
    m = gpu.robots.create_procedural_robot_model(0.75, (0,))
    s = gpu.robots.generate_uniform_random_states(m, gpu.robots.Mt19937(3), 40_000, 3.0, 4.0)
    f = gpu.check_motions(env["dscene"], env["drb"], s[:20_000], s[20_000:])
    r = gpu.check_motions(env["dscene"], env["drb"], s[20_000:], s[:20_000])
    assert np.array_equal(f["n_steps"], r["n_steps"])
    assert (f["hit"] != r["hit"]).mean() < 1e-3
    both = f["hit"] & r["hit"]
    # the first hit seen from a is the last hit seen from b: toi_f <= 1 - toi_r' for the *last* hit; weaker but exact:
    assert (f["toi"][both] >= 0).all() and (f["toi"][both] <= 1).all()


def test_engineered_collisions_are_found(gpu, oracle, env):
    """The reference's legacy engineered_collision test (test/planning/MyCollisionEnvTest.cpp:149-219), for the sampled checker:
    move a motion so that, at one of its sampled steps, a point inside a random link box lands on a random point of a
    random triangle.  The motion must be reported colliding, no later than that step."""
    rng = np.random.default_rng(17)
    orb, drb, mesh = env["orb"], env["drb"], env["mesh"]
    boxes, _ = orb.describe()
    states = oracle.gen_uniform_states(orb, 400, seed=23, h_range=3.0, v_range=4.0)
    a, b = states[:200].copy(), states[200:].copy()
    t_contact = np.zeros(200)
    for i in range(200):
        n = oracle.motion_steps(oracle.distance(a[i], b[i]))
        step = int(rng.integers(0, n + 1))
        t = step / n
        mid = oracle.interpolate(a[i], b[i], t)
        bx = boxes[rng.integers(0, len(boxes))]
        link_tf = orb.fk(mid)[int(bx[0])]
        local = (rng.uniform(-0.4, 0.4, 3)) * bx[1:4]                      # strictly inside the box
        p_box = oracle.tf_then(oracle.tf_then(link_tf, bx[4:11]), np.r_[local, 0, 0, 0, 1])[:3]
        tri = mesh.vertices[mesh.triangles[rng.integers(0, len(mesh.triangles))]]
        u, v = rng.uniform(0, 1, 2)
        if u + v > 1:
            u, v = 1 - u, 1 - v
        shift = tri[0] + u * (tri[1] - tri[0]) + v * (tri[2] - tri[0]) - p_box
        a[i, :3] += shift
        b[i, :3] += shift
        t_contact[i] = t
    res = gpu.check_motions(env["dscene"], drb, a, b)
    assert res["hit"].all()
    assert (res["toi"] <= t_contact + 1e-12).all()
    want = env["oscene"].check_motions(orb, a, b)
    assert np.array_equal(res["hit"], want["hit"]) and np.allclose(res["toi"], want["toi"], rtol=0, atol=0)


def test_goal_sampling_device_rng_and_distance(gpu, oracle, env):
    t = env["tree"].fruit_centers[:8]
    a = gpu.sample_goals(env["dscene"], env["drb"], t, 500, seed=1, want_states=True)
    b = gpu.sample_goals(env["dscene"], env["drb"], t, 500, seed=1, want_states=True)
    c = gpu.sample_goals(env["dscene"], env["drb"], t, 500, seed=2, want_states=True)
    assert np.array_equal(a["states"], b["states"]) and not np.array_equal(a["states"], c["states"])
    assert np.array_equal(a["collisions"], a["hit"].sum(axis=1))
    # every sampled state puts the end effector on its target (genGoalStateUniform's contract)
    st = a["states"].reshape(-1, 8)
    ee = np.stack([env["orb"].fk(s)[2, :3] for s in st[:200]])
    assert np.abs(ee - np.repeat(t, 500, axis=0)[:200]).max() < 1e-12
    # the sampled states' collision flags agree with the oracle
    want, cl = env["oscene"].check_states(env["orb"], st, clearance=True, threads=8)
    assert parity_report(a["hit"].reshape(-1), want, cl)["n_mismatch_outside_band"] == 0
    # distance_from_target > 0 backs the end effector off along the arm (goal_sampling.cpp:71-76)
    draws = oracle.goal_draws(env["orb"], 64, seed=9)
    d = gpu.sample_goals(env["dscene"], env["drb"], t[:2], 64, draws=draws, distance_from_target=0.2, want_states=True)
    want_states = np.stack([oracle.gen_goal_states(env["orb"], x, 64, seed=9, distance_from_target=0.2) for x in t[:2]])
    assert np.abs(d["states"] - want_states).max() < 1e-12


@pytest.mark.parametrize("causal", [False, True])
def test_knn_matches_exhaustive_oracle(gpu, oracle, env, causal):
    """§8f rank 1: exact k-NN under equal_weights_distance (the GNAT query before PRM edge validation)."""
    st = oracle.gen_uniform_states(env["orb"], 700, seed=12)
    idx, dist = gpu.capi.knn(st, 10, causal=causal)
    widx, wdist = oracle.knn(st, 10, causal=causal)
    assert np.allclose(np.nan_to_num(dist, nan=-1), np.nan_to_num(wdist, nan=-1), rtol=0, atol=1e-12)
    same = idx == widx
    # the device libm may order exact ties / 1-ulp near-ties differently; everything else must match
    assert same.mean() > 0.999
    bad = ~same
    assert np.abs(dist[bad] - wdist[bad]).max(initial=0) < 1e-12
    if causal:
        assert (idx[0] == -1).all() and (idx[5, :5] >= 0).all() and (idx[5, 5:] == -1).all()
        assert (idx < np.arange(700)[:, None]).all()


def test_knn_against_the_references_own_gnat(gpu, oracle):
    """tests/golden/gnat_knn.json: what the reference's NearestNeighborsGNAT returned, driven as its PRM builder drives it
    (tests/test_oracle.py explains the duplicates it lists): with duplicates removed, every row is the leading part of the
    library's causal k-NN row, in the same order."""
    with open(os.path.join(os.path.dirname(GOLDEN), "gnat_knn.json")) as f:
        case = json.load(f)["cases"][0]            # the default 0.75 m arm, 400 nodes from the PRM sampler, k = 10
    robot = oracle.Robot(case["arm_length"], tuple(case["joint_types"]))
    st = oracle.gen_uniform_states(robot, case["n"], seed=case["seed"], h_range=case["h_range"], v_range=case["v_range"])
    idx, _ = gpu.capi.knn(st, case["k"], causal=True)
    agree = 0
    for i, row in enumerate(case["gnat_idx"]):
        distinct = list(dict.fromkeys(x for x in row if x >= 0))
        agree += distinct == [x for x in idx[i].tolist() if x >= 0][:len(distinct)]
    assert agree >= case["n"] - 2                  # (a 1-ulp near-tie may be ordered differently by the device libm)


def test_knn_query_against_a_separate_database(gpu, oracle, env):
    """Goal samples / the start state look up their nearest *infrastructure* nodes without being part of the index
    (tsp_over_prm.cpp:297-344): exhaustive search with the oracle's equal_weights_distance."""
    orb = env["orb"]
    db = oracle.gen_uniform_states(orb, 300, seed=5)
    qs = oracle.gen_uniform_states(orb, 40, seed=6)
    idx, dist = gpu.capi.knn_query(db, qs, 7)
    for i, q in enumerate(qs):
        d = np.array([oracle.distance(q, x) for x in db])
        order = np.argsort(d, kind="stable")[:7]
        assert np.array_equal(idx[i], order) and np.allclose(dist[i], d[order], rtol=0, atol=1e-12)  # device acos: last ulp
    # fewer database rows than k: padded with -1 / NaN; an empty database gives only padding
    idx, dist = gpu.capi.knn_query(db[:3], qs[:2], 5)
    assert (idx[:, 3:] == -1).all() and np.isnan(dist[:, 3:]).all() and (idx[:, :3] >= 0).all()
    idx, _ = gpu.capi.knn_query(db[:0], qs[:2], 2)
    assert (idx == -1).all()


def test_prm_edge_validation_pipeline(gpu, oracle, env):
    """BASELINE config 4 in miniature: sample nodes -> state filter -> causal k-NN -> one batch of edge motions."""
    st = oracle.gen_uniform_states(env["orb"], 1500, seed=77, h_range=3.0, v_range=4.0)
    free = ~gpu.check_states(env["dscene"], env["drb"], st)
    nodes = st[free]
    idx, dist = gpu.capi.knn(nodes, 5, causal=True)
    src, dst = np.nonzero(idx >= 0)
    a, b = nodes[idx[src, dst]], nodes[src]  # motion_collides(neighbour, state), tsp_over_prm.cpp:51-53
    got = gpu.check_motions(env["dscene"], env["drb"], a, b)
    want = env["oscene"].check_motions(env["orb"], a, b, band=True, threads=8)
    ok = want["min_abs_clearance"] > BAND
    assert np.array_equal(got["hit"][ok], want["hit"][ok]) and np.array_equal(got["n_steps"], want["n_steps"])
    assert 0.05 < want["hit"].mean() < 0.95
