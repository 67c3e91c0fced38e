"""CPU tests of the host-side pieces: the Python mirror of the reference's samplers, batch sharding (incl. a
world_size-2 gloo run), and the C ABI library's exported surface."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(mg):
    """include/mgodpl_b200.h is the contract: every function it declares must be exported by the built library."""
    L = mg.capi.lib()
    header = open(os.path.join(ROOT, "include", "mgodpl_b200.h")).read()
    declared = set(re.findall(r"\b(mgodpl_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(mg.capi.EXPORTS)


def test_no_cpu_fallback_without_device(mg):
    if mg.capi.lib().mgodpl_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mg.MgodplError) as e:
        mg.Scene([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])
    assert e.value.code == 6 and "no CPU fallback" in e.value.message


def test_argument_validation_happens_before_device_use(mg):
    """Unsupported shapes / bad indices are rejected with the reference's error vocabulary."""
    with pytest.raises(mg.MgodplError) as e:
        mg.Robot(2, [(0, (1, 1, 1), mg.capi.IDENTITY_TF, mg.capi.SHAPE_MESH)], [], 0)
    assert e.value.code == 2 and "Only boxes are implemented" in e.value.message
    with pytest.raises(mg.MgodplError) as e:
        mg.Robot(2, [], [], 5)
    assert e.value.code == 3 and "Link not found" in e.value.message
    with pytest.raises(mg.MgodplError) as e:
        mg.Robot(2, [], [(0, 1, mg.capi.IDENTITY_TF, mg.capi.IDENTITY_TF, 7, (1, 0, 0), 0, 0)], 0)
    assert e.value.code == 4 and "Unknown joint type" in e.value.message
    with pytest.raises(mg.MgodplError) as e:
        mg.Scene([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 3]])
    assert e.value.code == 1


def test_procedural_robot_matches_oracle_description(mg, oracle):
    for arm, jt in [(0.75, (0,)), (1.0, (1,)), (1.0, (0, 1)), (0.9, (1, 0, 1))]:
        m = mg.robots.create_procedural_robot_model(arm, jt)
        boxes, joints = oracle.Robot(arm, jt).describe()
        assert len(m.shapes) == len(boxes) and len(m.joints) == len(joints)
        for s, b in zip(m.shapes, boxes):
            assert s[0] == int(b[0]) and np.array_equal(np.array(s[1]), b[1:4]) and np.array_equal(np.array(s[2]), b[4:11])
        for j, o in zip(m.joints, joints):
            assert (j[0], j[1], j[4]) == (int(o[0]), int(o[1]), int(o[2]))
            assert np.array_equal(np.array(j[2]), o[3:10]) and np.array_equal(np.array(j[3]), o[10:17])
            if j[4] == mg.capi.JOINT_REVOLUTE:
                assert np.array_equal(np.array(j[5]), o[17:20]) and (j[6], j[7]) == (o[20], o[21])
        assert m.find_link_by_name("flying_base") == 0 and m.find_link_by_name("end_effector") == len(jt) + 1
        with pytest.raises(RuntimeError, match="Link not found"):
            m.find_link_by_name("nope")


def test_mt19937_sampler_matches_std_mt19937(mg, oracle):
    """The Python sampler reproduces std::mt19937 + libstdc++ uniform_real_distribution draw for draw."""
    for arm, jt in [(1.0, (0,)), (0.75, (0, 1))]:
        m = mg.robots.create_procedural_robot_model(arm, jt)
        got = mg.robots.generate_uniform_random_states(m, mg.robots.Mt19937(42), 500, 5.0, 10.0)
        want = oracle.gen_uniform_states(oracle.Robot(arm, jt), 500, seed=42)
        assert np.array_equal(got[:, :3], want[:, :3]) and np.array_equal(got[:, 7:], want[:, 7:])
        assert np.abs(got[:, 3:7] - want[:, 3:7]).max() < 1e-15  # sin/cos of the same yaw, different libm
        d = mg.robots.goal_draws(m, mg.robots.Mt19937(42), 300)
        assert np.array_equal(d, oracle.goal_draws(oracle.Robot(arm, jt), 300, seed=42))


def test_synthetic_tree_shape(mg):
    t = mg.scenes.make_tree(seed=42, trunk_triangles=20000, leaf_triangles=8000, n_fruit=50)
    lo, hi = t.trunk_mesh.aabb()
    assert abs(lo[2]) < 0.1 and 2.5 < hi[2] < 3.5 and max(hi[0] - lo[0], hi[1] - lo[1]) < 5.0
    assert t.leaves_mesh.n_triangles == 8000 and abs(t.trunk_mesh.n_triangles - 20000) < 6000
    assert t.fruit_centers.shape == (50, 3) and 1.0 < t.fruit_centers[:, 2].mean() < 3.0
    assert t.collision_mesh(True).n_triangles == t.trunk_mesh.n_triangles + 8000
    t2 = mg.scenes.make_tree(seed=42, trunk_triangles=20000, leaf_triangles=8000, n_fruit=50)
    assert np.array_equal(t.trunk_mesh.vertices, t2.trunk_mesh.vertices)  # deterministic in the seed
    orchard = mg.scenes.make_orchard(n_trees=4, trunk_triangles=800, leaf_triangles=400, n_fruit=5)
    lo, hi = orchard.trunk_mesh.aabb()
    assert lo[0] < -3.5 and hi[0] > 1.5  # makeSingleRowOrchard: trees at x = -4, -2, 0, 2


def test_shard_ranges_cover_and_align(mg):
    for n in [0, 1, 31, 32, 33, 1000, 100_000, 10_000_001]:
        for world in [1, 2, 3, 4, 8]:
            parts = [mg.sharding.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (lo, hi), (lo2, _) in zip(parts, parts[1:]):
                assert hi == lo2 and (lo % 32 == 0 or lo == n) and lo <= hi
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 64 or n < 32 * world  # one 32-block of slack + the ragged tail


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import mgodpl_b200 as mg
from oracle import oracle as o
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
tree = mg.scenes.make_tree(seed=3, trunk_triangles=1500, leaf_triangles=0, n_fruit=2)
mesh = tree.trunk_mesh
robot, scene = o.Robot(0.75, (0,)), o.Scene(mesh.vertices, mesh.triangles)
n = 1000
states = o.gen_uniform_states(robot, n, seed=11, h_range=2.0, v_range=3.0)
lo, hi = mg.sharding.shard_range(n, rank, 2)
local = scene.check_states(robot, states[lo:hi])                     # the oracle stands in for the GPU on CPU boxes
words = np.packbits(local, bitorder="little")
words = np.concatenate([words, np.zeros((-len(words)) % 4, np.uint8)]).view(np.uint32)
full = mg.sharding.gather_words(torch.from_numpy(words.view(np.int32).copy()), n, rank, 2)
got = mg.capi.unpack_bits(full.numpy().view(np.uint32), n)
want = scene.check_states(robot, states)
assert np.array_equal(got, want), (rank, int((got != want).sum()))
# roadmap shortest paths shard by source (a shard of 32 = one source block): each rank solves its sources, rows are gathered
rng = np.random.default_rng(5)
nv, edges, w = 300, rng.integers(0, 300, (900, 2)).astype(np.uint32), rng.uniform(0.1, 2.0, 900)
sources = rng.integers(0, nv, 70).astype(np.uint32)
lo, hi = mg.sharding.shard_range(len(sources), rank, 2)
d_local, _ = o.roadmap_shortest_paths(nv, edges, w, sources[lo:hi])
d_full = mg.sharding.gather_rows(torch.from_numpy(d_local), len(sources), rank, 2).numpy()
d_want, _ = o.roadmap_shortest_paths(nv, edges, w, sources)
assert (lo, hi) == ((0, 64), (64, 70))[rank] and np.array_equal(d_full, d_want)
# motions: shards balanced by the predicted state checks per motion (known before launch), ragged gather of words and toi rows
m = 900
a, b = states[:m // 2 * 2:2][:400], states[1:m // 2 * 2:2][:400]
costs = mg.sharding.predicted_motion_costs(a, b)
whole = scene.check_motions(robot, a, b)
assert np.array_equal(costs, whole["n_steps"] + 1)
bounds = mg.sharding.balanced_bounds(costs, 2)
assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == len(a) and bounds[0][1] % 32 == 0
lo, hi = bounds[rank]
mine = scene.check_motions(robot, a[lo:hi], b[lo:hi])
w = np.packbits(mine["hit"], bitorder="little")
w = np.concatenate([w, np.zeros((-len(w)) % 4, np.uint8)]).view(np.uint32)
words = mg.sharding.gather_ragged(torch.from_numpy(w.view(np.int32).copy()), bounds, rank).numpy().view(np.uint32)
toi = mg.sharding.gather_ragged(torch.from_numpy(mine["toi"]), bounds, rank, per_query_words=False).numpy()
assert np.array_equal(mg.capi.unpack_bits(words, len(a)), whole["hit"]) and np.array_equal(toi, whole["toi"], equal_nan=True)
shard_cost = [costs[l:h].sum() for l, h in bounds]
assert max(shard_cost) <= 1.2 * (sum(shard_cost) / 2)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", int(want.sum()))
"""


def test_balanced_bounds_cover_align_and_balance(mg):
    rng = np.random.default_rng(0)
    for n in [0, 1, 31, 33, 1000, 100_003]:
        costs = rng.integers(1, 300, n)
        for world in [1, 2, 3, 8]:
            b = mg.sharding.balanced_bounds(costs, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi and (lo % 32 == 0 or lo == n)
            if n >= 32 * world * 50:
                tot = [costs[lo:hi].sum() for lo, hi in b]
                assert max(tot) < 1.02 * costs.sum() / world
    # all the cost in one place: the other shards are simply empty, nothing is lost
    b = mg.sharding.balanced_bounds([0] * 64 + [5] + [0] * 63, 4)
    assert b[0][0] == 0 and b[-1][1] == 128 and sum(hi - lo for lo, hi in b) == 128


def test_two_rank_gloo_shard_and_gather(tmp_path):
    """N > 1 path on CPU: two gloo ranks each check their 32-aligned shard, the gathered words equal the 1-rank mask."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
        assert "ok" in out


def test_cpp_facade_host_logic_against_the_oracle(mg):
    """tests/cpp/test_host.cpp: the C++ facade's device-free parts (FK, distance, interpolation, state tools, goal projection,
    closure-taking shortcutting over the oracle's motion check, sharding arithmetic, roadmap walking, ABI argument errors)."""
    mg.capi.lib()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    p = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_host")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    m = re.search(r"(\d+) checks, 0 failed", p.stdout)
    assert m and int(m.group(1)) > 4000
