#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.json by RUNNING THE REFERENCE'S OWN SOURCES (oracle/_ref, built from
/root/reference/src by `make -C oracle ref`).  Only possible in the build container; the GPU box and CI consume the
committed JSON.  Floats are stored as hex strings so the fixtures are bit-exact.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

assert o.have_ref(), "build oracle/_ref first: make -C oracle ref (needs /root/reference)"


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).reshape(-1)]


ROBOTS = [(1.0, (0,)), (0.75, (0,)), (0.5, (1,)), (1.0, (0, 0)), (0.75, (0, 1)), (1.0, (1, 0, 1))]
out = {"generator": "tests/golden/make_golden.py over oracle/_ref (reference sources, FCL/Eigen shim)", "robots": []}
for arm, jt in ROBOTS:
    rr = o.RefRobot(arm, jt)
    states = o.ref_gen_uniform_states(rr, 24, seed=42)            # generateUniformRandomState, state_tools.cpp:44-67
    fk = np.stack([rr.fk(s) for s in states])                     # forwardKinematics, RobotModel.cpp:17-90
    a, b = states[:12], states[12:]
    dist = [o.ref_distance(x, y) for x, y in zip(a, b)]           # equal_weights_distance, distance.cpp:17-25
    ts = [0.0, 0.125, 0.5, 0.9, 1.0]
    interp = np.stack([[o.ref_interpolate(x, y, t) for t in ts] for x, y in zip(a, b)])  # RobotState.cpp:14-24
    goals = o.ref_gen_goal_states(rr, [0.3, -0.2, 2.1], 16, seed=42)                      # goal_sampling.cpp:51-79
    goals_d = o.ref_gen_goal_states(rr, [0.3, -0.2, 2.1], 8, seed=7, distance_from_target=0.15)
    entry = dict(arm_length=arm, joint_types=list(jt), n_links=rr.n_links, n_joints=rr.n_joints,
                 states=hx(states), states_shape=list(states.shape), fk=hx(fk), fk_shape=list(fk.shape),
                 distance=hx(dist), interp_t=ts, interp=hx(interp), interp_shape=list(interp.shape),
                 goal_target=[0.3, -0.2, 2.1], goal_states=hx(goals), goal_shape=list(goals.shape),
                 goal_states_dist=hx(goals_d), goal_dist_shape=list(goals_d.shape))
    if len(jt) == 1:
        pts = np.array([[0.5, 0.1, 2.0], [-0.3, 0.4, 1.5]])
        vecs = np.array([[0.2, 0.3, 0.4], [-1.0, 0.5, -0.2]])
        entry["ee_vec_in"] = hx(np.concatenate([pts, vecs], axis=1))
        entry["ee_vec_out"] = hx(np.stack([o.ref_from_ee_and_vector(rr, p, v) for p, v in zip(pts, vecs)]))  # state_tools.cpp:14-42
    out["robots"].append(entry)

# One small scene: the reference's check_* loops (collision_detection.cpp:57-122) over the shim narrow phase.
sys.path.insert(0, ROOT)
import mgodpl_b200 as mg  # noqa: E402

tree = mg.scenes.make_tree(seed=5, trunk_triangles=3000, leaf_triangles=0, n_fruit=4)
mesh = tree.trunk_mesh
rr = o.RefRobot(0.75, (0,))
rs = o.RefScene(mesh.vertices, mesh.triangles)
st = o.ref_gen_uniform_states(rr, 400, seed=42, h_range=2.0, v_range=3.0)
hits = rs.check_states(rr, st)
mh, mt = rs.check_motions(rr, st[:100], st[100:200])
path = np.zeros((5, 9))
path[:, 6] = 1.0
path[:, 1] = [-10.0, -6.0, -3.0, 3.0, 10.0]   # trunk fly-through along y at z = 1 (test/planning/MyCollisionEnvTest.cpp:24-76)
path[:, 2] = 1.0
path_first = rs.check_path(rr, path)
out["scene"] = dict(tree=dict(seed=5, trunk_triangles=3000, leaf_triangles=0, n_fruit=4), robot=[0.75, [0]],
                    states_seed=42, h_range=2.0, v_range=3.0, n_states=400, state_hits="".join("1" if h else "0" for h in hits),
                    motion_hits="".join("1" if h else "0" for h in mh), motion_toi=hx(np.nan_to_num(mt, nan=-1.0)),
                    path=hx(path), path_shape=list(path.shape), path_first_colliding_segment=int(path_first),
                    mesh_checksum=float(mesh.vertices.sum()).hex(), n_triangles=int(mesh.n_triangles))
# Path shortcutting (local_optimization.cpp:23-174 driven as tsp_over_prm.cpp:594-612 does), same scene and robot.
short = []
for case, (first, seed, iters, mode) in enumerate([(200, 5, 30, o.SHORTCUT_GLOBALLY), (214, 6, 30, o.SHORTCUT_GLOBALLY),
                                                  (228, 7, 0, o.DELETE_EVERY_WAYPOINT), (242, 8, 25, o.SHORTCUT_LOCALLY),
                                                  (256, 9, 0, o.MIDPOINT_PULL)]):
    r = rs.shortcut_path(rr, st[first:first + 14], seed=seed, iterations=iters, mode=mode)
    short.append(dict(first_state=first, n_waypoints=14, seed=seed, iterations=iters, mode=mode, pull=0.5,
                      n_success=r["n_success"], n_checks=r["n_checks"], path=hx(r["path"]), path_shape=list(r["path"].shape)))
out["shortcutting"] = short
# a17 / a18: arm_vector_from_state (state_tools.cpp:69-75), generateUniformRandomArmVectorState and
# findGoalStateByUniformSampling (goal_sampling.cpp:81-124) on the same scene, one std::mt19937(seed + i) per target.
# (The three Gaussians of an arm vector are constructor arguments, evaluated in an unspecified order: these vectors hold
# what g++ — the compiler oracle/_ref is built with — produces; see oracle/mgodpl_oracle.hpp.)
targets = np.concatenate([tree.fruit_centers, np.array([[0.4, 0.3, 1.2], [-0.5, 0.2, 2.0], [0.1, -0.6, 2.6], [0.9, 0.9, 1.0]])])
av = o.ref_sample_arm_vector_states(rs, rr, targets, seed=11, max_attempts=40, ee_distance=0.05)
ug = o.ref_sample_uniform_goal_states(rs, rr, targets, 1, seed=5, max_attempts=12)
out["goal_states"] = dict(targets=hx(targets), n_targets=len(targets),
                          arm_vector=dict(seed=11, max_attempts=40, ee_distance=0.05, found=[int(x) for x in av["found"]],
                                          next_draw=[int(x) for x in av["next_draw"]], states=hx(np.nan_to_num(av["states"], nan=-7.0))),
                          uniform=dict(seed=5, max_attempts=12, found=[int(x) for x in ug["found"]], next_draw=[int(x) for x in ug["next_draw"]],
                                       states=hx(np.nan_to_num(ug["states"], nan=-7.0))),
                          arm_vector_of=hx(np.stack([o.ref_arm_vector(rr, x[:8]) for x in st[:16]])))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote reference_vectors.json:", len(out["robots"]), "robots;", int(hits.sum()), "/ 400 states collide;",
      int(mh.sum()), "/ 100 motions collide; path first =", path_first, "; shortcut successes",
      [c["n_success"] for c in short], "waypoints left", [c["path_shape"][0] for c in short])
