#!/usr/bin/env python
"""Generates tests/golden/gnat_knn.json by RUNNING THE REFERENCE'S OWN NearestNeighborsGNAT (oracle/_ref, built from
/root/reference/src by `make -C oracle ref`) the way add_and_connect_roadmap_node drives it (tsp_over_prm.cpp:33-66):
nearestK among the states inserted so far, then add.  Only possible in the build container; the committed JSON is what
every other machine checks the oracle's causal k-NN against.

    python tests/golden/make_golden_knn.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

assert o.have_ref(), "build oracle/_ref first: make -C oracle ref (needs /root/reference)"
out = {"generator": "tests/golden/make_golden_knn.py over oracle/_ref (the reference's NearestNeighborsGNAT.h + GreedyKCenters.h)", "cases": []}
for arm, jt, n, k, seed in [(0.75, (0,), 400, 10, 7), (1.0, (0, 1), 300, 5, 3), (0.5, (1,), 60, 16, 11)]:
    robot = o.Robot(arm, jt)
    states = o.gen_uniform_states(robot, n, seed=seed, h_range=5.0, v_range=10.0)      # the PRM sampler (tsp_over_prm.cpp:569-571)
    idx, dist = o.ref_gnat_causal_knn(states, k)
    out["cases"].append(dict(arm_length=arm, joint_types=list(jt), n=n, k=k, seed=seed, h_range=5.0, v_range=10.0,
                             gnat_idx=idx.tolist(), gnat_dist=[[float(x).hex() for x in row] for row in dist.tolist()]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnat_knn.json"), "w") as f:
    json.dump(out, f)
print({c["n"]: sum(len(set(r)) < len(r) for r in c["gnat_idx"]) for c in out["cases"]}, "rows with duplicated neighbours per case")
