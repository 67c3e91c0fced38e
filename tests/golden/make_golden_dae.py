#!/usr/bin/env python
"""Generates tests/golden/drone_dae_summary.json: what an assimp-style reading (tests/dae_reader.py) makes of the one real
COLLADA model that ships with the reference, test_robots/meshes/drone.dae (a Blender export; the `3d-models` tree set is a
git-lfs pointer and absent).  The model itself stays in /root/reference; only its summary is committed.

    python tests/golden/make_golden_dae.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import dae_reader  # noqa: E402

SRC = "/root/reference/test_robots/meshes/drone.dae"
v, t = dae_reader.read(SRC)
out = dict(generator="tests/golden/make_golden_dae.py (tests/dae_reader.py over the reference's test_robots/meshes/drone.dae)",
           file="test_robots/meshes/drone.dae", bytes=os.path.getsize(SRC), **dae_reader.summary(v, t))
with open(os.path.join(HERE, "drone_dae_summary.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
