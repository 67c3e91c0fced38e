"""The rarely-taken paths must give the same answers as the fast ones: work-list overflow (inline expansion / inline
traversal) and the exact test without the fp32 pre-filter.  Each variant runs in a fresh process because the knobs are
read once from the environment."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np
import mgodpl_b200 as mg
tree = mg.scenes.make_tree(seed=4, trunk_triangles=3000, leaf_triangles=6000, n_fruit=12)
mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model = mg.robots.create_procedural_robot_model(0.75, (0, 1))
robot = model.to_device()
st = mg.robots.generate_uniform_random_states(model, mg.robots.Mt19937(5), 60000, 2.5, 3.5)
out = {}
out["states"] = hashlib.sha1(mg.check_states(scene, robot, st, packed=True).tobytes()).hexdigest()
m = mg.check_motions(scene, robot, st[:8000], st[8000:16000])
out["motions"] = hashlib.sha1(m["words"].tobytes() + np.nan_to_num(m["toi"], nan=-1).tobytes()).hexdigest()
# large enough (>= 32768) for the chunked host-buffer path, ragged so that the last chunk ends inside a result word
mc = mg.check_motions(scene, robot, st[:40003], st[19997:60000])
out["motions_chunked"] = hashlib.sha1(mc["words"].tobytes() + np.nan_to_num(mc["toi"], nan=-1).tobytes() + mc["n_steps"].tobytes()).hexdigest()
g = mg.sample_goals(scene, robot, tree.fruit_centers, 512, draws=mg.robots.goal_draws(model, mg.robots.Mt19937(42), 512))
out["goals"] = hashlib.sha1(g["words"].tobytes() + g["collisions"].tobytes()).hexdigest()
rng = np.random.default_rng(1)
q = rng.normal(size=(20000, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
boxes = np.concatenate([rng.uniform([-2, -2, 0], [2, 2, 3.2], (20000, 3)), q, rng.uniform(0.02, 0.6, (20000, 3))], axis=1)
out["boxes"] = hashlib.sha1(np.packbits(mg.check_boxes(scene, boxes)).tobytes()).hexdigest()
out["hit_rate"] = float(mg.check_states(scene, robot, st).mean())
print("RESULT " + json.dumps(out))
""" % ROOT


def run_variant(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.gpu
def test_overflow_and_no_prefilter_paths_agree_with_the_fast_path(gpu):
    base = run_variant({})
    assert 0.02 < base["hit_rate"] < 0.9
    variants = {
        "box queue of 64 items (inline traversal in stage A)": {"MGODPL_QUEUE_MAX_ITEMS": "64"},
        "survivor list of 128 entries (inline expansion in stage A1)": {"MGODPL_SURVIVOR_MAX_ENTRIES": "128"},
        "both lists tiny": {"MGODPL_QUEUE_MAX_ITEMS": "1", "MGODPL_SURVIVOR_MAX_ENTRIES": "1"},
        "no fp32 pre-filter (every leaf triangle through the fp64 test)": {"MGODPL_NO_PREFILTER": "1"},
        "refill / leaf batching extremes": {"MGODPL_REFILL_MIN": "1", "MGODPL_LEAF_MIN": "32"},
        "high-occupancy traversal variant": {"MGODPL_TRAVERSE_BLOCKS": "8", "MGODPL_REFILL_MIN": "32", "MGODPL_LEAF_MIN": "1"},
        "binary records only (no 4-wide fold, no root-record cull)": {"MGODPL_WIDE_BVH": "0"},
        "binary records with tiny lists": {"MGODPL_WIDE_BVH": "0", "MGODPL_QUEUE_MAX_ITEMS": "64", "MGODPL_SURVIVOR_MAX_ENTRIES": "128"},
        "no range trimming / deep trimming": {"MGODPL_TRIM": "0"},
        "fixed head of 4 steps": {"MGODPL_HEAD_STEPS": "4"},
        "no clearance grid (bounds / root-record culls only)": {"MGODPL_CLEARANCE_GRID": "0"},
        "coarse clearance grid (4096 cells)": {"MGODPL_CLEARANCE_CELLS": "4096"},
        "host-buffer call in one chunk": {"MGODPL_E2E_CHUNKS": "1"},
        "host-buffer call in eight steeply tapered chunks, plain launches": {"MGODPL_E2E_CHUNKS": "8", "MGODPL_E2E_TAPER": "0.3", "MGODPL_E2E_GRAPHS": "0"},
        "host-buffer call in three equal chunks on prioritised streams": {"MGODPL_E2E_CHUNKS": "3", "MGODPL_E2E_TAPER": "1", "MGODPL_E2E_PRIO": "1"},
    }
    for name, env in variants.items():
        got = run_variant(env)
        for key in ("states", "motions", "motions_chunked", "goals", "boxes"):
            assert got[key] == base[key], f"{name}: {key} differs from the default configuration"
