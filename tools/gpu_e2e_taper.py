"""Dev tool: host-buffer motion call (100 k motions) under different chunk layouts — MGODPL_E2E_CHUNKS x MGODPL_E2E_TAPER x MGODPL_E2E_PRIO (arguments: chunks:taper[:prio] ...).
One subprocess per setting (the knobs are read once per process); every setting's result hash must equal the first one's."""
import hashlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch

    import bench
    import mgodpl_b200 as mg
    tree = mg.scenes.make_tree(**bench.TREE)
    scene = mg.Scene.from_mesh(tree.collision_mesh(True))
    model, a, b = bench.make_inputs(mg, 0)
    robot = model.to_device()
    m = len(a)
    pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    na, nb = pa.numpy(), pb.numpy()
    res = mg.capi.motion_result_buffers(m)
    ts = []
    for i in range(330):
        t0 = time.perf_counter()
        mg.capi.check_motions(scene, robot, na, nb, 0.1, True, out=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    h = hashlib.sha1(res["words"].tobytes() + res["n_steps"].tobytes() + res["toi"].tobytes()).hexdigest()[:12]
    # a ragged size as well (boundaries, partial last word)
    m2 = 77777
    res2 = mg.capi.motion_result_buffers(m2)
    mg.capi.check_motions(scene, robot, na[:m2], nb[:m2], 0.1, True, out=res2)
    h2 = hashlib.sha1(res2["words"].tobytes() + res2["n_steps"].tobytes() + res2["toi"].tobytes()).hexdigest()[:12]
    t = np.array(ts[130:])
    print(f"RESULT median {np.median(t):.4f} mean {t.mean():.4f} min {t.min():.4f} p90 {np.percentile(t, 90):.4f} ms | hash {h} {h2}", flush=True)
    sys.exit(0)

settings = [(3, 1.0), (3, 0.7), (3, 0.5), (4, 0.7), (4, 0.55), (5, 0.6), (6, 0.6), (4, 1.0), (3, 1.0)]
if len(sys.argv) > 1:
    settings = [tuple(float(x) for x in s.split(":")) for s in sys.argv[1:]]
first = None
for setting in settings:
    chunks, taper, prio = (tuple(setting) + (0,))[:3]
    env = dict(os.environ, MGODPL_E2E_CHUNKS=str(int(chunks)), MGODPL_E2E_TAPER=str(taper), MGODPL_E2E_PRIO=str(int(prio)))
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=120)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
    if not line:
        print(f"chunks {int(chunks)} taper {taper} prio {int(prio)}: FAILED rc {p.returncode}\n{p.stdout[-400:]}\n{p.stderr[-800:]}", flush=True)
        continue
    hashes = line[0].split("hash ")[1]
    first = first or hashes
    print(f"chunks {int(chunks)} taper {taper} prio {int(prio)}: {line[0][7:]} {'' if hashes == first else '  <-- RESULTS DIFFER'}", flush=True)
