"""Dev tool: device timeline of the host-buffer motion call's chunks (MGODPL_TRACE_E2E: plain launches, events around every
chunk's input copy, kernels and result copy).  Prints the last traced calls."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import time

    import torch

    import bench
    import mgodpl_b200 as mg
    tree = mg.scenes.make_tree(**bench.TREE)
    scene = mg.Scene.from_mesh(tree.collision_mesh(True))
    model, a, b = bench.make_inputs(mg, 0)
    robot = model.to_device()
    pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    res = mg.capi.motion_result_buffers(len(a))
    for i in range(150):
        t0 = time.perf_counter()
        mg.capi.check_motions(scene, robot, pa.numpy(), pb.numpy(), 0.1, True, out=res)
        print(f"[host] call {i}: {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr)
    sys.exit(0)
for extra in ({}, {"MGODPL_E2E_CHUNKS": "3", "MGODPL_E2E_TAPER": "1"}):
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, MGODPL_TRACE_E2E="1", **extra),
                       capture_output=True, text=True, timeout=120)
    lines = p.stderr.splitlines()
    print(f"---- {extra or 'default layout'} (rc {p.returncode})")
    print("\n".join(lines[-18:]))
