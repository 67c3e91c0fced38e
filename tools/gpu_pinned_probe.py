"""Dev tool: is the slow mode of the host-buffer call a property of WHERE a process's pinned input buffers landed?
Allocates several pinned input pairs, measures each one's H2D copy rate, then times the host-buffer motion call on the
fastest and on the slowest pair.  Run it several times in a row: the rates may differ from process to process."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import mgodpl_b200 as mg

tree = mg.scenes.make_tree(**bench.TREE)
scene = mg.Scene.from_mesh(tree.collision_mesh(True))
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
m = len(a)
res = mg.capi.motion_result_buffers(m)
d = torch.empty(a.shape, dtype=torch.float64, device="cuda")


def rate(t):
    ts = []
    for _ in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return t.numel() * 8 / np.median(ts[2:]) / 1e9


pairs, rates = [], []
for j in range(8):
    pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    pairs.append((pa, pb))
    rates.append((rate(pa), rate(pb)))
print("H2D GB/s per pinned pair:", " ".join(f"{x:.1f}/{y:.1f}" for x, y in rates), flush=True)


def e2e(pa, pb, n=160):
    na, nb = pa.numpy(), pb.numpy()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        mg.capi.check_motions(scene, robot, na, nb, 0.1, True, out=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts[100:]), min(ts)


order = np.argsort([min(r) for r in rates])
for tag, j in (("slowest pair", order[0]), ("fastest pair", order[-1]), ("slowest pair again", order[0])):
    med, mn = e2e(*pairs[j])
    print(f"e2e on the {tag} (#{j}, {rates[j][0]:.1f}/{rates[j][1]:.1f} GB/s): median {med:.3f} ms, min {mn:.3f} ms", flush=True)
print("rates afterwards:", " ".join(f"{rate(pa):.1f}/{rate(pb):.1f}" for pa, pb in pairs), flush=True)
