"""Dev tool: 60 host-buffer motion calls back to back, every call's wall time (looking for the slow mode of the host link)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model, a, b = bench.make_inputs(mg, 0); robot = model.to_device()
m = len(a)
pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
na, nb = pa.numpy(), pb.numpy()
res = mg.capi.motion_result_buffers(m)
mode = sys.argv[2] if len(sys.argv) > 2 else "none"
if mode == "sleep": time.sleep(1.0)          # let whatever follows the pinning settle
if mode == "busy":                            # keep the GPU busy for a second instead
    x = torch.zeros(64 << 20, device="cuda"); t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0: x.add_(1.0); torch.cuda.synchronize()
if mode == "spin":                            # keep the host busy for a second
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0: pass
ts = []
for i in range(60):
    t0 = time.perf_counter()
    mg.capi.check_motions(scene, robot, na, nb, 0.1, True, out=res)
    ts.append((time.perf_counter() - t0) * 1e3)
    if len(sys.argv) > 1 and i % 10 == 9: time.sleep(float(sys.argv[1]))
print("e2e ms:", " ".join(f"{t:.2f}" for t in ts))
d = torch.empty_like(pa, device="cuda")
cp = []
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pa, non_blocking=True); torch.cuda.synchronize(); cp.append((time.perf_counter() - t0) * 1e3)
print("plain 7.2 MB H2D ms:", " ".join(f"{t:.3f}" for t in cp))
