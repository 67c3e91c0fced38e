"""Dev tool: is the survivor-append atomic the bottleneck of k_cull_states?  Compare near vs far state batches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model = mg.robots.create_procedural_robot_model(1.0, (0,)); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
n = 10_000_000
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
st[:, 1] = (u[:, 1] * 2 - 1) * 5.0; st[:, 2] = u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
d_hit = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
for name, off in (("PRM box", 0.0), ("all far away (+1000 m)", 1000.0)):
    st[:, 0] = (u[:, 0] * 2 - 1) * 5.0 + off
    for _ in range(3):
        mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  ({n * 72 / ms / 1e6:.0f} GB/s of state bytes)")
