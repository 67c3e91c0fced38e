"""Dev tool: only the 10 M state-check workload (for ncu launch lists)."""
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
st[:, 0] = (u[:, 0] * 2 - 1) * 5.0; st[:, 1] = (u[:, 1] * 2 - 1) * 5.0; st[:, 2] = u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7] = (u[:, 4] * 2 - 1) * (np.pi / 2); st[:, 8] = (u[:, 5] * 2 - 1) * (np.pi / 2)
d_hit = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
for _ in range(3):
    mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
e1.record(); torch.cuda.synchronize()
print("10M states: %.3f ms" % (e0.elapsed_time(e1) / 5))
