"""Dev tool: config 5 (8-tree orchard, ~2 M triangles): 1 M motions, for ncu launch lists and counters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
orchard = mg.scenes.make_orchard(); mesh = orchard.collision_mesh(True)
counters = len(sys.argv) > 1
scene = mg.Scene.from_mesh(mesh, counters=counters)
model = mg.robots.create_procedural_robot_model(1.0, (0,)); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
mm = 1_000_000
gen = torch.Generator(device=dev).manual_seed(5)
u = torch.rand((2 * mm, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((2 * mm, 9), dtype=torch.float64, device=dev)
st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 10.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
a, b = st[:mm].contiguous(), st[mm:].contiguous()
d_h = torch.zeros((mm + 31) // 32, dtype=torch.int32, device=dev); d_t = torch.zeros(mm, dtype=torch.float64, device=dev)
d_s = torch.zeros(mm, dtype=torch.int32, device=dev)
for _ in range(2):
    mg.capi.check_motions_device(scene, robot, a, b, mm, 9, 2, d_h, 0.1, d_t, d_s, stream=stream)
torch.cuda.synchronize()
if counters: scene.counters(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    mg.capi.check_motions_device(scene, robot, a, b, mm, 9, 2, d_h, 0.1, d_t, d_s, stream=stream)
e1.record(); torch.cuda.synchronize()
print("orchard 1M motions: %.3f ms" % (e0.elapsed_time(e1) / 3), scene.info().n_nodes, "records")
if counters:
    c = scene.counters(); print({k: v / 3 for k, v in c.items()})
