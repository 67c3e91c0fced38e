"""Dev tool: work counters and kernel times of a state batch for the all-box robot and the capsule + polytope robot (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh, counters=True)
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
n = 10_000_000
lo, hi = tree.trunk_mesh.aabb()
h = max(hi[0] - lo[0], hi[1] - lo[1]) / 2 + 2.0
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
st[:, 0] = (u[:, 0] * 2 - 1) * h; st[:, 1] = (u[:, 1] * 2 - 1) * h; st[:, 2] = u[:, 2] * float(lo[2] + 1.0)
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7] = (u[:, 4] * 2 - 1) * (np.pi / 2); st[:, 8] = (u[:, 5] * 2 - 1) * (np.pi / 2)
d_hit = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
boxes = mg.robots.create_procedural_robot_model(1.0, (0,))
shaped = mg.robots.RobotModel()
base = shaped.insert_link("flying_base", (0.8, 0.8, 0.2)); arm = shaped.insert_link("stick")
s45 = float(np.sin(np.pi / 4))
kind = sys.argv[1] if len(sys.argv) > 1 else "both"
if kind in ("both", "capsule"): shaped.add_capsule(arm, 0.035, 0.93, (0.0, 0.0, 0.0, s45, 0.0, 0.0, s45))
else: shaped.shapes.append((arm, (0.05, 1.0, 0.05), mg.capi.IDENTITY_TF))
ee = shaped.insert_link("end_effector")
if kind in ("both", "convex"): shaped.add_convex_mesh(ee, [[-0.08, -0.1, -0.05], [0.08, -0.1, -0.05], [0.08, 0.1, -0.05], [-0.08, 0.1, -0.05], [0.0, 0.12, 0.09], [0.0, -0.06, 0.07]])
shaped.insert_joint(base, arm, (0, 0.4, 0, 0, 0, 0, 1), (0, -0.5, 0, 0, 0, 0, 1), mg.capi.JOINT_REVOLUTE, (1.0, 0.0, 0.0), -np.pi / 2, np.pi / 2)
shaped.insert_joint(arm, ee, (0, 0.5, 0, 0, 0, 0, 1), mg.capi.IDENTITY_TF, mg.capi.JOINT_REVOLUTE, (0.0, 0.0, 1.0), -np.pi / 2, np.pi / 2)
for name, model in (("boxes", boxes), ("shaped:" + kind, shaped)):
    robot = model.to_device()
    mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream); torch.cuda.synchronize()
    scene.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream); e1.record(); torch.cuda.synchronize()
    print(name, "%.3f ms" % e0.elapsed_time(e1), scene.counters(reset=True), flush=True)
