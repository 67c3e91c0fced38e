"""Dev tool: one variant (environment decides: MGODPL_PIPELINE, MGODPL_CLEARANCE_GRID, ...) of the bench workloads — timings,
device counters and result hashes, so that variants run as separate processes can be compared (run on the GPU box)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mgodpl_b200 as mg
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else "variant"
tree = mg.scenes.make_tree(**bench.TREE)
mesh = tree.collision_mesh(True)
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
m = len(a)
d_hit = torch.zeros((m + 31) // 32, dtype=torch.int32, device=dev)
d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
n = 10_000_000
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 5.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi)
st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
d_sh = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
f, s = len(tree.fruit_centers), 1000
d_t = torch.from_numpy(tree.fruit_centers.copy()).to(dev)
d_gh = torch.zeros(f * ((s + 31) // 32), dtype=torch.int32, device=dev)
d_cnt = torch.zeros(f, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
out = {"tag": tag}
for counters in (True, False):
    scene = mg.Scene.from_mesh(mesh, counters=counters)

    def motions():
        mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)

    def states():
        mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_sh, stream=stream)

    def goals():
        mg.capi.sample_goals_device(scene, robot, d_t, f, s, d_gh, d_collisions=d_cnt, stream=stream)
    for name, fn, bufs in (("motions100k", motions, (d_hit, d_toi)), ("states10M", states, (d_sh,)), ("goals500x1000", goals, (d_gh, d_cnt))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if counters:
            scene.counters(reset=True)
        ts = []
        for k in range(10):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        h = hashlib.sha1(b"".join(np.nan_to_num(x.cpu().numpy(), nan=-1.0).tobytes() for x in bufs)).hexdigest()[:12]
        msg = f"[{tag}] {name}: ms={np.mean(ts):.4f} (min {np.min(ts):.4f}) hash={h}"
        if counters:
            c = {x: v / 10 for x, v in scene.counters().items()}
            msg += (f" [counters on] states={c['states_checked']:.4g} boxes={c['boxes_traversed']:.4g} records={c['nodes_visited']:.4g} "
                    f"rec/box={c['nodes_visited'] / max(c['boxes_traversed'], 1):.2f} exact={c['leaf_tests']:.4g} maxrec={scene.counters()['max_nodes_per_box']}")
        else:
            out[name] = {"ms": float(np.mean(ts)), "min": float(np.min(ts)), "hash": h}
        print(msg, flush=True)
print("RESULT " + json.dumps(out), flush=True)
