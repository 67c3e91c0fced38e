"""Dev tool: the headline step replayed from a CUDA graph, as one call vs as k concurrent part-batch calls on k streams inside the graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE)
scene = mg.Scene.from_mesh(tree.collision_mesh(True))
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
m = len(a)
d_hit = torch.zeros((m + 31) // 32, dtype=torch.int32, device=dev)
d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
ref = None
for parts in (1, 2, 3, 4):
    bounds = [((m * i // parts + 31) // 32 * 32 if i < parts else m) for i in range(parts + 1)]
    cap = torch.cuda.Stream(device=dev)
    sides = [torch.cuda.Stream(device=dev) for _ in range(parts - 1)]

    def body():
        cur = torch.cuda.current_stream()
        for i in range(parts):
            lo, hi = bounds[i], bounds[i + 1]
            st = cur if i == 0 else sides[i - 1]
            if i: st.wait_stream(cur)
            with torch.cuda.stream(st):
                mg.capi.check_motions_device(scene, robot, d_a[lo:], d_b[lo:], hi - lo, 9, 2, d_hit[lo // 32:], 0.1, d_toi[lo:], d_steps[lo:], stream=st.cuda_stream)
        for s in sides: cur.wait_stream(s)
    with torch.cuda.stream(cap):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
        body()
    ts = []
    for k in range(14):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    h = d_hit.cpu().numpy().copy(); t = d_toi.cpu().numpy().copy()
    if ref is None: ref = (h, t)
    same = np.array_equal(h, ref[0]) and np.array_equal(t, ref[1], equal_nan=True)
    print(f"parts={parts}: replay ms median {sorted(ts[2:])[6]:.4f} min {min(ts[2:]):.4f}  identical={same}", flush=True)
