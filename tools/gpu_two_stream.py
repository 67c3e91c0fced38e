"""Dev tool: does splitting the headline 100k-motion batch over several streams hide the kernels' tails?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
stride, js, m = a.shape[1], a.shape[1] - 7, len(a)
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
words = (m + 31) // 32
d_hit = torch.zeros(words, dtype=torch.int32, device=dev); d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev); d_unc = torch.zeros(words, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
main = torch.cuda.current_stream()
ref = None
for parts in (1, 2, 3, 4, 8):
    streams = [main] + [torch.cuda.Stream() for _ in range(parts - 1)]
    per = ((m + parts - 1) // parts + 31) // 32 * 32
    def run():
        start = torch.cuda.Event(); start.record(main)
        done = []
        for p in range(parts):
            lo, hi = p * per, min(m, (p + 1) * per)
            if lo >= hi: break
            s = streams[p]
            if p: s.wait_event(start)
            mg.capi.check_motions_device(scene, robot, d_a[lo:hi], d_b[lo:hi], hi - lo, stride, js, d_hit[lo // 32:], bench.MAX_STEP,
                                         d_toi[lo:hi], d_steps[lo:hi], d_unc[lo // 32:], stream=s.cuda_stream)
            if p:
                e = torch.cuda.Event(); e.record(s); done.append(e)
        for e in done: main.wait_event(e)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for k in range(20):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); run(); e1.record(main); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    h = d_hit.cpu().numpy().copy()
    if ref is None: ref = h
    print("parts=%d: %.3f ms (min %.3f) same=%s" % (parts, sum(ts) / len(ts), min(ts), np.array_equal(ref, h)), flush=True)
