"""Dev tool: k-NN timings (causal / all-pairs) at several sizes (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
model = mg.robots.create_procedural_robot_model(0.75, (0,))
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [25_000, 50_000, 100_000]
for n in sizes:
    nodes = mg.robots.generate_uniform_random_states(model, mg.robots.Mt19937(7), n, 5.0, 10.0)
    d_nodes = torch.from_numpy(nodes).to(dev)
    for k in (10,):
        d_idx = torch.zeros((n, k), dtype=torch.int32, device=dev); d_nd = torch.zeros((n, k), dtype=torch.float64, device=dev)
        for causal in (True, False):
            for _ in range(2):
                mg.capi.knn_device(d_nodes, n, 9, 2, k, d_idx, d_nd, causal=causal, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                mg.capi.knn_device(d_nodes, n, 9, 2, k, d_idx, d_nd, causal=causal, stream=stream)
            e1.record(); torch.cuda.synchronize()
            print(f"knn n={n} k={k} causal={causal}: {e0.elapsed_time(e1) / 3:.3f} ms  mean kth dist {float(d_nd[n // 2:, -1].mean()):.3f}", flush=True)
