"""Dev tool: roadmap shortest paths on a 100 k-vertex k-NN roadmap (for timing sweeps and ncu launch lists)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nn, k = 100_000, 10
model = mg.robots.create_procedural_robot_model(0.75, (0,))
nodes = mg.robots.generate_uniform_random_states(model, mg.robots.Mt19937(7), nn, 5.0, 10.0)
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
d_nodes = torch.from_numpy(nodes).to(dev)
d_idx = torch.zeros((nn, k), dtype=torch.int32, device=dev); d_nd = torch.zeros((nn, k), dtype=torch.float64, device=dev)
mg.capi.knn_device(d_nodes, nn, 9, 2, k, d_idx, d_nd, causal=True, stream=stream)
idx = d_idx.cpu().numpy(); src, col = np.nonzero(idx >= 0)
eu, ev, ew = idx[src, col].astype(np.int64), src.astype(np.int64), d_nd.cpu().numpy()[src, col]
both_u, both_v, both_w = np.r_[eu, ev], np.r_[ev, eu], np.r_[ew, ew]
order = np.argsort(both_u, kind="stable")
row = np.zeros(nn + 1, np.int64); np.add.at(row, both_u + 1, 1)
d_row = torch.from_numpy(np.cumsum(row).astype(np.uint32).view(np.int32)).to(dev)
d_col = torch.from_numpy(both_v[order].astype(np.uint32).view(np.int32)).to(dev)
d_w = torch.from_numpy(np.ascontiguousarray(both_w[order])).to(dev)
sources = np.random.default_rng(4).choice(nn, g, replace=False).astype(np.uint32)
d_src = torch.from_numpy(sources.view(np.int32)).to(dev)
d_dist = torch.empty((g, nn), dtype=torch.float64, device=dev); d_pred = torch.empty((g, nn), dtype=torch.int32, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sweeps = mg.capi.roadmap_shortest_paths_device(d_row, d_col, d_w, nn, d_src, g, d_dist, d_pred, stream=stream)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
print("group bytes %s: %d sources, %.1f ms, %.0f sources/s, %d sweeps enqueued" % (os.environ.get("MGODPL_SSSP_GROUP_BYTES", "default"), g, ms, g / ms * 1e3, sweeps))
