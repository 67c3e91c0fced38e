"""Device-buffer calls recorded into a CUDA graph: they must neither block nor allocate once their scratch exists, and a replay
must give the results of a plain call.  Under capture the motion pipeline records its two rounds as a fork/join (the tail
expansion on a side stream behind the head traversal, launch_check_motions): this is the test that runs that path."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_captured_calls_replay_to_the_same_results(gpu):
    import torch
    mg = gpu
    tree = mg.scenes.make_tree(seed=9, trunk_triangles=4000, leaf_triangles=12000, n_fruit=16)
    scene = mg.Scene.from_mesh(tree.collision_mesh(True))
    model = mg.robots.create_procedural_robot_model(0.75, (0, 1))
    robot = model.to_device()
    n, m = 40000, 12000
    st = mg.robots.generate_uniform_random_states(model, mg.robots.Mt19937(3), n, 2.5, 3.5)
    dev = torch.device("cuda")
    d_st = torch.from_numpy(st).to(dev)
    d_a, d_b = d_st[:m].contiguous(), d_st[m:2 * m].contiguous()
    stride, js = st.shape[1], st.shape[1] - 7

    def buffers():
        return (torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev), torch.zeros((m + 31) // 32, dtype=torch.int32, device=dev),
                torch.full((m,), -1.0, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.int32, device=dev))

    def enqueue(bufs, stream):
        s_hit, m_hit, toi, steps = bufs
        mg.capi.check_states_device(scene, robot, d_st, n, stride, js, s_hit, stream=stream)
        mg.capi.check_motions_device(scene, robot, d_a, d_b, m, stride, js, m_hit, 0.1, toi, steps, stream=stream)

    plain = buffers()
    enqueue(plain, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cap = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(cap):  # the stream's scratch is allocated by its first call, before the capture
        enqueue(buffers(), cap.cuda_stream)
    torch.cuda.synchronize()
    replayed = buffers()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
        enqueue(replayed, torch.cuda.current_stream().cuda_stream)
    for b in replayed:  # nothing ran during the capture
        assert float(b.double().abs().sum()) in (0.0, float(m))
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for p, r in zip(plain, replayed):
        assert np.array_equal(p.cpu().numpy(), r.cpu().numpy(), equal_nan=True)
    hits = mg.capi.unpack_bits(plain[1].cpu().numpy().view(np.uint32), m)
    assert 0.05 < hits.mean() < 0.95
