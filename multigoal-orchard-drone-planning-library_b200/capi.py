"""ctypes binding of ``libmgodpl_b200.so`` (the C ABI declared in ``include/mgodpl_b200.h``).

Plumbing only: every computation on the collision-query path happens inside the CUDA library.  There is no CPU
fallback — if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmgodpl_b200.so")

MAX_LINKS, MAX_BOXES, MAX_JOINT_VALUES = 12, 12, 16
SCENE_COUNTERS, SCENE_NO_REFINE, SCENE_NO_CLEARANCE_GRID = 1, 2, 4
JOINT_REVOLUTE, JOINT_FIXED = 0, 1
SHAPE_BOX, SHAPE_MESH, SHAPE_CAPSULE, SHAPE_CONVEX_MESH = 0, 1, 2, 3

STATUS = {0: "OK", 1: "INVALID_ARGUMENT", 2: "UNSUPPORTED_SHAPE", 3: "LINK_NOT_FOUND", 4: "UNKNOWN_JOINT_TYPE",
          5: "CUDA", 6: "NO_DEVICE", 7: "OUT_OF_MEMORY", 8: "UNSUPPORTED_ROBOT"}

EXPORTS = [
    "mgodpl_last_error", "mgodpl_device_count", "mgodpl_set_device", "mgodpl_scene_device", "mgodpl_scene_create", "mgodpl_scene_create_u64",
    "mgodpl_scene_destroy", "mgodpl_scene_get_info", "mgodpl_scene_get_counters", "mgodpl_robot_create", "mgodpl_robot_create_ex",
    "mgodpl_robot_destroy", "mgodpl_robot_n_variables", "mgodpl_forward_kinematics", "mgodpl_check_boxes",
    "mgodpl_check_states", "mgodpl_check_states_device", "mgodpl_check_motions", "mgodpl_check_motions_device",
    "mgodpl_check_path", "mgodpl_sample_goals", "mgodpl_sample_goals_device", "mgodpl_knn", "mgodpl_knn_device", "mgodpl_knn_query",
    "mgodpl_roadmap_shortest_paths", "mgodpl_roadmap_shortest_paths_device", "mgodpl_roadmap_repair_predecessors",
    "mgodpl_mesh_components",
    "mgodpl_mesh_components_u64",
]


class MgodplError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mgodpl_b200 status {code} ({STATUS.get(code, '?')}): {message}")
        self.code = code
        self.message = message


class ShapeDesc(C.Structure):
    _fields_ = [("link", C.c_int32), ("shape_type", C.c_int32), ("size", C.c_double * 3), ("transform", C.c_double * 7)]


class JointDesc(C.Structure):
    _fields_ = [("link_a", C.c_int32), ("link_b", C.c_int32), ("joint_type", C.c_int32), ("reserved", C.c_int32),
                ("attachment_a", C.c_double * 7), ("attachment_b", C.c_double * 7), ("axis", C.c_double * 3),
                ("min_angle", C.c_double), ("max_angle", C.c_double)]


class SceneOpts(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("max_leaf_size", C.c_uint32)]


class SceneInfo(C.Structure):
    _fields_ = [("n_triangles", C.c_uint64), ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64),
                ("device_bytes", C.c_uint64), ("bounds_lo", C.c_double * 3), ("bounds_hi", C.c_double * 3),
                ("build_ms", C.c_double), ("sah_cost", C.c_double), ("max_depth", C.c_uint32),
                ("max_leaf_size", C.c_uint32), ("n_wide_nodes", C.c_uint64), ("traversal_bytes", C.c_uint64),
                ("clearance_cells", C.c_uint64), ("clearance_cell_size", C.c_double), ("clearance_build_ms", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("states_checked", C.c_uint64), ("boxes_traversed", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("motions_checked", C.c_uint64), ("uncertain_steps", C.c_uint64),
                ("max_nodes_per_box", C.c_uint64)]


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``libmgodpl_b200.so`` next to this file (nvcc cross-compiles
    without a GPU)."""
    cmd = ["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j4"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MgodplError(-1, f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                                  f"(make -C {os.path.join(PKG_DIR, 'csrc')}); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.mgodpl_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)  # AttributeError here means the header and the library disagree
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise MgodplError(rc, lib().mgodpl_last_error().decode())


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _stream(stream) -> C.c_void_p:
    return C.c_void_p(int(stream) if stream else 0)


def _dptr(x) -> C.c_void_p:
    if x is None:
        return C.c_void_p(0)
    return C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else int(x))


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    """Bit i of the packed result -> bool array of length n (bit (i & 31) of word (i >> 5))."""
    return np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)


class Scene:
    """Device-resident BVH over a triangle soup; stands where ``fcl::CollisionObjectd`` stood."""

    def __init__(self, vertices, triangles, counters: bool = False, max_leaf_size: int = 0, refine: bool = True,
                 clearance_grid: bool = True):
        v = _f64(vertices).reshape(-1, 3)
        t = np.ascontiguousarray(triangles)
        opts = SceneOpts((SCENE_COUNTERS if counters else 0) | (0 if refine else SCENE_NO_REFINE) |
                         (0 if clearance_grid else SCENE_NO_CLEARANCE_GRID), max_leaf_size)
        self.h = C.c_void_p()
        if t.dtype == np.uint64:
            t = t.reshape(-1, 3)
            _check(lib().mgodpl_scene_create_u64(_p(v), C.c_size_t(len(v)), _p(t, C.c_uint64), C.c_size_t(len(t)),
                                                  C.byref(opts), C.byref(self.h)))
        else:
            t = np.ascontiguousarray(t, dtype=np.uint32).reshape(-1, 3)
            _check(lib().mgodpl_scene_create(_p(v), C.c_size_t(len(v)), _p(t, C.c_uint32), C.c_size_t(len(t)),
                                              C.byref(opts), C.byref(self.h)))

    @classmethod
    def from_mesh(cls, mesh, **kw) -> "Scene":
        return cls(mesh.vertices, mesh.triangles, **kw)

    def info(self) -> SceneInfo:
        out = SceneInfo()
        _check(lib().mgodpl_scene_get_info(self.h, C.byref(out)))
        return out

    def counters(self, reset: bool = False) -> dict:
        out = Counters()
        _check(lib().mgodpl_scene_get_counters(self.h, C.byref(out), int(reset)))
        return {k: int(getattr(out, k)) for k, _ in Counters._fields_}

    def close(self):
        if getattr(self, "h", None):
            lib().mgodpl_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


IDENTITY_TF = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)


class Robot:
    """Flattened ``RobotModel``: links by index, box shapes, joints in insertion order."""

    def __init__(self, n_links: int, shapes, joints, root_link: int, end_effector_link: int = -1):
        """shapes: iterable of (link, size3, tf7[, shape_type[, vertices]]) — size3 is the box size, or (radius, -, length) of
        a SHAPE_CAPSULE; a SHAPE_CONVEX_MESH carries its vertices (k x 3) as the fifth element; joints: iterable of
        (link_a, link_b, attach_a7, attach_b7, joint_type, axis3, min, max)."""
        shapes, joints = list(shapes), list(joints)
        sd = (ShapeDesc * max(len(shapes), 1))()
        mesh_verts, mesh_off = [], [0]
        for i, s in enumerate(shapes):
            sd[i].link, sd[i].shape_type = int(s[0]), int(s[3]) if len(s) > 3 else SHAPE_BOX
            sd[i].size[:] = [float(x) for x in s[1]]
            sd[i].transform[:] = [float(x) for x in s[2]]
            if len(s) > 4 and s[4] is not None:
                mesh_verts.append(_f64(s[4]).reshape(-1, 3))
            mesh_off.append(mesh_off[-1] + (len(mesh_verts[-1]) if len(s) > 4 and s[4] is not None else 0))
        jd = (JointDesc * max(len(joints), 1))()
        for i, j in enumerate(joints):
            jd[i].link_a, jd[i].link_b, jd[i].joint_type = int(j[0]), int(j[1]), int(j[4])
            jd[i].attachment_a[:] = [float(x) for x in j[2]]
            jd[i].attachment_b[:] = [float(x) for x in j[3]]
            jd[i].axis[:] = [float(x) for x in j[5]]
            jd[i].min_angle, jd[i].max_angle = float(j[6]), float(j[7])
        self.h = C.c_void_p()
        if mesh_verts:
            mv = np.ascontiguousarray(np.concatenate(mesh_verts))
            mo = (C.c_int32 * len(mesh_off))(*mesh_off)
            _check(lib().mgodpl_robot_create_ex(n_links, sd, len(shapes), jd, len(joints), root_link, end_effector_link,
                                                _p(mv), mo, C.byref(self.h)))
        else:
            _check(lib().mgodpl_robot_create(n_links, sd, len(shapes), jd, len(joints), root_link, end_effector_link,
                                             C.byref(self.h)))
        self.n_links, self.n_joints = n_links, len(joints)
        self.n_variables = lib().mgodpl_robot_n_variables(self.h)
        self.root_link, self.end_effector_link = root_link, end_effector_link

    def close(self):
        if getattr(self, "h", None):
            lib().mgodpl_robot_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _states2d(x):
    x = _f64(x)
    if x.ndim == 1:
        x = x[None]
    return x


def forward_kinematics(robot: Robot, states, stream=None) -> np.ndarray:
    s = _states2d(states)
    out = np.zeros((len(s), robot.n_links, 7))
    _check(lib().mgodpl_forward_kinematics(robot.h, _p(s), C.c_size_t(len(s)), C.c_size_t(s.shape[1]),
                                           s.shape[1] - 7, _p(out), _stream(stream)))
    return out


def check_boxes(scene: Scene, boxes, stream=None) -> np.ndarray:
    b = _f64(boxes).reshape(-1, 10)
    words = np.zeros((len(b) + 31) // 32, np.uint32)
    _check(lib().mgodpl_check_boxes(scene.h, _p(b), C.c_size_t(len(b)), _p(words, C.c_uint32), _stream(stream)))
    return unpack_bits(words, len(b))


def check_states(scene: Scene, robot: Robot, states, stream=None, packed: bool = False) -> np.ndarray:
    s = _states2d(states)
    n = len(s)
    words = np.zeros((n + 31) // 32, np.uint32)
    _check(lib().mgodpl_check_states(scene.h, robot.h, _p(s), C.c_size_t(n), C.c_size_t(s.shape[1]), s.shape[1] - 7,
                                     _p(words, C.c_uint32), _stream(stream)))
    return words if packed else unpack_bits(words, n)


def check_motions(scene: Scene, robot: Robot, a, b, max_step: float = 0.1, want_toi: bool = True, stream=None,
                  out: dict | None = None) -> dict:
    """M x check_motion_collides.  Returns dict(hit[bool m], toi[m], n_steps[m], words[packed]).

    ``out``: a dict from ``motion_result_buffers(m)`` to write into instead of allocating (repeated calls on batches of
    one size); it is returned with words / toi / n_steps (uint32) filled and no unpacked ``hit``."""
    a, b = _states2d(a), _states2d(b)
    assert a.shape == b.shape
    m = len(a)
    if out is None:
        words = np.zeros((m + 31) // 32, np.uint32)
        toi = np.zeros(m) if want_toi else None
        steps = np.zeros(m, np.uint32)
    else:
        words, toi, steps = out["words"], out["toi"] if want_toi else None, out["n_steps"]
        if len(words) != (m + 31) // 32 or len(steps) != m or (toi is not None and len(toi) != m):
            raise MgodplError(-1, "result buffers do not match the batch size")
    _check(lib().mgodpl_check_motions(scene.h, robot.h, _p(a), _p(b), C.c_size_t(m), C.c_size_t(a.shape[1]),
                                      a.shape[1] - 7, C.c_double(max_step), _p(words, C.c_uint32), _p(toi),
                                      _p(steps, C.c_uint32), _stream(stream)))
    if out is not None:
        return out
    return dict(hit=unpack_bits(words, m), toi=toi, n_steps=steps.astype(np.int64), words=words)


def motion_result_buffers(m: int) -> dict:
    """Reusable result arrays for ``check_motions(..., out=...)``."""
    return dict(words=np.zeros((m + 31) // 32, np.uint32), toi=np.zeros(m), n_steps=np.zeros(m, np.uint32))


def check_path(scene: Scene, robot: Robot, states, max_step: float = 0.1, stream=None) -> int:
    s = _states2d(states)
    first = C.c_int64(-1)
    _check(lib().mgodpl_check_path(scene.h, robot.h, _p(s), C.c_size_t(len(s)), C.c_size_t(s.shape[1]),
                                   s.shape[1] - 7, C.c_double(max_step), C.byref(first), _stream(stream)))
    return int(first.value)


def sample_goals(scene: Scene, robot: Robot, targets, samples: int, draws=None, seed: int = 42,
                 distance_from_target: float = 0.0, want_states: bool = False, stream=None) -> dict:
    t = _f64(targets).reshape(-1, 3)
    f = len(t)
    wpr = (samples + 31) // 32
    words = np.zeros((f, wpr), np.uint32)
    counts = np.zeros(f, np.uint32)
    d = None
    if draws is not None:
        d = _f64(draws).reshape(samples, 1 + robot.n_variables)
    states = np.zeros((f, samples, 7 + robot.n_variables)) if want_states else None
    _check(lib().mgodpl_sample_goals(scene.h, robot.h, _p(t), C.c_size_t(f), _p(d), C.c_size_t(samples),
                                     C.c_uint64(seed), C.c_double(distance_from_target), _p(words, C.c_uint32),
                                     _p(counts, C.c_uint32), _p(states), _stream(stream)))
    hits = np.stack([unpack_bits(words[i], samples) for i in range(f)]) if f else np.zeros((0, samples), bool)
    return dict(hit=hits, collisions=counts, states=states, words=words)


def knn(states, k: int, causal: bool = False, want_dist: bool = True, stream=None):
    """(idx[n,k], dist[n,k]) nearest neighbours under equal_weights_distance; causal: only j < i (incremental PRM)."""
    s = _states2d(states)
    n = len(s)
    idx = np.full((n, k), -1, np.int32)
    dist = np.zeros((n, k)) if want_dist else None
    _check(lib().mgodpl_knn(_p(s), C.c_size_t(n), C.c_size_t(s.shape[1]), s.shape[1] - 7, k, int(causal), _p(idx, C.c_int32),
                            _p(dist), _stream(stream)))
    return idx, dist


def knn_query(database, queries, k: int, want_dist: bool = True, stream=None):
    """(idx[n_q,k], dist[n_q,k]): the k nearest rows of `database` for every row of `queries` (equal_weights_distance)."""
    db, q = _states2d(database), _states2d(queries)
    assert db.shape[1] == q.shape[1]
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.zeros((len(q), k)) if want_dist else None
    _check(lib().mgodpl_knn_query(_p(db), C.c_size_t(len(db)), _p(q), C.c_size_t(len(q)), C.c_size_t(q.shape[1]), q.shape[1] - 7, k,
                                  _p(idx, C.c_int32), _p(dist), _stream(stream)))
    return idx, dist


def knn_device(d_states, n: int, stride: int, n_joint_values: int, k: int, d_idx, d_dist=None, causal: bool = False, stream=None):
    _check(lib().mgodpl_knn_device(_dptr(d_states), C.c_size_t(n), C.c_size_t(stride), n_joint_values, k, int(causal),
                                   _dptr(d_idx), _dptr(d_dist), _stream(stream)))


def mesh_components(triangles, n_vertices: int, stream=None):
    """labels[v] = smallest vertex index of v's connected component (connected_vertex_components, canonical form)."""
    tris = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    labels = np.zeros(n_vertices, np.uint32)
    _check(lib().mgodpl_mesh_components(_p(tris, C.c_uint32), C.c_size_t(len(tris)), C.c_size_t(n_vertices), _p(labels, C.c_uint32),
                                        _stream(stream)))
    return labels


def roadmap_shortest_paths(n_vertices: int, edges, weights, sources, want_pred: bool = True, stream=None):
    """Shortest paths from every source over an undirected roadmap (runDijkstra per source, tsp_over_prm.cpp:80-97).

    Returns dict(dist[n_sources, n_vertices], pred[n_sources, n_vertices] | None, sweeps)."""
    edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
    weights = _f64(weights).reshape(-1)
    sources = np.ascontiguousarray(sources, dtype=np.uint32).reshape(-1)
    if len(weights) != len(edges):
        raise MgodplError(-1, "one weight per edge")
    dist = np.zeros((len(sources), n_vertices))
    pred = np.zeros((len(sources), n_vertices), np.uint32) if want_pred else None
    sweeps = C.c_uint32(0)
    _check(lib().mgodpl_roadmap_shortest_paths(C.c_uint32(n_vertices), _p(edges, C.c_uint32), _p(weights), C.c_size_t(len(edges)),
                                               _p(sources, C.c_uint32), C.c_size_t(len(sources)), _p(dist), _p(pred, C.c_uint32),
                                               C.byref(sweeps), _stream(stream)))
    return dict(dist=dist, pred=pred, sweeps=int(sweeps.value))


def roadmap_repair_predecessors(row_offsets, neighbours, weights, dist, pred):
    """Host-side: one source's predecessors made a tree again where vanishing (zero-weight) edges let vertices at one distance
    name each other.  CSR graph with both directions of every edge; returns the repaired copy of `pred`."""
    row = np.ascontiguousarray(row_offsets, dtype=np.uint32).reshape(-1)
    col = np.ascontiguousarray(neighbours, dtype=np.uint32).reshape(-1)
    w = _f64(weights).reshape(-1)
    dist = _f64(dist).reshape(-1)
    pred = np.array(pred, dtype=np.uint32).reshape(-1)
    n = len(row) - 1
    if n < 0 or len(dist) != n or len(pred) != n or len(col) != len(w) or (n >= 0 and len(col) != int(row[-1])):
        raise MgodplError(-1, "row_offsets: n_vertices + 1 entries; dist / pred: n_vertices; neighbours / weights: row_offsets[-1]")
    _check(lib().mgodpl_roadmap_repair_predecessors(C.c_uint32(n), _p(row, C.c_uint32), _p(col, C.c_uint32), _p(w), _p(dist),
                                                    _p(pred, C.c_uint32)))
    return pred


def roadmap_shortest_paths_device(d_row_offsets, d_neighbours, d_weights, n_vertices: int, d_sources, n_sources: int, d_dist,
                                  d_pred=None, stream=None) -> int:
    """CSR graph and outputs in device memory; synchronises the stream; returns the number of sweeps enqueued."""
    sweeps = C.c_uint32(0)
    _check(lib().mgodpl_roadmap_shortest_paths_device(_dptr(d_row_offsets), _dptr(d_neighbours), _dptr(d_weights),
                                                      C.c_uint32(n_vertices), _dptr(d_sources), C.c_size_t(n_sources),
                                                      _dptr(d_dist), _dptr(d_pred), C.byref(sweeps), _stream(stream)))
    return int(sweeps.value)


# ---- device-pointer variants (torch tensors or raw pointers; asynchronous on `stream`) -------------------------------
def check_states_device(scene: Scene, robot: Robot, d_states, n: int, stride: int, n_joint_values: int, d_hit_bits,
                        stream=None):
    _check(lib().mgodpl_check_states_device(scene.h, robot.h, _dptr(d_states), C.c_size_t(n), C.c_size_t(stride),
                                            n_joint_values, _dptr(d_hit_bits), _stream(stream)))


def check_motions_device(scene: Scene, robot: Robot, d_a, d_b, m: int, stride: int, n_joint_values: int,
                         d_hit_bits, max_step: float = 0.1, d_toi=None, d_n_steps=None, d_uncertain=None,
                         d_override=None, stream=None):
    _check(lib().mgodpl_check_motions_device(scene.h, robot.h, _dptr(d_a), _dptr(d_b), C.c_size_t(m),
                                             C.c_size_t(stride), n_joint_values, C.c_double(max_step),
                                             _dptr(d_override), _dptr(d_hit_bits), _dptr(d_toi), _dptr(d_n_steps),
                                             _dptr(d_uncertain), _stream(stream)))


def sample_goals_device(scene: Scene, robot: Robot, d_targets, f: int, samples: int, d_hit_bits, d_draws=None,
                        seed: int = 42, distance_from_target: float = 0.0, d_collisions=None, d_states_out=None,
                        stream=None):
    _check(lib().mgodpl_sample_goals_device(scene.h, robot.h, _dptr(d_targets), C.c_size_t(f), _dptr(d_draws),
                                            C.c_size_t(samples), C.c_uint64(seed), C.c_double(distance_from_target),
                                            _dptr(d_hit_bits), _dptr(d_collisions), _dptr(d_states_out),
                                            _stream(stream)))
