"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end over ``oracle/liboracle.so`` (the fp64 CPU restatement, see ``mgodpl_oracle.hpp``) and, when it was
built, ``oracle/_ref/libmgodpl_ref.so`` (the reference's own sources over the FCL/Eigen shim, see ``ref_capi.cpp``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the product
package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libmgodpl_ref.so")

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(ref: bool = True) -> None:
    """Compile the restatement and, where /root/reference exists, the reference-source library."""
    subprocess.check_call(["make", "-s", "-C", HERE])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _d(a):
    return a.ctypes.data_as(_dp)


def _arr(x, dtype=np.float64):
    return np.ascontiguousarray(x, dtype=dtype)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_robot_procedural.restype = C.c_void_p
        L.orc_robot_new.restype = C.c_void_p
        L.orc_scene_create.restype = C.c_void_p
        L.orc_distance.restype = C.c_double
        L.orc_angular_distance.restype = C.c_double
        L.orc_motion_steps.restype = C.c_uint64
        L.orc_check_path.restype = C.c_long
        L.orc_shortcut_path.restype = C.c_long
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_PATH)
        R.ref_robot_procedural.restype = C.c_void_p
        R.ref_scene_create.restype = C.c_void_p
        R.ref_distance.restype = C.c_double
        R.ref_check_path.restype = C.c_long
        R.ref_shortcut_path.restype = C.c_long
        _ref = R
    return _ref


HORIZONTAL, VERTICAL = 0, 1
SHORTCUT_GLOBALLY, DELETE_EVERY_WAYPOINT, SHORTCUT_LOCALLY, MIDPOINT_PULL = 0, 1, 2, 3
REVOLUTE, FIXED = 0, 1


class Robot:
    """Procedural drone (createProceduralRobotModel) or a hand-built link/joint graph."""

    def __init__(self, arm_length: float | None = 1.0, joint_types=(HORIZONTAL,)):
        L = lib()
        if arm_length is None:
            self.h = C.c_void_p(L.orc_robot_new())
        else:
            jt = (C.c_int * len(joint_types))(*joint_types)
            self.h = C.c_void_p(L.orc_robot_procedural(C.c_double(arm_length), jt, len(joint_types)))
        self.arm_length, self.joint_types = arm_length, tuple(joint_types)

    # generic construction ---------------------------------------------------------------------------------
    def add_link(self, name: str) -> int:
        return lib().orc_robot_add_link(self.h, name.encode())

    def add_box(self, link: int, size, tf=(0, 0, 0, 0, 0, 0, 1)):
        lib().orc_robot_add_box(self.h, link, _d(_arr(size)), _d(_arr(tf)))

    # link shapes the reference does not have; `other_shapes` records them for helpers.device_robot_from_oracle
    def add_capsule(self, link: int, radius: float, length: float, tf=(0, 0, 0, 0, 0, 0, 1)):
        lib().orc_robot_add_capsule(self.h, link, C.c_double(radius), C.c_double(length), _d(_arr(tf)))
        self.__dict__.setdefault("other_shapes", []).append((link, (radius, radius, length), tuple(tf), 2, None))

    def add_convex(self, link: int, verts, tf=(0, 0, 0, 0, 0, 0, 1)):
        v = _arr(verts).reshape(-1, 3)
        lib().orc_robot_add_convex(self.h, link, _d(v), len(v), _d(_arr(tf)))
        self.__dict__.setdefault("other_shapes", []).append((link, (0.0, 0.0, 0.0), tuple(tf), 3, v.copy()))

    def add_joint(self, link_a, link_b, attach_a, attach_b, kind=FIXED, axis=(1, 0, 0), lo=0.0, hi=0.0) -> int:
        return lib().orc_robot_add_joint(self.h, link_a, link_b, _d(_arr(attach_a)), _d(_arr(attach_b)), kind,
                                         _d(_arr(axis)), C.c_double(lo), C.c_double(hi))

    # introspection ------------------------------------------------------------------------------------------
    @property
    def n_links(self):
        return lib().orc_robot_n_links(self.h)

    @property
    def n_joints(self):
        return lib().orc_robot_n_joints(self.h)

    @property
    def n_variables(self):
        return lib().orc_robot_n_variables(self.h)

    def find_link(self, name: str) -> int:
        return lib().orc_robot_find_link(self.h, name.encode())

    def describe(self):
        """(boxes[nb,11], joints[nj,22]) flat description — see orc_robot_describe."""
        boxes = np.zeros((64, 11))
        joints = np.zeros((max(self.n_joints, 1), 22))
        nb = lib().orc_robot_describe(self.h, _d(boxes), 64, _d(joints), self.n_joints)
        return boxes[:nb].copy(), joints[: self.n_joints].copy()

    # path functions -------------------------------------------------------------------------------------------
    def fk(self, state, root: int | None = None):
        state = _arr(state)
        out = np.zeros((self.n_links, 7))
        if root is None:
            root = self.find_link("flying_base")
        lib().orc_fk(self.h, _d(state), len(state) - 7, root, _d(out))
        return out

    def __del__(self):
        try:
            lib().orc_robot_free(self.h)
        except Exception:
            pass


def distance(a, b) -> float:
    a, b = _arr(a), _arr(b)
    return lib().orc_distance(_d(a), _d(b), len(a) - 7)


def angular_distance(qa, qb) -> float:
    return lib().orc_angular_distance(_d(_arr(qa)), _d(_arr(qb)))


def interpolate(a, b, t: float):
    a, b = _arr(a), _arr(b)
    out = np.zeros_like(a)
    lib().orc_interpolate(_d(a), _d(b), len(a) - 7, C.c_double(t), _d(out))
    return out


def slerp(qa, qb, t: float):
    out = np.zeros(4)
    lib().orc_slerp(_d(_arr(qa)), _d(_arr(qb)), C.c_double(t), _d(out))
    return out


def tf_then(a, b):
    out = np.zeros(7)
    lib().orc_tf_then(_d(_arr(a)), _d(_arr(b)), _d(out))
    return out


def tf_inverse(a):
    out = np.zeros(7)
    lib().orc_tf_inverse(_d(_arr(a)), _d(out))
    return out


def motion_steps(d: float, max_step: float = 0.1) -> int:
    return int(lib().orc_motion_steps(C.c_double(d), C.c_double(max_step)))


def shape_tri(kind, tf7, tri9, radius=0.0, length=0.0, verts=None):
    """(hit, signed clearance) of one capsule (kind 1) or convex polytope (kind 2) against one triangle (world frame)."""
    cl = C.c_double()
    v = _arr(verts).reshape(-1, 3) if verts is not None else np.zeros((0, 3))
    hit = lib().orc_shape_tri(kind, _d(_arr(tf7)), C.c_double(radius), C.c_double(length), _d(v), len(v), _d(_arr(tri9).reshape(-1)), C.byref(cl))
    return bool(hit), cl.value


def box_tri(tf7, size3, tri9):
    """(hit, signed clearance) of one oriented box against one triangle."""
    cl = C.c_double()
    hit = lib().orc_box_tri(_d(_arr(tf7)), _d(_arr(size3)), _d(_arr(tri9).reshape(-1)), C.byref(cl))
    return bool(hit), cl.value


MPR_STAGES = ("portal discovery: apart", "portal discovery: centre on the origin ray", "refinement: origin enclosed",
              "refinement: support plane short of the origin", "refinement: stopped at the tolerance", "iteration guard")


def box_tri_batch(tf7, size3, tri9, tolerance: float = 1e-6):
    """n box / triangle pairs: the exact answer (hit, signed clearance) and the libccd-style MPR second opinion
    (mpr_second_opinion.hpp; mpr_stage indexes MPR_STAGES).  Never "FCL": a restatement of the published algorithm."""
    tf7, size3 = _arr(tf7).reshape(-1, 7), _arr(size3).reshape(-1, 3)
    tri9 = _arr(tri9).reshape(-1, 9)
    n = len(tf7)
    assert len(size3) == n and len(tri9) == n
    hit, mpr_hit = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    cl, it, st = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    i32p = C.POINTER(C.c_int32)
    lib().orc_box_tri_batch(C.c_size_t(n), _d(tf7), _d(size3), _d(tri9), C.c_double(tolerance), hit.ctypes.data_as(_u8p), _d(cl),
                            mpr_hit.ctypes.data_as(_u8p), it.ctypes.data_as(i32p), st.ctypes.data_as(i32p))
    return dict(hit=hit.astype(bool), clearance=cl, mpr_hit=mpr_hit.astype(bool), mpr_iterations=it, mpr_stage=st)


class Scene:
    def __init__(self, verts, tris):
        self.verts = _arr(verts).reshape(-1, 3)
        self.tris = _arr(tris, np.uint32).reshape(-1, 3)
        self.h = C.c_void_p(lib().orc_scene_create(_d(self.verts), C.c_size_t(len(self.verts)),
                                                   self.tris.ctypes.data_as(_u32p), C.c_size_t(len(self.tris))))
        if not self.h:
            raise ValueError("bad mesh")

    def bounds(self):
        out = np.zeros(6)
        lib().orc_scene_bounds(self.h, _d(out))
        return out[:3], out[3:]

    def check_obbs(self, boxes, clearance=False, reach=1e-3):
        boxes = _arr(boxes).reshape(-1, 10)
        hits = np.zeros(len(boxes), np.uint8)
        cl = np.zeros(len(boxes)) if clearance else None
        lib().orc_check_obbs(self.h, _d(boxes), C.c_size_t(len(boxes)), hits.ctypes.data_as(_u8p),
                             _d(cl) if clearance else None, C.c_double(reach))
        return (hits.astype(bool), cl) if clearance else hits.astype(bool)

    def check_states(self, robot: Robot, states, clearance=False, reach=1e-3, threads=1, counters=False):
        states = _arr(states)
        states = states.reshape(-1, states.shape[-1])
        n, stride = states.shape
        hits = np.zeros(n, np.uint8)
        cl = np.zeros(n) if clearance else None
        cnt = np.zeros(3, np.uint64)
        lib().orc_check_states(self.h, robot.h, _d(states), C.c_size_t(n), C.c_size_t(stride), stride - 7,
                               hits.ctypes.data_as(_u8p), _d(cl) if clearance else None, C.c_double(reach), threads,
                               cnt.ctypes.data_as(_u64p))
        out = [hits.astype(bool)]
        if clearance:
            out.append(cl)
        if counters:
            out.append(cnt)
        return out[0] if len(out) == 1 else tuple(out)

    def check_motions(self, robot: Robot, a, b, max_step=0.1, band=False, reach=1e-3, threads=1):
        """dict(hit, toi, n_steps, states_checked[, min_abs_clearance])."""
        a, b = _arr(a), _arr(b)
        a = a.reshape(-1, a.shape[-1])
        b = b.reshape(-1, b.shape[-1])
        m, stride = a.shape
        hits = np.zeros(m, np.uint8)
        toi = np.zeros(m)
        ns = np.zeros(m, np.uint64)
        sc = np.zeros(m, np.uint64)
        mc = np.zeros(m) if band else None
        lib().orc_check_motions(self.h, robot.h, _d(a), _d(b), C.c_size_t(m), C.c_size_t(stride), stride - 7,
                                C.c_double(max_step), hits.ctypes.data_as(_u8p), _d(toi), ns.ctypes.data_as(_u64p),
                                sc.ctypes.data_as(_u64p), _d(mc) if band else None, C.c_double(reach), threads)
        out = dict(hit=hits.astype(bool), toi=toi, n_steps=ns.astype(np.int64), states_checked=sc.astype(np.int64))
        if band:
            out["min_abs_clearance"] = mc
        return out

    def check_path(self, robot: Robot, states, max_step=0.1) -> int:
        states = _arr(states)
        w, stride = states.shape
        return lib().orc_check_path(self.h, robot.h, _d(states), C.c_size_t(w), C.c_size_t(stride), stride - 7,
                                    C.c_double(max_step))

    def shortcut_path(self, robot: Robot, states, seed=42, iterations=100, mode=SHORTCUT_GLOBALLY, pull=0.5, max_step=0.1):
        """local_optimization.cpp driven as tsp_over_prm.cpp:594-612 does; returns dict(path, n_success, n_checks)."""
        states = _arr(states)
        w, stride = states.shape
        out = np.zeros((w + 2 * iterations + 2, stride))
        out_w, checks = C.c_size_t(0), C.c_size_t(0)
        ok = lib().orc_shortcut_path(self.h, robot.h, _d(states), C.c_size_t(w), C.c_size_t(stride), stride - 7, C.c_uint(seed),
                                     C.c_size_t(iterations), mode, C.c_double(pull), C.c_double(max_step), _d(out),
                                     C.c_size_t(len(out)), C.byref(out_w), C.byref(checks))
        return dict(path=out[:out_w.value].copy(), n_success=int(ok), n_checks=int(checks.value))

    def goal_sample_counts(self, robot: Robot, targets, samples: int, seed=42, threads=1, want_hits=False):
        targets = _arr(targets).reshape(-1, 3)
        f = len(targets)
        counts = np.zeros(f, np.uint32)
        hits = np.zeros((f, samples), np.uint8) if want_hits else None
        lib().orc_goal_sample_counts(self.h, robot.h, C.c_uint(seed), _d(targets), C.c_size_t(f), C.c_size_t(samples),
                                     counts.ctypes.data_as(_u32p), hits.ctypes.data_as(_u8p) if want_hits else None,
                                     threads)
        return (counts, hits.astype(bool)) if want_hits else counts

    def __del__(self):
        try:
            lib().orc_scene_free(self.h)
        except Exception:
            pass


def gen_uniform_states(robot: Robot, n: int, seed=42, h_range=5.0, v_range=10.0, joint_range=np.pi / 2):
    out = np.zeros((n, 7 + robot.n_joints))
    lib().orc_gen_uniform_states(robot.h, C.c_uint(seed), C.c_double(h_range), C.c_double(v_range),
                                 C.c_double(joint_range), C.c_size_t(n), _d(out))
    return out


def goal_draws(robot: Robot, n: int, seed=42):
    out = np.zeros((n, 1 + robot.n_variables))
    lib().orc_goal_draws(robot.h, C.c_uint(seed), C.c_size_t(n), _d(out))
    return out


def gen_goal_states(robot: Robot, target, n: int, seed=42, distance_from_target=0.0):
    out = np.zeros((n, 7 + robot.n_variables))
    lib().orc_gen_goal_states(robot.h, C.c_uint(seed), _d(_arr(target)), C.c_double(distance_from_target),
                              C.c_size_t(n), _d(out))
    return out


def knn(states, k: int, causal: bool = False):
    states = _arr(states)
    n, stride = states.shape
    idx = np.zeros((n, k), np.int32)
    dist = np.zeros((n, k))
    lib().orc_knn(_d(states), C.c_size_t(n), C.c_size_t(stride), stride - 7, k, int(causal),
                  idx.ctypes.data_as(C.POINTER(C.c_int32)), _d(dist))
    return idx, dist


def roadmap_shortest_paths(n_vertices: int, edges, weights, sources, threads=1):
    """runDijkstra (tsp_over_prm.cpp:80-97) from every source; returns (dist, pred), each n_sources x n_vertices."""
    edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
    weights = _arr(weights).reshape(-1)
    sources = np.ascontiguousarray(sources, dtype=np.uint32).reshape(-1)
    dist = np.zeros((len(sources), n_vertices))
    pred = np.zeros((len(sources), n_vertices), np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().orc_roadmap_shortest_paths(C.c_uint32(n_vertices), edges.ctypes.data_as(u32p), _d(weights), C.c_size_t(len(edges)),
                                     sources.ctypes.data_as(u32p), C.c_size_t(len(sources)), _d(dist), pred.ctypes.data_as(u32p),
                                     threads)
    return dist, pred


def mesh_components(triangles, n_vertices: int):
    """connected_vertex_components (mesh_connected_components.cpp:25-74), canonical labels (smallest vertex index)."""
    tris = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    labels = np.zeros(n_vertices, np.uint32)
    lib().orc_mesh_components(tris.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(len(tris)), C.c_size_t(n_vertices),
                              labels.ctypes.data_as(C.POINTER(C.c_uint32)))
    return labels


def ref_mesh_components(triangles, n_vertices: int):
    tris = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    labels = np.zeros(n_vertices, np.uint32)
    ref().ref_mesh_components.restype = C.c_size_t
    n = ref().ref_mesh_components(tris.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(len(tris)), C.c_size_t(n_vertices),
                                  labels.ctypes.data_as(C.POINTER(C.c_uint32)))
    return labels, int(n)


def from_ee_and_vector(robot: Robot, point, vec):
    out = np.zeros(8)
    lib().orc_from_ee_and_vector(robot.h, _d(_arr(point)), _d(_arr(vec)), _d(out))
    return out


def arm_vector(robot: Robot, state):
    state = _arr(state)
    out = np.zeros(3)
    lib().orc_arm_vector(robot.h, _d(state), len(state) - 7, _d(out))
    return out


def _sample_goal_states(L, fn, scene, robot, which, targets, seed, max_attempts, ee_distance, row):
    targets = _arr(targets).reshape(-1, 3)
    n = len(targets)
    out = np.zeros((n, row))
    found = np.zeros(n, np.uint8)
    nxt = np.zeros(n, np.uint32)
    getattr(L, fn)(scene.h, robot.h, which, C.c_uint(seed), _d(targets), C.c_size_t(n), int(max_attempts), C.c_double(ee_distance),
                   _d(out), found.ctypes.data_as(_u8p), nxt.ctypes.data_as(_u32p))
    return dict(states=out, found=found.astype(bool), next_draw=nxt)


def sample_arm_vector_states(scene: "Scene", robot: Robot, targets, seed=42, max_attempts=100, ee_distance=0.0):
    """generateUniformRandomArmVectorState (goal_sampling.cpp:101-124) per target with std::mt19937(seed + i)."""
    return _sample_goal_states(lib(), "orc_sample_goal_states", scene, robot, 0, targets, seed, max_attempts, ee_distance, 8)


def sample_uniform_goal_states(scene: "Scene", robot: Robot, targets, seed=42, max_attempts=100):
    """findGoalStateByUniformSampling (goal_sampling.cpp:81-99) per target with std::mt19937(seed + i)."""
    return _sample_goal_states(lib(), "orc_sample_goal_states", scene, robot, 1, targets, seed, max_attempts, 0.0, 7 + robot.n_variables)


# ---------------------------------------------------------------------------------------------------------------
# The reference's own sources (oracle/_ref) — same call shapes, for cross-checking the restatement.
# ---------------------------------------------------------------------------------------------------------------
class RefRobot:
    def __init__(self, arm_length=1.0, joint_types=(HORIZONTAL,)):
        jt = (C.c_int * len(joint_types))(*joint_types)
        self.h = C.c_void_p(ref().ref_robot_procedural(C.c_double(arm_length), jt, len(joint_types)))
        self.n_links = ref().ref_robot_n_links(self.h)
        self.n_joints = ref().ref_robot_n_joints(self.h)
        self.n_variables = len(joint_types)

    def fk(self, state):
        state = _arr(state)
        out = np.zeros((self.n_links, 7))
        ref().ref_fk(self.h, _d(state), len(state) - 7, _d(out))
        return out

    def __del__(self):
        try:
            ref().ref_robot_free(self.h)
        except Exception:
            pass


def ref_gnat_causal_knn(states, k: int, seed: int = 42):
    """The reference's own NearestNeighborsGNAT (greedy-k-centres pivots, equal_weights_distance) used as
    add_and_connect_roadmap_node uses it: nearestK among the states added before, then add.  Returns (idx, dist), n x k."""
    states = _arr(states)
    n, stride = states.shape
    idx = np.zeros((n, k), np.int32)
    dist = np.zeros((n, k))
    ref().ref_gnat_causal_knn(_d(states), C.c_size_t(n), C.c_size_t(stride), stride - 7, k, C.c_uint(seed),
                              idx.ctypes.data_as(C.POINTER(C.c_int32)), _d(dist))
    return idx, dist


def ref_distance(a, b):
    a, b = _arr(a), _arr(b)
    return ref().ref_distance(_d(a), _d(b), len(a) - 7)


def ref_interpolate(a, b, t):
    a, b = _arr(a), _arr(b)
    out = np.zeros_like(a)
    ref().ref_interpolate(_d(a), _d(b), len(a) - 7, C.c_double(t), _d(out))
    return out


class RefScene:
    def __init__(self, verts, tris):
        verts = _arr(verts).reshape(-1, 3)
        tris = _arr(tris, np.uint32).reshape(-1, 3)
        self.h = C.c_void_p(ref().ref_scene_create(_d(verts), C.c_size_t(len(verts)), tris.ctypes.data_as(_u32p),
                                                   C.c_size_t(len(tris))))

    def check_states(self, robot: RefRobot, states):
        states = _arr(states)
        n, stride = states.shape
        hits = np.zeros(n, np.uint8)
        ref().ref_check_states(self.h, robot.h, _d(states), C.c_size_t(n), C.c_size_t(stride), stride - 7,
                               hits.ctypes.data_as(_u8p))
        return hits.astype(bool)

    def check_motions(self, robot: RefRobot, a, b):
        a, b = _arr(a), _arr(b)
        m, stride = a.shape
        hits = np.zeros(m, np.uint8)
        toi = np.zeros(m)
        ref().ref_check_motions(self.h, robot.h, _d(a), _d(b), C.c_size_t(m), C.c_size_t(stride), stride - 7,
                                hits.ctypes.data_as(_u8p), _d(toi))
        return hits.astype(bool), toi

    def check_path(self, robot: RefRobot, states) -> int:
        states = _arr(states)
        w, stride = states.shape
        return ref().ref_check_path(self.h, robot.h, _d(states), C.c_size_t(w), C.c_size_t(stride), stride - 7)

    def shortcut_path(self, robot: RefRobot, states, seed=42, iterations=100, mode=SHORTCUT_GLOBALLY, pull=0.5):
        states = _arr(states)
        w, stride = states.shape
        out = np.zeros((w + 2 * iterations + 2, stride))
        out_w, checks = C.c_size_t(0), C.c_size_t(0)
        ok = ref().ref_shortcut_path(self.h, robot.h, _d(states), C.c_size_t(w), C.c_size_t(stride), stride - 7, C.c_uint(seed),
                                     C.c_size_t(iterations), mode, C.c_double(pull), _d(out), C.c_size_t(len(out)),
                                     C.byref(out_w), C.byref(checks))
        return dict(path=out[:out_w.value].copy(), n_success=int(ok), n_checks=int(checks.value))

    def __del__(self):
        try:
            ref().ref_scene_free(self.h)
        except Exception:
            pass


def ref_gen_uniform_states(robot: RefRobot, n, seed=42, h_range=5.0, v_range=10.0, joint_range=np.pi / 2):
    out = np.zeros((n, 7 + robot.n_joints))
    ref().ref_gen_uniform_states(robot.h, C.c_uint(seed), C.c_double(h_range), C.c_double(v_range),
                                 C.c_double(joint_range), C.c_size_t(n), _d(out))
    return out


def ref_gen_goal_states(robot: RefRobot, target, n, seed=42, distance_from_target=0.0):
    out = np.zeros((n, 7 + robot.n_variables))
    ref().ref_gen_goal_states(robot.h, C.c_uint(seed), _d(_arr(target)), C.c_double(distance_from_target),
                              C.c_size_t(n), _d(out))
    return out


def ref_from_ee_and_vector(robot: RefRobot, point, vec):
    out = np.zeros(8)
    ref().ref_from_ee_and_vector(robot.h, _d(_arr(point)), _d(_arr(vec)), _d(out))
    return out


def ref_arm_vector(robot: RefRobot, state):
    state = _arr(state)
    out = np.zeros(3)
    ref().ref_arm_vector(robot.h, _d(state), len(state) - 7, _d(out))
    return out


def ref_sample_arm_vector_states(scene: "RefScene", robot: RefRobot, targets, seed=42, max_attempts=100, ee_distance=0.0):
    return _sample_goal_states(ref(), "ref_sample_goal_states", scene, robot, 0, targets, seed, max_attempts, ee_distance, 8)


def ref_sample_uniform_goal_states(scene: "RefScene", robot: RefRobot, targets, n_variables: int, seed=42, max_attempts=100):
    return _sample_goal_states(ref(), "ref_sample_goal_states", scene, robot, 1, targets, seed, max_attempts, 0.0, 7 + n_variables)
