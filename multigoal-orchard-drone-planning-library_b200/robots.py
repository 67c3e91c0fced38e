"""Host-side mirror (Python flavour) of the reference's robot construction and state samplers for this path.

The C++ facade in ``host/mgodpl/`` is the primary mirror (the reference is C++); this module exists so that
``bench.py`` and the pytest suite can describe robots and draw states without a compiler.  Nothing here computes
collisions — that is the CUDA library's job.

Reference (paths relative to its checkout):
  createProceduralRobotModel      src/experiment_utils/procedural_robot_models.cpp:86-154
  generateUniformRandomState      src/planning/state_tools.cpp:44-67
  genUprightState                 src/planning/goal_sampling.cpp:13-49
  RandomNumberGenerator           src/planning/RandomNumberGenerator.h:21-77
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi

HORIZONTAL, VERTICAL = 0, 1


@dataclass
class RobotModel:
    """Link / joint graph in the reference's insertion order (RobotModel.h:36-196)."""
    link_names: list = field(default_factory=list)
    shapes: list = field(default_factory=list)  # (link, size3, tf7)
    joints: list = field(default_factory=list)  # (link_a, link_b, attach_a7, attach_b7, type, axis3, min, max)

    def insert_link(self, name: str, box_size=None) -> int:
        self.link_names.append(name)
        if box_size is not None:
            self.shapes.append((len(self.link_names) - 1, tuple(box_size), capi.IDENTITY_TF))
        return len(self.link_names) - 1

    def add_capsule(self, link: int, radius: float, length: float, tf=None):
        """A capsule link shape (not in the reference): the points within ``radius`` of a segment of ``length`` along the
        shape's z axis, centred (fcl::Capsule's convention)."""
        self.shapes.append((link, (radius, radius, length), tuple(tf) if tf is not None else capi.IDENTITY_TF, capi.SHAPE_CAPSULE))

    def add_convex_mesh(self, link: int, vertices, tf=None):
        """A convex-polytope link shape (not in the reference): the hull of 4..16 points given in the shape frame."""
        self.shapes.append((link, (0.0, 0.0, 0.0), tuple(tf) if tf is not None else capi.IDENTITY_TF, capi.SHAPE_CONVEX_MESH,
                            np.asarray(vertices, dtype=np.float64).reshape(-1, 3)))

    def insert_joint(self, link_a, link_b, attach_a, attach_b, joint_type, axis=(1.0, 0.0, 0.0), lo=0.0, hi=0.0) -> int:
        self.joints.append((link_a, link_b, tuple(attach_a), tuple(attach_b), joint_type, tuple(axis), lo, hi))
        return len(self.joints) - 1

    def find_link_by_name(self, name: str) -> int:
        for i, n in enumerate(self.link_names):
            if n == name:
                return i
        raise RuntimeError("Link not found: " + name)  # RobotModel.cpp:119

    def count_joint_variables(self) -> int:
        return sum(1 for j in self.joints if j[4] == capi.JOINT_REVOLUTE)

    def revolute_limits(self):
        return [(j[6], j[7]) for j in self.joints if j[4] == capi.JOINT_REVOLUTE]

    def to_device(self) -> capi.Robot:
        ee = self.link_names.index("end_effector") if "end_effector" in self.link_names else -1
        return capi.Robot(len(self.link_names), self.shapes, self.joints, self.find_link_by_name("flying_base"), ee)


def create_procedural_robot_model(total_arm_length: float = 0.75, joint_types=(HORIZONTAL,)) -> RobotModel:
    """The drone: base box 0.8 x 0.8 x 0.2, one stick box 0.05 x l x 0.05 per joint (l = arm length / joints),
    revolute joints about x (HORIZONTAL) or z (VERTICAL) limited to +-pi/2, fixed end effector at the arm's tip."""
    m = RobotModel()
    stick = total_arm_length / float(len(joint_types))
    attach_to = m.insert_link("flying_base", (0.8, 0.8, 0.2))
    attach_a = (0.0, 0.2, 0.0, 0.0, 0.0, 0.0, 1.0)
    for jt in joint_types:
        link = m.insert_link("stick", (0.05, stick, 0.05))
        stick_tf = (0.0, -stick / 2, 0.0, 0.0, 0.0, 0.0, 1.0)
        m.insert_joint(attach_to, link, attach_a, stick_tf, capi.JOINT_REVOLUTE,
                       (1.0, 0.0, 0.0) if jt == HORIZONTAL else (0.0, 0.0, 1.0), -np.pi / 2.0, np.pi / 2.0)
        # inverse of a pure translation (Transform.h:52-58 with an identity rotation): the negated vector,
        # produced as 0 - t so that signed zeros match the reference's quaternion sandwich.
        attach_a = (0.0, 0.0 - stick_tf[1], 0.0, -0.0, -0.0, -0.0, 1.0)
        attach_to = link
    ee = m.insert_link("end_effector")
    m.insert_joint(attach_to, ee, attach_a, capi.IDENTITY_TF, capi.JOINT_FIXED)
    return m


class Mt19937:
    """``random_numbers::RandomNumberGenerator``: std::mt19937 plus libstdc++'s uniform_real_distribution<double>
    (two 32-bit draws per double: (r1 + r2 * 2^32) / 2^64)."""

    def __init__(self, seed: int = 42):
        self.bitgen = np.random.MT19937()
        self.bitgen._legacy_seeding(seed)  # init_genrand, what std::mt19937(seed) does

    def canonical(self, n: int) -> np.ndarray:
        raw = self.bitgen.random_raw(2 * n).astype(np.float64)
        c = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
        return np.where(c >= 1.0, np.nextafter(1.0, 0.0), c)

    def uniform_real(self, lo, hi, n: int) -> np.ndarray:
        return self.canonical(n) * (hi - lo) + lo


def generate_uniform_random_states(model: RobotModel, rng: Mt19937, n: int, h_range=5.0, v_range=10.0,
                                   joint_angle_range=np.pi / 2.0) -> np.ndarray:
    """n states; per state the draws come in the reference's order: x, y, z, yaw, then one value per *joint*
    (fixed joints included)."""
    nj = len(model.joints)
    per = 4 + nj
    c = rng.canonical(n * per).reshape(n, per)
    out = np.zeros((n, 7 + nj))
    out[:, 0] = c[:, 0] * (2 * h_range) + -h_range
    out[:, 1] = c[:, 1] * (2 * h_range) + -h_range
    out[:, 2] = c[:, 2] * (v_range - 0) + 0
    yaw = c[:, 3] * (2 * np.pi - 0) + 0
    out[:, 5] = 1.0 * np.sin(yaw / 2.0)
    out[:, 6] = np.cos(yaw / 2.0)
    out[:, 7:] = c[:, 4:] * (2 * joint_angle_range) + -joint_angle_range
    return out


def goal_draws(model: RobotModel, rng: Mt19937, n: int) -> np.ndarray:
    """The draws genUprightState consumes, per sample: yaw ~ U(-pi, pi), then one value per revolute joint."""
    lim = model.revolute_limits()
    per = 1 + len(lim)
    c = rng.canonical(n * per).reshape(n, per)
    out = np.zeros((n, per))
    out[:, 0] = c[:, 0] * (np.pi - -np.pi) + -np.pi
    for k, (lo, hi) in enumerate(lim):
        out[:, 1 + k] = c[:, 1 + k] * (hi - lo) + lo
    return out
