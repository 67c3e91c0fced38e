"""mgodpl_b200 — B200-native collision-query path of the Multigoal Orchard Drone Planning Library.

The product is ``libmgodpl_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI in ``include/mgodpl_b200.h``)
plus the C++ facade in ``host/`` that restores the reference's planning-layer signatures.  This Python package is
the thin binding used by ``bench.py`` and the test-suite:

    capi     ctypes binding of the C ABI (Scene, Robot, check_states, check_motions, sample_goals, ...)
    robots   createProceduralRobotModel + the reference's state samplers (std::mt19937-compatible)
    scenes   procedural fruit trees / orchard rows (the reference's meshes are an un-fetched LFS blob)
    sharding contiguous 32-aligned batch shards for one-process-per-GPU runs

The directory name contains hyphens, so import it with ``importlib.import_module`` (see ``mgodpl_b200.py`` at the
repository root for a ready-made alias).
"""
from . import capi, robots, scenes, sharding  # noqa: F401
from .capi import (MgodplError, Robot, Scene, check_boxes, check_motions, check_path, check_states,  # noqa: F401
                   forward_kinematics, sample_goals)
