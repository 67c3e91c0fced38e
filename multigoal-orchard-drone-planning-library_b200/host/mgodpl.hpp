// Umbrella header of the host-side mirror: the reference's planning-layer API for the collision-query path.
#pragma once
#include "mgodpl/collision_detection.hpp"
#include "mgodpl/local_optimization.hpp"
#include "mgodpl/math.hpp"
#include "mgodpl/multi_gpu.hpp"
#include "mgodpl/prm.hpp"
#include "mgodpl/roadmap.hpp"
#include "mgodpl/robot_model.hpp"
#include "mgodpl/rrt.hpp"
#include "mgodpl/sampling.hpp"
#include "mgodpl/scanning_motions.hpp"
#include "mgodpl/tree_meshes.hpp"
