// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE's own sources, compiled from where they
// lie under /root/reference/src by oracle/Makefile (target `ref`) into oracle/_ref/libmgodpl_ref.so.  The
// reference's planning-layer code runs unmodified: forwardKinematics, interpolate, equal_weights_distance,
// check_robot_collision, check_motion_collides, check_path_collides, generateUniformRandomState,
// genGoalStateUniform, fromEndEffectorAndVector, createProceduralRobotModel, meshToFclBVH, and its nearest-neighbour structure
// (NearestNeighborsGNAT.h + GreedyKCenters.h, header-only, standard library only).  Its two absent
// third-party dependencies (FCL, Eigen) are satisfied by oracle/shim/, which forwards to the oracle's restated
// narrow phase and slerp; so this library pins everything ABOVE the fcl::collide call, not FCL itself.
//
// Used (a) here in the build container to validate oracle/mgodpl_oracle.hpp and to generate tests/golden/*.json
// (tests/golden/make_golden.py), (b) on the GPU box — as a prebuilt .so — as a second checker in tests.
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "experiment_utils/procedural_robot_models.h"
#include "experiment_utils/mesh_from_dae.h"
#include "experiment_utils/mesh_connected_components.h"
#include "planning/Mesh.h"
#include "planning/RobotModel.h"
#include "planning/RobotPath.h"
#include "planning/RobotState.h"
#include "planning/collision_detection.h"
#include "planning/distance.h"
#include "planning/fcl_utils.h"
#include "planning/goal_sampling.h"
#include "planning/local_optimization.h"
#include "planning/nearest_neighbours/GreedyKCenters.h"
#include "planning/nearest_neighbours/NearestNeighborsGNAT.h"
#include "planning/state_tools.h"

#include <fcl/narrowphase/collision_object.h>
#include <fcl/geometry/bvh/BVH_model.h>

// procedural_robot_models.cpp:93 loads the drone's *visual* mesh through assimp (absent).  Collision geometry is
// procedural, so an empty visual mesh changes nothing on this path.
namespace mgodpl {
Mesh from_dae(const std::string &) { return {}; }
// local_optimization.cpp:117-136 defines this overload; the header only declares a (robot, path, obstacle) one.
bool tryDeletingEveryWaypoint(RobotPath &path, const std::function<bool(const RobotState &, const RobotState &)> &check_motion_collides);
}

using namespace mgodpl;

namespace {
RobotState load_state(const double *p, int js) {
	RobotState s;
	s.base_tf.translation = math::Vec3d(p[0], p[1], p[2]);
	s.base_tf.orientation = {p[3], p[4], p[5], p[6]};
	s.joint_values.assign(p + 7, p + 7 + js);
	return s;
}
void store_state(const RobotState &s, double *p) {
	p[0] = s.base_tf.translation.x(), p[1] = s.base_tf.translation.y(), p[2] = s.base_tf.translation.z();
	p[3] = s.base_tf.orientation.x, p[4] = s.base_tf.orientation.y, p[5] = s.base_tf.orientation.z, p[6] = s.base_tf.orientation.w;
	for (size_t i = 0; i < s.joint_values.size(); ++i) p[7 + i] = s.joint_values[i];
}
struct RefScene {
	fcl::CollisionObjectd object;
};
}  // namespace

extern "C" {

void *ref_robot_procedural(double arm_length, const int *joint_types, int n) {
	experiments::RobotArmParameters p{arm_length, {}, false};
	for (int i = 0; i < n; ++i) p.joint_types.push_back(joint_types[i] == 0 ? experiments::HORIZONTAL : experiments::VERTICAL);
	return new robot_model::RobotModel(experiments::createProceduralRobotModel(p));
}
void ref_robot_free(void *r) { delete (robot_model::RobotModel *) r; }
int ref_robot_n_links(void *r) { return (int) ((robot_model::RobotModel *) r)->getLinks().size(); }
int ref_robot_n_joints(void *r) { return (int) ((robot_model::RobotModel *) r)->getJoints().size(); }

void ref_fk(void *r, const double *state, int js, double *out_link_tfs7) {
	auto &robot = *(robot_model::RobotModel *) r;
	RobotState s = load_state(state, js);
	auto fk = robot_model::forwardKinematics(robot, s.joint_values, robot.findLinkByName("flying_base"), s.base_tf);
	for (size_t i = 0; i < fk.link_transforms.size(); ++i) {
		double *o = out_link_tfs7 + 7 * i;
		const auto &t = fk.link_transforms[i];
		o[0] = t.translation.x(), o[1] = t.translation.y(), o[2] = t.translation.z();
		o[3] = t.orientation.x, o[4] = t.orientation.y, o[5] = t.orientation.z, o[6] = t.orientation.w;
	}
}
double ref_distance(const double *a, const double *b, int js) {
	return equal_weights_distance(load_state(a, js), load_state(b, js));
}
void ref_interpolate(const double *a, const double *b, int js, double t, double *out) {
	store_state(interpolate(load_state(a, js), load_state(b, js), t), out);
}

void *ref_scene_create(const double *verts, size_t nv, const uint32_t *tris, size_t nt) {
	Mesh m;
	for (size_t i = 0; i < nv; ++i) m.vertices.emplace_back(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
	for (size_t i = 0; i < nt; ++i) m.triangles.push_back({tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]});
	return new RefScene{fcl::CollisionObjectd(fcl_utils::meshToFclBVH(m))};
}
void ref_scene_free(void *s) { delete (RefScene *) s; }

void ref_check_states(void *scene, void *robot, const double *states, size_t n, size_t stride, int js, uint8_t *hits) {
	auto &rb = *(robot_model::RobotModel *) robot;
	auto &sc = ((RefScene *) scene)->object;
	for (size_t i = 0; i < n; ++i) hits[i] = check_robot_collision(rb, sc, load_state(states + i * stride, js));
}
/// Degenerate motions (identical endpoints) are skipped by the caller: the reference evaluates 0/0 there.
void ref_check_motions(void *scene, void *robot, const double *a, const double *b, size_t m, size_t stride, int js,
					   uint8_t *hits, double *toi) {
	auto &rb = *(robot_model::RobotModel *) robot;
	auto &sc = ((RefScene *) scene)->object;
	for (size_t i = 0; i < m; ++i) {
		double t = std::numeric_limits<double>::quiet_NaN();
		hits[i] = check_motion_collides(rb, sc, load_state(a + i * stride, js), load_state(b + i * stride, js), t);
		toi[i] = hits[i] ? t : std::numeric_limits<double>::quiet_NaN();
	}
}
long ref_check_path(void *scene, void *robot, const double *states, size_t w, size_t stride, int js) {
	RobotPath path;
	for (size_t i = 0; i < w; ++i) path.append(load_state(states + i * stride, js));
	PathPoint pt{};
	pt.segment_i = (size_t) -1;
	bool hit = check_path_collides(*(robot_model::RobotModel *) robot, ((RefScene *) scene)->object, path, pt);
	return hit ? (long) pt.segment_i : -1;
}

void ref_gen_uniform_states(void *robot, unsigned seed, double h_range, double v_range, double joint_range, size_t n, double *out) {
	auto &rb = *(robot_model::RobotModel *) robot;
	random_numbers::RandomNumberGenerator rng(seed);
	size_t row = 7 + rb.getJoints().size();
	for (size_t i = 0; i < n; ++i) store_state(generateUniformRandomState(rb, rng, h_range, v_range, joint_range), out + i * row);
}
void ref_gen_goal_states(void *robot, unsigned seed, const double *target3, double distance_from_target, size_t n, double *out) {
	auto &rb = *(robot_model::RobotModel *) robot;
	random_numbers::RandomNumberGenerator rng(seed);
	auto fb = rb.findLinkByName("flying_base"), ee = rb.findLinkByName("end_effector");
	size_t row = 7 + rb.count_joint_variables();
	for (size_t i = 0; i < n; ++i)
		store_state(genGoalStateUniform(rng, math::Vec3d(target3[0], target3[1], target3[2]), distance_from_target, rb, fb, ee),
					out + i * row);
}
void ref_from_ee_and_vector(void *robot, const double *point3, const double *vec3, double *out8) {
	store_state(fromEndEffectorAndVector(*(robot_model::RobotModel *) robot, math::Vec3d(point3[0], point3[1], point3[2]),
										 math::Vec3d(vec3[0], vec3[1], vec3[2])), out8);
}

void ref_arm_vector(void *robot, const double *state, int js, double *out3) {
	math::Vec3d v = arm_vector_from_state(*(robot_model::RobotModel *) robot, load_state(state, js));
	out3[0] = v.x(), out3[1] = v.y(), out3[2] = v.z();
}
/// Same contract as orc_sample_goal_states, over the reference's own goal_sampling.cpp (generateUniformRandomArmVectorState,
/// findGoalStateByUniformSampling) and collision_detection.cpp.
void ref_sample_goal_states(void *scene, void *robot, int which, unsigned seed, const double *targets, size_t n, int max_attempts,
							double ee_distance, double *out, uint8_t *found, uint32_t *next_draw) {
	auto &rb = *(robot_model::RobotModel *) robot;
	auto &sc = *(RefScene *) scene;
	const auto fb = rb.findLinkByName("flying_base"), ee = rb.findLinkByName("end_effector");
	const size_t row = 7 + (which == 0 ? 1 : rb.count_joint_variables());
	for (size_t i = 0; i < n; ++i) {
		random_numbers::RandomNumberGenerator rng(seed + (unsigned) i);
		const math::Vec3d t(targets[3 * i], targets[3 * i + 1], targets[3 * i + 2]);
		std::optional<RobotState> s = which == 0 ? generateUniformRandomArmVectorState(rb, sc.object, t, rng, max_attempts, ee_distance)
												 : findGoalStateByUniformSampling(t, rb, fb, ee, sc.object, rng, (size_t) max_attempts);
		found[i] = s.has_value();
		for (size_t k = 0; k < row; ++k) out[i * row + k] = std::numeric_limits<double>::quiet_NaN();
		if (s) store_state(*s, out + i * row);
		next_draw[i] = (uint32_t) rng.uniformInteger(0, std::numeric_limits<int>::max());
	}
}

/// Same contract as orc_shortcut_path, over the reference's own local_optimization.cpp / RobotPath.cpp.
long ref_shortcut_path(void *scene, void *robot, const double *states, size_t w, size_t stride, int js, unsigned seed,
					   size_t iterations, int mode, double pull, double *out_states, size_t out_cap, size_t *out_w,
					   size_t *n_checks) {
	auto &rb = *(robot_model::RobotModel *) robot;
	auto &sc = ((RefScene *) scene)->object;
	RobotPath path;
	for (size_t i = 0; i < w; ++i) path.append(load_state(states + i * stride, js));
	random_numbers::RandomNumberGenerator rng(seed);
	size_t checks = 0;
	std::function<bool(const RobotState &, const RobotState &)> motion = [&](const RobotState &a, const RobotState &b) {
		++checks;
		return check_motion_collides(rb, sc, a, b);
	};
	long ok = 0;
	if (mode == 0) {
		for (size_t i = 0; i < iterations; ++i) {  // tsp_over_prm.cpp:594-612
			if (path.states.size() <= 2) break;
			ok += tryShortcuttingRandomlyGlobally(path, motion, rng);
		}
	} else if (mode == 1) ok = tryDeletingEveryWaypoint(path, motion);
	else if (mode == 2) {
		for (size_t i = 0; i < iterations && path.states.size() > 2; ++i) ok += tryShortcuttingRandomlyLocally(path, rng, motion);
	} else {
		for (size_t i = 1; i + 1 < path.states.size(); ++i) ok += tryMidpointPull(path, i, pull, motion);
	}
	for (size_t i = 0; i < path.states.size() && i < out_cap; ++i) store_state(path.states[i], out_states + i * stride);
	*out_w = path.states.size();
	if (n_checks) *n_checks = checks;
	return ok;
}

/// connected_vertex_components of the reference (mesh_connected_components.cpp:25-74), reported in canonical form:
/// labels[v] = smallest vertex index of v's component; returns the number of components.
size_t ref_mesh_components(const uint32_t *tris, size_t nt, size_t nv, uint32_t *labels) {
	Mesh mesh;
	mesh.vertices.resize(nv);
	for (size_t t = 0; t < nt; ++t) mesh.triangles.push_back({tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]});
	auto cc = connected_vertex_components(mesh);
	for (auto &c: cc) {
		size_t lo = *std::min_element(c.begin(), c.end());
		for (size_t v: c) labels[v] = (uint32_t) lo;
	}
	return cc.size();
}

/// The reference's OWN nearest-neighbour structure (nearest_neighbours/NearestNeighborsGNAT.h with greedy_k_centers pivots, set up
/// as init_empty_spatial_index does, tsp_over_prm.cpp:357-390) driven the way add_and_connect_roadmap_node drives it
/// (tsp_over_prm.cpp:33-66): for every state in turn, nearestK among the states added BEFORE it, then add it.  idx: n x k, -1
/// padded; dist: n x k (NaN padded), in the order nearestK returns them.
void ref_gnat_causal_knn(const double *states, size_t n, size_t stride, int js, int k, unsigned seed, int32_t *idx, double *dist) {
	using Point = std::pair<RobotState, size_t>;
	random_numbers::RandomNumberGenerator rng(seed);
	const std::function<double(const Point &, const Point &)> d = [](const Point &a, const Point &b) { return equal_weights_distance(a.first, b.first); };
	ompl::NearestNeighborsGNAT<Point>::SelectPivotFn pivots = [d, &rng](const std::vector<Point> &data, unsigned int kk) {
		return greedy_k_centers<Point>(data, kk, d, rng, std::nullopt);
	};
	ompl::NearestNeighborsGNAT<Point> gnat(pivots);
	gnat.setDistanceFunction(d);
	std::vector<Point> near;
	for (size_t i = 0; i < n; ++i) {
		const RobotState s = load_state(states + i * stride, js);
		gnat.nearestK({s, 0}, (size_t) k, near);
		for (int j = 0; j < k; ++j) {
			const bool have = (size_t) j < near.size();
			idx[i * k + j] = have ? (int32_t) near[j].second : -1;
			dist[i * k + j] = have ? equal_weights_distance(s, near[j].first) : std::nan("");
		}
		gnat.add({s, i});
	}
}

}  // extern "C"
