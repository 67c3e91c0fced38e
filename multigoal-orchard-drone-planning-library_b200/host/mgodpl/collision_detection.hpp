// The reference's planning-layer collision API (src/planning/collision_detection.h:29-111 and
// collision_detection.cppm:19-119) over the CUDA library, header-only C++17.
//
// Drop-in rule: every signature keeps the reference's name and argument order; `mgodpl::gpu::Scene` stands in the slot
// where `fcl::CollisionObjectd` stood, and `gpu::meshToScene` / `gpu::treeMeshesToScene` replace
// `fcl_utils::meshToFclBVH` / `treeMeshesToFclCollisionObject` (src/planning/fcl_utils.h:33,41).  Single queries are
// batches of one; the batched entry points (check_states, check_motions, sample_goal_states) are the fast path.
// Errors surface as std::runtime_error with the reference's messages; there is no CPU fallback.
#pragma once
#include <cstdint>
#include <functional>
#include <optional>

#include "robot_model.hpp"

namespace mgodpl {

namespace gpu {

/// Device-resident BVH over a triangle mesh (identity pose), shared by copies, safe for concurrent queries.
class Scene {
	struct Holder {
		mgodpl_scene *h = nullptr;
		~Holder() { mgodpl_scene_destroy(h); }
	};
	std::shared_ptr<Holder> holder_;

public:
	Scene() = default;
	explicit Scene(const Mesh &mesh, const mgodpl_scene_opts *opts = nullptr) : holder_(std::make_shared<Holder>()) {
		std::vector<double> v;
		v.reserve(mesh.vertices.size() * 3);
		for (auto &p: mesh.vertices) v.insert(v.end(), {p.x(), p.y(), p.z()});
		static_assert(sizeof(size_t) == sizeof(uint64_t), "Mesh::triangles is passed through as uint64");
		throw_status(mgodpl_scene_create_u64(v.data(), mesh.vertices.size(), reinterpret_cast<const uint64_t *>(mesh.triangles.data()),
											 mesh.triangles.size(), opts, &holder_->h));
	}
	mgodpl_scene *handle() const {
		if (!holder_) throw std::runtime_error("mgodpl::gpu::Scene is empty");
		return holder_->h;
	}
	mgodpl_scene_info info() const {
		mgodpl_scene_info i{};
		throw_status(mgodpl_scene_get_info(handle(), &i));
		return i;
	}
};

/// Replaces fcl_utils::meshToFclBVH(const Mesh&).
inline Scene meshToScene(const Mesh &mesh) { return Scene(mesh); }

/// Mirror of tree_meshes::TreeMeshes (src/experiment_utils/TreeMeshes.h:19-24).
struct TreeMeshes {
	std::string tree_name;
	Mesh leaves_mesh, trunk_mesh;
	std::vector<Mesh> fruit_meshes;
};
/// Replaces fcl_utils::treeMeshesToFclCollisionObject: like the reference, only the trunk mesh collides (fcl_utils.cpp:58-60).
inline Scene treeMeshesToScene(const TreeMeshes &meshes) { return Scene(meshes.trunk_mesh); }

inline void pack_state(const RobotState &s, size_t js, double *o) {
	to7(s.base_tf, o);
	for (size_t i = 0; i < js; ++i) o[7 + i] = i < s.joint_values.size() ? s.joint_values[i] : 0.0;
}
inline size_t common_joint_values(const RobotState *s, size_t n) {
	size_t js = n ? s[0].joint_values.size() : 0;
	for (size_t i = 1; i < n; ++i)
		if (s[i].joint_values.size() != js) throw std::runtime_error("states of one batch must carry the same number of joint values");
	return js;
}
}  // namespace gpu

// ---- batched variants (the fast path) ----------------------------------------------------------------------------------
/// N x check_robot_collision.  Returns packed bits: state i -> bit (i & 31) of word (i >> 5).
inline std::vector<uint32_t> check_states(const robot_model::RobotModel &robot, const gpu::Scene &scene, const RobotState *states,
										  size_t n, void *stream = nullptr) {
	const size_t js = gpu::common_joint_values(states, n), stride = 7 + js;
	std::vector<double> buf(n * stride);
	for (size_t i = 0; i < n; ++i) gpu::pack_state(states[i], js, buf.data() + i * stride);
	std::vector<uint32_t> bits((n + 31) / 32);
	gpu::throw_status(mgodpl_check_states(scene.handle(), robot.device_handle(), buf.data(), n, stride, (int32_t) js, bits.data(), stream));
	return bits;
}
inline std::vector<uint32_t> check_states(const robot_model::RobotModel &robot, const gpu::Scene &scene, const std::vector<RobotState> &states) {
	return check_states(robot, scene, states.data(), states.size());
}
inline bool bit(const std::vector<uint32_t> &bits, size_t i) { return (bits[i >> 5] >> (i & 31)) & 1u; }

struct MotionResults {
	std::vector<uint32_t> hit_bits;
	std::vector<double> toi;        // first colliding interpolation parameter; NaN when free
	std::vector<uint32_t> n_steps;  // ceil(distance / max_step) per motion
};
/// M x check_motion_collides (from[i] -> to[i]).  max_step defaults to the reference's hard-coded 0.1.
inline MotionResults check_motions(const robot_model::RobotModel &robot, const gpu::Scene &scene, const RobotState *from,
								   const RobotState *to, size_t m, double max_step = 0.1, void *stream = nullptr) {
	const size_t js = gpu::common_joint_values(from, m), stride = 7 + js;
	if (gpu::common_joint_values(to, m) != js) throw std::runtime_error("states of one batch must carry the same number of joint values");
	std::vector<double> a(m * stride), b(m * stride);
	for (size_t i = 0; i < m; ++i) gpu::pack_state(from[i], js, a.data() + i * stride), gpu::pack_state(to[i], js, b.data() + i * stride);
	MotionResults r{std::vector<uint32_t>((m + 31) / 32), std::vector<double>(m), std::vector<uint32_t>(m)};
	gpu::throw_status(mgodpl_check_motions(scene.handle(), robot.device_handle(), a.data(), b.data(), m, stride, (int32_t) js, max_step,
										   r.hit_bits.data(), r.toi.data(), r.n_steps.data(), stream));
	return r;
}

/// k nearest neighbours of every state among the states that precede it (causal = true: the roadmap as it stood when
/// the reference inserted that node, tsp_over_prm.cpp:33-66) or among all other states, under equal_weights_distance —
/// what NearestNeighborsGNAT::nearestK answers (tsp_over_prm.cpp:44-45, 383-386), for the whole batch at once.
struct NearestNeighbours {
	size_t k = 0;
	std::vector<int32_t> index;    // n x k, ascending by distance, -1 pads
	std::vector<double> distance;  // n x k, NaN pads
};
inline NearestNeighbours nearest_k(const RobotState *states, size_t n, size_t k, bool causal = true, void *stream = nullptr) {
	const size_t js = gpu::common_joint_values(states, n), stride = 7 + js;
	std::vector<double> buf(n * stride);
	for (size_t i = 0; i < n; ++i) gpu::pack_state(states[i], js, buf.data() + i * stride);
	NearestNeighbours r{k, std::vector<int32_t>(n * k), std::vector<double>(n * k)};
	gpu::throw_status(mgodpl_knn(buf.data(), n, stride, (int32_t) js, (int32_t) k, causal ? 1 : 0, r.index.data(), r.distance.data(), stream));
	return r;
}

/// PRM edge validation in one batch (add_and_connect_roadmap_node, tsp_over_prm.cpp:33-66, for every node at once):
/// for node i and each of its k nearest earlier nodes j, motion_collides(nodes[j], nodes[i]); collision-free pairs
/// become edges weighted by equal_weights_distance.
struct RoadmapEdge {
	size_t from, to;
	double weight;
};
inline std::vector<RoadmapEdge> validate_prm_edges(const robot_model::RobotModel &robot, const gpu::Scene &scene,
												   const std::vector<RobotState> &nodes, size_t k) {
	NearestNeighbours nn = nearest_k(nodes.data(), nodes.size(), k, true);
	std::vector<RobotState> a, b;
	std::vector<RoadmapEdge> cand;
	for (size_t i = 0; i < nodes.size(); ++i)
		for (size_t q = 0; q < k; ++q) {
			const int32_t j = nn.index[i * k + q];
			if (j < 0) continue;
			a.push_back(nodes[(size_t) j]), b.push_back(nodes[i]);
			cand.push_back({(size_t) j, i, nn.distance[i * k + q]});
		}
	if (cand.empty()) return {};
	MotionResults r = check_motions(robot, scene, a.data(), b.data(), a.size());
	std::vector<RoadmapEdge> edges;
	for (size_t e = 0; e < cand.size(); ++e)
		if (!bit(r.hit_bits, e)) edges.push_back(cand[e]);
	return edges;
}

// ---- the reference's free functions --------------------------------------------------------------------------------------
/// collision_detection.h:29-31.  Throws "Only boxes are implemented for collision geometry." for non-box shapes.
inline bool check_link_collision(const robot_model::RobotModel::Link &link, const gpu::Scene &tree_trunk_object, const math::Transformd &link_tf) {
	std::vector<double> boxes;
	for (const auto &g: link.collision_geometry) {
		const Box *box = std::get_if<Box>(&g.shape);
		if (!box) throw std::runtime_error("Only boxes are implemented for collision geometry.");
		double row[10];
		gpu::to7(link_tf.then(g.transform), row);
		row[7] = box->size.x(), row[8] = box->size.y(), row[9] = box->size.z();
		boxes.insert(boxes.end(), row, row + 10);
	}
	const size_t n = boxes.size() / 10;
	if (!n) return false;
	std::vector<uint32_t> bits((n + 31) / 32);
	gpu::throw_status(mgodpl_check_boxes(tree_trunk_object.handle(), boxes.data(), n, bits.data(), nullptr));
	for (uint32_t w: bits)
		if (w) return true;
	return false;
}

/// collision_detection.h:41-43.
inline bool check_robot_collision(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object, const RobotState &state) {
	return check_states(robot, tree_trunk_object, &state, 1)[0] & 1u;
}

/// collision_detection.h:59-63: toi receives the first colliding interpolation parameter (undefined if free).
inline bool check_motion_collides(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object, const RobotState &state1,
								  const RobotState &state2, double &toi) {
	MotionResults r = check_motions(robot, tree_trunk_object, &state1, &state2, 1);
	if (r.hit_bits[0] & 1u) {
		toi = r.toi[0];
		return true;
	}
	return false;
}
inline bool check_motion_collides(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object, const RobotState &state1,
								  const RobotState &state2) {
	double dummy_toi;
	return check_motion_collides(robot, tree_trunk_object, state1, state2, dummy_toi);
}

/// collision_detection.h:93-96: sets only collision_point.segment_i, like the reference (collision_detection.cpp:115-118).
/// All segments are validated in one batch.
inline bool check_path_collides(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object, const RobotPath &path,
								PathPoint &collision_point) {
	if (path.states.size() < 2) return false;
	const size_t m = path.states.size() - 1;
	MotionResults r = check_motions(robot, tree_trunk_object, path.states.data(), path.states.data() + 1, m);
	for (size_t i = 0; i < m; ++i)
		if (bit(r.hit_bits, i)) {
			collision_point.segment_i = i;
			return true;
		}
	return false;
}
inline bool check_path_collides(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object, const RobotPath &path) {
	PathPoint dummy{};
	return check_path_collides(robot, tree_trunk_object, path, dummy);
}

// ---- the functional plug API the planners consume (collision_detection.cppm:19-119) -------------------------------------
struct CollisionEnvironment {
	const robot_model::RobotModel &robot;
	const gpu::Scene &tree_collision;
};
using CollisionDetectionFn = std::function<bool(const RobotState &)>;
using MotionCollisionDetectionFn = std::function<bool(const RobotState &, const RobotState &)>;

/// Note (as in the reference): the closures keep a reference to the environment, which must outlive them.
inline CollisionDetectionFn collision_check_fn_in_environment(const CollisionEnvironment &env) {
	return [&](const RobotState &from) { return check_robot_collision(env.robot, env.tree_collision, from); };
}
inline MotionCollisionDetectionFn motion_collision_check_fn_in_environment(const CollisionEnvironment &env) {
	return [&](const RobotState &from, const RobotState &to) { return check_motion_collides(env.robot, env.tree_collision, from, to); };
}
/// Sampled motion check over an arbitrary state predicate (collision_detection.cppm:57-75), host-side loop.
template<typename StateCollisionFn>
MotionCollisionDetectionFn motion_collision_check_fn_from_state_collision(StateCollisionFn state_collision_fn) {
	return [state_collision_fn](const RobotState &from, const RobotState &to) {
		const double distance = equal_weights_distance(from, to), MAX_STEP = 0.1;
		const size_t n_steps = (size_t) std::ceil(distance / MAX_STEP);
		for (size_t step_i = 0; step_i <= n_steps; ++step_i) {
			const double t = n_steps ? (double) step_i / (double) n_steps : 0.0;
			if (state_collision_fn(interpolate(from, to, t))) return true;
		}
		return false;
	};
}
struct CollisionFunctions {
	CollisionDetectionFn state_collides;
	MotionCollisionDetectionFn motion_collides;
};
inline CollisionFunctions collision_functions_in_environment(const CollisionEnvironment &env) {
	return {collision_check_fn_in_environment(env), motion_collision_check_fn_in_environment(env)};
}
using calls_t = size_t;
struct CollisionFunctionsCounts {
	calls_t &collision_check_invocations;
	calls_t &motion_collision_check_invocations;
};
/// Invocation-counting decorators (functional_utils.cppm:27-32).
inline CollisionFunctions collision_functions_in_environment_counting(const CollisionFunctions &f, CollisionFunctionsCounts counts) {
	return {[fn = f.state_collides, &c = counts.collision_check_invocations](const RobotState &s) {
				++c;
				return fn(s);
			},
			[fn = f.motion_collides, &c = counts.motion_collision_check_invocations](const RobotState &a, const RobotState &b) {
				++c;
				return fn(a, b);
			}};
}

}  // namespace mgodpl
