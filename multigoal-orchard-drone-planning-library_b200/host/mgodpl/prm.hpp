// Roadmap construction of plan_path_tsp_over_prm, with the collision checks and nearest-neighbour queries batched:
//   GoalSampleParams, TspOverPrmParameters                src/planning/tsp_over_prm.h:62-81
//   add_and_connect_roadmap_node                          src/planning/tsp_over_prm.cpp:33-66
//   sample_and_connect_infrastucture_node                 src/planning/tsp_over_prm.cpp:244-281
//   sample_and_connect_goal_states                        src/planning/tsp_over_prm.cpp:297-344
//   build_infrastructure_roadmap                          src/planning/tsp_over_prm.cpp:411-429
//   sample_goal_states                                    src/planning/tsp_over_prm.cpp:447-483
//   construct_final_path                                  src/planning/tsp_over_prm.cpp:189-234
// The reference grows the roadmap one node at a time: sample, check the state, ask the GNAT for the k nearest
// *infrastructure* nodes, check k motions, insert.  None of the random draws depends on a collision result (except the
// early stop of goal sampling, handled by drawing ahead on a copy of the generator), and the k nearest nodes of node i
// among the nodes inserted before it are a property of the node set — so the same graph, vertex for vertex and edge for
// edge, comes out of a handful of batched GPU calls.  The GNAT is replaced by the exact k-NN kernels (same answers, ties
// aside); the nodes that take part in nearest-neighbour queries are the first `n_infrastructure` vertices of the graph.
#pragma once
#include "local_optimization.hpp"
#include "roadmap.hpp"
#include "sampling.hpp"

namespace mgodpl {

/// tsp_over_prm.h:62-69.
struct GoalSampleParams {
	size_t k_neighbors = 5;
	size_t max_valid_samples = 2;
	size_t max_attempts = 100;
};
/// tsp_over_prm.h:71-81.
struct TspOverPrmParameters {
	size_t n_neighbours = 5;
	size_t max_samples = 100;
	GoalSampleParams goal_sample_params;
};

/// The roadmap plus the number of leading vertices that are infrastructure nodes (the contents of the reference's GNAT).
struct PRM {
	PRMGraph graph;
	size_t n_infrastructure = 0;
};

/// k nearest infrastructure nodes of each query state (NearestNeighborsGNAT::nearestK on an index the queries are not in).
inline NearestNeighbours nearest_infrastructure_nodes(const PRM &prm, const RobotState *queries, size_t n_queries, size_t k, void *stream = nullptr) {
	NearestNeighbours r{k, std::vector<int32_t>(n_queries * k, -1), std::vector<double>(n_queries * k)};
	if (!n_queries) return r;
	const size_t js = gpu::common_joint_values(queries, n_queries), stride = 7 + js, n_db = prm.n_infrastructure;
	if (n_db && gpu::common_joint_values(prm.graph.states.data(), n_db) != js)
		throw std::runtime_error("states of one batch must carry the same number of joint values");
	std::vector<double> db(n_db * stride), q(n_queries * stride);
	for (size_t i = 0; i < n_db; ++i) gpu::pack_state(prm.graph.states[i], js, db.data() + i * stride);
	for (size_t i = 0; i < n_queries; ++i) gpu::pack_state(queries[i], js, q.data() + i * stride);
	gpu::throw_status(mgodpl_knn_query(db.data(), n_db, q.data(), n_queries, stride, (int32_t) js, (int32_t) k, r.index.data(), r.distance.data(), stream));
	return r;
}

/// add_and_connect_roadmap_node for a batch of states that are not infrastructure nodes (goal samples, the start state):
/// each becomes a vertex and is linked to those of its k nearest infrastructure nodes it can reach by a free motion
/// (motion_collides(neighbour, state), edge weight = equal_weights_distance).  Returns the new vertex ids.
///
/// Joint-vector lengths: infrastructure nodes come from generateUniformRandomState (one value per joint, fixed joints
/// included) while goal samples come from genUprightState (one per revolute joint).  The reference mixes the two and
/// its distance / interpolation loops then read past the end of the shorter vector (distance.cpp:20-22,
/// RobotState.cpp:19-21: undefined behaviour).  Here a shorter state is padded with zeros to the infrastructure length
/// — the values FK never looks at — so the result is defined.
inline std::vector<PRMGraph::vertex_descriptor> add_and_connect_roadmap_nodes(std::vector<RobotState> states, PRM &prm, size_t k_neighbors,
																			  const robot_model::RobotModel &robot, const gpu::Scene &scene) {
	std::vector<PRMGraph::vertex_descriptor> ids;
	if (states.empty()) return ids;
	if (prm.n_infrastructure)
		for (auto &s: states)
			if (s.joint_values.size() < prm.graph.states[0].joint_values.size()) s.joint_values.resize(prm.graph.states[0].joint_values.size(), 0.0);
	const NearestNeighbours nn = nearest_infrastructure_nodes(prm, states.data(), states.size(), k_neighbors);
	std::vector<RobotState> from, to;
	for (size_t i = 0; i < states.size(); ++i)
		for (size_t q = 0; q < k_neighbors; ++q)
			if (nn.index[i * k_neighbors + q] >= 0) from.push_back(prm.graph.states[(size_t) nn.index[i * k_neighbors + q]]), to.push_back(states[i]);
	MotionResults r;
	if (!from.empty()) r = check_motions(robot, scene, from.data(), to.data(), from.size());
	size_t e = 0;
	for (size_t i = 0; i < states.size(); ++i) {
		const auto v = prm.graph.add_vertex(states[i]);
		ids.push_back(v);
		for (size_t q = 0; q < k_neighbors; ++q) {
			const int32_t j = nn.index[i * k_neighbors + q];
			if (j < 0) continue;
			if (!bit(r.hit_bits, e)) prm.graph.add_edge(v, (size_t) j, nn.distance[i * k_neighbors + q]);
			++e;
		}
	}
	return ids;
}
/// tsp_over_prm.cpp:33-66 for one state.
inline PRMGraph::vertex_descriptor add_and_connect_roadmap_node(const RobotState &state, PRM &prm, size_t k_neighbors,
																const robot_model::RobotModel &robot, const gpu::Scene &scene) {
	return add_and_connect_roadmap_nodes({state}, prm, k_neighbors, robot, scene)[0];
}

/// build_infrastructure_roadmap (tsp_over_prm.cpp:411-429) with the sampler plan_path_tsp_over_prm passes in
/// (generateUniformRandomState(robot, rng, 5.0, 10.0), :569-571): max_samples draws, the collision-free ones become
/// infrastructure nodes in draw order, each linked to its k nearest predecessors by free motions.  Three GPU calls.
inline PRM build_infrastructure_roadmap(const robot_model::RobotModel &robot, const gpu::Scene &scene, random_numbers::RandomNumberGenerator &rng,
										size_t max_samples, size_t n_neighbours, double h_range = 5.0, double v_range = 10.0) {
	std::vector<RobotState> drawn;
	drawn.reserve(max_samples);
	for (size_t i = 0; i < max_samples; ++i) drawn.push_back(generateUniformRandomState(robot, rng, h_range, v_range));
	const std::vector<uint32_t> collides = drawn.empty() ? std::vector<uint32_t>() : check_states(robot, scene, drawn);
	std::vector<RobotState> nodes;
	for (size_t i = 0; i < drawn.size(); ++i)
		if (!bit(collides, i)) nodes.push_back(std::move(drawn[i]));
	PRM prm;
	prm.n_infrastructure = nodes.size();
	const std::vector<RoadmapEdge> edges = validate_prm_edges(robot, scene, nodes, n_neighbours);
	prm.graph.states = std::move(nodes);
	for (const auto &e: edges) prm.graph.add_edge(e.to, e.from, e.weight);  // (new vertex, neighbour), as boost::add_edge is called
	return prm;
}

/// sample_goal_states (tsp_over_prm.cpp:447-483) over sample_and_connect_goal_states (:297-344): per fruit up to
/// max_attempts draws of genGoalStateUniform until max_valid_samples are collision-free; the valid ones become vertices
/// linked to their k nearest infrastructure nodes.  The draws of one fruit are made ahead on a copy of the generator and
/// checked as one batch; `rng` is then advanced by exactly the attempts the sequential loop would have made.  All
/// fruits' valid samples are connected in one k-NN call and one motion batch.  Returns {goal_nodes, group_sizes}.
inline std::pair<std::vector<PRMGraph::vertex_descriptor>, std::vector<size_t>> sample_goal_states(
	const std::vector<math::Vec3d> &fruit_positions, PRM &prm, const GoalSampleParams &params, const robot_model::RobotModel &robot,
	const gpu::Scene &scene, random_numbers::RandomNumberGenerator &rng) {
	const auto base_link = robot.findLinkByName("flying_base"), end_effector_link = robot.findLinkByName("end_effector");
	std::vector<RobotState> valid;
	std::vector<size_t> group_sizes(fruit_positions.size(), 0);
	for (size_t g = 0; g < fruit_positions.size(); ++g) {
		if (!params.max_attempts || !params.max_valid_samples) continue;
		random_numbers::RandomNumberGenerator ahead = rng;
		std::vector<RobotState> attempts;
		for (size_t a = 0; a < params.max_attempts; ++a) attempts.push_back(genGoalStateUniform(ahead, fruit_positions[g], robot, base_link, end_effector_link));
		const std::vector<uint32_t> collides = check_states(robot, scene, attempts);
		size_t used = 0;
		while (used < params.max_attempts && group_sizes[g] < params.max_valid_samples) {
			if (!bit(collides, used)) valid.push_back(attempts[used]), ++group_sizes[g];
			++used;
		}
		for (size_t a = 0; a < used; ++a) (void) genGoalStateUniform(rng, fruit_positions[g], robot, base_link, end_effector_link);
	}
	return {add_and_connect_roadmap_nodes(valid, prm, params.k_neighbors, robot, scene), group_sizes};
}

/// construct_final_path (tsp_over_prm.cpp:189-234): the path from the start node through the goal samples of `tour` (indices
/// into goal_nodes, in visiting order — the reference gets it from its OR-Tools TSP, pick_visitation_order :156-175, which is
/// outside this library), each leg retraced through the predecessor tables and handed to `optimize_segment`.
inline RobotPath construct_final_path(const PRMGraph &graph, const std::vector<std::vector<PRMGraph::vertex_descriptor>> &predecessor_lookup,
									  const std::vector<PRMGraph::vertex_descriptor> &start_to_goals_predecessors,
									  const std::vector<PRMGraph::vertex_descriptor> &goal_nodes, const std::vector<size_t> &tour,
									  const std::function<bool(RobotPath &)> &optimize_segment) {
	RobotPath path;
	if (tour.empty()) return path;
	RobotPath initial_path = retrace_path(graph, start_to_goals_predecessors, goal_nodes[tour[0]]);
	optimize_segment(initial_path);
	path.append(initial_path);
	for (size_t i = 1; i < tour.size(); ++i) {
		RobotPath goal_to_goal_path = retrace_path(graph, predecessor_lookup[tour[i - 1]], goal_nodes[tour[i]]);
		optimize_segment(goal_to_goal_path);
		path.append(goal_to_goal_path);
	}
	return path;
}

/// The `optimize_path_segment` closure of plan_path_tsp_over_prm (tsp_over_prm.cpp:594-612) on the batched shortcutting
/// driver: 100 x tryShortcuttingRandomlyGlobally per leg, same path and generator state, a few GPU calls.
inline std::function<bool(RobotPath &)> optimize_path_segment_fn(const robot_model::RobotModel &robot, const gpu::Scene &scene,
																  random_numbers::RandomNumberGenerator &rng, size_t iterations = 100) {
	return [&robot, &scene, &rng, iterations](RobotPath &path) {
		if (path.states.size() > 2) shortcut_path_globally(robot, scene, path, rng, iterations);
		return false;  // the reference's lambda never sets its `successful` flag either (tsp_over_prm.cpp:595-611)
	};
}

}  // namespace mgodpl
