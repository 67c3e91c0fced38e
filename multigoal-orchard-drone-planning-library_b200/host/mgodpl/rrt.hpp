// Rapidly-exploring random tree — a planner that consumes the collision path through the two closures of
// CollisionFunctions (src/planning/rrt.cppm:82-89):
//   RRTNode, linear_nearest, RRTCallbackFn, rrt, retrace, rrt_path_to_acceptable     src/planning/rrt.cppm:22-188
//   DistanceFn                                                                      src/planning/distance.h:18
// `rrt` keeps the reference's signature and works with any closures (each call is then a GPU batch of one).
// `rrt_batched` grows the same tree — node for node, and leaves the generator in the same state — with two GPU calls per
// window of samples: the samples of a window are drawn ahead on a copy of the generator, their state checks do not depend
// on the tree at all, and their motion checks are issued against the nearest node of the tree as it stands at the start
// of the window.  While the window is committed in order, a candidate whose nearest node has been overtaken by a node
// inserted earlier in the same window is re-checked on its own; everything else was already answered.
#pragma once
#include <functional>
#include <optional>

#include "sampling.hpp"

namespace mgodpl {

using DistanceFn = std::function<double(const RobotState &, const RobotState &)>;

/// rrt.cppm:22-25.
struct RRTNode {
	RobotState state;
	size_t parent_index;
};

/// rrt.cppm:35-49: first node at the smallest distance.
inline size_t linear_nearest(const std::vector<RRTNode> &nodes, const RobotState &new_state, const DistanceFn &distance) {
	assert(!nodes.empty());
	size_t nearest_index = 0;
	double nearest_distance = distance(new_state, nodes[0].state);
	for (size_t i = 1; i < nodes.size(); ++i) {
		const double d = distance(new_state, nodes[i].state);
		if (d < nearest_distance) nearest_distance = d, nearest_index = i;
	}
	return nearest_index;
}

/// rrt.cppm:61: called after every expansion with the tree (new node last); return true to stop.
using RRTCallbackFn = std::function<bool(const std::vector<RRTNode> &nodes)>;

/// rrt.cppm:82-126.
inline void rrt(const RobotState &root, const std::function<RobotState()> &sample_state, const std::function<bool(const RobotState &)> &collides,
				const std::function<bool(const RobotState &, const RobotState &)> &motion_collides, const DistanceFn &distance,
				size_t max_iterations, const RRTCallbackFn &tree_expanded) {
	std::vector<RRTNode> nodes{{root, 0}};
	for (size_t iteration = 0; iteration < max_iterations; ++iteration) {
		RobotState new_state = sample_state();
		if (collides(new_state)) continue;
		const size_t nearest_index = linear_nearest(nodes, new_state, distance);
		if (motion_collides(nodes[nearest_index].state, new_state)) continue;
		nodes.push_back({std::move(new_state), nearest_index});
		if (tree_expanded(nodes)) break;
	}
}

/// rrt.cppm:139-151: the states from the last node back to the root (in that order).
inline RobotPath retrace(const std::vector<RRTNode> &nodes) {
	RobotPath path;
	size_t current_index = nodes.size() - 1;
	do {
		path.states.push_back(nodes[current_index].state);
		current_index = nodes[current_index].parent_index;
	} while (current_index != 0);
	path.states.push_back(nodes[0].state);
	return path;
}

/// rrt.cppm:166-188.
inline std::optional<RobotPath> rrt_path_to_acceptable(const RobotState &root, const std::function<RobotState()> &sample_state,
													   const std::function<bool(const RobotState &)> &collides,
													   const std::function<bool(const RobotState &, const RobotState &)> &motion_collides,
													   const DistanceFn &distance, size_t max_iterations,
													   const std::function<bool(const RobotState &)> &state_acceptable) {
	std::optional<RobotPath> path;
	rrt(root, sample_state, collides, motion_collides, distance, max_iterations, [&](const std::vector<RRTNode> &nodes) {
		if (!state_acceptable(nodes.back().state)) return false;
		path = retrace(nodes);
		return true;
	});
	return path;
}

struct RrtStats {
	size_t iterations = 0;  // samples the sequential loop would have drawn
	size_t gpu_calls = 0;
	size_t respeculated = 0;  // candidates whose nearest node changed inside their window
};

/// The same tree as rrt(root, [&] { return sample_state(rng); }, check_robot_collision, check_motion_collides,
/// equal_weights_distance, ...), built a window of samples at a time (see the header comment).  Returns the nodes.
inline std::vector<RRTNode> rrt_batched(const RobotState &root, const std::function<RobotState(random_numbers::RandomNumberGenerator &)> &sample_state,
										random_numbers::RandomNumberGenerator &rng, const robot_model::RobotModel &robot, const gpu::Scene &scene,
										size_t max_iterations, const RRTCallbackFn &tree_expanded, size_t window = 64, RrtStats *stats_out = nullptr) {
	const DistanceFn distance = [](const RobotState &a, const RobotState &b) { return equal_weights_distance(a, b); };
	std::vector<RRTNode> nodes{{root, 0}};
	RrtStats stats;
	bool stop = false;
	while (!stop && stats.iterations < max_iterations) {
		// a node inserted inside a window overtakes a later candidate's nearest node with probability ~ 1 / tree size each:
		// keep the window a fraction of the tree so most speculated motion checks stay valid
		const size_t w = std::min(std::min(window ? window : 1, std::max<size_t>(4, nodes.size() / 2)), max_iterations - stats.iterations);
		random_numbers::RandomNumberGenerator ahead = rng;
		std::vector<RobotState> candidates;
		for (size_t j = 0; j < w; ++j) candidates.push_back(sample_state(ahead));
		const std::vector<uint32_t> state_hit = check_states(robot, scene, candidates);
		++stats.gpu_calls;
		// speculate: nearest node of the tree as it stands now
		const size_t tree_at_start = nodes.size();
		std::vector<size_t> nearest(w, 0), slot(w, SIZE_MAX);
		std::vector<double> nearest_d(w, 0.0);
		std::vector<RobotState> from, to;
		for (size_t j = 0; j < w; ++j) {
			if (bit(state_hit, j)) continue;
			nearest[j] = linear_nearest(nodes, candidates[j], distance);
			nearest_d[j] = distance(candidates[j], nodes[nearest[j]].state);
			slot[j] = from.size();
			from.push_back(nodes[nearest[j]].state), to.push_back(candidates[j]);
		}
		MotionResults motions;
		if (!from.empty()) motions = check_motions(robot, scene, from.data(), to.data(), from.size()), ++stats.gpu_calls;
		// commit in order
		size_t consumed = 0;
		for (size_t j = 0; j < w && !stop; ++j) {
			consumed = j + 1;
			if (bit(state_hit, j)) continue;
			bool overtaken = false;
			for (size_t n = tree_at_start; n < nodes.size(); ++n) {
				const double d = distance(candidates[j], nodes[n].state);
				if (d < nearest_d[j]) nearest_d[j] = d, nearest[j] = n, overtaken = true;  // strictly nearer, as linear_nearest decides
			}
			bool collides;
			if (overtaken) {
				collides = check_motion_collides(robot, scene, nodes[nearest[j]].state, candidates[j]);
				++stats.gpu_calls, ++stats.respeculated;
			} else {
				collides = bit(motions.hit_bits, slot[j]);
			}
			if (collides) continue;
			nodes.push_back({candidates[j], nearest[j]});
			stop = tree_expanded(nodes);
		}
		for (size_t j = 0; j < consumed; ++j) (void) sample_state(rng);  // the draws the sequential loop made
		stats.iterations += consumed;
	}
	if (stats_out) *stats_out = stats;
	return nodes;
}

}  // namespace mgodpl
