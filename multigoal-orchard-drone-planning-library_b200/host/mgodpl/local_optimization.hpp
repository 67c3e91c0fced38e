// Path shortcutting on top of the batched motion check — the direct consumers of check_motion_collides:
//   tryShortcutByDeletingWaypoint, tryMidpointPull, tryShortcutBetweenPathPoints, generateRandomPathPoint,
//   tryDeletingEveryWaypoint, tryShortcuttingRandomlyLocally, tryShortcuttingRandomlyGlobally
//                                                                      src/planning/local_optimization.{h,cpp}
//   the "100 iterations of shortcutting" loop                         src/planning/tsp_over_prm.cpp:594-612
// The single-attempt functions keep the reference's signatures (any motion-check closure works, e.g.
// collision_functions_in_environment(env).motion_collides).  The batched drivers below give the SAME final path and
// leave the random generator in the SAME state as the reference's sequential loops, but issue far fewer GPU calls:
// a failed attempt does not change the path, so the attempts that follow it can be drawn and validated ahead of time;
// every attempt up to and including the first success of a batch is exactly what the sequential loop would have done.
#pragma once
#include <functional>

#include "sampling.hpp"

namespace mgodpl {

using MotionCollidesFn = std::function<bool(const RobotState &, const RobotState &)>;

/// local_optimization.cpp:23-45.
inline bool tryShortcutByDeletingWaypoint(RobotPath &path, size_t waypointIndex, const MotionCollidesFn &check_motion_collides) {
	assert(waypointIndex > 0 && waypointIndex + 1 < path.states.size());
	if (check_motion_collides(path.states[waypointIndex - 1], path.states[waypointIndex + 1])) return false;
	path.states.erase(path.states.begin() + (std::ptrdiff_t) waypointIndex);
	return true;
}

/// local_optimization.cpp:47-76 (the motions validated are current -> pulled and pulled -> next, as there).
inline bool tryMidpointPull(RobotPath &path, size_t waypointIndex, double pull_factor, const MotionCollidesFn &check_motion_collides) {
	assert(waypointIndex > 0 && waypointIndex + 1 < path.states.size());
	const RobotState midpoint = interpolate(path.states[waypointIndex - 1], path.states[waypointIndex + 1], 0.5);
	RobotState pulled = interpolate(path.states[waypointIndex], midpoint, pull_factor);
	if (check_motion_collides(path.states[waypointIndex], pulled) || check_motion_collides(pulled, path.states[waypointIndex + 1]))
		return false;
	path.states[waypointIndex] = std::move(pulled);
	return true;
}

namespace detail {
/// The splice of tryShortcutBetweenPathPoints (local_optimization.cpp:96-110), once the motion st1 -> st2 is known free.
inline void splice_shortcut(RobotPath &path, const PathPoint &start, const PathPoint &end, RobotState st1, RobotState st2) {
	std::vector<RobotState> kept;
	kept.reserve(start.segment_i + 2 + (path.states.size() - end.segment_i - 1));
	kept.insert(kept.end(), path.states.begin(), path.states.begin() + (std::ptrdiff_t) start.segment_i);
	kept.push_back(std::move(st1));
	kept.push_back(std::move(st2));
	kept.insert(kept.end(), path.states.begin() + (std::ptrdiff_t) end.segment_i + 1, path.states.end());
	path.states = std::move(kept);
}
}  // namespace detail

/// local_optimization.cpp:78-111.
inline bool tryShortcutBetweenPathPoints(RobotPath &path, const PathPoint &start, const PathPoint &end,
										 const MotionCollidesFn &check_motion_collides) {
	assert(start < end);
	RobotState st1 = interpolate(start, path), st2 = interpolate(end, path);
	if (check_motion_collides(st1, st2)) return false;
	detail::splice_shortcut(path, start, end, std::move(st1), std::move(st2));
	return true;
}

/// local_optimization.cpp:113-117.
inline PathPoint generateRandomPathPoint(const RobotPath &path, random_numbers::RandomNumberGenerator &rng) {
	PathPoint p;
	p.segment_i = (size_t) rng.uniformInteger(0, (int) (path.states.size() - 2));
	p.segment_t = rng.uniformReal(0.0, 1.0);
	return p;
}

/// local_optimization.cpp:119-136.
inline bool tryDeletingEveryWaypoint(RobotPath &path, const MotionCollidesFn &check_motion_collides) {
	bool shortened = false;
	for (size_t cursor = 1; cursor + 1 < path.states.size();)
		if (tryShortcutByDeletingWaypoint(path, cursor, check_motion_collides)) shortened = true;
		else ++cursor;
	return shortened;
}

/// local_optimization.cpp:138-153.
inline bool tryShortcuttingRandomlyLocally(RobotPath &path, random_numbers::RandomNumberGenerator &rng,
										   const MotionCollidesFn &check_motion_collides) {
	const PathPoint middle = generateRandomPathPoint(path, rng);
	const double radius = rng.uniformReal(0.5, 2.0);  // below 0.5 the two ends might not span a waypoint
	return tryShortcutBetweenPathPoints(path, middle.adjustByScalar(-radius, path), middle.adjustByScalar(radius, path),
										check_motion_collides);
}

/// local_optimization.cpp:155-173.
inline bool tryShortcuttingRandomlyGlobally(RobotPath &path, const MotionCollidesFn &check_motion,
											random_numbers::RandomNumberGenerator &rng) {
	assert(path.states.size() >= 3);
	PathPoint start = generateRandomPathPoint(path, rng), end = generateRandomPathPoint(path, rng);
	if (start > end) std::swap(start, end);
	return tryShortcutBetweenPathPoints(path, start, end, check_motion);
}

// ---- batched drivers --------------------------------------------------------------------------------------------------
struct ShortcutStats {
	size_t attempts = 0;         // attempts the sequential loop would have made
	size_t successes = 0;        // attempts that shortened the path
	size_t motions_checked = 0;  // motions validated on the GPU, speculative ones included
	size_t batches = 0;          // GPU calls
};

namespace detail {
struct ShortcutProposal {
	PathPoint start, end;
};
using ProposalFn = std::function<ShortcutProposal(const RobotPath &, random_numbers::RandomNumberGenerator &)>;

/// `iterations` shortcut attempts drawn by `propose`, validated a batch at a time.
inline ShortcutStats shortcut_speculatively(const robot_model::RobotModel &robot, const gpu::Scene &scene, RobotPath &path,
											random_numbers::RandomNumberGenerator &rng, size_t iterations, const ProposalFn &propose) {
	ShortcutStats stats;
	std::vector<ShortcutProposal> proposals;
	std::vector<RobotState> from, to;
	while (stats.attempts < iterations && path.states.size() > 2) {
		const size_t width = iterations - stats.attempts;
		// what the sequential loop would draw next if every attempt from here on failed
		random_numbers::RandomNumberGenerator ahead = rng;
		proposals.clear(), from.clear(), to.clear();
		for (size_t k = 0; k < width; ++k) {
			proposals.push_back(propose(path, ahead));
			from.push_back(interpolate(proposals.back().start, path));
			to.push_back(interpolate(proposals.back().end, path));
		}
		const MotionResults r = check_motions(robot, scene, from.data(), to.data(), width);
		stats.motions_checked += width, ++stats.batches;
		size_t first_free = 0;
		while (first_free < width && bit(r.hit_bits, first_free)) ++first_free;
		if (first_free == width) {  // all failed: the look-ahead generator is where the sequential loop ends up
			rng = ahead, stats.attempts += width;
			break;
		}
		// replay the draws of the attempts actually consumed, then apply the successful one
		for (size_t k = 0; k <= first_free; ++k) (void) propose(path, rng);
		splice_shortcut(path, proposals[first_free].start, proposals[first_free].end, std::move(from[first_free]), std::move(to[first_free]));
		stats.attempts += first_free + 1, ++stats.successes;
	}
	return stats;
}
}  // namespace detail

/// The reference's optimize_path_segment loop (tsp_over_prm.cpp:594-612): `iterations` x tryShortcuttingRandomlyGlobally,
/// stopping once the path has two states left.  Same resulting path and generator state, (successes + 1) GPU calls.
inline ShortcutStats shortcut_path_globally(const robot_model::RobotModel &robot, const gpu::Scene &scene, RobotPath &path,
											random_numbers::RandomNumberGenerator &rng, size_t iterations = 100) {
	return detail::shortcut_speculatively(robot, scene, path, rng, iterations, [](const RobotPath &p, random_numbers::RandomNumberGenerator &g) {
		detail::ShortcutProposal s{generateRandomPathPoint(p, g), generateRandomPathPoint(p, g)};
		if (s.start > s.end) std::swap(s.start, s.end);
		return s;
	});
}

/// `iterations` x tryShortcuttingRandomlyLocally, batched the same way.
inline ShortcutStats shortcut_path_locally(const robot_model::RobotModel &robot, const gpu::Scene &scene, RobotPath &path,
										   random_numbers::RandomNumberGenerator &rng, size_t iterations = 100) {
	return detail::shortcut_speculatively(robot, scene, path, rng, iterations, [](const RobotPath &p, random_numbers::RandomNumberGenerator &g) {
		const PathPoint middle = generateRandomPathPoint(p, g);
		const double radius = g.uniformReal(0.5, 2.0);
		return detail::ShortcutProposal{middle.adjustByScalar(-radius, p), middle.adjustByScalar(radius, p)};
	});
}

/// tryDeletingEveryWaypoint against a GPU scene (the overload local_optimization.h:35-45 declares): every remaining
/// deletion candidate is validated in one batch; the first free one is applied and the sweep resumes from there.
inline bool tryDeletingEveryWaypoint(const robot_model::RobotModel &robot, RobotPath &path, const gpu::Scene &obstacle,
									 ShortcutStats *stats_out = nullptr) {
	ShortcutStats stats;
	std::vector<RobotState> from, to;
	size_t cursor = 1;
	while (cursor + 1 < path.states.size()) {
		from.clear(), to.clear();
		for (size_t i = cursor; i + 1 < path.states.size(); ++i) from.push_back(path.states[i - 1]), to.push_back(path.states[i + 1]);
		const MotionResults r = check_motions(robot, obstacle, from.data(), to.data(), from.size());
		stats.motions_checked += from.size(), ++stats.batches;
		size_t first_free = 0;
		while (first_free < from.size() && bit(r.hit_bits, first_free)) ++first_free;
		stats.attempts += std::min(first_free + 1, from.size());
		if (first_free == from.size()) break;
		cursor += first_free;  // the sweep failed at cursor .. cursor+first_free-1 and deletes the next waypoint
		path.states.erase(path.states.begin() + (std::ptrdiff_t) cursor);
		++stats.successes;
	}
	if (stats_out) *stats_out = stats;
	return stats.successes > 0;
}

}  // namespace mgodpl
