// Scanning motions around a fruit — another direct consumer of check_robot_collision:
//   Direction, generateSidewaysScanningMotion, createLeftRightScanningMotion     src/planning/scanning_motions.{h,cpp}
//   spherical_geometry::RelativeVertex::to_cartesian                             src/planning/spherical_geometry.h:221-227
// The reference walks along an arc of constant latitude around the fruit, one fromEndEffectorAndVector state and one
// collision check at a time, and stops at the first collision.  Here the (at most 31) states of the arc are checked as
// one batch and the arc is cut at the first colliding one: same path, one GPU call.
#pragma once
#include "sampling.hpp"

namespace mgodpl {

enum class Direction { LEFT, RIGHT };

namespace spherical_geometry {
struct RelativeVertex {
	double longitude = 0.0, latitude = 0.0;
	math::Vec3d to_cartesian() const { return {std::cos(latitude) * std::cos(longitude), std::cos(latitude) * std::sin(longitude), std::sin(latitude)}; }
};
}  // namespace spherical_geometry

/// scanning_motions.cpp:15-56.
inline RobotPath generateSidewaysScanningMotion(const robot_model::RobotModel &robot, const gpu::Scene &treeTrunkObject, const math::Vec3d &fruitCenter,
												double initialLongitude, double initialLatitude, double scanDistance, Direction direction) {
	constexpr int n_steps = 32;  // the arc spans half a turn in 32 steps; step 0 (the start itself) is not part of the motion
	const double step = (direction == Direction::LEFT ? -M_PI : M_PI) / n_steps;
	std::vector<RobotState> arc;
	arc.reserve(n_steps - 1);
	for (int i = 1; i < n_steps; ++i) {
		const math::Vec3d outward = spherical_geometry::RelativeVertex{initialLongitude + i * step, initialLatitude}.to_cartesian();
		arc.push_back(fromEndEffectorAndVector(robot, fruitCenter + outward.normalized() * scanDistance, outward));
	}
	const std::vector<uint32_t> collides = check_states(robot, treeTrunkObject, arc);
	size_t keep = 0;
	while (keep < arc.size() && !bit(collides, keep)) ++keep;  // stop generating states after the first collision
	arc.resize(keep);
	return RobotPath{std::move(arc)};
}

/// scanning_motions.cpp:58-97: left and back, right and back.
inline RobotPath createLeftRightScanningMotion(const robot_model::RobotModel &robot, const gpu::Scene &treeTrunkObject, const math::Vec3d &fruitCenter,
											   double initialLongitude, double initialLatitude, double scanDistance) {
	RobotPath path;
	for (Direction d: {Direction::LEFT, Direction::RIGHT}) {
		const RobotPath scan = generateSidewaysScanningMotion(robot, treeTrunkObject, fruitCenter, initialLongitude, initialLatitude, scanDistance, d);
		path.states.insert(path.states.end(), scan.states.begin(), scan.states.end());
		path.states.insert(path.states.end(), scan.states.rbegin(), scan.states.rend());
	}
	return path;
}

}  // namespace mgodpl
