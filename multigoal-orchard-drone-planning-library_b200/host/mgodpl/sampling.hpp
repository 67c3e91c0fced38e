// The reference's state samplers for this path, over std::mt19937 exactly like its RandomNumberGenerator facade:
//   random_numbers::RandomNumberGenerator   src/planning/RandomNumberGenerator.h:21-77
//   generateUniformRandomState, fromEndEffectorAndVector, arm_vector_from_state   src/planning/state_tools.cpp:14-75
//   genUprightState, genGoalStateUniform, findGoalStateByUniformSampling, generateUniformRandomArmVectorState
//                                                                                 src/planning/goal_sampling.cpp:13-124
//   try_at_valid_goal_samples, project_to_goal, goal_region_sampler               src/planning/goal_sampling.cppm:37-106
// plus the batched sample_goal_states (F targets x S samples on the GPU).
#pragma once
#include <random>

#include "collision_detection.hpp"

namespace random_numbers {
class RandomNumberGenerator {
	std::mt19937 generator;

public:
	explicit RandomNumberGenerator(unsigned int seed) : generator(seed) {}
	RandomNumberGenerator() : generator(std::random_device()()) {}
	double uniform01() { return std::generate_canonical<double, 10>(generator); }
	int uniformInteger(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(generator); }
	double gaussian01() { return std::normal_distribution<double>(0.0, 1.0)(generator); }
	double uniformReal(double lo, double hi) { return std::uniform_real_distribution<double>(lo, hi)(generator); }
	mgodpl::math::Vec3d random_unit_vector() { return mgodpl::math::Vec3d(gaussian01(), gaussian01(), gaussian01()).normalized(); }
};
}  // namespace random_numbers

namespace mgodpl {

namespace spherical_geometry {
inline double latitude(const math::Vec3d &p, const math::Vec3d &c = {0, 0, 0}) {
	math::Vec3d d = p - c;
	return std::atan2(d.z(), std::sqrt(d.x() * d.x() + d.y() * d.y()));
}
inline double longitude(const math::Vec3d &p, const math::Vec3d &c = {0, 0, 0}) {
	math::Vec3d d = p - c;
	return std::atan2(d.y(), d.x());
}
}  // namespace spherical_geometry

inline RobotState generateUniformRandomState(const robot_model::RobotModel &robot, random_numbers::RandomNumberGenerator &rng,
											 const double h_range = 5.0, const double v_range = 10.0, const double joint_angle_range = M_PI / 2.0) {
	// draw order: x, y, z, yaw, then one value per joint (fixed joints included — state_tools.cpp:61-64)
	const double x = rng.uniformReal(-h_range, h_range), y = rng.uniformReal(-h_range, h_range), z = rng.uniformReal(0, v_range);
	RobotState s{{{x, y, z}, math::Quaterniond::fromAxisAngle(math::Vec3d::UnitZ(), rng.uniformReal(0, 2 * M_PI))}, {}};
	s.joint_values.reserve(robot.getJoints().size());
	for (size_t i = 0; i < robot.getJoints().size(); ++i) s.joint_values.push_back(rng.uniformReal(-joint_angle_range, joint_angle_range));
	return s;
}

inline RobotState fromEndEffectorAndVector(const robot_model::RobotModel &robot, const math::Vec3d &endEffectorPoint, const math::Vec3d &vector) {
	std::vector<double> arm_angles{-spherical_geometry::latitude(vector)};
	math::Transformd base{{0, 0, 0}, math::Quaterniond::fromAxisAngle(math::Vec3d::UnitZ(), spherical_geometry::longitude(vector) + M_PI / 2.0)};
	const auto fk = robot_model::forwardKinematics(robot, arm_angles, robot.findLinkByName("flying_base"), base);
	const math::Vec3d ee = fk.forLink(robot.findLinkByName("end_effector")).translation;
	base.translation = base.translation + endEffectorPoint - ee;
	return {base, arm_angles};
}

inline math::Vec3d arm_vector_from_state(const robot_model::RobotModel &robot, const RobotState &state) {
	return robot_model::forwardKinematics(robot, state).forLink(robot.findLinkByName("end_effector")).orientation.rotate({0.0, -1.0, 0.0});
}

inline RobotState genUprightState(const robot_model::RobotModel &robot, random_numbers::RandomNumberGenerator &rng) {
	RobotState s{{{0, 0, 0}, math::Quaterniond::fromAxisAngle(math::Vec3d::UnitZ(), rng.uniformReal(-M_PI, M_PI))}, {}};
	for (const auto &j: robot.getJoints())
		if (auto *r = std::get_if<robot_model::RobotModel::RevoluteJoint>(&j.type_specific))
			s.joint_values.push_back(rng.uniformReal(r->min_angle, r->max_angle));
	return s;
}

/// Deterministic half of genGoalStateUniform: shift the base so that the end effector reaches `target`.
inline RobotState project_to_goal(const robot_model::RobotModel &robot, RobotState to_project, const robot_model::RobotModel::LinkId flying_base,
								  const robot_model::RobotModel::LinkId end_effector, const math::Vec3d target, double distance_from_target = 0.0) {
	const auto fk = robot_model::forwardKinematics(robot, to_project.joint_values, flying_base, to_project.base_tf);
	const auto &ee = fk.forLink(end_effector);
	math::Vec3d delta = target - ee.translation;
	if (distance_from_target > 0.0) delta += ee.orientation.rotate({0.0, -distance_from_target, 0.0});
	to_project.base_tf.translation = to_project.base_tf.translation + delta;
	return to_project;
}

inline RobotState genGoalStateUniform(random_numbers::RandomNumberGenerator &rng, const math::Vec3d &target, double distance_from_target,
									  const robot_model::RobotModel &robot, const robot_model::RobotModel::LinkId &flying_base,
									  const robot_model::RobotModel::LinkId &end_effector) {
	return project_to_goal(robot, genUprightState(robot, rng), flying_base, end_effector, target, distance_from_target);
}
inline RobotState genGoalStateUniform(random_numbers::RandomNumberGenerator &rng, const math::Vec3d &target, const robot_model::RobotModel &robot,
									  const robot_model::RobotModel::LinkId &flying_base, const robot_model::RobotModel::LinkId &end_effector) {
	return genGoalStateUniform(rng, target, 0.0, robot, flying_base, end_effector);
}

/// First collision-free goal sample out of at most max_attempts.  Candidates are drawn from a *copy* of rng and checked
/// in batches on the GPU; rng itself is then advanced by exactly the draws the reference's sequential loop would have
/// consumed, so results and generator state match the reference call for call.
inline std::optional<RobotState> findGoalStateByUniformSampling(const math::Vec3d &target, const robot_model::RobotModel &robot,
																const robot_model::RobotModel::LinkId &flying_base,
																const robot_model::RobotModel::LinkId &end_effector, const gpu::Scene &tree_trunk_object,
																random_numbers::RandomNumberGenerator &rng, size_t max_attempts) {
	size_t done = 0, batch = 32;
	while (done < max_attempts) {
		const size_t n = std::min(batch, max_attempts - done);
		random_numbers::RandomNumberGenerator lookahead = rng;
		std::vector<RobotState> cand;
		cand.reserve(n);
		for (size_t i = 0; i < n; ++i) cand.push_back(genGoalStateUniform(lookahead, target, robot, flying_base, end_effector));
		const auto bits = check_states(robot, tree_trunk_object, cand);
		for (size_t i = 0; i < n; ++i)
			if (!bit(bits, i)) {
				for (size_t k = 0; k <= i; ++k) (void) genUprightState(robot, rng);  // consume what the sequential loop consumed
				return cand[i];
			}
		rng = lookahead;
		done += n;
		batch = std::min<size_t>(batch * 4, 4096);
	}
	return std::nullopt;
}

inline std::optional<RobotState> generateUniformRandomArmVectorState(const robot_model::RobotModel &robot, const gpu::Scene &tree_trunk_object,
																	 const math::Vec3d &fruit_center, random_numbers::RandomNumberGenerator &rng,
																	 const int max_attempts, const double ee_distance) {
	int done = 0, batch = 32;
	while (done < max_attempts) {
		const int n = std::min(batch, max_attempts - done);
		random_numbers::RandomNumberGenerator lookahead = rng;
		std::vector<RobotState> cand;
		for (int i = 0; i < n; ++i) {
			math::Vec3d v(lookahead.gaussian01(), lookahead.gaussian01(), lookahead.gaussian01());
			v = v.normalized();
			cand.push_back(fromEndEffectorAndVector(robot, fruit_center - v * ee_distance, -v));
		}
		const auto bits = check_states(robot, tree_trunk_object, cand);
		for (int i = 0; i < n; ++i)
			if (!bit(bits, (size_t) i)) {
				for (int k = 0; k <= i; ++k) (void) rng.gaussian01(), (void) rng.gaussian01(), (void) rng.gaussian01();
				return cand[(size_t) i];
			}
		rng = lookahead;
		done += n;
		batch = std::min(batch * 4, 4096);
	}
	return std::nullopt;
}

template<typename T>
std::optional<T> try_at_valid_goal_samples(const std::function<bool(const RobotState &)> &collision, const std::function<RobotState()> &goal_sample,
										   int max_samples, const std::function<std::optional<T>(const RobotState &)> &operation) {
	for (int i = 0; i < max_samples; ++i) {
		RobotState goal = goal_sample();
		if (!collision(goal))
			if (auto r = operation(goal)) return r;
	}
	return std::nullopt;
}

using GoalTargetSampleFunctionFn = std::function<RobotState(const math::Vec3d &target)>;
inline GoalTargetSampleFunctionFn goal_region_sampler(const robot_model::RobotModel &robot_model, random_numbers::RandomNumberGenerator &rng) {
	return [&](const math::Vec3d &target) {
		return genGoalStateUniform(rng, target, 0.0, robot_model, robot_model.findLinkByName("flying_base"), robot_model.findLinkByName("end_effector"));
	};
}

/// Batched goal sampling (goal_sample_success_rate.cpp:121-142 for F fruit at once): every target gets `samples` states
/// from a fresh std::mt19937(seed) — the reference re-seeds per problem — FK + collision filter on the GPU.
struct GoalSamples {
	size_t n_targets = 0, n_samples = 0;
	std::vector<uint32_t> hit_bits;     // n_targets x ceil(n_samples / 32) words
	std::vector<uint32_t> collisions;   // per target: how many samples collide (the benchmark's "collisions")
	std::vector<RobotState> states;     // n_targets x n_samples, row-major (empty unless requested)
	bool collides(size_t target, size_t sample) const {
		const size_t wpr = (n_samples + 31) / 32;
		return (hit_bits[target * wpr + (sample >> 5)] >> (sample & 31)) & 1u;
	}
};
inline GoalSamples sample_goal_states(const robot_model::RobotModel &robot, const gpu::Scene &scene, const std::vector<math::Vec3d> &targets,
									  size_t samples, unsigned seed = 42, double distance_from_target = 0.0, bool want_states = false) {
	const size_t nv = robot.count_joint_variables(), f = targets.size(), wpr = (samples + 31) / 32;
	random_numbers::RandomNumberGenerator rng(seed);
	std::vector<double> draws(samples * (1 + nv));
	for (size_t s = 0; s < samples; ++s) {  // the draws genUprightState makes, in its order
		draws[s * (1 + nv)] = rng.uniformReal(-M_PI, M_PI);
		size_t k = 1;
		for (const auto &j: robot.getJoints())
			if (auto *r = std::get_if<robot_model::RobotModel::RevoluteJoint>(&j.type_specific))
				draws[s * (1 + nv) + k++] = rng.uniformReal(r->min_angle, r->max_angle);
	}
	std::vector<double> t(f * 3);
	for (size_t i = 0; i < f; ++i) t[3 * i] = targets[i].x(), t[3 * i + 1] = targets[i].y(), t[3 * i + 2] = targets[i].z();
	GoalSamples out{f, samples, std::vector<uint32_t>(f * wpr), std::vector<uint32_t>(f), {}};
	std::vector<double> st(want_states ? f * samples * (7 + nv) : 0);
	gpu::throw_status(mgodpl_sample_goals(scene.handle(), robot.device_handle(), t.data(), f, draws.data(), samples, seed, distance_from_target,
										  out.hit_bits.data(), out.collisions.data(), want_states ? st.data() : nullptr, nullptr));
	if (want_states) {
		out.states.resize(f * samples);
		for (size_t i = 0; i < f * samples; ++i) {
			const double *r = st.data() + i * (7 + nv);
			out.states[i] = {{{r[0], r[1], r[2]}, {r[3], r[4], r[5], r[6]}}, std::vector<double>(r + 7, r + 7 + nv)};
		}
	}
	return out;
}

}  // namespace mgodpl
