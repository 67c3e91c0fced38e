// One host process driving several GPUs: the C++ counterpart of sharding.py + torchrun.
//
// The path shards with no exchange step: every state / motion / goal sample is independent given the
// read-only scene, so a DeviceGroup holds one replica of the scene BVH per device and splits each batch into contiguous
// shards whose boundaries are multiples of 32 (result words never straddle two devices).  States: equal shards.  Motions:
// shards balanced by the number of state checks a collision-free motion costs, ceil(equal_weights_distance / max_step) + 1,
// which is known before launch (src/planning/collision_detection.cpp:88-92).  Goal sampling: sharded by target.  One host
// thread per device issues that device's shard through the ordinary C ABI; results land in place in the caller's arrays.
// The reference's analogue is its std::for_each(std::execution::par, ...) over independent problems
// (src/benchmarks/goal_sample_success_rate.cpp:165-172).
#pragma once
#include <exception>
#include <thread>
#include <utility>

#include "sampling.hpp"

namespace mgodpl::gpu {

/// [lo, hi) of shard `rank` of n queries; lo is a multiple of `align`; shards differ by one block of `align` at most, and the last
/// non-empty shard may additionally end inside a block (so sizes differ by less than 2 x align).
inline std::pair<size_t, size_t> shard_range(size_t n, size_t rank, size_t world, size_t align = 32) {
	if (!world || rank >= world) throw std::invalid_argument("bad rank / world size");
	const size_t blocks = (n + align - 1) / align, per = blocks / world, extra = blocks % world;
	const size_t lo_b = rank * per + std::min(rank, extra), hi_b = lo_b + per + (rank < extra ? 1 : 0);
	return {std::min(lo_b * align, n), std::min(hi_b * align, n)};
}

/// Contiguous shards with (nearly) equal total cost; every boundary is a multiple of `align`.
inline std::vector<std::pair<size_t, size_t>> balanced_bounds(const std::vector<double> &costs, size_t world, size_t align = 32) {
	if (!world) throw std::invalid_argument("bad world size");
	const size_t n = costs.size(), blocks = (n + align - 1) / align;
	std::vector<double> before(blocks + 1, 0.0);  // cost before block k
	for (size_t k = 0; k < blocks; ++k) {
		double c = 0.0;
		for (size_t i = k * align; i < std::min(n, (k + 1) * align); ++i) c += costs[i];
		before[k + 1] = before[k] + c;
	}
	const double total = before[blocks];
	std::vector<size_t> cuts{0};
	for (size_t r = 1; r < world; ++r) {
		size_t k = total > 0.0 ? (size_t) (std::lower_bound(before.begin(), before.end(), total * (double) r / (double) world) - before.begin())
							   : blocks * r / world;
		cuts.push_back(std::min(std::max(k, cuts.back()), blocks));
	}
	cuts.push_back(blocks);
	std::vector<std::pair<size_t, size_t>> out;
	for (size_t r = 0; r < world; ++r) out.push_back({std::min(cuts[r] * align, n), std::min(cuts[r + 1] * align, n)});
	return out;
}

/// State checks a collision-free motion costs (collision_detection.cpp:88-92): what the motion shards are balanced by.
inline double predicted_motion_cost(const RobotState &from, const RobotState &to, double max_step = 0.1) {
	const double d = equal_weights_distance(from, to);
	return std::isfinite(d) ? std::ceil(d / max_step) + 1.0 : 1.0;
}

class DeviceGroup {
	std::vector<int> devices_;
	std::vector<Scene> scenes_;

	/// Run f(r) on one host thread per replica; the first exception (if any) is rethrown on the caller's thread.
	template<class F>
	void for_each_replica(F f) const {
		std::vector<std::thread> pool;
		std::vector<std::exception_ptr> errors(scenes_.size());
		for (size_t r = 0; r < scenes_.size(); ++r)
			pool.emplace_back([&, r] {
				try {
					f(r);
				} catch (...) { errors[r] = std::current_exception(); }
			});
		for (auto &t: pool) t.join();
		for (auto &e: errors)
			if (e) std::rethrow_exception(e);
	}

public:
	/// One replica of `mesh` per listed device (default: every visible device).  A device may be listed more than once
	/// (several replicas on one GPU: used by the tests on single-GPU boxes).
	explicit DeviceGroup(const Mesh &mesh, std::vector<int> devices = {}, const mgodpl_scene_opts *opts = nullptr) : devices_(std::move(devices)) {
		if (devices_.empty())
			for (int d = 0; d < mgodpl_device_count(); ++d) devices_.push_back(d);
		if (devices_.empty()) throw std::runtime_error("no CUDA device: this library has no CPU fallback");
		scenes_.resize(devices_.size());
		for_each_replica([&](size_t r) {
			throw_status(mgodpl_set_device(devices_[r]));  // a fresh thread: the scene is built on (and stays bound to) this device
			scenes_[r] = Scene(mesh, opts);
		});
	}
	size_t size() const { return scenes_.size(); }
	const Scene &scene(size_t replica) const { return scenes_.at(replica); }
	int device(size_t replica) const { return devices_.at(replica); }

	/// N x check_robot_collision over all devices; same packed bits as mgodpl::check_states on one device.
	std::vector<uint32_t> check_states(const robot_model::RobotModel &robot, const RobotState *states, size_t n) const {
		std::vector<uint32_t> bits((n + 31) / 32);
		robot.device_handle();  // build the (device-independent) robot handle once, before the threads race for it
		for_each_replica([&](size_t r) {
			const auto [lo, hi] = shard_range(n, r, size());
			if (hi <= lo) return;
			const auto part = mgodpl::check_states(robot, scenes_[r], states + lo, hi - lo);
			std::copy(part.begin(), part.end(), bits.begin() + lo / 32);
		});
		return bits;
	}
	std::vector<uint32_t> check_states(const robot_model::RobotModel &robot, const std::vector<RobotState> &states) const {
		return check_states(robot, states.data(), states.size());
	}

	/// M x check_motion_collides over all devices, shards balanced by predicted step counts.
	MotionResults check_motions(const robot_model::RobotModel &robot, const RobotState *from, const RobotState *to, size_t m,
								double max_step = 0.1) const {
		MotionResults out{std::vector<uint32_t>((m + 31) / 32), std::vector<double>(m), std::vector<uint32_t>(m)};
		std::vector<double> costs(m);
		for (size_t i = 0; i < m; ++i) costs[i] = predicted_motion_cost(from[i], to[i], max_step);
		const auto bounds = balanced_bounds(costs, size());
		robot.device_handle();
		for_each_replica([&](size_t r) {
			const auto [lo, hi] = bounds[r];
			if (hi <= lo) return;
			const MotionResults part = mgodpl::check_motions(robot, scenes_[r], from + lo, to + lo, hi - lo, max_step);
			std::copy(part.hit_bits.begin(), part.hit_bits.end(), out.hit_bits.begin() + lo / 32);
			std::copy(part.toi.begin(), part.toi.end(), out.toi.begin() + lo);
			std::copy(part.n_steps.begin(), part.n_steps.end(), out.n_steps.begin() + lo);
		});
		return out;
	}

	/// sample_goal_states sharded by target (every target draws from the same per-problem generator, so a shard's result
	/// does not depend on which device computes it).
	GoalSamples sample_goal_states(const robot_model::RobotModel &robot, const std::vector<math::Vec3d> &targets, size_t samples,
								   unsigned seed = 42, double distance_from_target = 0.0, bool want_states = false) const {
		const size_t f = targets.size(), wpr = (samples + 31) / 32;
		GoalSamples out{f, samples, std::vector<uint32_t>(f * wpr), std::vector<uint32_t>(f), {}};
		if (want_states) out.states.resize(f * samples);
		robot.device_handle();
		for_each_replica([&](size_t r) {
			const auto [lo, hi] = shard_range(f, r, size(), 1);
			if (hi <= lo) return;
			const std::vector<math::Vec3d> mine(targets.begin() + lo, targets.begin() + hi);
			GoalSamples part = mgodpl::sample_goal_states(robot, scenes_[r], mine, samples, seed, distance_from_target, want_states);
			std::copy(part.hit_bits.begin(), part.hit_bits.end(), out.hit_bits.begin() + lo * wpr);
			std::copy(part.collisions.begin(), part.collisions.end(), out.collisions.begin() + lo);
			if (want_states) std::move(part.states.begin(), part.states.end(), out.states.begin() + lo * samples);
		});
		return out;
	}
};

}  // namespace mgodpl::gpu
