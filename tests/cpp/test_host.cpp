// Host-only C++ tests of the facade (multigoal-orchard-drone-planning-library_b200/host/mgodpl.hpp): everything here runs
// WITHOUT a CUDA device — forward kinematics, interpolation and distance, the state tools and goal projections, the
// closure-taking shortcutting functions (the oracle answers their motion checks), sharding arithmetic, roadmap walking,
// argument errors of the ABI — each against the fp64 oracle (oracle/mgodpl_oracle.hpp, test infrastructure) or a stated
// property.  The device-side twin is test_facade.cpp; this one is run by tests/test_host_side.py in the CPU suite.
#include <algorithm>
#include <cstdio>
#include <numeric>

#include "../../multigoal-orchard-drone-planning-library_b200/host/mgodpl.hpp"
#include "../../oracle/mgodpl_oracle.hpp"

using namespace mgodpl;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                        \
	do {                                                                   \
		++g_checked;                                                       \
		if (!(cond)) {                                                     \
			++g_failed;                                                    \
			std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
		}                                                                  \
	} while (0)

static orc::State to_orc(const RobotState &s) {
	orc::State o;
	o.base = {{s.base_tf.translation.x(), s.base_tf.translation.y(), s.base_tf.translation.z()},
			  {s.base_tf.orientation.x, s.base_tf.orientation.y, s.base_tf.orientation.z, s.base_tf.orientation.w}};
	o.joints = s.joint_values;
	return o;
}
static bool same_state(const RobotState &a, const orc::State &b) {
	return a.base_tf.translation.x() == b.base.t.x && a.base_tf.translation.y() == b.base.t.y && a.base_tf.translation.z() == b.base.t.z &&
		   a.base_tf.orientation.x == b.base.q.x && a.base_tf.orientation.y == b.base.q.y && a.base_tf.orientation.z == b.base.q.z &&
		   a.base_tf.orientation.w == b.base.q.w && a.joint_values == b.joints;
}
static bool same_path(const RobotPath &p, const orc::Path &o) {
	if (p.states.size() != o.size()) return false;
	for (size_t i = 0; i < o.size(); ++i)
		if (!same_state(p.states[i], o[i])) return false;
	return true;
}

/// A scene for the oracle only: one slab of two triangles standing in the middle of the sampling volume.
static orc::Scene wall() {
	const double v[] = {-0.6, 0.0, 0.0, 0.6, 0.0, 0.0, 0.6, 0.0, 2.4, -0.6, 0.0, 2.4, 0.0, -0.7, 1.0, 0.0, 0.7, 1.0, 0.0, 0.0, 2.2};
	const uint32_t t[] = {0, 1, 2, 0, 2, 3, 4, 5, 6};
	return orc::Scene(v, 7, t, 3);
}

int main() {
	// -- robots of several shapes: host FK, distance and interpolation equal the oracle's restatement bit for bit -----------
	struct Variant {
		double arm;
		std::vector<experiments::JointType> joints;
		std::vector<int> ojoints;
	};
	const std::vector<Variant> variants = {{0.75, {experiments::HORIZONTAL}, {0}},
										   {1.0, {experiments::VERTICAL}, {1}},
										   {0.5, {experiments::HORIZONTAL, experiments::VERTICAL}, {0, 1}},
										   {1.0, {experiments::VERTICAL, experiments::HORIZONTAL, experiments::HORIZONTAL}, {1, 0, 0}}};
	for (const auto &var: variants) {
		const auto robot = experiments::createProceduralRobotModel({var.arm, var.joints, false});
		const orc::Robot orobot = orc::procedural_robot(var.arm, var.ojoints);
		const auto flying_base = robot.findLinkByName("flying_base"), end_effector = robot.findLinkByName("end_effector");
		CHECK(robot.count_joint_variables() == (size_t) orobot.n_variables());
		random_numbers::RandomNumberGenerator rng(42);
		orc::Rng orng(42);
		RobotState prev;
		for (int i = 0; i < 60; ++i) {
			const RobotState s = generateUniformRandomState(robot, rng, 2.5, 3.5);
			const orc::State os = orc::generate_uniform_random_state(orobot, orng, 2.5, 3.5, M_PI_2);
			CHECK(same_state(s, os));  // same draws in the same order (state_tools.cpp:44-67)
			const auto fk = robot_model::forwardKinematics(robot, s, flying_base);
			const auto ofk = orc::forward_kinematics(orobot, os.joints.data(), 0, os.base);
			CHECK(fk.link_transforms.size() == ofk.size());
			for (size_t l = 0; l < ofk.size(); ++l)
				CHECK(fk.link_transforms[l].translation.x() == ofk[l].t.x && fk.link_transforms[l].translation.z() == ofk[l].t.z &&
					  fk.link_transforms[l].orientation.w == ofk[l].q.w && fk.link_transforms[l].orientation.y == ofk[l].q.y);
			if (i) {
				CHECK(equal_weights_distance(prev, s) == orc::equal_weights_distance(to_orc(prev), os));
				for (double t: {0.0, 0.125, 0.5, 0.9, 1.0}) CHECK(same_state(interpolate(prev, s, t), orc::interpolate(to_orc(prev), os, t)));
				CHECK(gpu::predicted_motion_cost(prev, s) == (double) orc::motion_steps(orc::equal_weights_distance(to_orc(prev), os), 0.1) + 1.0);
			}
			prev = s;
			// state tools and goal projection (state_tools.cpp:14-42,69-75; goal_sampling.cppm:58-79)
			const math::Vec3d p{rng.uniformReal(-2, 2), rng.uniformReal(-2, 2), rng.uniformReal(0, 3)};
			const math::Vec3d v{rng.gaussian01(), rng.gaussian01(), rng.gaussian01()};
			(void) orng.uniform_real(-2, 2), (void) orng.uniform_real(-2, 2), (void) orng.uniform_real(0, 3);
			(void) orng.gaussian01(), (void) orng.gaussian01(), (void) orng.gaussian01();
			if (var.joints.size() == 1) {  // fromEndEffectorAndVector writes ONE joint value: single-joint arms only (state_tools.cpp:19)
				const RobotState st = fromEndEffectorAndVector(robot, p, v);
				CHECK(same_state(st, orc::from_end_effector_and_vector(orobot, {p.x(), p.y(), p.z()}, {v.x(), v.y(), v.z()})));
				CHECK((robot_model::forwardKinematics(robot, st, flying_base).forLink(end_effector).translation - p).norm() < 1e-12);
			}
			const math::Vec3d av = arm_vector_from_state(robot, s);
			const orc::V3 oav = orc::arm_vector_from_state(orobot, os);
			CHECK(av.x() == oav.x && av.y() == oav.y && av.z() == oav.z);
			RobotState up = genUprightState(robot, rng);
			const orc::State oup = orc::gen_upright_state(orobot, orng);
			CHECK(same_state(up, oup));
			CHECK(same_state(project_to_goal(robot, up, flying_base, end_effector, p),
							 orc::project_upright_to_goal(orobot, oup, {p.x(), p.y(), p.z()}, 0.0, 0, (int) end_effector)));
		}
		// goal_region_sampler (goal_sampling.cppm:93-106) draws from the caller's generator, like genGoalStateUniform
		random_numbers::RandomNumberGenerator g1(5);
		orc::Rng g2(5);
		auto sampler = goal_region_sampler(robot, g1);
		for (int i = 0; i < 40; ++i) {
			const math::Vec3d t{0.1 * i - 2.0, 0.3, 1.5};
			CHECK(same_state(sampler(t), orc::gen_goal_state_uniform(g2, {t.x(), t.y(), t.z()}, 0.0, orobot, 0, (int) end_effector)));
		}
		CHECK(g1.uniformReal(0, 1) == g2.uniform_real(0, 1));
	}

	// -- the closure-taking path functions (local_optimization.cpp) with the ORACLE as the motion check: same paths, same
	//    generator state as the oracle's sequential loops ----------------------------------------------------------------------
	{
		const auto robot = experiments::createProceduralRobotModel();
		const orc::Robot orobot = orc::procedural_robot(0.75, {0});
		const orc::Scene oscene = wall();
		const orc::MotionFn omotion = [&](const orc::State &a, const orc::State &b) { return orc::check_motion_collides(orobot, oscene, a, b).hit; };
		size_t calls = 0;
		const auto motion = [&](const RobotState &a, const RobotState &b) {
			++calls;
			return omotion(to_orc(a), to_orc(b));
		};
		size_t shortcuts = 0, blocked = 0;
		for (unsigned seed = 0; seed < 8; ++seed) {
			random_numbers::RandomNumberGenerator make(300 + seed);
			RobotPath path0;
			for (int i = 0; i < 12; ++i) path0.append(generateUniformRandomState(robot, make, 1.8, 2.6));
			orc::Path opath0;
			for (auto &st: path0.states) opath0.push_back(to_orc(st));
			for (size_t i = 0; i + 1 < opath0.size(); ++i) blocked += omotion(opath0[i], opath0[i + 1]);
			for (double d: {0.0, 0.25, 3.75, 10.999, 11.5, 40.0}) {  // PathPoint arithmetic (RobotPath.h:116-170)
				const PathPoint p = PathPoint::fromScalar(d, path0);
				const orc::PathPoint q = orc::PathPoint::from_scalar(d, opath0);
				CHECK(p.segment_i == q.segment_i && p.segment_t == q.segment_t && p.toScalar() == q.to_scalar());
				if (p.segment_t <= 1.0) CHECK(same_state(interpolate(p, path0), orc::interpolate(q, opath0)));
			}
			{  // tryShortcuttingRandomlyGlobally, one attempt at a time (tsp_over_prm.cpp:594-612 drives it in a loop)
				RobotPath path = path0;
				orc::Path opath = opath0;
				random_numbers::RandomNumberGenerator rng(seed + 11);
				orc::Rng orng(seed + 11);
				const size_t want = orc::optimize_path_segment(opath, omotion, orng, 50);
				size_t got = 0;
				for (int i = 0; i < 50 && path.states.size() > 2; ++i) got += tryShortcuttingRandomlyGlobally(path, motion, rng);
				CHECK(got == want && same_path(path, opath));
				CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));
				shortcuts += got;
			}
			{  // locally
				RobotPath path = path0;
				orc::Path opath = opath0;
				random_numbers::RandomNumberGenerator rng(seed + 23);
				orc::Rng orng(seed + 23);
				size_t want = 0, got = 0;
				for (int i = 0; i < 40 && opath.size() > 2; ++i) want += orc::try_shortcutting_randomly_locally(opath, orng, omotion);
				for (int i = 0; i < 40 && path.states.size() > 2; ++i) got += tryShortcuttingRandomlyLocally(path, rng, motion);
				CHECK(got == want && same_path(path, opath));
				CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));
			}
			{  // deleting waypoints, midpoint pull
				RobotPath a = path0, b = path0;
				orc::Path opath = opath0, opath2 = opath0;
				CHECK(tryDeletingEveryWaypoint(a, motion) == orc::try_deleting_every_waypoint(opath, omotion) && same_path(a, opath));
				for (size_t i = 1; i + 1 < opath2.size(); ++i) CHECK(tryMidpointPull(b, i, 0.5, motion) == orc::try_midpoint_pull(opath2, i, 0.5, omotion));
				CHECK(same_path(b, opath2));
			}
		}
		std::printf("path functions over the oracle's motion check: %zu motion calls, %zu shortcuts taken, %zu blocked segments in the inputs\n", calls,
					shortcuts, blocked);
		CHECK(shortcuts > 0 && blocked > 0);  // the wall matters: some proposals are refused, some accepted
	}

	// -- sharding arithmetic (multi_gpu.hpp): cover, align, balance ------------------------------------------------------------------
	{
		for (size_t n: {0u, 1u, 31u, 32u, 33u, 1000u, 100000u, 1000003u})
			for (size_t world: {1u, 2u, 3u, 4u, 8u}) {
				size_t at = 0, smallest = SIZE_MAX, largest = 0;
				for (size_t r = 0; r < world; ++r) {
					const auto [lo, hi] = gpu::shard_range(n, r, world);
					CHECK(lo == at && hi >= lo && (lo % 32 == 0 || lo == n));
					at = hi, smallest = std::min(smallest, hi - lo), largest = std::max(largest, hi - lo);
				}
				CHECK(at == n && largest - smallest < 64);  // one block of 32, plus the ragged end of the batch
			}
		bool threw = false;
		try {
			gpu::shard_range(10, 2, 2);
		} catch (const std::invalid_argument &) {
			threw = true;
		}
		CHECK(threw);
		// costs that grow along the batch: equal-count shards would be 1 : 3 : 5 : 7 in cost
		std::vector<double> costs(64000);
		for (size_t i = 0; i < costs.size(); ++i) costs[i] = 1.0 + (double) i * 1e-3;
		const double total = std::accumulate(costs.begin(), costs.end(), 0.0);
		for (size_t world: {1u, 2u, 4u, 8u}) {
			const auto b = gpu::balanced_bounds(costs, world);
			CHECK(b.size() == world && b.front().first == 0 && b.back().second == costs.size());
			double worst = 0.0;
			for (size_t r = 0; r < world; ++r) {
				CHECK(b[r].first % 32 == 0 && (r + 1 == world || b[r].second == b[r + 1].first));
				worst = std::max(worst, std::accumulate(costs.begin() + b[r].first, costs.begin() + b[r].second, 0.0));
			}
			CHECK(worst <= 1.01 * total / (double) world);
		}
		// all-zero costs and fewer blocks than ranks still cover the batch
		const auto z = gpu::balanced_bounds(std::vector<double>(100, 0.0), 8);
		CHECK(z.size() == 8 && z.front().first == 0 && z.back().second == 100);
		for (size_t r = 0; r + 1 < z.size(); ++r) CHECK(z[r].second == z[r + 1].first && z[r].first <= z[r].second);
	}

	// -- roadmap containers and the predecessor walk (tsp_over_prm.cpp:111-128) ---------------------------------------------------
	{
		const auto robot = experiments::createProceduralRobotModel();
		random_numbers::RandomNumberGenerator rng(9);
		PRMGraph g;
		for (int i = 0; i < 6; ++i) CHECK(g.add_vertex(generateUniformRandomState(robot, rng, 2.0, 3.0)) == (size_t) i);
		g.add_edge(0, 1, equal_weights_distance(g[0], g[1]));
		g.add_edge(1, 2, equal_weights_distance(g[1], g[2]));
		CHECK(g.num_vertices() == 6 && g.num_edges() == 2 && g.edges == (std::vector<uint32_t>{0, 1, 1, 2}));
		// source 0: 0 <- 1 <- 2; vertices 3, 4, 5 name each other (what only a malformed map does): the walk must end
		const std::vector<size_t> pred = {0, 0, 1, 4, 5, 3};
		const RobotPath p = retrace_path(g, pred, 2);
		CHECK(p.states.size() == 2 && p.states[0] == g[1] && p.states[1] == g[2]);  // the start node itself is not emitted
		CHECK(retrace_path(g, pred, 0).states.empty());
		CHECK(retrace_path(g, pred, 4).states.size() <= pred.size());
		const PRMGraph h = roadmap_from_edges(g.states, {{0, 1, 0.5}, {2, 1, 0.25}});
		CHECK(h.num_edges() == 2 && h.edges == (std::vector<uint32_t>{0, 1, 2, 1}) && h.weights == (std::vector<double>{0.5, 0.25}));
	}

	// -- the ABI refuses bad arguments before it looks for a device; without one it says so (no CPU fallback) -------------------------
	{
		const uint32_t edges[] = {0, 3};
		const double w[] = {1.0};
		const uint32_t src[] = {0};
		double dist[3];
		CHECK(mgodpl_roadmap_shortest_paths(3, edges, w, 1, src, 1, dist, nullptr, nullptr, nullptr) == MGODPL_E_INVALID_ARGUMENT);
		CHECK(std::string(mgodpl_last_error()).find("out of range") != std::string::npos);
		// the host-side predecessor repair needs no device at all: 1 <-> 2 joined by a zero-weight edge behind vertex 0
		const uint32_t row[] = {0, 1, 3, 4}, col[] = {1, 0, 2, 1};
		const double cw[] = {1.0, 1.0, 0.0, 0.0}, d[] = {0.0, 1.0, 1.0};
		uint32_t pred[] = {0, 2, 1};
		CHECK(mgodpl_roadmap_repair_predecessors(3, row, col, cw, d, pred) == MGODPL_OK);
		CHECK(pred[0] == 0 && pred[1] == 0 && pred[2] == 1);
		if (mgodpl_device_count() <= 0) {
			const Mesh one{{{0, 0, 0}, {1, 0, 0}, {0, 1, 0}}, {{0, 1, 2}}};
			bool refused = false;
			try {
				gpu::meshToScene(one);
			} catch (const std::runtime_error &e) {
				refused = std::string(e.what()).find("device") != std::string::npos;
			}
			CHECK(refused);
		}
	}

	std::printf("%d checks, %d failed\n", g_checked, g_failed);
	return g_failed ? 1 : 0;
}
