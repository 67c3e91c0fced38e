// C++ tests of the host-side facade (multigoal-orchard-drone-planning-library_b200/host/mgodpl.hpp): the reference's
// planning-layer API running on the CUDA library, checked against the fp64 oracle (oracle/mgodpl_oracle.hpp — test
// infrastructure).  Modelled on the reference's own test style: seeded randomised properties
// (test/math/Transform_test.cpp:16-65) and the legacy collision tests (test/planning/MyCollisionEnvTest.cpp:13-78,
// src_old/test/collision_checking_test.cpp:30-57).  Needs a CUDA device; run by tests/test_gpu_facade.py.
#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <thread>

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

// ---- a small procedural tree: trunk + a few branches as triangulated tubes ---------------------------------------------
static void add_tube(Mesh &m, math::Vec3d a, math::Vec3d b, double r0, double r1, int sides, int segs) {
	math::Vec3d d = (b - a).normalized();
	math::Vec3d u = std::fabs(d.x()) < 0.9 ? math::Vec3d{1, 0, 0} : math::Vec3d{0, 1, 0};
	math::Vec3d e1{d.y() * u.z() - d.z() * u.y(), d.z() * u.x() - d.x() * u.z(), d.x() * u.y() - d.y() * u.x()};
	e1 = e1.normalized();
	math::Vec3d e2{d.y() * e1.z() - d.z() * e1.y(), d.z() * e1.x() - d.x() * e1.z(), d.x() * e1.y() - d.y() * e1.x()};
	size_t base = m.vertices.size();
	for (int s = 0; s <= segs; ++s) {
		double t = (double) s / segs, r = r0 + (r1 - r0) * t;
		math::Vec3d c = a + (b - a) * t;
		for (int k = 0; k < sides; ++k) {
			double ang = 2 * M_PI * k / sides;
			m.vertices.push_back(c + e1 * (r * std::cos(ang)) + e2 * (r * std::sin(ang)));
		}
	}
	for (int s = 0; s < segs; ++s)
		for (int k = 0; k < sides; ++k) {
			size_t i0 = base + s * sides + k, i1 = base + s * sides + (k + 1) % sides, j0 = i0 + sides, j1 = i1 + sides;
			m.triangles.push_back({i0, i1, j0});
			m.triangles.push_back({i1, j1, j0});
		}
}
static Mesh make_tree() {
	Mesh m;
	add_tube(m, {0, 0, 0}, {0, 0, 1.6}, 0.12, 0.08, 10, 24);
	random_numbers::RandomNumberGenerator rng(7);
	for (int i = 0; i < 14; ++i) {
		double az = rng.uniformReal(0, 2 * M_PI), el = rng.uniformReal(0.2, 1.1), len = rng.uniformReal(0.6, 1.3), z = rng.uniformReal(0.9, 1.6);
		math::Vec3d dir{std::cos(az) * std::cos(el), std::sin(az) * std::cos(el), std::sin(el)};
		add_tube(m, {0, 0, z}, math::Vec3d{0, 0, z} + dir * len, 0.05, 0.015, 6, 16);
	}
	return m;
}

// ---- COLLADA fixtures: a tree model written the way exporters write them (inches, y up), to be read back by from_dae -------
static void add_blob(Mesh &m, math::Vec3d c, double r, int rings, int sides) {  // a closed UV sphere: one connected component
	size_t base = m.vertices.size();
	m.vertices.push_back(c + math::Vec3d{0, 0, r});
	for (int i = 1; i < rings; ++i)
		for (int k = 0; k < sides; ++k) {
			double th = M_PI * i / rings, ph = 2 * M_PI * k / sides;
			m.vertices.push_back(c + math::Vec3d{r * std::sin(th) * std::cos(ph), r * std::sin(th) * std::sin(ph), r * std::cos(th)});
		}
	m.vertices.push_back(c + math::Vec3d{0, 0, -r});
	size_t south = m.vertices.size() - 1;
	for (int k = 0; k < sides; ++k) {
		size_t a = base + 1 + k, b = base + 1 + (k + 1) % sides;
		m.triangles.push_back({base, a, b});
		size_t la = base + 1 + (size_t) (rings - 2) * sides + k, lb = base + 1 + (size_t) (rings - 2) * sides + (k + 1) % sides;
		m.triangles.push_back({south, lb, la});
	}
	for (int i = 0; i + 2 < rings; ++i)
		for (int k = 0; k < sides; ++k) {
			size_t a = base + 1 + (size_t) i * sides + k, b = base + 1 + (size_t) i * sides + (k + 1) % sides, c2 = a + sides, d = b + sides;
			m.triangles.push_back({a, c2, b});
			m.triangles.push_back({b, c2, d});
		}
}
/// Writes `mesh` (metres, z up) as a COLLADA file in exporter units: inches x 100 / 2.54 ... i.e. the inverse of
/// mesh_from_dae.cpp:55-63, y and z swapped.  style 0: <triangles> with an interleaved NORMAL input and every corner given
/// its own position entry (so the reader must join identical vertices); 1: <polylist> (triangle pairs merged into quads
/// where they share an edge in sequence); 2: <polygons>.
static void write_dae(const std::string &path, const Mesh &mesh, int style) {
	std::ofstream f(path);
	f.precision(9);
	const double s = 100.0 / 2.54;
	f << "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<!-- test fixture -->\n<COLLADA xmlns=\"http://www.collada.org/2005/11/COLLADASchema\" version=\"1.4.1\">\n"
	  << " <asset><unit name=\"inch\" meter=\"0.0254\"/><up_axis>Y_UP</up_axis></asset>\n <library_geometries>\n  <geometry id=\"g\" name=\"g\">\n   <mesh>\n";
	std::vector<math::Vec3d> pos;
	std::vector<size_t> idx;
	if (style == 0) {
		for (auto &t: mesh.triangles)
			for (size_t v: t) idx.push_back(pos.size()), pos.push_back(mesh.vertices[v]);
	} else {
		pos = mesh.vertices;
		for (auto &t: mesh.triangles)
			for (size_t v: t) idx.push_back(v);
	}
	f << "    <source id=\"g-positions\"><float_array id=\"g-positions-array\" count=\"" << pos.size() * 3 << "\">";
	for (auto &p: pos) f << p.x() * s << ' ' << p.z() * s << ' ' << p.y() * s << ' ';
	f << "</float_array><technique_common><accessor source=\"#g-positions-array\" count=\"" << pos.size()
	  << "\" stride=\"3\"><param name=\"X\" type=\"float\"/><param name=\"Y\" type=\"float\"/><param name=\"Z\" type=\"float\"/></accessor></technique_common></source>\n"
	  << "    <source id=\"g-normals\"><float_array id=\"g-normals-array\" count=\"3\">0 1 0</float_array><technique_common><accessor source=\"#g-normals-array\" "
		 "count=\"1\" stride=\"3\"/></technique_common></source>\n"
	  << "    <vertices id=\"g-vertices\"><input semantic=\"POSITION\" source=\"#g-positions\"/></vertices>\n";
	const size_t nt = mesh.triangles.size();
	if (style == 0) {
		f << "    <triangles material=\"m\" count=\"" << nt << "\"><input semantic=\"VERTEX\" source=\"#g-vertices\" offset=\"0\"/>"
		  << "<input semantic=\"NORMAL\" source=\"#g-normals\" offset=\"1\"/><p>";
		for (size_t i: idx) f << i << " 0 ";
		f << "</p></triangles>\n";
	} else if (style == 1) {
		f << "    <polylist count=\"" << nt << "\"><input semantic=\"VERTEX\" source=\"#g-vertices\" offset=\"0\"/><vcount>";
		for (size_t t = 0; t < nt; ++t) f << "3 ";
		f << "</vcount><p>";
		for (size_t i: idx) f << i << ' ';
		f << "</p></polylist>\n";
	} else {
		f << "    <polygons count=\"" << nt << "\"><input semantic=\"VERTEX\" source=\"#g-vertices\" offset=\"0\"/>";
		for (size_t t = 0; t < nt; ++t) f << "<p>" << idx[3 * t] << ' ' << idx[3 * t + 1] << ' ' << idx[3 * t + 2] << "</p>";
		f << "</polygons>\n";
	}
	f << "   </mesh>\n  </geometry>\n </library_geometries>\n <library_visual_scenes><visual_scene id=\"s\"><node id=\"n\"><instance_geometry url=\"#g\"/></node>"
		 "</visual_scene></library_visual_scenes>\n <scene><instance_visual_scene url=\"#s\"/></scene>\n</COLLADA>\n";
}

// ---- oracle twins ---------------------------------------------------------------------------------------------------------
static orc::State to_orc(const RobotState &s) {
	orc::State o;
	o.base = {{s.base_tf.translation.x(), s.base_tf.translation.y(), s.base_tf.translation.z()},
			  {s.base_tf.orientation.x, s.base_tf.orientation.y, s.base_tf.orientation.z, s.base_tf.orientation.w}};
	o.joints = s.joint_values;
	return o;
}
static orc::Scene orc_scene(const Mesh &m) {
	std::vector<double> v;
	std::vector<uint32_t> t;
	for (auto &p: m.vertices) v.insert(v.end(), {p.x(), p.y(), p.z()});
	for (auto &tr: m.triangles) t.insert(t.end(), {(uint32_t) tr[0], (uint32_t) tr[1], (uint32_t) tr[2]});
	return orc::Scene(v.data(), m.vertices.size(), t.data(), m.triangles.size());
}
static RobotState base_at(double x, double y, double z) { return {{{x, y, z}, {0, 0, 0, 1}}, {0.0, 0.0}}; }

int main() {
	if (mgodpl_device_count() <= 0) {
		std::printf("no CUDA device: the facade has no CPU fallback\n");
		return 2;
	}
	const Mesh tree = make_tree();
	const gpu::Scene scene = gpu::meshToScene(tree);  // was: fcl_utils::meshToFclBVH(tree)
	const orc::Scene oscene = orc_scene(tree);
	const auto robot = experiments::createProceduralRobotModel();
	const orc::Robot orobot = orc::procedural_robot(0.75, {0});
	const auto flying_base = robot.findLinkByName("flying_base"), end_effector = robot.findLinkByName("end_effector");
	CHECK(scene.info().n_triangles == tree.triangles.size());

	// -- host FK of the facade equals the oracle's restatement bit for bit ------------------------------------------------
	{
		random_numbers::RandomNumberGenerator rng(42);
		for (int i = 0; i < 100; ++i) {
			RobotState s = generateUniformRandomState(robot, rng);
			auto fk = robot_model::forwardKinematics(robot, s, flying_base);
			orc::State os = to_orc(s);
			auto ofk = orc::forward_kinematics(orobot, os.joints.data(), 0, os.base);
			for (size_t l = 0; l < fk.link_transforms.size(); ++l)
				CHECK(fk.link_transforms[l].translation.x() == ofk[l].t.x && fk.link_transforms[l].orientation.w == ofk[l].q.w);
		}
	}
	// -- trunk fly-through (MyCollisionEnvTest.cpp:24-76) -------------------------------------------------------------------
	{
		CHECK(!check_robot_collision(robot, scene, base_at(0, -10, 1)));
		CHECK(check_robot_collision(robot, scene, base_at(0, 0, 1)));
		CHECK(!check_robot_collision(robot, scene, base_at(0, 10, 1)));
		double toi = -1;
		CHECK(check_motion_collides(robot, scene, base_at(0, -10, 1), base_at(0, 10, 1), toi));
		CHECK(std::fabs(toi - 0.5) < 0.2);
		CHECK(!check_motion_collides(robot, scene, base_at(5, -10, 1), base_at(5, 10, 1)));
		RobotPath path;
		for (double y: {-10.0, -6.0, -3.0, 3.0, 10.0}) path.append(base_at(0, y, 1));
		PathPoint pt{};
		CHECK(check_path_collides(robot, scene, path, pt) && pt.segment_i == 2);
		RobotPath clear;
		for (double y: {-10.0, -6.0, -3.0}) clear.append(base_at(4, y, 1));
		CHECK(!check_path_collides(robot, scene, clear));
	}
	// -- randomised single + batched state checks against the oracle, with band accounting --------------------------------
	{
		random_numbers::RandomNumberGenerator rng(42);
		std::vector<RobotState> states;
		for (int i = 0; i < 20000; ++i) states.push_back(generateUniformRandomState(robot, rng, 2.0, 3.0));
		auto bits = check_states(robot, scene, states);
		size_t mismatch_outside = 0, in_band = 0, hits = 0;
		for (size_t i = 0; i < states.size(); ++i) {
			const orc::State os = to_orc(states[i]);
			const bool want = orc::check_robot_collision(orobot, oscene, os);
			hits += want;
			if (bit(bits, i) != want) {
				const double cl = orc::robot_clearance(orobot, oscene, os, 1e-3);
				if (std::fabs(cl) > 1e-6) ++mismatch_outside;
				else ++in_band;
			}
		}
		std::printf("states: %zu / %zu collide, mismatches outside band %zu, inside %zu\n", hits, states.size(), mismatch_outside, in_band);
		CHECK(mismatch_outside == 0);
		CHECK(hits > 1000 && hits < 19000);
		for (size_t i = 0; i < 200; ++i) CHECK(check_robot_collision(robot, scene, states[i]) == bit(bits, i));
		// per-link entry point
		auto fk = robot_model::forwardKinematics(robot, states[5], flying_base);
		bool any = false;
		for (size_t l = 0; l < fk.link_transforms.size(); ++l) any |= check_link_collision(robot.getLinks()[l], scene, fk.link_transforms[l]);
		CHECK(any == bit(bits, 5));
	}
	// -- motions: boolean, toi and step count equal the oracle's loop -----------------------------------------------------
	{
		random_numbers::RandomNumberGenerator rng(43);
		std::vector<RobotState> a, b;
		for (int i = 0; i < 3000; ++i) a.push_back(generateUniformRandomState(robot, rng, 3.0, 4.0)), b.push_back(generateUniformRandomState(robot, rng, 3.0, 4.0));
		MotionResults r = check_motions(robot, scene, a.data(), b.data(), a.size());
		size_t bad = 0, hits = 0;
		for (size_t i = 0; i < a.size(); ++i) {
			orc::MotionResult want = orc::check_motion_collides(orobot, oscene, to_orc(a[i]), to_orc(b[i]));
			hits += want.hit;
			if (want.n_steps != r.n_steps[i]) ++bad;
			if (want.hit != bit(r.hit_bits, i) || (want.hit && want.toi != r.toi[i])) ++bad;  // grazing cases would show up here
		}
		std::printf("motions: %zu / %zu collide, %zu disagreements\n", hits, a.size(), bad);
		CHECK(bad == 0);
		double toi = -1;
		for (size_t i = 0; i < 50; ++i) {
			bool h = check_motion_collides(robot, scene, a[i], b[i], toi);
			CHECK(h == bit(r.hit_bits, i));
			if (h) CHECK(toi == r.toi[i]);
		}
		// symmetry property (src_old/test/collision_checking_test.cpp:30-57)
		MotionResults rev = check_motions(robot, scene, b.data(), a.data(), a.size());
		size_t asym = 0;
		for (size_t i = 0; i < a.size(); ++i) asym += bit(r.hit_bits, i) != bit(rev.hit_bits, i);
		CHECK(asym <= 2);
	}
	// -- the functional plug API + counting decorators ----------------------------------------------------------------------
	{
		CollisionEnvironment env{robot, scene};
		CollisionFunctions fns = collision_functions_in_environment(env);
		calls_t n_state = 0, n_motion = 0;
		CollisionFunctions counted = collision_functions_in_environment_counting(fns, {n_state, n_motion});
		CHECK(counted.state_collides(base_at(0, 0, 1)));
		CHECK(!counted.state_collides(base_at(0, 9, 1)));
		CHECK(counted.motion_collides(base_at(0, -10, 1), base_at(0, 10, 1)));
		CHECK(n_state == 2 && n_motion == 1);
		auto sampled = motion_collision_check_fn_from_state_collision(fns.state_collides);
		CHECK(sampled(base_at(0, -10, 1), base_at(0, 10, 1)) && !sampled(base_at(5, -10, 1), base_at(5, 10, 1)));
	}
	// -- goal sampling: same result and same generator state as the reference's sequential loop ---------------------------
	{
		const math::Vec3d target{0.45, 0.3, 1.7};
		random_numbers::RandomNumberGenerator rng(42), twin(42);
		auto got = findGoalStateByUniformSampling(target, robot, flying_base, end_effector, scene, rng, 1000);
		std::optional<RobotState> want;
		for (int i = 0; i < 1000 && !want; ++i) {  // goal_sampling.cpp:90-96 with the oracle as collision check
			RobotState s = genGoalStateUniform(twin, target, robot, flying_base, end_effector);
			if (!orc::check_robot_collision(orobot, oscene, to_orc(s))) want = s;
		}
		CHECK(got.has_value() == want.has_value());
		if (got && want) CHECK(*got == *want);
		CHECK(rng.uniformReal(0, 1) == twin.uniformReal(0, 1));  // both generators consumed the same number of draws
		auto ee = robot_model::forwardKinematics(robot, *got, flying_base).forLink(end_effector).translation;
		CHECK((ee - target).norm() < 1e-12);

		std::vector<math::Vec3d> targets{{0.45, 0.3, 1.7}, {-0.6, 0.2, 1.9}, {0.1, -0.7, 1.4}, {3.0, 3.0, 2.0}};
		GoalSamples gs = sample_goal_states(robot, scene, targets, 1000, 42, 0.0, true);
		for (size_t f = 0; f < targets.size(); ++f) {
			random_numbers::RandomNumberGenerator r2(42);  // goal_sample_success_rate.cpp:121: fresh rng(42) per problem
			uint32_t collisions = 0;
			for (size_t s = 0; s < 1000; ++s) {
				RobotState st = genGoalStateUniform(r2, targets[f], robot, flying_base, end_effector);
				if (s < 5) CHECK((st.base_tf.translation - gs.states[f * 1000 + s].base_tf.translation).norm() < 1e-12);
				collisions += orc::check_robot_collision(orobot, oscene, to_orc(st));
			}
			std::printf("goal target %zu: %u collisions (oracle %u)\n", f, gs.collisions[f], collisions);
			CHECK(gs.collisions[f] == collisions);
		}
		CHECK(gs.collisions[3] == 0);  // a target far from the tree is always reachable
	}
	// -- a17 / a19: fromEndEffectorAndVector, arm_vector_from_state, project_to_goal, goal_region_sampler,
	//    try_at_valid_goal_samples, generateUniformRandomArmVectorState against the oracle's restatements (pinned to the
	//    reference's state_tools.cpp / goal_sampling.cpp through oracle/_ref and tests/golden) ------------------------------------
	{
		auto same_state = [](const RobotState &a, const orc::State &b) {
			return a.base_tf.translation.x() == b.base.t.x && a.base_tf.translation.y() == b.base.t.y && a.base_tf.translation.z() == b.base.t.z &&
				   a.base_tf.orientation.x == b.base.q.x && a.base_tf.orientation.y == b.base.q.y && a.base_tf.orientation.z == b.base.q.z &&
				   a.base_tf.orientation.w == b.base.q.w && a.joint_values == b.joints;
		};
		random_numbers::RandomNumberGenerator rng(77);
		for (int i = 0; i < 200; ++i) {
			const math::Vec3d p{rng.uniformReal(-2, 2), rng.uniformReal(-2, 2), rng.uniformReal(0, 3)};
			const math::Vec3d v{rng.gaussian01(), rng.gaussian01(), rng.gaussian01()};
			const RobotState st = fromEndEffectorAndVector(robot, p, v);                                   // state_tools.cpp:14-42
			CHECK(same_state(st, orc::from_end_effector_and_vector(orobot, {p.x(), p.y(), p.z()}, {v.x(), v.y(), v.z()})));
			const math::Vec3d av = arm_vector_from_state(robot, st);                                         // state_tools.cpp:69-75
			const orc::V3 oav = orc::arm_vector_from_state(orobot, to_orc(st));
			CHECK(av.x() == oav.x && av.y() == oav.y && av.z() == oav.z);
			// the arm points back along `v`, and the end effector sits on p
			CHECK((av.normalized() - v.normalized()).norm() < 1e-9);
			CHECK((robot_model::forwardKinematics(robot, st, flying_base).forLink(end_effector).translation - p).norm() < 1e-12);
			// project_to_goal (goal_sampling.cppm:58-79) = the deterministic half of genGoalStateUniform
			RobotState up = genUprightState(robot, rng);
			const RobotState pg = project_to_goal(robot, up, flying_base, end_effector, p);
			CHECK(same_state(pg, orc::project_upright_to_goal(orobot, to_orc(up), {p.x(), p.y(), p.z()}, 0.0, 0, (int) end_effector)));
		}
		// goal_region_sampler (goal_sampling.cppm:93-106): the closure draws from the caller's generator
		random_numbers::RandomNumberGenerator g1(5);
		orc::Rng g2(5);
		auto sampler = goal_region_sampler(robot, g1);
		for (int i = 0; i < 50; ++i) {
			const math::Vec3d t{0.1 * i - 2.0, 0.3, 1.5};
			CHECK(same_state(sampler(t), orc::gen_goal_state_uniform(g2, {t.x(), t.y(), t.z()}, 0.0, orobot, 0, (int) end_effector)));
		}
		// try_at_valid_goal_samples (goal_sampling.cppm:37-56) over the GPU state check vs over the oracle's
		CollisionEnvironment env{robot, scene};
		const CollisionFunctions fns = collision_functions_in_environment(env);
		const orc::StateFn ocollides = [&](const orc::State &st) { return orc::check_robot_collision(orobot, oscene, st); };
		size_t found = 0, none = 0;
		for (int trial = 0; trial < 24; ++trial) {
			const math::Vec3d t{0.5 * std::cos(trial), 0.5 * std::sin(trial), 1.0 + 0.05 * trial};
			random_numbers::RandomNumberGenerator r1(900 + trial);
			orc::Rng r2(900 + trial);
			int ops1 = 0, ops2 = 0;
			// the operation accepts every third valid sample: exercises "valid but operation declined"
			auto got = try_at_valid_goal_samples<RobotState>(
					fns.state_collides, [&] { return genGoalStateUniform(r1, t, robot, flying_base, end_effector); }, 12,
					[&](const RobotState &st) -> std::optional<RobotState> { return ++ops1 % 3 == 0 ? std::optional<RobotState>(st) : std::nullopt; });
			auto want = orc::try_at_valid_goal_samples<orc::State>(
					ocollides, [&] { return orc::gen_goal_state_uniform(r2, {t.x(), t.y(), t.z()}, 0.0, orobot, 0, (int) end_effector); }, 12,
					[&](const orc::State &st) -> std::optional<orc::State> { return ++ops2 % 3 == 0 ? std::optional<orc::State>(st) : std::nullopt; });
			CHECK(got.has_value() == want.has_value() && ops1 == ops2);
			if (got && want) CHECK(same_state(*got, *want));
			CHECK(r1.uniformReal(0, 1) == r2.uniform_real(0, 1));
			found += got.has_value(), none += !got.has_value();
		}
		CHECK(found > 0 && none > 0);
		// generateUniformRandomArmVectorState (goal_sampling.cpp:101-124): batched look-ahead vs the sequential loop
		size_t av_found = 0, av_none = 0;
		for (int trial = 0; trial < 40; ++trial) {
			const math::Vec3d fruit{0.6 * std::cos(0.7 * trial), 0.6 * std::sin(0.7 * trial), 0.8 + 0.03 * trial};
			const int attempts = trial % 4 == 0 ? 2 : 60;
			random_numbers::RandomNumberGenerator r1(300 + trial);
			orc::Rng r2(300 + trial);
			auto got = generateUniformRandomArmVectorState(robot, scene, fruit, r1, attempts, 0.05);
			auto want = orc::generate_uniform_random_arm_vector_state(orobot, ocollides, {fruit.x(), fruit.y(), fruit.z()}, r2, attempts, 0.05);
			CHECK(got.has_value() == want.has_value());
			if (got && want) CHECK(same_state(*got, *want));
			CHECK(r1.uniformReal(0, 1) == r2.uniform_real(0, 1));  // both generators consumed the same draws
			av_found += got.has_value(), av_none += !got.has_value();
		}
		std::printf("goal helpers: try_at_valid_goal_samples %zu found / %zu none; arm-vector states %zu found / %zu none\n", found, none, av_found, av_none);
		CHECK(av_found > 0);
	}
	// -- batched PRM edge validation = the reference's sequential insert-and-connect loop -----------------------------------
	{
		random_numbers::RandomNumberGenerator rng(11);
		std::vector<RobotState> nodes;
		while (nodes.size() < 300) {
			RobotState s = generateUniformRandomState(robot, rng, 3.0, 4.0);
			if (!orc::check_robot_collision(orobot, oscene, to_orc(s))) nodes.push_back(s);
		}
		const size_t k = 5;
		std::vector<RoadmapEdge> edges = validate_prm_edges(robot, scene, nodes, k);
		// reference semantics (tsp_over_prm.cpp:33-66): node i connects to its k nearest among nodes 0..i-1 when the motion is free
		size_t expected = 0, found = 0;
		for (size_t i = 0; i < nodes.size(); ++i) {
			std::vector<std::pair<double, size_t>> d;
			for (size_t j = 0; j < i; ++j) d.push_back({equal_weights_distance(nodes[i], nodes[j]), j});
			std::stable_sort(d.begin(), d.end(), [](auto &x, auto &y) { return x.first < y.first; });
			for (size_t q = 0; q < std::min(k, d.size()); ++q) {
				if (orc::check_motion_collides(orobot, oscene, to_orc(nodes[d[q].second]), to_orc(nodes[i])).hit) continue;
				++expected;
				for (auto &e: edges) found += e.from == d[q].second && e.to == i && std::fabs(e.weight - d[q].first) < 1e-12;
			}
		}
		std::printf("PRM: %zu edges validated on the GPU, %zu expected, %zu matched\n", edges.size(), expected, found);
		CHECK(edges.size() == expected && found == expected);
		// -- goal-to-goal shortest paths on that roadmap (tsp_over_prm.cpp:497-517) against the oracle's Dijkstra ---------
		const PRMGraph graph = roadmap_from_edges(nodes, edges);
		orc::Roadmap ograph(nodes.size());
		for (auto &e: edges) ograph.add_edge((uint32_t) e.from, (uint32_t) e.to, e.weight);
		std::vector<PRMGraph::vertex_descriptor> goals;
		for (size_t v = 3; v < nodes.size(); v += 7) goals.push_back(v);
		const GoalToGoalPathResults g2g = calculate_goal_to_goal_paths(graph, goals);
		CHECK(g2g.distance_lookup.size() == goals.size() && g2g.predecessor_lookup.size() == goals.size());
		size_t bad_dist = 0, bad_pred = 0, reachable = 0;
		auto check_source = [&](uint32_t source, const std::vector<double> &dist, const std::vector<PRMGraph::vertex_descriptor> &pred) {
			std::vector<double> od;
			std::vector<uint32_t> op;
			orc::run_dijkstra(ograph, source, od, op);
			for (size_t v = 0; v < nodes.size(); ++v) {
				bad_dist += dist[v] != od[v];
				reachable += od[v] < std::numeric_limits<double>::max();
				bad_pred += pred[v] != op[v];  // generic weights: no exact ties, so boost, the oracle and the GPU agree
			}
		};
		for (size_t i = 0; i < goals.size(); ++i) check_source((uint32_t) goals[i], g2g.distance_lookup[i], g2g.predecessor_lookup[i]);
		check_source(0, g2g.start_to_goals_distances, g2g.start_to_goals_predecessors);
		std::printf("roadmap: %zu sources x %zu vertices, %zu reachable labels, %zu distance / %zu predecessor mismatches\n", goals.size() + 1,
					nodes.size(), reachable, bad_dist, bad_pred);
		CHECK(bad_dist == 0 && bad_pred == 0 && reachable > nodes.size());
		auto [d0, p0] = runDijkstra(graph, goals[0]);
		CHECK(d0 == g2g.distance_lookup[0] && p0 == g2g.predecessor_lookup[0]);
		// retrace_path: same vertex sequence as the oracle's, start node omitted
		for (size_t v = 0; v < nodes.size(); v += 11) {
			std::vector<uint32_t> op(p0.begin(), p0.end());
			const auto want = orc::retrace_path(op, (uint32_t) v);
			const RobotPath got = retrace_path(graph, p0, v);
			bool same = got.states.size() == want.size();
			for (size_t i = 0; same && i < want.size(); ++i) same = got.states[i] == nodes[want[i]];
			CHECK(same);
		}
	}
	// -- path shortcutting: batched drivers == the reference's sequential loops (oracle as motion check) ------------------
	{
		auto same_state = [](const RobotState &s, const orc::State &o) {
			return s.base_tf.translation.x() == o.base.t.x && s.base_tf.translation.y() == o.base.t.y && s.base_tf.translation.z() == o.base.t.z &&
				   s.base_tf.orientation.x == o.base.q.x && s.base_tf.orientation.y == o.base.q.y && s.base_tf.orientation.z == o.base.q.z &&
				   s.base_tf.orientation.w == o.base.q.w && s.joint_values == o.joints;
		};
		auto same_path = [&](const RobotPath &p, const orc::Path &o) {
			if (p.states.size() != o.size()) return false;
			for (size_t i = 0; i < o.size(); ++i)
				if (!same_state(p.states[i], o[i])) return false;
			return true;
		};
		const orc::MotionFn omotion = [&](const orc::State &a, const orc::State &b) { return orc::check_motion_collides(orobot, oscene, a, b).hit; };
		const CollisionEnvironment env{robot, scene};
		const CollisionFunctions fns = collision_functions_in_environment(env);
		size_t total_success = 0, total_batches = 0, total_attempts = 0;
		for (unsigned seed = 0; seed < 6; ++seed) {
			random_numbers::RandomNumberGenerator make(300 + seed);
			RobotPath path0;
			for (int i = 0; i < 14; ++i) path0.append(generateUniformRandomState(robot, make, 1.8, 2.6));
			orc::Path opath0;
			for (auto &st: path0.states) opath0.push_back(to_orc(st));
			// PathPoint arithmetic and interpolation, bit for bit
			for (double d: {0.0, 0.25, 3.75, 12.999, 13.5, 40.0}) {
				PathPoint p = PathPoint::fromScalar(d, path0);
				orc::PathPoint q = orc::PathPoint::from_scalar(d, opath0);
				CHECK(p.segment_i == q.segment_i && p.segment_t == q.segment_t && p.toScalar() == q.to_scalar());
				if (p.segment_t <= 1.0) CHECK(same_state(interpolate(p, path0), orc::interpolate(q, opath0)));
			}
			{  // globally: tsp_over_prm.cpp:594-612
				RobotPath path = path0;
				orc::Path opath = opath0;
				random_numbers::RandomNumberGenerator rng(seed + 11);
				orc::Rng orng(seed + 11);
				const size_t want = orc::optimize_path_segment(opath, omotion, orng, 60);
				const ShortcutStats st = shortcut_path_globally(robot, scene, path, rng, 60);
				CHECK(st.successes == want && same_path(path, opath));
				CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));  // generator left where the sequential loop leaves it
				CHECK(st.batches <= st.successes + 1 && st.attempts <= 60);
				total_success += st.successes, total_batches += st.batches, total_attempts += st.attempts;
				// the one-attempt-at-a-time API through the planner-facing closure gives the same path
				RobotPath seq = path0;
				random_numbers::RandomNumberGenerator rng2(seed + 11);
				for (int i = 0; i < 60 && seq.states.size() > 2; ++i) tryShortcuttingRandomlyGlobally(seq, fns.motion_collides, rng2);
				CHECK(seq.states == path.states);
			}
			{  // locally
				RobotPath path = path0;
				orc::Path opath = opath0;
				random_numbers::RandomNumberGenerator rng(seed + 23);
				orc::Rng orng(seed + 23);
				size_t want = 0;
				for (int i = 0; i < 40 && opath.size() > 2; ++i) want += orc::try_shortcutting_randomly_locally(opath, orng, omotion);
				const ShortcutStats st = shortcut_path_locally(robot, scene, path, rng, 40);
				CHECK(st.successes == want && same_path(path, opath));
				CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));
			}
			{  // deleting waypoints: closure overload and batched scene overload
				RobotPath a = path0, b = path0;
				orc::Path opath = opath0;
				const bool want = orc::try_deleting_every_waypoint(opath, omotion);
				ShortcutStats st;
				CHECK(tryDeletingEveryWaypoint(a, fns.motion_collides) == want && same_path(a, opath));
				CHECK(tryDeletingEveryWaypoint(robot, b, scene, &st) == want && same_path(b, opath));
				CHECK(st.batches <= st.successes + 1);
			}
			{  // midpoint pull sweep
				RobotPath path = path0;
				orc::Path opath = opath0;
				for (size_t i = 1; i + 1 < opath.size(); ++i)
					CHECK(tryMidpointPull(path, i, 0.5, fns.motion_collides) == orc::try_midpoint_pull(opath, i, 0.5, omotion));
				CHECK(same_path(path, opath));
			}
		}
		std::printf("shortcutting: %zu attempts, %zu successes, %zu GPU batches\n", total_attempts, total_success, total_batches);
		CHECK(total_success > 0 && total_batches < total_attempts);
	}
	// -- tree-model ingest: COLLADA -> Mesh -> fruit components -> scene ------------------------------------------------------
	{
		namespace fs = std::filesystem;
		const std::string dir = (fs::temp_directory_path() / "mgodpl_b200_models").string() + "/";
		fs::create_directories(dir);
		Mesh fruit;
		std::vector<math::Vec3d> centres;
		random_numbers::RandomNumberGenerator frng(5);
		for (int i = 0; i < 23; ++i) {
			centres.push_back({frng.uniformReal(-1.2, 1.2), frng.uniformReal(-1.2, 1.2), frng.uniformReal(1.0, 2.6)});
			add_blob(fruit, centres.back(), 0.04, 6, 8);                                   // 42 vertices: a fruit
			add_blob(fruit, centres.back() + math::Vec3d{0, 0, 0.06}, 0.005, 3, 5);      // 12 vertices: its stem, to be dropped
		}
		Mesh leaves;
		for (int i = 0; i < 40; ++i) {
			math::Vec3d c{frng.uniformReal(-1, 1), frng.uniformReal(-1, 1), frng.uniformReal(1, 2.5)};
			size_t b = leaves.vertices.size();
			leaves.vertices.insert(leaves.vertices.end(), {c, c + math::Vec3d{0.05, 0, 0}, c + math::Vec3d{0.05, 0.03, 0.01}});
			leaves.triangles.push_back({b, b + 1, b + 2});
		}
		write_dae(dir + "synthetic_trunk.dae", tree, 0);
		write_dae(dir + "synthetic_leaves.dae", leaves, 1);
		write_dae(dir + "synthetic_apples.dae", fruit, 2);  // the legacy file name (TreeMeshes.cpp:38-47)
		const auto names = tree_meshes::getTreeModelNames(dir);
		CHECK(names.size() == 1 && names[0] == "synthetic");
		const tree_meshes::TreeMeshes loaded = tree_meshes::loadTreeMeshes("synthetic", dir);
		// identical positions joined, triangles kept, coordinates back in metres (float32 precision), z up
		CHECK(loaded.trunk_mesh.vertices.size() == tree.vertices.size() && loaded.trunk_mesh.triangles.size() == tree.triangles.size());
		double worst = 0;
		for (size_t t = 0; t < tree.triangles.size() && t < loaded.trunk_mesh.triangles.size(); ++t)
			for (int k = 0; k < 3; ++k)
				worst = std::max(worst, (loaded.trunk_mesh.vertices[loaded.trunk_mesh.triangles[t][k]] - tree.vertices[tree.triangles[t][k]]).norm());
		CHECK(worst < 1e-6);
		CHECK(loaded.leaves_mesh.triangles.size() == leaves.triangles.size());
		// fruit: 23 components of 42 vertices survive, 23 stems (12 vertices) are dropped; positions = AABB centres
		CHECK(loaded.fruit_meshes.size() == centres.size());
		const auto positions = tree_meshes::computeFruitPositions(loaded);
		size_t matched = 0;
		for (size_t i = 0; i < positions.size() && i < centres.size(); ++i)
			matched += loaded.fruit_meshes[i].vertices.size() == 42 && (positions[i] - centres[i]).norm() < 1e-5;  // components come in first-vertex order
		CHECK(matched == centres.size());
		// components agree with the reference's merging rule (oracle restatement, pinned against mesh_connected_components.cpp)
		{
			std::vector<uint32_t> t32;
			for (auto &t: fruit.triangles) t32.insert(t32.end(), {(uint32_t) t[0], (uint32_t) t[1], (uint32_t) t[2]});
			const auto want = orc::mesh_component_labels(t32.data(), fruit.triangles.size(), fruit.vertices.size());
			const auto cc = connected_vertex_components(fruit);
			size_t agree = 0;
			for (auto &c: cc)
				for (size_t v: c) agree += want[v] == c.front();
			CHECK(cc.size() == 46 && agree == fruit.vertices.size());
		}
		// the loaded trunk answers collision queries like the mesh it was written from
		const gpu::Scene loaded_scene = gpu::treeMeshesToScene(loaded);
		random_numbers::RandomNumberGenerator qrng(77);
		std::vector<RobotState> qs;
		for (int i = 0; i < 2000; ++i) qs.push_back(generateUniformRandomState(robot, qrng, 1.5, 2.5));
		const auto b0 = check_states(robot, scene, qs), b1 = check_states(robot, loaded_scene, qs);
		size_t differ = 0;
		for (size_t i = 0; i < qs.size(); ++i) differ += bit(b0, i) != bit(b1, i);
		CHECK(differ <= 2);  // float32 storage moves vertices by < 1e-6 m: only grazing contacts may flip
		// orchard assembly: two metres apart, centred (TreeMeshes.cpp:197-206), one merged scene
		std::vector<tree_meshes::TreeMeshes> models{loaded, loaded, loaded};
		const auto orchard = tree_meshes::makeSingleRowOrchard(models);
		CHECK(orchard.trees.size() == 3 && orchard.trees[0].first.x() == -3.0 && orchard.trees[2].first.x() == 1.0);
		const gpu::Scene orchard_scene = gpu::orchardToScene(orchard);
		CHECK(orchard_scene.info().n_triangles == 3 * tree.triangles.size());
		CHECK(check_robot_collision(robot, orchard_scene, base_at(-3.0, 0, 0.8)) && check_robot_collision(robot, orchard_scene, base_at(1.0, 0, 0.8)) &&
			  !check_robot_collision(robot, orchard_scene, base_at(-3.0, 0, 6.0)));
		bool threw = false;
		try {
			from_dae(dir + "missing_trunk.dae");
		} catch (const std::runtime_error &e) { threw = std::string(e.what()).find("Failed to load mesh from") == 0; }
		CHECK(threw);
		fs::remove_all(dir);
	}
	// -- roadmap construction in batches == the reference's node-by-node loops (tsp_over_prm.cpp:33-66, 244-344, 411-483) ------
	{
		auto same_state = [](const RobotState &s, const orc::State &o) {
			return s.base_tf.translation.x() == o.base.t.x && s.base_tf.translation.y() == o.base.t.y && s.base_tf.translation.z() == o.base.t.z &&
				   s.base_tf.orientation.x == o.base.q.x && s.base_tf.orientation.y == o.base.q.y && s.base_tf.orientation.z == o.base.q.z &&
				   s.base_tf.orientation.w == o.base.q.w && s.joint_values == o.joints;
		};
		struct Edge {
			size_t u, v;
			double w;
		};
		const size_t max_samples = 400, k = 5;
		const GoalSampleParams gp{4, 3, 40};
		std::vector<math::Vec3d> fruit;
		random_numbers::RandomNumberGenerator frng(9);
		for (int i = 0; i < 12; ++i) fruit.push_back({frng.uniformReal(-1.0, 1.0), frng.uniformReal(-1.0, 1.0), frng.uniformReal(0.8, 2.2)});
		// the reference, one node at a time, with the oracle as state / motion check and an exhaustive nearest-neighbour search
		orc::Rng orng(31);
		std::vector<orc::State> onodes;
		std::vector<Edge> oedges;
		size_t n_infra = 0;
		auto connect = [&](const orc::State &s, size_t kk) {
			std::vector<std::pair<double, size_t>> d;
			for (size_t j = 0; j < n_infra; ++j) d.push_back({orc::equal_weights_distance(s, onodes[j]), j});
			std::stable_sort(d.begin(), d.end(), [](auto &x, auto &y) { return x.first < y.first; });
			const size_t me = onodes.size();
			for (size_t q = 0; q < std::min(kk, d.size()); ++q)
				if (!orc::check_motion_collides(orobot, oscene, onodes[d[q].second], s).hit) oedges.push_back({me, d[q].second, d[q].first});
			onodes.push_back(s);
		};
		for (size_t i = 0; i < max_samples; ++i) {
			const orc::State s = orc::generate_uniform_random_state(orobot, orng, 2.0, 3.0);
			if (orc::check_robot_collision(orobot, oscene, s)) continue;
			connect(s, k);
			n_infra = onodes.size();  // infrastructure nodes join the index
		}
		std::vector<size_t> want_groups(fruit.size(), 0);
		const int ofb = orobot.find_link("flying_base"), oee = orobot.find_link("end_effector");
		for (size_t g = 0; g < fruit.size(); ++g)
			for (size_t attempt = 0; attempt < gp.max_attempts && want_groups[g] < gp.max_valid_samples; ++attempt) {
				orc::State s = orc::gen_goal_state_uniform(orng, {fruit[g].x(), fruit[g].y(), fruit[g].z()}, 0.0, orobot, ofb, oee);
				if (orc::check_robot_collision(orobot, oscene, s)) continue;
				s.joints.resize(onodes[0].joints.size(), 0.0);  // defined stand-in for the reference's out-of-bounds read (see prm.hpp)
				connect(s, gp.k_neighbors);  // goal samples are not added to the index
				++want_groups[g];
			}
		// the facade: a handful of GPU calls
		random_numbers::RandomNumberGenerator rng(31);
		PRM prm = build_infrastructure_roadmap(robot, scene, rng, max_samples, k, 2.0, 3.0);
		CHECK(prm.n_infrastructure == n_infra && prm.graph.num_vertices() == n_infra);
		const auto [goal_nodes, group_sizes] = sample_goal_states(fruit, prm, gp, robot, scene, rng);
		CHECK(group_sizes == want_groups && prm.graph.num_vertices() == onodes.size());
		CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));
		size_t same_vertices = 0, same_edges = 0, n_goals = 0;
		for (size_t v = 0; v < std::min(onodes.size(), prm.graph.num_vertices()); ++v) same_vertices += same_state(prm.graph[v], onodes[v]);
		for (size_t e = 0; e < std::min(oedges.size(), prm.graph.num_edges()); ++e)
			same_edges += prm.graph.edges[2 * e] == oedges[e].u && prm.graph.edges[2 * e + 1] == oedges[e].v &&
						  std::fabs(prm.graph.weights[e] - oedges[e].w) < 1e-12;  // device vs host libm: last-ulp differences in acos
		for (size_t g: group_sizes) n_goals += g;
		std::printf("roadmap construction: %zu infrastructure + %zu goal vertices, %zu edges (reference loop: %zu), %zu identical\n", n_infra, n_goals,
					prm.graph.num_edges(), oedges.size(), same_edges);
		CHECK(same_vertices == onodes.size() && prm.graph.num_edges() == oedges.size() && same_edges == oedges.size());
		CHECK(goal_nodes.size() == n_goals && n_goals > 0 && (goal_nodes.empty() || goal_nodes.front() == n_infra));
		// the start state joins like a goal sample (tsp_over_prm.cpp:621-629); then the planner's shortest-path tables come from one call
		const auto start = add_and_connect_roadmap_node(base_at(1.9, 1.9, 2.5), prm, k, robot, scene);
		CHECK(start == onodes.size());
		const GoalToGoalPathResults g2g = calculate_goal_to_goal_paths(prm.graph, goal_nodes);
		CHECK(g2g.distance_lookup.size() == goal_nodes.size() && g2g.start_to_goals_distances.size() == prm.graph.num_vertices());
		// construct_final_path (tsp_over_prm.cpp:189-234) over a fixed visiting order (the TSP itself is OR-Tools, out of scope):
		// legs retraced and shortcut on the GPU == the oracle's Dijkstra, retrace and sequential shortcutting loop
		std::vector<size_t> tour;
		for (size_t i = 0; i < goal_nodes.size(); i += 3) tour.push_back(i);
		const RobotPath final_path = construct_final_path(prm.graph, g2g.predecessor_lookup, g2g.start_to_goals_predecessors, goal_nodes, tour,
														  optimize_path_segment_fn(robot, scene, rng, 40));
		orc::Roadmap ograph(prm.graph.num_vertices());
		for (size_t e = 0; e < prm.graph.num_edges(); ++e) ograph.add_edge(prm.graph.edges[2 * e], prm.graph.edges[2 * e + 1], prm.graph.weights[e]);
		const orc::MotionFn omotion = [&](const orc::State &a, const orc::State &b) { return orc::check_motion_collides(orobot, oscene, a, b).hit; };
		onodes.push_back(to_orc(prm.graph[start]));
		orc::Path want_path;
		for (size_t i = 0; i < tour.size(); ++i) {
			std::vector<double> od;
			std::vector<uint32_t> op;
			orc::run_dijkstra(ograph, i ? (uint32_t) goal_nodes[tour[i - 1]] : 0u, od, op);
			orc::Path leg;
			for (uint32_t v: orc::retrace_path(op, (uint32_t) goal_nodes[tour[i]])) leg.push_back(onodes[v]);
			if (leg.size() > 2) orc::optimize_path_segment(leg, omotion, orng, 40);
			want_path.insert(want_path.end(), leg.begin(), leg.end());
		}
		bool same = final_path.states.size() == want_path.size();
		for (size_t i = 0; same && i < want_path.size(); ++i) same = same_state(final_path.states[i], want_path[i]);
		std::printf("final path: %zu legs, %zu waypoints (reference pipeline: %zu)\n", tour.size(), final_path.states.size(), want_path.size());
		CHECK(same && !want_path.empty());
		CHECK(rng.uniformReal(0, 1) == orng.uniform_real(0, 1));
	}
	// -- RRT (rrt.cppm:82-188): the windowed driver grows the tree of the sequential loop ---------------------------------
	{
		const RobotState root = base_at(1.9, 1.9, 2.5);
		const auto sampler = [&](random_numbers::RandomNumberGenerator &g) { return generateUniformRandomState(robot, g, 2.0, 3.0); };
		const DistanceFn dist = [](const RobotState &a, const RobotState &b) { return equal_weights_distance(a, b); };
		for (size_t stop_at: {size_t(0), size_t(45)}) {
			// reference loop with the oracle as state / motion check
			random_numbers::RandomNumberGenerator rng_seq(77);
			std::vector<RRTNode> want;
			rrt(root, [&] { return sampler(rng_seq); }, [&](const RobotState &s) { return orc::check_robot_collision(orobot, oscene, to_orc(s)); },
				[&](const RobotState &a, const RobotState &b) { return orc::check_motion_collides(orobot, oscene, to_orc(a), to_orc(b)).hit; }, dist, 400,
				[&](const std::vector<RRTNode> &nodes) {
					want = nodes;
					return stop_at && nodes.size() >= stop_at;
				});
			random_numbers::RandomNumberGenerator rng_b(77);
			RrtStats st;
			const std::vector<RRTNode> got = rrt_batched(root, sampler, rng_b, robot, scene, 400,
														 [&](const std::vector<RRTNode> &nodes) { return stop_at && nodes.size() >= stop_at; }, 64, &st);
			bool same = got.size() == want.size();
			for (size_t i = 0; same && i < want.size(); ++i) same = got[i].state == want[i].state && got[i].parent_index == want[i].parent_index;
			std::printf("rrt: %zu nodes after %zu samples, %zu GPU calls, %zu re-speculated (sequential loop: %zu nodes)\n", got.size(), st.iterations,
						st.gpu_calls, st.respeculated, want.size());
			CHECK(same && want.size() > 20);
			CHECK(rng_b.uniformReal(0, 1) == rng_seq.uniformReal(0, 1));
			CHECK(st.gpu_calls < st.iterations);  // the sequential loop makes up to two calls per sample
			if (stop_at) {
				const RobotPath back = retrace(got);
				CHECK(back.states.front() == got.back().state && back.states.back() == root);
			}
		}
		// the closure-based rrt on the GPU closures and rrt_path_to_acceptable
		const CollisionEnvironment env{robot, scene};
		const CollisionFunctions fns = collision_functions_in_environment(env);
		random_numbers::RandomNumberGenerator rng_c(77);
		const auto path = rrt_path_to_acceptable(root, [&] { return sampler(rng_c); }, fns.state_collides, fns.motion_collides, dist, 60,
												 [](const RobotState &s) { return s.base_tf.translation.z() < 0.6; });
		CHECK(!path || (path->states.back() == root && path->states.front().base_tf.translation.z() < 0.6));
	}
	// -- scanning motions (scanning_motions.cpp:15-97): one batch per arc == the reference's check-and-stop loop ---------
	{
		size_t arcs = 0, states_total = 0, cut_short = 0;
		random_numbers::RandomNumberGenerator srng(3);
		for (int trial = 0; trial < 12; ++trial) {
			const math::Vec3d fruit{srng.uniformReal(-0.9, 0.9), srng.uniformReal(-0.9, 0.9), srng.uniformReal(0.9, 2.0)};
			const double lon = srng.uniformReal(-M_PI, M_PI), lat = srng.uniformReal(-0.3, 0.6), scan_distance = 0.4;
			for (Direction dir: {Direction::LEFT, Direction::RIGHT}) {
				const RobotPath got = generateSidewaysScanningMotion(robot, scene, fruit, lon, lat, scan_distance, dir);
				// reference loop with the oracle as collision check
				std::vector<RobotState> want;
				const double step = (dir == Direction::LEFT ? -1.0 : 1.0) * M_PI / 32.0;
				for (int i = 1; i < 32; ++i) {
					const math::Vec3d rel = spherical_geometry::RelativeVertex{lon + i * step, lat}.to_cartesian();
					const RobotState st = fromEndEffectorAndVector(robot, fruit + rel.normalized() * scan_distance, rel);
					if (orc::check_robot_collision(orobot, oscene, to_orc(st))) break;
					want.push_back(st);
				}
				CHECK(got.states == want);
				++arcs, states_total += want.size(), cut_short += want.size() < 31;
			}
			const RobotPath both = createLeftRightScanningMotion(robot, scene, fruit, lon, lat, scan_distance);
			CHECK(both.states.size() % 2 == 0);
		}
		std::printf("scanning motions: %zu arcs, %zu states kept, %zu arcs cut at a collision\n", arcs, states_total, cut_short);
		CHECK(cut_short > 0 && cut_short < arcs);
	}
	// -- one process, several GPUs: a DeviceGroup shards every batch over its scene replicas (one per device; on a single-GPU
	//    box the same device is listed twice, which exercises the same threads, shard bounds and in-place merging) -----------------
	{
		std::vector<int> devs;
		for (int d = 0; d < std::min(mgodpl_device_count(), 4); ++d) devs.push_back(d);
		if (devs.size() < 2) devs = {0, 0, 0};
		const gpu::DeviceGroup group(tree, devs);
		CHECK(group.size() == devs.size());
		int dev = -1;
		CHECK(mgodpl_scene_device(group.scene(group.size() - 1).handle(), &dev) == MGODPL_OK && dev == devs.back());
		random_numbers::RandomNumberGenerator rng(2024);
		std::vector<RobotState> st;
		for (int i = 0; i < 5003; ++i) st.push_back(generateUniformRandomState(robot, rng, 2.5, 3.5));
		CHECK(group.check_states(robot, st) == check_states(robot, scene, st));
		const size_t m = 2001;
		const MotionResults one = check_motions(robot, scene, st.data(), st.data() + m, m), many = group.check_motions(robot, st.data(), st.data() + m, m);
		CHECK(one.hit_bits == many.hit_bits && one.n_steps == many.n_steps);
		size_t toi_diff = 0;
		for (size_t i = 0; i < m; ++i) toi_diff += !(one.toi[i] == many.toi[i] || (std::isnan(one.toi[i]) && std::isnan(many.toi[i])));
		CHECK(toi_diff == 0);
		// the motion shards are balanced by predicted state checks, and the prediction is the step count the device uses
		std::vector<double> costs(m);
		for (size_t i = 0; i < m; ++i) costs[i] = gpu::predicted_motion_cost(st[i], st[m + i]);
		for (size_t i = 0; i < m; ++i) CHECK(costs[i] == (double) one.n_steps[i] + 1.0);
		const auto bounds = gpu::balanced_bounds(costs, 3);
		CHECK(bounds.front().first == 0 && bounds.back().second == m && bounds[0].second % 32 == 0 && bounds[0].second == bounds[1].first);
		std::vector<math::Vec3d> targets;
		for (int i = 0; i < 7; ++i) targets.push_back({0.5 * std::cos(i), 0.5 * std::sin(i), 1.2 + 0.1 * i});
		const GoalSamples g1 = sample_goal_states(robot, scene, targets, 300, 42, 0.0, true), gn = group.sample_goal_states(robot, targets, 300, 42, 0.0, true);
		CHECK(g1.hit_bits == gn.hit_bits && g1.collisions == gn.collisions && g1.states.size() == gn.states.size());
		for (size_t i = 0; i < g1.states.size(); i += 97) CHECK(g1.states[i] == gn.states[i]);
		std::printf("device group: %zu replicas on devices", group.size());
		for (size_t r = 0; r < group.size(); ++r) std::printf(" %d", group.device(r));
		std::printf("\n");
	}
	// -- capsule and convex-polytope links (this library's additions to the Shape variant) against the oracle's definitions ---------
	{
		using RM = robot_model::RobotModel;
		const std::vector<math::Vec3d> wedge{{-0.08, -0.1, -0.05}, {0.08, -0.1, -0.05}, {0.08, 0.1, -0.05}, {-0.08, 0.1, -0.05}, {0.0, 0.12, 0.09}, {0.0, -0.06, 0.07}};
		const math::Transformd lying{{0, 0, 0}, math::Quaterniond::fromAxisAngle({1, 0, 0}, M_PI / 2)};  // capsule axis along the link's y
		RM shaped;
		auto base = shaped.insertLink(RM::Link::namedBox("flying_base", {0.8, 0.8, 0.2}));
		auto arm = shaped.insertLink({"stick", {}, {PositionedShape{Capsule{0.04, 0.7}, lying}}, {}});
		auto ee = shaped.insertLink({"end_effector", {}, {PositionedShape::untransformed(ConvexMesh{wedge})}, {}});
		shaped.insertJoint({"j0", math::Transformd{{0, 0.4, 0}, {0, 0, 0, 1}}, math::Transformd{{0, -0.4, 0}, {0, 0, 0, 1}}, base, arm, RM::RevoluteJoint{{1, 0, 0}, -1.5, 1.5}});
		shaped.insertJoint({"j1", math::Transformd{{0, 0.4, 0}, {0, 0, 0, 1}}, math::Transformd{{0, 0, 0}, {0, 0, 0, 1}}, arm, ee, RM::RevoluteJoint{{0, 0, 1}, -1.0, 1.0}});
		orc::Robot oshaped;
		orc::BoxGeom obase;
		obase.size = {0.8, 0.8, 0.2}, obase.tf = orc::tf_identity();
		oshaped.add_link({"flying_base", {}, {obase}});
		orc::BoxGeom cap;
		cap.size = {0, 0, 0}, cap.tf = {{0, 0, 0}, {lying.orientation.x, lying.orientation.y, lying.orientation.z, lying.orientation.w}};
		cap.kind = orc::GEOM_CAPSULE, cap.radius = 0.04, cap.length = 0.7;
		oshaped.add_link({"stick", {}, {cap}});
		orc::BoxGeom cvx;
		cvx.size = {0, 0, 0}, cvx.tf = orc::tf_identity();
		cvx.kind = orc::GEOM_CONVEX;
		for (auto &v: wedge) cvx.verts.push_back({v.x(), v.y(), v.z()});
		oshaped.add_link({"end_effector", {}, {cvx}});
		orc::Joint j0, j1;
		j0.attach_a = {{0, 0.4, 0}, {0, 0, 0, 1}}, j0.attach_b = {{0, -0.4, 0}, {0, 0, 0, 1}}, j0.link_a = 0, j0.link_b = 1, j0.kind = orc::REVOLUTE, j0.axis = {1, 0, 0}, j0.min_angle = -1.5, j0.max_angle = 1.5;
		j1.attach_a = {{0, 0.4, 0}, {0, 0, 0, 1}}, j1.attach_b = {{0, 0, 0}, {0, 0, 0, 1}}, j1.link_a = 1, j1.link_b = 2, j1.kind = orc::REVOLUTE, j1.axis = {0, 0, 1}, j1.min_angle = -1.0, j1.max_angle = 1.0;
		oshaped.add_joint(j0), oshaped.add_joint(j1);
		random_numbers::RandomNumberGenerator rng(7);
		std::vector<RobotState> states;
		for (int i = 0; i < 8000; ++i) states.push_back(generateUniformRandomState(shaped, rng, 2.0, 3.0));
		auto bits = check_states(shaped, scene, states);
		size_t mismatch_outside = 0, hits = 0;
		for (size_t i = 0; i < states.size(); ++i) {
			const orc::State os = to_orc(states[i]);
			const bool want = orc::check_robot_collision(oshaped, oscene, os);
			hits += want;
			if (bit(bits, i) != want && std::fabs(orc::robot_clearance(oshaped, oscene, os, 1e-3)) > 1e-6) ++mismatch_outside;
		}
		std::printf("capsule + convex links: %zu / %zu collide, mismatches outside band %zu\n", hits, states.size(), mismatch_outside);
		CHECK(mismatch_outside == 0);
		CHECK(hits > 300 && hits < 7700);
	}
	// -- error vocabulary -------------------------------------------------------------------------------------------------------
	{
		robot_model::RobotModel bad;
		bad.insertLink({"flying_base", {}, {PositionedShape::untransformed(Mesh{})}, {}});
		bool threw = false;
		try {
			check_robot_collision(bad, scene, base_at(0, 0, 1));
		} catch (const std::runtime_error &e) { threw = std::string(e.what()).find("Only boxes are implemented") != std::string::npos; }
		CHECK(threw);
		robot_model::RobotModel nameless;
		nameless.insertLink(robot_model::RobotModel::Link::namedBox("base", {1, 1, 1}));
		threw = false;
		try {
			check_robot_collision(nameless, scene, base_at(0, 0, 1));
		} catch (const std::runtime_error &e) { threw = std::string(e.what()) == "Link not found: flying_base"; }
		CHECK(threw);
	}
	// -- concurrent callers on one shared scene (the reference runs this path from TBB workers) ---------------------------
	{
		std::vector<std::thread> pool;
		std::vector<int> ok(6, 0);
		for (int t = 0; t < 6; ++t)
			pool.emplace_back([&, t] {
				random_numbers::RandomNumberGenerator rng(100 + t);
				int good = 0;
				for (int i = 0; i < 150; ++i) {
					RobotState s = generateUniformRandomState(robot, rng, 2.0, 3.0);
					good += check_robot_collision(robot, scene, s) == orc::check_robot_collision(orobot, oscene, to_orc(s));
				}
				ok[t] = good;
			});
		for (auto &th: pool) th.join();
		for (int t = 0; t < 6; ++t) CHECK(ok[t] >= 149);
	}
	std::printf("%d checks, %d failed\n", g_checked, g_failed);
	return g_failed ? 1 : 0;
}
