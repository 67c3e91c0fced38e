// ORACLE — TEST INFRASTRUCTURE ONLY (see mgodpl_oracle.hpp).  Plain C entry points over the fp64 restatement so
// that tests/ and bench.py's cpu_baseline leg can drive it through ctypes.  Nothing in the product links this.
#include "mgodpl_oracle.hpp"
#include "mpr_second_opinion.hpp"

#include <atomic>
#include <cstring>
#include <thread>

using namespace orc;

namespace {
State load_state(const double *p, int js) {
	State s;
	s.base.t = {p[0], p[1], p[2]};
	s.base.q = {p[3], p[4], p[5], p[6]};
	s.joints.assign(p + 7, p + 7 + js);
	return s;
}
void store_state(const State &s, double *p) {
	p[0] = s.base.t.x, p[1] = s.base.t.y, p[2] = s.base.t.z;
	p[3] = s.base.q.x, p[4] = s.base.q.y, p[5] = s.base.q.z, p[6] = s.base.q.w;
	for (size_t i = 0; i < s.joints.size(); ++i) p[7 + i] = s.joints[i];
}
Tf load_tf(const double *p) { return {{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}}; }

template<class F>
void parallel_for(size_t n, int threads, F f) {
	if (threads <= 1 || n < 2) {
		f(0, n, 0);
		return;
	}
	std::vector<std::thread> pool;
	size_t chunk = (n + threads - 1) / threads;
	for (int t = 0; t < threads; ++t) {
		size_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
		if (lo < hi) pool.emplace_back(f, lo, hi, t);
	}
	for (auto &th: pool) th.join();
}
}  // namespace

extern "C" {

// ---- robots ----------------------------------------------------------------------------------------------
void *orc_robot_procedural(double arm_length, const int *joint_types, int n_joint_types) {
	return new Robot(procedural_robot(arm_length, std::vector<int>(joint_types, joint_types + n_joint_types)));
}
/// Generic builder mirroring RobotModel::insertLink / insertJoint (RobotModel.cpp:96-112).
void *orc_robot_new() { return new Robot(); }
int orc_robot_add_link(void *r, const char *name) { return ((Robot *) r)->add_link({name, {}, {}}); }
void orc_robot_add_box(void *r, int link, const double *size3, const double *tf7) {
	BoxGeom g;
	g.size = {size3[0], size3[1], size3[2]}, g.tf = load_tf(tf7);
	((Robot *) r)->links[link].boxes.push_back(g);
}
/// Link shapes the reference does not have (capsule: fcl::Capsule's convention, axis along the shape's z; convex: hull of the points).
void orc_robot_add_capsule(void *r, int link, double radius, double length, const double *tf7) {
	BoxGeom g;
	g.size = {2 * radius, 2 * radius, length + 2 * radius}, g.tf = load_tf(tf7);
	g.kind = GEOM_CAPSULE, g.radius = radius, g.length = length;
	((Robot *) r)->links[link].boxes.push_back(g);
}
void orc_robot_add_convex(void *r, int link, const double *verts, int n, const double *tf7) {
	BoxGeom g;
	g.size = {0, 0, 0}, g.tf = load_tf(tf7);
	g.kind = GEOM_CONVEX;
	for (int i = 0; i < n; ++i) g.verts.push_back({verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]});
	((Robot *) r)->links[link].boxes.push_back(g);
}
/// One shape against one triangle (world frame): kind 1 = capsule (params radius, length), 2 = convex (verts); returns hit,
/// *clearance = signed clearance.
int orc_shape_tri(int kind, const double *tf7, double radius, double length, const double *verts, int n, const double *tri9, double *clearance) {
	BoxGeom g;
	g.size = {0, 0, 0}, g.tf = tf_identity();
	g.kind = kind, g.radius = radius, g.length = length;
	for (int i = 0; i < n; ++i) g.verts.push_back({verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]});
	std::array<V3, 3> t{V3{tri9[0], tri9[1], tri9[2]}, V3{tri9[3], tri9[4], tri9[5]}, V3{tri9[6], tri9[7], tri9[8]}};
	const double cl = geom_tri_clearance(load_tf(tf7), g, t);
	if (clearance) *clearance = cl;
	return cl <= 0.0;
}
int orc_robot_add_joint(void *r, int link_a, int link_b, const double *attach_a7, const double *attach_b7, int kind,
						const double *axis3, double min_angle, double max_angle) {
	Joint j;
	j.attach_a = load_tf(attach_a7);
	j.attach_b = load_tf(attach_b7);
	j.link_a = link_a;
	j.link_b = link_b;
	j.kind = kind;
	j.axis = {axis3[0], axis3[1], axis3[2]};
	j.min_angle = min_angle;
	j.max_angle = max_angle;
	return ((Robot *) r)->add_joint(j);
}
void orc_robot_free(void *r) { delete (Robot *) r; }
int orc_robot_n_links(void *r) { return (int) ((Robot *) r)->links.size(); }
int orc_robot_n_joints(void *r) { return (int) ((Robot *) r)->joints.size(); }
int orc_robot_n_variables(void *r) { return ((Robot *) r)->n_variables(); }
int orc_robot_find_link(void *r, const char *name) {
	try {
		return ((Robot *) r)->find_link(name);
	} catch (...) { return -1; }
}
/// Flat description of the robot, for handing the same model to the C ABI under test.
/// boxes: per box [link, sx,sy,sz, tf7] (11 doubles); joints: per joint [link_a, link_b, kind, a7, b7, axis3, min, max] (22).
int orc_robot_describe(void *rp, double *boxes, int max_boxes, double *joints, int max_joints) {
	Robot &r = *(Robot *) rp;
	int nb = 0;
	for (size_t l = 0; l < r.links.size(); ++l)
		for (auto &g: r.links[l].boxes) {
			if (g.kind != GEOM_BOX) continue;  // capsules / convex links: the Python wrapper keeps its own record of what it added
			if (nb < max_boxes) {
				double *o = boxes + 11 * nb;
				o[0] = (double) l, o[1] = g.size.x, o[2] = g.size.y, o[3] = g.size.z;
				o[4] = g.tf.t.x, o[5] = g.tf.t.y, o[6] = g.tf.t.z;
				o[7] = g.tf.q.x, o[8] = g.tf.q.y, o[9] = g.tf.q.z, o[10] = g.tf.q.w;
			}
			++nb;
		}
	for (size_t i = 0; i < r.joints.size() && (int) i < max_joints; ++i) {
		const Joint &j = r.joints[i];
		double *o = joints + 22 * i;
		o[0] = j.link_a, o[1] = j.link_b, o[2] = j.kind;
		const Tf *tfs[2] = {&j.attach_a, &j.attach_b};
		for (int k = 0; k < 2; ++k) {
			double *t = o + 3 + 7 * k;
			t[0] = tfs[k]->t.x, t[1] = tfs[k]->t.y, t[2] = tfs[k]->t.z;
			t[3] = tfs[k]->q.x, t[4] = tfs[k]->q.y, t[5] = tfs[k]->q.z, t[6] = tfs[k]->q.w;
		}
		o[17] = j.axis.x, o[18] = j.axis.y, o[19] = j.axis.z, o[20] = j.min_angle, o[21] = j.max_angle;
	}
	return nb;
}

// ---- math / FK / state helpers -----------------------------------------------------------------------------
void orc_fk(void *r, const double *state, int js, int root, double *out_link_tfs7) {
	State s = load_state(state, js);
	auto fk = forward_kinematics(*(Robot *) r, s.joints.data(), root, s.base);
	for (size_t i = 0; i < fk.size(); ++i) {
		double *o = out_link_tfs7 + 7 * i;
		o[0] = fk[i].t.x, o[1] = fk[i].t.y, o[2] = fk[i].t.z;
		o[3] = fk[i].q.x, o[4] = fk[i].q.y, o[5] = fk[i].q.z, o[6] = fk[i].q.w;
	}
}
double orc_distance(const double *a, const double *b, int js) {
	return equal_weights_distance(load_state(a, js), load_state(b, js));
}
double orc_angular_distance(const double *qa, const double *qb) {
	return angular_distance({qa[0], qa[1], qa[2], qa[3]}, {qb[0], qb[1], qb[2], qb[3]});
}
void orc_interpolate(const double *a, const double *b, int js, double t, double *out) {
	store_state(interpolate(load_state(a, js), load_state(b, js), t), out);
}
void orc_slerp(const double *qa, const double *qb, double t, double *out) {
	Quat q = slerp({qa[0], qa[1], qa[2], qa[3]}, {qb[0], qb[1], qb[2], qb[3]}, t);
	out[0] = q.x, out[1] = q.y, out[2] = q.z, out[3] = q.w;
}
void orc_tf_then(const double *a7, const double *b7, double *out7) {
	Tf r = then(load_tf(a7), load_tf(b7));
	out7[0] = r.t.x, out7[1] = r.t.y, out7[2] = r.t.z, out7[3] = r.q.x, out7[4] = r.q.y, out7[5] = r.q.z, out7[6] = r.q.w;
}
void orc_tf_inverse(const double *a7, double *out7) {
	Tf r = inverse(load_tf(a7));
	out7[0] = r.t.x, out7[1] = r.t.y, out7[2] = r.t.z, out7[3] = r.q.x, out7[4] = r.q.y, out7[5] = r.q.z, out7[6] = r.q.w;
}
uint64_t orc_motion_steps(double d, double max_step) { return motion_steps(d, max_step); }

// ---- narrow phase ------------------------------------------------------------------------------------------
/// Box given by world pose tf7 and full extents size3; triangle by 9 world coordinates.
int orc_box_tri(const double *tf7, const double *size3, const double *tri9, double *clearance_or_null) {
	Obb b(load_tf(tf7), {size3[0], size3[1], size3[2]});
	V3 p0 = b.to_local({tri9[0], tri9[1], tri9[2]}), p1 = b.to_local({tri9[3], tri9[4], tri9[5]}),
	   p2 = b.to_local({tri9[6], tri9[7], tri9[8]});
	if (clearance_or_null) *clearance_or_null = box_tri_clearance(b.h, p0, p1, p2);
	return box_tri_sat(b.h, p0, p1, p2).hit;
}

/// n box / triangle pairs at once: the exact answer (hit, signed clearance) and the libccd-style MPR second opinion
/// (mpr_second_opinion.hpp: contact reported, refinement rounds, stage at which it decided).  Any output may be null.
void orc_box_tri_batch(size_t n, const double *tf7, const double *size3, const double *tri9, double tolerance, uint8_t *hit, double *clearance,
					   uint8_t *mpr_hit, int32_t *mpr_iterations, int32_t *mpr_stage) {
	for (size_t i = 0; i < n; ++i) {
		const Tf tf = load_tf(tf7 + 7 * i);
		const V3 size{size3[3 * i], size3[3 * i + 1], size3[3 * i + 2]};
		const double *t = tri9 + 9 * i;
		const V3 a{t[0], t[1], t[2]}, b{t[3], t[4], t[5]}, c{t[6], t[7], t[8]};
		Obb o(tf, size);
		const V3 p0 = o.to_local(a), p1 = o.to_local(b), p2 = o.to_local(c);
		if (hit) hit[i] = box_tri_sat(o.h, p0, p1, p2).hit;
		if (clearance) clearance[i] = box_tri_clearance(o.h, p0, p1, p2);
		if (mpr_hit || mpr_iterations || mpr_stage) {
			const mpr::Outcome r = mpr::box_triangle(tf, size, a, b, c, tolerance);
			if (mpr_hit) mpr_hit[i] = (uint8_t) r.intersect;
			if (mpr_iterations) mpr_iterations[i] = r.iterations;
			if (mpr_stage) mpr_stage[i] = r.stage;
		}
	}
}

// ---- scene -------------------------------------------------------------------------------------------------
void *orc_scene_create(const double *verts, size_t nv, const uint32_t *tris, size_t nt) {
	try {
		return new Scene(verts, nv, tris, nt);
	} catch (...) { return nullptr; }
}
void orc_scene_free(void *s) { delete (Scene *) s; }
void orc_scene_bounds(void *s, double *lo3hi3) {
	Scene &sc = *(Scene *) s;
	lo3hi3[0] = sc.bounds.lo.x, lo3hi3[1] = sc.bounds.lo.y, lo3hi3[2] = sc.bounds.lo.z;
	lo3hi3[3] = sc.bounds.hi.x, lo3hi3[4] = sc.bounds.hi.y, lo3hi3[5] = sc.bounds.hi.z;
}

// ---- the check_* loops, batched for convenience (each element is an independent reference-style call) ------
/// hits: one byte per box; boxes: n × (tf7 + size3) = 10 doubles.
void orc_check_obbs(void *scene, const double *boxes, size_t n, uint8_t *hits, double *clearance_or_null, double reach) {
	Scene &sc = *(Scene *) scene;
	for (size_t i = 0; i < n; ++i) {
		const double *b = boxes + 10 * i;
		Obb o(load_tf(b), {b[7], b[8], b[9]});
		hits[i] = obb_hits_scene(sc, o);
		if (clearance_or_null) clearance_or_null[i] = obb_scene_clearance(sc, o, reach);
	}
}
/// states: n rows of `stride` doubles (tx ty tz qx qy qz qw j0..j{js-1}).  counters3 (optional): states, nodes, leaf tests.
void orc_check_states(void *scene, void *robot, const double *states, size_t n, size_t stride, int js, uint8_t *hits,
					  double *clearance_or_null, double reach, int threads, uint64_t *counters3_or_null) {
	Scene &sc = *(Scene *) scene;
	Robot &rb = *(Robot *) robot;
	std::atomic<uint64_t> c0{0}, c1{0}, c2{0};
	parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
		Counters cnt;
		for (size_t i = lo; i < hi; ++i) {
			State s = load_state(states + i * stride, js);
			hits[i] = check_robot_collision(rb, sc, s, &cnt);
			if (clearance_or_null) clearance_or_null[i] = robot_clearance(rb, sc, s, reach);
		}
		c0 += cnt.states_checked, c1 += cnt.nodes_visited, c2 += cnt.leaf_tests;
	});
	if (counters3_or_null) counters3_or_null[0] = c0, counters3_or_null[1] = c1, counters3_or_null[2] = c2;
}
/// Motions a[i] → b[i].  Outputs (each optional except hits): toi (first colliding t), n_steps, states_checked,
/// min_clearance over the states the reference loop visits (for band accounting, SURVEY §8d "parity report").
void orc_check_motions(void *scene, void *robot, const double *a, const double *b, size_t m, size_t stride, int js,
					   double max_step, uint8_t *hits, double *toi, uint64_t *n_steps, uint64_t *states_checked,
					   double *min_abs_clearance, double reach, int threads) {
	Scene &sc = *(Scene *) scene;
	Robot &rb = *(Robot *) robot;
	parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
		for (size_t i = lo; i < hi; ++i) {
			State sa = load_state(a + i * stride, js), sb = load_state(b + i * stride, js);
			MotionResult r = check_motion_collides(rb, sc, sa, sb, max_step);
			hits[i] = r.hit;
			if (toi) toi[i] = r.hit ? r.toi : std::numeric_limits<double>::quiet_NaN();
			if (n_steps) n_steps[i] = r.n_steps;
			if (states_checked) states_checked[i] = r.states_checked;
			if (min_abs_clearance) {
				double best = reach;
				for (size_t k = 0; k < r.states_checked; ++k) {
					double t = r.n_steps ? (double) k / (double) r.n_steps : 0.0;
					State s = r.n_steps ? interpolate(sa, sb, t) : sa;
					best = std::min(best, std::fabs(robot_clearance(rb, sc, s, reach)));
				}
				min_abs_clearance[i] = best;
			}
		}
	});
}
/// check_path_collides over one path of w states: returns first colliding segment index or -1.
long orc_check_path(void *scene, void *robot, const double *states, size_t w, size_t stride, int js, double max_step) {
	std::vector<State> path;
	for (size_t i = 0; i < w; ++i) path.push_back(load_state(states + i * stride, js));
	return check_path_collides(*(Robot *) robot, *(Scene *) scene, path, max_step);
}

// ---- samplers ----------------------------------------------------------------------------------------------
/// n states from generateUniformRandomState with one std::mt19937(seed) stream; row = 7 + n_joints doubles.
void orc_gen_uniform_states(void *robot, unsigned seed, double h_range, double v_range, double joint_range, size_t n,
							double *out) {
	Robot &rb = *(Robot *) robot;
	Rng rng(seed);
	size_t row = 7 + rb.joints.size();
	for (size_t i = 0; i < n; ++i) store_state(generate_uniform_random_state(rb, rng, h_range, v_range, joint_range), out + i * row);
}
/// The raw draws genUprightState consumes (goal_sampling.cpp:13-49): per sample [yaw, θ_0..θ_{V-1}].
void orc_goal_draws(void *robot, unsigned seed, size_t n, double *out) {
	Robot &rb = *(Robot *) robot;
	Rng rng(seed);
	size_t row = 1 + rb.n_variables();
	for (size_t i = 0; i < n; ++i) {
		out[i * row] = rng.uniform_real(-M_PI, M_PI);
		size_t k = 1;
		for (auto &j: rb.joints)
			if (j.kind == REVOLUTE) out[i * row + k++] = rng.uniform_real(j.min_angle, j.max_angle);
	}
}
/// n goal states for one target from genGoalStateUniform with a fresh mt19937(seed); row = 7 + n_variables.
void orc_gen_goal_states(void *robot, unsigned seed, const double *target3, double distance_from_target, size_t n,
						 double *out) {
	Robot &rb = *(Robot *) robot;
	Rng rng(seed);
	int fb = rb.find_link("flying_base"), ee = rb.find_link("end_effector");
	size_t row = 7 + rb.n_variables();
	for (size_t i = 0; i < n; ++i)
		store_state(gen_goal_state_uniform(rng, {target3[0], target3[1], target3[2]}, distance_from_target, rb, fb, ee),
					out + i * row);
}
/// goal_sample_success_rate.cpp:121-142 for F targets: per target, S samples from a fresh mt19937(seed), count collisions.
void orc_goal_sample_counts(void *scene, void *robot, unsigned seed, const double *targets, size_t f, size_t s_per,
							uint32_t *collisions, uint8_t *hits_or_null, int threads) {
	Scene &sc = *(Scene *) scene;
	Robot &rb = *(Robot *) robot;
	int fb = rb.find_link("flying_base"), ee = rb.find_link("end_effector");
	parallel_for(f, threads, [&](size_t lo, size_t hi, int) {
		for (size_t i = lo; i < hi; ++i) {
			Rng rng(seed);
			uint32_t c = 0;
			for (size_t k = 0; k < s_per; ++k) {
				State st = gen_goal_state_uniform(rng, {targets[3 * i], targets[3 * i + 1], targets[3 * i + 2]}, 0.0, rb, fb, ee);
				bool h = check_robot_collision(rb, sc, st);
				c += h;
				if (hits_or_null) hits_or_null[i * s_per + k] = h;
			}
			collisions[i] = c;
		}
	});
}
/// k nearest neighbours under equal_weights_distance by exhaustive search — what NearestNeighborsGNAT::nearestK returns
/// (src/planning/nearest_neighbours/NearestNeighborsGNAT.h:180-189 with the metric set at tsp_over_prm.cpp:383-386).
/// causal: state i only sees states j < i (the roadmap when node i was inserted, tsp_over_prm.cpp:33-66).
void orc_knn(const double *states, size_t n, size_t stride, int js, int k, int causal, int32_t *idx, double *dist) {
	std::vector<State> st;
	for (size_t i = 0; i < n; ++i) st.push_back(load_state(states + i * stride, js));
	for (size_t i = 0; i < n; ++i) {
		std::vector<std::pair<double, int32_t>> c;
		for (size_t j = 0; j < (causal ? i : n); ++j)
			if (j != i) c.push_back({equal_weights_distance(st[i], st[j]), (int32_t) j});
		std::stable_sort(c.begin(), c.end(), [](auto &a, auto &b) { return a.first < b.first; });
		for (int q = 0; q < k; ++q) {
			idx[i * k + q] = q < (int) c.size() ? c[q].second : -1;
			dist[i * k + q] = q < (int) c.size() ? c[q].first : std::numeric_limits<double>::quiet_NaN();
		}
	}
}
void orc_from_ee_and_vector(void *robot, const double *point3, const double *vec3, double *out8) {
	store_state(from_end_effector_and_vector(*(Robot *) robot, {point3[0], point3[1], point3[2]}, {vec3[0], vec3[1], vec3[2]}), out8);
}

void orc_arm_vector(void *robot, const double *state, int js, double *out3) {
	V3 v = arm_vector_from_state(*(Robot *) robot, load_state(state, js));
	out3[0] = v.x, out3[1] = v.y, out3[2] = v.z;
}
/// generateUniformRandomArmVectorState / findGoalStateByUniformSampling (goal_sampling.cpp:81-124) for n targets, each with a
/// fresh std::mt19937(seed + i) and the oracle's check_robot_collision.  out: n x 8 (found ? state : NaN), found: n flags,
/// next_draw: n x the generator's next raw output (pins how many draws the loop consumed).  which: 0 arm-vector, 1 uniform.
void orc_sample_goal_states(void *scene, void *robot, int which, unsigned seed, const double *targets, size_t n, int max_attempts,
							double ee_distance, double *out, uint8_t *found, uint32_t *next_draw) {
	auto &rb = *(Robot *) robot;
	auto &sc = *(Scene *) scene;
	const int fb = rb.find_link("flying_base"), ee = rb.find_link("end_effector");
	const size_t row = 7 + (which == 0 ? 1 : (size_t) rb.n_variables());
	StateFn collides = [&](const State &s) { return check_robot_collision(rb, sc, s); };
	for (size_t i = 0; i < n; ++i) {
		Rng rng(seed + (unsigned) i);
		const V3 t{targets[3 * i], targets[3 * i + 1], targets[3 * i + 2]};
		std::optional<State> s = which == 0 ? generate_uniform_random_arm_vector_state(rb, collides, t, rng, max_attempts, ee_distance)
											: find_goal_state_by_uniform_sampling(t, rb, fb, ee, collides, rng, (size_t) max_attempts);
		found[i] = s.has_value();
		for (size_t k = 0; k < row; ++k) out[i * row + k] = std::numeric_limits<double>::quiet_NaN();
		if (s) store_state(*s, out + i * row);
		next_draw[i] = (uint32_t) rng.uniform_integer(0, std::numeric_limits<int>::max());
	}
}

// ---- local optimisation (shortcutting) ---------------------------------------------------------------------
/// Runs one of the reference's path optimisers on a path of w states with one std::mt19937(seed) stream and the
/// oracle's check_motion_collides as the motion check.  mode 0: `iterations` x tryShortcuttingRandomlyGlobally
/// (tsp_over_prm.cpp:594-612); 1: tryDeletingEveryWaypoint; 2: `iterations` x tryShortcuttingRandomlyLocally;
/// 3: one tryMidpointPull(pull) sweep over the interior waypoints.  Writes the resulting path (out_cap rows of
/// `stride`), its length to *out_w, the number of motion checks to *n_checks; returns the number of successes.
long orc_shortcut_path(void *scene, void *robot, const double *states, size_t w, size_t stride, int js, unsigned seed,
					   size_t iterations, int mode, double pull, double max_step, double *out_states, size_t out_cap,
					   size_t *out_w, size_t *n_checks) {
	Path path;
	for (size_t i = 0; i < w; ++i) path.push_back(load_state(states + i * stride, js));
	Rng rng(seed);
	size_t checks = 0;
	MotionFn motion = [&](const State &a, const State &b) {
		++checks;
		return check_motion_collides(*(Robot *) robot, *(Scene *) scene, a, b, max_step).hit;
	};
	long ok = 0;
	if (mode == 0) ok = (long) optimize_path_segment(path, motion, rng, iterations);
	else if (mode == 1) ok = try_deleting_every_waypoint(path, motion);
	else if (mode == 2) {
		for (size_t i = 0; i < iterations && path.size() > 2; ++i) ok += try_shortcutting_randomly_locally(path, rng, motion);
	} else {
		for (size_t i = 1; i + 1 < path.size(); ++i) ok += try_midpoint_pull(path, i, pull, motion);
	}
	for (size_t i = 0; i < path.size() && i < out_cap; ++i) store_state(path[i], out_states + i * stride);
	*out_w = path.size();
	if (n_checks) *n_checks = checks;
	return ok;
}

// ---- roadmap shortest paths --------------------------------------------------------------------------------
/// Dijkstra from every source over an undirected edge list; dist / pred: n_sources x n_vertices, source-major.
void orc_roadmap_shortest_paths(uint32_t n_vertices, const uint32_t *edges, const double *weights, size_t n_edges,
								const uint32_t *sources, size_t n_sources, double *dist, uint32_t *pred, int threads) {
	Roadmap g(n_vertices);
	for (size_t e = 0; e < n_edges; ++e) g.add_edge(edges[2 * e], edges[2 * e + 1], weights[e]);
	parallel_for(n_sources, threads, [&](size_t lo, size_t hi, int) {
		std::vector<double> d;
		std::vector<uint32_t> p;
		for (size_t s = lo; s < hi; ++s) {
			run_dijkstra(g, sources[s], d, p);
			std::copy(d.begin(), d.end(), dist + s * n_vertices);
			if (pred) std::copy(p.begin(), p.end(), pred + s * n_vertices);
		}
	});
}

// ---- mesh connected components ----------------------------------------------------------------------------
void orc_mesh_components(const uint32_t *tris, size_t nt, size_t nv, uint32_t *labels) {
	auto l = mesh_component_labels(tris, nt, nv);
	std::copy(l.begin(), l.end(), labels);
}

}  // extern "C"
