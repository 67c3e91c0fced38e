// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// A CPU, IEEE-fp64 restatement of the collision-query hot path of
// werner291/Multigoal-Orchard-Drone-Planning-Library ("the reference").  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or call this.
// The product path (libmgodpl_b200.so) never includes or links anything under oracle/.
//
// PARITY STATUS: the planning-layer arithmetic (FK, interpolate, distance, n_steps, samplers, RNG draws) is
// PINNED against the reference's own sources compiled here (oracle/_ref, see oracle/Makefile and
// tests/test_oracle_vs_ref.py) and against SURVEY.md Appendix B known-answer vectors.  The narrow phase
// (box vs triangle) lives in FCL 0.7.x + libccd 2.1, which are absent from /root/reference and from this image:
// that part is "parity unpinned" — it restates the published semantics (closed box ∩ closed triangle, union over
// triangles) with an exact separating-axis test and reports a signed clearance so that queries inside the
// reference's 1e-6 m libccd tolerance band can be counted separately.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <numeric>
#include <optional>
#include <functional>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------------------
// L0 math: src/math/Vec3.h:22-24, src/math/Quaternion.h:25-85, src/math/Transform.h:20-78
// ---------------------------------------------------------------------------------------------------------
struct V3 {
	double x = 0, y = 0, z = 0;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline double comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/// Unit quaternion stored x,y,z,w (Quaternion.h:25-27).
struct Quat {
	double x = 0, y = 0, z = 0, w = 1;
};
/// Conjugate doubles as the inverse; no renormalisation anywhere (Quaternion.h:28-34).
inline Quat conj(Quat q) { return {-q.x, -q.y, -q.z, q.w}; }
/// Hamilton product, term order as Quaternion.h:36-44.
inline Quat qmul(Quat a, Quat b) {
	return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
			a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
			a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
			a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
/// Rotation by sandwich product q (v,0) q* (Quaternion.h:46-51).
inline V3 qrot(Quat q, V3 v) {
	Quat r = qmul(qmul(q, Quat{v.x, v.y, v.z, 0.0}), conj(q));
	return {r.x, r.y, r.z};
}
/// Quaternion.h:53-62.
inline Quat from_axis_angle(V3 axis, double angle) {
	double half = angle / 2.0;
	double s = std::sin(half);
	return {axis.x * s, axis.y * s, axis.z * s, std::cos(half)};
}
/// Quaternion.h:78-85: acos(2 clamp(a.b)^2 - 1); dot accumulates w first.
inline double angular_distance(Quat a, Quat b) {
	double c = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
	c = std::clamp(c, -1.0, 1.0);
	return std::acos(2.0 * c * c - 1.0);
}
/// Quaternion.cpp:12-20 hands off to Eigen::Quaterniond::slerp.  Eigen (3.4 series; pinned only through
/// flake.lock's nixpkgs 24.05) is absent; this restates its published rule: shortest arc, linear weights when
/// |dot| >= 1-eps, weights sin((1-t)θ)/sinθ and sin(tθ)/sinθ otherwise, no renormalisation.
inline Quat slerp(Quat a, Quat b, double t) {
	const double one = 1.0 - std::numeric_limits<double>::epsilon();
	double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
	double ad = std::fabs(d);
	double s0, s1;
	if (ad >= one) {
		s0 = 1.0 - t;
		s1 = t;
	} else {
		double theta = std::acos(ad);
		double st = std::sin(theta);
		s0 = std::sin((1.0 - t) * theta) / st;
		s1 = std::sin(t * theta) / st;
	}
	if (d < 0.0) s1 = -s1;
	return {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

/// Rigid transform {translation, orientation} (Transform.h:20-23).
struct Tf {
	V3 t;
	Quat q;
};
inline Tf tf_identity() { return {{0, 0, 0}, {0, 0, 0, 1}}; }
/// Transform.h:31-36.
inline Tf then(const Tf &a, const Tf &b) { return {a.t + qrot(a.q, b.t), qmul(a.q, b.q)}; }
/// Transform.h:52-58.
inline Tf inverse(const Tf &a) {
	Quat qi = conj(a.q);
	return {qrot(qi, -a.t), qi};
}
/// Transform.h:69-78: translation (1-t)a + t b, orientation slerp.
inline Tf interpolate(const Tf &a, const Tf &b, double t) {
	return {{(1.0 - t) * a.t.x + t * b.t.x, (1.0 - t) * a.t.y + t * b.t.y, (1.0 - t) * a.t.z + t * b.t.z},
			slerp(a.q, b.q, t)};
}

// ---------------------------------------------------------------------------------------------------------
// Robot model: src/planning/RobotModel.h:36-196, RobotModel.cpp:17-161
// ---------------------------------------------------------------------------------------------------------
enum JointKind : int { REVOLUTE = 0, FIXED = 1 };

enum GeomKind : int { GEOM_BOX = 0, GEOM_CAPSULE = 1, GEOM_CONVEX = 2 };
struct BoxGeom {  // src/experiment_utils/shapes.h:20-22 + positioned_shape.h:12-30
	V3 size{0, 0, 0};      // full extents, centred
	Tf tf = tf_identity(); // relative to the link frame
	// Not in the reference (it collides boxes only, collision_detection.cpp:21-22,49-51): capsule and convex-polytope links of
	// the CUDA library (BASELINE north_star (3)), defined here from first principles; parity for them is against this file only.
	int kind = GEOM_BOX;
	double radius = 0, length = 0;  // capsule: points within `radius` of the segment (0, 0, -+length / 2) of the shape frame
	std::vector<V3> verts;          // convex: the hull of these points (shape frame)
};
struct Link {
	std::string name;
	std::vector<int> joints;  // back-links in insertion order (RobotModel.cpp:103-112)
	std::vector<BoxGeom> boxes;
};
struct Joint {
	Tf attach_a, attach_b;
	int link_a = 0, link_b = 0;
	int kind = FIXED;
	V3 axis{1, 0, 0};
	double min_angle = 0, max_angle = 0;
};
struct Robot {
	std::vector<Link> links;
	std::vector<Joint> joints;
	int add_link(const Link &l) {
		links.push_back(l);
		return (int) links.size() - 1;
	}
	int add_joint(const Joint &j) {
		joints.push_back(j);
		links[j.link_a].joints.push_back((int) joints.size() - 1);
		links[j.link_b].joints.push_back((int) joints.size() - 1);
		return (int) joints.size() - 1;
	}
	int find_link(const std::string &name) const {  // RobotModel.cpp:114-120
		for (size_t i = 0; i < links.size(); ++i)
			if (links[i].name == name) return (int) i;
		throw std::runtime_error("Link not found: " + name);
	}
	int n_variables() const {  // RobotModel.cpp:130-134
		int n = 0;
		for (auto &j: joints) n += j.kind == REVOLUTE ? 1 : 0;
		return n;
	}
};

/// Explicit-stack DFS of RobotModel.cpp:17-90.  Quirks kept: the skip test looks only at visited[linkB]
/// (:56-57); the variable cursor advances when a joint is *expanded*, not when its link is popped (:64-69);
/// a link pushed twice keeps the transform of its first pop (:43-50).
inline std::vector<Tf> forward_kinematics(const Robot &r, const double *joint_values, int root, const Tf &root_tf) {
	std::vector<Tf> out(r.links.size(), root_tf);
	std::vector<char> visited(r.links.size(), 0);
	std::vector<std::pair<int, Tf>> stack{{root, root_tf}};
	size_t var = 0;
	while (!stack.empty()) {
		auto [link, tf] = stack.back();
		stack.pop_back();
		if (visited[link]) continue;
		visited[link] = 1;
		out[link] = tf;
		for (int ji: r.links[link].joints) {
			const Joint &j = r.joints[ji];
			if (visited[j.link_b]) continue;
			bool from_a = j.link_a == link;
			Tf to_joint = from_a ? j.attach_a : j.attach_b;
			Tf var_tf = j.kind == REVOLUTE ? Tf{{0, 0, 0}, from_axis_angle(j.axis, joint_values[var])} : tf_identity();
			if (j.link_b == link) var_tf = inverse(var_tf);
			var += j.kind == REVOLUTE ? 1 : 0;
			Tf after = inverse(from_a ? j.attach_b : j.attach_a);
			Tf next = then(then(then(tf, to_joint), var_tf), after);
			stack.emplace_back(from_a ? j.link_b : j.link_a, next);
		}
	}
	return out;
}

/// createProceduralRobotModel, src/experiment_utils/procedural_robot_models.cpp:86-154.
/// joint_types: 0 = HORIZONTAL (axis x), 1 = VERTICAL (axis z).
inline Robot procedural_robot(double total_arm_length, const std::vector<int> &joint_types) {
	Robot m;
	double stick = total_arm_length / (double) joint_types.size();
	BoxGeom base_box;
	base_box.size = {0.8, 0.8, 0.2};
	int base = m.add_link({"flying_base", {}, {base_box}});
	int attach_to = base;
	Tf attach_a{{0, 0.2, 0}, {0, 0, 0, 1}};
	for (int jt: joint_types) {
		Tf stick_tf{{0, -stick / 2, 0}, {0, 0, 0, 1}};
		BoxGeom stick_box;
		stick_box.size = {0.05, stick, 0.05};
		int link = m.add_link({"stick", {}, {stick_box}});
		Joint j;
		j.attach_a = attach_a;
		j.attach_b = stick_tf;
		j.link_a = attach_to;
		j.link_b = link;
		j.kind = REVOLUTE;
		j.axis = jt == 0 ? V3{1, 0, 0} : V3{0, 0, 1};
		j.min_angle = -M_PI / 2.0;
		j.max_angle = M_PI / 2.0;
		m.add_joint(j);
		attach_a = inverse(stick_tf);
		attach_to = link;
	}
	int ee = m.add_link({"end_effector", {}, {}});
	Joint f;
	f.attach_a = attach_a;
	f.attach_b = tf_identity();
	f.link_a = attach_to;
	f.link_b = ee;
	f.kind = FIXED;
	m.add_joint(f);
	return m;
}

// ---------------------------------------------------------------------------------------------------------
// Robot state: src/planning/RobotState.h:16-23, RobotState.cpp:14-24, distance.cpp:17-25
// ---------------------------------------------------------------------------------------------------------
struct State {
	Tf base;
	std::vector<double> joints;
};
/// RobotState.cpp:14-24: joints interpolate as a(1-t) + b t (operand order differs from the translation lerp).
inline State interpolate(const State &a, const State &b, double t) {
	State s;
	s.base = interpolate(a.base, b.base, t);
	s.joints.reserve(a.joints.size());
	for (size_t i = 0; i < a.joints.size(); ++i) s.joints.push_back(a.joints[i] * (1 - t) + b.joints[i] * t);
	return s;
}
/// distance.cpp:17-25: |Δt| + angular_distance + Σ|Δθ_i| over *all* stored joint values.
inline double equal_weights_distance(const State &a, const State &b) {
	double d = norm(a.base.t - b.base.t);
	d += angular_distance(a.base.q, b.base.q);
	for (size_t i = 0; i < a.joints.size(); ++i) d += std::abs(a.joints[i] - b.joints[i]);
	return d;
}

// ---------------------------------------------------------------------------------------------------------
// Narrow phase (stands where FCL + libccd stood; semantics per SURVEY §8c "follow semantics, not source"):
// closed oriented box vs closed triangle.  Exact 13-axis separating-axis test in the box frame, plus the
// exact Euclidean distance for the separated case and the SAT penetration depth for the overlapping case,
// so that callers can classify a query against the 1e-6 m band.
// ---------------------------------------------------------------------------------------------------------
struct SatResult {
	bool hit;         // closed sets intersect (touching counts)
	double sat_gap;   // max over axes of normalised separation: > 0 separated (a lower bound on the distance),
	                  // <= 0 overlapping (then -sat_gap is the penetration depth)
};

/// p0,p1,p2: triangle vertices already expressed in the box frame; h: box half extents.
inline SatResult box_tri_sat(const V3 &h, const V3 &p0, const V3 &p1, const V3 &p2) {
	const V3 p[3] = {p0, p1, p2};
	const V3 f[3] = {p1 - p0, p2 - p1, p0 - p2};
	double gap = -std::numeric_limits<double>::infinity();
	double scale2 = std::max({dot(f[0], f[0]), dot(f[1], f[1]), dot(f[2], f[2])});
	auto axis = [&](const V3 &a, double len2_floor) {
		double l2 = dot(a, a);
		if (!(l2 > len2_floor)) return;  // degenerate axis carries no information
		double q0 = dot(a, p[0]), q1 = dot(a, p[1]), q2 = dot(a, p[2]);
		double lo = std::min({q0, q1, q2}), hi = std::max({q0, q1, q2});
		double r = h.x * std::fabs(a.x) + h.y * std::fabs(a.y) + h.z * std::fabs(a.z);
		double s = std::max(lo - r, -hi - r) / std::sqrt(l2);
		gap = std::max(gap, s);
	};
	axis({1, 0, 0}, 0.0);
	axis({0, 1, 0}, 0.0);
	axis({0, 0, 1}, 0.0);
	// Triangle normal; degenerate when the triangle has (numerically) no area.
	axis(cross(f[0], f[1]), 1e-28 * scale2 * scale2);
	// 9 edge-edge axes e_j x f_k; degenerate when the triangle edge is parallel to the box axis.
	for (int k = 0; k < 3; ++k) {
		double fl2 = dot(f[k], f[k]);
		axis({0, -f[k].z, f[k].y}, 1e-24 * fl2);
		axis({f[k].z, 0, -f[k].x}, 1e-24 * fl2);
		axis({-f[k].y, f[k].x, 0}, 1e-24 * fl2);
	}
	return {gap <= 0.0, gap};
}

inline V3 clamp_to_box(const V3 &h, const V3 &p) {
	return {std::clamp(p.x, -h.x, h.x), std::clamp(p.y, -h.y, h.y), std::clamp(p.z, -h.z, h.z)};
}
/// Closest point on triangle (a,b,c) to p — region classification by barycentric signs.
inline V3 closest_on_triangle(const V3 &p, const V3 &a, const V3 &b, const V3 &c) {
	V3 ab = b - a, ac = c - a, ap = p - a;
	double d1 = dot(ab, ap), d2 = dot(ac, ap);
	if (d1 <= 0 && d2 <= 0) return a;
	V3 bp = p - b;
	double d3 = dot(ab, bp), d4 = dot(ac, bp);
	if (d3 >= 0 && d4 <= d3) return b;
	double vc = d1 * d4 - d3 * d2;
	if (vc <= 0 && d1 >= 0 && d3 <= 0) return a + ab * (d1 / (d1 - d3));
	V3 cp = p - c;
	double d5 = dot(ab, cp), d6 = dot(ac, cp);
	if (d6 >= 0 && d5 <= d6) return c;
	double vb = d5 * d2 - d1 * d6;
	if (vb <= 0 && d2 >= 0 && d6 <= 0) return a + ac * (d2 / (d2 - d6));
	double va = d3 * d6 - d5 * d4;
	if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
	double denom = va + vb + vc;
	if (!(std::fabs(denom) > 0)) return a;
	return a + ab * (vb / denom) + ac * (vc / denom);
}
/// Squared distance between segments [p1,q1] and [p2,q2].
inline double seg_seg_dist2(const V3 &p1, const V3 &q1, const V3 &p2, const V3 &q2) {
	V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
	double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
	double s, t;
	const double tiny = 1e-300;
	if (a <= tiny && e <= tiny) return dot(r, r);
	if (a <= tiny) {
		s = 0;
		t = std::clamp(f / e, 0.0, 1.0);
	} else {
		double c = dot(d1, r);
		if (e <= tiny) {
			t = 0;
			s = std::clamp(-c / a, 0.0, 1.0);
		} else {
			double b = dot(d1, d2), den = a * e - b * b;
			s = den > 1e-14 * a * e ? std::clamp((b * f - c * e) / den, 0.0, 1.0) : 0.0;
			t = (b * s + f) / e;
			if (t < 0) {
				t = 0;
				s = std::clamp(-c / a, 0.0, 1.0);
			} else if (t > 1) {
				t = 1;
				s = std::clamp((b - c) / a, 0.0, 1.0);
			}
		}
	}
	V3 c1 = p1 + d1 * s, c2 = p2 + d2 * t;
	V3 dd = c1 - c2;
	return dot(dd, dd);
}
/// Exact Euclidean distance between a centred box (half extents h) and a triangle that does NOT intersect it:
/// min over (triangle vertex → box), (box corner → triangle), (box edge ↔ triangle edge).
inline double box_tri_distance_separated(const V3 &h, const V3 &p0, const V3 &p1, const V3 &p2) {
	const V3 p[3] = {p0, p1, p2};
	double best = std::numeric_limits<double>::infinity();
	for (auto &v: p) {
		V3 d = v - clamp_to_box(h, v);
		best = std::min(best, dot(d, d));
	}
	V3 corner[8];
	for (int i = 0; i < 8; ++i) corner[i] = {(i & 1 ? h.x : -h.x), (i & 2 ? h.y : -h.y), (i & 4 ? h.z : -h.z)};
	for (auto &c: corner) {
		V3 d = c - closest_on_triangle(c, p0, p1, p2);
		best = std::min(best, dot(d, d));
	}
	for (int i = 0; i < 8; ++i)
		for (int bit = 1; bit < 8; bit <<= 1)
			if (!(i & bit))
				for (int k = 0; k < 3; ++k)
					best = std::min(best, seg_seg_dist2(corner[i], corner[i | bit], p[k], p[(k + 1) % 3]));
	return std::sqrt(best);
}
/// Signed clearance of one box/triangle pair: +distance when separated, -penetration depth when overlapping.
inline double box_tri_clearance(const V3 &h, const V3 &p0, const V3 &p1, const V3 &p2) {
	SatResult s = box_tri_sat(h, p0, p1, p2);
	if (s.hit) return s.sat_gap;  // <= 0
	return box_tri_distance_separated(h, p0, p1, p2);
}

/// Capsule vs triangle in the capsule's frame: signed clearance = (distance from the axis segment to the triangle) - radius.
/// The closest approach of a segment and a triangle is 0 if the segment pierces the triangle, otherwise it is attained at a
/// segment end point or on a triangle edge.
inline double capsule_tri_clearance(double radius, double length, const V3 &p0, const V3 &p1, const V3 &p2) {
	const V3 a{0, 0, -0.5 * length}, b{0, 0, 0.5 * length};
	const V3 tri[3] = {p0, p1, p2};
	double best = std::numeric_limits<double>::infinity();
	for (const V3 &e: {a, b}) {
		V3 d = e - closest_on_triangle(e, p0, p1, p2);
		best = std::min(best, dot(d, d));
	}
	for (int k = 0; k < 3; ++k) best = std::min(best, seg_seg_dist2(a, b, tri[k], tri[(k + 1) % 3]));
	// piercing: solve a + t (b - a) = p0 + u (p1 - p0) + v (p2 - p0) (Moeller-Trumbore form)
	V3 dir = b - a, e1 = p1 - p0, e2 = p2 - p0, pv = cross(dir, e2);
	double det = dot(e1, pv);
	if (std::fabs(det) > 0) {
		V3 tv = a - p0;
		double u = dot(tv, pv) / det;
		V3 qv = cross(tv, e1);
		double v = dot(dir, qv) / det, t = dot(e2, qv) / det;
		if (u >= 0 && v >= 0 && u + v <= 1 && t >= 0 && t <= 1) best = 0;
	}
	return std::sqrt(best) - radius;
}

/// Convex polytope (hull of `verts`, shape frame) vs triangle (same frame) by separating axes, with a deliberately generous
/// axis set found by exhaustive search at every call: the normals of ALL supporting planes through three vertices, the triangle
/// normal, and (v_i - v_j) x (triangle edge) for ALL vertex pairs — a superset of the hull's real edges, so nothing depends on
/// a hull construction being right.  sat_gap: largest normalised separation (> 0 separated, a lower bound of the distance;
/// <= 0 overlapping, minus the penetration depth along the best axis).
inline SatResult convex_tri_sat(const std::vector<V3> &verts, const V3 &p0, const V3 &p1, const V3 &p2) {
	const V3 p[3] = {p0, p1, p2};
	const V3 f[3] = {p1 - p0, p2 - p1, p0 - p2};
	double gap = -std::numeric_limits<double>::infinity();
	double scale = 0;
	for (auto &v: verts) scale = std::max({scale, std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)});
	auto axis = [&](const V3 &a, double len2_floor) {
		double l2 = dot(a, a);
		if (!(l2 > len2_floor)) return;
		double lo = 1e300, hi = -1e300;
		for (auto &v: verts) {
			double q = dot(a, v);
			lo = std::min(lo, q), hi = std::max(hi, q);
		}
		double q0 = dot(a, p[0]), q1 = dot(a, p[1]), q2 = dot(a, p[2]);
		double tlo = std::min({q0, q1, q2}), thi = std::max({q0, q1, q2});
		gap = std::max(gap, std::max(tlo - hi, lo - thi) / std::sqrt(l2));
	};
	const size_t n = verts.size();
	for (size_t i = 0; i < n; ++i)
		for (size_t j = i + 1; j < n; ++j)
			for (size_t k = j + 1; k < n; ++k) {
				V3 nn = cross(verts[j] - verts[i], verts[k] - verts[i]);
				if (!(dot(nn, nn) > 1e-24 * scale * scale * scale * scale)) continue;
				double lo = 0, hi = 0, len = std::sqrt(dot(nn, nn));
				for (auto &v: verts) {
					double sd = dot(nn, v - verts[i]) / len;
					lo = std::min(lo, sd), hi = std::max(hi, sd);
				}
				if (hi > 1e-10 * scale && lo < -1e-10 * scale) continue;  // not a supporting plane
				axis(nn, 0.0);
			}
	double scale2 = std::max({dot(f[0], f[0]), dot(f[1], f[1]), dot(f[2], f[2])});
	axis(cross(f[0], f[1]), 1e-28 * scale2 * scale2);
	for (size_t i = 0; i < n; ++i)
		for (size_t j = i + 1; j < n; ++j) {
			V3 d = verts[i] - verts[j];
			double dl2 = dot(d, d);
			if (!(dl2 > 0)) continue;
			for (int k = 0; k < 3; ++k) axis(cross(d, f[k]), 1e-24 * dl2 * dot(f[k], f[k]));
		}
	return {gap <= 0.0, gap};
}

// ---------------------------------------------------------------------------------------------------------
// Scene (stands where fcl::BVHModel<OBBd> stood): the triangle soup of Mesh (src/planning/Mesh.h:19-22) as
// handed to meshToFclBVH (src/planning/fcl_utils.cpp:23-56), identity pose, surface only.
// Broad phase: median-split AABB tree, 4 triangles per leaf — any correct BVH does.
// ---------------------------------------------------------------------------------------------------------
struct Aabb {
	V3 lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
	void grow(const V3 &p) {
		lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
		hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
	}
	void grow(const Aabb &o) {
		grow(o.lo);
		grow(o.hi);
	}
};

struct Scene {
	std::vector<std::array<V3, 3>> tris;
	struct Node {
		Aabb box;
		int left = -1, right = -1;  // children, or (first, count) into order[] when leaf
		bool leaf = false;
	};
	std::vector<Node> nodes;
	std::vector<int> order;
	Aabb bounds;

	Scene(const double *verts, size_t nv, const uint32_t *idx, size_t nt) {
		tris.resize(nt);
		for (size_t i = 0; i < nt; ++i)
			for (int k = 0; k < 3; ++k) {
				size_t v = idx[3 * i + k];
				if (v >= nv) throw std::runtime_error("triangle index out of range");
				tris[i][k] = {verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]};
			}
		order.resize(nt);
		std::iota(order.begin(), order.end(), 0);
		std::vector<Aabb> tb(nt);
		std::vector<V3> ctr(nt);
		for (size_t i = 0; i < nt; ++i) {
			for (auto &v: tris[i]) tb[i].grow(v);
			ctr[i] = (tb[i].lo + tb[i].hi) * 0.5;
			bounds.grow(tb[i]);
		}
		nodes.reserve(nt / 2 + 16);
		if (nt) build(0, (int) nt, tb, ctr);
	}

	int build(int first, int last, const std::vector<Aabb> &tb, const std::vector<V3> &ctr) {
		int id = (int) nodes.size();
		nodes.emplace_back();
		Aabb box, cb;
		for (int i = first; i < last; ++i) {
			box.grow(tb[order[i]]);
			cb.grow(ctr[order[i]]);
		}
		nodes[id].box = box;
		if (last - first <= 4) {
			nodes[id].leaf = true;
			nodes[id].left = first;
			nodes[id].right = last - first;
			return id;
		}
		V3 e = cb.hi - cb.lo;
		int ax = e.x >= e.y ? (e.x >= e.z ? 0 : 2) : (e.y >= e.z ? 1 : 2);
		int mid = (first + last) / 2;
		std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + last,
						 [&](int a, int b) { return comp(ctr[a], ax) < comp(ctr[b], ax); });
		int l = build(first, mid, tb, ctr);
		int r = build(mid, last, tb, ctr);
		nodes[id].left = l;
		nodes[id].right = r;
		return id;
	}
};

/// An oriented box in the world: centre c, rotation columns from q, half extents h.
struct Obb {
	V3 c;
	Quat q;
	V3 h;
	double R[3][3];  // world = R * local
	Obb(const Tf &tf, const V3 &size) : c(tf.t), q(tf.q), h{size.x * 0.5, size.y * 0.5, size.z * 0.5} {
		// standard unit-quaternion rotation matrix (what Eigen's Transform::rotate builds).
		double x = q.x, y = q.y, z = q.z, w = q.w;
		R[0][0] = 1 - 2 * (y * y + z * z);
		R[0][1] = 2 * (x * y - z * w);
		R[0][2] = 2 * (x * z + y * w);
		R[1][0] = 2 * (x * y + z * w);
		R[1][1] = 1 - 2 * (x * x + z * z);
		R[1][2] = 2 * (y * z - x * w);
		R[2][0] = 2 * (x * z - y * w);
		R[2][1] = 2 * (y * z + x * w);
		R[2][2] = 1 - 2 * (x * x + y * y);
	}
	V3 to_local(const V3 &p) const {
		V3 d = p - c;
		return {R[0][0] * d.x + R[1][0] * d.y + R[2][0] * d.z, R[0][1] * d.x + R[1][1] * d.y + R[2][1] * d.z,
				R[0][2] * d.x + R[1][2] * d.y + R[2][2] * d.z};
	}
	/// Conservative overlap with an AABB grown by `pad`: 3 world axes + 3 box axes.
	bool may_touch(const Aabb &b, double pad) const {
		V3 bc = (b.lo + b.hi) * 0.5, bh = (b.hi - b.lo) * 0.5 + V3{pad, pad, pad};
		V3 d = bc - c;
		const double dd[3] = {d.x, d.y, d.z}, hh[3] = {h.x, h.y, h.z}, bhh[3] = {bh.x, bh.y, bh.z};
		for (int i = 0; i < 3; ++i) {
			double r = std::fabs(R[i][0]) * hh[0] + std::fabs(R[i][1]) * hh[1] + std::fabs(R[i][2]) * hh[2];
			if (std::fabs(dd[i]) > r + bhh[i]) return false;
		}
		for (int j = 0; j < 3; ++j) {
			double proj = R[0][j] * dd[0] + R[1][j] * dd[1] + R[2][j] * dd[2];
			double r = std::fabs(R[0][j]) * bhh[0] + std::fabs(R[1][j]) * bhh[1] + std::fabs(R[2][j]) * bhh[2];
			if (std::fabs(proj) > r + hh[j]) return false;
		}
		return true;
	}
};

struct Counters {  // the reference's own observability hook is an invocation counter (functional_utils.cppm:27-32)
	uint64_t states_checked = 0, nodes_visited = 0, leaf_tests = 0;
};

/// Boolean box-vs-scene query with first-hit exit (what fcl::collide + isCollision answer at
/// collision_detection.cpp:36-45).
inline bool obb_hits_scene(const Scene &s, const Obb &b, Counters *cnt = nullptr) {
	if (s.nodes.empty()) return false;
	int stack[128];
	int sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const Scene::Node &n = s.nodes[stack[--sp]];
		if (cnt) cnt->nodes_visited++;
		if (!b.may_touch(n.box, 0.0)) continue;
		if (n.leaf) {
			for (int i = 0; i < n.right; ++i) {
				const auto &t = s.tris[s.order[n.left + i]];
				if (cnt) cnt->leaf_tests++;
				if (box_tri_sat(b.h, b.to_local(t[0]), b.to_local(t[1]), b.to_local(t[2])).hit) return true;
			}
		} else {
			stack[sp++] = n.right;
			stack[sp++] = n.left;
		}
	}
	return false;
}

/// Signed clearance of a box against the whole scene, exact wherever |clearance| <= `reach`:
///   any triangle overlapping  → -(deepest penetration depth)
///   none overlapping          → +(smallest distance), or +reach if nothing lies within reach.
inline double obb_scene_clearance(const Scene &s, const Obb &b, double reach) {
	double min_dist = reach, max_pen = -1.0;
	if (s.nodes.empty()) return reach;
	int stack[128];
	int sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const Scene::Node &n = s.nodes[stack[--sp]];
		if (!b.may_touch(n.box, reach)) continue;
		if (n.leaf) {
			for (int i = 0; i < n.right; ++i) {
				const auto &t = s.tris[s.order[n.left + i]];
				V3 p0 = b.to_local(t[0]), p1 = b.to_local(t[1]), p2 = b.to_local(t[2]);
				SatResult r = box_tri_sat(b.h, p0, p1, p2);
				if (r.hit) max_pen = std::max(max_pen, -r.sat_gap);
				else if (r.sat_gap < min_dist) min_dist = std::min(min_dist, box_tri_distance_separated(b.h, p0, p1, p2));
			}
		} else {
			stack[sp++] = n.right;
			stack[sp++] = n.left;
		}
	}
	return max_pen >= 0.0 ? -max_pen : min_dist;
}

/// The same two queries for any link shape.  Boxes go through the functions above; a capsule or convex polytope walks the tree
/// with its bounding box (conservative) and runs its own exact test on the triangles of the leaves it reaches.
inline Obb geom_bounding_box(const Tf &shape_tf, const BoxGeom &g) {
	if (g.kind == GEOM_CAPSULE) return Obb(shape_tf, {2 * g.radius, 2 * g.radius, g.length + 2 * g.radius});
	if (g.kind == GEOM_CONVEX) {
		Aabb bb;
		for (auto &v: g.verts) bb.grow(v);
		return Obb(then(shape_tf, Tf{(bb.lo + bb.hi) * 0.5, {0, 0, 0, 1}}), bb.hi - bb.lo);
	}
	return Obb(shape_tf, g.size);
}
/// Signed clearance of one shape/triangle pair (triangle given in the world frame); > 0 separated, <= 0 in contact.
inline double geom_tri_clearance(const Tf &shape_tf, const BoxGeom &g, const std::array<V3, 3> &t) {
	const Obb fr(shape_tf, g.kind == GEOM_BOX ? g.size : V3{0, 0, 0});
	const V3 p0 = fr.to_local(t[0]), p1 = fr.to_local(t[1]), p2 = fr.to_local(t[2]);
	if (g.kind == GEOM_CAPSULE) return capsule_tri_clearance(g.radius, g.length, p0, p1, p2);
	if (g.kind == GEOM_CONVEX) return convex_tri_sat(g.verts, p0, p1, p2).sat_gap;
	return box_tri_clearance(fr.h, p0, p1, p2);
}
inline bool geom_hits_scene(const Scene &s, const Tf &shape_tf, const BoxGeom &g, Counters *cnt = nullptr) {
	if (g.kind == GEOM_BOX) return obb_hits_scene(s, Obb(shape_tf, g.size), cnt);
	if (s.nodes.empty()) return false;
	const Obb b = geom_bounding_box(shape_tf, g);
	int stack[128];
	int sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const Scene::Node &n = s.nodes[stack[--sp]];
		if (cnt) cnt->nodes_visited++;
		if (!b.may_touch(n.box, 0.0)) continue;
		if (n.leaf) {
			for (int i = 0; i < n.right; ++i) {
				if (cnt) cnt->leaf_tests++;
				if (geom_tri_clearance(shape_tf, g, s.tris[s.order[n.left + i]]) <= 0.0) return true;
			}
		} else {
			stack[sp++] = n.right;
			stack[sp++] = n.left;
		}
	}
	return false;
}
inline double geom_scene_clearance(const Scene &s, const Tf &shape_tf, const BoxGeom &g, double reach) {
	if (g.kind == GEOM_BOX) return obb_scene_clearance(s, Obb(shape_tf, g.size), reach);
	double min_dist = reach, max_pen = -1.0;
	if (s.nodes.empty()) return reach;
	const Obb b = geom_bounding_box(shape_tf, g);
	int stack[128];
	int sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const Scene::Node &n = s.nodes[stack[--sp]];
		if (!b.may_touch(n.box, reach)) continue;
		if (n.leaf) {
			for (int i = 0; i < n.right; ++i) {
				const double cl = geom_tri_clearance(shape_tf, g, s.tris[s.order[n.left + i]]);
				if (cl <= 0.0) max_pen = std::max(max_pen, -cl);
				else min_dist = std::min(min_dist, cl);
			}
		} else {
			stack[sp++] = n.right;
			stack[sp++] = n.left;
		}
	}
	return max_pen >= 0.0 ? -max_pen : min_dist;
}

// ---------------------------------------------------------------------------------------------------------
// The four check_* loops: src/planning/collision_detection.cpp:16-122
// ---------------------------------------------------------------------------------------------------------
/// collision_detection.cpp:16-55 — every box of the link, world pose = link_tf.then(shape_tf), OR with early break.
inline bool check_link_collision(const Link &link, const Scene &scene, const Tf &link_tf, Counters *cnt = nullptr) {
	for (auto &g: link.boxes)
		if (geom_hits_scene(scene, then(link_tf, g.tf), g, cnt)) return true;
	return false;
}
/// collision_detection.cpp:57-80 — FK from "flying_base", links in index order, early break.
inline bool check_robot_collision(const Robot &robot, const Scene &scene, const State &st, Counters *cnt = nullptr) {
	auto fk = forward_kinematics(robot, st.joints.data(), robot.find_link("flying_base"), st.base);
	if (cnt) cnt->states_checked++;
	for (size_t i = 0; i < fk.size(); ++i)
		if (check_link_collision(robot.links[i], scene, fk[i], cnt)) return true;
	return false;
}
/// Signed clearance of a whole robot state (min over boxes; see obb_scene_clearance).  Test-side only: used to
/// decide whether a query sits inside the 1e-6 m band.
inline double robot_clearance(const Robot &robot, const Scene &scene, const State &st, double reach) {
	auto fk = forward_kinematics(robot, st.joints.data(), robot.find_link("flying_base"), st.base);
	double best = reach;
	for (size_t i = 0; i < fk.size(); ++i)
		for (auto &g: robot.links[i].boxes)
			best = std::min(best, geom_scene_clearance(scene, then(fk[i], g.tf), g, reach));
	return best;
}

/// n_steps exactly as collision_detection.cpp:88-92: size_t(ceil(d / 0.1)).
inline size_t motion_steps(double d, double max_step) { return (size_t) std::ceil(d / max_step); }

struct MotionResult {
	bool hit = false;
	double toi = 0;           // first colliding t
	size_t n_steps = 0;
	size_t states_checked = 0;
};
/// collision_detection.cpp:82-105.  Deviation (documented in DESIGN.md): for d == 0 the reference
/// evaluates t = 0/0 = NaN and checks a NaN pose once; here (and in the CUDA path) that case checks state1 once
/// at t = 0.
inline MotionResult check_motion_collides(const Robot &robot, const Scene &scene, const State &a, const State &b,
										  double max_step = 0.1, Counters *cnt = nullptr) {
	MotionResult r;
	double d = equal_weights_distance(a, b);
	r.n_steps = motion_steps(d, max_step);
	for (size_t i = 0; i <= r.n_steps; ++i) {
		double t = r.n_steps ? (double) i / (double) r.n_steps : 0.0;
		State s = r.n_steps ? interpolate(a, b, t) : a;
		r.states_checked++;
		if (check_robot_collision(robot, scene, s, cnt)) {
			r.hit = true;
			r.toi = t;
			return r;
		}
	}
	return r;
}
/// collision_detection.cpp:107-122 — first colliding segment index, or -1.
inline long check_path_collides(const Robot &robot, const Scene &scene, const std::vector<State> &path,
								double max_step = 0.1) {
	for (size_t i = 0; i + 1 < path.size(); ++i)
		if (check_motion_collides(robot, scene, path[i], path[i + 1], max_step).hit) return (long) i;
	return -1;
}

// ---------------------------------------------------------------------------------------------------------
// Samplers: src/planning/RandomNumberGenerator.h:21-77, state_tools.cpp:14-75, goal_sampling.cpp:13-124,
// spherical_geometry.h:142-164
// ---------------------------------------------------------------------------------------------------------
struct Rng {  // std::mt19937 façade; distribution objects are built per call exactly as the reference does
	std::mt19937 gen;
	explicit Rng(unsigned seed) : gen(seed) {}
	double uniform_real(double lo, double hi) {
		std::uniform_real_distribution<double> d(lo, hi);
		return d(gen);
	}
	double gaussian01() {
		std::normal_distribution<double> d(0.0, 1.0);
		return d(gen);
	}
	int uniform_integer(int lo, int hi) {  // RandomNumberGenerator.h:48-52
		std::uniform_int_distribution<int> d(lo, hi);
		return d(gen);
	}
};

/// state_tools.cpp:44-67 — one joint value per *joint* (fixed joints included).
inline State generate_uniform_random_state(const Robot &robot, Rng &rng, double h_range, double v_range,
										   double joint_range = M_PI / 2.0) {
	State s;
	double tx = rng.uniform_real(-h_range, h_range);
	double ty = rng.uniform_real(-h_range, h_range);
	double tz = rng.uniform_real(0, v_range);
	s.base.t = {tx, ty, tz};
	s.base.q = from_axis_angle({0, 0, 1}, rng.uniform_real(0, 2 * M_PI));
	for (size_t i = 0; i < robot.joints.size(); ++i) s.joints.push_back(rng.uniform_real(-joint_range, joint_range));
	return s;
}
/// goal_sampling.cpp:13-49 — yaw U(-π,π); one value per revolute joint in [min,max].
inline State gen_upright_state(const Robot &robot, Rng &rng) {
	State s;
	s.base.t = {0, 0, 0};
	s.base.q = from_axis_angle({0, 0, 1}, rng.uniform_real(-M_PI, M_PI));
	for (auto &j: robot.joints)
		if (j.kind == REVOLUTE) s.joints.push_back(rng.uniform_real(j.min_angle, j.max_angle));
	return s;
}
/// The deterministic half of genGoalStateUniform (goal_sampling.cpp:51-79): given an upright state, shift the
/// base so the end effector lands on `target` (minus distance_from_target along the arm's -y).
inline State project_upright_to_goal(const Robot &robot, State s, const V3 &target, double distance_from_target,
									 int flying_base, int end_effector) {
	auto fk = forward_kinematics(robot, s.joints.data(), flying_base, s.base);
	const Tf &ee = fk[end_effector];
	V3 delta = target - ee.t;
	if (distance_from_target > 0.0) delta = delta + qrot(ee.q, V3{0.0, -distance_from_target, 0.0});
	s.base.t = s.base.t + delta;
	return s;
}
inline State gen_goal_state_uniform(Rng &rng, const V3 &target, double distance_from_target, const Robot &robot,
									int flying_base, int end_effector) {
	return project_upright_to_goal(robot, gen_upright_state(robot, rng), target, distance_from_target, flying_base,
								   end_effector);
}
/// spherical_geometry.h:142-164.
inline double latitude(const V3 &v) { return std::atan2(v.z, std::sqrt(v.x * v.x + v.y * v.y)); }
inline double longitude(const V3 &v) { return std::atan2(v.y, v.x); }
/// state_tools.cpp:14-42 — single-joint robots only.
inline State from_end_effector_and_vector(const Robot &robot, const V3 &ee_point, const V3 &vec) {
	State s;
	s.joints = {-latitude(vec)};
	s.base.t = {0, 0, 0};
	s.base.q = from_axis_angle({0, 0, 1}, longitude(vec) + M_PI / 2.0);
	auto fk = forward_kinematics(robot, s.joints.data(), robot.find_link("flying_base"), s.base);
	V3 ee = fk[robot.find_link("end_effector")].t;
	s.base.t = s.base.t + ee_point - ee;
	return s;
}

/// state_tools.cpp:69-75 — the direction the arm points in: the end effector's orientation applied to -y.
inline V3 arm_vector_from_state(const Robot &robot, const State &s) {
	auto fk = forward_kinematics(robot, s.joints.data(), robot.find_link("flying_base"), s.base);
	return qrot(fk[robot.find_link("end_effector")].q, V3{0.0, -1.0, 0.0});
}
using StateFn = std::function<bool(const State &)>;
/// goal_sampling.cpp:101-124 — up to max_attempts x (three fresh Gaussians -> unit vector -> fromEndEffectorAndVector ->
/// state check); the first collision-free candidate is returned.
inline std::optional<State> generate_uniform_random_arm_vector_state(const Robot &robot, const StateFn &collides, const V3 &fruit_center,
																	 Rng &rng, int max_attempts, double ee_distance) {
	for (int attempt = 0; attempt < max_attempts; ++attempt) {
		// The reference writes `math::Vec3d arm_vec(rng.gaussian01(), rng.gaussian01(), rng.gaussian01())`: the order in which
		// the three arguments are evaluated is unspecified in C++.  clang (the reference's own toolchain, flake.nix:19-28)
		// goes left to right, GCC right to left; the restatement follows whichever compiler builds it, so that it agrees
		// with oracle/_ref (the reference's sources under the same compiler) draw for draw.
#if defined(__clang__)
		const double gx = rng.gaussian01(), gy = rng.gaussian01(), gz = rng.gaussian01();
#else
		const double gz = rng.gaussian01(), gy = rng.gaussian01(), gx = rng.gaussian01();
#endif
		const double norm = std::sqrt(gx * gx + gy * gy + gz * gz);  // Vec3d::normalized (Vec3.h:276-278): component / norm
		const V3 v{gx / norm, gy / norm, gz / norm};
		State cand = from_end_effector_and_vector(robot, fruit_center - v * ee_distance, V3{-v.x, -v.y, -v.z});
		if (!collides(cand)) return cand;
	}
	return std::nullopt;
}
/// goal_sampling.cpp:81-99 — findGoalStateByUniformSampling.
inline std::optional<State> find_goal_state_by_uniform_sampling(const V3 &target, const Robot &robot, int flying_base, int end_effector,
																const StateFn &collides, Rng &rng, size_t max_attempts) {
	for (size_t i = 0; i < max_attempts; ++i) {
		State s = gen_goal_state_uniform(rng, target, 0.0, robot, flying_base, end_effector);
		if (!collides(s)) return s;
	}
	return std::nullopt;
}
/// goal_sampling.cppm:37-56 — at most max_samples x (sample -> state check -> operation); the first non-empty result wins.
template<class T>
std::optional<T> try_at_valid_goal_samples(const StateFn &collides, const std::function<State()> &goal_sample, int max_samples,
										   const std::function<std::optional<T>(const State &)> &operation) {
	for (int i = 0; i < max_samples; ++i) {
		State goal = goal_sample();
		if (!collides(goal))
			if (auto r = operation(goal)) return r;
	}
	return std::nullopt;
}

// ---------------------------------------------------------------------------------------------------------
// Path points and local optimisation (shortcutting): src/planning/RobotPath.h:116-170, RobotPath.cpp:13-23,
// local_optimization.cpp:23-174, and the 100-iteration loop of tsp_over_prm.cpp:594-612.
// ---------------------------------------------------------------------------------------------------------
using Path = std::vector<State>;
using MotionFn = std::function<bool(const State &, const State &)>;

struct PathPoint {  // RobotPath.h:116-170
	size_t segment_i = 0;
	double segment_t = 0.0;
	double to_scalar() const { return (double) segment_i + segment_t; }
	static PathPoint from_scalar(double d, const Path &path) {
		PathPoint p;
		p.segment_i = (size_t) std::floor(d);
		p.segment_i = std::min(p.segment_i, path.size() - 2);
		p.segment_t = d - (double) p.segment_i;
		return p;
	}
	PathPoint adjust_by_scalar(double d, const Path &path) const { return from_scalar(to_scalar() + d, path); }
};
/// RobotPath.cpp:13-23.
inline State interpolate(const PathPoint &pt, const Path &path) {
	return interpolate(path[pt.segment_i], path[pt.segment_i + 1], pt.segment_t);
}
/// local_optimization.cpp:23-45.
inline bool try_shortcut_by_deleting_waypoint(Path &path, size_t waypoint, const MotionFn &motion_collides) {
	if (motion_collides(path[waypoint - 1], path[waypoint + 1])) return false;
	path.erase(path.begin() + (long) waypoint);
	return true;
}
/// local_optimization.cpp:47-76 — note the two motions checked are (current → pulled) and (pulled → next).
inline bool try_midpoint_pull(Path &path, size_t waypoint, double pull_factor, const MotionFn &motion_collides) {
	State midpoint = interpolate(path[waypoint - 1], path[waypoint + 1], 0.5);
	State pulled = interpolate(path[waypoint], midpoint, pull_factor);
	if (motion_collides(path[waypoint], pulled) || motion_collides(pulled, path[waypoint + 1])) return false;
	path[waypoint] = pulled;
	return true;
}
/// local_optimization.cpp:78-111.
inline bool try_shortcut_between_path_points(Path &path, const PathPoint &start, const PathPoint &end,
											 const MotionFn &motion_collides) {
	State st1 = interpolate(start, path), st2 = interpolate(end, path);
	if (motion_collides(st1, st2)) return false;
	Path out(path.begin(), path.begin() + (long) start.segment_i);
	out.push_back(st1);
	out.push_back(st2);
	out.insert(out.end(), path.begin() + (long) end.segment_i + 1, path.end());
	path = std::move(out);
	return true;
}
/// local_optimization.cpp:113-117.
inline PathPoint generate_random_path_point(const Path &path, Rng &rng) {
	PathPoint p;
	p.segment_i = (size_t) rng.uniform_integer(0, (int) (path.size() - 2));
	p.segment_t = rng.uniform_real(0.0, 1.0);
	return p;
}
/// local_optimization.cpp:119-136.
inline bool try_deleting_every_waypoint(Path &path, const MotionFn &motion_collides) {
	size_t cursor = 1;
	bool shortened = false;
	while (cursor + 1 < path.size()) {
		if (try_shortcut_by_deleting_waypoint(path, cursor, motion_collides)) shortened = true;
		else cursor++;
	}
	return shortened;
}
/// local_optimization.cpp:138-153.
inline bool try_shortcutting_randomly_locally(Path &path, Rng &rng, const MotionFn &motion_collides) {
	PathPoint middle = generate_random_path_point(path, rng);
	double radius = rng.uniform_real(0.5, 2.0);
	PathPoint start = middle.adjust_by_scalar(-radius, path), end = middle.adjust_by_scalar(radius, path);
	return try_shortcut_between_path_points(path, start, end, motion_collides);
}
/// local_optimization.cpp:155-173.
inline bool try_shortcutting_randomly_globally(Path &path, const MotionFn &motion_collides, Rng &rng) {
	PathPoint start = generate_random_path_point(path, rng), end = generate_random_path_point(path, rng);
	if (start.to_scalar() > end.to_scalar()) std::swap(start, end);
	return try_shortcut_between_path_points(path, start, end, motion_collides);
}
/// tsp_over_prm.cpp:594-612 — `iterations` attempts, stopping early once the path is trivial; returns the number
/// of attempts that shortened the path.
inline size_t optimize_path_segment(Path &path, const MotionFn &motion_collides, Rng &rng, size_t iterations = 100) {
	size_t n_success = 0;
	for (size_t i = 0; i < iterations; ++i) {
		if (path.size() <= 2) break;
		if (try_shortcutting_randomly_globally(path, motion_collides, rng)) ++n_success;
	}
	return n_success;
}

// ---------------------------------------------------------------------------------------------------------
// Roadmap shortest paths: runDijkstra (src/planning/tsp_over_prm.cpp:80-97) = boost::dijkstra_shortest_paths on
// the undirected PRMGraph (tsp_over_prm.h:25-26), retrace_path (:111-128).  Boost.Graph is a third-party
// dependency absent from /root/reference (nixpkgs 24.05 pins boost 1.81): PARITY UNPINNED for this function.
// Restated from its published algorithm: labels start at numeric_limits<double>::max() ("infinity") with every
// vertex its own predecessor; the closest unsettled vertex is settled next; an edge (u, v, w) relaxes v when
// d[u] + w < d[v] (strict), recording u as predecessor.  Distances do not depend on the heap's tie order;
// predecessors do when two relaxations tie exactly.
// ---------------------------------------------------------------------------------------------------------
struct Roadmap {
	struct Arc {
		uint32_t to;
		double weight;
	};
	std::vector<std::vector<Arc>> out;  // both directions of every undirected edge, in insertion order
	explicit Roadmap(size_t n) : out(n) {}
	void add_edge(uint32_t a, uint32_t b, double w) {
		out[a].push_back({b, w});
		out[b].push_back({a, w});
	}
};
inline void run_dijkstra(const Roadmap &g, uint32_t source, std::vector<double> &dist, std::vector<uint32_t> &pred) {
	const double inf = std::numeric_limits<double>::max();
	dist.assign(g.out.size(), inf);
	pred.resize(g.out.size());
	std::iota(pred.begin(), pred.end(), 0u);
	dist[source] = 0.0;
	using Entry = std::pair<double, uint32_t>;
	std::priority_queue<Entry, std::vector<Entry>, std::greater<Entry>> heap;
	heap.push({0.0, source});
	std::vector<char> settled(g.out.size(), 0);
	while (!heap.empty()) {
		auto [d, u] = heap.top();
		heap.pop();
		if (settled[u]) continue;
		settled[u] = 1;
		for (const auto &arc: g.out[u]) {
			const double cand = d + arc.weight;
			if (cand < dist[arc.to]) {
				dist[arc.to] = cand;
				pred[arc.to] = u;
				heap.push({cand, arc.to});
			}
		}
	}
}
/// tsp_over_prm.cpp:111-128 — vertices from just after the start node up to the goal (the start itself is not emitted).
inline std::vector<uint32_t> retrace_path(const std::vector<uint32_t> &pred, uint32_t goal) {
	std::vector<uint32_t> path;
	for (uint32_t cur = goal; pred[cur] != cur; cur = pred[cur]) path.push_back(cur);
	std::reverse(path.begin(), path.end());
	return path;
}

// ---------------------------------------------------------------------------------------------------------
// Fruit separation: connected_vertex_components (src/experiment_utils/mesh_connected_components.cpp:25-74).
// The reference merges vertex lists triangle by triangle (edges 0-1 and 1-2 of each triangle) and returns the
// components in unordered_map order; restated with the same merging rule, reported in canonical form:
// label[v] = smallest vertex index of v's component.
// ---------------------------------------------------------------------------------------------------------
inline std::vector<uint32_t> mesh_component_labels(const uint32_t *tris, size_t nt, size_t nv) {
	std::vector<uint32_t> id(nv);
	std::iota(id.begin(), id.end(), 0u);
	std::vector<std::vector<uint32_t>> members(nv);
	for (size_t v = 0; v < nv; ++v) members[v] = {(uint32_t) v};
	for (size_t t = 0; t < nt; ++t)
		for (int i: {0, 1}) {
			const uint32_t a = id[tris[3 * t + i]], b = id[tris[3 * t + i + 1]];
			if (a == b) continue;
			for (uint32_t v: members[b]) members[a].push_back(v), id[v] = a;
			members[b].clear();
		}
	std::vector<uint32_t> label(nv);
	for (size_t c = 0; c < nv; ++c) {
		if (members[c].empty()) continue;
		const uint32_t lo = *std::min_element(members[c].begin(), members[c].end());
		for (uint32_t v: members[c]) label[v] = lo;
	}
	return label;
}

}  // namespace orc
