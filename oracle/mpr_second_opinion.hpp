// ORACLE — TEST INFRASTRUCTURE ONLY.  A "libccd-style" second opinion on box / triangle contact (last row).
//
// What the reference's narrow phase does for one (box link, mesh triangle) pair, as far as it can be told without the sources
// (FCL 0.7.x + libccd 2.1, both absent here and in /root/reference): fcl::collide of a Boxd against a BVHModel<OBBd> with the
// default CollisionRequest (no contacts wanted) ends, per leaf, in the GJK solver's shape-triangle test, which for the default
// libccd solver is libccd's Minkowski Portal Refinement *intersection* query (tolerance 1e-6 — the number behind the band of
// BASELINE.json's north_star).  This file restates that PUBLISHED algorithm (Snethen's MPR / XenoCollide as libccd implements it:
// portal discovery from the centre difference, portal refinement until the portal encloses the origin or cannot advance by more
// than the tolerance) with the support mappings FCL gives the two shapes: the box's sign corner in the box frame — a direction
// component within machine epsilon of zero yields the face / edge centre, not a corner — and the triangle's arg-max vertex,
// centres = box centre and triangle centroid.
//
// It is NOT FCL and NOT pinned to anything: its output is labelled "libccd-style", never "FCL".  The product and the parity
// verdicts use the exact separating-axis test (mgodpl_oracle.hpp); this exists to EXPLAIN how an MPR-based narrow phase answers
// the cases inside the band, and to measure how wide the band of disagreement really is (tests/test_oracle.py).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "mgodpl_oracle.hpp"

namespace orc::mpr {

constexpr double EPS = DBL_EPSILON;
inline bool is_zero(double x) { return std::fabs(x) < EPS; }
/// Relative equality, the way libccd compares reals.
inline bool nearly_equal(double a, double b) {
	const double ab = std::fabs(a - b);
	if (ab < EPS) return true;
	const double m = std::max(std::fabs(a), std::fabs(b));
	return ab < EPS * m;
}
inline bool same_point(const V3 &a, const V3 &b) { return nearly_equal(a.x, b.x) && nearly_equal(a.y, b.y) && nearly_equal(a.z, b.z); }
inline V3 unit(const V3 &v) { return v * (1.0 / norm(v)); }
inline int sign0(double x) { return is_zero(x) ? 0 : (x < 0 ? -1 : 1); }

struct Pair {
	Tf box;      // world pose of the box
	V3 half;     // half extents
	V3 tri[3];   // world coordinates
	V3 centroid;
	/// Support point of (box - triangle) in direction d.
	V3 support(const V3 &d) const {
		const V3 l = qrot(conj(box.q), d);
		const V3 corner{sign0(l.x) * half.x, sign0(l.y) * half.y, sign0(l.z) * half.z};
		const V3 on_box = qrot(box.q, corner) + box.t;
		const V3 nd = -d;
		int best = 0;
		double best_dot = -DBL_MAX;
		for (int i = 0; i < 3; ++i) {
			const double s = dot(nd, tri[i] - centroid);
			if (s > best_dot) best_dot = s, best = i;
		}
		return on_box - tri[best];
	}
};

struct Outcome {
	int intersect;    // 1: the query reports contact
	int iterations;   // refinement rounds taken (0: decided during portal discovery)
	int stage;        // 0 discovery said "apart", 1 discovery said "touching / centre on the origin ray", 2 refinement enclosed the
	                  // origin, 3 refinement could not reach the origin, 4 refinement stopped at the tolerance, 5 iteration guard
};

inline Outcome intersect(const Pair &p, double tolerance = 1e-6, int max_iterations = 500) {
	V3 v0 = p.box.t - p.centroid, v1, v2, v3;
	if (same_point(v0, V3{0, 0, 0})) v0.x += EPS * 10.0;  // centres coincide: nudge, as the published code does
	// -- portal discovery ------------------------------------------------------------------------------------------------------
	V3 dir = unit(-v0);
	v1 = p.support(dir);
	double d = dot(v1, dir);
	if (is_zero(d) || d < 0) return {0, 0, 0};
	dir = cross(v0, v1);
	if (is_zero(dot(dir, dir))) return {1, 0, 1};  // the origin lies on the ray through v0 and v1 (or on v1 itself)
	dir = unit(dir);
	v2 = p.support(dir);
	d = dot(v2, dir);
	if (is_zero(d) || d < 0) return {0, 0, 0};
	dir = unit(cross(v1 - v0, v2 - v0));
	if (dot(dir, v0) > 0) {  // orient the portal away from the centre
		std::swap(v1, v2);
		dir = -dir;
	}
	for (int guard = 0;; ++guard) {
		if (guard > max_iterations) return {0, 0, 5};
		v3 = p.support(dir);
		d = dot(v3, dir);
		if (is_zero(d) || d < 0) return {0, 0, 0};
		// the origin ray must pass through the triangle (v1, v2, v3): otherwise replace a vertex and look again
		double side = dot(cross(v1, v3), v0);
		bool again = false;
		if (side < 0 && !is_zero(side)) v2 = v3, again = true;
		if (!again) {
			side = dot(cross(v3, v2), v0);
			if (side < 0 && !is_zero(side)) v1 = v3, again = true;
		}
		if (!again) break;
		dir = unit(cross(v1 - v0, v2 - v0));
	}
	// -- portal refinement -----------------------------------------------------------------------------------------------------
	for (int it = 1;; ++it) {
		dir = unit(cross(v2 - v1, v3 - v1));
		d = dot(dir, v1);
		if (is_zero(d) || d > 0) return {1, it, 2};  // the origin is on the inner side of the portal
		const V3 v4 = p.support(dir);
		const double d4 = dot(v4, dir);
		if (!(is_zero(d4) || d4 > 0)) return {0, it, 3};  // the support plane does not reach the origin
		const double advance = std::min({d4 - dot(v1, dir), d4 - dot(v2, dir), d4 - dot(v3, dir)});
		if (nearly_equal(advance, tolerance) || advance < tolerance) return {0, it, 4};
		if (it >= max_iterations) return {0, it, 5};
		// keep the origin ray inside the new portal
		const V3 c = cross(v4, v0);
		if (dot(v1, c) > 0) {
			if (dot(v2, c) > 0) v1 = v4;
			else v3 = v4;
		} else {
			if (dot(v3, c) > 0) v2 = v4;
			else v1 = v4;
		}
	}
}

inline Outcome box_triangle(const Tf &box_tf, const V3 &size, const V3 &a, const V3 &b, const V3 &c, double tolerance = 1e-6) {
	Pair p;
	p.box = box_tf, p.half = size * 0.5;
	p.tri[0] = a, p.tri[1] = b, p.tri[2] = c;
	p.centroid = (a + b + c) * (1.0 / 3.0);
	return intersect(p, tolerance);
}

}  // namespace orc::mpr
