// fp64 rigid-body math, forward kinematics and the exact box/triangle test; fp32 conservative culling.
//
// Semantics follow the reference (paths relative to its checkout): src/math/Quaternion.h:25-85, Transform.h:20-78,
// src/planning/RobotModel.cpp:17-90, RobotState.cpp:14-24.  The arithmetic is free to differ in the last bits
// (rotate-by-cross-products instead of the sandwich product, fused multiply-adds): every boolean this path returns is
// decided in fp64 with ~1e-15 m of slack, against a 1e-6 m parity band.
#pragma once
#include "device_types.cuh"

namespace mgodpl_b200 {

struct D3 {
	double x, y, z;
};
struct Q4 {
	double x, y, z, w;
};

__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
	return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

/// Hamilton product (Quaternion.h:36-44).
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
	return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
			a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ Q4 qconj(Q4 q) { return {-q.x, -q.y, -q.z, q.w}; }
/// q (v,0) q* for a unit quaternion (Quaternion.h:46-51), evaluated as v + 2w(u x v) + 2u x (u x v).
__device__ __forceinline__ D3 qrot(Q4 q, D3 v) {
	D3 u{q.x, q.y, q.z};
	D3 t = cross(u, v);
	t = t + t;
	return v + t * q.w + cross(u, t);
}
/// Rotation matrix of a unit quaternion, world = R * local (what Eigen builds for fcl::Transform3d::rotate,
/// collision_detection.cpp:29-30).  Row-major r[3*i+j].
__device__ __forceinline__ void qmat(Q4 q, double *r) {
	double x = q.x, y = q.y, z = q.z, w = q.w;
	r[0] = 1 - 2 * (y * y + z * z), r[1] = 2 * (x * y - z * w), r[2] = 2 * (x * z + y * w);
	r[3] = 2 * (x * y + z * w), r[4] = 1 - 2 * (x * x + z * z), r[5] = 2 * (y * z - x * w);
	r[6] = 2 * (x * z - y * w), r[7] = 2 * (y * z + x * w), r[8] = 1 - 2 * (x * x + y * y);
}

struct Pose {
	D3 t;
	Q4 q;
};
/// Transformd::then (Transform.h:31-36).
__device__ __forceinline__ Pose then(const Pose &a, D3 bt, Q4 bq, bool b_rot_identity) {
	Pose r;
	r.t = a.t + qrot(a.q, bt);
	r.q = b_rot_identity ? a.q : qmul(a.q, bq);
	return r;
}

/// One op of the host-unrolled forwardKinematics DFS: tf[child] = tf[parent] . pre . J(theta) . post.
__device__ __forceinline__ Pose fk_apply(const FkOp &op, const Pose &parent, const double *joint_values) {
	Pose p = then(parent, D3{op.pre_t[0], op.pre_t[1], op.pre_t[2]}, Q4{op.pre_q[0], op.pre_q[1], op.pre_q[2], op.pre_q[3]},
				  op.flags & FK_PRE_ROT_ID);
	if (op.var >= 0) {
		// RobotModel::variableTransform (RobotModel.cpp:136-154) + Quaterniond::fromAxisAngle (Quaternion.h:53-62)
		double s, c;
		sincos(joint_values[op.var] / 2.0, &s, &c);
		if (op.flags & FK_INVERT) s = -s;
		p.q = qmul(p.q, Q4{op.axis[0] * s, op.axis[1] * s, op.axis[2] * s, c});
	}
	return then(p, D3{op.post_t[0], op.post_t[1], op.post_t[2]}, Q4{op.post_q[0], op.post_q[1], op.post_q[2], op.post_q[3]},
				op.flags & FK_POST_ROT_ID);
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 culling box: an oriented box expressed relative to SceneDev::origin, half extents already inflated by the
// scene's conservative margin.  19 registers.
// ---------------------------------------------------------------------------------------------------------------
struct CullBox {
	float cx, cy, cz;
	float r[9];        // row-major, world = R * local
	float hx, hy, hz;  // inflated half extents
	float wx, wy, wz;  // half extents of the box's world-axis AABB (from the inflated h)
};

/// Second half of make_cull_box: inflated half extents and the world-axis AABB, given centre and rotation.
__device__ __forceinline__ void finish_cull_box(const SceneDev &sc, CullBox &b, float h0, float h1, float h2) {
	b.hx = h0 + sc.margin, b.hy = h1 + sc.margin, b.hz = h2 + sc.margin;
	b.wx = fabsf(b.r[0]) * b.hx + fabsf(b.r[1]) * b.hy + fabsf(b.r[2]) * b.hz;
	b.wy = fabsf(b.r[3]) * b.hx + fabsf(b.r[4]) * b.hy + fabsf(b.r[5]) * b.hz;
	b.wz = fabsf(b.r[6]) * b.hx + fabsf(b.r[7]) * b.hy + fabsf(b.r[8]) * b.hz;
}

__device__ __forceinline__ CullBox make_cull_box(const SceneDev &sc, D3 c, Q4 q, const double *half) {
	CullBox b;
	b.cx = (float) (c.x - sc.origin[0]), b.cy = (float) (c.y - sc.origin[1]), b.cz = (float) (c.z - sc.origin[2]);
	// fp32 rotation from the rounded quaternion: entries are good to ~3e-7, i.e. ~3e-6 m over a 10 m lever — the
	// scene margin (>= 2e-5 m, growing with the extent) absorbs it; the exact test never uses this matrix.
	const float x = (float) q.x, y = (float) q.y, z = (float) q.z, w = (float) q.w;
	b.r[0] = 1.f - 2.f * (y * y + z * z), b.r[1] = 2.f * (x * y - z * w), b.r[2] = 2.f * (x * z + y * w);
	b.r[3] = 2.f * (x * y + z * w), b.r[4] = 1.f - 2.f * (x * x + z * z), b.r[5] = 2.f * (y * z - x * w);
	b.r[6] = 2.f * (x * z - y * w), b.r[7] = 2.f * (y * z + x * w), b.r[8] = 1.f - 2.f * (x * x + y * y);
	finish_cull_box(sc, b, (float) half[0], (float) half[1], (float) half[2]);
	return b;
}

/// Conservative oriented-box vs AABB(centre, half) overlap: the 3 world axes and the 3 box axes of the
/// separating-axis test (the 9 edge-edge axes are skipped, which can only report extra overlaps).
__device__ __forceinline__ bool cull_overlap(const CullBox &b, float ncx, float ncy, float ncz, float nhx, float nhy,
											 float nhz) {
	float dx = ncx - b.cx, dy = ncy - b.cy, dz = ncz - b.cz;
	bool sep = fabsf(dx) > nhx + b.wx;
	sep |= fabsf(dy) > nhy + b.wy;
	sep |= fabsf(dz) > nhz + b.wz;
	float p0 = b.r[0] * dx + b.r[3] * dy + b.r[6] * dz;
	float p1 = b.r[1] * dx + b.r[4] * dy + b.r[7] * dz;
	float p2 = b.r[2] * dx + b.r[5] * dy + b.r[8] * dz;
	float e0 = fabsf(b.r[0]) * nhx + fabsf(b.r[3]) * nhy + fabsf(b.r[6]) * nhz + b.hx;
	float e1 = fabsf(b.r[1]) * nhx + fabsf(b.r[4]) * nhy + fabsf(b.r[7]) * nhz + b.hy;
	float e2 = fabsf(b.r[2]) * nhx + fabsf(b.r[5]) * nhy + fabsf(b.r[8]) * nhz + b.hz;
	sep |= fabsf(p0) > e0;
	sep |= fabsf(p1) > e1;
	sep |= fabsf(p2) > e2;
	return !sep;
}

/// fp32 pre-filter of the narrow phase: the same 13 separating axes as the exact test, evaluated in fp32 on
/// origin-relative coordinates against the *inflated* culling box.  It may only answer "certainly separated": the
/// inflation (SceneDev::margin, >= 10x the fp32 error of these projections at scene scale) turns every rounding
/// error into a slightly larger box, and any axis — even an inexact one — that separates the triangle from the larger
/// box separates it from the real one.  Everything else (overlap or a near miss) goes to the fp64 test.
enum TriVerdict : int { TRI_SEPARATED = 0, TRI_HIT = 1, TRI_UNKNOWN = 2 };
/// The pre-filter can also answer "certainly hit" in the one situation where that is trivially sound: a triangle
/// vertex lies inside the box shrunk by the margin.  (Most real contacts of small tree triangles with a link box are
/// of this kind, so the fp64 test is left with grazing configurations only.)
__device__ __forceinline__ int tri_classify(const CullBox &b, float margin, float4 v0, float4 v1, float4 v2) {
	float px[3], py[3], pz[3];
	const float4 v[3] = {v0, v1, v2};
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float dx = v[k].x - b.cx, dy = v[k].y - b.cy, dz = v[k].z - b.cz;
		px[k] = b.r[0] * dx + b.r[3] * dy + b.r[6] * dz;
		py[k] = b.r[1] * dx + b.r[4] * dy + b.r[7] * dz;
		pz[k] = b.r[2] * dx + b.r[5] * dy + b.r[8] * dz;
	}
	if (fminf(px[0], fminf(px[1], px[2])) > b.hx || fmaxf(px[0], fmaxf(px[1], px[2])) < -b.hx) return TRI_SEPARATED;
	if (fminf(py[0], fminf(py[1], py[2])) > b.hy || fmaxf(py[0], fmaxf(py[1], py[2])) < -b.hy) return TRI_SEPARATED;
	if (fminf(pz[0], fminf(pz[1], pz[2])) > b.hz || fmaxf(pz[0], fmaxf(pz[1], pz[2])) < -b.hz) return TRI_SEPARATED;
	{
		const float sx = b.hx - 2.f * margin, sy = b.hy - 2.f * margin, sz = b.hz - 2.f * margin;  // true extent minus the margin
#pragma unroll
		for (int k = 0; k < 3; ++k)
			if (fabsf(px[k]) < sx && fabsf(py[k]) < sy && fabsf(pz[k]) < sz) return TRI_HIT;
	}
	const float fx[3] = {px[1] - px[0], px[2] - px[1], px[0] - px[2]};
	const float fy[3] = {py[1] - py[0], py[2] - py[1], py[0] - py[2]};
	const float fz[3] = {pz[1] - pz[0], pz[2] - pz[1], pz[0] - pz[2]};
	{  // triangle plane
		const float nx = fy[0] * fz[1] - fz[0] * fy[1], ny = fz[0] * fx[1] - fx[0] * fz[1], nz = fx[0] * fy[1] - fy[0] * fx[1];
		const float q0 = nx * px[0] + ny * py[0] + nz * pz[0], q1 = nx * px[1] + ny * py[1] + nz * pz[1],
					q2 = nx * px[2] + ny * py[2] + nz * pz[2];
		const float rad = b.hx * fabsf(nx) + b.hy * fabsf(ny) + b.hz * fabsf(nz);
		if (fminf(q0, fminf(q1, q2)) > rad || fmaxf(q0, fmaxf(q1, q2)) < -rad) return TRI_SEPARATED;
	}
#pragma unroll
	for (int k = 0; k < 3; ++k) {  // box axis x edge k: project all three vertices (no "two are equal" shortcut)
		{
			const float q0 = py[0] * fz[k] - pz[0] * fy[k], q1 = py[1] * fz[k] - pz[1] * fy[k], q2 = py[2] * fz[k] - pz[2] * fy[k];
			const float rad = b.hy * fabsf(fz[k]) + b.hz * fabsf(fy[k]);
			if (fminf(q0, fminf(q1, q2)) > rad || fmaxf(q0, fmaxf(q1, q2)) < -rad) return TRI_SEPARATED;
		}
		{
			const float q0 = pz[0] * fx[k] - px[0] * fz[k], q1 = pz[1] * fx[k] - px[1] * fz[k], q2 = pz[2] * fx[k] - px[2] * fz[k];
			const float rad = b.hx * fabsf(fz[k]) + b.hz * fabsf(fx[k]);
			if (fminf(q0, fminf(q1, q2)) > rad || fmaxf(q0, fmaxf(q1, q2)) < -rad) return TRI_SEPARATED;
		}
		{
			const float q0 = px[0] * fy[k] - py[0] * fx[k], q1 = px[1] * fy[k] - py[1] * fx[k], q2 = px[2] * fy[k] - py[2] * fx[k];
			const float rad = b.hx * fabsf(fy[k]) + b.hy * fabsf(fx[k]);
			if (fminf(q0, fminf(q1, q2)) > rad || fmaxf(q0, fmaxf(q1, q2)) < -rad) return TRI_SEPARATED;
		}
	}
	return TRI_UNKNOWN;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact narrow phase: closed oriented box vs closed triangle, 13-axis separating-axis test in the box frame, fp64.
// Stands where FCL's box/triangle test stood behind fcl::collide (collision_detection.cpp:36-45); touching counts
// as contact.  Axes that are numerically null (degenerate triangle, edge parallel to a box axis) carry no
// information and are skipped — the remaining axes are complete for those configurations.
// ---------------------------------------------------------------------------------------------------------------
struct ExactBox {
	double cx, cy, cz;
	double r[9];
	double hx, hy, hz;
};

__device__ __forceinline__ bool axis_separates(double ax, double ay, double az, double floor2, const ExactBox &b, D3 p0,
											   D3 p1, D3 p2) {
	double l2 = ax * ax + ay * ay + az * az;
	double q0 = ax * p0.x + ay * p0.y + az * p0.z;
	double q1 = ax * p1.x + ay * p1.y + az * p1.z;
	double q2 = ax * p2.x + ay * p2.y + az * p2.z;
	double lo = fmin(q0, fmin(q1, q2)), hi = fmax(q0, fmax(q1, q2));
	double rad = b.hx * fabs(ax) + b.hy * fabs(ay) + b.hz * fabs(az);
	return (l2 > floor2) && (fmax(lo - rad, -hi - rad) > 0.0);
}

static __device__ __noinline__ bool box_triangle_hit(const ExactBox &b, const double *__restrict__ tri) {
	D3 p[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		double dx = __ldg(tri + 3 * k) - b.cx, dy = __ldg(tri + 3 * k + 1) - b.cy, dz = __ldg(tri + 3 * k + 2) - b.cz;
		p[k] = {b.r[0] * dx + b.r[3] * dy + b.r[6] * dz, b.r[1] * dx + b.r[4] * dy + b.r[7] * dz,
				b.r[2] * dx + b.r[5] * dy + b.r[8] * dz};
	}
	// box face normals
	if (fmin(p[0].x, fmin(p[1].x, p[2].x)) > b.hx || fmax(p[0].x, fmax(p[1].x, p[2].x)) < -b.hx) return false;
	if (fmin(p[0].y, fmin(p[1].y, p[2].y)) > b.hy || fmax(p[0].y, fmax(p[1].y, p[2].y)) < -b.hy) return false;
	if (fmin(p[0].z, fmin(p[1].z, p[2].z)) > b.hz || fmax(p[0].z, fmax(p[1].z, p[2].z)) < -b.hz) return false;
	D3 f0 = p[1] - p[0], f1 = p[2] - p[1], f2 = p[0] - p[2];
	double s0 = dot(f0, f0), s1 = dot(f1, f1), s2 = dot(f2, f2);
	double scale2 = fmax(s0, fmax(s1, s2));
	D3 n = cross(f0, f1);
	if (axis_separates(n.x, n.y, n.z, 1e-28 * scale2 * scale2, b, p[0], p[1], p[2])) return false;
	const D3 f[3] = {f0, f1, f2};
	const double fl[3] = {s0, s1, s2};
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		double fl2 = 1e-24 * fl[k];
		if (axis_separates(0.0, -f[k].z, f[k].y, fl2, b, p[0], p[1], p[2])) return false;
		if (axis_separates(f[k].z, 0.0, -f[k].x, fl2, b, p[0], p[1], p[2])) return false;
		if (axis_separates(-f[k].y, f[k].x, 0.0, fl2, b, p[0], p[1], p[2])) return false;
	}
	return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Other link shapes (north_star (3); the reference itself only collides boxes, collision_detection.cpp:21-22,49-51).
// Both walk the tree with their bounding box (the same fp32 culling as a box link) and differ at the leaves only.
// ---------------------------------------------------------------------------------------------------------------

/// Squared distance between the segments [p1, q1] and [p2, q2] (clamped closest-point parameters).
__device__ __forceinline__ double segment_segment_dist2(D3 p1, D3 q1, D3 p2, D3 q2) {
	const D3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
	const double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
	double s = 0.0, t = 0.0;
	if (a > 1e-300 && e > 1e-300) {
		const double den = a * e - b * b;
		s = den > 1e-14 * a * e ? fmin(fmax((b * f - c * e) / den, 0.0), 1.0) : 0.0;
		t = (b * s + f) / e;
		if (t < 0.0) t = 0.0, s = fmin(fmax(-c / a, 0.0), 1.0);
		else if (t > 1.0) t = 1.0, s = fmin(fmax((b - c) / a, 0.0), 1.0);
	} else if (a > 1e-300) {
		s = fmin(fmax(-c / a, 0.0), 1.0);
	} else if (e > 1e-300) {
		t = fmin(fmax(f / e, 0.0), 1.0);
	}
	const D3 dd = (p1 + d1 * s) - (p2 + d2 * t);
	return dot(dd, dd);
}

/// Squared distance from point p to triangle (a, b, c): Voronoi-region classification of the closest point.
__device__ __forceinline__ double point_triangle_dist2(D3 p, D3 a, D3 b, D3 c) {
	const D3 ab = b - a, ac = c - a, ap = p - a;
	const double d1 = dot(ab, ap), d2 = dot(ac, ap);
	D3 cl = a;
	do {
		if (d1 <= 0.0 && d2 <= 0.0) break;
		const D3 bp = p - b;
		const double d3 = dot(ab, bp), d4 = dot(ac, bp);
		if (d3 >= 0.0 && d4 <= d3) { cl = b; break; }
		const double vc = d1 * d4 - d3 * d2;
		if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) { cl = a + ab * (d1 / (d1 - d3)); break; }
		const D3 cp = p - c;
		const double d5 = dot(ab, cp), d6 = dot(ac, cp);
		if (d6 >= 0.0 && d5 <= d6) { cl = c; break; }
		const double vb = d5 * d2 - d1 * d6;
		if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { cl = a + ac * (d2 / (d2 - d6)); break; }
		const double va = d3 * d6 - d5 * d4;
		if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) { cl = b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))); break; }
		const double den = va + vb + vc;
		if (fabs(den) > 0.0) cl = a + ab * (vb / den) + ac * (vc / den);
	} while (false);
	const D3 d = p - cl;
	return dot(d, d);
}

/// Capsule (the points within `radius` of the segment (0, 0, -+half_len) of the shape frame; fcl::Capsule's convention) vs
/// closed triangle: contact iff the segment comes within `radius` of the triangle.  The segment either pierces the triangle
/// (distance 0) or its closest approach is attained at a segment end point or on a triangle edge.
static __device__ __noinline__ bool capsule_triangle_hit(const ExactBox &b, double radius, double half_len, const double *__restrict__ tri) {
	D3 p[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const double dx = __ldg(tri + 3 * k) - b.cx, dy = __ldg(tri + 3 * k + 1) - b.cy, dz = __ldg(tri + 3 * k + 2) - b.cz;
		p[k] = {b.r[0] * dx + b.r[3] * dy + b.r[6] * dz, b.r[1] * dx + b.r[4] * dy + b.r[7] * dz, b.r[2] * dx + b.r[5] * dy + b.r[8] * dz};
	}
	const D3 s0{0.0, 0.0, -half_len}, s1{0.0, 0.0, half_len};
	const double r2 = radius * radius;
	double best = fmin(point_triangle_dist2(s0, p[0], p[1], p[2]), point_triangle_dist2(s1, p[0], p[1], p[2]));
	if (best <= r2) return true;
#pragma unroll
	for (int k = 0; k < 3; ++k) best = fmin(best, segment_segment_dist2(s0, s1, p[k], p[(k + 1) % 3]));
	if (best <= r2) return true;
	// piercing: the end points lie on opposite sides of the plane and the crossing point falls inside the triangle
	const D3 e0 = p[1] - p[0], e1 = p[2] - p[0], n = cross(e0, e1);
	const double h0 = dot(n, s0 - p[0]), h1 = dot(n, s1 - p[0]);
	if ((h0 > 0.0 && h1 > 0.0) || (h0 < 0.0 && h1 < 0.0) || h0 == h1) return false;
	const D3 x = s0 + (s1 - s0) * (h0 / (h0 - h1));
	const D3 w = x - p[0];
	const double d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(w, e0), d21 = dot(w, e1);
	const double den = d00 * d11 - d01 * d01;
	if (!(den > 0.0)) return false;  // degenerate triangle: its edges were already measured
	const double v = (d11 * d20 - d01 * d21) / den, u = (d00 * d21 - d01 * d20) / den;
	return v >= 0.0 && u >= 0.0 && v + u <= 1.0;
}

// ---- the same distance in fp32, as a pre-filter (the exact test above decides what this cannot) -----------------------------
struct V3f {
	float x, y, z;
};
__device__ __forceinline__ V3f operator+(V3f a, V3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3f operator-(V3f a, V3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3f operator*(V3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3f cross(V3f a, V3f b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ float segment_segment_dist2_f(V3f p1, V3f q1, V3f p2, V3f q2) {
	const V3f d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
	const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
	float s = 0.f, t = 0.f;
	if (a > 1e-20f && e > 1e-20f) {
		const float den = a * e - b * b;
		s = den > 1e-6f * a * e ? clamp01((b * f - c * e) / den) : 0.f;
		t = (b * s + f) / e;
		if (t < 0.f) t = 0.f, s = clamp01(-c / a);
		else if (t > 1.f) t = 1.f, s = clamp01((b - c) / a);
	} else if (a > 1e-20f) {
		s = clamp01(-c / a);
	} else if (e > 1e-20f) {
		t = clamp01(f / e);
	}
	const V3f dd = (p1 + d1 * s) - (p2 + d2 * t);
	return dot(dd, dd);  // the distance between two points ON the segments: never below the true minimum by more than rounding
}
__device__ __forceinline__ float point_triangle_dist2_f(V3f p, V3f a, V3f b, V3f c) {
	const V3f ab = b - a, ac = c - a, ap = p - a;
	const float d1 = dot(ab, ap), d2 = dot(ac, ap);
	V3f cl = a;
	do {
		if (d1 <= 0.f && d2 <= 0.f) break;
		const V3f bp = p - b;
		const float d3 = dot(ab, bp), d4 = dot(ac, bp);
		if (d3 >= 0.f && d4 <= d3) { cl = b; break; }
		const float vc = d1 * d4 - d3 * d2;
		if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { cl = a + ab * (d1 / (d1 - d3)); break; }
		const V3f cp = p - c;
		const float d5 = dot(ab, cp), d6 = dot(ac, cp);
		if (d6 >= 0.f && d5 <= d6) { cl = c; break; }
		const float vb = d5 * d2 - d1 * d6;
		if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { cl = a + ac * (d2 / (d2 - d6)); break; }
		const float va = d3 * d6 - d5 * d4;
		if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) { cl = b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))); break; }
		const float den = va + vb + vc;
		if (fabsf(den) > 0.f) cl = a + ab * (vb / den) + ac * (vc / den);
	} while (false);
	const V3f d = p - cl;
	return dot(d, d);
}

/// fp32 verdict for a capsule (axis (0, 0, -+half_len) of the frame the triangle is given in, radius `rad`) against a triangle.
/// Every candidate distance below is measured between a point of the segment and a point of the triangle, so the smallest of
/// them can only OVERestimate the true distance: "smallest < rad - slack" is a certain hit.  "Certainly separated" needs the
/// true minimum, which the candidates attain unless the segment pierces the triangle's interior — that case is told apart by
/// the plane-side test, and left to the exact test whenever it is not clear-cut.  `slack` covers the fp32 error of the inputs.
__device__ __forceinline__ int capsule_triangle_classify(float rad, float half_len, float slack, V3f p0, V3f p1, V3f p2) {
	const V3f s0{0.f, 0.f, -half_len}, s1{0.f, 0.f, half_len};
	float best = fminf(point_triangle_dist2_f(s0, p0, p1, p2), point_triangle_dist2_f(s1, p0, p1, p2));
	best = fminf(best, fminf(segment_segment_dist2_f(s0, s1, p0, p1), fminf(segment_segment_dist2_f(s0, s1, p1, p2), segment_segment_dist2_f(s0, s1, p2, p0))));
	const float d = sqrtf(best), tol = slack + 1e-5f * (d + rad);
	if (d < rad - tol) return TRI_HIT;
	if (!(d > rad + tol)) return TRI_UNKNOWN;
	// far by every candidate: separated unless the axis passes through the triangle.  Both end points strictly on one side of the
	// plane (by more than the tolerance, measured along the unit normal) rule that out.
	const V3f n = cross(p1 - p0, p2 - p0);
	const float nn = dot(n, n);
	if (!(nn > 1e-20f)) return TRI_SEPARATED;  // no area: its edges were measured
	const float inv = rsqrtf(nn), h0 = dot(n, s0 - p0) * inv, h1 = dot(n, s1 - p0) * inv;
	if ((h0 > tol && h1 > tol) || (h0 < -tol && h1 < -tol)) return TRI_SEPARATED;
	return TRI_UNKNOWN;
}

/// Convex polytope vs closed triangle, separating-axis test in the shape frame.  `data` (mgodpl_robot_create_ex lays it out):
/// [n_verts, n_faces, n_edges, -] [verts: 3 each] [faces: outward unit normal + support value, 4 each] [edge directions: 3 each].
/// Axes: the polytope's face normals, the triangle normal, edge x edge.  Touching counts as contact; numerically null axes
/// (parallel edges, degenerate triangle) carry no information and are skipped.
static __device__ __noinline__ bool convex_triangle_hit(const ExactBox &b, const double *__restrict__ data, const double *__restrict__ tri) {
	D3 p[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const double dx = __ldg(tri + 3 * k) - b.cx, dy = __ldg(tri + 3 * k + 1) - b.cy, dz = __ldg(tri + 3 * k + 2) - b.cz;
		p[k] = {b.r[0] * dx + b.r[3] * dy + b.r[6] * dz, b.r[1] * dx + b.r[4] * dy + b.r[7] * dz, b.r[2] * dx + b.r[5] * dy + b.r[8] * dz};
	}
	const int nv = (int) __ldg(data), nf = (int) __ldg(data + 1), ne = (int) __ldg(data + 2);
	const double *verts = data + 4, *faces = verts + 3 * nv, *edges = faces + 4 * nf;
	for (int f = 0; f < nf; ++f) {
		const D3 n{__ldg(faces + 4 * f), __ldg(faces + 4 * f + 1), __ldg(faces + 4 * f + 2)};
		if (fmin(dot(n, p[0]), fmin(dot(n, p[1]), dot(n, p[2]))) > __ldg(faces + 4 * f + 3)) return false;
	}
	const D3 tf[3] = {p[1] - p[0], p[2] - p[1], p[0] - p[2]};
	const double scale2 = fmax(dot(tf[0], tf[0]), fmax(dot(tf[1], tf[1]), dot(tf[2], tf[2])));
	auto separates = [&](D3 a, double floor2) {
		if (!(dot(a, a) > floor2)) return false;
		double lo = 1e300, hi = -1e300;
		for (int v = 0; v < nv; ++v) {
			const double q = a.x * __ldg(verts + 3 * v) + a.y * __ldg(verts + 3 * v + 1) + a.z * __ldg(verts + 3 * v + 2);
			lo = fmin(lo, q), hi = fmax(hi, q);
		}
		const double q0 = dot(a, p[0]), q1 = dot(a, p[1]), q2 = dot(a, p[2]);
		return fmin(q0, fmin(q1, q2)) > hi || fmax(q0, fmax(q1, q2)) < lo;
	};
	if (separates(cross(tf[0], tf[1]), 1e-28 * scale2 * scale2)) return false;
	for (int e = 0; e < ne; ++e) {
		const D3 d{__ldg(edges + 3 * e), __ldg(edges + 3 * e + 1), __ldg(edges + 3 * e + 2)};  // unit length
#pragma unroll
		for (int k = 0; k < 3; ++k)
			if (separates(cross(d, tf[k]), 1e-24 * dot(tf[k], tf[k]))) return false;
	}
	return true;
}

/// One exact shape/triangle test, by kind.  `half`: the box's half extents (capsule: radius = half[0], segment half length =
/// half[2] - half[0]; convex: its bounding box, the vertices live in `data`).
__device__ __forceinline__ bool shape_triangle_hit(int kind, const ExactBox &b, const double *data, const double *__restrict__ tri) {
	if (kind == SHAPE_KIND_BOX) return box_triangle_hit(b, tri);
	if (kind == SHAPE_KIND_CAPSULE) return capsule_triangle_hit(b, b.hx, b.hz - b.hx, tri);
	return convex_triangle_hit(b, data, tri);
}

}  // namespace mgodpl_b200
