// Host-side mirror of the reference's math layer for the collision-query path (header-only, C++17).
// Same names and semantics as src/math/Vec3.h:22-24, Quaternion.h:25-85, Transform.h:20-78 of the reference; written
// from their documented behaviour, not copied.  Quaternions are stored x, y, z, w.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <limits>

namespace mgodpl::math {

struct Vec3d {
	std::array<double, 3> components{0.0, 0.0, 0.0};
	Vec3d() = default;
	Vec3d(double x, double y, double z) : components{x, y, z} {}
	double x() const { return components[0]; }
	double y() const { return components[1]; }
	double z() const { return components[2]; }
	double getX() const { return components[0]; }
	double getY() const { return components[1]; }
	double getZ() const { return components[2]; }
	double &operator[](size_t i) { return components[i]; }
	double operator[](size_t i) const { return components[i]; }
	Vec3d operator+(const Vec3d &o) const { return {x() + o.x(), y() + o.y(), z() + o.z()}; }
	Vec3d operator-(const Vec3d &o) const { return {x() - o.x(), y() - o.y(), z() - o.z()}; }
	Vec3d operator-() const { return {-x(), -y(), -z()}; }
	Vec3d operator*(double s) const { return {x() * s, y() * s, z() * s}; }
	Vec3d &operator+=(const Vec3d &o) { return *this = *this + o; }
	double dot(const Vec3d &o) const { return x() * o.x() + y() * o.y() + z() * o.z(); }
	double squaredNorm() const { return dot(*this); }
	double norm() const { return std::sqrt(squaredNorm()); }
	Vec3d normalized() const {
		double n = norm();
		return {x() / n, y() / n, z() / n};
	}
	bool operator==(const Vec3d &o) const { return components == o.components; }
	static Vec3d UnitX() { return {1, 0, 0}; }
	static Vec3d UnitY() { return {0, 1, 0}; }
	static Vec3d UnitZ() { return {0, 0, 1}; }
};

struct Quaterniond {
	double x = 0, y = 0, z = 0, w = 1;
	Quaterniond conjugate() const { return {-x, -y, -z, w}; }
	Quaterniond inverse() const { return conjugate(); }  // unit quaternions only, never renormalised
	Quaterniond operator*(const Quaterniond &o) const {
		return {w * o.x + x * o.w + y * o.z - z * o.y, w * o.y - x * o.z + y * o.w + z * o.x,
				w * o.z + x * o.y - y * o.x + z * o.w, w * o.w - x * o.x - y * o.y - z * o.z};
	}
	Vec3d rotate(const Vec3d &v) const {
		Quaterniond r = (*this * Quaterniond{v.x(), v.y(), v.z(), 0.0}) * conjugate();
		return {r.x, r.y, r.z};
	}
	static Quaterniond fromAxisAngle(const Vec3d &axis, double angle) {
		double h = angle / 2.0, s = std::sin(h);
		return {axis.x() * s, axis.y() * s, axis.z() * s, std::cos(h)};
	}
	bool operator==(const Quaterniond &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
};

inline double angular_distance(const Quaterniond &a, const Quaterniond &b) {
	double c = std::clamp(a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z, -1.0, 1.0);
	return std::acos(2.0 * c * c - 1.0);
}

/// The reference delegates to Eigen::Quaterniond::slerp (Quaternion.cpp:12-20): shortest arc, linear weights when
/// |dot| >= 1 - eps, no renormalisation.
inline Quaterniond slerp(const Quaterniond &a, const Quaterniond &b, double t) {
	const double one = 1.0 - std::numeric_limits<double>::epsilon();
	double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w, ad = std::fabs(d), s0, s1;
	if (ad >= one) {
		s0 = 1.0 - t, s1 = t;
	} else {
		double th = std::acos(ad), st = std::sin(th);
		s0 = std::sin((1.0 - t) * th) / st, s1 = std::sin(t * th) / st;
	}
	if (d < 0.0) s1 = -s1;
	return {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

struct Transformd {
	Vec3d translation;
	Quaterniond orientation;
	static Transformd identity() { return {{0, 0, 0}, {0, 0, 0, 1}}; }
	static Transformd fromTranslation(const Vec3d &t) { return {t, {0, 0, 0, 1}}; }
	static Transformd fromRotation(const Quaterniond &q) { return {{0, 0, 0}, q}; }
	Transformd then(const Transformd &o) const { return {translation + orientation.rotate(o.translation), orientation * o.orientation}; }
	Transformd inverse() const {
		Quaterniond qi = orientation.inverse();
		return {qi.rotate(-translation), qi};
	}
	Vec3d apply(const Vec3d &v) const { return orientation.rotate(v) + translation; }
	bool operator==(const Transformd &o) const { return translation == o.translation && orientation == o.orientation; }
};

inline Transformd interpolate(const Transformd &a, const Transformd &b, double t) {
	return {{(1.0 - t) * a.translation.x() + t * b.translation.x(), (1.0 - t) * a.translation.y() + t * b.translation.y(),
			 (1.0 - t) * a.translation.z() + t * b.translation.z()},
			slerp(a.orientation, b.orientation, t)};
}

}  // namespace mgodpl::math
