// ORACLE — TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the FCL 0.7 types named at
// src/planning/collision_detection.cpp:25-30 (FCL is absent from this image; see oracle/mgodpl_oracle.hpp header).
#pragma once
#include <memory>
#include "../../Eigen/Geometry"
namespace fcl {
struct Vector3d {
	double v[3]{0, 0, 0};
	Vector3d() = default;
	Vector3d(double x, double y, double z) : v{x, y, z} {}
};
using Quaterniond = Eigen::Quaterniond;
/// Isometry: translation + rotation kept as a quaternion; rotate() post-multiplies like Eigen::Transform::rotate.
class Transform3d {
	Vector3d t_;
	orc::Quat q_;
public:
	void setIdentity() {
		t_ = Vector3d();
		q_ = orc::Quat{0, 0, 0, 1};
	}
	Vector3d &translation() { return t_; }
	const Vector3d &translation() const { return t_; }
	void rotate(const Quaterniond &q) { q_ = orc::qmul(q_, q.raw()); }
	orc::Tf as_tf() const { return {{t_.v[0], t_.v[1], t_.v[2]}, q_}; }
	static Transform3d Identity() {
		Transform3d t;
		t.setIdentity();
		return t;
	}
};
template<typename S> struct OBB {};
using OBBd = OBB<double>;
template<typename S>
class CollisionGeometry {
public:
	virtual ~CollisionGeometry() = default;
};
using CollisionGeometryd = CollisionGeometry<double>;
}  // namespace fcl
