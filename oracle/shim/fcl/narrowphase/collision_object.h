// ORACLE — TEST INFRASTRUCTURE ONLY.  fcl::CollisionObjectd: geometry + pose (collision_detection.cpp:32-34,
// fcl_utils.cpp:58-60).
#pragma once
#include "../common/types.h"
namespace fcl {
template<typename S>
class CollisionObject {
	std::shared_ptr<CollisionGeometry<S>> geom_;
	Transform3d tf_;
public:
	explicit CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g) : geom_(g) { tf_.setIdentity(); }
	CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Transform3d &tf) : geom_(g), tf_(tf) {}
	const std::shared_ptr<CollisionGeometry<S>> &collisionGeometry() const { return geom_; }
	const Transform3d &getTransform() const { return tf_; }
};
using CollisionObjectd = CollisionObject<double>;
}  // namespace fcl
