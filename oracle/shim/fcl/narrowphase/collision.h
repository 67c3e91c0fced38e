// ORACLE — TEST INFRASTRUCTURE ONLY.  fcl::collide(mesh object, box object) → boolean, answered by the oracle's
// exact box/triangle SAT over its own AABB tree (FCL + libccd are absent; semantics per SURVEY §8c).
#pragma once
#include <stdexcept>
#include "../geometry/bvh/BVH_model.h"
#include "../geometry/shape/box.h"
#include "collision_object.h"
#include "collision_request.h"
namespace fcl {
inline std::size_t collide(const CollisionObjectd *o1, const CollisionObjectd *o2, const CollisionRequestd &,
						   CollisionResultd &result) {
	auto *mesh = dynamic_cast<const BVHModel<OBBd> *>(o1->collisionGeometry().get());
	auto *box = dynamic_cast<const Boxd *>(o2->collisionGeometry().get());
	if (!mesh || !box) throw std::runtime_error("fcl shim: only (BVHModel<OBBd>, Boxd) pairs are supported");
	orc::Obb obb(o2->getTransform().as_tf(), {box->side[0], box->side[1], box->side[2]});
	result.collision = orc::obb_hits_scene(*mesh->scene, obb);
	return result.collision ? 1 : 0;
}
}  // namespace fcl
