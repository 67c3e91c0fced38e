// ORACLE — TEST INFRASTRUCTURE ONLY.  Default-constructed request/result pair (collision_detection.cpp:36-45).
#pragma once
namespace fcl {
struct CollisionRequestd {};
struct CollisionResultd {
	bool collision = false;
	bool isCollision() const { return collision; }
};
}  // namespace fcl
