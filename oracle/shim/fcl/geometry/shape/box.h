// ORACLE — TEST INFRASTRUCTURE ONLY.  fcl::Boxd(x, y, z): centred box, full extents (collision_detection.cpp:32-34).
#pragma once
#include "../../common/types.h"
namespace fcl {
template<typename S>
class Box : public CollisionGeometry<S> {
public:
	S side[3];
	Box(S x, S y, S z) : side{x, y, z} {}
};
using Boxd = Box<double>;
}  // namespace fcl
