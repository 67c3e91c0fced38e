// ORACLE — TEST INFRASTRUCTURE ONLY.  fcl::BVHModel<BV> as driven by src/planning/fcl_utils.cpp:25-55
// (beginModel / addSubModel / endModel); the tree itself is the oracle's AABB tree, not FCL's OBB tree.
#pragma once
#include <vector>
#include "../../common/types.h"
#include "../../math/triangle.h"
namespace fcl {
template<typename BV>
class BVHModel : public CollisionGeometry<double> {
	std::vector<double> verts_;
	std::vector<uint32_t> tris_;
public:
	std::shared_ptr<orc::Scene> scene;
	int beginModel() {
		verts_.clear();
		tris_.clear();
		return 0;
	}
	int addSubModel(const std::vector<Vector3d> &points, const std::vector<Triangle> &tris) {
		uint32_t base = (uint32_t) (verts_.size() / 3);
		for (auto &p: points) verts_.insert(verts_.end(), p.v, p.v + 3);
		for (auto &t: tris)
			for (int k = 0; k < 3; ++k) tris_.push_back(base + (uint32_t) t.i[k]);
		return 0;
	}
	int endModel() {
		scene = std::make_shared<orc::Scene>(verts_.data(), verts_.size() / 3, tris_.data(), tris_.size() / 3);
		return 0;
	}
};
}  // namespace fcl
