// Host-side mirror of the reference's robot layer for the collision-query path (header-only, C++17):
//   Mesh / Box / Shape / PositionedShape   src/planning/Mesh.h:19-22, src/experiment_utils/shapes.h:20-29, positioned_shape.h:12-30
//   robot_model::RobotModel, forwardKinematics  src/planning/RobotModel.h:36-237, RobotModel.cpp:17-161
//   RobotState, interpolate, equal_weights_distance  src/planning/RobotState.h:16-23, RobotState.cpp:14-24, distance.cpp:17-25
//   RobotPath, PathPoint  src/planning/RobotPath.h
//   experiments::createProceduralRobotModel  src/experiment_utils/procedural_robot_models.cpp:86-169
// Host FK is kept because callers (goal samplers) need it; the batched, GPU-side FK lives behind the C ABI.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../../include/mgodpl_b200.h"
#include "math.hpp"

namespace mgodpl {

struct Mesh {
	std::vector<math::Vec3d> vertices;
	std::vector<std::array<size_t, 3>> triangles;
};
struct Box {
	math::Vec3d size;
};
/// Additions of this library to the reference's Shape variant (Box | Mesh, shapes.h:24-29): link shapes the CUDA path collides
/// beside boxes.  Capsule: the points within `radius` of a segment of `length` along the shape's z axis, centred (fcl::Capsule's
/// convention).  ConvexMesh: the convex hull of 4 .. MGODPL_MAX_CONVEX_VERTICES points.  A general Mesh link still throws the
/// reference's "Only boxes are implemented for collision geometry.".
struct Capsule {
	double radius, length;
};
struct ConvexMesh {
	std::vector<math::Vec3d> vertices;
};
using Shape = std::variant<Box, Mesh, Capsule, ConvexMesh>;
struct PositionedShape {
	Shape shape;
	math::Transformd transform;
	static PositionedShape untransformed(const Shape &s) { return {s, math::Transformd::identity()}; }
};

struct RobotState {
	math::Transformd base_tf;
	std::vector<double> joint_values;
	bool operator==(const RobotState &o) const { return base_tf == o.base_tf && joint_values == o.joint_values; }
};
inline RobotState interpolate(const RobotState &a, const RobotState &b, double t) {
	RobotState r{math::interpolate(a.base_tf, b.base_tf, t), {}};
	r.joint_values.reserve(a.joint_values.size());
	for (size_t i = 0; i < a.joint_values.size(); ++i) r.joint_values.push_back(a.joint_values[i] * (1 - t) + b.joint_values[i] * t);
	return r;
}
inline double equal_weights_distance(const RobotState &a, const RobotState &b) {
	double d = (a.base_tf.translation - b.base_tf.translation).norm();
	d += math::angular_distance(a.base_tf.orientation, b.base_tf.orientation);
	for (size_t i = 0; i < a.joint_values.size(); ++i) d += std::abs(a.joint_values[i] - b.joint_values[i]);
	assert(std::isfinite(d));
	return d;
}

struct RobotPath {
	std::vector<RobotState> states;
	void append(const RobotState &s) { states.push_back(s); }
	void append(const RobotPath &p) { states.insert(states.end(), p.states.begin(), p.states.end()); }
	static RobotPath singleton(RobotState s) { return {{std::move(s)}}; }
	size_t n_waypoints() const { return states.size(); }
};
struct PathPoint {  // RobotPath.h:116-170
	size_t segment_i = 0;
	double segment_t = 0.0;
	static PathPoint fromScalar(double d, const RobotPath &path) {
		PathPoint p;
		p.segment_i = std::min((size_t) std::floor(d), path.states.size() - 2);
		p.segment_t = d - (double) p.segment_i;
		return p;
	}
	double toScalar() const { return (double) segment_i + segment_t; }
	PathPoint adjustByScalar(double d, const RobotPath &path) const { return fromScalar(toScalar() + d, path); }
	bool operator<(const PathPoint &o) const { return toScalar() < o.toScalar(); }
	bool operator>(const PathPoint &o) const { return toScalar() > o.toScalar(); }
};
/// RobotPath.cpp:13-23.
inline RobotState interpolate(const PathPoint &path_point, const RobotPath &robot_path) {
	assert(path_point.segment_i + 1 < robot_path.states.size() && "Segment index is out of bounds");
	return interpolate(robot_path.states[path_point.segment_i], robot_path.states[path_point.segment_i + 1], path_point.segment_t);
}

namespace gpu {
/// Owner of a mgodpl_robot handle; built lazily from a RobotModel and shared by its copies.
struct DeviceRobot {
	mgodpl_robot *handle = nullptr;
	~DeviceRobot() { mgodpl_robot_destroy(handle); }
};
inline void throw_status(int rc) {
	if (rc != MGODPL_OK) throw std::runtime_error(mgodpl_last_error());
}
inline void to7(const math::Transformd &t, double *o) {
	o[0] = t.translation.x(), o[1] = t.translation.y(), o[2] = t.translation.z();
	o[3] = t.orientation.x, o[4] = t.orientation.y, o[5] = t.orientation.z, o[6] = t.orientation.w;
}
}  // namespace gpu

namespace robot_model {

class RobotModel {
public:
	using LinkId = size_t;
	using JointId = size_t;
	struct RevoluteJoint {
		math::Vec3d axis;
		double min_angle, max_angle;
	};
	struct FixedJoint {};
	using JointTypeSpecific = std::variant<RevoluteJoint, FixedJoint>;
	struct Joint {
		std::string name;
		math::Transformd attachmentA, attachmentB;
		LinkId linkA, linkB;
		JointTypeSpecific type_specific;
	};
	struct Link {
		std::string name;
		std::vector<JointId> joints{};
		std::vector<PositionedShape> collision_geometry{};
		std::vector<PositionedShape> visual_geometry{};
		static Link namedBox(const std::string &name, const math::Vec3d &size) {
			return Link{name, {}, {PositionedShape::untransformed(Box{size})}, {}};
		}
	};

	static math::Transformd variableTransform(const JointTypeSpecific &v, const double *joint_values) {
		if (auto *r = std::get_if<RevoluteJoint>(&v))
			return math::Transformd::fromRotation(math::Quaterniond::fromAxisAngle(r->axis, joint_values[0]));
		return math::Transformd::identity();
	}
	static size_t n_variables(const JointTypeSpecific &v) { return v.index() == 0 ? 1 : 0; }

	LinkId insertLink(const Link &l) {
		links.push_back(l);
		device.reset();
		return links.size() - 1;
	}
	JointId insertJoint(const Joint &j) {
		joints.push_back(j);
		links[j.linkA].joints.push_back(joints.size() - 1);
		links[j.linkB].joints.push_back(joints.size() - 1);
		device.reset();
		return joints.size() - 1;
	}
	LinkId findLinkByName(const std::string &name) const {
		for (size_t i = 0; i < links.size(); ++i)
			if (links[i].name == name) return i;
		throw std::runtime_error("Link not found: " + name);
	}
	JointId findJointByName(const std::string &name) const {
		for (size_t i = 0; i < joints.size(); ++i)
			if (joints[i].name == name) return i;
		throw std::runtime_error("Joint not found: " + name);
	}
	size_t count_joint_variables() const {
		size_t n = 0;
		for (auto &j: joints) n += n_variables(j.type_specific);
		return n;
	}
	const std::vector<Link> &getLinks() const { return links; }
	const std::vector<Joint> &getJoints() const { return joints; }

	/// Device copy of this model for the CUDA path (built on first use; thread-safe).  Throws the reference's
	/// messages: "Only boxes are implemented for collision geometry.", "Link not found: flying_base".
	mgodpl_robot *device_handle() const {
		std::lock_guard<std::mutex> g(*device_mu);
		if (!device) {
			std::vector<mgodpl_shape_desc> sd;
			std::vector<double> mesh_vertices;
			std::vector<int32_t> mesh_offsets{0};
			for (size_t l = 0; l < links.size(); ++l)
				for (auto &g2: links[l].collision_geometry) {
					mgodpl_shape_desc d{};
					d.link = (int32_t) l;
					if (auto *b = std::get_if<Box>(&g2.shape)) {
						d.shape_type = MGODPL_SHAPE_BOX;
						d.size[0] = b->size.x(), d.size[1] = b->size.y(), d.size[2] = b->size.z();
					} else if (auto *c = std::get_if<Capsule>(&g2.shape)) {
						d.shape_type = MGODPL_SHAPE_CAPSULE;
						d.size[0] = d.size[1] = c->radius, d.size[2] = c->length;
					} else if (auto *m = std::get_if<ConvexMesh>(&g2.shape)) {
						d.shape_type = MGODPL_SHAPE_CONVEX_MESH;
						for (auto &v: m->vertices) mesh_vertices.insert(mesh_vertices.end(), {v.x(), v.y(), v.z()});
					} else {
						d.shape_type = MGODPL_SHAPE_MESH;
					}
					mesh_offsets.push_back((int32_t) (mesh_vertices.size() / 3));
					gpu::to7(g2.transform, d.transform);
					sd.push_back(d);
				}
			std::vector<mgodpl_joint_desc> jd;
			for (auto &j: joints) {
				mgodpl_joint_desc d{};
				d.link_a = (int32_t) j.linkA, d.link_b = (int32_t) j.linkB;
				gpu::to7(j.attachmentA, d.attachment_a), gpu::to7(j.attachmentB, d.attachment_b);
				if (auto *r = std::get_if<RevoluteJoint>(&j.type_specific)) {
					d.joint_type = MGODPL_JOINT_REVOLUTE;
					d.axis[0] = r->axis.x(), d.axis[1] = r->axis.y(), d.axis[2] = r->axis.z();
					d.min_angle = r->min_angle, d.max_angle = r->max_angle;
				} else {
					d.joint_type = MGODPL_JOINT_FIXED;
				}
				jd.push_back(d);
			}
			int32_t root = (int32_t) findLinkByName("flying_base");  // collision_detection.cpp:65
			int32_t ee = -1;
			for (size_t i = 0; i < links.size(); ++i)
				if (links[i].name == "end_effector") ee = (int32_t) i;
			auto dr = std::make_shared<gpu::DeviceRobot>();
			gpu::throw_status(mgodpl_robot_create_ex((int32_t) links.size(), sd.data(), (int32_t) sd.size(), jd.data(), (int32_t) jd.size(),
													  root, ee, mesh_vertices.data(), mesh_offsets.data(), &dr->handle));
			device = dr;
		}
		return device->handle;
	}

private:
	std::vector<Link> links;
	std::vector<Joint> joints;
	mutable std::shared_ptr<gpu::DeviceRobot> device;
	std::shared_ptr<std::mutex> device_mu = std::make_shared<std::mutex>();
};

struct ForwardKinematicsResult {
	std::vector<math::Transformd> link_transforms;
	const math::Transformd &forLink(RobotModel::LinkId l) const { return link_transforms[l]; }
};

/// Host FK with the reference's traversal semantics (explicit-stack DFS; the variable cursor advances when a joint
/// is expanded; only visited[linkB] is consulted when skipping).
inline ForwardKinematicsResult forwardKinematics(const RobotModel &model, const std::vector<double> &joint_values,
												 const RobotModel::LinkId &root_link,
												 const math::Transformd &root_tf = math::Transformd::identity()) {
	const auto &links = model.getLinks();
	const auto &joints = model.getJoints();
	ForwardKinematicsResult out{std::vector<math::Transformd>(links.size(), root_tf)};
	std::vector<char> seen(links.size(), 0);
	struct Pending {
		RobotModel::LinkId link;
		math::Transformd tf;
	};
	std::vector<Pending> todo{{root_link, root_tf}};
	size_t cursor = 0;
	while (!todo.empty()) {
		Pending cur = todo.back();
		todo.pop_back();
		if (seen[cur.link]) continue;
		seen[cur.link] = 1;
		out.link_transforms[cur.link] = cur.tf;
		for (auto ji: links[cur.link].joints) {
			const auto &j = joints[ji];
			if (seen[j.linkB]) continue;
			const bool fromA = j.linkA == cur.link;
			math::Transformd var = RobotModel::variableTransform(j.type_specific, joint_values.data() + cursor);
			if (j.linkB == cur.link) var = var.inverse();
			cursor += RobotModel::n_variables(j.type_specific);
			const math::Transformd next = cur.tf.then(fromA ? j.attachmentA : j.attachmentB).then(var)
												  .then((fromA ? j.attachmentB : j.attachmentA).inverse());
			todo.push_back({fromA ? j.linkB : j.linkA, next});
		}
	}
	return out;
}
inline ForwardKinematicsResult forwardKinematics(const RobotModel &model, const RobotState &state, const RobotModel::LinkId &root_link = 0) {
	return forwardKinematics(model, state.joint_values, root_link, state.base_tf);
}

}  // namespace robot_model

namespace experiments {
static const double ARM_LENGTH = 0.75;
enum JointType { HORIZONTAL, VERTICAL };
struct RobotArmParameters {
	double total_arm_length;
	std::vector<JointType> joint_types;
	bool add_spherical_wrist;
	double arm_segment_length() const { return total_arm_length / (double) joint_types.size(); }
};
/// The procedural drone: 0.8 x 0.8 x 0.2 base, one 0.05 x l x 0.05 stick per joint, revolute about x (HORIZONTAL) or
/// z (VERTICAL) within +-pi/2, fixed end effector at the tip.  (The visual drone mesh of the reference is not needed
/// on this path.)
inline robot_model::RobotModel createProceduralRobotModel(const RobotArmParameters &p) {
	using robot_model::RobotModel;
	const double stick = p.arm_segment_length();
	RobotModel m;
	RobotModel::LinkId attach_to = m.insertLink(RobotModel::Link::namedBox("flying_base", {0.8, 0.8, 0.2}));
	math::Transformd attachA = math::Transformd::fromTranslation({0, 0.2, 0});
	int counter = 1;
	for (JointType jt: p.joint_types) {
		const math::Transformd stick_tf = math::Transformd::fromTranslation({0, -stick / 2, 0});
		RobotModel::LinkId link = m.insertLink(RobotModel::Link::namedBox("stick", {0.05, stick, 0.05}));
		m.insertJoint({"stick_joint_" + std::to_string(counter++), attachA, stick_tf, attach_to, link,
					   RobotModel::RevoluteJoint{jt == HORIZONTAL ? math::Vec3d{1, 0, 0} : math::Vec3d{0, 0, 1}, -M_PI / 2.0, M_PI / 2.0}});
		attachA = stick_tf.inverse();
		attach_to = link;
	}
	RobotModel::LinkId ee = m.insertLink({"end_effector", {}, {}, {}});
	m.insertJoint({"stick_end_attachment", attachA, math::Transformd::identity(), attach_to, ee, RobotModel::FixedJoint{}});
	return m;
}
inline robot_model::RobotModel createProceduralRobotModel() { return createProceduralRobotModel({ARM_LENGTH, {HORIZONTAL}, false}); }
inline robot_model::RobotModel createStraightArmRobotModel(double length) { return createProceduralRobotModel({length, {HORIZONTAL}, false}); }
}  // namespace experiments

}  // namespace mgodpl
