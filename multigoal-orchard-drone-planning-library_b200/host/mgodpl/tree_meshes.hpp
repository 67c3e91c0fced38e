// Tree-model ingest: the data format on the input side of the scene builder.
//   from_dae                                       src/experiment_utils/mesh_from_dae.cpp:17-85
//   connected_vertex_components,
//   break_down_to_connected_components             src/experiment_utils/mesh_connected_components.cpp:25-118
//   mesh_aabb                                      src/planning/Mesh.cpp:11-17
//   TreeMeshes, loadTreeMeshes, getTreeModelNames, SimplifiedOrchard, makeSingleRowOrchard, computeFruitPositions
//                                                  src/experiment_utils/TreeMeshes.{h,cpp}
// The reference reads COLLADA through assimp (aiProcess_Triangulate | JoinIdenticalVertices | SortByPType |
// RemoveComponent with everything but positions removed) and then ignores the node hierarchy: it concatenates the raw
// vertices of every aiMesh, applying one hard-coded unit/axis conversion.  assimp is absent here, so `from_dae` is a
// small COLLADA 1.4 reader that produces the same kind of mesh: positions only, geometries in the order the visual scene's
// nodes first instantiate them (assimp's aiMesh order), polygons fan-triangulated, identical positions merged per primitive
// block in first-seen order, float32 coordinates (assimp's aiVector3D) converted exactly as
// mesh_from_dae.cpp:55-69 does.  Fruit separation runs on the GPU (mgodpl_mesh_components).
#pragma once
#include <algorithm>
#include <array>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "collision_detection.hpp"

namespace mgodpl {

namespace math {
/// src/math/AABB.h: the two members this path uses.
struct AABBd {
	Vec3d _min{INFINITY, INFINITY, INFINITY}, _max{-INFINITY, -INFINITY, -INFINITY};
	void expand(const Vec3d &p) {
		_min = {std::min(_min.x(), p.x()), std::min(_min.y(), p.y()), std::min(_min.z(), p.z())};
		_max = {std::max(_max.x(), p.x()), std::max(_max.y(), p.y()), std::max(_max.z(), p.z())};
	}
	Vec3d center() const { return (_min + _max) * 0.5; }
	Vec3d size() const { return _max - _min; }
};
}  // namespace math

/// Mesh.cpp:11-17.
inline math::AABBd mesh_aabb(const Mesh &mesh) {
	math::AABBd box;
	for (const auto &v: mesh.vertices) box.expand(v);
	return box;
}

namespace dae_detail {
/// Just enough XML for COLLADA: elements, attributes, character data; comments / declarations / CDATA are skipped.
struct Element {
	std::string name, text;
	std::map<std::string, std::string> attributes;
	std::vector<Element> children;
	const Element *child(const std::string &n) const {
		for (const auto &c: children)
			if (c.name == n) return &c;
		return nullptr;
	}
	std::string attribute(const std::string &k) const {
		auto it = attributes.find(k);
		return it == attributes.end() ? std::string() : it->second;
	}
};
class Parser {
	const std::string &s;
	size_t i = 0;
	[[noreturn]] void fail(const char *what) const { throw std::runtime_error(std::string("malformed COLLADA document: ") + what); }
	void skip_space() {
		while (i < s.size() && std::isspace((unsigned char) s[i])) ++i;
	}
	bool starts(const char *lit) const { return s.compare(i, std::strlen(lit), lit) == 0; }
	void skip_past(const char *lit) {
		size_t j = s.find(lit, i);
		if (j == std::string::npos) fail("unterminated markup");
		i = j + std::strlen(lit);
	}
	std::string name() {
		size_t b = i;
		while (i < s.size() && !std::isspace((unsigned char) s[i]) && s[i] != '>' && s[i] != '/' && s[i] != '=') ++i;
		return s.substr(b, i - b);
	}

public:
	explicit Parser(const std::string &text) : s(text) {}
	/// Parses the next element (skipping prolog / comments before it).
	Element element() {
		for (;;) {
			skip_space();
			if (starts("<?")) skip_past("?>");
			else if (starts("<!--")) skip_past("-->");
			else if (starts("<!")) skip_past(">");
			else break;
		}
		if (i >= s.size() || s[i] != '<') fail("element expected");
		++i;
		Element e;
		e.name = name();
		for (;;) {
			skip_space();
			if (i >= s.size()) fail("unterminated tag");
			if (s[i] == '/') {
				if (i + 1 >= s.size() || s[i + 1] != '>') fail("bad empty-element tag");
				i += 2;
				return e;
			}
			if (s[i] == '>') {
				++i;
				break;
			}
			std::string key = name();
			skip_space();
			if (i >= s.size() || s[i] != '=') fail("attribute without value");
			++i;
			skip_space();
			if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail("unquoted attribute");
			const char quote = s[i++];
			size_t b = i;
			while (i < s.size() && s[i] != quote) ++i;
			if (i >= s.size()) fail("unterminated attribute");
			e.attributes[key] = s.substr(b, i - b);
			++i;
		}
		for (;;) {  // content
			size_t lt = s.find('<', i);
			if (lt == std::string::npos) fail("unterminated element");
			e.text.append(s, i, lt - i);
			i = lt;
			if (starts("</")) {
				skip_past(">");
				return e;
			}
			if (starts("<!--")) skip_past("-->");
			else if (starts("<![CDATA[")) {
				size_t b = i + 9;
				skip_past("]]>");
				e.text.append(s, b, i - 3 - b);
			} else if (starts("<?")) skip_past("?>");
			else e.children.push_back(element());
		}
	}
};
template<class T>
inline std::vector<T> numbers(const std::string &text) {
	std::vector<T> out;
	const char *p = text.c_str();
	char *end = nullptr;
	for (;;) {
		while (*p && std::isspace((unsigned char) *p)) ++p;
		if (!*p) break;
		if constexpr (std::is_floating_point_v<T>) out.push_back((T) std::strtod(p, &end));
		else out.push_back((T) std::strtoull(p, &end, 10));
		if (end == p) throw std::runtime_error("malformed COLLADA document: number expected");
		p = end;
	}
	return out;
}
inline std::string strip_hash(const std::string &ref) { return !ref.empty() && ref[0] == '#' ? ref.substr(1) : ref; }
inline void collect(const Element &e, const std::string &name, std::vector<const Element *> &out) {
	if (e.name == name) out.push_back(&e);
	for (const auto &c: e.children) collect(c, name, out);
}
}  // namespace dae_detail

/// mesh_from_dae.cpp:17-85: positions and triangles of every geometry of a COLLADA file, converted to metres, z up.
/// Files whose name contains "apples", "fruit", "leaves" or "trunk" are tree meshes (inch-scaled, y/z swapped); anything
/// else is divided by 10 (mesh_from_dae.cpp:44-69).
inline Mesh from_dae(const std::string &dae_file) {
	using namespace dae_detail;
	std::ifstream in(dae_file, std::ios::binary);
	if (!in)
		throw std::runtime_error("Failed to load mesh from " + dae_file + ": Unable to open file"
								 " (Are you in the right working directory, and are the mesh files available in it?)");
	std::stringstream buffer;
	buffer << in.rdbuf();
	const std::string text = buffer.str();
	const Element root = Parser(text).element();
	if (root.name != "COLLADA") throw std::runtime_error("Failed to load mesh from " + dae_file + ": not a COLLADA document");

	const bool is_tree_mesh = dae_file.find("apples") != std::string::npos || dae_file.find("fruit") != std::string::npos ||
							  dae_file.find("leaves") != std::string::npos || dae_file.find("trunk") != std::string::npos;
	constexpr double CENTIMETERS_PER_INCH = 2.54;
	auto convert = [&](float x, float y, float z) {
		if (is_tree_mesh)  // z and y are swapped intentionally (mesh_from_dae.cpp:55-63)
			return math::Vec3d(x / (100.0 / CENTIMETERS_PER_INCH), z / (100.0 / CENTIMETERS_PER_INCH), y / (100.0 / CENTIMETERS_PER_INCH));
		return math::Vec3d(x / 10.0, y / 10.0, z / 10.0);
	};

	Mesh mesh;
	std::vector<const Element *> geometries;
	collect(root, "geometry", geometries);
	// assimp's COLLADA loader creates its aiMeshes while it walks the node hierarchy of the instantiated visual scene (depth
	// first, a node's own geometry instances before its children), one per geometry the first time a node instantiates it:
	// scene->mMeshes — which mesh_from_dae.cpp:52 concatenates, ignoring the nodes' transforms — is in THAT order, and a
	// geometry no node refers to is not in it at all.  A document without a visual scene (assimp refuses those) is read in
	// library order.
	{
		std::vector<const Element *> scenes, nodes, ordered;
		collect(root, "visual_scene", scenes);
		collect(root, "node", nodes);
		const Element *start = nullptr;
		if (const Element *sc = root.child("scene"))
			if (const Element *inst = sc->child("instance_visual_scene"))
				for (const Element *vs: scenes)
					if (vs->attribute("id") == strip_hash(inst->attribute("url"))) start = vs;
		if (!start && !scenes.empty()) start = scenes.front();
		std::vector<const Element *> active;  // guards against <instance_node> cycles
		auto walk = [&](auto &&self, const Element &node) -> void {
			if (std::find(active.begin(), active.end(), &node) != active.end()) return;
			active.push_back(&node);
			for (const auto &c: node.children)
				if (c.name == "instance_geometry")
					for (const Element *g: geometries)
						if (g->attribute("id") == strip_hash(c.attribute("url")) && std::find(ordered.begin(), ordered.end(), g) == ordered.end())
							ordered.push_back(g);
			for (const auto &c: node.children) {
				if (c.name == "node") self(self, c);
				else if (c.name == "instance_node")
					for (const Element *n: nodes)
						if (n->attribute("id") == strip_hash(c.attribute("url"))) self(self, *n);
			}
			active.pop_back();
		};
		if (start) {
			for (const auto &c: start->children)
				if (c.name == "node") walk(walk, c);
			geometries = ordered;
		}
	}
	for (const Element *geometry: geometries) {
		const Element *m = geometry->child("mesh");
		if (!m) continue;
		std::map<std::string, std::vector<float>> sources;  // source id -> xyz triples
		std::map<std::string, std::string> vertices_of;     // <vertices> id -> position source id
		for (const auto &c: m->children) {
			if (c.name == "source") {
				const Element *arr = c.child("float_array");
				if (!arr) continue;
				std::vector<float> values = numbers<float>(arr->text);
				size_t stride = 3;
				if (const Element *tc = c.child("technique_common"))
					if (const Element *acc = tc->child("accessor"))
						if (!acc->attribute("stride").empty()) stride = std::stoul(acc->attribute("stride"));
				if (stride < 3) continue;  // not a position source
				std::vector<float> xyz;
				for (size_t k = 0; k + stride <= values.size(); k += stride) xyz.insert(xyz.end(), {values[k], values[k + 1], values[k + 2]});
				sources[c.attribute("id")] = std::move(xyz);
			} else if (c.name == "vertices") {
				for (const auto &inp: c.children)
					if (inp.name == "input" && inp.attribute("semantic") == "POSITION") vertices_of[c.attribute("id")] = strip_hash(inp.attribute("source"));
			}
		}
		for (const auto &prim: m->children) {
			const bool tri = prim.name == "triangles", plist = prim.name == "polylist", pgons = prim.name == "polygons";
			if (!tri && !plist && !pgons) continue;  // lines / points: dropped, as aiProcess_SortByPType + the triangle check would
			size_t n_inputs = 0, vertex_offset = 0;
			const std::vector<float> *positions = nullptr;
			for (const auto &inp: prim.children) {
				if (inp.name != "input") continue;
				const size_t off = inp.attribute("offset").empty() ? 0 : std::stoul(inp.attribute("offset"));
				n_inputs = std::max(n_inputs, off + 1);
				if (inp.attribute("semantic") == "VERTEX") {
					vertex_offset = off;
					auto v = vertices_of.find(strip_hash(inp.attribute("source")));
					if (v != vertices_of.end() && sources.count(v->second)) positions = &sources[v->second];
				}
			}
			if (!positions || !n_inputs) continue;
			// one sub-mesh per primitive block (assimp: one aiMesh), identical positions joined in first-seen order
			const size_t index_offset = mesh.vertices.size();
			std::map<std::array<uint32_t, 3>, size_t> seen;
			auto corner = [&](size_t position_index) {
				if (3 * position_index + 2 >= positions->size()) throw std::runtime_error("Failed to load mesh from " + dae_file + ": index out of range");
				const float x = (*positions)[3 * position_index], y = (*positions)[3 * position_index + 1], z = (*positions)[3 * position_index + 2];
				std::array<uint32_t, 3> key;
				std::memcpy(&key[0], &x, 4), std::memcpy(&key[1], &y, 4), std::memcpy(&key[2], &z, 4);
				auto [it, fresh] = seen.try_emplace(key, mesh.vertices.size());
				if (fresh) mesh.vertices.push_back(convert(x, y, z));
				return it->second;
			};
			auto polygon = [&](const size_t *p, size_t n_corners) {  // fan triangulation (aiProcess_Triangulate on convex input)
				for (size_t k = 1; k + 1 < n_corners; ++k) {
					const size_t a = corner(p[vertex_offset]), b = corner(p[k * n_inputs + vertex_offset]), c = corner(p[(k + 1) * n_inputs + vertex_offset]);
					mesh.triangles.push_back({a, b, c});
				}
			};
			(void) index_offset;
			if (pgons) {
				for (const auto &p: prim.children)
					if (p.name == "p") {
						const std::vector<size_t> idx = numbers<size_t>(p.text);
						polygon(idx.data(), idx.size() / n_inputs);
					}
				continue;
			}
			const Element *p = prim.child("p");
			if (!p) continue;
			const std::vector<size_t> idx = numbers<size_t>(p->text);
			if (tri) {
				for (size_t k = 0; k + 3 * n_inputs <= idx.size(); k += 3 * n_inputs) polygon(idx.data() + k, 3);
			} else {
				const Element *vc = prim.child("vcount");
				if (!vc) continue;
				size_t k = 0;
				for (size_t n_corners: numbers<size_t>(vc->text)) {
					if (k + n_corners * n_inputs > idx.size()) throw std::runtime_error("Failed to load mesh from " + dae_file + ": truncated polylist");
					polygon(idx.data() + k, n_corners);
					k += n_corners * n_inputs;
				}
			}
		}
	}
	return mesh;
}

/// mesh_connected_components.cpp:25-74, on the GPU.  Components are ordered by their smallest vertex index and list
/// their vertices in ascending order (the reference returns the same partition in unordered_map iteration order).
inline std::vector<std::vector<size_t>> connected_vertex_components(const Mesh &mesh) {
	static_assert(sizeof(size_t) == sizeof(uint64_t), "Mesh::triangles is passed through as uint64");
	std::vector<uint32_t> labels(mesh.vertices.size());
	gpu::throw_status(mgodpl_mesh_components_u64(reinterpret_cast<const uint64_t *>(mesh.triangles.data()), mesh.triangles.size(),
												 mesh.vertices.size(), labels.data(), nullptr));
	std::vector<std::vector<size_t>> components;
	std::vector<size_t> slot(mesh.vertices.size(), SIZE_MAX);
	for (size_t v = 0; v < labels.size(); ++v) {
		size_t &s = slot[labels[v]];
		if (s == SIZE_MAX) s = components.size(), components.emplace_back();  // labels[v] <= v, so roots come first
		components[s].push_back(v);
	}
	return components;
}

/// mesh_connected_components.cpp:77-118.
inline std::vector<Mesh> break_down_to_connected_components(const Mesh &combined_mesh) {
	const auto cc = connected_vertex_components(combined_mesh);
	std::vector<Mesh> parts(cc.size());
	std::vector<std::pair<size_t, size_t>> index_map(combined_mesh.vertices.size());
	for (size_t i = 0; i < cc.size(); ++i)
		for (size_t j = 0; j < cc[i].size(); ++j) {
			index_map[cc[i][j]] = {i, j};
			parts[i].vertices.push_back(combined_mesh.vertices[cc[i][j]]);
		}
	for (const auto &t: combined_mesh.triangles) {
		const auto &a = index_map[t[0]], &b = index_map[t[1]], &c = index_map[t[2]];
		assert(a.first == b.first && b.first == c.first);
		parts[a.first].triangles.push_back({a.second, b.second, c.second});
	}
	return parts;
}

namespace tree_meshes {
using TreeMeshes = gpu::TreeMeshes;  // TreeMeshes.h:19-24

/// TreeMeshes.cpp:28-66.  `directory` defaults to the reference's "<cwd>/3d-models/".
inline TreeMeshes loadTreeMeshes(const std::string &treeName, const std::string &directory = "3d-models/") {
	TreeMeshes meshes;
	meshes.tree_name = treeName;
	meshes.leaves_mesh = from_dae(directory + treeName + "_leaves.dae");
	meshes.trunk_mesh = from_dae(directory + treeName + "_trunk.dae");
	// for legacy reasons some models call the file "fruit", others "apples"
	const std::string fruit = directory + treeName + "_fruit.dae";
	meshes.fruit_meshes = break_down_to_connected_components(from_dae(std::filesystem::exists(fruit) ? fruit : directory + treeName + "_apples.dae"));
	const size_t n_before = meshes.fruit_meshes.size();
	// tiny sliver components are the stems of the fruit: ignored (<= 20 vertices)
	meshes.fruit_meshes.erase(std::remove_if(meshes.fruit_meshes.begin(), meshes.fruit_meshes.end(), [](const Mesh &m) { return m.vertices.size() <= 20; }),
							  meshes.fruit_meshes.end());
	if (n_before != meshes.fruit_meshes.size())
		std::cout << "Removed " << n_before - meshes.fruit_meshes.size() << " out of " << n_before << " fruit meshes that were too small from model "
				  << treeName << std::endl;
	return meshes;
}

/// TreeMeshes.cpp:68-88: names of all "<name>_trunk.dae" models in a directory.
inline std::vector<std::string> getTreeModelNames(const std::string &directory = "./3d-models/") {
	std::vector<std::string> names;
	for (const auto &entry: std::filesystem::directory_iterator(directory)) {
		const std::string stem = entry.path().stem().string();
		if (entry.path().extension() == ".dae" && stem.find("_trunk") != std::string::npos) names.push_back(stem.substr(0, stem.find("_trunk")));
	}
	return names;
}

/// TreeMeshes.h:57-59, TreeMeshes.cpp:197-206: trees two metres apart along x, centred on the origin.
struct SimplifiedOrchard {
	std::vector<std::pair<math::Vec3d, TreeMeshes>> trees;
};
inline SimplifiedOrchard makeSingleRowOrchard(std::vector<TreeMeshes> &tree_models) {
	SimplifiedOrchard orchard;
	const double first_x = -1.0 * (double) tree_models.size();  // n trees, 2 m apart, starting n metres left of the origin
	for (size_t i = 0; i < tree_models.size(); ++i) orchard.trees.push_back({{first_x + 2.0 * (double) i, 0.0, 0.0}, tree_models[i]});
	return orchard;
}

/// TreeMeshes.cpp:208-214.
inline std::vector<math::Vec3d> computeFruitPositions(const TreeMeshes &tree_meshes) {
	std::vector<math::Vec3d> positions;
	for (const auto &fruit: tree_meshes.fruit_meshes) positions.push_back(mesh_aabb(fruit).center());
	return positions;
}
}  // namespace tree_meshes

namespace gpu {
/// One collision scene for a whole orchard: every tree's trunk mesh moved to its root position and merged (the
/// reference checks each tree separately; one BVH answers the same union).
inline Scene orchardToScene(const tree_meshes::SimplifiedOrchard &orchard) {
	Mesh merged;
	for (const auto &[root, tree]: orchard.trees) {
		const size_t base = merged.vertices.size();
		for (const auto &v: tree.trunk_mesh.vertices) merged.vertices.push_back(v + root);
		for (const auto &t: tree.trunk_mesh.triangles) merged.triangles.push_back({t[0] + base, t[1] + base, t[2] + base});
	}
	return Scene(merged);
}
}  // namespace gpu

}  // namespace mgodpl
