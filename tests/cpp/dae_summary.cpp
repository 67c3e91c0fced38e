// Prints what mgodpl::from_dae (host code, no device involved) makes of a COLLADA file, as one JSON object: the counts, the
// bounding box, and order-sensitive checksums of the vertex and triangle arrays.  tests/test_ingest.py compares it with an
// independent reading of the same file that follows assimp's documented behaviour (mesh_from_dae.cpp:17-85).
#include <cinttypes>
#include <cstdio>
#include <string>

#include "../../multigoal-orchard-drone-planning-library_b200/host/mgodpl/tree_meshes.hpp"

int main(int argc, char **argv) {
	if (argc != 2) {
		std::fprintf(stderr, "usage: dae_summary file.dae\n");
		return 2;
	}
	try {
		const mgodpl::Mesh mesh = mgodpl::from_dae(argv[1]);
		auto box = mgodpl::mesh_aabb(mesh);
		if (mesh.vertices.empty()) box._min = box._max = {0, 0, 0};  // (an empty mesh has infinite bounds: not JSON)
		// position-weighted sums: sensitive to the order of vertices and triangles, exactly reproducible in numpy (fp64, sequential)
		double vsum = 0.0;
		for (size_t i = 0; i < mesh.vertices.size(); ++i) {
			const auto &v = mesh.vertices[i];
			vsum += (double) (i % 97 + 1) * (v.x() + 2.0 * v.y() + 3.0 * v.z());
		}
		uint64_t tsum = 0;
		for (size_t t = 0; t < mesh.triangles.size(); ++t)
			tsum += (uint64_t) (t % 89 + 1) * (mesh.triangles[t][0] + 3 * mesh.triangles[t][1] + 7 * mesh.triangles[t][2]);
		std::printf("{\"n_vertices\": %zu, \"n_triangles\": %zu, \"lo\": [%.17g, %.17g, %.17g], \"hi\": [%.17g, %.17g, %.17g], "
					"\"vertex_checksum\": %.17g, \"triangle_checksum\": %" PRIu64 "}\n",
					mesh.vertices.size(), mesh.triangles.size(), box._min.x(), box._min.y(), box._min.z(), box._max.x(), box._max.y(), box._max.z(),
					vsum, tsum);
	} catch (const std::exception &e) {
		std::string what = e.what();
		for (char &c: what)
			if (c == '"' || c == '\\' || (unsigned char) c < 32) c = '\'';
		std::printf("{\"error\": \"%s\"}\n", what.c_str());
		return 1;
	}
	return 0;
}
