// Shortest paths on the probabilistic roadmap — what plan_path_tsp_over_prm does right after edge validation:
//   PRMGraph                                    src/planning/tsp_over_prm.h:25-26 (boost::adjacency_list, undirected, RobotState
//                                               vertices, fp64 edge weights)
//   runDijkstra, retrace_path                   src/planning/tsp_over_prm.cpp:80-128
//   GoalToGoalPathResults                       src/planning/tsp_over_prm.h:87-97
//   calculate_goal_to_goal_paths                src/planning/tsp_over_prm.cpp:497-517
// Boost is not needed: the graph is a vertex array plus an edge list, and every source is solved in one GPU call.
#pragma once
#include <limits>

#include "collision_detection.hpp"

namespace mgodpl {

/// The roadmap: states at the vertices, undirected weighted edges in insertion order.
struct PRMGraph {
	using vertex_descriptor = size_t;
	std::vector<RobotState> states;
	std::vector<uint32_t> edges;  // (u, v) pairs
	std::vector<double> weights;
	vertex_descriptor add_vertex(const RobotState &s) {
		states.push_back(s);
		return states.size() - 1;
	}
	void add_edge(vertex_descriptor u, vertex_descriptor v, double weight) {
		edges.push_back((uint32_t) u), edges.push_back((uint32_t) v), weights.push_back(weight);
	}
	size_t num_vertices() const { return states.size(); }
	size_t num_edges() const { return weights.size(); }
	const RobotState &operator[](vertex_descriptor v) const { return states[v]; }
};

/// A roadmap from validated PRM edges (validate_prm_edges): vertex i = nodes[i].
inline PRMGraph roadmap_from_edges(const std::vector<RobotState> &nodes, const std::vector<RoadmapEdge> &edges) {
	PRMGraph g;
	g.states = nodes;
	for (const auto &e: edges) g.add_edge(e.from, e.to, e.weight);
	return g;
}

struct ShortestPathTables {
	std::vector<std::vector<double>> distance;                          // [source][vertex]; numeric_limits<double>::max() if unreachable
	std::vector<std::vector<PRMGraph::vertex_descriptor>> predecessor;  // [source][vertex]; the vertex itself for sources / unreachable
};
/// runDijkstra from every vertex of `sources`, in one GPU call.
inline ShortestPathTables shortest_paths_from(const PRMGraph &graph, const std::vector<PRMGraph::vertex_descriptor> &sources, void *stream = nullptr) {
	const size_t n = graph.num_vertices(), g = sources.size();
	std::vector<uint32_t> src(sources.begin(), sources.end()), pred(n * g);
	std::vector<double> dist(n * g);
	gpu::throw_status(mgodpl_roadmap_shortest_paths((uint32_t) n, graph.edges.data(), graph.weights.data(), graph.num_edges(), src.data(), g,
													dist.data(), pred.data(), nullptr, stream));
	ShortestPathTables t;
	t.distance.resize(g), t.predecessor.resize(g);
	for (size_t s = 0; s < g; ++s) {
		t.distance[s].assign(dist.begin() + (std::ptrdiff_t) (s * n), dist.begin() + (std::ptrdiff_t) ((s + 1) * n));
		t.predecessor[s].assign(pred.begin() + (std::ptrdiff_t) (s * n), pred.begin() + (std::ptrdiff_t) ((s + 1) * n));
	}
	return t;
}

/// tsp_over_prm.cpp:80-97: {distances, predecessors} from one start node.
inline std::pair<std::vector<double>, std::vector<PRMGraph::vertex_descriptor>> runDijkstra(const PRMGraph &graph,
																							 PRMGraph::vertex_descriptor start_node) {
	ShortestPathTables t = shortest_paths_from(graph, {start_node});
	return {std::move(t.distance[0]), std::move(t.predecessor[0])};
}

/// tsp_over_prm.cpp:111-128: the states from just after the start node up to goal_node (the start node — the one that is
/// its own predecessor — is not emitted, as in the reference).
inline RobotPath retrace_path(const PRMGraph &graph, const std::vector<PRMGraph::vertex_descriptor> &predecessor_lookup,
							  PRMGraph::vertex_descriptor goal_node) {
	RobotPath path;
	// (a predecessor map is a tree for positive edge weights; zero-weight edges between distinct vertices can tie — the walk is
	// bounded by the vertex count so that such a map cannot hang the caller)
	size_t guard = predecessor_lookup.size();
	for (auto cur = goal_node; predecessor_lookup[cur] != cur && guard-- > 0; cur = predecessor_lookup[cur]) path.states.push_back(graph[cur]);
	std::reverse(path.states.begin(), path.states.end());
	return path;
}

/// tsp_over_prm.h:87-97.
struct GoalToGoalPathResults {
	std::vector<std::vector<double>> distance_lookup;
	std::vector<double> start_to_goals_distances;
	std::vector<std::vector<PRMGraph::vertex_descriptor>> predecessor_lookup;
	std::vector<PRMGraph::vertex_descriptor> start_to_goals_predecessors;
};
/// tsp_over_prm.cpp:497-517: Dijkstra from every goal node and from the start node (vertex 0) — one batched call.
inline GoalToGoalPathResults calculate_goal_to_goal_paths(const PRMGraph &graph, const std::vector<PRMGraph::vertex_descriptor> &goal_nodes) {
	std::vector<PRMGraph::vertex_descriptor> sources = goal_nodes;
	sources.push_back(0);
	ShortestPathTables t = shortest_paths_from(graph, sources);
	GoalToGoalPathResults r;
	r.start_to_goals_distances = std::move(t.distance.back()), r.start_to_goals_predecessors = std::move(t.predecessor.back());
	t.distance.pop_back(), t.predecessor.pop_back();
	r.distance_lookup = std::move(t.distance), r.predecessor_lookup = std::move(t.predecessor);
	return r;
}

}  // namespace mgodpl
