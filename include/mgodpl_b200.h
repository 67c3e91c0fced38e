/*
 * mgodpl_b200 — C ABI of the B200-native collision-query path.
 *
 * This is the drop-in boundary for ONE hot path of werner291/Multigoal-Orchard-Drone-Planning-Library ("the
 * reference", paths below are relative to its checkout): RobotState -> forward kinematics -> box-vs-tree-mesh
 * collision, i.e. check_link_collision / check_robot_collision / check_motion_collides / check_path_collides
 * (src/planning/collision_detection.h:29-111), the goal sampler built on them (src/planning/goal_sampling.cpp:51-99)
 * and the PRM edge-validation call sites (src/planning/tsp_over_prm.cpp:33-66).  The reference has no FFI of its
 * own (single C++ process); these entry points are what a binding for this path would have to expose, and the C++
 * facade in multigoal-orchard-drone-planning-library_b200/host/ re-creates the reference's signatures on top of them.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an int status (MGODPL_OK == 0) and never throws.
 *    mgodpl_last_error() gives a thread-local message for the last non-OK status on the calling thread.
 *  - a rigid transform is 7 doubles: translation x y z, then the unit quaternion x y z w
 *    (src/math/Transform.h:20-23, src/math/Quaternion.h:25-27).
 *  - a robot state is `stride` doubles: base transform (7) then `n_joint_values` joint values
 *    (src/planning/RobotState.h:16-23).  n_joint_values may exceed the number of joint *variables*: the surplus is
 *    ignored by FK but counts in the motion distance, exactly like the reference.
 *  - result bitmasks: query i -> bit (i & 31) of word (i >> 5); tail bits are zero.  Bit set == collides.
 *  - "_device" variants take device pointers and only enqueue work on `stream` (a cudaStream_t passed as void*);
 *    the others take host pointers, copy in, run, copy out and synchronise `stream` before returning.  (One exception:
 *    a device-variant batch whose worst case exceeds the internal box queue — more than 2^25 candidate boxes — reads one
 *    counter block back, i.e. synchronises `stream` once, to learn how many passes it needs.)  Batches below that size never
 *    block the host and allocate nothing once the stream's scratch exists (after the first call of that size on the stream),
 *    so they can be recorded into a CUDA graph; replays use the capture stream's scratch, so do not issue other calls on
 *    that stream while a replay is in flight.
 *  - handles are immutable after creation and may be used from several host threads at once.
 *  - there is no CPU fallback: without a CUDA device every entry point returns MGODPL_E_NO_DEVICE.
 */
#ifndef MGODPL_B200_H
#define MGODPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgodpl_scene mgodpl_scene; /* stands where fcl::CollisionObjectd stood (collision_detection.h:30) */
typedef struct mgodpl_robot mgodpl_robot; /* device copy of robot_model::RobotModel (RobotModel.h:36-196)       */

enum mgodpl_status {
	MGODPL_OK = 0,
	MGODPL_E_INVALID_ARGUMENT = 1,
	MGODPL_E_UNSUPPORTED_SHAPE = 2, /* "Only boxes are implemented for collision geometry." collision_detection.cpp:50 */
	MGODPL_E_LINK_NOT_FOUND = 3,    /* "Link not found: ..." RobotModel.cpp:119                                      */
	MGODPL_E_UNKNOWN_JOINT_TYPE = 4,/* "Unknown joint type" RobotModel.h:104                                         */
	MGODPL_E_CUDA = 5,
	MGODPL_E_NO_DEVICE = 6,
	MGODPL_E_OUT_OF_MEMORY = 7,
	MGODPL_E_UNSUPPORTED_ROBOT = 8  /* more links / boxes / joint values than MGODPL_MAX_* */
};

#define MGODPL_MAX_LINKS 12
#define MGODPL_MAX_BOXES 12
#define MGODPL_MAX_JOINT_VALUES 16

/* BOX and MESH are the reference's Shape variant (shapes.h:24-29); it collides boxes only and so does a MESH link here (rejected
 * with the reference's message).  CAPSULE and CONVEX_MESH are additions of this library (BASELINE north_star (3)): they walk the
 * tree as their bounding box and get their own exact test at the triangle leaves. */
enum mgodpl_shape_type { MGODPL_SHAPE_BOX = 0, MGODPL_SHAPE_MESH = 1, MGODPL_SHAPE_CAPSULE = 2, MGODPL_SHAPE_CONVEX_MESH = 3 };
#define MGODPL_MAX_CONVEX_VERTICES 16
enum mgodpl_joint_type { MGODPL_JOINT_REVOLUTE = 0, MGODPL_JOINT_FIXED = 1 }; /* RobotModel.h:48-62 variant order */

/* PositionedShape (positioned_shape.h:12-30) attached to a link; only boxes can be collided (collision_detection.cpp:21-22). */
typedef struct mgodpl_shape_desc {
	int32_t link;
	int32_t shape_type;  /* mgodpl_shape_type */
	double size[3];      /* BOX: Box::size, full extents, centred (shapes.h:20-22).  CAPSULE: size[0] = radius, size[2] = length of
	                      * the cylindrical part along the shape's z axis, centred (fcl::Capsule's convention); size[1] unused.
	                      * CONVEX_MESH: unused (the vertices go to mgodpl_robot_create_ex). */
	double transform[7]; /* relative to the link frame */
} mgodpl_shape_desc;

/* RobotModel::Joint (RobotModel.h:67-84).  Joints must be listed in insertJoint order: the order decides the
 * DFS expansion order of forwardKinematics and with it which joint value feeds which joint (RobotModel.cpp:53-69). */
typedef struct mgodpl_joint_desc {
	int32_t link_a, link_b;
	int32_t joint_type; /* mgodpl_joint_type */
	int32_t reserved;
	double attachment_a[7], attachment_b[7];
	double axis[3];                 /* revolute only */
	double min_angle, max_angle;    /* revolute only */
} mgodpl_joint_desc;

typedef struct mgodpl_scene_opts {
	uint32_t flags;          /* MGODPL_SCENE_* */
	uint32_t max_leaf_size;  /* triangles per BVH leaf after collapsing, 0 = default */
} mgodpl_scene_opts;
#define MGODPL_SCENE_COUNTERS 1u   /* keep device-side work counters (states, nodes visited, leaf tests) */
#define MGODPL_SCENE_NO_REFINE 2u  /* plain Morton LBVH, skip the SAH refinement passes */
#define MGODPL_SCENE_NO_CLEARANCE_GRID 4u /* skip the clearance grid (the one-byte-per-cell "certainly free" broad phase) */

typedef struct mgodpl_scene_info {
	uint64_t n_triangles, n_nodes, n_leaves, device_bytes;
	double bounds_lo[3], bounds_hi[3];
	double build_ms;       /* device time of the BVH build */
	double sah_cost;       /* surface-area-heuristic cost of the final tree (C_node = 1, C_tri = 1) */
	uint32_t max_depth, max_leaf_size;
	uint64_t n_wide_nodes;  /* 128-byte 4-wide records the traversal walks (0: it walks the n_nodes 64-byte binary records) */
	uint64_t traversal_bytes; /* bytes one pass over the traversal's own data reads: its records + both triangle copies */
	uint64_t clearance_cells; /* cells of the clearance grid (one byte each), 0 = none */
	double clearance_cell_size, clearance_build_ms;
} mgodpl_scene_info;

/* Deterministic cost counters: the reference's own observability hook on this path is an invocation counter
 * (src/planning/functional_utils.cppm:27-32, approach_planning.cpp:505-533). */
typedef struct mgodpl_counters {
	uint64_t states_checked;   /* robot states that went through FK + collision */
	uint64_t boxes_traversed;  /* boxes that survived the root cull and entered the BVH */
	uint64_t nodes_visited;
	uint64_t leaf_tests;       /* exact box/triangle tests */
	uint64_t motions_checked;
	uint64_t uncertain_steps;  /* motions whose step count sat within 1e-9 of an integer boundary (re-derived on host) */
	uint64_t max_nodes_per_box;/* longest single box descent (BVH records visited): the traversal's critical path */
} mgodpl_counters;

const char *mgodpl_last_error(void);
int mgodpl_device_count(void);
/* Multi-GPU hosts: a scene is built on the calling thread's current CUDA device and stays bound to it (every later call on
 * the handle runs there, whatever device is current in the caller, and restores the caller's device).  Robot handles hold no
 * device memory and work with scenes on any device.  mgodpl_set_device = cudaSetDevice for callers that do not link the CUDA
 * runtime themselves (one scene replica per device, one host thread per replica: the reference's analogue is its
 * std::for_each(std::execution::par, ...) over problems, src/benchmarks/goal_sample_success_rate.cpp:165-172). */
int mgodpl_set_device(int device);
int mgodpl_scene_device(const mgodpl_scene *scene, int *device);

/* ---- scene: replaces meshToFclBVH(const Mesh&) / treeMeshesToFclCollisionObject (src/planning/fcl_utils.cpp:23-60) ----
 * verts: nv x 3 doubles; tris: nt x 3 vertex indices.  Builds the linear BVH on the current CUDA device. */
int mgodpl_scene_create(const double *verts, size_t nv, const uint32_t *tris, size_t nt, const mgodpl_scene_opts *opts,
						mgodpl_scene **out);
/* Same, with the reference's index type (Mesh::triangles is vector<array<size_t,3>>, src/planning/Mesh.h:19-22). */
int mgodpl_scene_create_u64(const double *verts, size_t nv, const uint64_t *tris, size_t nt,
							const mgodpl_scene_opts *opts, mgodpl_scene **out);
int mgodpl_scene_destroy(mgodpl_scene *scene);
int mgodpl_scene_get_info(const mgodpl_scene *scene, mgodpl_scene_info *info);
int mgodpl_scene_get_counters(mgodpl_scene *scene, mgodpl_counters *out, int reset);

/* ---- robot: flattened RobotModel (links in insertLink order, joints in insertJoint order) -------------------------
 * root_link is what robot.findLinkByName("flying_base") returns (collision_detection.cpp:65); end_effector_link is
 * findLinkByName("end_effector") or -1 when goal sampling is not needed. */
int mgodpl_robot_create(int32_t n_links, const mgodpl_shape_desc *shapes, int32_t n_shapes,
						const mgodpl_joint_desc *joints, int32_t n_joints, int32_t root_link, int32_t end_effector_link,
						mgodpl_robot **out);
/* The same with convex-polytope links: shape i of type MGODPL_SHAPE_CONVEX_MESH is the convex hull of the vertices
 * mesh_vertices[3 * mesh_vertex_offsets[i] .. 3 * mesh_vertex_offsets[i + 1]) (shape frame; 4 .. MGODPL_MAX_CONVEX_VERTICES points
 * with non-zero volume); mesh_vertex_offsets has n_shapes + 1 entries.  Both arrays may be NULL when no shape needs them. */
int mgodpl_robot_create_ex(int32_t n_links, const mgodpl_shape_desc *shapes, int32_t n_shapes,
						   const mgodpl_joint_desc *joints, int32_t n_joints, int32_t root_link, int32_t end_effector_link,
						   const double *mesh_vertices, const int32_t *mesh_vertex_offsets, mgodpl_robot **out);
int mgodpl_robot_destroy(mgodpl_robot *robot);
int mgodpl_robot_n_variables(const mgodpl_robot *robot); /* RobotModel::count_joint_variables, RobotModel.cpp:130-134 */

/* ---- forwardKinematics (src/planning/RobotModel.cpp:17-90), batched: out_link_tfs is n x n_links x 7 doubles ---- */
int mgodpl_forward_kinematics(const mgodpl_robot *robot, const double *states, size_t n, size_t stride,
							  int32_t n_joint_values, double *out_link_tfs, void *stream);

/* ---- check_link_collision (collision_detection.cpp:16-55) for explicit boxes: n x (world transform 7 + size 3) ---- */
int mgodpl_check_boxes(mgodpl_scene *scene, const double *boxes, size_t n, uint32_t *hit_bits, void *stream);

/* ---- check_robot_collision (collision_detection.cpp:57-80), N states ---------------------------------------------- */
int mgodpl_check_states(mgodpl_scene *scene, const mgodpl_robot *robot, const double *states, size_t n, size_t stride,
						int32_t n_joint_values, uint32_t *hit_bits, void *stream);
int mgodpl_check_states_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_states, size_t n,
							   size_t stride, int32_t n_joint_values, uint32_t *d_hit_bits, void *stream);

/* ---- check_motion_collides (collision_detection.cpp:82-105), M motions a[i] -> b[i] --------------------------------
 * max_step: the reference hard-codes 0.1 (collision_detection.cpp:90, collision_detection.cppm:61).
 * toi (optional): first colliding interpolation parameter (NaN where the motion is free).
 * n_steps_out (optional): the step count used for each motion.
 * The host variant guarantees the reference's step count bit-for-bit: motions whose distance/max_step lands within
 * 1e-9 of an integer are re-derived with the host libm and re-run if the count differs. */
int mgodpl_check_motions(mgodpl_scene *scene, const mgodpl_robot *robot, const double *a, const double *b, size_t m,
						 size_t stride, int32_t n_joint_values, double max_step, uint32_t *hit_bits, double *toi,
						 uint32_t *n_steps_out, void *stream);
/* Device variant: d_n_steps_override (optional) forces the step count of motion i when its entry is not 0xFFFFFFFF;
 * d_uncertain_bits (optional) receives one bit per motion whose step count is boundary-sensitive. */
int mgodpl_check_motions_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_a, const double *d_b,
								size_t m, size_t stride, int32_t n_joint_values, double max_step,
								const uint32_t *d_n_steps_override, uint32_t *d_hit_bits, double *d_toi,
								uint32_t *d_n_steps_out, uint32_t *d_uncertain_bits, void *stream);

/* ---- check_path_collides (collision_detection.cpp:107-122): w waypoints, returns first colliding segment or -1 ---- */
int mgodpl_check_path(mgodpl_scene *scene, const mgodpl_robot *robot, const double *states, size_t w, size_t stride,
					  int32_t n_joint_values, double max_step, int64_t *first_colliding_segment, void *stream);

/* ---- goal sampling: genGoalStateUniform + collision filter (goal_sampling.cpp:51-99,
 *      goal_sample_success_rate.cpp:121-142) for F targets x S samples ------------------------------------------------
 * draws: S x (1 + n_variables) doubles = the yaw and joint values genUprightState would draw (goal_sampling.cpp:13-49),
 *        produced by the caller's std::mt19937 so sampled states match the reference bit for bit; shared by all targets
 *        (the reference re-seeds per problem).  Pass NULL to draw on the device from a counter-based generator (`seed`).
 * hit_bits: F x ceil(S/32) words; collisions: F counts; states_out (optional): F x S x (7 + n_variables) doubles. */
int mgodpl_sample_goals(mgodpl_scene *scene, const mgodpl_robot *robot, const double *targets, size_t f,
						const double *draws, size_t s, uint64_t seed, double distance_from_target, uint32_t *hit_bits,
						uint32_t *collisions, double *states_out, void *stream);
int mgodpl_sample_goals_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_targets, size_t f,
							   const double *d_draws, size_t s, uint64_t seed, double distance_from_target,
							   uint32_t *d_hit_bits, uint32_t *d_collisions, double *d_states_out, void *stream);

/* ---- k nearest neighbours under equal_weights_distance (src/planning/distance.cpp:17-25): the NearestNeighborsGNAT
 *      query issued before every batch of PRM edge checks (src/planning/tsp_over_prm.cpp:44-45, 383-386), for all n
 *      states at once.  causal != 0: state i only sees states j < i, i.e. the roadmap as it stood when the reference
 *      inserted node i.  idx: n x k indices ascending by distance (-1 pads when fewer than k candidates exist);
 *      dist (optional): n x k distances (NaN pads).  1 <= k <= 32.  The k neighbours are DISTINCT; the reference's GNAT can list
 *      one neighbour twice (its Node::split never finds a point's closest pivot, NearestNeighborsGNAT.h:398-405): with the
 *      duplicates removed its list is the leading part of this one. ---------------------------------------------------------- */
int mgodpl_knn(const double *states, size_t n, size_t stride, int32_t n_joint_values, int32_t k, int32_t causal, int32_t *idx,
			   double *dist, void *stream);
int mgodpl_knn_device(const double *d_states, size_t n, size_t stride, int32_t n_joint_values, int32_t k, int32_t causal,
					  int32_t *d_idx, double *d_dist, void *stream);
/* The same query for states that are not part of the index: goal samples and the start state are connected to their k
 * nearest *infrastructure* nodes but never inserted into the GNAT (sample_and_connect_goal_states,
 * src/planning/tsp_over_prm.cpp:297-344; :621-629).  idx: n_queries x k indices into `database`. */
int mgodpl_knn_query(const double *database, size_t n_database, const double *queries, size_t n_queries, size_t stride,
					 int32_t n_joint_values, int32_t k, int32_t *idx, double *dist, void *stream);

/* ---- shortest paths on the roadmap: runDijkstra (src/planning/tsp_over_prm.cpp:80-97, boost::dijkstra_shortest_paths on
 *      the undirected PRMGraph of tsp_over_prm.h:25-26) as calculate_goal_to_goal_paths calls it once per goal sample and
 *      once from the start node (tsp_over_prm.cpp:497-517) — all sources in one call.  edges: n_edges x (u, v), undirected;
 *      weights: fp64, finite, >= 0 (equal_weights_distance of the endpoints, tsp_over_prm.cpp:56-58).  dist: n_sources x
 *      n_vertices, source-major (distance_lookup[i][v], tsp_over_prm.h:90-97) — bit-identical to Dijkstra's labels, DBL_MAX
 *      where unreachable (boost's default infinity); pred (optional): same shape, the vertex before v on a shortest path,
 *      v itself for the source and for unreachable vertices.  Where several neighbours tie exactly, the one nearer the
 *      source, then the lower index, is reported (boost reports whichever its heap relaxed first).  With positive weights
 *      the predecessors form a tree.  Distinct vertices joined by edges that VANISH in the sum (zero weight, or below half an
 *      ulp of the distance) sit at one distance: the host-buffer variant passes its result through
 *      mgodpl_roadmap_repair_predecessors, so that it is a tree again; the _device variant leaves the predecessors as the
 *      kernel picked them, and three or more such vertices may name each other (distances stay exact) — copy them to the host
 *      and repair them there, or bound any walk over `pred` by n_vertices, as the facade does.  sweeps_out (optional):
 *      label-correcting sweeps enqueued.  The _device variant takes the graph in CSR form (row_offsets: n_vertices + 1,
 *      neighbours / weights: both directions of every edge) in device memory and synchronises the stream before returning. */
int mgodpl_roadmap_shortest_paths(uint32_t n_vertices, const uint32_t *edges, const double *weights, size_t n_edges,
								  const uint32_t *sources, size_t n_sources, double *dist, uint32_t *pred, uint32_t *sweeps_out,
								  void *stream);
int mgodpl_roadmap_shortest_paths_device(const uint32_t *d_row_offsets, const uint32_t *d_neighbours, const double *d_weights,
										 uint32_t n_vertices, const uint32_t *d_sources, size_t n_sources, double *d_dist,
										 uint32_t *d_pred, uint32_t *sweeps_out, void *stream);

/* Host-side utility, no device involved: makes ONE source's predecessors a tree where vanishing edges let vertices at one
 * distance name each other.  row_offsets / neighbours / weights: the CSR graph as the _device variant takes it, in host memory;
 * dist / pred: one source's n_vertices labels and predecessors (pred is rewritten in place).  Vertices whose predecessor is
 * strictly nearer the source are left alone; the others take a strictly nearer neighbour on a shortest path if they have one,
 * and are otherwise reached breadth-first over the vanishing edges, so every walk over `pred` ends at the source. */
int mgodpl_roadmap_repair_predecessors(uint32_t n_vertices, const uint32_t *row_offsets, const uint32_t *neighbours,
									   const double *weights, const double *dist, uint32_t *pred);

/* ---- connected components of a mesh's vertex graph: connected_vertex_components
 *      (src/experiment_utils/mesh_connected_components.cpp:25-74), which loadTreeMeshes uses to split a tree model's
 *      combined fruit mesh into individual fruit (src/experiment_utils/TreeMeshes.cpp:28-66).  tris: nt x 3 vertex indices
 *      (uint32, or uint64 = the reference's size_t, Mesh.h:19-22); labels[v] = the smallest vertex index of v's component
 *      (the reference returns the same partition in unordered_map order). ------------------------------------------------ */
int mgodpl_mesh_components(const uint32_t *tris, size_t nt, size_t nv, uint32_t *labels, void *stream);
int mgodpl_mesh_components_u64(const uint64_t *tris, size_t nt, size_t nv, uint32_t *labels, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MGODPL_B200_H */
