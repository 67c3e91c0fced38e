// Host-callable launchers of the device kernels (query_kernels.cu, scene_build.cu, knn_kernels.cu, roadmap_kernels.cu, mesh_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace mgodpl_b200 {

/// Upper bound on queued boxes per launch (64 B each); beyond it stage A traverses inline.
constexpr size_t QUEUE_MAX_ITEMS = 32u << 20;
/// Bytes of device scratch a query launch wants for `worst_case_boxes` candidate boxes (queue header + items).
/// Upper bound on survivor-list entries (8 B each) between the broad cull and FK expansion.
constexpr size_t SURVIVOR_MAX_ENTRIES = 64u << 20;
size_t query_scratch_bytes(size_t worst_case_boxes, size_t survivor_entries = 0);
size_t motion_overflow_bytes(size_t m);
size_t motion_tail_queue_bytes(size_t m, int n_boxes);
/// Does a motion batch of this size run without the survivor-count read-back (i.e. can its launches be recorded into a graph)?
bool motion_batch_is_one_pass(size_t m, int n_boxes);
/// What lets launch_check_motions overlap its two rounds (see there): a side stream and two events, owned by the caller's
/// per-stream slot.
struct MotionOverlap {
	cudaStream_t side = nullptr;
	cudaEvent_t ev_head = nullptr, ev_tail = nullptr;
};
size_t state_survivor_slots(size_t n);

cudaError_t launch_check_states(const SceneDev &sc, const RobotDev &rb, const double *d_states, size_t n, size_t stride,
								uint32_t *d_hit_bits, void *scratch, size_t scratch_bytes, cudaStream_t stream);
cudaError_t launch_check_boxes(const SceneDev &sc, const double *d_boxes, size_t n, uint32_t *d_hit_bits, void *scratch,
							   size_t scratch_bytes, cudaStream_t stream);
/// d_n_steps, d_first_step (m x u32 each) and d_slerp (m x 32 B) are mandatory work buffers; d_toi / d_uncertain /
/// d_override are optional.
cudaError_t launch_check_motions(const SceneDev &sc, const RobotDev &rb, const double *d_a, const double *d_b, size_t m,
								 size_t stride, int js, double max_step, const uint32_t *d_override, uint32_t *d_hit_bits,
								 double *d_toi, uint32_t *d_n_steps, uint32_t *d_uncertain, unsigned *d_first_step, void *d_slerp,
								 void *scratch, size_t scratch_bytes, size_t survivor_entries, cudaStream_t stream, const MotionOverlap *overlap = nullptr);
cudaError_t launch_sample_goals(const SceneDev &sc, const RobotDev &rb, const double *d_targets, size_t f,
								const double *d_draws, size_t s, unsigned long long seed,
								double distance_from_target, uint32_t *d_hit_bits, uint32_t *d_collisions,
								double *d_states_out, void *scratch, size_t scratch_bytes, cudaStream_t stream);
cudaError_t launch_forward_kinematics(const RobotDev &rb, const double *d_states, size_t n, size_t stride, double *d_out,
									  cudaStream_t stream);

/// Exact k-NN over equal_weights_distance; causal != 0 restricts node i to neighbours j < i (incremental PRM insertion).
/// `scratch`: knn_scratch_bytes(n_queries, n_candidates, k) bytes of device memory (the per-chunk partial lists).
size_t knn_scratch_bytes(size_t n_q, size_t n_c, int k);
cudaError_t launch_knn(const double *d_states, size_t n, size_t stride, int js, int k, int causal, int32_t *d_idx, double *d_dist, void *scratch,
					   cudaStream_t stream);
/// The same for n_q query states against a separate database of n_db states (both with the same row layout).
cudaError_t launch_knn_query(const double *d_database, size_t n_db, const double *d_queries, size_t n_q, size_t stride, int js, int k,
							 int32_t *d_idx, double *d_dist, void *scratch, cudaStream_t stream);

/// Shortest paths from many sources at once over a CSR roadmap (both directions of every undirected edge present).
/// d_dist / d_pred: n_sources x n, source-major; d_pred may be null.  Synchronises `stream` (convergence read-backs).
size_t roadmap_scratch_bytes(unsigned n, size_t n_sources);
cudaError_t launch_roadmap_shortest_paths(const uint32_t *d_row, const uint32_t *d_col, const double *d_weight, unsigned n,
										  const uint32_t *d_sources, size_t n_sources, double *d_dist, uint32_t *d_pred, void *scratch,
										  unsigned *h_flag_pinned, unsigned *sweeps_out, cudaStream_t stream);

/// Connected components of the vertex graph of a triangle mesh; d_labels[v] = smallest vertex index of v's component.
template<class Index>
cudaError_t launch_mesh_components(const Index *d_tris, size_t nt, unsigned nv, unsigned *d_labels, cudaStream_t stream);

/// Result of the device BVH build: device buffers owned by the caller (cudaFree).
struct BuiltScene {
	float4 *nodes = nullptr;
	float4 *wide = nullptr;       // 4-wide records folded from `nodes` (nullptr when not requested)
	float4 top[8] = {};          // host copy of wide record 0
	uint32_t n_wide = 0;
	double *tris = nullptr;
	float4 *tris32 = nullptr;
	uint32_t *tri_ids = nullptr;  // original triangle index of every stored triangle slot
	uint32_t n_nodes = 0, n_tris = 0, n_leaves = 0, max_depth = 0, max_leaf = 0;
	double sah_cost = 0, build_ms = 0;
	float root_c[3] = {0, 0, 0}, root_h[3] = {-1, -1, -1};  // scene bounds relative to the origin, rounded outward
	size_t device_bytes = 0, device_bytes_extra = 0;
	unsigned char *clear = nullptr;  // clearance grid (SceneDev::clear), nullptr when not requested
	int clear_n[3] = {0, 0, 0};
	float clear_lo[3] = {0, 0, 0}, clear_cell = 0, clear_max = 0, clear_quantum = 0, clear_build_ms = 0;
};
/// How far beyond the scene bounds the clearance grid reaches = the largest clearance it can certify (metres).
constexpr double CLEAR_PAD = 1.6;
/// d_verts: nv x 3 doubles, d_tris: nt x 3 uint32 (device).  lo/hi: scene bounds (host, from the vertices used).
cudaError_t build_scene(const double *d_verts, size_t nv, const uint32_t *d_tris, size_t nt, const double *lo,
						const double *hi, const double *origin, unsigned max_leaf_size, bool refine, bool wide, bool clearance_grid,
						BuiltScene *out, cudaStream_t stream);

}  // namespace mgodpl_b200
