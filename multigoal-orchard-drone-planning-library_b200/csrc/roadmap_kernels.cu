// Batched single-source shortest paths on the roadmap — the step after PRM edge validation.  The reference runs one
// boost::dijkstra_shortest_paths per goal sample plus one from the start node (runDijkstra,
// src/planning/tsp_over_prm.cpp:80-97; calculate_goal_to_goal_paths, :497-517) over an undirected graph whose edge
// weights are equal_weights_distance values (:56-58).  All sources are independent, so they are solved together.
//
// Method: label correcting.  d[s][v] starts at +inf (0 at the source) and is lowered by d[s][v] = min(d[s][v],
// d[s][u] + w(u,v)) until no label moves.  With non-negative weights IEEE addition is monotone and never decreases a
// label, so the fixed point is the minimum over all paths of the left-to-right sum of their weights — bit for bit the
// value Dijkstra's algorithm returns (it evaluates the same sums, d[u] + w, along the same minimising paths).
//
// Layout: sources are packed 32 to a *source block*; labels are stored [block][vertex][32 sources], so a warp that
// owns (block, vertex) reads one 256-byte line per neighbour, lanes = sources.  Source blocks are swept to convergence a
// group at a time (256 MB of labels per group: large enough to amortise launches and convergence read-backs — smaller,
// L2-resident groups measured slower — small enough to bound scratch memory).  Per (block, vertex) "moved in the
// previous sweep" flags let a warp skip vertices none of whose neighbours moved.
#include <cfloat>

#include "kernels.h"

namespace mgodpl_b200 {

namespace {

constexpr int SSSP_BLOCK = 256;  // 8 warps; one warp per (source block, vertex)
constexpr double SSSP_INF = __builtin_huge_val();

__global__ void __launch_bounds__(SSSP_BLOCK)
k_sssp_init(double *__restrict__ labels, unsigned char *__restrict__ moved, unsigned n, unsigned n_blocks) {
	const size_t total = (size_t) n_blocks * n * 32, tid = (size_t) blockIdx.x * SSSP_BLOCK + threadIdx.x;
	const size_t step = (size_t) gridDim.x * SSSP_BLOCK;
	for (size_t i = tid; i < total; i += step) labels[i] = SSSP_INF;
	for (size_t i = tid; i < (size_t) 3 * n_blocks * n; i += step) moved[i] = 0;
}

__global__ void k_sssp_seed(double *__restrict__ labels, unsigned char *__restrict__ moved, unsigned n, unsigned n_blocks,
							const uint32_t *__restrict__ sources, size_t first_source, size_t n_sources) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;  // local source slot
	if (i >= (size_t) n_blocks * 32 || first_source + i >= n_sources) return;
	const unsigned b = (unsigned) (i >> 5), lane = (unsigned) (i & 31), v = sources[first_source + i];
	labels[((size_t) b * n + v) * 32 + lane] = 0.0;
	moved[(size_t) b * n + v] = 1;  // buffer 0 = "moved before sweep 0"
}

/// One sweep.  `t` selects the roles of the three flag buffers: read (moved before this sweep), write (moved in this
/// sweep) and clear (the buffer the next sweep writes).  sweep_flags[k + 1] is raised when any label moved; a sweep
/// whose predecessor raised nothing returns at once, so the host can enqueue sweeps in batches without a sync each.
__global__ void __launch_bounds__(SSSP_BLOCK)
k_sssp_sweep(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, const double *__restrict__ weight, unsigned n,
			 unsigned n_blocks, double *labels, unsigned char *moved, unsigned t, unsigned k, unsigned *sweep_flags) {
	if (__ldcg(sweep_flags + k) == 0) return;
	const unsigned lane = threadIdx.x & 31, b = blockIdx.y;
	const size_t plane = (size_t) n_blocks * n;
	const unsigned char *moved_before = moved + (size_t) (t % 3) * plane + (size_t) b * n;
	unsigned char *moved_now = moved + (size_t) ((t + 1) % 3) * plane + (size_t) b * n;
	unsigned char *moved_next = moved + (size_t) ((t + 2) % 3) * plane + (size_t) b * n;
	const double *base = labels + (size_t) b * n * 32 + lane;
	bool raised = false;
	for (unsigned v = blockIdx.x * (SSSP_BLOCK / 32) + (threadIdx.x >> 5); v < n; v += gridDim.x * (SSSP_BLOCK / 32)) {
		if (lane == 0) moved_next[v] = 0;
		const unsigned e0 = __ldg(row + v), e1 = __ldg(row + v + 1);
		bool any = false;
		for (unsigned e = e0 + lane; e < e1; e += 32) any |= __ldcg(moved_before + __ldg(col + e)) != 0;
		if (!__any_sync(0xffffffffu, any)) continue;
		double *mine = labels + ((size_t) b * n + v) * 32 + lane;
		const double before = __ldcg(mine);
		double best = before;
#pragma unroll 4
		for (unsigned e = e0; e < e1; ++e) best = fmin(best, __ldcg(base + (size_t) __ldg(col + e) * 32) + __ldg(weight + e));
		const bool lower = best < before;
		if (lower) __stcg(mine, best);
		if (__any_sync(0xffffffffu, lower) && lane == 0) moved_now[v] = 1, raised = true;
	}
	// test before set: a hundred thousand warps storing to one address serialise in L2 and stall every load behind them
	if (raised && __ldcg(sweep_flags + k + 1) == 0) sweep_flags[k + 1] = 1;
}

/// Final pass for one group: pick each vertex's predecessor (the neighbour u with d[u] + w == d[v]; ties go to the
/// smaller d[u], then the smaller index), and write distances / predecessors source-major as the reference stores them
/// (distance_lookup[i][v], tsp_over_prm.h:90-97).  Unreachable: distance DBL_MAX, predecessor = the vertex itself, as
/// boost leaves them.  32 vertices x 32 sources per CTA, transposed through shared memory for coalesced rows.
__global__ void __launch_bounds__(1024)
k_sssp_finish(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, const double *__restrict__ weight, unsigned n,
			  const double *__restrict__ labels, const uint32_t *__restrict__ sources, size_t first_source, size_t n_sources,
			  double *__restrict__ dist_out, uint32_t *__restrict__ pred_out) {
	__shared__ double s_dist[32][33];
	__shared__ uint32_t s_pred[32][33];
	const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y, v0 = blockIdx.x * 32;
	const unsigned v = v0 + w;
	const size_t source_slot = first_source + (size_t) b * 32 + lane;
	if (v < n) {
		const double *base = labels + (size_t) b * n * 32 + lane;
		const double d = __ldg(base + (size_t) v * 32);
		uint32_t pred = v;
		if (pred_out && d < SSSP_INF && source_slot < n_sources && __ldg(sources + source_slot) != v) {
			double best_du = SSSP_INF;
			const unsigned e0 = __ldg(row + v), e1 = __ldg(row + v + 1);
			for (unsigned e = e0; e < e1; ++e) {
				const uint32_t u = __ldg(col + e);
				const double du = __ldg(base + (size_t) u * 32);
				if (u != v && du + __ldg(weight + e) == d && (du < best_du || (du == best_du && u < pred))) best_du = du, pred = u;
			}
		}
		s_dist[w][lane] = d < SSSP_INF ? d : DBL_MAX;
		s_pred[w][lane] = pred;
	}
	__syncthreads();
	// warp w now owns source (b, w) and writes 32 consecutive vertices
	const size_t my_source = first_source + (size_t) b * 32 + w;
	if (my_source < n_sources && v0 + lane < n) {
		dist_out[my_source * n + v0 + lane] = s_dist[lane][w];
		if (pred_out) pred_out[my_source * n + v0 + lane] = s_pred[lane][w];
	}
}

}  // namespace

size_t roadmap_group_blocks(unsigned n, size_t n_sources) {
	static size_t l2_budget = 0;
	if (!l2_budget) {
		const char *e = getenv("MGODPL_SSSP_GROUP_BYTES");
		l2_budget = e && atoll(e) > 0 ? (size_t) atoll(e) : (size_t) 256 << 20;
	}
	const size_t total_blocks = (n_sources + 31) / 32;
	size_t g = l2_budget / ((size_t) n * 256 + 1);
	if (g < 1) g = 1;
	return g < total_blocks ? g : total_blocks;
}

size_t roadmap_scratch_bytes(unsigned n, size_t n_sources) {
	const size_t g = roadmap_group_blocks(n, n_sources);
	return g * n * 256 + 3 * g * n + 64 * sizeof(unsigned) + 1024;
}

cudaError_t launch_roadmap_shortest_paths(const uint32_t *d_row, const uint32_t *d_col, const double *d_weight, unsigned n,
										  const uint32_t *d_sources, size_t n_sources, double *d_dist, uint32_t *d_pred, void *scratch,
										  unsigned *h_flag_pinned, unsigned *sweeps_out, cudaStream_t stream) {
	if (!n || !n_sources) return cudaSuccess;
	const size_t group = roadmap_group_blocks(n, n_sources), total_blocks = (n_sources + 31) / 32;
	double *labels = (double *) scratch;
	unsigned char *moved = (unsigned char *) (labels + group * n * 32);
	unsigned *flags = (unsigned *) (((uintptr_t) (moved + 3 * group * n) + 255) & ~(uintptr_t) 255);
	int sm_dev = 0, sms = 148;  // SM count of the current device (a process may drive several)
	cudaGetDevice(&sm_dev);
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sm_dev) != cudaSuccess || sms <= 0) sms = 148;
	constexpr unsigned BATCH = 8;  // sweeps enqueued between two convergence read-backs
	unsigned sweeps_total = 0;
	cudaError_t err = cudaSuccess;
	for (size_t b0 = 0; b0 < total_blocks; b0 += group) {
		const unsigned nb = (unsigned) (b0 + group <= total_blocks ? group : total_blocks - b0);
		const size_t first_source = b0 * 32;
		k_sssp_init<<<(unsigned) sms * 8, SSSP_BLOCK, 0, stream>>>(labels, moved, n, nb);
		k_sssp_seed<<<(nb * 32 + 127) / 128, 128, 0, stream>>>(labels, moved, n, nb, d_sources, first_source, n_sources);
		// a few resident waves per source block: a sweep that finds the batch already converged costs microseconds, not a full grid
		const unsigned per_block = (n + SSSP_BLOCK / 32 - 1) / (SSSP_BLOCK / 32), cap = ((unsigned) sms * 8u * 4u + nb - 1) / nb;
		const dim3 grid(per_block < cap ? per_block : cap, nb);
		unsigned t = 0;
		for (;;) {
			// flags[0] != 0 lets the batch start; flags[k + 1] is raised by sweep k when it moved a label
			if ((err = cudaMemsetAsync(flags, 1, sizeof(unsigned), stream)) != cudaSuccess) return err;
			if ((err = cudaMemsetAsync(flags + 1, 0, BATCH * sizeof(unsigned), stream)) != cudaSuccess) return err;
			for (unsigned k = 0; k < BATCH; ++k, ++t)
				k_sssp_sweep<<<grid, SSSP_BLOCK, 0, stream>>>(d_row, d_col, d_weight, n, nb, labels, moved, t, k, flags);
			if ((err = cudaGetLastError()) != cudaSuccess) return err;
			if ((err = cudaMemcpyAsync(h_flag_pinned, flags + BATCH, sizeof(unsigned), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return err;
			if ((err = cudaStreamSynchronize(stream)) != cudaSuccess) return err;
			if (*h_flag_pinned == 0) break;  // some sweep of this batch moved nothing: converged
		}
		sweeps_total += t;
		k_sssp_finish<<<dim3((n + 31) / 32, nb), 1024, 0, stream>>>(d_row, d_col, d_weight, n, labels, d_sources, first_source, n_sources, d_dist,
																	d_pred);
		if ((err = cudaGetLastError()) != cudaSuccess) return err;
	}
	if (sweeps_out) *sweeps_out = sweeps_total;
	return cudaSuccess;
}

}  // namespace mgodpl_b200
