// Exact k-nearest-neighbour search over equal_weights_distance (src/planning/distance.cpp:17-25) — the step right
// before PRM edge validation.  The reference keeps a GNAT (src/planning/nearest_neighbours/NearestNeighborsGNAT.h,
// queried at src/planning/tsp_over_prm.cpp:44-45) and inserts nodes one at a time, so node i is connected to its k
// nearest among the nodes inserted *before* it; `causal` reproduces exactly that neighbourhood for a whole batch.
// Brute force, tiled through shared memory: one thread per query, 128 candidates staged per tile, top-k kept sorted.
#include "geometry.cuh"
#include "kernels.h"

namespace mgodpl_b200 {

constexpr int KNN_BLOCK = 128;
constexpr int KNN_MAX_K = 32;
constexpr int KNN_ROW_MAX = 7 + MGODPL_MAX_JOINT_VALUES;

__global__ void __launch_bounds__(KNN_BLOCK)
k_knn(const double *__restrict__ states, size_t n, const double *__restrict__ cands, size_t n_cands, size_t stride, int js, int k,
	  int mode, int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
	// mode 0: every other row of the same array; 1 (causal): rows before the query; 2: a separate database, nothing excluded
	const int causal = mode == 1;
	extern __shared__ double tile[];  // KNN_BLOCK rows of (7 + js) doubles
	const int row = 7 + js;
	const size_t q = (size_t) blockIdx.x * KNN_BLOCK + threadIdx.x;
	double me[KNN_ROW_MAX];
	if (q < n)
		for (int i = 0; i < row; ++i) me[i] = __ldg(states + q * stride + i);
	double best_d[KNN_MAX_K];
	int32_t best_i[KNN_MAX_K];
	for (int i = 0; i < k; ++i) best_d[i] = __longlong_as_double(0x7ff0000000000000ll), best_i[i] = -1;
	// causal: candidates j < q only; the block's last query bounds the tiles worth visiting
	const size_t last_q = min(n, (size_t) (blockIdx.x + 1) * KNN_BLOCK);
	const size_t n_cand = causal ? last_q : n_cands;
	for (size_t t0 = 0; t0 < n_cand; t0 += KNN_BLOCK) {
		__syncthreads();
		const size_t tn = min((size_t) KNN_BLOCK, n_cand - t0);
		for (size_t e = threadIdx.x; e < tn * row; e += KNN_BLOCK) tile[e] = __ldg(cands + (t0 + e / row) * stride + e % row);
		__syncthreads();
		if (q >= n) continue;
		for (size_t c = 0; c < tn; ++c) {
			const size_t j = t0 + c;
			if (causal ? j >= q : (mode == 0 && j == q)) continue;
			const double *o = tile + c * row;
			// equal_weights_distance in the reference's operand order (distance.cpp:17-25, Quaternion.h:78-85)
			const double dx = me[0] - o[0], dy = me[1] - o[1], dz = me[2] - o[2];
			double d = sqrt(dx * dx + dy * dy + dz * dz);
			if (d >= best_d[k - 1]) continue;  // the remaining terms are non-negative
			const double cq = fmin(fmax(me[6] * o[6] + me[3] * o[3] + me[4] * o[4] + me[5] * o[5], -1.0), 1.0);
			d += acos(2.0 * cq * cq - 1.0);
			for (int i = 0; i < js; ++i) d += fabs(me[7 + i] - o[7 + i]);
			if (!(d < best_d[k - 1])) continue;
			int pos = k - 1;  // insertion into the ascending list; ties keep the earlier index first
			while (pos > 0 && best_d[pos - 1] > d) {
				best_d[pos] = best_d[pos - 1], best_i[pos] = best_i[pos - 1];
				--pos;
			}
			best_d[pos] = d, best_i[pos] = (int32_t) j;
		}
	}
	if (q < n)
		for (int i = 0; i < k; ++i) {
			out_idx[q * k + i] = best_i[i];
			if (out_dist) out_dist[q * k + i] = best_i[i] >= 0 ? best_d[i] : __longlong_as_double(0x7ff8000000000000ll);
		}
}

cudaError_t launch_knn(const double *d_states, size_t n, size_t stride, int js, int k, int causal, int32_t *d_idx, double *d_dist,
					   cudaStream_t stream) {
	if (!n) return cudaSuccess;
	if (k < 1 || k > KNN_MAX_K) return cudaErrorInvalidValue;
	const size_t smem = (size_t) KNN_BLOCK * (7 + js) * sizeof(double);
	k_knn<<<(unsigned) ((n + KNN_BLOCK - 1) / KNN_BLOCK), KNN_BLOCK, smem, stream>>>(d_states, n, d_states, n, stride, js, k, causal ? 1 : 0, d_idx,
																					 d_dist);
	return cudaGetLastError();
}

cudaError_t launch_knn_query(const double *d_database, size_t n_db, const double *d_queries, size_t n_q, size_t stride, int js, int k,
							 int32_t *d_idx, double *d_dist, cudaStream_t stream) {
	if (!n_q) return cudaSuccess;
	if (k < 1 || k > KNN_MAX_K) return cudaErrorInvalidValue;
	const size_t smem = (size_t) KNN_BLOCK * (7 + js) * sizeof(double);
	k_knn<<<(unsigned) ((n_q + KNN_BLOCK - 1) / KNN_BLOCK), KNN_BLOCK, smem, stream>>>(d_queries, n_q, d_database, n_db, stride, js, k, 2, d_idx,
																					   d_dist);
	return cudaGetLastError();
}

}  // namespace mgodpl_b200
