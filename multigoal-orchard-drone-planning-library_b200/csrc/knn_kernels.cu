// Exact k-nearest-neighbour search over equal_weights_distance (src/planning/distance.cpp:17-25) — the step right
// before PRM edge validation.  The reference keeps a GNAT (src/planning/nearest_neighbours/NearestNeighborsGNAT.h,
// queried at src/planning/tsp_over_prm.cpp:44-45) and inserts nodes one at a time, so node i is connected to its k
// nearest among the nodes inserted *before* it; `causal` reproduces exactly that neighbourhood for a whole batch.
//
// Brute force, tiled in two dimensions: block (x, y) holds 128 queries (one per thread) and scans candidate chunk y,
// staged 128 rows at a time through shared memory.  Work units are uniform — in causal mode the chunks that lie
// entirely behind a query tile are simply not scanned, instead of the last query tiles scanning 100 k candidates while
// the first scan 128 — and the grid is as large as the batch allows.  Each thread keeps its best K of the chunk sorted
// in registers (K a template parameter: no runtime-indexed arrays); a second kernel merges the per-chunk lists in
// chunk order, which preserves the tie rule of the sequential scan (equal distances: lower index first).
//
// Nearly every candidate is rejected by a partial sum of the metric — translation distance plus the differences of the first
// two joint values (every other term is non-negative).  That test runs in fp32, squared so that it needs no root, on
// per-component arrays of the tile (five broadcast 16-byte loads bring four candidates: 9 + 5 instructions per candidate where
// the first version spent 20, half of them re-deriving a shared-memory address and re-converting the query), with margins that
// cover the fp32 rounding: it can only reject candidates the fp64 test would reject too.  Whatever passes is measured in fp64
// in the reference's operand order.  Tried without gain: cp.async double buffering of the tile (the stall samples on the tile
// load are the block's warps waiting for each other at the barrier, not for memory), 5 or 6 blocks per SM (spills).
#include "geometry.cuh"
#include "kernels.h"

namespace mgodpl_b200 {

constexpr int KNN_BLOCK = 128;   // queries per block = candidates per shared-memory tile
constexpr int KNN_MAX_K = 32;
constexpr int KNN_MAX_CHUNKS = 64;

template<int K>
struct TopK {
	// Ascending list in registers.  Only k <= K entries are wanted: the first K - k slots hold -inf sentinels that nothing
	// can pass, so the k real entries live in slots [K - k, K), the k-th best is always d[K - 1], and no register is ever
	// indexed with a runtime value.
	double d[K];
	int32_t i[K];
	__device__ __forceinline__ void init(int k) {
#pragma unroll
		for (int p = 0; p < K; ++p) d[p] = __longlong_as_double(p < K - k ? 0xfff0000000000000ll : 0x7ff0000000000000ll), i[p] = -1;
	}
	__device__ __forceinline__ double thr() const { return d[K - 1]; }
	/// Precondition: dist < thr().  The newcomer replaces the worst entry and bubbles up; it stops behind equal distances,
	/// so among ties the entry inserted first (the lower index) stays first.
	__device__ __forceinline__ void insert(double dist, int32_t idx) {
		d[K - 1] = dist, i[K - 1] = idx;
#pragma unroll
		for (int s = 0; s < K - 1; ++s) {
			const int p = K - 1 - s;
			const bool sw = d[p - 1] > d[p];
			const double td = d[p];
			const int32_t ti = i[p];
			d[p] = sw ? d[p - 1] : td, i[p] = sw ? i[p - 1] : ti;
			d[p - 1] = sw ? td : d[p - 1], i[p - 1] = sw ? ti : i[p - 1];
		}
	}
	/// Real entry number o (0 = best) sits in slot K - k + o.
	template<class F>
	__device__ __forceinline__ void for_each(int k, F f) const {
#pragma unroll
		for (int p = 0; p < K; ++p) {
			const int o = p - (K - k);
			if (o >= 0) f(o, d[p], i[p]);
		}
	}
};

/// grid (query tiles, candidate chunks).  mode 0: every other row of the same array; 1 (causal): rows before the query;
/// 2: a separate database, nothing excluded.  part_*: [query][chunk][k] partial lists.
///
/// Per 128-candidate tile a thread works in two phases.  Phase 1 (fp32, every candidate): reject on translation distance +
/// first joint differences (squared form), then on a lower bound of the whole metric (translation + 2 sqrt(2 (1 - |qa . qb|)) <= the angular term, since
/// acos(x) >= sqrt(2 (1 - x)), + joint differences), both with margins that cover the fp32 rounding; survivors are noted in a
/// per-thread list in shared memory.  Phase 2 (fp64, the few survivors): the exact metric in the reference's operand order and
/// the insertion.  Separating the phases matters on a SIMT machine: interleaved, one passing lane makes the whole warp walk
/// through the fp64 code (acos, sqrt) for almost every candidate — measured 5.8 active lanes, 47 ms for 100 k causal k = 10.
template<int K>
__global__ void __launch_bounds__(KNN_BLOCK)
k_knn_scan(const double *__restrict__ queries, size_t n_q, const double *__restrict__ cands, size_t n_c, size_t stride, int js, int k, int mode,
		   size_t chunk, int n_chunks, int32_t *__restrict__ part_idx, double *__restrict__ part_dist) {
	extern __shared__ __align__(16) unsigned char knn_smem[];
	const int row = 7 + js, row32 = (row + 3) & ~3;  // fp32 rows padded to whole float4s
	const float inv_row = 1.0f / (float) row;
	double *tile = reinterpret_cast<double *>(knn_smem);                              // KNN_BLOCK x row     (fp64 candidates)
	double *qj = tile + (size_t) KNN_BLOCK * row;                                     // KNN_BLOCK x js      (the queries' joint values)
	float *tile32 = reinterpret_cast<float *>(qj + (size_t) KNN_BLOCK * js);          // KNN_BLOCK x row32   (fp32 candidates)
	float *qj32 = tile32 + (size_t) KNN_BLOCK * row32;                                // KNN_BLOCK x js      (the same in fp32)
	// the candidates' translations and first two joint values once more, one array per component (phase 1a); joints a robot does
	// not have read as 0 on both sides
	float *xs = qj32 + (size_t) KNN_BLOCK * js, *ys = xs + KNN_BLOCK, *zs = ys + KNN_BLOCK, *j0s = zs + KNN_BLOCK, *j1s = j0s + KNN_BLOCK;
	unsigned char *pending = reinterpret_cast<unsigned char *>(j1s + KNN_BLOCK);                      // KNN_BLOCK entries x KNN_BLOCK threads
	const unsigned pend = (unsigned) __cvta_generic_to_shared(pending) + threadIdx.x;                // this thread's list: entry p at pend + p * KNN_BLOCK
	const size_t q = (size_t) blockIdx.x * KNN_BLOCK + threadIdx.x;
	const bool live = q < n_q;
	double me[7] = {0, 0, 0, 0, 0, 0, 1};
	if (live) {
#pragma unroll
		for (int a = 0; a < 7; ++a) me[a] = __ldg(queries + q * stride + a);
		for (int a = 0; a < js; ++a) {
			const double v = __ldg(queries + q * stride + 7 + a);
			qj[threadIdx.x * js + a] = v, qj32[threadIdx.x * js + a] = (float) v;
		}
	}
	float m32[7];
#pragma unroll
	for (int a = 0; a < 7; ++a) {
		m32[a] = (float) me[a];
		asm volatile("" : "+f"(m32[a]));  // opaque: otherwise the compiler re-converts from the fp64 copy inside the candidate loop
	}
	const float mag = fabsf(m32[0]) + fabsf(m32[1]) + fabsf(m32[2]) + 1.f;
	float mj0 = live && js > 0 ? (float) __ldg(queries + q * stride + 7) : 0.f, mj1 = live && js > 1 ? (float) __ldg(queries + q * stride + 8) : 0.f;
	asm volatile("" : "+f"(mj0), "+f"(mj1));
	TopK<K> best;
	best.init(k);
	// fp32 rejection threshold on the metric (+inf while the list is not full)
	float thr1 = __int_as_float(0x7f800000);
	// causal: only candidates j < q; the block's last query bounds what is worth scanning
	const size_t last_q = min(n_q, ((size_t) blockIdx.x + 1) * KNN_BLOCK);
	const size_t c_end = min(mode == 1 ? min(n_c, last_q) : n_c, ((size_t) blockIdx.y + 1) * chunk);
	for (size_t t0 = (size_t) blockIdx.y * chunk; t0 < c_end; t0 += KNN_BLOCK) {
		__syncthreads();
		const size_t tn = min((size_t) KNN_BLOCK, c_end - t0);
		for (unsigned e = threadIdx.x; e < (unsigned) tn * (unsigned) row; e += KNN_BLOCK) {
			const unsigned r = __float2uint_rz(((float) e + 0.5f) * inv_row), a = e - r * (unsigned) row;  // e / row (exact: e < 2^12), no integer division
			const double v = __ldg(cands + (t0 + r) * stride + a);
			tile[e] = v, tile32[r * row32 + a] = (float) v;
			if (a < 3) (a == 0 ? xs : a == 1 ? ys : zs)[r] = (float) v;
			else if (a == 7) j0s[r] = (float) v;
			else if (a == 8) j1s[r] = (float) v;
		}
		if (js < 2) {  // (uniform) absent joints: zeros
			if (js < 1) j0s[threadIdx.x] = 0.f;
			j1s[threadIdx.x] = 0.f;
		}
		__syncthreads();
		if (!live) continue;
		const unsigned tn_me = (unsigned) (mode == 1 ? (q > t0 ? min(tn, q - t0) : 0) : tn);
		// ---- phase 1a: every candidate of the tile against a lower bound of the metric that needs no square root and no
		//      quaternion: translation distance + the differences of the first two joint values <= threshold, tested as
		//      |dt|^2 <= (threshold - joint differences)^2.  Five broadcast 16-byte loads bring four candidates; who passes is noted
		//      as a bit (phase 1b walks the bits).  Under the PRM sampler the joint terms alone reject three candidates in four
		//      that the translation would let through — at full lane occupancy here instead of 12 of 32 lanes in phase 1b.
		//      Rows beyond tn_me (ragged last tile, causal diagonal) are computed and masked off. ------------------------------------
		unsigned near[KNN_BLOCK / 32];
#pragma unroll
		for (int g = 0; g < KNN_BLOCK / 32; ++g) {
			unsigned m = 0;
#pragma unroll
			for (int c4 = 0; c4 < 8; ++c4) {
				const int at = g * 32 + c4 * 4;
				const float4 X = *reinterpret_cast<const float4 *>(xs + at), Y = *reinterpret_cast<const float4 *>(ys + at), Z = *reinterpret_cast<const float4 *>(zs + at);
				const float4 A = *reinterpret_cast<const float4 *>(j0s + at), B = *reinterpret_cast<const float4 *>(j1s + at);
				const float x[4] = {X.x, X.y, X.z, X.w}, y[4] = {Y.x, Y.y, Y.z, Y.w}, z[4] = {Z.x, Z.y, Z.z, Z.w};
				const float ja[4] = {A.x, A.y, A.z, A.w}, jb[4] = {B.x, B.y, B.z, B.w};
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const float fx = m32[0] - x[u], fy = m32[1] - y[u], fz = m32[2] - z[u];
					// what the translation may still use: phase 1b keeps a candidate only if 0.99999 |dt| + joints <= thr1, so 1.00002
					// on the square (and 1e-6 for the rounding of the joint sum) keeps everything phase 1b would keep
					const float rem = thr1 - (fabsf(mj0 - ja[u]) + fabsf(mj1 - jb[u])) + 1e-6f;
					if (fx * fx + fy * fy + fz * fz <= rem * fabsf(rem) * 1.00003f) m |= 1u << (c4 * 4 + u);  // (rem < 0: the joints alone rule it out)
				}
			}
			const unsigned lo = (unsigned) g * 32u;
			near[g] = tn_me >= lo + 32u ? m : (tn_me > lo ? m & ((1u << (tn_me - lo)) - 1u) : 0u);
		}
		// ---- phase 1b: lower bound of the whole metric for the near ones (every lane walks its own bits: the lanes stay
		//      together, where testing inline would drag the warp through this code for almost every candidate) ----------------------
		unsigned n_pending = 0;
#pragma unroll
		for (int g = 0; g < KNN_BLOCK / 32; ++g) {
			for (unsigned m = near[g]; m; m &= m - 1) {
				const unsigned c = (unsigned) g * 32u + (unsigned) __ffs(m) - 1u;
				const float4 o0 = *reinterpret_cast<const float4 *>(tile32 + c * row32);      // x y z qx
				const float4 o1 = *reinterpret_cast<const float4 *>(tile32 + c * row32 + 4);  // qy qz qw j0
				const float fx = m32[0] - o0.x, fy = m32[1] - o0.y, fz = m32[2] - o0.z, t2 = fx * fx + fy * fy + fz * fz;
				const float dotq = fabsf(m32[3] * o0.w + m32[4] * o1.x + m32[5] * o1.y + m32[6] * o1.z);
				// x * rsqrt(x) (2 ulp) instead of an IEEE square root: the 0.99999 keeps the sum a lower bound
				const float u = 2.f * fmaxf(1.f - dotq - 1e-6f, 0.f);
				float lb = 0.99999f * (t2 * rsqrtf(fmaxf(t2, 1e-30f)) + 2.f * u * rsqrtf(fmaxf(u, 1e-30f)));
				for (int a = 0; a < js; ++a) lb += fabsf(qj32[threadIdx.x * js + a] - tile32[c * row32 + 7 + a]);
				if (lb <= thr1) {
					asm volatile("st.shared.u8 [%0], %1;" ::"r"(pend + n_pending * KNN_BLOCK), "r"(c) : "memory");
					++n_pending;
				}
			}
		}
		// ---- phase 2: the exact metric for what is left, in candidate order (equal distances: lower index first) ------------------
		for (unsigned p = 0; p < n_pending; ++p) {
			unsigned c;
			asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(pend + p * KNN_BLOCK) : "memory");
			const size_t j = t0 + c;
			if (mode == 0 && j == q) continue;
			const double *o = tile + c * row;
			// equal_weights_distance in the reference's operand order (distance.cpp:17-25, Quaternion.h:78-85)
			const double dx = me[0] - o[0], dy = me[1] - o[1], dz = me[2] - o[2];
			double d = sqrt(dx * dx + dy * dy + dz * dz);
			if (d >= best.thr()) continue;  // the remaining terms are non-negative
			const double cq = fmin(fmax(me[6] * o[6] + me[3] * o[3] + me[4] * o[4] + me[5] * o[5], -1.0), 1.0);
			d += acos(2.0 * cq * cq - 1.0);
			for (int a = 0; a < js; ++a) d += fabs(qj[threadIdx.x * js + a] - o[7 + a]);
			if (!(d < best.thr())) continue;
			best.insert(d, (int32_t) j);
		}
		if (n_pending && best.thr() < 1e30) {
			// fp32 bounds: coordinates carry <= 2^-24 relative rounding each, and only candidates within thr of the query matter;
			// 2e-5 more covers the joint sums and the 1e-6 taken off the quaternion dot product above
			thr1 = __double2float_ru(best.thr()) * 1.00001f + 4e-7f * (mag + __double2float_ru(best.thr())) + 2e-5f;
		}
	}
	if (live) {
		int32_t *pi = part_idx + (q * n_chunks + blockIdx.y) * (size_t) k;
		double *pd = part_dist + (q * n_chunks + blockIdx.y) * (size_t) k;
		best.for_each(k, [&](int o, double d, int32_t j) { pi[o] = j, pd[o] = d; });
	}
}

/// One thread per query: merge the per-chunk lists (ascending chunk = ascending candidate index) into the final k.
template<int K>
__global__ void __launch_bounds__(KNN_BLOCK)
k_knn_merge(size_t n_q, int k, int n_chunks, const int32_t *__restrict__ part_idx, const double *__restrict__ part_dist,
			int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
	const size_t q = (size_t) blockIdx.x * KNN_BLOCK + threadIdx.x;
	if (q >= n_q) return;
	TopK<K> best;
	best.init(k);
	for (int c = 0; c < n_chunks; ++c) {
		const int32_t *pi = part_idx + (q * n_chunks + c) * (size_t) k;
		const double *pd = part_dist + (q * n_chunks + c) * (size_t) k;
		for (int p = 0; p < k; ++p) {
			const int32_t j = pi[p];
			const double d = pd[p];
			if (j < 0 || !(d < best.thr())) break;  // lists are ascending: nothing further in this chunk can enter
			best.insert(d, j);
		}
	}
	best.for_each(k, [&](int o, double d, int32_t j) {
		out_idx[q * k + o] = j;
		if (out_dist) out_dist[q * k + o] = j >= 0 ? d : __longlong_as_double(0x7ff8000000000000ll);
	});
}

namespace {
struct KnnPlan {
	size_t chunk;
	int n_chunks;
};
KnnPlan knn_plan(size_t n_q, size_t n_c) {
	int sms = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t q_tiles = (n_q + KNN_BLOCK - 1) / KNN_BLOCK, c_tiles = std::max<size_t>(1, (n_c + KNN_BLOCK - 1) / KNN_BLOCK);
	// enough blocks to fill the GPU a few times over, but no more chunks than there are candidate tiles (or KNN_MAX_CHUNKS)
	size_t want = (8 * (size_t) std::max(sms, 1) + q_tiles - 1) / std::max<size_t>(q_tiles, 1);
	want = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, c_tiles), KNN_MAX_CHUNKS));
	const size_t tiles_per_chunk = (c_tiles + want - 1) / want;
	KnnPlan p;
	p.chunk = tiles_per_chunk * KNN_BLOCK;
	p.n_chunks = (int) ((c_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
	return p;
}

template<int K>
cudaError_t run_knn(const double *d_q, size_t n_q, const double *d_c, size_t n_c, size_t stride, int js, int k, int mode, int32_t *d_idx,
					double *d_dist, void *scratch, cudaStream_t stream) {
	const KnnPlan p = knn_plan(n_q, n_c);
	int32_t *part_idx = reinterpret_cast<int32_t *>(scratch);
	double *part_dist = reinterpret_cast<double *>(reinterpret_cast<char *>(scratch) + ((n_q * p.n_chunks * k * sizeof(int32_t) + 255) & ~(size_t) 255));
	const size_t row = 7 + (size_t) js, row32 = (row + 3) & ~(size_t) 3;
	const size_t smem = (size_t) KNN_BLOCK * (row * sizeof(double) + js * sizeof(double) + row32 * sizeof(float) + js * sizeof(float) + 5 * sizeof(float) + KNN_BLOCK);
	const dim3 grid((unsigned) ((n_q + KNN_BLOCK - 1) / KNN_BLOCK), (unsigned) p.n_chunks);
	cudaError_t ea = cudaFuncSetAttribute(k_knn_scan<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (ea != cudaSuccess) return ea;
	k_knn_scan<K><<<grid, KNN_BLOCK, smem, stream>>>(d_q, n_q, d_c, n_c, stride, js, k, mode, p.chunk, p.n_chunks, part_idx, part_dist);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	k_knn_merge<K><<<grid.x, KNN_BLOCK, 0, stream>>>(n_q, k, p.n_chunks, part_idx, part_dist, d_idx, d_dist);
	return cudaGetLastError();
}

cudaError_t dispatch_knn(const double *d_q, size_t n_q, const double *d_c, size_t n_c, size_t stride, int js, int k, int mode, int32_t *d_idx,
						 double *d_dist, void *scratch, cudaStream_t stream) {
	if (k < 1 || k > KNN_MAX_K) return cudaErrorInvalidValue;
	if (k <= 8) return run_knn<8>(d_q, n_q, d_c, n_c, stride, js, k, mode, d_idx, d_dist, scratch, stream);
	if (k <= 16) return run_knn<16>(d_q, n_q, d_c, n_c, stride, js, k, mode, d_idx, d_dist, scratch, stream);
	return run_knn<32>(d_q, n_q, d_c, n_c, stride, js, k, mode, d_idx, d_dist, scratch, stream);
}
}  // namespace

size_t knn_scratch_bytes(size_t n_q, size_t n_c, int k) {
	const KnnPlan p = knn_plan(n_q, n_c);
	return ((n_q * p.n_chunks * k * sizeof(int32_t) + 255) & ~(size_t) 255) + n_q * p.n_chunks * k * sizeof(double) + 256;
}

cudaError_t launch_knn(const double *d_states, size_t n, size_t stride, int js, int k, int causal, int32_t *d_idx, double *d_dist, void *scratch,
					   cudaStream_t stream) {
	if (!n) return cudaSuccess;
	return dispatch_knn(d_states, n, d_states, n, stride, js, k, causal ? 1 : 0, d_idx, d_dist, scratch, stream);
}

cudaError_t launch_knn_query(const double *d_database, size_t n_db, const double *d_queries, size_t n_q, size_t stride, int js, int k,
							 int32_t *d_idx, double *d_dist, void *scratch, cudaStream_t stream) {
	if (!n_q) return cudaSuccess;
	return dispatch_knn(d_queries, n_q, d_database, n_db, stride, js, k, 2, d_idx, d_dist, scratch, stream);
}

}  // namespace mgodpl_b200
