// GPU scene builder: triangle soup -> linear BVH (Morton sort + Karras topology) -> bottom-up refit with leaf
// collapsing -> depth-first preorder flattening with skip links.  Stands where meshToFclBVH stood
// (src/planning/fcl_utils.cpp:23-56): same input (the Mesh's vertices and index triples, identity pose), a different
// acceleration structure.  Runs once per scene; everything it produces stays resident in HBM.
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "geometry.cuh"
#include "kernels.h"

namespace mgodpl_b200 {

namespace {

constexpr int TB = 256;

struct BuildStats {
	unsigned long long n_leaves;
	unsigned max_depth, max_leaf;
	double sah;
};

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {  // 21 bits -> every third bit
	v &= 0x1fffffull;
	v = (v | v << 32) & 0x1f00000000ffffull;
	v = (v | v << 16) & 0x1f0000ff0000ffull;
	v = (v | v << 8) & 0x100f00f00f00f00full;
	v = (v | v << 4) & 0x10c30c30c30c30c3ull;
	v = (v | v << 2) & 0x1249249249249249ull;
	return v;
}

/// Per triangle: fp32 bounds relative to `origin` (rounded outward) and the 63-bit Morton key of the bounds centre.
__global__ void k_prepare(const double *__restrict__ verts, const uint32_t *__restrict__ tris, unsigned nt, D3 origin, D3 lo,
						  D3 inv_extent, float4 *__restrict__ tlo, float4 *__restrict__ thi,
						  unsigned long long *__restrict__ keys, uint32_t *__restrict__ ids) {
	unsigned t = blockIdx.x * TB + threadIdx.x;
	if (t >= nt) return;
	double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
	for (int k = 0; k < 3; ++k) {
		const double *v = verts + 3ull * tris[3ull * t + k];
		for (int a = 0; a < 3; ++a) mn[a] = fmin(mn[a], v[a]), mx[a] = fmax(mx[a], v[a]);
	}
	tlo[t] = make_float4(__double2float_rd(mn[0] - origin.x), __double2float_rd(mn[1] - origin.y),
						 __double2float_rd(mn[2] - origin.z), 0.f);
	thi[t] = make_float4(__double2float_ru(mx[0] - origin.x), __double2float_ru(mx[1] - origin.y),
						 __double2float_ru(mx[2] - origin.z), 0.f);
	const double scale = 2097151.0;  // 2^21 - 1
	double cx = fmin(fmax((0.5 * (mn[0] + mx[0]) - lo.x) * inv_extent.x, 0.0), 1.0);
	double cy = fmin(fmax((0.5 * (mn[1] + mx[1]) - lo.y) * inv_extent.y, 0.0), 1.0);
	double cz = fmin(fmax((0.5 * (mn[2] + mx[2]) - lo.z) * inv_extent.z, 0.0), 1.0);
	keys[t] = spread21((unsigned long long) (cx * scale)) << 2 | spread21((unsigned long long) (cy * scale)) << 1 |
			  spread21((unsigned long long) (cz * scale));
	ids[t] = t;
}

// Node ids during construction: internal nodes 0 .. n-2, leaves n-1 .. 2n-2 (leaf j = sorted triangle j).
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
	if (j < 0 || j >= n) return -1;
	unsigned long long a = keys[i], b = keys[j];
	return a == b ? 64 + __clz(i ^ j) : __clzll((long long) (a ^ b));
}

/// Karras 2012: one thread per internal node finds its key range and split.
__global__ void k_topology(const unsigned long long *__restrict__ keys, int n, int *__restrict__ left, int *__restrict__ right,
						   int *__restrict__ parent) {
	int i = blockIdx.x * TB + threadIdx.x;
	if (i >= n - 1) return;
	int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
	int dmin = delta(keys, n, i, i - d);
	int lmax = 2;
	while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t; t >>= 1)
		if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
	int j = i + l * d;
	int dnode = delta(keys, n, i, j);
	int s = 0;
	for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
		if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
		if (t == 1) break;
	}
	int gamma = i + s * d + min(d, 0);
	int lc = min(i, j) == gamma ? gamma + n - 1 : gamma;
	int rc = max(i, j) == gamma + 1 ? gamma + 1 + n - 1 : gamma + 1;
	left[i] = lc, right[i] = rc;
	parent[lc] = i, parent[rc] = i;
	if (i == 0) parent[0] = -1;
}

__device__ __forceinline__ float half_area6(float4 alo, float4 ahi, float4 blo, float4 bhi) {
	float x = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), y = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y),
		  z = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
	return x * y + y * z + z * x;
}
__device__ __forceinline__ float half_area1(float4 lo, float4 hi) {
	float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
	return x * y + y * z + z * x;
}

// ---------------------------------------------------------------------------------------------------------------
// SAH-driven topology: parallel locally-ordered clustering (Meister & Bittner 2018) over the Morton-sorted leaves.
// Every round, each cluster looks `PLOC_RADIUS` neighbours to either side for the partner that minimises the surface
// area of the merged box; mutual nearest neighbours merge; the survivors are compacted in order.  Compared with
// splitting on Morton bits this builds the tree bottom-up by the quantity the traversal cost depends on.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PLOC_RADIUS = 16;

__global__ void k_ploc_init(int n, const uint32_t *__restrict__ ids, const float4 *__restrict__ tlo, const float4 *__restrict__ thi,
							int *__restrict__ cn, float4 *__restrict__ clo, float4 *__restrict__ chi) {
	int j = blockIdx.x * TB + threadIdx.x;
	if (j >= n) return;
	const uint32_t t = ids[j];
	cn[j] = j + n - 1, clo[j] = tlo[t], chi[j] = thi[t];
}

__global__ void k_ploc_nearest(int c, int radius, const float4 *__restrict__ clo, const float4 *__restrict__ chi, int *__restrict__ nn) {
	int i = blockIdx.x * TB + threadIdx.x;
	if (i >= c) return;
	const float4 lo = clo[i], hi = chi[i];
	float best = 3.0e38f;
	int arg = -1;
	const int j0 = max(0, i - radius), j1 = min(c - 1, i + radius);
	for (int j = j0; j <= j1; ++j) {
		if (j == i) continue;
		const float a = half_area6(lo, hi, clo[j], chi[j]);
		if (a < best) best = a, arg = j;  // ties keep the smaller index: deterministic
	}
	nn[i] = arg;
}

__global__ void k_ploc_merge(int c, int *__restrict__ cn, float4 *__restrict__ clo, float4 *__restrict__ chi, const int *__restrict__ nn,
							 int *__restrict__ next_id, int *__restrict__ left, int *__restrict__ right, int *__restrict__ parent,
							 unsigned *__restrict__ keep) {
	int i = blockIdx.x * TB + threadIdx.x;
	if (i >= c) return;
	const int j = nn[i];
	const bool mutual = j >= 0 && nn[j] == i;
	if (mutual && i > j) {
		keep[i] = 0;  // absorbed by the lower-indexed partner
		return;
	}
	keep[i] = 1;
	if (!mutual) return;
	const int id = atomicSub(next_id, 1);  // internal ids count down: the last merge of the build receives id 0 = the root
	const int a = cn[i], b = cn[j];
	left[id] = a, right[id] = b;
	parent[a] = id, parent[b] = id;
	const float4 alo = clo[i], ahi = chi[i], blo2 = clo[j], bhi2 = chi[j];
	cn[i] = id;
	clo[i] = make_float4(fminf(alo.x, blo2.x), fminf(alo.y, blo2.y), fminf(alo.z, blo2.z), 0.f);
	chi[i] = make_float4(fmaxf(ahi.x, bhi2.x), fmaxf(ahi.y, bhi2.y), fmaxf(ahi.z, bhi2.z), 0.f);
}

__global__ void k_ploc_compact(int c, const unsigned *__restrict__ keep, const unsigned *__restrict__ pos, const int *__restrict__ cn,
							   const float4 *__restrict__ clo, const float4 *__restrict__ chi, int *__restrict__ cn2,
							   float4 *__restrict__ clo2, float4 *__restrict__ chi2) {
	int i = blockIdx.x * TB + threadIdx.x;
	if (i >= c || !keep[i]) return;
	const unsigned p = pos[i];
	cn2[p] = cn[i], clo2[p] = clo[i], chi2[p] = chi[i];
}

/// Bottom-up pass over the binary tree (atomic arrival flags: the second child to arrive owns the parent):
///  * bounds, triangle counts and — with leaves collapsed at `max_leaf` triangles — the flattened subtree size;
///  * optionally the SAH refinement: at every internal node, the best of the four child <-> grandchild rotations
///    (Kensler 2008) is applied if it lowers the surface area of the restructured child.  Everything a rotation
///    touches lies inside the node's own, already finished subtree, so the sweep stays race-free.
__global__ void k_refit(int n, const uint32_t *__restrict__ ids, const float4 *__restrict__ tlo, const float4 *__restrict__ thi,
						int *left, int *right, int *parent, float4 *blo, float4 *bhi, unsigned *tri_count, unsigned *eff_size,
						unsigned *__restrict__ flags, unsigned max_leaf, int rotate, int init_leaves) {
	int j = blockIdx.x * TB + threadIdx.x;
	if (j >= n) return;
	int node = j + n - 1;
	if (init_leaves) {
		uint32_t t = ids[j];
		blo[node] = tlo[t], bhi[node] = thi[t];
		tri_count[node] = 1, eff_size[node] = 1;
	}
	__threadfence();
	int p = n > 1 ? __ldcg(parent + node) : -1;
	while (p >= 0) {
		if (atomicAdd(&flags[p], 1u) == 0) return;  // first child to arrive leaves; the second one merges
		__threadfence();
		int l = __ldcg(left + p), r = __ldcg(right + p);
		float4 llo = __ldcg(blo + l), lhi = __ldcg(bhi + l), rlo = __ldcg(blo + r), rhi = __ldcg(bhi + r);
		if (rotate) {
			// candidates: (child kept, child restructured, grandchild promoted, grandchild staying)
			float best = 0.f;
			int kind = -1;
			int rl = -1, rr = -1, ll = -1, lr = -1;
			float4 rllo, rlhi, rrlo, rrhi, lllo, llhi, lrlo, lrhi;
			if (r < n - 1) {
				rl = __ldcg(left + r), rr = __ldcg(right + r);
				rllo = __ldcg(blo + rl), rlhi = __ldcg(bhi + rl), rrlo = __ldcg(blo + rr), rrhi = __ldcg(bhi + rr);
				const float base = half_area1(rlo, rhi);
				float d0 = half_area6(llo, lhi, rrlo, rrhi) - base;  // swap L <-> RL : R' = (L, RR)
				float d1 = half_area6(llo, lhi, rllo, rlhi) - base;  // swap L <-> RR : R' = (RL, L)
				if (d0 < best) best = d0, kind = 0;
				if (d1 < best) best = d1, kind = 1;
			}
			if (l < n - 1) {
				ll = __ldcg(left + l), lr = __ldcg(right + l);
				lllo = __ldcg(blo + ll), llhi = __ldcg(bhi + ll), lrlo = __ldcg(blo + lr), lrhi = __ldcg(bhi + lr);
				const float base = half_area1(llo, lhi);
				float d2 = half_area6(rlo, rhi, lrlo, lrhi) - base;  // swap R <-> LL : L' = (R, LR)
				float d3 = half_area6(rlo, rhi, lllo, llhi) - base;  // swap R <-> LR : L' = (LL, R)
				if (d2 < best) best = d2, kind = 2;
				if (d3 < best) best = d3, kind = 3;
			}
			if (kind >= 0) {
				// restructured child c keeps grandchild `stay` and receives the other child `in`; grandchild `up` moves to p
				const bool right_side = kind < 2;
				const int c = right_side ? r : l, in = right_side ? l : r;
				const int up = kind == 0 ? rl : kind == 1 ? rr : kind == 2 ? ll : lr;
				const int stay = kind == 0 ? rr : kind == 1 ? rl : kind == 2 ? lr : ll;
				left[c] = in, right[c] = stay;
				parent[in] = c, parent[up] = p;
				left[p] = up, right[p] = c;
				const float4 ilo = right_side ? llo : rlo, ihi = right_side ? lhi : rhi;
				const float4 slo = kind == 0 ? rrlo : kind == 1 ? rllo : kind == 2 ? lrlo : lllo;
				const float4 shi = kind == 0 ? rrhi : kind == 1 ? rlhi : kind == 2 ? lrhi : llhi;
				blo[c] = make_float4(fminf(ilo.x, slo.x), fminf(ilo.y, slo.y), fminf(ilo.z, slo.z), 0.f);
				bhi[c] = make_float4(fmaxf(ihi.x, shi.x), fmaxf(ihi.y, shi.y), fmaxf(ihi.z, shi.z), 0.f);
				const unsigned cnt = __ldcg(tri_count + in) + __ldcg(tri_count + stay);
				tri_count[c] = cnt;
				eff_size[c] = cnt <= max_leaf ? 1u : 1u + __ldcg(eff_size + in) + __ldcg(eff_size + stay);
				__threadfence();
				l = up, r = c;
				llo = __ldcg(blo + l), lhi = __ldcg(bhi + l), rlo = __ldcg(blo + r), rhi = __ldcg(bhi + r);
			}
		}
		blo[p] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
		bhi[p] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
		unsigned cnt = __ldcg(tri_count + l) + __ldcg(tri_count + r);
		tri_count[p] = cnt;
		eff_size[p] = cnt <= max_leaf ? 1u : 1u + __ldcg(eff_size + l) + __ldcg(eff_size + r);
		__threadfence();
		p = __ldcg(parent + p);
	}
}

__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
	float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
	return x * y + y * z + z * x;
}

__device__ __forceinline__ void centre_half(float4 lo, float4 hi, float *c, float *h) {
	const float lo3[3] = {lo.x, lo.y, lo.z}, hi3[3] = {hi.x, hi.y, hi.z};
	for (int a = 0; a < 3; ++a) {
		double cd = 0.5 * ((double) lo3[a] + (double) hi3[a]);
		c[a] = (float) cd;
		h[a] = __double2float_ru(0.5 * ((double) hi3[a] - (double) lo3[a]) + fabs(cd - (double) c[a]));  // rounded outward
	}
}

/// One thread per construction node.  Walking to the root yields the node's preorder rank among the surviving internal
/// nodes and its leaf rank (triangles before it), for whatever shape the rotations left the tree in.
///  * every surviving internal node emits one 64-byte traversal record holding BOTH children's boxes and references
///    (see device_types.cuh), so a visit tests two boxes per dependent fetch and leaves never cost a fetch of their own;
///  * every original leaf places its triangle — 9 fp64 world coordinates — at its leaf-order slot, so each (collapsed)
///    leaf owns a contiguous run of slots.
__global__ void k_flatten(int n, const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent,
						  const float4 *__restrict__ blo, const float4 *__restrict__ bhi, const unsigned *__restrict__ tri_count,
						  const unsigned *__restrict__ eff_size, unsigned max_leaf, const double *__restrict__ verts,
						  const uint32_t *__restrict__ tris, const uint32_t *__restrict__ ids, float4 *__restrict__ nodes,
						  double *__restrict__ tri_out, float4 *__restrict__ tri32_out, D3 origin, uint32_t *__restrict__ slot_ids,
						  float4 *__restrict__ root_box, unsigned *__restrict__ even_depth,
						  BuildStats *__restrict__ stats) {
	int x = blockIdx.x * TB + threadIdx.x;
	if (x >= 2 * n - 1) return;
	const int root = 0;  // internal node 0, or — when n == 1 — the single leaf (id n-1 == 0)
	int p = parent[x];   // -1 for the root
	const bool swallowed = p >= 0 && tri_count[p] <= max_leaf;  // inside a collapsed leaf
	const bool original_leaf = x >= n - 1;
	if (swallowed && !original_leaf) return;
	unsigned pre = 0, depth = 0, tri_before = 0;  // pre: internal records before this node in preorder
	for (int c = x; p >= 0; c = p, p = parent[p]) {
		const int l = left[p];
		if (c != l) pre += (eff_size[l] - 1u) / 2u, tri_before += tri_count[l];
		pre += 1u;
		++depth;
	}
	if (original_leaf) {
		const uint32_t t = ids[x - (n - 1)];
		slot_ids[tri_before] = t;
		for (int k = 0; k < 3; ++k) {
			const double *v = verts + 3ull * tris[3ull * t + k];
			tri_out[9ull * tri_before + 3 * k] = v[0], tri_out[9ull * tri_before + 3 * k + 1] = v[1], tri_out[9ull * tri_before + 3 * k + 2] = v[2];
			tri32_out[3ull * tri_before + k] = make_float4((float) (v[0] - origin.x), (float) (v[1] - origin.y), (float) (v[2] - origin.z), 0.f);
		}
	}
	if (swallowed) return;
	const unsigned cnt = tri_count[x];
	const bool leaf = cnt <= max_leaf;
	const float4 lo = blo[x], hi = bhi[x];
	// statistics
	const float root_area = half_area(blo[root], bhi[root]);
	double cost = root_area > 0.f ? (double) half_area(lo, hi) / (double) root_area * (leaf ? (double) cnt : 1.0) : 0.0;
	atomicAdd(&stats->sah, cost);
	if (leaf) {
		atomicAdd(&stats->n_leaves, 1ull);
		atomicMax(&stats->max_leaf, cnt);
		atomicMax(&stats->max_depth, depth);
	}
	if (x == root) root_box[0] = lo, root_box[1] = hi;
	float c0[3], h0[3], c1[3], h1[3];
	int lref, rref;
	if (leaf) {
		if (x != root) return;  // leaves live inside their parent's record
		// the whole scene fits one leaf: a single record whose second child can never overlap
		centre_half(lo, hi, c0, h0);
		lref = -1 - (int) (0u << 4 | (cnt - 1));
		c1[0] = c1[1] = c1[2] = 0.f, h1[0] = h1[1] = h1[2] = -3.0e38f;
		rref = lref;
	} else {
		const int l = left[x], r = right[x];
		const unsigned cl = tri_count[l], cr = tri_count[r];
		centre_half(blo[l], bhi[l], c0, h0);
		centre_half(blo[r], bhi[r], c1, h1);
		lref = cl <= max_leaf ? -1 - (int) (tri_before << 4 | (cl - 1)) : (int) (pre + 1u);
		rref = cr <= max_leaf ? -1 - (int) ((tri_before + cl) << 4 | (cr - 1)) : (int) (pre + 1u + (eff_size[l] - 1u) / 2u);
	}
	even_depth[pre] = (depth & 1u) == 0u;  // these records become the 4-wide records (k_fold_wide)
	float4 *rec = nodes + 4ull * pre;
	rec[0] = make_float4(c0[0], c0[1], c0[2], h0[0]);
	rec[1] = make_float4(h0[1], h0[2], c1[0], c1[1]);
	rec[2] = make_float4(c1[2], h1[0], h1[1], h1[2]);
	rec[3] = make_float4(__int_as_float(lref), __int_as_float(rref), 0.f, 0.f);
}

/// Folds every second level of the binary records into its parent (4-wide records, device_types.cuh): one thread per
/// even-depth binary record; its wide index is its rank among the even-depth records in preorder (an exclusive scan of
/// the even-depth flags), so the wide records are in depth-first preorder too.
__global__ void k_fold_wide(const float4 *__restrict__ nodes, unsigned n_nodes, const unsigned *__restrict__ even_depth,
							const unsigned *__restrict__ wide_index, float4 *__restrict__ wide) {
	const unsigned x = blockIdx.x * TB + threadIdx.x;
	if (x >= n_nodes || !even_depth[x]) return;
	float4 *out = wide + 8ull * wide_index[x];
	int slot = 0;
	auto put = [&](float cx, float cy, float cz, float hx, float hy, float hz, int ref) {
		if (ref >= 0) ref = (int) wide_index[ref];  // an internal grandchild: itself an even-depth record
		out[2 * slot] = make_float4(cx, cy, cz, __int_as_float(ref));
		out[2 * slot + 1] = make_float4(hx, hy, hz, 0.f);
		++slot;
	};
	auto children = [&](unsigned rec, bool expand) {
		const float4 a = nodes[4ull * rec], b = nodes[4ull * rec + 1], c = nodes[4ull * rec + 2], d = nodes[4ull * rec + 3];
		const float box[2][6] = {{a.x, a.y, a.z, a.w, b.x, b.y}, {b.z, b.w, c.x, c.y, c.z, c.w}};
		const int ref[2] = {__float_as_int(d.x), __float_as_int(d.y)};
		for (int k = 0; k < 2; ++k) {
			if (box[k][3] < 0.f) continue;  // the empty second child of a single-leaf scene
			if (expand && ref[k] >= 0) {
				// an internal child sits at odd depth: it is absorbed, its two children become children of this record
				const unsigned r2 = (unsigned) ref[k];
				const float4 a2 = nodes[4ull * r2], b2 = nodes[4ull * r2 + 1], c2 = nodes[4ull * r2 + 2], d2 = nodes[4ull * r2 + 3];
				put(a2.x, a2.y, a2.z, a2.w, b2.x, b2.y, __float_as_int(d2.x));
				put(b2.z, b2.w, c2.x, c2.y, c2.z, c2.w, __float_as_int(d2.y));
			} else {
				put(box[k][0], box[k][1], box[k][2], box[k][3], box[k][4], box[k][5], ref[k]);
			}
		}
	};
	children(x, true);
	for (; slot < 4; ++slot) {
		out[2 * slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(WIDE_EMPTY));
		out[2 * slot + 1] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
	}
}


// ---------------------------------------------------------------------------------------------------------------
// Clearance grid (SceneDev::clear): one thread per cell runs a nearest-triangle query from the cell centre over the
// binary records (nearest child first, subtrees beyond the best distance so far pruned), in fp32 on origin-relative
// coordinates.  Everything errs towards SMALLER clearances: the triangle distance is the minimum over the three edge
// segments and (where the projection falls inside, judged generously) the plane — each the distance to some point of
// the triangle or less; then the cell's half diagonal (the distance is 1-Lipschitz) and a safety term are subtracted
// and the value is rounded down to the u8 quantum.
// ---------------------------------------------------------------------------------------------------------------
struct F3 {
	float x, y, z;
};
__device__ __forceinline__ F3 f3sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float f3dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float seg_dist2(F3 p, F3 a, F3 b) {
	const F3 ab = f3sub(b, a), ap = f3sub(p, a);
	const float l2 = f3dot(ab, ab);
	float t = l2 > 0.f ? f3dot(ap, ab) / l2 : 0.f;
	t = fminf(fmaxf(t, 0.f), 1.f);
	const F3 d{ap.x - t * ab.x, ap.y - t * ab.y, ap.z - t * ab.z};
	return f3dot(d, d);
}
__device__ __forceinline__ float tri_dist2_lower(F3 p, F3 a, F3 b, F3 c) {
	float d2 = fminf(seg_dist2(p, a, b), fminf(seg_dist2(p, b, c), seg_dist2(p, c, a)));
	const F3 ab = f3sub(b, a), ac = f3sub(c, a), ap = f3sub(p, a);
	const F3 n{ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
	const float n2 = f3dot(n, n);
	if (n2 > 0.f) {
		// barycentric sign tests of the projection, with slack: a projection wrongly taken for "inside" can only lower the result
		const F3 bp = f3sub(p, b), cp = f3sub(p, c), bc = f3sub(c, b), ca = f3sub(a, c);
		auto side = [&](F3 e, F3 q) { return (e.y * q.z - e.z * q.y) * n.x + (e.z * q.x - e.x * q.z) * n.y + (e.x * q.y - e.y * q.x) * n.z; };
		const float slack = -1e-3f * n2;
		if (side(ab, ap) >= slack && side(bc, bp) >= slack && side(ca, cp) >= slack) {
			const float h = f3dot(ap, n);
			d2 = fminf(d2, h * h / n2);
		}
	}
	return d2;
}
__device__ __forceinline__ float box_dist2(F3 p, float cx, float cy, float cz, float hx, float hy, float hz) {
	const float dx = fmaxf(fabsf(p.x - cx) - hx, 0.f), dy = fmaxf(fabsf(p.y - cy) - hy, 0.f), dz = fmaxf(fabsf(p.z - cz) - hz, 0.f);
	return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(TB)
k_clearance_grid(const float4 *__restrict__ nodes, const float4 *__restrict__ tris32, int nx, int ny, int nz, F3 lo, float cell,
				 float cmax, float quantum, unsigned char *__restrict__ out) {
	const size_t i = (size_t) blockIdx.x * TB + threadIdx.x;
	if (i >= (size_t) nx * ny * nz) return;
	const int ix = (int) (i % nx), iy = (int) ((i / nx) % ny), iz = (int) (i / ((size_t) nx * ny));
	const F3 p{lo.x + (ix + 0.5f) * cell, lo.y + (iy + 0.5f) * cell, lo.z + (iz + 0.5f) * cell};
	const float half_diag = 0.8660255f * cell, safety = 2e-4f + 1e-5f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z));
	const float cutoff = cmax + half_diag + safety;
	float best2 = cutoff * cutoff;
	int stack[64];
	float stack_d2[64];
	int sp = 0, cur = 0;
	while (true) {
		if (cur >= 0) {
			const float4 *rec = nodes + 4ull * (unsigned) cur;
			const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1), n2 = __ldg(rec + 2), n3 = __ldg(rec + 3);
			const float dl = box_dist2(p, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y), dr = box_dist2(p, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
			const int lref = __float_as_int(n3.x), rref = __float_as_int(n3.y);
			const bool tl = dl < best2, tr = dr < best2;
			if (tl && tr) {
				const bool right_first = dr < dl;
				stack[sp] = right_first ? lref : rref, stack_d2[sp] = right_first ? dl : dr, ++sp;
				cur = right_first ? rref : lref;
				continue;
			}
			if (tl || tr) {
				cur = tl ? lref : rref;
				continue;
			}
		} else {
			const unsigned info = (unsigned) (-1 - cur), first = info >> 4, cnt = (info & 15u) + 1u;
			for (unsigned k = 0; k < cnt; ++k) {
				const float4 a = __ldg(tris32 + 3ull * (first + k)), b = __ldg(tris32 + 3ull * (first + k) + 1), c = __ldg(tris32 + 3ull * (first + k) + 2);
				best2 = fminf(best2, tri_dist2_lower(p, F3{a.x, a.y, a.z}, F3{b.x, b.y, b.z}, F3{c.x, c.y, c.z}));
			}
		}
		bool found = false;
		while (sp > 0) {
			--sp;
			if (stack_d2[sp] < best2) {
				cur = stack[sp], found = true;
				break;
			}
		}
		if (!found) break;
	}
	const float d = fminf(fmaxf(sqrtf(best2) - half_diag - safety, 0.f), cmax);
	out[i] = (unsigned char) fminf(floorf(d / quantum), 255.f);
}

struct Scratch {
	std::vector<void *> ptrs;
	template<class T>
	cudaError_t alloc(T **p, size_t n) {
		cudaError_t e = cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T));
		if (e == cudaSuccess) ptrs.push_back(*p);
		return e;
	}
	~Scratch() {
		for (void *p: ptrs) cudaFree(p);
	}
};

}  // namespace

#define BS_CHECK(x)                       \
	do {                                  \
		cudaError_t e_ = (x);             \
		if (e_ != cudaSuccess) return e_; \
	} while (0)

cudaError_t build_scene(const double *d_verts, size_t nv, const uint32_t *d_tris, size_t nt, const double *lo, const double *hi,
						const double *origin, unsigned max_leaf_size, bool refine, bool wide, bool clearance_grid, BuiltScene *out,
						cudaStream_t stream) {
	(void) nv;
	(void) refine;
	*out = BuiltScene{};
	if (nt == 0) return cudaSuccess;
	if (max_leaf_size < 1) max_leaf_size = 1;
	if (max_leaf_size > 16) max_leaf_size = 16;
	const int n = (int) nt;
	const unsigned gt = (unsigned) ((nt + TB - 1) / TB);
	Scratch tmp;
	float4 *tlo, *thi, *blo, *bhi;
	unsigned long long *keys, *keys2;
	uint32_t *ids, *ids0, *ids2;  // final slot -> triangle (kept) | unsorted | Morton-sorted
	int *left, *right, *parent;
	unsigned *tri_count, *eff_size, *flags;
	BuildStats *d_stats;
	BS_CHECK(tmp.alloc(&tlo, nt));
	BS_CHECK(tmp.alloc(&thi, nt));
	BS_CHECK(tmp.alloc(&keys, nt));
	BS_CHECK(tmp.alloc(&keys2, nt));
	BS_CHECK(tmp.alloc(&ids0, nt));
	BS_CHECK(tmp.alloc(&ids2, nt));
	BS_CHECK(tmp.alloc(&left, nt));
	BS_CHECK(tmp.alloc(&right, nt));
	BS_CHECK(tmp.alloc(&parent, 2 * nt));
	BS_CHECK(tmp.alloc(&blo, 2 * nt));
	BS_CHECK(tmp.alloc(&bhi, 2 * nt));
	BS_CHECK(tmp.alloc(&tri_count, 2 * nt));
	BS_CHECK(tmp.alloc(&eff_size, 2 * nt));
	BS_CHECK(tmp.alloc(&flags, nt));
	BS_CHECK(tmp.alloc(&d_stats, 1));
	BS_CHECK(cudaMalloc((void **) &ids, nt * sizeof(uint32_t)));  // kept: slot -> original triangle
	out->tri_ids = ids;

	struct Events {  // destroyed on every exit path
		cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
		~Events() {
			for (cudaEvent_t x: e)
				if (x) cudaEventDestroy(x);
		}
	} evs;
	BS_CHECK(cudaEventCreate(&evs.e[0]));
	BS_CHECK(cudaEventCreate(&evs.e[1]));
	cudaEvent_t &ev0 = evs.e[0], &ev1 = evs.e[1], &ev2 = evs.e[2];
	BS_CHECK(cudaEventRecord(ev0, stream));

	D3 inv{hi[0] > lo[0] ? 1.0 / (hi[0] - lo[0]) : 0.0, hi[1] > lo[1] ? 1.0 / (hi[1] - lo[1]) : 0.0,
		   hi[2] > lo[2] ? 1.0 / (hi[2] - lo[2]) : 0.0};
	k_prepare<<<gt, TB, 0, stream>>>(d_verts, d_tris, (unsigned) nt, D3{origin[0], origin[1], origin[2]},
									 D3{lo[0], lo[1], lo[2]}, inv, tlo, thi, keys, ids0);
	BS_CHECK(cudaGetLastError());
	{
		size_t bytes = 0;
		BS_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys2, ids0, ids2, n, 0, 63, stream));
		void *ws;
		BS_CHECK(tmp.alloc((char **) &ws, bytes));
		BS_CHECK(cub::DeviceRadixSort::SortPairs(ws, bytes, keys, keys2, ids0, ids2, n, 0, 63, stream));
	}
	BS_CHECK(cudaMemsetAsync(flags, 0, nt * sizeof(unsigned), stream));
	BS_CHECK(cudaMemsetAsync(d_stats, 0, sizeof(BuildStats), stream));
	BS_CHECK(cudaMemsetAsync(parent, 0xff, 2 * nt * sizeof(int), stream));
	if (n > 1 && !refine) {
		k_topology<<<gt, TB, 0, stream>>>(keys2, n, left, right, parent);  // plain LBVH: split on Morton bits
		BS_CHECK(cudaGetLastError());
	} else if (n > 1) {
		int *cn[2], *nn, *d_next;
		float4 *clo[2], *chi[2];
		unsigned *keep, *pos;
		for (int b = 0; b < 2; ++b) {
			BS_CHECK(tmp.alloc(&cn[b], nt));
			BS_CHECK(tmp.alloc(&clo[b], nt));
			BS_CHECK(tmp.alloc(&chi[b], nt));
		}
		BS_CHECK(tmp.alloc(&nn, nt));
		BS_CHECK(tmp.alloc(&keep, nt + 1));
		BS_CHECK(tmp.alloc(&pos, nt + 1));
		BS_CHECK(tmp.alloc(&d_next, 1));
		size_t scan_bytes = 0;
		BS_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, pos, n + 1, stream));
		void *scan_ws;
		BS_CHECK(tmp.alloc((char **) &scan_ws, scan_bytes));
		const int first_id = n - 2;
		BS_CHECK(cudaMemcpyAsync(d_next, &first_id, sizeof(int), cudaMemcpyHostToDevice, stream));
		k_ploc_init<<<gt, TB, 0, stream>>>(n, ids2, tlo, thi, cn[0], clo[0], chi[0]);
		BS_CHECK(cudaGetLastError());
		static const int ploc_radius = [] {
			const char *e = getenv("MGODPL_PLOC_RADIUS");
			return e && atoi(e) > 0 ? atoi(e) : PLOC_RADIUS;
		}();
		int c = n, cur = 0;
		while (c > 1) {
			const unsigned g = (unsigned) ((c + TB - 1) / TB);
			k_ploc_nearest<<<g, TB, 0, stream>>>(c, ploc_radius, clo[cur], chi[cur], nn);
			k_ploc_merge<<<g, TB, 0, stream>>>(c, cn[cur], clo[cur], chi[cur], nn, d_next, left, right, parent, keep);
			BS_CHECK(cudaGetLastError());
			BS_CHECK(cudaMemsetAsync(keep + c, 0, sizeof(unsigned), stream));
			BS_CHECK(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, keep, pos, c + 1, stream));
			k_ploc_compact<<<g, TB, 0, stream>>>(c, keep, pos, cn[cur], clo[cur], chi[cur], cn[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1]);
			BS_CHECK(cudaGetLastError());
			unsigned c_new = 0;
			BS_CHECK(cudaMemcpyAsync(&c_new, pos + c, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
			BS_CHECK(cudaStreamSynchronize(stream));
			if ((int) c_new >= c) return cudaErrorUnknown;  // every round merges at least the globally closest pair
			c = (int) c_new, cur ^= 1;
		}
	}
	// Pass 0 fits the bounds; with refinement on, it and two more sweeps also apply SAH-lowering rotations.
	const int sweeps = refine && n > 2 ? 3 : 1;
	for (int sweep = 0; sweep < sweeps; ++sweep) {
		if (sweep) BS_CHECK(cudaMemsetAsync(flags, 0, nt * sizeof(unsigned), stream));
		k_refit<<<gt, TB, 0, stream>>>(n, ids2, tlo, thi, left, right, parent, blo, bhi, tri_count, eff_size, flags, max_leaf_size,
									   refine && n > 2, sweep == 0);
		BS_CHECK(cudaGetLastError());
	}
	// The flattened size is the root's effective size; fetch it before allocating the node array.
	unsigned root_size = 0;
	BS_CHECK(cudaMemcpyAsync(&root_size, eff_size, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
	BS_CHECK(cudaStreamSynchronize(stream));
	const unsigned n_nodes = root_size > 1 ? (root_size - 1) / 2 : 1;  // one 64-byte record per surviving internal node
	float4 *d_root_box;
	BS_CHECK(tmp.alloc(&d_root_box, 2));
	unsigned *even_depth, *wide_index;
	BS_CHECK(tmp.alloc(&even_depth, (size_t) n_nodes + 1));
	BS_CHECK(tmp.alloc(&wide_index, (size_t) n_nodes + 1));
	BS_CHECK(cudaMemsetAsync(even_depth, 0, ((size_t) n_nodes + 1) * sizeof(unsigned), stream));
	BS_CHECK(cudaMalloc((void **) &out->nodes, (size_t) n_nodes * 4 * sizeof(float4)));
	BS_CHECK(cudaMalloc((void **) &out->tris, nt * 9 * sizeof(double)));
	BS_CHECK(cudaMalloc((void **) &out->tris32, nt * 3 * sizeof(float4)));
	k_flatten<<<(unsigned) ((2 * nt - 1 + TB - 1) / TB), TB, 0, stream>>>(n, left, right, parent, blo, bhi, tri_count, eff_size,
																		  max_leaf_size, d_verts, d_tris, ids2, out->nodes, out->tris, out->tris32,
																		  D3{origin[0], origin[1], origin[2]}, ids, d_root_box, even_depth, d_stats);
	BS_CHECK(cudaGetLastError());
	unsigned n_wide = 0;
	if (wide) {
		size_t scan_bytes = 0;
		BS_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, even_depth, wide_index, n_nodes + 1, stream));
		void *scan_ws;
		BS_CHECK(tmp.alloc((char **) &scan_ws, scan_bytes));
		BS_CHECK(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, even_depth, wide_index, n_nodes + 1, stream));
		BS_CHECK(cudaMemcpyAsync(&n_wide, wide_index + n_nodes, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		BS_CHECK(cudaStreamSynchronize(stream));
		BS_CHECK(cudaMalloc((void **) &out->wide, (size_t) n_wide * 8 * sizeof(float4)));
		k_fold_wide<<<(n_nodes + TB - 1) / TB, TB, 0, stream>>>(out->nodes, n_nodes, even_depth, wide_index, out->wide);
		BS_CHECK(cudaGetLastError());
		BS_CHECK(cudaMemcpyAsync(out->top, out->wide, sizeof(out->top), cudaMemcpyDeviceToHost, stream));
	}
	float4 h_root[2];
	BS_CHECK(cudaMemcpyAsync(h_root, d_root_box, sizeof(h_root), cudaMemcpyDeviceToHost, stream));
	BS_CHECK(cudaEventRecord(ev1, stream));
	if (clearance_grid) {
		// grid over the scene bounds grown by CLEAR_PAD; cell size from a cell budget, never finer than 2 cm
		const double pad = CLEAR_PAD;
		double ext[3], vol = 1.0;
		for (int a = 0; a < 3; ++a) ext[a] = (hi[a] - lo[a]) + 2.0 * pad, vol *= ext[a];
		static const size_t max_cells = [] {  // initialised once, thread-safe
			const char *e = getenv("MGODPL_CLEARANCE_CELLS");
			const size_t v = e && atoll(e) > 0 ? (size_t) atoll(e) : (size_t) 6 << 20;
			return std::min<size_t>(v, (size_t) 1 << 30);  // cell indices are 32-bit on the device
		}();
		double cell = std::max(0.02, std::cbrt(vol / (double) max_cells));
		int dims[3];
		for (int a = 0; a < 3; ++a) dims[a] = std::max(1, (int) std::ceil(ext[a] / cell));
		const size_t cells = (size_t) dims[0] * dims[1] * dims[2];
		BS_CHECK(cudaMalloc((void **) &out->clear, cells));
		for (int a = 0; a < 3; ++a) out->clear_n[a] = dims[a], out->clear_lo[a] = (float) (lo[a] - pad - origin[a]);
		out->clear_cell = (float) cell, out->clear_max = (float) pad, out->clear_quantum = (float) (pad / 255.0);
		k_clearance_grid<<<(unsigned) ((cells + TB - 1) / TB), TB, 0, stream>>>(out->nodes, out->tris32, dims[0], dims[1], dims[2],
																				   F3{out->clear_lo[0], out->clear_lo[1], out->clear_lo[2]}, out->clear_cell,
																				   out->clear_max, out->clear_quantum, out->clear);
		BS_CHECK(cudaGetLastError());
		BS_CHECK(cudaEventCreate(&ev2));
		BS_CHECK(cudaEventRecord(ev2, stream));
		out->device_bytes_extra = cells;
	}
	BuildStats st{};
	BS_CHECK(cudaMemcpyAsync(&st, d_stats, sizeof(st), cudaMemcpyDeviceToHost, stream));
	BS_CHECK(cudaStreamSynchronize(stream));
	float ms = 0;
	cudaEventElapsedTime(&ms, ev0, ev1);
	if (ev2) {
		float ms2 = 0;
		cudaEventElapsedTime(&ms2, ev1, ev2);
		out->clear_build_ms = ms2;
	}
	out->n_nodes = n_nodes;
	out->n_tris = (uint32_t) nt;
	out->n_leaves = (uint32_t) st.n_leaves;
	out->max_depth = st.max_depth;
	out->max_leaf = st.max_leaf;
	out->sah_cost = st.sah;
	out->build_ms = ms;
	out->n_wide = n_wide;
	out->device_bytes = (size_t) n_nodes * 64 + (size_t) n_wide * 128 + nt * 72 + nt * 48 + nt * 4 + out->device_bytes_extra;
	float rc[3], rh[3];
	{
		// same outward rounding as the device (centre in fp32, half extent padded by the centre's rounding error)
		const float lo3[3] = {h_root[0].x, h_root[0].y, h_root[0].z}, hi3[3] = {h_root[1].x, h_root[1].y, h_root[1].z};
		for (int a = 0; a < 3; ++a) {
			double cd = 0.5 * ((double) lo3[a] + (double) hi3[a]);
			rc[a] = (float) cd;
			rh[a] = std::nextafter((float) (0.5 * ((double) hi3[a] - (double) lo3[a]) + std::fabs(cd - (double) rc[a])), 3.0e38f);
		}
	}
	for (int a = 0; a < 3; ++a) out->root_c[a] = rc[a], out->root_h[a] = rh[a];
	return cudaSuccess;
}

}  // namespace mgodpl_b200
