// The collision-query kernels: batched forward kinematics, box-vs-BVH traversal, motion validation, goal sampling.
//
// Every query runs as a two-stage pipeline on one stream (DESIGN.md "Kernels"):
//
//   stage A  "expand"   one thread per candidate robot state (an explicit state, one interpolation step of a motion,
//                       or one goal sample).  Cull the whole robot against the scene bounds with its reach sphere
//                       (translation only), then — for the few that survive — interpolate / run FK in fp64, cull each
//                       link box against the scene bounds and append the survivors (fp64 box pose + owner) to a
//                       work queue in HBM, one warp-aggregated atomic per warp.
//   stage B  "traverse" one thread per queued box, full warps: stackless skip-link traversal of the flattened BVH,
//                       fp32 conservative node tests, exact fp64 box/triangle tests at the leaves, first hit wins.
//                       Hits are reported with atomicOr (state bitmasks) or atomicMin (first colliding motion step).
//   stage C  "finish"   motions only: first colliding step -> hit bit + toi.
//
// Far-away states (the majority under PRM-style sampling) cost a 24-byte read and a handful of flops in stage A;
// boxes that do need a tree descent are compacted, so stage B runs converged and at full occupancy.  If the queue
// would overflow, stage A traverses the box on the spot — capacity is a performance knob, never a correctness one.
#include <algorithm>
#include <cooperative_groups.h>
#include <cstdlib>

#include "geometry.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace mgodpl_b200 {

constexpr int BLOCK_A = 256;
constexpr int BLOCK_C = 128;  // stage A2 for chain robots: 5 blocks x 4 warps per SM at <= 102 registers
constexpr int BLOCK_B = 128;
constexpr unsigned NO_HIT = 0xFFFFFFFFu;
constexpr int ROW_MAX = 7 + MGODPL_MAX_JOINT_VALUES;

// ---------------------------------------------------------------------------------------------------------------
// Work queue: 64-byte items = the box's exact fp64 pose (centre + quaternion) and its owner.  Stage B derives the
// fp32 culling box from it when a lane picks the item up (a dozen conversions and ~25 fp32 flops) and re-reads the
// same 64 bytes in the rare case a leaf needs the exact test; half the queue traffic of carrying both forms.
// ---------------------------------------------------------------------------------------------------------------
struct __align__(16) Item {
	double c[3];
	double q[4];
	unsigned owner;
	unsigned step_box;  // step | box << 28
};
static_assert(sizeof(Item) == 64, "queue item must be 64 bytes");

enum ResultMode : int { RESULT_BITS = 0, RESULT_FIRST_STEP = 1 };

struct Results {
	int mode;
	uint32_t *hit_bits;      // RESULT_BITS: bit `owner`
	unsigned *first_step;    // RESULT_FIRST_STEP: min colliding step of motion `owner`
};

struct LocalCounters {
	unsigned states = 0, boxes = 0, nodes = 0, leaf = 0;
};

__device__ __forceinline__ void flush_counters(const SceneDev &sc, const LocalCounters &lc) {
	if (!sc.counters) return;
	unsigned v[4] = {lc.states, lc.boxes, lc.nodes, lc.leaf};
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		unsigned s = v[k];
		for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if ((threadIdx.x & 31) == 0 && s) atomicAdd(sc.counters + k, (unsigned long long) s);
	}
}

/// The same from a single lane in divergent code (rare paths only).
__device__ __forceinline__ void flush_counters_lane(const SceneDev &sc, const LocalCounters &lc) {
	if (!sc.counters) return;
	if (lc.states) atomicAdd(sc.counters + 0, (unsigned long long) lc.states);
	if (lc.boxes) atomicAdd(sc.counters + 1, (unsigned long long) lc.boxes);
	if (lc.nodes) atomicAdd(sc.counters + 2, (unsigned long long) lc.nodes);
	if (lc.leaf) atomicAdd(sc.counters + 3, (unsigned long long) lc.leaf);
}

__device__ __forceinline__ bool already_decided(const Results &res, unsigned owner, unsigned step) {
	if (res.mode == RESULT_BITS) return (__ldcg(res.hit_bits + (owner >> 5)) >> (owner & 31)) & 1u;
	return __ldcg(res.first_step + owner) <= step;  // an earlier (or this) step already collides
}
__device__ __forceinline__ void report_hit(const Results &res, unsigned owner, unsigned step) {
	if (res.mode == RESULT_BITS) atomicOr(res.hit_bits + (owner >> 5), 1u << (owner & 31));
	else atomicMin(res.first_step + owner, step);
}

__device__ __forceinline__ ExactBox make_exact_box(D3 c, Q4 q, const double *half) {
	ExactBox eb;
	eb.cx = c.x, eb.cy = c.y, eb.cz = c.z;
	qmat(q, eb.r);
	eb.hx = half[0], eb.hy = half[1], eb.hz = half[2];
	return eb;
}

/// fp32 classification of a leaf's triangles against one box: returns TRI_HIT as soon as one triangle certainly
/// touches, otherwise the set of triangles that still need the exact test (bit k of *unknown).
/// `is_box` false (capsule / convex link: `cb` is the shape's bounding box): a vertex inside the bounding box proves nothing,
/// so "certain hit" is demoted to "needs the exact test".
__device__ __forceinline__ bool leaf_prefilter(const SceneDev &sc, const CullBox &cb, int ref, unsigned *unknown, bool is_box = true) {
	const unsigned info = (unsigned) (-1 - ref), first = info >> 4, cnt = (info & 15u) + 1u;
	*unknown = 0;
	for (unsigned k = 0; k < cnt; ++k) {
		const float4 *t = sc.tris32 + 3ull * (first + k);
		const int v = tri_classify(cb, sc.margin, __ldg(t), __ldg(t + 1), __ldg(t + 2));
		if (v == TRI_HIT && is_box) return true;
		if (v != TRI_SEPARATED) *unknown |= 1u << k;
	}
	return false;
}

/// The same for a capsule / convex link (`cb` is the shape's bounding box), out of line so that box robots pay nothing for it.
/// "Certainly separated" carries over from the bounding box; beyond that a capsule gets an fp32 copy of its exact test (the
/// axis-segment / triangle distance against the radius, with error bars: capsule_triangle_classify), a polytope a "certain
/// hit" when a triangle vertex lies behind every face plane shrunk by the margin.
/// (Without it every contact went to the fp64 test at one active lane: 10 M states took 2.7 ms instead of 0.57 for boxes.)
__device__ __noinline__ bool leaf_prefilter_shape(const SceneDev &sc, const CullBox &cb, int ref, unsigned *unknown, const BoxDev *shape,
												 const double *shape_data_base) {
	const unsigned info = (unsigned) (-1 - ref), first = info >> 4, cnt = (info & 15u) + 1u;
	const float slack = 2.f * sc.margin + 1e-5f;
	const bool capsule = shape->kind == SHAPE_KIND_CAPSULE;
	// capsule: true radius = inflated hx - margin, the axis segment reaches +- (hz - hx) along the local z axis
	const float half_len = cb.hz - cb.hx;
	const double *hull = capsule ? nullptr : shape_data_base + shape->data_off;
	const int n_faces = capsule ? 0 : (int) __ldg(hull + 1);
	const double *faces = capsule ? nullptr : hull + 4 + 3 * (int) __ldg(hull);
	*unknown = 0;
	for (unsigned k = 0; k < cnt; ++k) {
		const float4 *t = sc.tris32 + 3ull * (first + k);
		const float4 v[3] = {__ldg(t), __ldg(t + 1), __ldg(t + 2)};
		if (tri_classify(cb, sc.margin, v[0], v[1], v[2]) == TRI_SEPARATED) continue;
		bool inside = false;
		if (capsule) {
			V3f p[3];
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float dx = v[c].x - cb.cx, dy = v[c].y - cb.cy, dz = v[c].z - cb.cz;
				p[c] = V3f{cb.r[0] * dx + cb.r[3] * dy + cb.r[6] * dz, cb.r[1] * dx + cb.r[4] * dy + cb.r[7] * dz, cb.r[2] * dx + cb.r[5] * dy + cb.r[8] * dz};
			}
			const int verdict = capsule_triangle_classify(cb.hx - sc.margin, half_len, slack, p[0], p[1], p[2]);
			if (verdict == TRI_SEPARATED) continue;
			inside = verdict == TRI_HIT;
		} else {
			// polytope: one pass over its face planes settles both easy cases — all three vertices beyond one plane (separated), or
			// one vertex behind every plane (hit); everything else goes to the exact test
			float px[3], py[3], pz[3];
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float dx = v[c].x - cb.cx, dy = v[c].y - cb.cy, dz = v[c].z - cb.cz;
				px[c] = cb.r[0] * dx + cb.r[3] * dy + cb.r[6] * dz, py[c] = cb.r[1] * dx + cb.r[4] * dy + cb.r[7] * dz, pz[c] = cb.r[2] * dx + cb.r[5] * dy + cb.r[8] * dz;
			}
			bool in0 = true, in1 = true, in2 = true, separated = false;
			for (int f = 0; f < n_faces && !separated; ++f) {
				const float nx = (float) __ldg(faces + 4 * f), ny = (float) __ldg(faces + 4 * f + 1), nz = (float) __ldg(faces + 4 * f + 2), off = (float) __ldg(faces + 4 * f + 3);
				const float s0 = nx * px[0] + ny * py[0] + nz * pz[0] - off, s1 = nx * px[1] + ny * py[1] + nz * pz[1] - off, s2 = nx * px[2] + ny * py[2] + nz * pz[2] - off;
				separated = fminf(s0, fminf(s1, s2)) > slack;
				in0 &= s0 < -slack, in1 &= s1 < -slack, in2 &= s2 < -slack;
			}
			if (separated) continue;
			inside = in0 | in1 | in2;
		}
		if (inside) return true;
		*unknown |= 1u << k;
	}
	return false;
}

/// The exact phase of one parked lane, kept out of line: re-read the fp64 pose of the queued box, build the exact box
/// and test the triangles the pre-filter could not decide.  Rare and register-hungry (fp64): as a separate function its
/// register demand does not shape the traversal loop.
__device__ __noinline__ bool exact_leaf_test(const SceneDev &sc, const void *exact_pose, const double *half, int ref, unsigned unknown,
											 unsigned *n_tests, const BoxDev *shape = nullptr, const double *shape_data_base = nullptr) {
	const int kind = shape ? shape->kind : (int) SHAPE_KIND_BOX;
	const double *shape_data = kind == SHAPE_KIND_CONVEX ? shape_data_base + shape->data_off : nullptr;
	const unsigned info = (unsigned) (-1 - ref), first = info >> 4, cnt = (info & 15u) + 1u;
	const double2 *ex = reinterpret_cast<const double2 *>(exact_pose);
	const double2 e0 = __ldcg(ex), e1 = __ldcg(ex + 1), e2 = __ldcg(ex + 2), e3 = __ldcg(ex + 3);
	const ExactBox eb = make_exact_box(D3{e0.x, e0.y, e1.x}, Q4{e1.y, e2.x, e2.y, e3.x}, half);
	for (unsigned k = 0; k < cnt; ++k) {
		if (!(unknown >> k & 1u)) continue;
		++*n_tests;
		if (shape_triangle_hit(kind, eb, shape_data, sc.tris + 9ull * (first + k))) return true;
	}
	return false;
}

constexpr int STACK_DEPTH = 52;        // scene creation rejects trees deeper than 48 levels
constexpr int WIDE_STACK_DEPTH = 72;   // 4-wide records: at most 24 levels, up to three pushes each
constexpr int REF_END = (int) 0x80000000;  // "nothing left to visit"

/// Test both children of record `cur`; returns the next reference to visit (pushing the other overlapping child).
struct LocalStack {
	int *data;
	int sp;
	__device__ __forceinline__ void push(int v) { data[sp++] = v; }
	__device__ __forceinline__ int pop() { return data[--sp]; }
};
template<class Stack>
__device__ __forceinline__ int visit_record(const SceneDev &sc, const CullBox &cb, int cur, Stack &st) {
	const float4 *rec = sc.nodes + 4ull * (unsigned) cur;
	const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1), n2 = __ldg(rec + 2), n3 = __ldg(rec + 3);
	const bool ov_l = cull_overlap(cb, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
	const bool ov_r = cull_overlap(cb, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
	const int lref = __float_as_int(n3.x), rref = __float_as_int(n3.y);
	int next = REF_END;
	if (ov_l && ov_r) {
		// any-hit query: walk into the child whose centre is nearer to the box first — a colliding box finds its hit sooner
		const float lx = n0.x - cb.cx, ly = n0.y - cb.cy, lz = n0.z - cb.cz, rx = n1.z - cb.cx, ry = n1.w - cb.cy, rz = n2.x - cb.cz;
		const bool right_first = rx * rx + ry * ry + rz * rz < lx * lx + ly * ly + lz * lz;
		next = right_first ? rref : lref;
		st.push(right_first ? lref : rref);
	} else if (ov_l) {
		next = lref;
	} else if (ov_r) {
		next = rref;
	} else if (st.sp > 0) {
		next = st.pop();
	}
	return next;
}

/// Per-lane traversal stack: the first STACK_SHARED entries live in shared memory ([depth][thread]: conflict-free), only
/// deeper ones in a local array (trees of ordinary depth never get there).
constexpr int STACK_SHARED = 16;
struct LaneStack {
	int *shared;  // &s_stack[0][threadIdx.x]; entries are blockDim.x ints apart
	int *local;
	int stride, sp;
	// The shared array has one spare row: entries beyond STACK_SHARED are (also) written there, so the store needs no branch;
	// the local copy is what counts for them.
	__device__ __forceinline__ void push(int v) {
		shared[min(sp, STACK_SHARED) * stride] = v;
		if (sp >= STACK_SHARED) local[sp - STACK_SHARED] = v;
		++sp;
	}
	__device__ __forceinline__ int pop() {
		--sp;
		int v = shared[min(sp, STACK_SHARED) * stride];
		if (sp >= STACK_SHARED) v = local[sp - STACK_SHARED];
		return v;
	}
};

/// The same for a 4-wide record: up to four boxes per dependent fetch; the nearest overlapping child is visited next, the
/// others are pushed.  Written without per-child branches: the four overlap tests, distances and the nearest-child selection
/// are straight-line code for every lane of the warp, only the (short) pushes are predicated.
template<class Stack>
__device__ __forceinline__ int visit_wide(const SceneDev &sc, const CullBox &cb, int cur, Stack &st) {
	const float4 *rec = sc.wide + 8ull * (unsigned) cur;
	float4 a[4], b[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) a[k] = __ldg(rec + 2 * k), b[k] = __ldg(rec + 2 * k + 1);
	int ref[4];
	float d[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		ref[k] = __float_as_int(a[k].w);
		const bool ov = (ref[k] != WIDE_EMPTY) & cull_overlap(cb, a[k].x, a[k].y, a[k].z, b[k].x, b[k].y, b[k].z);
		const float dx = a[k].x - cb.cx, dy = a[k].y - cb.cy, dz = a[k].z - cb.cz;
		d[k] = ov ? dx * dx + dy * dy + dz * dz : __int_as_float(0x7f800000);
	}
	int next = REF_END;
	float next_d = __int_as_float(0x7f800000);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const bool nearer = d[k] < next_d;
		next = nearer ? ref[k] : next, next_d = nearer ? d[k] : next_d;
	}
#pragma unroll
	for (int k = 0; k < 4; ++k)
		if (d[k] < __int_as_float(0x7f800000) && ref[k] != next) st.push(ref[k]);
	if (next == REF_END && st.sp > 0) next = st.pop();
	return next;
}

/// Exact test of the box against every triangle of one leaf reference.
__device__ __forceinline__ bool leaf_hit(const SceneDev &sc, const CullBox &cb, const ExactBox &eb, int ref, LocalCounters &lc,
										 int kind, const double *shape_data) {
	const unsigned info = (unsigned) (-1 - ref), first = info >> 4, cnt = (info & 15u) + 1u;
	unsigned unknown;
	if (leaf_prefilter(sc, cb, ref, &unknown, kind == SHAPE_KIND_BOX)) return true;  // (overflow path: non-box shapes go straight to the exact test)
	for (unsigned k = 0; k < cnt; ++k) {
		if (!(unknown >> k & 1u)) continue;
		lc.leaf++;
		if (shape_triangle_hit(kind, eb, shape_data, sc.tris + 9ull * (first + k))) return true;
	}
	return false;
}

/// Plain one-thread traversal (stage A's queue-overflow path): does the box touch any triangle of the scene?
__device__ __noinline__ bool box_hits_scene(const SceneDev &sc, const CullBox &cb, D3 c, Q4 q, const double *half,
											LocalCounters &lc, const BoxDev *shape = nullptr, const double *shape_data_base = nullptr) {
	if (!sc.n_nodes) return false;
	const int kind = shape ? shape->kind : (int) SHAPE_KIND_BOX;
	const double *shape_data = kind == SHAPE_KIND_CONVEX ? shape_data_base + shape->data_off : nullptr;
	const ExactBox eb = make_exact_box(c, q, half);
	int stack[STACK_DEPTH], cur = 0;
	LocalStack st{stack, 0};
	lc.boxes++;
	while (cur != REF_END) {
		if (cur >= 0) {
			lc.nodes++;
			cur = visit_record(sc, cb, cur, st);
		} else {
			if (leaf_hit(sc, cb, eb, cur, lc, kind, shape_data)) return true;
			cur = st.sp > 0 ? st.pop() : REF_END;
		}
	}
	return false;
}

// ---------------------------------------------------------------------------------------------------------------
// Sharded append lists.  A single append counter serialises in L2: ~300 k warp-level atomics on one address cost
// ~0.3 ms, 3x the HBM time of a 10 M-state batch.  Both work lists therefore keep SHARDS counters; shard s owns the
// 32-slot blocks s, s + SHARDS, s + 2 SHARDS, ... of the list, a producer warp always appends to the same shard, and
// consumers walk the flat slot range, skipping the holes (slots beyond their shard's count).  Blocks of 32 keep a
// warp's consecutive appends contiguous, so neighbouring consumer lanes still see neighbouring queries.
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned SHARDS = 64, SHARD_BLOCK = 32;
constexpr unsigned COUNTER_STRIDE = 32;  // one 128-byte line per counter: neighbouring counters would share an L2 atomic unit
constexpr size_t QUEUE_HEADER = SHARDS * COUNTER_STRIDE * sizeof(unsigned) + 256;  // queue counters + cursor: reset before every pass
constexpr size_t SCRATCH_HEADER = QUEUE_HEADER + 2 * SHARDS * COUNTER_STRIDE * sizeof(unsigned);  // + two survivor lists

__device__ __forceinline__ unsigned my_shard() { return (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (SHARDS - 1); }
__device__ __forceinline__ unsigned shard_slot(unsigned shard, unsigned local) {
	return ((local / SHARD_BLOCK) * SHARDS + shard) * SHARD_BLOCK + (local % SHARD_BLOCK);
}
/// Consumers copy the SHARDS counters of a finished list into shared memory once per block (block-wide call).
__device__ __forceinline__ void load_shard_counts(unsigned *s_counts, const unsigned *counts) {
	if (threadIdx.x < SHARDS) s_counts[threadIdx.x] = __ldcg(counts + threadIdx.x * COUNTER_STRIDE);
	__syncthreads();
}
/// Is flat slot `idx` filled?
__device__ __forceinline__ bool slot_filled(const unsigned *s_counts, unsigned idx) {
	const unsigned blk = idx / SHARD_BLOCK, shard = blk % SHARDS, local = (blk / SHARDS) * SHARD_BLOCK + idx % SHARD_BLOCK;
	return local < s_counts[shard];
}
/// Flat slot range that covers every filled slot (warp-uniform; all 32 lanes must call).
__device__ __forceinline__ unsigned slot_range(const unsigned *s_counts, unsigned cap) {
	const unsigned lane = threadIdx.x & 31;
	unsigned m = max(s_counts[lane], s_counts[lane + 32]);
	for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	const unsigned long long r = (unsigned long long) ((m + SHARD_BLOCK - 1) / SHARD_BLOCK) * SHARDS * SHARD_BLOCK;
	return (unsigned) (r < cap ? r : cap);
}

struct QueueDev {
	Item *items;
	unsigned *counts;  // SHARDS append counters (a shard may run past cap: the excess was traversed inline)
	unsigned *cursor;  // stage B's chunk dispenser
	unsigned cap;
};

/// Stage A tail: cull one box against the scene bounds and hand it to stage B (or traverse it here on overflow).
__device__ __forceinline__ void emit_box(const SceneDev &sc, const QueueDev &Q, const Results &res, D3 c, Q4 q,
										 const double *half, const BoxDev *spheres, unsigned owner, unsigned step, unsigned box, LocalCounters &lc,
										 const double *shape_data = nullptr) {
	const CullBox cb = make_cull_box(sc, c, q, half);
	if (!cull_overlap(cb, sc.root_c[0], sc.root_c[1], sc.root_c[2], sc.root_h[0], sc.root_h[1], sc.root_h[2])) return;
	if (sc.clear) {
		// clearance grid: the box is certainly free when each of the spheres that cover it is (independent one-byte loads)
		bool free_of_scene = true;
		if (spheres) {
			const float need = spheres->sphere_r + sc.margin + 1e-4f;
#pragma unroll
			for (int k = 0; k < BOX_SPHERES; ++k) {
				const float ox = spheres->sphere_c[k][0], oy = spheres->sphere_c[k][1], oz = spheres->sphere_c[k][2];
				free_of_scene &= scene_clearance(sc, cb.cx + cb.r[0] * ox + cb.r[1] * oy + cb.r[2] * oz, cb.cy + cb.r[3] * ox + cb.r[4] * oy + cb.r[5] * oz,
												 cb.cz + cb.r[6] * ox + cb.r[7] * oy + cb.r[8] * oz) > need;
			}
		} else {
			free_of_scene = scene_clearance(sc, cb.cx, cb.cy, cb.cz) > sqrtf(cb.hx * cb.hx + cb.hy * cb.hy + cb.hz * cb.hz) + 1e-4f;
		}
		if (free_of_scene) return;
	} else if (sc.wide) {
		// (without a clearance grid) One level further while the box is in registers anyway: the root record's (up to four) children travel with the
		// kernel parameters.  A box that misses all of them would cost stage B a queue item, a refill and a record fetch.
		bool any = false;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			any |= __float_as_int(sc.top[2 * k].w) != WIDE_EMPTY &&
				   cull_overlap(cb, sc.top[2 * k].x, sc.top[2 * k].y, sc.top[2 * k].z, sc.top[2 * k + 1].x, sc.top[2 * k + 1].y, sc.top[2 * k + 1].z);
		if (!any) return;
	}
	cg::coalesced_group g = cg::coalesced_threads();
	unsigned base = 0;
	const unsigned shard = my_shard();
	if (g.thread_rank() == 0) base = atomicAdd(Q.counts + shard * COUNTER_STRIDE, g.size());
	const unsigned pos = shard_slot(shard, g.shfl(base, 0) + g.thread_rank());
	if (pos < Q.cap) {
		double2 *dst = reinterpret_cast<double2 *>(Q.items + pos);
		__stcg(dst + 0, make_double2(c.x, c.y));
		__stcg(dst + 1, make_double2(c.z, q.x));
		__stcg(dst + 2, make_double2(q.y, q.z));
		__stcg(dst + 3, make_double2(q.w, __hiloint2double((int) (step | box << 28), (int) owner)));
	} else if (!already_decided(res, owner, step)) {
		if (box_hits_scene(sc, cb, c, q, half, lc, spheres, shape_data)) report_hit(res, owner, step);
	}
}

__device__ __forceinline__ bool finite7(const double *p) { return isfinite(p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6]); }

/// Whole-robot cull: does the sphere (base origin, rb.reach) miss the scene bounds — and, one level down, the boxes of
/// all of the root's children?  (A tree's bounding box is mostly air around the trunk; the children's boxes are not.)
__device__ __forceinline__ bool robot_out_of_reach(const SceneDev &sc, const RobotDev &rb, D3 t, bool one_level_down = true) {
	const double px = t.x - sc.origin[0], py = t.y - sc.origin[1], pz = t.z - sc.origin[2];
	double dx = fmax(fabs(px - (double) sc.root_c[0]) - (double) sc.root_h[0], 0.0);
	double dy = fmax(fabs(py - (double) sc.root_c[1]) - (double) sc.root_h[1], 0.0);
	double dz = fmax(fabs(pz - (double) sc.root_c[2]) - (double) sc.root_h[2], 0.0);
	const double r = rb.reach + (double) sc.margin;
	if (!(dx * dx + dy * dy + dz * dz <= r * r)) return true;  // NaN counts as out of reach
	const float fx = (float) px, fy = (float) py, fz = (float) pz, fr = (float) r + 1e-4f;  // fp32 is enough from here on: conservative radius
	if (sc.clear) return scene_clearance(sc, fx, fy, fz) > fr;  // the grid knows the distance to the triangles themselves
	if (!sc.wide || !one_level_down) return false;
	bool near = false;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (__float_as_int(sc.top[2 * k].w) == WIDE_EMPTY) continue;
		const float ex = fmaxf(fabsf(fx - sc.top[2 * k].x) - sc.top[2 * k + 1].x, 0.f);
		const float ey = fmaxf(fabsf(fy - sc.top[2 * k].y) - sc.top[2 * k + 1].y, 0.f);
		const float ez = fmaxf(fabsf(fz - sc.top[2 * k].z) - sc.top[2 * k + 1].z, 0.f);
		near |= ex * ex + ey * ey + ez * ez <= fr * fr;
	}
	return !near;
}

/// FK for the links that carry collision geometry + one emit per box (check_robot_collision,
/// collision_detection.cpp:57-80; check_link_collision :16-55).  Ops run in the reference's DFS pop order; for chain
/// robots (every procedural drone) the running pose stays in registers and the link table is never touched.
__device__ __forceinline__ void emit_link_boxes(const SceneDev &sc, const RobotDev &rb, const QueueDev &Q, const Results &res,
												const Pose &link, int lo, int hi, unsigned owner, unsigned step, LocalCounters &lc) {
	for (int bi = lo; bi < hi; ++bi) {
		const BoxDev &bx = rb.boxes[bi];
		const Pose w = bx.tf_identity ? link : then(link, D3{bx.t[0], bx.t[1], bx.t[2]}, Q4{bx.q[0], bx.q[1], bx.q[2], bx.q[3]}, false);
		emit_box(sc, Q, res, w.t, w.q, bx.half, &bx, owner, step, (unsigned) bi, lc, rb.shape_data);
	}
}
__device__ __forceinline__ void expand_state(const SceneDev &sc, const RobotDev &rb, const QueueDev &Q, const Results &res,
											 const Pose &base, const double *joints, unsigned owner, unsigned step,
											 LocalCounters &lc) {
	Pose link_tf[MGODPL_MAX_LINKS];
	link_tf[rb.root] = base;
	emit_link_boxes(sc, rb, Q, res, base, rb.root_box_begin, rb.root_box_end, owner, step, lc);
	Pose cur = base;
	for (int i = 0; i < rb.n_ops; ++i) {
		const FkOp &op = rb.ops[i];
		if (!(op.flags & FK_FOR_COLLISION)) continue;
		cur = fk_apply(op, (op.flags & FK_PARENT_IS_PREV) ? cur : link_tf[op.parent], joints);
		if (op.flags & FK_STORE) link_tf[op.child] = cur;
		emit_link_boxes(sc, rb, Q, res, cur, op.box_begin, op.box_end, owner, step, lc);
	}
}

__device__ __forceinline__ double u01_from_bits(unsigned long long x) { return (double) (x >> 11) * 0x1.0p-53; }
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/// Per-motion constants of the Eigen slerp rule (Quaternion.cpp:12-20), computed once in stage A1: with theta =
/// acos(|dot|) the weights sin((1-t) theta) / sin(theta) and sin(t theta) / sin(theta) become, by the angle-difference
/// identity, cos(t theta) - cot(theta) sin(t theta) and sin(t theta) / sin(theta): one sincos per step, no division.
/// theta < 0 flags the linear-weights case; the sign of inv_sin carries "dot < 0" (second weight negated).
struct __align__(16) SlerpConst {
	double theta, inv_sin, cot, pad;
};
__device__ __forceinline__ SlerpConst load_slerp(const SlerpConst *p) {
	const double2 lo = __ldg(reinterpret_cast<const double2 *>(p)), hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
	return SlerpConst{lo.x, lo.y, hi.x, 0.0};
}

}  // namespace mgodpl_b200
#include "expand.cuh"
namespace mgodpl_b200 {

// ---------------------------------------------------------------------------------------------------------------
// Stage A kernels
// ---------------------------------------------------------------------------------------------------------------
/// Survivor list between the broad cull (A1) and FK expansion (A2): 8-byte entries (owner, step).
struct SurvivorsDev {
	uint2 *entries;
	unsigned *counts;  // SHARDS append counters; slots at or beyond cap are expanded inline by the producer
	unsigned cap;
};

/// Overflow path of the survivor list (badly skewed batches only), kept out of line so that the streaming kernel
/// below stays small: expand one state on the spot.
__device__ __noinline__ void expand_state_now(const SceneDev &sc, const RobotDev &rb, const QueueDev &Q, const Results &res,
											  const double *row, unsigned owner) {
	LocalCounters lc;
	if (finite7(row)) expand_state(sc, rb, Q, res, Pose{D3{row[0], row[1], row[2]}, Q4{row[3], row[4], row[5], row[6]}}, row + 7, owner, 0u, lc);
}

/// A1 for explicit states: stream the state rows through shared memory with fully coalesced loads (a warp's 32 rows
/// are one contiguous span; warps run independently, so one warp's append round trip never stalls another's loads),
/// then keep the states whose reach sphere touches the scene bounds.  This
/// stage is the HBM-bound part of a state batch: 8 (7 + J_s) bytes per state, read exactly once.
__global__ void __launch_bounds__(BLOCK_A, 8)
k_cull_states(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ states,
			  size_t n, size_t stride, SurvivorsDev S, QueueDev Q, Results res) {
	extern __shared__ double s_rows[];  // BLOCK_A * stride: one 32-row span per warp, no block-wide barrier
	unsigned n_states = 0;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double *rows = s_rows + (size_t) warp * 32 * stride;
	const size_t n_spans = (n + 31) / 32, warps_total = (size_t) gridDim.x * (BLOCK_A / 32);
	const bool aligned16 = (reinterpret_cast<size_t>(states) & 15) == 0;  // a caller's slice may start on an odd row: 8-byte loads then
	for (size_t span = (size_t) blockIdx.x * (BLOCK_A / 32) + warp; span < n_spans; span += warps_total) {
		const size_t i0 = span * 32;
		const unsigned cnt = (unsigned) min((size_t) 32, n - i0);
		const double *src = states + i0 * stride;
		const unsigned total = cnt * (unsigned) stride;
		__syncwarp();
		if (cnt == 32 && aligned16) {  // 32 * stride doubles = a whole number of 16-byte pairs: 16-byte loads
			const double2 *src2 = reinterpret_cast<const double2 *>(src);
			double2 *rows2 = reinterpret_cast<double2 *>(rows);
			for (unsigned e = lane; e < total / 2; e += 32) rows2[e] = __ldcs(src2 + e);
		} else {
			for (unsigned e = lane; e < total; e += 32) rows[e] = __ldcs(src + e);
		}
		__syncwarp();
		if (lane >= cnt) continue;
		const double *r = rows + lane * stride;
		n_states++;
		if (robot_out_of_reach(sc, rb, D3{r[0], r[1], r[2]})) continue;
		cg::coalesced_group g = cg::coalesced_threads();
		unsigned base = 0;
		const unsigned shard = my_shard();
		if (g.thread_rank() == 0) base = atomicAdd(S.counts + shard * COUNTER_STRIDE, g.size());
		const unsigned pos = shard_slot(shard, g.shfl(base, 0) + g.thread_rank());
		const unsigned i = (unsigned) (i0 + lane);
		if (pos < S.cap) {
			S.entries[pos] = make_uint2(i, 0u);
		} else {
			expand_state_now(sc, rb, Q, res, r, i);  // list full (badly skewed batch)
		}
	}
	LocalCounters lc;
	lc.states = n_states;
	flush_counters(sc, lc);
}

// ---- the same stream through the bulk-copy engine (sm_90+ / sm_100a: cp.async.bulk + mbarrier) ---------------------------------
// One elected lane per warp asks the copy engine for a whole 32-row span (32 * stride * 8 bytes, one descriptor-less bulk
// copy global -> shared) and the warp keeps BULK_STAGES spans in flight; completion is signalled on an mbarrier by byte count,
// so no thread spends load instructions or registers on the stream.  MGODPL_CULL_BULK=1 selects it.
constexpr int BULK_STAGES = 3;
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
				 "r"(bytes), "r"(smem_addr(bar))
				 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
	unsigned done = 0;
	while (!done)
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(BLOCK_A, 4)
k_cull_states_bulk(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ states,
				   size_t n, size_t stride, SurvivorsDev S, QueueDev Q, Results res) {
	extern __shared__ __align__(128) unsigned char s_bulk[];
	constexpr int WARPS = BLOCK_A / 32;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned span_bytes = 32u * (unsigned) stride * 8u, span_pad = (span_bytes + 127u) & ~127u;
	unsigned long long *bars = reinterpret_cast<unsigned long long *>(s_bulk) + warp * BULK_STAGES;                 // WARPS x BULK_STAGES barriers
	unsigned char *bufs = s_bulk + ((WARPS * BULK_STAGES * 8 + 127) & ~127) + (size_t) warp * BULK_STAGES * span_pad;  // this warp's stages
	if (lane == 0)
		for (int st = 0; st < BULK_STAGES; ++st) mbar_init(bars + st, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	unsigned n_states = 0;
	const size_t full_spans = n / 32, warps_total = (size_t) gridDim.x * WARPS, first = (size_t) blockIdx.x * WARPS + warp;
	auto issue = [&](size_t j) {  // span number j of this warp -> stage j % BULK_STAGES (lane 0 only)
		const size_t span = first + j * warps_total;
		if (span >= full_spans) return;
		const int st = (int) (j % BULK_STAGES);
		mbar_expect_tx(bars + st, span_bytes);
		bulk_copy_g2s(bufs + (size_t) st * span_pad, states + span * 32 * stride, span_bytes, bars + st);
	};
	if (lane == 0)
		for (int j = 0; j < BULK_STAGES; ++j) issue((size_t) j);
	for (size_t j = 0; first + j * warps_total < full_spans; ++j) {
		const int st = (int) (j % BULK_STAGES);
		mbar_wait(bars + st, (unsigned) ((j / BULK_STAGES) & 1));
		const double *r = reinterpret_cast<const double *>(bufs + (size_t) st * span_pad) + lane * stride;
		const D3 t{r[0], r[1], r[2]};
		const size_t i0 = (first + j * warps_total) * 32;
		n_states++;
		const bool keep = !robot_out_of_reach(sc, rb, t);
		bool overflow = false;
		unsigned pos = 0;
		if (keep) {
			cg::coalesced_group g = cg::coalesced_threads();
			unsigned base = 0;
			const unsigned shard = my_shard();
			if (g.thread_rank() == 0) base = atomicAdd(S.counts + shard * COUNTER_STRIDE, g.size());
			pos = shard_slot(shard, g.shfl(base, 0) + g.thread_rank());
			if (pos < S.cap) S.entries[pos] = make_uint2((unsigned) (i0 + lane), 0u);
			else overflow = true;
		}
		if (overflow) expand_state_now(sc, rb, Q, res, r, (unsigned) (i0 + lane));  // list full (badly skewed batch); reads the row before it is recycled
		__syncwarp();  // every lane is done with this stage: hand it back to the copy engine
		if (lane == 0) issue(j + BULK_STAGES);
	}
	// the ragged last span (n % 32 rows) goes through ordinary loads, by one warp of the grid
	if (n % 32 && first == 0 && lane < n % 32) {
		const size_t i = full_spans * 32 + lane;
		const double *r = states + i * stride;
		n_states++;
		if (!robot_out_of_reach(sc, rb, D3{__ldg(r), __ldg(r + 1), __ldg(r + 2)})) {
			cg::coalesced_group g = cg::coalesced_threads();
			unsigned base = 0;
			const unsigned shard = my_shard();
			if (g.thread_rank() == 0) base = atomicAdd(S.counts + shard * COUNTER_STRIDE, g.size());
			const unsigned pos = shard_slot(shard, g.shfl(base, 0) + g.thread_rank());
			if (pos < S.cap) S.entries[pos] = make_uint2((unsigned) i, 0u);
			else expand_state_now(sc, rb, Q, res, r, (unsigned) i);
		}
	}
	LocalCounters lc;
	lc.states = n_states;
	flush_counters(sc, lc);
}

/// A2 for explicit states: FK + per-box cull + queue append, one thread per surviving state
/// (check_robot_collision, collision_detection.cpp:57-80).  CHAIN: the robot is a chain (every procedural drone), so no
/// link table and no joint array exist — the running pose and the joint value an op needs are all that is live.
template<bool CHAIN>
__global__ void __launch_bounds__(CHAIN ? BLOCK_C : BLOCK_A, CHAIN ? 5 : 2)
k_expand_states(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ states,
				size_t stride, SurvivorsDev S, unsigned pass_lo, unsigned pass_hi, QueueDev Q, Results res) {
	LocalCounters lc;
	__shared__ unsigned s_counts[SHARDS];
	load_shard_counts(s_counts, S.counts);
	const unsigned n = min(slot_range(s_counts, S.cap), pass_hi);
	DeferredBox deferred[MGODPL_MAX_BOXES];
	QueueSink sink{sc, Q, res, rb, lc, deferred, 0};
	for (unsigned k = pass_lo + blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		if (!slot_filled(s_counts, k)) continue;
		const unsigned i = __ldcg(&S.entries[k]).x;
		const double *r = states + (size_t) i * stride;
		double r7[7];
#pragma unroll
		for (int j = 0; j < 7; ++j) r7[j] = __ldg(r + j);
		if (!finite7(r7)) continue;
		expand_boxes<CHAIN, false>(sc, rb, sink, true, Pose{D3{r7[0], r7[1], r7[2]}, Q4{r7[3], r7[4], r7[5], r7[6]}}, JointsFromRow{r + 7}, i, 0u, 0, MGODPL_MAX_BOXES);
		sink.flush();
	}
	flush_counters(sc, lc);
}

/// Explicit world-frame boxes: check_link_collision (collision_detection.cpp:16-55), n x (tf7, size3).
__global__ void __launch_bounds__(BLOCK_A)
k_expand_boxes(const __grid_constant__ SceneDev sc, const double *__restrict__ boxes, size_t pass_lo, size_t n, QueueDev Q, Results res) {
	LocalCounters lc;
	for (size_t i = pass_lo + (size_t) blockIdx.x * BLOCK_A + threadIdx.x; i < n; i += (size_t) gridDim.x * BLOCK_A) {
		double v[10];
#pragma unroll
		for (int k = 0; k < 10; ++k) v[k] = __ldg(boxes + i * 10 + k);
		if (!finite7(v)) continue;
		const double half[3] = {v[7] * 0.5, v[8] * 0.5, v[9] * 0.5};
		emit_box(sc, Q, res, D3{v[0], v[1], v[2]}, Q4{v[3], v[4], v[5], v[6]}, half, nullptr, (unsigned) i, 0u, 0u, lc);
	}
	flush_counters(sc, lc);
}

/// Goal samples: genGoalStateUniform (goal_sampling.cpp:51-79) for `rows` targets x `per_row` samples, then the collision
/// filter.  Result bit index = row * (32 * words_per_row) + sample.
template<bool CHAIN>
__global__ void __launch_bounds__(CHAIN ? BLOCK_C : BLOCK_A, CHAIN ? 5 : 2)
k_expand_goals(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ targets,
			   size_t rows, size_t per_row, const double *__restrict__ draws,
			   unsigned long long seed, double distance_from_target, double *__restrict__ states_out, size_t pass_lo, size_t pass_hi,
			   QueueDev Q, Results res) {
	LocalCounters lc;
	DeferredBox deferred[MGODPL_MAX_BOXES];
	QueueSink sink{sc, Q, res, rb, lc, deferred, 0};
	const size_t total = min(rows * per_row, pass_hi);
	for (size_t i = pass_lo + (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
		Pose base{};
		JointsGoal jg{nullptr, rb.limits, seed, 0ull};
		unsigned owner = 0;
		if (!make_goal_sample(sc, rb, targets, per_row, draws, seed, distance_from_target, states_out, i, base, jg, owner, lc)) continue;
		expand_boxes<CHAIN, false>(sc, rb, sink, true, base, jg, owner, 0u, 0, MGODPL_MAX_BOXES);
		sink.flush();
	}
	flush_counters(sc, lc);
}

/// One interpolation step of one motion: interpolate(a, b, t) (RobotState.cpp:14-24, Transform.h:69-78, Eigen slerp
/// rule of Quaternion.cpp:12-20), exact reach cull, FK, box emit.  A/B point at the two endpoint rows.
__device__ __forceinline__ void expand_motion_step(const SceneDev &sc, const RobotDev &rb, const QueueDev &Q, const Results &res,
												   const double *__restrict__ A, const double *__restrict__ B, unsigned n_steps,
												   SlerpConst slerp_k, unsigned motion, unsigned step, LocalCounters &lc,
												   const unsigned *skip_if_first_hit_le = nullptr) {
	// the "an earlier step of this motion already collides" probe travels with the endpoint rows: one round trip, not two
	unsigned first_hit = NO_HIT;
	if (skip_if_first_hit_le) first_hit = __ldcg(skip_if_first_hit_le);
	double a[7], b[7];
#pragma unroll
	for (int k = 0; k < 7; ++k) a[k] = __ldg(A + k), b[k] = __ldg(B + k);
	if (first_hit <= step) return;
	lc.states++;
	// for n_steps == 0 the reference evaluates t = 0/0; here the start state is checked once.
	const double t = n_steps ? (double) step / (double) n_steps : 0.0;
	const D3 tr{(1.0 - t) * a[0] + t * b[0], (1.0 - t) * a[1] + t * b[1], (1.0 - t) * a[2] + t * b[2]};
	if (robot_out_of_reach(sc, rb, tr, false)) return;
	Pose base{tr, Q4{a[3], a[4], a[5], a[6]}};
	double jv[MGODPL_MAX_JOINT_VALUES];
	if (n_steps) {
		double w0, w1;
		if (slerp_k.theta < 0.0) {
			w0 = 1.0 - t, w1 = t;
		} else {
			double s1, c1;
			sincos(t * slerp_k.theta, &s1, &c1);
			w0 = c1 - slerp_k.cot * s1;
			w1 = s1 * fabs(slerp_k.inv_sin);
		}
		if (signbit(slerp_k.inv_sin)) w1 = -w1;
		base.q = Q4{w0 * a[3] + w1 * b[3], w0 * a[4] + w1 * b[4], w0 * a[5] + w1 * b[5], w0 * a[6] + w1 * b[6]};
		for (int k = 0; k < rb.n_vars; ++k) jv[k] = __ldg(A + 7 + k) * (1 - t) + __ldg(B + 7 + k) * t;
	} else {
		for (int k = 0; k < rb.n_vars; ++k) jv[k] = __ldg(A + 7 + k);
	}
	if (!isfinite(base.q.x + base.q.y + base.q.z + base.q.w)) return;
	expand_state(sc, rb, Q, res, base, jv, motion, step, lc);
}

/// Steps that found their survivor list full (badly skewed batches only): k_motion_ranges records them as ranges — the slots
/// a warp writes for one motion and list ascend, so what does not fit is a suffix — and k_expand_overflow_ranges, a kernel of its
/// own right behind it, expands them.  (The expansion used to be an out-of-line call inside k_motion_ranges; a call anywhere in
/// that kernel cost it 330 bytes of spills and 13 of its 45 us on every batch, used or not.)
struct OverflowRanges {
	uint4 *entries;   // (motion, first step, count, -): at most one per motion and list
	unsigned *count;
	unsigned cap;
};

/// Expand one interpolation step on the spot (out of line: rare paths only).
__device__ __noinline__ void expand_motion_step_now(const SceneDev &sc, const RobotDev &rb, const QueueDev &Q, const Results &res,
													const double *A, const double *B, unsigned n_steps, const SlerpConst *slerp_k, unsigned motion, unsigned step) {
	LocalCounters lc;
	expand_motion_step(sc, rb, Q, res, A, B, n_steps, *slerp_k, motion, step, lc);
	flush_counters_lane(sc, lc);
}

/// A1 for motions, one thread per motion: the step count of check_motion_collides (collision_detection.cpp:88-92)
/// and — instead of visiting every step — the contiguous range of steps whose base position can lie within the
/// robot's reach of the scene bounds (slab clip of the translation segment, widened by one step either side).
/// The surviving (motion, step) pairs go to the survivor list, written warp-cooperatively.
template<int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
k_motion_ranges(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ a,
				const double *__restrict__ b, size_t m, size_t stride, int js, double max_step,
				const uint32_t *__restrict__ n_steps_override, uint32_t *__restrict__ n_steps_out,
				uint32_t *__restrict__ uncertain_bits, SlerpConst *__restrict__ slerp_out, SurvivorsDev S, SurvivorsDev S_tail,
				OverflowRanges O, unsigned head_steps, int trim, Results res) {
	const unsigned lane = threadIdx.x & 31;
	LocalCounters lc;
	unsigned long long n_uncertain = 0, n_motions = 0;
	const size_t m_pad = (m + 31) & ~(size_t) 31;
	for (size_t i = (size_t) blockIdx.x * BLOCK + threadIdx.x; i < m_pad; i += (size_t) gridDim.x * BLOCK) {
		unsigned s_lo = 0, len = 0, n_steps = 0, head = head_steps;
		bool uncertain = false;
		if (i < m) {
			res.first_step[i] = NO_HIT;  // (this kernel initialises the per-motion result: no memset in front of it)
			const double *A = a + i * stride, *B = b + i * stride;
			double ar[7], br[7];
#pragma unroll
			for (int k = 0; k < 7; ++k) ar[k] = __ldg(A + k), br[k] = __ldg(B + k);
			// equal_weights_distance (distance.cpp:17-25, Quaternion.h:78-85) in the reference's operand order
			const double dx = ar[0] - br[0], dy = ar[1] - br[1], dz = ar[2] - br[2];
			double d = sqrt(dx * dx + dy * dy + dz * dz);
			const double c = ar[6] * br[6] + ar[3] * br[3] + ar[4] * br[4] + ar[5] * br[5];
			const double cc = fmin(fmax(c, -1.0), 1.0);
			d += acos(2.0 * cc * cc - 1.0);
			for (int k = 0; k < js; ++k) d += fabs(__ldg(A + 7 + k) - __ldg(B + 7 + k));
			const double trans = sqrt(dx * dx + dy * dy + dz * dz);
			const double x = d / max_step;
			double nd = ceil(x);
			uncertain = fabs(x - rint(x)) < 1e-9 * fmax(1.0, x);
			if (!(nd >= 0.0 && nd < 268435455.0)) nd = 0.0;  // non-finite / absurd distances: one check of the start state
			n_steps = (unsigned) nd;
			if (n_steps_override) {
				const unsigned o = __ldg(n_steps_override + i);
				if (o != NO_HIT) n_steps = o, uncertain = false;
			}
			if (n_steps_out) n_steps_out[i] = n_steps;
			{
				// Eigen slerp rule (Quaternion.cpp:12-20): dot over (x, y, z, w); linear weights when |dot| >= 1 - eps
				const double dq = ar[3] * br[3] + ar[4] * br[4] + ar[5] * br[5] + ar[6] * br[6];
				const double ad = fabs(dq);
				SlerpConst k{-1.0, 1.0, 0.0, 0.0};
				if (ad < 1.0 - 2.220446049250313e-16) {
					k.theta = acos(ad);
					const double st = sin(k.theta);
					k.inv_sin = 1.0 / st, k.cot = ad / st;
				}
				if (dq < 0.0) k.inv_sin = -k.inv_sin;
				slerp_out[i] = k;
			}
			n_uncertain += uncertain;
			n_motions++;
			// slab clip of P(t) = A + t (B - A), t in [0, 1], against the scene bounds grown by the reach
			const double grow = rb.reach + (double) sc.margin + 1e-6;
			double t0 = 0.0, t1 = 1.0;
			bool empty = false;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const double p = ar[k] - (sc.origin[k] + (double) sc.root_c[k]);
				const double e = (double) sc.root_h[k] + grow;
				const double dl = n_steps ? br[k] - ar[k] : 0.0;
				if (fabs(dl) < 1e-300) {
					empty |= !(fabs(p) <= e);
				} else {
					double ta = (-e - p) / dl, tb = (e - p) / dl;
					if (ta > tb) {
						const double tmp = ta;
						ta = tb, tb = tmp;
					}
					t0 = fmax(t0, ta), t1 = fmin(t1, tb);
				}
			}
			if (!empty && t0 <= t1) {
				const double lo = floor(t0 * (double) n_steps) - 1.0, hi = ceil(t1 * (double) n_steps) + 1.0;
				s_lo = (unsigned) fmax(lo, 0.0);
				unsigned s_hi = (unsigned) fmin(hi, (double) n_steps);
				// The slab range is the box-shaped superset of the rounded "within reach" region, which is convex: trim
				// both ends with the exact test stage A2 applies (same t, same lerp), so A2 mostly sees real survivors.
				// `gap(s)` > 0: at step s the whole robot clears the scene by at least that much (distance from the base origin
				// to the scene bounds, or to the triangles themselves where the clearance grid knows it, minus the reach).  The
				// base moves step_len per step and distances are 1-Lipschitz, so a gap of g rules out floor(g / step_len) steps at
				// once: a few look-ups walk each end of the range up to the first step that can touch anything.
				// (fp32 on origin-relative coordinates with a reciprocal instead of divisions: this runs per look-up on a few lanes
				// of the warp; 1e-3 m of extra reach and a 1e-3 relative cut of the stride swallow every rounding it introduces.)
				const float fa[3] = {(float) (ar[0] - sc.origin[0]), (float) (ar[1] - sc.origin[1]), (float) (ar[2] - sc.origin[2])};
				const float fd[3] = {(float) (br[0] - ar[0]), (float) (br[1] - ar[1]), (float) (br[2] - ar[2])};
				const float inv_n = n_steps ? 1.0f / (float) n_steps : 0.f;
				const float fr = (float) (rb.reach + (double) sc.margin) + 1e-3f;
				const float step_len = n_steps ? (float) trans * inv_n : 0.f;
				const float inv_step = step_len > 0.f ? 0.999f / step_len : 0.f;
				auto gap = [&](unsigned s) {
					const float t = (float) s * inv_n;
					const float px = fa[0] + t * fd[0], py = fa[1] + t * fd[1], pz = fa[2] + t * fd[2];
					const float ex = fmaxf(fabsf(px - sc.root_c[0]) - sc.root_h[0], 0.f), ey = fmaxf(fabsf(py - sc.root_c[1]) - sc.root_h[1], 0.f),
								ez = fmaxf(fabsf(pz - sc.root_c[2]) - sc.root_h[2], 0.f);
					float g = sqrtf(ex * ex + ey * ey + ez * ez) - fr;
					if (sc.clear && g <= 0.f) g = scene_clearance(sc, px, py, pz) - fr;
					return g;
				};
				auto steps_cleared = [&](float g, unsigned room) {  // NaN positions count as clear, one step at a time
					if (!(g > 0.f) || !(inv_step > 0.f)) return 1u;
					return (unsigned) fminf(fmaxf(floorf(g * inv_step), 1.f), (float) max(room, 1u));
				};
				// both ends walk inwards in the same loop: their look-ups are independent, so the two dependent chains of (cold)
				// grid loads overlap instead of running one after the other
				bool lo_open = true, hi_open = true;
				for (int guard = 0; guard < trim && (lo_open || hi_open) && s_lo < s_hi; ++guard) {
					const float g_lo = lo_open ? gap(s_lo) : 0.f, g_hi = hi_open ? gap(s_hi) : 0.f;
					if (lo_open) {
						if (g_lo <= 0.f) lo_open = false;
						else s_lo += steps_cleared(g_lo, s_hi - s_lo);
					}
					if (hi_open && s_hi > s_lo) {
						if (g_hi <= 0.f) hi_open = false;
						else s_hi -= steps_cleared(g_hi, s_hi - s_lo);
					}
				}
				len = (s_lo == s_hi && !(gap(s_lo) <= 0.f)) ? 0u : s_hi - s_lo + 1;
				// head_steps == 0: adaptive head — the steps it takes to cross the reach shell around the scene bounds plus
				// ~0.75 m of canopy, where a colliding motion typically finds its first hit
				if (!head_steps) head = (unsigned) fmin(fmax((rb.reach + 0.75) * (double) n_steps / fmax(trans, 1e-3 * (double) n_steps), 8.0), 256.0);
			}
		}
		{
			const unsigned ub = __ballot_sync(0xffffffffu, uncertain);  // one whole word per warp iteration: no memset, no atomics
			if (lane == 0 && uncertain_bits) uncertain_bits[i >> 5] = ub;
		}
		__syncwarp();  // slerp_out[] / first_step[] of this warp's motions are touched by other lanes on the overflow path
		// cooperative append.  One allocation per warp and list: a prefix sum over the lanes' range lengths and a single
		// atomicAdd of the total (a chain of per-motion atomics, each a dependent L2 round trip, used to be this kernel's
		// critical path).  The first `head` steps of a range go to the head list, the rest to the tail list: the head is
		// expanded and traversed first, and tail entries of motions that already collided by then are never expanded.
		const unsigned l_head = min(len, head), l_tail = len - l_head;
		unsigned incl_h = l_head, incl_t = l_tail;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned x = __shfl_up_sync(0xffffffffu, incl_h, o), y = __shfl_up_sync(0xffffffffu, incl_t, o);
			if ((int) lane >= o) incl_h += x, incl_t += y;
		}
		const unsigned tot_h = __shfl_sync(0xffffffffu, incl_h, 31), tot_t = __shfl_sync(0xffffffffu, incl_t, 31);
		if (tot_h + tot_t) {
			const unsigned shard = my_shard();
			unsigned base_h = 0, base_t = 0;
			if (lane == 0) {
				if (tot_h) base_h = atomicAdd(S.counts + shard * COUNTER_STRIDE, tot_h);
				if (tot_t) base_t = atomicAdd(S_tail.counts + shard * COUNTER_STRIDE, tot_t);
			}
			base_h = __shfl_sync(0xffffffffu, base_h, 0) + incl_h - l_head;
			base_t = __shfl_sync(0xffffffffu, base_t, 0) + incl_t - l_tail;
			unsigned todo = __ballot_sync(0xffffffffu, len > 0);
			while (todo) {  // for every lane with a non-empty range, the whole warp writes its (motion, step) pairs
				const int src = __ffs(todo) - 1;
				todo &= todo - 1;
				const unsigned lo = __shfl_sync(0xffffffffu, s_lo, src), lh = __shfl_sync(0xffffffffu, l_head, src);
				const unsigned mot = (unsigned) (i - lane + src);
				const unsigned cnts[2] = {lh, __shfl_sync(0xffffffffu, l_tail, src)};
				const unsigned bases[2] = {__shfl_sync(0xffffffffu, base_h, src), __shfl_sync(0xffffffffu, base_t, src)};
				for (int part = 0; part < 2; ++part) {
					const SurvivorsDev &L = part ? S_tail : S;
					const unsigned first = part ? lh : 0u;
					unsigned k_full = cnts[part];  // first entry of this lane's share that found the list full
					for (unsigned k = lane; k < cnts[part]; k += 32) {
						const unsigned pos = shard_slot(shard, bases[part] + k);
						if (pos < L.cap) L.entries[pos] = make_uint2(mot, lo + first + k);
						else k_full = min(k_full, k);
					}
					if (__any_sync(0xffffffffu, k_full < cnts[part])) {  // list full: hand the rest of the range to k_expand_overflow_ranges
						k_full = __reduce_min_sync(0xffffffffu, k_full);
						if (lane == 0) {
							const unsigned at = atomicAdd(O.count, 1u);
							if (at < O.cap) O.entries[at] = make_uint4(mot, lo + first + k_full, cnts[part] - k_full, 0u);
						}
					}
				}
			}
		}
	}
	flush_counters(sc, lc);
	if (sc.counters) {
		if (n_motions) atomicAdd(sc.counters + CNT_MOTIONS, n_motions);
		if (n_uncertain) atomicAdd(sc.counters + CNT_UNCERTAIN, n_uncertain);
	}
}

/// The ranges k_motion_ranges could not list (see OverflowRanges): one warp per range, expanded on the spot.  Launched behind
/// every k_motion_ranges; without overflow it reads one counter and exits.
__global__ void __launch_bounds__(128)
k_expand_overflow_ranges(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ a,
						 const double *__restrict__ b, size_t stride, const uint32_t *__restrict__ n_steps, const SlerpConst *__restrict__ slerp_k,
						 OverflowRanges O, QueueDev Q, Results res) {
	const unsigned n = min(__ldcg(O.count), O.cap);
	if (!n) return;
	const unsigned lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (unsigned t = warp; t < n; t += n_warps) {
		const uint4 e = __ldcg(O.entries + t);
		for (unsigned k = lane; k < e.z; k += 32)
			expand_motion_step_now(sc, rb, Q, res, a + (size_t) e.x * stride, b + (size_t) e.x * stride, __ldg(n_steps + e.x), slerp_k + e.x, e.x, e.y + k);
	}
}

/// A2 for motions: one thread per surviving (motion, step).
template<bool CHAIN>
__global__ void __launch_bounds__(CHAIN ? BLOCK_C : BLOCK_A, CHAIN ? 5 : 2)
k_expand_motion_steps(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ a,
					  const double *__restrict__ b, size_t stride, const uint32_t *__restrict__ n_steps,
					  const SlerpConst *__restrict__ slerp_k, SurvivorsDev S, unsigned pass_lo, unsigned pass_hi, int skip_decided, QueueDev Q,
					  Results res) {
	LocalCounters lc;
	__shared__ unsigned s_counts[SHARDS];
	load_shard_counts(s_counts, S.counts);
	const unsigned n = min(slot_range(s_counts, S.cap), pass_hi);
	DeferredBox deferred[MGODPL_MAX_BOXES];
	QueueSink sink{sc, Q, res, rb, lc, deferred, 0};
	for (unsigned k = pass_lo + blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		if (!slot_filled(s_counts, k)) continue;
		const uint2 e = __ldcg(&S.entries[k]);
		Pose base{};
		JointsLerp jl{nullptr, nullptr, 0.0, false};
		if (!make_motion_step(sc, rb, a + (size_t) e.x * stride, b + (size_t) e.x * stride, __ldg(n_steps + e.x), slerp_k + e.x, e.y,
							  skip_decided ? res.first_step + e.x : nullptr, base, jl, lc))
			continue;
		expand_boxes<CHAIN, false>(sc, rb, sink, true, base, jl, e.x, e.y, 0, MGODPL_MAX_BOXES);
		sink.flush();
	}
	flush_counters(sc, lc);
}

// ---------------------------------------------------------------------------------------------------------------
// Stage B: drain the queue.  Persistent warps, one box per lane, "while-while" with lane refill:
//   refill   lanes whose box is finished take the next queued boxes (warp-local chunks of the queue, so neighbouring
//            lanes hold neighbouring interpolation steps / samples and walk the tree together);
//   descend  every lane walks the BVH (one 64-byte record = both children per visit, small per-lane stack) until it
//            reaches an overlapping leaf (it parks there) or runs out of nodes (it goes idle); the warp leaves this
//            loop once a refill or a leaf batch is worthwhile;
//   leaves   lanes parked on a leaf run the exact fp64 box/triangle test together, in batches of >= leaf_min lanes
//            (the test is ~400 instructions; running it for 3 lanes at a time cost as much as all descents).
// Without refill a warp's time is its slowest lane (a few boxes deep in the canopy among many that die at the top
// of the tree): measured 3.6 active threads per instruction; this loop keeps most lanes busy.
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned QUEUE_CHUNK = 64;    // items a warp claims per cursor atomic
struct TraverseTuning {
	int refill_min;   // idle lanes that trigger a refill
	int leaf_min;     // parked lanes (on an overlapping leaf) that trigger a batch of exact tests
	int no_prefilter; // MGODPL_NO_PREFILTER=1: every leaf triangle goes to the fp64 test (tests compare both modes)
};

template<int MIN_BLOCKS, bool WIDE>
__global__ void __launch_bounds__(BLOCK_B, MIN_BLOCKS)
k_traverse(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ explicit_boxes,
		   QueueDev Q, Results res, TraverseTuning tune) {
	const int REFILL_MIN = tune.refill_min, LEAF_MIN = tune.leaf_min;
	constexpr unsigned FULL = 0xffffffffu;
	const unsigned lane = threadIdx.x & 31;
	__shared__ unsigned s_counts[SHARDS];
	load_shard_counts(s_counts, Q.counts);
	const unsigned n_items = sc.n_nodes ? slot_range(s_counts, Q.cap) : 0u;  // flat slot range, holes included
	// few items per warp: claim one refill at a time so that the queue spreads over all the warps (a short queue is a latency problem)
	const unsigned claim = n_items > gridDim.x * (BLOCK_B / 32) * QUEUE_CHUNK * 4u ? QUEUE_CHUNK : 32u;
	LocalCounters lc;
	unsigned chunk_lo = 0, chunk_hi = 0;  // warp-uniform: unclaimed part of this warp's chunk
	bool exhausted = false;               // warp-uniform: the queue has nothing left to claim
	// lane state: cur >= 0 walking (record index) | cur < 0 parked on a leaf reference | REF_END idle
	CullBox cb{};
	unsigned owner = 0, step_box = 0, item = 0;
	int cur = REF_END;
	__shared__ int s_stack[STACK_SHARED + 1][BLOCK_B];
	int stack_overflow[(WIDE ? WIDE_STACK_DEPTH : STACK_DEPTH) - STACK_SHARED];
	LaneStack st{&s_stack[0][threadIdx.x], stack_overflow, BLOCK_B, 0};
	unsigned box_nodes = 0, max_box_nodes = 0;  // longest single descent (counters only)

	for (;;) {
		const bool idle = cur == REF_END;
		const unsigned idle_mask = __ballot_sync(FULL, idle);
		if (idle_mask == FULL && exhausted) break;
		// ---- refill ---------------------------------------------------------------------------------------------
		if (!exhausted && (__popc(idle_mask) >= REFILL_MIN)) {
			if (chunk_lo == chunk_hi) {
				unsigned lo = 0;
				if (lane == 0) lo = atomicAdd(Q.cursor, claim);
				lo = __shfl_sync(FULL, lo, 0);
				chunk_lo = min(lo, n_items), chunk_hi = min(lo + claim, n_items);
				exhausted = chunk_lo >= n_items;
			}
			const unsigned avail = chunk_hi - chunk_lo, rank = __popc(idle_mask & ((1u << lane) - 1u));
			if (idle && rank < avail && slot_filled(s_counts, chunk_lo + rank)) {
				item = chunk_lo + rank;
				const double2 *src = reinterpret_cast<const double2 *>(Q.items + item);
				const double2 v0 = __ldcg(src), v1 = __ldcg(src + 1), v2 = __ldcg(src + 2), v3 = __ldcg(src + 3);
				owner = (unsigned) __double2loint(v3.y), step_box = (unsigned) __double2hiint(v3.y);
				// Motions: skip boxes of steps that can no longer be the first colliding one.  (For state bitmasks the
				// same check only runs before a leaf test: it would cost a dependent L2 round trip per refill here.)
				// The probe is issued first and consumed last, so the cull box is built while it is in flight; the lane
				// is idle, so overwriting `cb` for a box that turns out to be skipped costs nothing.
				unsigned first_hit = NO_HIT;
				if (res.mode != RESULT_BITS) first_hit = __ldcg(res.first_step + owner);
				double half_d[3];
				if (explicit_boxes) {
					const double *bx = explicit_boxes + (size_t) owner * 10;
					half_d[0] = __ldg(bx + 7) * 0.5, half_d[1] = __ldg(bx + 8) * 0.5, half_d[2] = __ldg(bx + 9) * 0.5;
				} else {
					const double *hh = rb.boxes[step_box >> 28].half;
					half_d[0] = hh[0], half_d[1] = hh[1], half_d[2] = hh[2];
				}
				cb = make_cull_box(sc, D3{v0.x, v0.y, v1.x}, Q4{v1.y, v2.x, v2.y, v3.x}, half_d);
				if (res.mode == RESULT_BITS || !(first_hit <= (step_box & 0x0FFFFFFFu))) {
					cur = 0, st.sp = 0;
					lc.boxes++;
					max_box_nodes = max(max_box_nodes, box_nodes), box_nodes = 0;
				}
			}
			chunk_lo += min(avail, (unsigned) __popc(idle_mask));
			continue;  // re-evaluate who is idle (a refill may have come up short at a chunk boundary)
		}
		// ---- descend: until another action becomes worthwhile (enough idle lanes to refill, enough parked lanes for a
		//      leaf batch) or nobody is left walking ------------------------------------------------------------------
		unsigned go = __ballot_sync(FULL, cur >= 0), parked = __ballot_sync(FULL, cur < 0 && cur != REF_END);
		while (go) {
			if (cur >= 0) {
				lc.nodes++, box_nodes++;
				cur = WIDE ? visit_wide(sc, cb, cur, st) : visit_record(sc, cb, cur, st);
			}
			go = __ballot_sync(FULL, cur >= 0);
			parked = __ballot_sync(FULL, cur < 0 && cur != REF_END);
			if (__popc(parked) >= LEAF_MIN) break;
			if (!exhausted && 32 - __popc(go) - __popc(parked) >= REFILL_MIN) break;
		}
		// ---- leaves: parked lanes run the exact test together, once there are enough of them or nothing else can move -----
		if (__popc(parked) >= LEAF_MIN || (!go && parked)) {
			if (cur < 0 && cur != REF_END) {
				const unsigned step = step_box & 0x0FFFFFFFu;
				if (already_decided(res, owner, step)) {
					cur = REF_END;
				} else {
					unsigned unknown = 0xffffu;
					const bool is_box = explicit_boxes || rb.all_boxes || rb.boxes[step_box >> 28].kind == SHAPE_KIND_BOX;
					bool hit = false;
					if (!tune.no_prefilter) hit = is_box ? leaf_prefilter(sc, cb, cur, &unknown) : leaf_prefilter_shape(sc, cb, cur, &unknown, &rb.boxes[step_box >> 28], rb.shape_data);
					if (!hit && unknown) {
						double half_buf[3];
						const double *half;
						const BoxDev *shape = nullptr;
						if (explicit_boxes) {
							const double *bx = explicit_boxes + (size_t) owner * 10;
							half_buf[0] = __ldg(bx + 7) * 0.5, half_buf[1] = __ldg(bx + 8) * 0.5, half_buf[2] = __ldg(bx + 9) * 0.5;
							half = half_buf;
						} else {
							shape = &rb.boxes[step_box >> 28];
							half = shape->half;
						}
						unsigned n_exact = 0;
						hit = exact_leaf_test(sc, reinterpret_cast<const char *>(Q.items + item), half, cur, unknown, &n_exact, shape, rb.shape_data);
						lc.leaf += n_exact;
					}
					if (hit) {
						report_hit(res, owner, step);
						cur = REF_END;
					} else {
						cur = st.sp > 0 ? st.pop() : REF_END;
					}
				}
			}
		}
	}
	flush_counters(sc, lc);
	if (sc.counters) atomicMax(sc.counters + CNT_MAX_BOX_NODES, (unsigned long long) max(max_box_nodes, box_nodes));
}

}  // namespace mgodpl_b200
#include "pipeline.cuh"
namespace mgodpl_b200 {

// ---------------------------------------------------------------------------------------------------------------
// Stage C (motions): first colliding step -> packed hit bit and toi = step / n_steps (collision_detection.cpp:95-101).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_finish_motions(const unsigned *__restrict__ first_step, const uint32_t *__restrict__ n_steps, size_t m,
				 uint32_t *__restrict__ hit_bits, double *__restrict__ toi) {
	const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
	const unsigned fs = i < m ? first_step[i] : NO_HIT;
	const bool hit = fs != NO_HIT;
	const unsigned word = __ballot_sync(0xffffffffu, hit);
	if (i < m) {
		if ((threadIdx.x & 31) == 0) hit_bits[i >> 5] = word;
		if (toi) {
			const unsigned n = n_steps[i];
			toi[i] = hit ? (n ? (double) fs / (double) n : 0.0) : __longlong_as_double(0x7ff8000000000000ll);
		}
	}
}

/// Per-row population count of a packed hit matrix (collision counts per goal target).
__global__ void k_count_rows(const uint32_t *__restrict__ bits, size_t rows, size_t words_per_row, uint32_t *__restrict__ counts) {
	const size_t row = blockIdx.x;
	if (row >= rows) return;
	unsigned s = 0;
	for (size_t w = threadIdx.x; w < words_per_row; w += blockDim.x) s += __popc(bits[row * words_per_row + w]);
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	__shared__ unsigned acc;
	if (threadIdx.x == 0) acc = 0;
	__syncthreads();
	if ((threadIdx.x & 31) == 0 && s) atomicAdd(&acc, s);
	__syncthreads();
	if (threadIdx.x == 0) counts[row] = acc;
}

/// Batched forwardKinematics (RobotModel.cpp:17-90): all link transforms of every state, n x n_links x 7 doubles.
__global__ void __launch_bounds__(256)
k_forward_kinematics(const __grid_constant__ RobotDev rb, const double *__restrict__ states, size_t n, size_t stride,
					 double *__restrict__ out) {
	const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const double *r = states + i * stride;
	double jv[MGODPL_MAX_JOINT_VALUES];
	for (int k = 0; k < rb.n_vars; ++k) jv[k] = r[7 + k];
	Pose tf[MGODPL_MAX_LINKS];
	const Pose base{{r[0], r[1], r[2]}, {r[3], r[4], r[5], r[6]}};
	// links the DFS never reaches keep the root transform (RobotModel.cpp:23-24)
	for (int l = 0; l < rb.n_links; ++l) tf[l] = base;
	for (int k = 0; k < rb.n_ops; ++k) tf[rb.ops[k].child] = fk_apply(rb.ops[k], tf[rb.ops[k].parent], jv);
	for (int l = 0; l < rb.n_links; ++l) {
		double *o = out + (i * rb.n_links + l) * 7;
		o[0] = tf[l].t.x, o[1] = tf[l].t.y, o[2] = tf[l].t.z;
		o[3] = tf[l].q.x, o[4] = tf[l].q.y, o[5] = tf[l].q.z, o[6] = tf[l].q.w;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------------------
static int sm_count() {  // of the current device (a process may drive several)
	int dev = 0, n = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
	return n > 0 ? n : 148;
}

/// Environment knobs, read once (thread-safe: function-local static of a struct initialised by one expression).
struct Knobs {
	int overlap, pipeline, pipe_refill_min, pipe_leaf_min, pipe_blocks, pipe_claim, cull_bulk, ranges_variant, no_prefilter, traverse_blocks, refill_min, leaf_min, trim;
	unsigned head_steps;
	size_t queue_max_items, survivor_max_entries;
	static int geti(const char *name, int dflt) {
		const char *e = getenv(name);
		return e ? atoi(e) : dflt;
	}
	static size_t getz(const char *name, size_t dflt) {
		const char *e = getenv(name);
		return e && atoll(e) > 0 ? (size_t) atoll(e) : dflt;
	}
	Knobs()
		: overlap(geti("MGODPL_OVERLAP", 1)), pipeline(geti("MGODPL_PIPELINE", 0)), pipe_refill_min(geti("MGODPL_PIPE_REFILL_MIN", 8)), pipe_leaf_min(geti("MGODPL_PIPE_LEAF_MIN", 32)), pipe_blocks(geti("MGODPL_PIPE_BLOCKS", 4)), pipe_claim(geti("MGODPL_PIPE_CLAIM", 0)), cull_bulk(geti("MGODPL_CULL_BULK", 0)), ranges_variant(geti("MGODPL_RANGES_VARIANT", 1)),
		  no_prefilter(geti("MGODPL_NO_PREFILTER", 0)), traverse_blocks(geti("MGODPL_TRAVERSE_BLOCKS", 4)), refill_min(geti("MGODPL_REFILL_MIN", 16)),
		  leaf_min(geti("MGODPL_LEAF_MIN", 8)), trim(geti("MGODPL_TRIM", 16)),
		  head_steps(geti("MGODPL_HEAD_STEPS", 0) > 0 ? (unsigned) geti("MGODPL_HEAD_STEPS", 0) : 0u),
		  queue_max_items(getz("MGODPL_QUEUE_MAX_ITEMS", QUEUE_MAX_ITEMS)), survivor_max_entries(getz("MGODPL_SURVIVOR_MAX_ENTRIES", SURVIVOR_MAX_ENTRIES)) {}
};
static const Knobs &knobs() {
	static const Knobs k;
	return k;
}
/// The experimental fused kernels (pipeline.cuh) only know box links: robots with capsules / convex links take the default path.
static int pipeline_mode(const RobotDev &rb) { return rb.all_boxes ? knobs().pipeline : 0; }
/// Grids are a multiple of the SM count (148 on B200) times the resident blocks per SM, or smaller for small inputs.
static unsigned grid_for(size_t work_items, int block, int blocks_per_sm) {
	size_t blocks = (work_items + block - 1) / block, cap = (size_t) sm_count() * blocks_per_sm;
	return (unsigned) (blocks < cap ? (blocks ? blocks : 1) : cap);
}

/// Survivor slots for n explicit states: the sharded list has holes, so leave room for imbalance between shards.
size_t state_survivor_slots(size_t n) { return n + n / 4 + SHARDS * SHARD_BLOCK; }

/// Capacity limits of the two work lists.  MGODPL_QUEUE_MAX_ITEMS / MGODPL_SURVIVOR_MAX_ENTRIES shrink them (tests use
/// this to drive every query through the inline overflow paths).
static size_t queue_max_items() { return knobs().queue_max_items; }
static size_t survivor_max_entries() { return knobs().survivor_max_entries; }

bool motion_batch_is_one_pass(size_t m, int n_boxes) {
	return knobs().pipeline == 0 && (m * 32 + 4096) * (size_t) std::max(n_boxes, 1) <= queue_max_items() && m * 64 <= survivor_max_entries();
}

/// Room for the tail round's own queue (8 boxes per motion; launch_check_motions carves it off the end of its scratch).
size_t motion_tail_queue_bytes(size_t m, int n_boxes) { return QUEUE_HEADER + (8 * m + 4096) * (size_t) std::max(n_boxes, 1) * sizeof(Item); }

/// Room for one overflow-range record per motion and list (launch_check_motions carves it off the end of its scratch).
size_t motion_overflow_bytes(size_t m) { return ((2 * m + 64) * sizeof(uint4) + 255) & ~(size_t) 255; }

size_t query_scratch_bytes(size_t worst_case_boxes, size_t survivor_entries) {
	if (knobs().pipeline == 1) worst_case_boxes = 0;  // the fused kernel keeps its boxes in shared memory: no queue in HBM
	size_t cap = worst_case_boxes < queue_max_items() ? worst_case_boxes : queue_max_items();
	size_t scap = survivor_entries < survivor_max_entries() ? survivor_entries : survivor_max_entries();
	return SCRATCH_HEADER + ((scap * sizeof(uint2) + 255) & ~(size_t) 255) + (cap ? cap : 1) * sizeof(Item);
}

/// Scratch layout: [header: queue shard counters | survivor shard counters | queue cursor][survivor entries][queue items].
static void carve(void *scratch, size_t scratch_bytes, size_t survivor_entries, SurvivorsDev *S, QueueDev *Q,
				  SurvivorsDev *S_tail = nullptr, size_t head_entries = 0) {
	char *p = reinterpret_cast<char *>(scratch);
	unsigned *hdr = reinterpret_cast<unsigned *>(p);
	Q->counts = hdr;
	Q->cursor = hdr + SHARDS * COUNTER_STRIDE;
	size_t scap = survivor_entries < survivor_max_entries() ? survivor_entries : survivor_max_entries();
	size_t sbytes = (scap * sizeof(uint2) + 255) & ~(size_t) 255;
	if (S) {
		S->counts = hdr + QUEUE_HEADER / sizeof(unsigned);
		S->entries = reinterpret_cast<uint2 *>(p + SCRATCH_HEADER);
		S->cap = (unsigned) scap;
	}
	if (S && S_tail) {  // split the entry region: [head | tail]
		const size_t grain = (size_t) SHARDS * SHARD_BLOCK;
		size_t hcap = head_entries < scap / 2 ? head_entries : scap / 2;
		hcap = hcap >= grain ? hcap / grain * grain : std::min(grain, scap);  // (down: a queue sized for the list's worst case then takes it in one pass)
		S->cap = (unsigned) hcap;
		S_tail->counts = S->counts + SHARDS * COUNTER_STRIDE;
		S_tail->entries = S->entries + hcap;
		S_tail->cap = (unsigned) (scap - hcap);
	}
	Q->items = reinterpret_cast<Item *>(p + SCRATCH_HEADER + sbytes);
	size_t cap = scratch_bytes > SCRATCH_HEADER + sbytes ? (scratch_bytes - SCRATCH_HEADER - sbytes) / sizeof(Item) : 0;
	Q->cap = (unsigned) (cap < queue_max_items() ? cap : queue_max_items());
}
static QueueDev carve_queue(void *scratch, size_t scratch_bytes) {
	QueueDev Q;
	carve(scratch, scratch_bytes, 0, nullptr, &Q);
	return Q;
}

#define LQ_CHECK(x)                       \
	do {                                  \
		cudaError_t e_ = (x);             \
		if (e_ != cudaSuccess) return e_; \
	} while (0)

/// The fused expansion + traversal kernel (pipeline.cuh) over `bound` candidates at most.
template<int KIND, bool CHAIN, bool WIDE>
static cudaError_t launch_pipeline_inst(const SceneDev &sc, const RobotDev &rb, const SourceArgs &src, unsigned *cursor, const Results &res,
										size_t bound, cudaStream_t stream) {
	constexpr size_t smem = (BLOCK_P / 32) * sizeof(WarpMem);
	auto kernel = k_pipeline<KIND, CHAIN, WIDE>;
	LQ_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));  // per device, idempotent
	PipeTuning tune{knobs().pipe_refill_min, (unsigned) std::max(1, std::min(knobs().pipe_leaf_min, 32)), knobs().no_prefilter, 32u * CLAIM_ROUNDS};
	const size_t resident = (size_t) sm_count() * (size_t) std::max(1, std::min(knobs().pipe_blocks, 4));
	// small batches: one claim of 32 per warp so that the work spreads over the SMs; large ones: fewer cursor atomics
	if (bound < resident * (BLOCK_P / 32) * tune.claim) tune.claim = 32u;
	if (knobs().pipe_claim > 0) tune.claim = 32u * (unsigned) knobs().pipe_claim;
	size_t blocks = (bound + (BLOCK_P / 32) * tune.claim - 1) / ((BLOCK_P / 32) * tune.claim);
	blocks = std::max<size_t>(1, std::min(blocks, resident));
	kernel<<<(unsigned) blocks, BLOCK_P, smem, stream>>>(sc, rb, src, cursor, res, tune);
	return cudaGetLastError();
}
template<int KIND>
static cudaError_t launch_pipeline(const SceneDev &sc, const RobotDev &rb, const SourceArgs &src, unsigned *cursor, const Results &res, size_t bound,
								   cudaStream_t stream) {
	if (!sc.n_nodes && KIND != SRC_GOALS) return cudaSuccess;  // an empty scene collides with nothing (goal sampling still writes its states)
	if (bound >= 0xFFF00000ull) return cudaErrorInvalidValue;   // the candidate cursor is 32 bits wide
	constexpr bool one_box = KIND == SRC_BOXES || KIND == SRC_QUEUE;  // no FK in the kernel: a single instantiation
	const bool wide = sc.wide != nullptr, chain = one_box || rb.chain;
	if (wide) return chain ? launch_pipeline_inst<KIND, true, true>(sc, rb, src, cursor, res, bound, stream)
						   : launch_pipeline_inst<KIND, one_box, true>(sc, rb, src, cursor, res, bound, stream);
	return chain ? launch_pipeline_inst<KIND, true, false>(sc, rb, src, cursor, res, bound, stream)
				 : launch_pipeline_inst<KIND, one_box, false>(sc, rb, src, cursor, res, bound, stream);
}

static cudaError_t run_traverse(const SceneDev &sc, const RobotDev &rb, const double *explicit_boxes, const QueueDev &Q,
								const Results &res, size_t worst_case_items, cudaStream_t stream) {
	size_t bound = worst_case_items < Q.cap ? worst_case_items : Q.cap;
	if (pipeline_mode(rb) == 2) {  // the pipeline kernel fed from the queue: warp-wide leaf batches and shared-memory stacks, expansion stays a kernel of its own
		SourceArgs src{};
		src.queue = Q, src.boxes = explicit_boxes;
		return launch_pipeline<SRC_QUEUE>(sc, rb, src, Q.cursor, res, bound, stream);
	}
	const int min_blocks = knobs().traverse_blocks;
	const TraverseTuning tune{knobs().refill_min, knobs().leaf_min, knobs().no_prefilter};
	const bool wide = sc.wide != nullptr;
#define MGODPL_LAUNCH_TRAVERSE(MB)                                                                                                      \
	if (wide) k_traverse<MB, true><<<grid_for(bound, BLOCK_B, MB), BLOCK_B, 0, stream>>>(sc, rb, explicit_boxes, Q, res, tune);       \
	else k_traverse<MB, false><<<grid_for(bound, BLOCK_B, MB), BLOCK_B, 0, stream>>>(sc, rb, explicit_boxes, Q, res, tune)
	switch (min_blocks) {
		case 8: MGODPL_LAUNCH_TRAVERSE(8); break;
		case 3: MGODPL_LAUNCH_TRAVERSE(3); break;
		case 2: MGODPL_LAUNCH_TRAVERSE(2); break;
		case 6: MGODPL_LAUNCH_TRAVERSE(6); break;
		case 5: MGODPL_LAUNCH_TRAVERSE(5); break;
		default: MGODPL_LAUNCH_TRAVERSE(4); break;
	}
#undef MGODPL_LAUNCH_TRAVERSE
	return cudaGetLastError();
}

/// Stage A2 + B run in passes over the candidates so that a pass can never produce more boxes than the queue holds
/// (pass size = queue capacity / boxes per state).  A batch whose worst case fits runs one pass with no host
/// involvement; otherwise the survivor count is read back once and only the passes that have work are launched.  (Before this, a 1 M-motion orchard batch pushed 3/4 of
/// its boxes through stage A's one-thread-per-box overflow path: 14 ms instead of ~2.)
/// How far the survivor list actually got (flat slot range).  Only called when a batch cannot be guaranteed to fit one
/// pass: a small synchronous read-back of the shard counters is cheaper than launching passes over an empty tail.
static cudaError_t survivor_slot_ranges(const SurvivorsDev *lists, int n_lists, cudaStream_t stream, size_t *ranges) {
	// the lists' counter blocks are adjacent in the scratch header: one copy, one synchronisation
	static thread_local unsigned h_counts[2 * SHARDS * COUNTER_STRIDE];
	cudaError_t e = cudaMemcpyAsync(h_counts, lists[0].counts, (size_t) n_lists * SHARDS * COUNTER_STRIDE * sizeof(unsigned),
									cudaMemcpyDeviceToHost, stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
	if (e != cudaSuccess) return e;
	for (int l = 0; l < n_lists; ++l) {
		unsigned m = 0;
		for (unsigned s = 0; s < SHARDS; ++s) m = std::max(m, h_counts[(l * SHARDS + s) * COUNTER_STRIDE]);
		size_t r = (size_t) ((m + SHARD_BLOCK - 1) / SHARD_BLOCK) * SHARDS * SHARD_BLOCK;
		ranges[l] = r < lists[l].cap ? r : lists[l].cap;
	}
	return cudaSuccess;
}

static size_t pass_size(const QueueDev &Q, int boxes_per_state) {
	size_t r = (size_t) Q.cap / (size_t) (boxes_per_state > 0 ? boxes_per_state : 1);
	const size_t grain = (size_t) SHARDS * SHARD_BLOCK;
	r = r / grain * grain;
	return r ? r : grain;
}

cudaError_t launch_check_states(const SceneDev &sc, const RobotDev &rb, const double *d_states, size_t n, size_t stride,
								uint32_t *d_hit_bits, void *scratch, size_t scratch_bytes, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	SurvivorsDev S;
	QueueDev Q;
	carve(scratch, scratch_bytes, state_survivor_slots(n), &S, &Q);
	Results res{RESULT_BITS, d_hit_bits, nullptr};
	LQ_CHECK(cudaMemsetAsync(Q.counts, 0, SCRATCH_HEADER, stream));
	LQ_CHECK(cudaMemsetAsync(d_hit_bits, 0, ((n + 31) / 32) * 4, stream));
	if (pipeline_mode(rb) == 1) Q.cap = 0;  // the fused kernel keeps its boxes on chip; only stage A1's overflow path would queue, and traverses inline instead
	if (knobs().cull_bulk && (reinterpret_cast<size_t>(d_states) & 15) == 0 && (32 * stride * 8) % 16 == 0) {
		const size_t span_pad = (32 * stride * 8 + 127) & ~(size_t) 127;
		const size_t smem = (((BLOCK_A / 32) * BULK_STAGES * 8 + 127) & ~(size_t) 127) + (BLOCK_A / 32) * BULK_STAGES * span_pad;
		LQ_CHECK(cudaFuncSetAttribute(k_cull_states_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		const unsigned per_sm = (unsigned) std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem));
		k_cull_states_bulk<<<grid_for(n, BLOCK_A, (int) per_sm), BLOCK_A, smem, stream>>>(sc, rb, d_states, n, stride, S, Q, res);
	} else {
		k_cull_states<<<grid_for(n, BLOCK_A, 8), BLOCK_A, BLOCK_A * stride * sizeof(double), stream>>>(sc, rb, d_states, n, stride, S, Q, res);
	}
	LQ_CHECK(cudaGetLastError());
	if (pipeline_mode(rb) == 1) {
		SourceArgs src{};
		src.list[0] = S, src.n_lists = 1, src.states = d_states, src.stride = stride;
		return launch_pipeline<SRC_STATES>(sc, rb, src, Q.cursor, res, std::min<size_t>(S.cap, n + SHARDS * SHARD_BLOCK) * (size_t) ((rb.n_boxes + SPL - 1) / SPL > 0 ? (rb.n_boxes + SPL - 1) / SPL : 1), stream);
	}
	// The list's slot range has holes (sharded appends) but at most n filled slots: a queue that holds n * n_boxes items takes the
	// whole list in one pass, with no read-back (the call stays asynchronous and can be captured into a CUDA graph).
	const bool one_pass = (size_t) Q.cap >= n * (size_t) std::max(rb.n_boxes, 1);
	const size_t R = one_pass ? (size_t) S.cap : pass_size(Q, rb.n_boxes);
	size_t end = S.cap;
	if (end > R) LQ_CHECK(survivor_slot_ranges(&S, 1, stream, &end));
	for (size_t lo = 0; lo < end || lo == 0; lo += R) {
		const size_t hi = lo + R < end ? lo + R : end;
		if (lo) LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
		if (rb.chain) k_expand_states<true><<<grid_for(hi - lo, BLOCK_C, 5), BLOCK_C, 0, stream>>>(sc, rb, d_states, stride, S, (unsigned) lo, (unsigned) hi, Q, res);
		else k_expand_states<false><<<grid_for(hi - lo, BLOCK_A, 2), BLOCK_A, 0, stream>>>(sc, rb, d_states, stride, S, (unsigned) lo, (unsigned) hi, Q, res);
		LQ_CHECK(cudaGetLastError());
		LQ_CHECK(run_traverse(sc, rb, nullptr, Q, res, (hi - lo) * (size_t) rb.n_boxes, stream));
	}
	return cudaSuccess;
}

cudaError_t launch_check_boxes(const SceneDev &sc, const double *d_boxes, size_t n, uint32_t *d_hit_bits, void *scratch,
							   size_t scratch_bytes, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	QueueDev Q = carve_queue(scratch, scratch_bytes);
	Results res{RESULT_BITS, d_hit_bits, nullptr};
	RobotDev none{};
	none.all_boxes = 1;
	LQ_CHECK(cudaMemsetAsync(d_hit_bits, 0, ((n + 31) / 32) * 4, stream));
	if (knobs().pipeline == 1) {
		LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
		SourceArgs src{};
		src.boxes = d_boxes, src.total = n;
		return launch_pipeline<SRC_BOXES>(sc, none, src, Q.cursor, res, n, stream);
	}
	const size_t R = pass_size(Q, 1);
	for (size_t lo = 0; lo < n; lo += R) {
		const size_t hi = lo + R < n ? lo + R : n;
		LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
		k_expand_boxes<<<grid_for(hi - lo, BLOCK_A, 8), BLOCK_A, 0, stream>>>(sc, d_boxes, lo, hi, Q, res);
		LQ_CHECK(cudaGetLastError());
		LQ_CHECK(run_traverse(sc, none, d_boxes, Q, res, hi - lo, stream));
	}
	return cudaSuccess;
}

cudaError_t launch_check_motions(const SceneDev &sc, const RobotDev &rb, const double *d_a, const double *d_b, size_t m,
								 size_t stride, int js, double max_step, const uint32_t *d_override, uint32_t *d_hit_bits,
								 double *d_toi, uint32_t *d_n_steps, uint32_t *d_uncertain, unsigned *d_first_step, void *d_slerp,
								 void *scratch, size_t scratch_bytes, size_t survivor_entries, cudaStream_t stream, const MotionOverlap *overlap) {
	if (!m) return cudaSuccess;
	SurvivorsDev S, S_tail;
	QueueDev Q;
	// (optional) a second, smaller queue for the tail round at the very end of the scratch: with it the tail expansion does not have
	// to wait for the head traversal to hand the queue back
	QueueDev Q2{};
	if (overlap && knobs().overlap) {
		const size_t tail_bytes = motion_tail_queue_bytes(m, rb.n_boxes);
		if (scratch_bytes > tail_bytes + SCRATCH_HEADER) {
			scratch_bytes -= tail_bytes;
			char *p = reinterpret_cast<char *>(scratch) + scratch_bytes;
			Q2.counts = reinterpret_cast<unsigned *>(p), Q2.cursor = Q2.counts + SHARDS * COUNTER_STRIDE;
			Q2.items = reinterpret_cast<Item *>(p + QUEUE_HEADER);
			Q2.cap = (unsigned) std::min<size_t>((tail_bytes - QUEUE_HEADER) / sizeof(Item), queue_max_items());
		}
	}
	const unsigned head_arg = knobs().head_steps;  // 0: adaptive per motion
	const int trim = knobs().trim;
	// the overflow-range records sit behind the queue: [header][survivor lists][queue items][overflow ranges]
	const size_t over_bytes = motion_overflow_bytes(m);
	if (scratch_bytes < over_bytes + SCRATCH_HEADER) return cudaErrorInvalidValue;
	scratch_bytes -= over_bytes;
	carve(scratch, scratch_bytes, survivor_entries, &S, &Q, &S_tail, m * (size_t) (head_arg ? head_arg : 48u));
	const OverflowRanges O{reinterpret_cast<uint4 *>(reinterpret_cast<char *>(scratch) + scratch_bytes), Q.cursor + 16, (unsigned) (2 * m + 64)};
	Results res{RESULT_FIRST_STEP, nullptr, d_first_step};
	// Overlapped rounds (see below) pay off when the launches are being recorded into a CUDA graph, where a cross-stream
	// dependency is an edge; between plain launches each one costs several microseconds and the overlap loses (0.291 -> 0.313 ms
	// per 100 k motions, against 0.277 -> 0.258 ms replayed).  MGODPL_OVERLAP=2 forces it on, 0 off.
	const size_t R_all = pass_size(Q, rb.n_boxes);
	cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
	if (Q2.items) LQ_CHECK(cudaStreamIsCapturing(stream, &capturing));
	const bool overlapped = Q2.items && pipeline_mode(rb) == 0 && S.cap <= R_all && (capturing == cudaStreamCaptureStatusActive || knobs().overlap == 2);
	LQ_CHECK(cudaMemsetAsync(Q.counts, 0, SCRATCH_HEADER, stream));
	if (pipeline_mode(rb) == 1) Q.cap = 0;
	if (knobs().ranges_variant == 1)  // one wave for 100 k motions: 6 blocks x 4 warps per SM at <= 85 registers
		k_motion_ranges<128, 6><<<grid_for(m, 128, 6), 128, 0, stream>>>(sc, rb, d_a, d_b, m, stride, js, max_step, d_override, d_n_steps, d_uncertain,
																		  (SlerpConst *) d_slerp, S, S_tail, O, head_arg, trim, res);
	else
		k_motion_ranges<BLOCK_A, 2><<<grid_for(m, BLOCK_A, 2), BLOCK_A, 0, stream>>>(sc, rb, d_a, d_b, m, stride, js, max_step, d_override, d_n_steps, d_uncertain,
																					  (SlerpConst *) d_slerp, S, S_tail, O, head_arg, trim, res);
	LQ_CHECK(cudaGetLastError());
	k_expand_overflow_ranges<<<(unsigned) sm_count(), 128, 0, stream>>>(sc, rb, d_a, d_b, stride, d_n_steps, (const SlerpConst *) d_slerp, O, Q, res);
	LQ_CHECK(cudaGetLastError());
	if (pipeline_mode(rb) == 1) {
		// one launch walks the head list, then the tail list: by the time a warp reaches a motion's tail steps, the head steps of
		// (almost) every motion have been decided, and tail steps behind a hit are dropped when they are claimed
		SourceArgs src{};
		src.list[0] = S, src.list[1] = S_tail, src.n_lists = 2, src.a = d_a, src.b = d_b, src.stride = stride;
		src.n_steps = d_n_steps, src.slerp = (const SlerpConst *) d_slerp;
		const size_t groups = (size_t) std::max(1, (rb.n_boxes + SPL - 1) / SPL);
		LQ_CHECK(launch_pipeline<SRC_MOTION_STEPS>(sc, rb, src, Q.cursor, res, ((size_t) S.cap + S_tail.cap) * groups, stream));
		k_finish_motions<<<(unsigned) ((m + 255) / 256), 256, 0, stream>>>(d_first_step, d_n_steps, m, d_hit_bits, d_toi);
		return cudaGetLastError();
	}
	// Head list first (the steps right after a motion comes within reach of the scene), then the tail with the
	// "already collided" skip: most of a colliding motion's in-reach steps lie after its first hit.
	const size_t R = pass_size(Q, rb.n_boxes);
	bool first_pass = true;
	const SurvivorsDev lists[2] = {S, S_tail};
	size_t ends[2] = {S.cap, S_tail.cap};
	auto expand = [&](const SurvivorsDev &L, size_t lo, size_t hi, int part, const QueueDev &q, cudaStream_t st) {
		if (rb.chain)
			k_expand_motion_steps<true><<<grid_for(hi - lo, BLOCK_C, 5), BLOCK_C, 0, st>>>(sc, rb, d_a, d_b, stride, d_n_steps, (const SlerpConst *) d_slerp, L,
																						   (unsigned) lo, (unsigned) hi, part, q, res);
		else
			k_expand_motion_steps<false><<<grid_for(hi - lo, BLOCK_A, 2), BLOCK_A, 0, st>>>(sc, rb, d_a, d_b, stride, d_n_steps, (const SlerpConst *) d_slerp, L,
																							(unsigned) lo, (unsigned) hi, part, q, res);
		return cudaGetLastError();
	};
	if (overlapped) {
		// Overlapped rounds (batches whose head list fits one pass).  The tail expansion runs on a side stream that waits for the
		// head expansion only.  The head traversal follows the head expansion on the caller's stream, so it is runnable first and
		// its persistent blocks fill every SM; the tail expansion's blocks — runnable a cross-stream hop later — get on as those
		// retire, i.e. during the traversal's drawn-out end, when most first hits are known and the "already collided" skip works
		// almost as well as after the kernel boundary.  A tail step expanded too early costs its FK, never a wrong answer: the
		// traversal re-checks the first hit before it walks a box.  The tail queue is small (8 boxes per motion); what does not
		// fit is traversed by its producer, as everywhere.  (Measured alternatives: the traversal on a high-priority side stream
		// instead — the tail expansion then starts FIRST and expands every tail step blind: the orchard shard went 0.77 -> 1.14 ms;
		// the whole pipeline on a high-priority stream — right order, but four cross-stream hops cost plain launches 8 %.)
		LQ_CHECK(cudaMemsetAsync(Q2.counts, 0, QUEUE_HEADER, stream));
		LQ_CHECK(expand(S, 0, ends[0], 0, Q, stream));
		LQ_CHECK(cudaEventRecord(overlap->ev_head, stream));
		LQ_CHECK(run_traverse(sc, rb, nullptr, Q, res, ends[0] * (size_t) rb.n_boxes, stream));
		LQ_CHECK(cudaStreamWaitEvent(overlap->side, overlap->ev_head, 0));
		LQ_CHECK(expand(S_tail, 0, ends[1], 1, Q2, overlap->side));
		LQ_CHECK(cudaEventRecord(overlap->ev_tail, overlap->side));
		LQ_CHECK(cudaStreamWaitEvent(stream, overlap->ev_tail, 0));
		LQ_CHECK(run_traverse(sc, rb, nullptr, Q2, res, ends[1] * (size_t) rb.n_boxes, stream));
	} else {
		if (ends[0] > R || ends[1] > R) LQ_CHECK(survivor_slot_ranges(lists, 2, stream, ends));
		for (int part = 0; part < 2; ++part) {
			const SurvivorsDev &L = lists[part];
			const size_t end = ends[part];
			for (size_t lo = 0; lo < end; lo += R) {
				const size_t hi = lo + R < end ? lo + R : end;
				// the first pass shares the queue with stage A1's own overflow path (a survivor list that ran full expands inline)
				if (!first_pass) LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
				first_pass = false;
				LQ_CHECK(expand(L, lo, hi, part, Q, stream));
				LQ_CHECK(run_traverse(sc, rb, nullptr, Q, res, (hi - lo) * (size_t) rb.n_boxes, stream));
			}
		}
	}
	k_finish_motions<<<(unsigned) ((m + 255) / 256), 256, 0, stream>>>(d_first_step, d_n_steps, m, d_hit_bits, d_toi);
	return cudaGetLastError();
}

cudaError_t launch_sample_goals(const SceneDev &sc, const RobotDev &rb, const double *d_targets, size_t f,
								const double *d_draws, size_t s, unsigned long long seed,
								double distance_from_target, uint32_t *d_hit_bits, uint32_t *d_collisions,
								double *d_states_out, void *scratch, size_t scratch_bytes, cudaStream_t stream) {
	if (!f || !s) return cudaSuccess;
	QueueDev Q = carve_queue(scratch, scratch_bytes);
	Results res{RESULT_BITS, d_hit_bits, nullptr};
	const size_t wpr = (s + 31) / 32, total = f * s;
	LQ_CHECK(cudaMemsetAsync(d_hit_bits, 0, f * wpr * 4, stream));
	if (pipeline_mode(rb) == 1) {
		LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
		SourceArgs src{};
		src.targets = d_targets, src.draws = d_draws, src.rows = f, src.per_row = s, src.seed = seed;
		src.distance_from_target = distance_from_target, src.states_out = d_states_out, src.total = total;
		LQ_CHECK(launch_pipeline<SRC_GOALS>(sc, rb, src, Q.cursor, res, total * (size_t) std::max(1, (rb.n_boxes + SPL - 1) / SPL), stream));
		if (d_collisions) {
			k_count_rows<<<(unsigned) f, 128, 0, stream>>>(d_hit_bits, f, wpr, d_collisions);
			LQ_CHECK(cudaGetLastError());
		}
		return cudaSuccess;
	}
	const size_t R = pass_size(Q, rb.n_boxes);  // the callers size the queue with 1/8 head-room for shard imbalance
	for (size_t lo = 0; lo < total; lo += R) {
		const size_t hi = lo + R < total ? lo + R : total;
		LQ_CHECK(cudaMemsetAsync(Q.counts, 0, QUEUE_HEADER, stream));
		if (rb.chain)
			k_expand_goals<true><<<grid_for(hi - lo, BLOCK_C, 5), BLOCK_C, 0, stream>>>(sc, rb, d_targets, f, s, d_draws, seed, distance_from_target, d_states_out, lo, hi, Q, res);
		else
			k_expand_goals<false><<<grid_for(hi - lo, BLOCK_A, 2), BLOCK_A, 0, stream>>>(sc, rb, d_targets, f, s, d_draws, seed, distance_from_target, d_states_out, lo, hi, Q, res);
		LQ_CHECK(cudaGetLastError());
		LQ_CHECK(run_traverse(sc, rb, nullptr, Q, res, (hi - lo) * (size_t) rb.n_boxes, stream));
	}
	if (d_collisions) {
		k_count_rows<<<(unsigned) f, 128, 0, stream>>>(d_hit_bits, f, wpr, d_collisions);
		LQ_CHECK(cudaGetLastError());
	}
	return cudaSuccess;
}

cudaError_t launch_forward_kinematics(const RobotDev &rb, const double *d_states, size_t n, size_t stride, double *d_out,
									  cudaStream_t stream) {
	if (!n) return cudaSuccess;
	k_forward_kinematics<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(rb, d_states, n, stride, d_out);
	return cudaGetLastError();
}

}  // namespace mgodpl_b200
