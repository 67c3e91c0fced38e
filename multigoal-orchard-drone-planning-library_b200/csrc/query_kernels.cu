// The collision-query kernels: batched forward kinematics, box-vs-BVH traversal, motion validation, goal sampling.
//
// Shape of every query kernel (DESIGN.md "Kernels"):
//   phase A  one thread per candidate robot state: load/generate the state, cull the whole robot against the scene
//            bounds with its reach sphere, run FK in fp64, cull each link box against the scene bounds, and push the
//            survivors (fp64 box pose + owner) into a block-wide queue in shared memory;
//   phase B  the block's threads drain the queue, one box per thread: stackless skip-link traversal of the flattened
//            BVH with fp32 conservative node tests, exact fp64 box/triangle tests at the leaves, first hit wins;
//   phase C  hits are OR-ed into a per-block bitmask and written out as packed words.
// Far-away states (the majority in PRM-style sampling) never leave phase A; the boxes that do need a tree descent are
// compacted so that phase B runs with full warps instead of a few live lanes per warp.
#include "geometry.cuh"
#include "kernels.h"

namespace mgodpl_b200 {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr int ROUND_BOXES = 2;            // link boxes pushed per state per round
constexpr int QCAP = BLOCK * ROUND_BOXES;

struct Queue {
	double c[3][QCAP];
	double q[4][QCAP];
	uint16_t owner[QCAP];
	uint8_t box[QCAP];
	unsigned count;
	unsigned hitmask[WARPS];
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

__device__ __forceinline__ bool finite7(const double *p) {
	double s = p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6];
	return isfinite(s);
}

/// Whole-robot cull: does the sphere (base origin, rb.reach) miss the scene bounds?
__device__ __forceinline__ bool robot_out_of_reach(const SceneDev &sc, const RobotDev &rb, D3 t) {
	double dx = fmax(fabs(t.x - (sc.origin[0] + (double) sc.root_c[0])) - (double) sc.root_h[0], 0.0);
	double dy = fmax(fabs(t.y - (sc.origin[1] + (double) sc.root_c[1])) - (double) sc.root_h[1], 0.0);
	double dz = fmax(fabs(t.z - (sc.origin[2] + (double) sc.root_c[2])) - (double) sc.root_h[2], 0.0);
	double r = rb.reach + (double) sc.margin;
	return dx * dx + dy * dy + dz * dz > r * r;
}

__device__ __forceinline__ void push_item(Queue &Q, const SceneDev &sc, D3 c, Q4 q, const double *half, unsigned owner,
										  unsigned box) {
	CullBox cb = make_cull_box(sc, c, q, half);
	if (!cull_overlap(cb, sc.root_c[0], sc.root_c[1], sc.root_c[2], sc.root_h[0], sc.root_h[1], sc.root_h[2])) return;
	unsigned pos = atomicAdd(&Q.count, 1u);
	Q.c[0][pos] = c.x, Q.c[1][pos] = c.y, Q.c[2][pos] = c.z;
	Q.q[0][pos] = q.x, Q.q[1][pos] = q.y, Q.q[2][pos] = q.z, Q.q[3][pos] = q.w;
	Q.owner[pos] = (uint16_t) owner;
	Q.box[pos] = (uint8_t) box;
}

/// FK for the links that carry collision geometry (check_robot_collision, collision_detection.cpp:57-80).
__device__ __forceinline__ void fk_collision_links(const RobotDev &rb, const Pose &base, const double *joints, Pose *link_tf) {
	link_tf[rb.root] = base;
	for (int i = 0; i < rb.n_ops; ++i) {
		const FkOp &op = rb.ops[i];
		if (op.flags & FK_FOR_COLLISION) link_tf[op.child] = fk_apply(op, link_tf[op.parent], joints);
	}
}

/// Phase B: drain the block queue.  `half_of(item)` yields the fp64 half extents of a queued box.
template<class HalfFn>
__device__ __forceinline__ void traverse_queue(Queue &Q, const SceneDev &sc, LocalCounters &lc, HalfFn half_of) {
	const unsigned n = Q.count;
	for (unsigned i = threadIdx.x; i < n; i += BLOCK) {
		const unsigned owner = Q.owner[i];
		if ((((volatile unsigned *) Q.hitmask)[owner >> 5] >> (owner & 31)) & 1u) continue;  // another box already hit
		D3 c{Q.c[0][i], Q.c[1][i], Q.c[2][i]};
		Q4 q{Q.q[0][i], Q.q[1][i], Q.q[2][i], Q.q[3][i]};
		const double *half = half_of(i, owner);
		const CullBox cb = make_cull_box(sc, c, q, half);
		ExactBox eb;
		eb.cx = c.x, eb.cy = c.y, eb.cz = c.z;
		qmat(q, eb.r);
		eb.hx = half[0], eb.hy = half[1], eb.hz = half[2];
		bool hit = false;
		unsigned node = 0;
		const unsigned n_nodes = sc.n_nodes;
		lc.boxes++;
		while (node < n_nodes) {
			const float4 a = __ldg(sc.nodes + 2 * node);
			const float4 b = __ldg(sc.nodes + 2 * node + 1);
			const bool ov = cull_overlap(cb, a.x, a.y, a.z, b.x, b.y, b.z);
			const int leaf = __float_as_int(b.w);
			lc.nodes++;
			if (ov && leaf >= 0) {
				const unsigned first = (unsigned) leaf >> 4, cnt = ((unsigned) leaf & 15u) + 1u;
				for (unsigned k = 0; k < cnt && !hit; ++k) {
					lc.leaf++;
					hit = box_triangle_hit(eb, sc.tris + 9ull * (first + k));
				}
				if (hit) break;
			}
			node = (ov && leaf < 0) ? node + 1 : __float_as_uint(a.w);
		}
		if (hit) atomicOr(&Q.hitmask[owner >> 5], 1u << (owner & 31));
	}
}

// ---------------------------------------------------------------------------------------------------------------
// State sources for the tiled state kernel.  A tile is BLOCK consecutive samples of one row; tile -> (row, s0).
// ---------------------------------------------------------------------------------------------------------------
/// Explicit states (check_robot_collision over N states): rows = 1.
struct ExplicitStates {
	const double *states;
	size_t stride;
	__device__ __forceinline__ void stage(double *smem, size_t row, size_t s0, unsigned cnt) const {
		const double *src = states + s0 * stride;
		for (unsigned i = threadIdx.x; i < cnt * stride; i += BLOCK) smem[i] = __ldg(src + i);
	}
	__device__ __forceinline__ bool get(const double *smem, const RobotDev &, size_t, size_t, unsigned tid, Pose &base,
										const double *&joints, double *) const {
		const double *r = smem + tid * stride;
		base.t = {r[0], r[1], r[2]};
		base.q = {r[3], r[4], r[5], r[6]};
		joints = r + 7;
		return finite7(r);
	}
};

__device__ __forceinline__ double u01_from_bits(unsigned long long x) { return (double) (x >> 11) * 0x1.0p-53; }
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/// genGoalStateUniform (goal_sampling.cpp:51-79): upright state from draws (yaw, joint values), FK to the end
/// effector, base shifted so the end effector lands on the row's target.  rows = targets, per_row = samples.
struct GoalStates {
	const double *targets;   // rows x 3
	const double *draws;     // per_row x (1 + n_vars), or nullptr -> counter-based generator
	const double *limits;    // n_vars x (min, max), device copy, used only by the generator
	unsigned long long seed;
	double distance_from_target;
	double *states_out;      // optional rows x per_row x (7 + n_vars)
	size_t per_row;
	__device__ __forceinline__ void stage(double *, size_t, size_t, unsigned) const {}
	__device__ __forceinline__ bool get(const double *, const RobotDev &rb, size_t row, size_t s, unsigned, Pose &base,
										const double *&joints, double *jv) const {
		const int nv = rb.n_vars;
		double yaw;
		if (draws) {
			const double *d = draws + s * (size_t) (1 + nv);
			yaw = __ldg(d);
			for (int k = 0; k < nv; ++k) jv[k] = __ldg(d + 1 + k);
		} else {
			unsigned long long ctr = (row * per_row + s) * (unsigned long long) (1 + nv);
			yaw = -M_PI + 2.0 * M_PI * u01_from_bits(mix64(seed ^ mix64(ctr)));
			for (int k = 0; k < nv; ++k) {
				double lo = limits[2 * k], hi = limits[2 * k + 1];
				jv[k] = lo + (hi - lo) * u01_from_bits(mix64(seed ^ mix64(ctr + 1 + k)));
			}
		}
		// genUprightState (goal_sampling.cpp:13-49): base at the origin, rotation about z.
		double sy, cy;
		sincos(yaw / 2.0, &sy, &cy);
		Pose tf[MGODPL_MAX_LINKS];
		tf[rb.root] = Pose{{0, 0, 0}, {0 * sy, 0 * sy, 1 * sy, cy}};
		for (int i = 0; i < rb.n_ops; ++i) tf[rb.ops[i].child] = fk_apply(rb.ops[i], tf[rb.ops[i].parent], jv);
		const Pose ee = tf[rb.ee];
		D3 tgt{__ldg(targets + 3 * row), __ldg(targets + 3 * row + 1), __ldg(targets + 3 * row + 2)};
		D3 delta = tgt - ee.t;
		if (distance_from_target > 0.0) delta = delta + qrot(ee.q, D3{0.0, -distance_from_target, 0.0});
		base.q = tf[rb.root].q;
		base.t = D3{0, 0, 0} + delta;
		joints = jv;
		if (states_out) {
			double *o = states_out + (row * per_row + s) * (size_t) (7 + nv);
			o[0] = base.t.x, o[1] = base.t.y, o[2] = base.t.z;
			o[3] = base.q.x, o[4] = base.q.y, o[5] = base.q.z, o[6] = base.q.w;
			for (int k = 0; k < nv; ++k) o[7 + k] = jv[k];
		}
		return isfinite(base.t.x + base.t.y + base.t.z);
	}
};

template<class Source>
__global__ void __launch_bounds__(BLOCK)
k_check_states(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const Source src, size_t rows,
			   size_t per_row, uint32_t *__restrict__ hit_bits) {
	__shared__ Queue Q;
	extern __shared__ double s_rows[];
	const unsigned tid = threadIdx.x;
	const size_t tiles_per_row = (per_row + BLOCK - 1) / BLOCK;
	const size_t words_per_row = (per_row + 31) / 32;
	const size_t n_tiles = rows * tiles_per_row;
	LocalCounters lc;
	for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const size_t row = tile / tiles_per_row, s0 = (tile - row * tiles_per_row) * BLOCK;
		const unsigned cnt = (unsigned) min((size_t) BLOCK, per_row - s0);
		if (tid < WARPS) Q.hitmask[tid] = 0;
		src.stage(s_rows, row, s0, cnt);
		__syncthreads();
		Pose link_tf[MGODPL_MAX_LINKS];
		double jv[MGODPL_MAX_JOINT_VALUES];
		const double *joints = nullptr;
		bool active = false;
		if (tid < cnt) {
			Pose base;
			active = src.get(s_rows, rb, row, s0 + tid, tid, base, joints, jv);
			lc.states++;
			active = active && !robot_out_of_reach(sc, rb, base.t);
			if (active) fk_collision_links(rb, base, joints, link_tf);
		}
		for (int g = 0; g < rb.n_boxes; g += ROUND_BOXES) {
			if (tid == 0) Q.count = 0;
			__syncthreads();
			if (active) {
				const int e = min(g + ROUND_BOXES, rb.n_boxes);
				for (int bi = g; bi < e; ++bi) {
					const BoxDev &bx = rb.boxes[bi];
					const Pose &l = link_tf[bx.link];
					Pose w = bx.tf_identity ? l : then(l, D3{bx.t[0], bx.t[1], bx.t[2]}, Q4{bx.q[0], bx.q[1], bx.q[2], bx.q[3]}, false);
					push_item(Q, sc, w.t, w.q, bx.half, tid, bi);
				}
			}
			__syncthreads();
			traverse_queue(Q, sc, lc, [&](unsigned i, unsigned) { return rb.boxes[Q.box[i]].half; });
			__syncthreads();
		}
		if (tid < WARPS && s0 + 32u * tid < per_row) hit_bits[row * words_per_row + s0 / 32 + tid] = Q.hitmask[tid];
		__syncthreads();
	}
	flush_counters(sc, lc);
}

/// check_link_collision for explicit world-frame boxes (collision_detection.cpp:16-55): n x (tf7, size3).
__global__ void __launch_bounds__(BLOCK)
k_check_boxes(const __grid_constant__ SceneDev sc, const double *__restrict__ boxes, size_t n, uint32_t *__restrict__ hit_bits) {
	__shared__ Queue Q;
	__shared__ double s_half[BLOCK][3];
	const unsigned tid = threadIdx.x;
	const size_t n_tiles = (n + BLOCK - 1) / BLOCK;
	LocalCounters lc;
	for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		if (tid < WARPS) Q.hitmask[tid] = 0;
		if (tid == 0) Q.count = 0;
		__syncthreads();
		const size_t idx = tile * BLOCK + tid;
		if (idx < n) {
			const double *r = boxes + idx * 10;
			double v[10];
#pragma unroll
			for (int k = 0; k < 10; ++k) v[k] = __ldg(r + k);
			s_half[tid][0] = v[7] * 0.5, s_half[tid][1] = v[8] * 0.5, s_half[tid][2] = v[9] * 0.5;
			if (finite7(v)) push_item(Q, sc, D3{v[0], v[1], v[2]}, Q4{v[3], v[4], v[5], v[6]}, s_half[tid], tid, 0);
		}
		__syncthreads();
		traverse_queue(Q, sc, lc, [&](unsigned, unsigned owner) { return (const double *) s_half[owner]; });
		__syncthreads();
		if (tid < WARPS && tile * BLOCK + 32u * tid < n) hit_bits[tile * WARPS + tid] = Q.hitmask[tid];
		__syncthreads();
	}
	flush_counters(sc, lc);
}

// ---------------------------------------------------------------------------------------------------------------
// Motion validation: check_motion_collides (collision_detection.cpp:82-105).  One warp per motion, lanes = 32
// consecutive interpolation steps; the block's 8 warps share the traversal queue; a motion stops at the first chunk
// that contains a hit and reports the lowest colliding step as toi.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ROW_MAX = 7 + MGODPL_MAX_JOINT_VALUES;

struct MotionSetup {
	unsigned n_steps;
	bool linear;     // slerp falls back to linear weights
	bool negate;     // quaternion dot < 0: second weight negated
	double theta, sin_theta;
};

__global__ void __launch_bounds__(BLOCK)
k_check_motions(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const double *__restrict__ a,
				const double *__restrict__ b, size_t m, size_t stride, int js, double max_step,
				const uint32_t *__restrict__ n_steps_override, uint32_t *__restrict__ hit_bits, double *__restrict__ toi,
				uint32_t *__restrict__ n_steps_out, uint32_t *__restrict__ uncertain_bits,
				unsigned long long *__restrict__ next_motion) {
	__shared__ Queue Q;
	__shared__ double s_a[WARPS][ROW_MAX], s_b[WARPS][ROW_MAX];
	const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	LocalCounters lc;
	unsigned long long motion = ~0ull;
	unsigned chunk = 0;
	MotionSetup ms{};
	unsigned long long n_motions_done = 0, n_uncertain = 0;

	auto fetch = [&]() {
		unsigned long long mi = 0;
		if (lane == 0) mi = atomicAdd(next_motion, 1ull);
		mi = __shfl_sync(0xffffffffu, mi, 0);
		motion = mi < m ? mi : ~0ull;
		chunk = 0;
		if (motion == ~0ull) return;
		__syncwarp();
		if (lane < stride) {
			s_a[warp][lane] = __ldg(a + motion * stride + lane);
			s_b[warp][lane] = __ldg(b + motion * stride + lane);
		}
		__syncwarp();
		const double *A = s_a[warp], *B = s_b[warp];
		// equal_weights_distance (distance.cpp:17-25) in the reference's operand order
		double dx = A[0] - B[0], dy = A[1] - B[1], dz = A[2] - B[2];
		double d = sqrt(dx * dx + dy * dy + dz * dz);
		double c = A[6] * B[6] + A[3] * B[3] + A[4] * B[4] + A[5] * B[5];
		double cc = fmin(fmax(c, -1.0), 1.0);
		d += acos(2.0 * cc * cc - 1.0);
		for (int i = 0; i < js; ++i) d += fabs(A[7 + i] - B[7 + i]);
		double x = d / max_step;
		double n = ceil(x);
		bool uncertain = fabs(x - rint(x)) < 1e-9 * fmax(1.0, x);
		if (!(n >= 0.0 && n < 4.0e9)) n = 0.0;  // non-finite or absurd distances: single check of the start state
		ms.n_steps = (unsigned) n;
		if (n_steps_override) {
			unsigned o = __ldg(n_steps_override + motion);
			if (o != 0xFFFFFFFFu) {
				ms.n_steps = o;
				uncertain = false;
			}
		}
		// Eigen slerp rule (Quaternion.cpp:12-20): dot over (x,y,z,w)
		double dq = A[3] * B[3] + A[4] * B[4] + A[5] * B[5] + A[6] * B[6];
		double ad = fabs(dq);
		ms.negate = dq < 0.0;
		ms.linear = ad >= 1.0 - 2.220446049250313e-16;
		ms.theta = ms.linear ? 0.0 : acos(ad);
		ms.sin_theta = ms.linear ? 1.0 : sin(ms.theta);
		if (lane == 0) {
			if (n_steps_out) n_steps_out[motion] = ms.n_steps;
			if (uncertain) {
				if (uncertain_bits) atomicOr(uncertain_bits + (motion >> 5), 1u << (motion & 31));
				n_uncertain++;
			}
		}
	};

	fetch();
	while (true) {
		if (tid < WARPS) Q.hitmask[tid] = 0;
		if (!__syncthreads_or(motion != ~0ull)) break;
		// ---- phase A: interpolate + FK for this lane's step --------------------------------------------------
		Pose link_tf[MGODPL_MAX_LINKS];
		double jv[MGODPL_MAX_JOINT_VALUES];
		bool active = false;
		const unsigned long long step = (unsigned long long) chunk * 32ull + lane;
		double t = 0.0;
		if (motion != ~0ull && step <= ms.n_steps) {
			const double *A = s_a[warp], *B = s_b[warp];
			Pose base;
			if (ms.n_steps == 0) {  // the reference evaluates 0/0 here; we check the start state once
				base.t = {A[0], A[1], A[2]};
				base.q = {A[3], A[4], A[5], A[6]};
				for (int k = 0; k < rb.n_vars; ++k) jv[k] = A[7 + k];
			} else {
				t = (double) step / (double) ms.n_steps;
				// Transform.h:69-78 (translation) and RobotState.cpp:19-21 (joints): two distinct lerp forms
				base.t = {(1.0 - t) * A[0] + t * B[0], (1.0 - t) * A[1] + t * B[1], (1.0 - t) * A[2] + t * B[2]};
				double w0, w1;
				if (ms.linear) {
					w0 = 1.0 - t, w1 = t;
				} else {
					w0 = sin((1.0 - t) * ms.theta) / ms.sin_theta;
					w1 = sin(t * ms.theta) / ms.sin_theta;
				}
				if (ms.negate) w1 = -w1;
				base.q = {w0 * A[3] + w1 * B[3], w0 * A[4] + w1 * B[4], w0 * A[5] + w1 * B[5], w0 * A[6] + w1 * B[6]};
				for (int k = 0; k < rb.n_vars; ++k) jv[k] = A[7 + k] * (1 - t) + B[7 + k] * t;
			}
			lc.states++;
			active = isfinite(base.t.x + base.t.y + base.t.z + base.q.x + base.q.y + base.q.z + base.q.w) &&
					 !robot_out_of_reach(sc, rb, base.t);
			if (active) fk_collision_links(rb, base, jv, link_tf);
		}
		// ---- phase A2/B: per box group, push survivors and drain the shared queue ------------------------------
		for (int g = 0; g < rb.n_boxes; g += ROUND_BOXES) {
			if (tid == 0) Q.count = 0;
			__syncthreads();
			if (active) {
				const int e = min(g + ROUND_BOXES, rb.n_boxes);
				for (int bi = g; bi < e; ++bi) {
					const BoxDev &bx = rb.boxes[bi];
					const Pose &l = link_tf[bx.link];
					Pose w = bx.tf_identity ? l : then(l, D3{bx.t[0], bx.t[1], bx.t[2]}, Q4{bx.q[0], bx.q[1], bx.q[2], bx.q[3]}, false);
					push_item(Q, sc, w.t, w.q, bx.half, tid, bi);
				}
			}
			__syncthreads();
			traverse_queue(Q, sc, lc, [&](unsigned i, unsigned) { return rb.boxes[Q.box[i]].half; });
			__syncthreads();
		}
		// ---- phase C: first colliding step of this chunk, early exit ---------------------------------------------
		if (motion != ~0ull) {
			const unsigned mask = Q.hitmask[warp];
			bool done = false;
			if (mask) {
				const unsigned first = __ffs(mask) - 1;
				if (lane == first) {
					atomicOr(hit_bits + (motion >> 5), 1u << (motion & 31));
					if (toi) toi[motion] = t;
				}
				done = true;
			} else if ((unsigned long long) chunk * 32ull + 31ull >= ms.n_steps) {
				if (lane == 0 && toi) toi[motion] = __longlong_as_double(0x7ff8000000000000ll);
				done = true;
			} else {
				chunk++;
			}
			if (done) {
				n_motions_done += lane == 0;
				fetch();
			}
		}
		__syncthreads();  // hitmask is reset at the top of the loop
	}
	flush_counters(sc, lc);
	if (sc.counters && lane == 0) {
		if (n_motions_done) atomicAdd(sc.counters + CNT_MOTIONS, n_motions_done);
		if (n_uncertain) atomicAdd(sc.counters + CNT_UNCERTAIN, n_uncertain);
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
__global__ void __launch_bounds__(BLOCK)
k_forward_kinematics(const __grid_constant__ RobotDev rb, const double *__restrict__ states, size_t n, size_t stride,
					 double *__restrict__ out) {
	const size_t i = (size_t) blockIdx.x * BLOCK + threadIdx.x;
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
static int g_sm_count = 0;
static int sm_count() {
	if (!g_sm_count) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
		if (g_sm_count <= 0) g_sm_count = 148;
	}
	return g_sm_count;
}
/// Persistent-style grids: a multiple of the SM count (148 on B200), several resident blocks per SM.
static unsigned grid_for(size_t tiles, int blocks_per_sm) {
	size_t cap = (size_t) sm_count() * blocks_per_sm;
	return (unsigned) (tiles < cap ? (tiles ? tiles : 1) : cap);
}

cudaError_t launch_check_states(const SceneDev &sc, const RobotDev &rb, const double *d_states, size_t n, size_t stride,
								uint32_t *d_hit_bits, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	ExplicitStates src{d_states, stride};
	size_t smem = (size_t) BLOCK * stride * sizeof(double);
	static bool attr = false;
	if (!attr) {
		cudaFuncSetAttribute(k_check_states<ExplicitStates>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLOCK * ROW_MAX * 8);
		attr = true;
	}
	k_check_states<ExplicitStates><<<grid_for((n + BLOCK - 1) / BLOCK, 6), BLOCK, smem, stream>>>(sc, rb, src, 1, n, d_hit_bits);
	return cudaGetLastError();
}

cudaError_t launch_check_boxes(const SceneDev &sc, const double *d_boxes, size_t n, uint32_t *d_hit_bits, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	k_check_boxes<<<grid_for((n + BLOCK - 1) / BLOCK, 6), BLOCK, 0, stream>>>(sc, d_boxes, n, d_hit_bits);
	return cudaGetLastError();
}

cudaError_t launch_check_motions(const SceneDev &sc, const RobotDev &rb, const double *d_a, const double *d_b, size_t m,
								 size_t stride, int js, double max_step, const uint32_t *d_override, uint32_t *d_hit_bits,
								 double *d_toi, uint32_t *d_n_steps, uint32_t *d_uncertain, unsigned long long *d_next_motion,
								 cudaStream_t stream) {
	if (!m) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(d_hit_bits, 0, ((m + 31) / 32) * 4, stream);
	if (e != cudaSuccess) return e;
	if (d_uncertain && (e = cudaMemsetAsync(d_uncertain, 0, ((m + 31) / 32) * 4, stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(d_next_motion, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
	k_check_motions<<<grid_for((m + WARPS - 1) / WARPS, 6), BLOCK, 0, stream>>>(
			sc, rb, d_a, d_b, m, stride, js, max_step, d_override, d_hit_bits, d_toi, d_n_steps, d_uncertain, d_next_motion);
	return cudaGetLastError();
}

cudaError_t launch_sample_goals(const SceneDev &sc, const RobotDev &rb, const double *d_targets, size_t f,
								const double *d_draws, const double *d_limits, size_t s, unsigned long long seed,
								double distance_from_target, uint32_t *d_hit_bits, uint32_t *d_collisions,
								double *d_states_out, cudaStream_t stream) {
	if (!f || !s) return cudaSuccess;
	GoalStates src{d_targets, d_draws, d_limits, seed, distance_from_target, d_states_out, s};
	size_t tiles = f * ((s + BLOCK - 1) / BLOCK);
	k_check_states<GoalStates><<<grid_for(tiles, 6), BLOCK, 0, stream>>>(sc, rb, src, f, s, d_hit_bits);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	if (d_collisions) {
		k_count_rows<<<(unsigned) f, 128, 0, stream>>>(d_hit_bits, f, (s + 31) / 32, d_collisions);
		e = cudaGetLastError();
	}
	return e;
}

cudaError_t launch_forward_kinematics(const RobotDev &rb, const double *d_states, size_t n, size_t stride, double *d_out,
									  cudaStream_t stream) {
	if (!n) return cudaSuccess;
	k_forward_kinematics<<<(unsigned) ((n + BLOCK - 1) / BLOCK), BLOCK, 0, stream>>>(rb, d_states, n, stride, d_out);
	return cudaGetLastError();
}

}  // namespace mgodpl_b200
