// Candidate expansion shared by the stage A2 kernels (query_kernels.cu) and the fused pipeline kernel (pipeline.cuh):
// joint-value sources, forward kinematics over the host-unrolled op list, the per-box culls, and the box sinks.
#pragma once

namespace mgodpl_b200 {

// ---- joint-value sources: a joint value is fetched when its FK op needs it (no per-thread joint array) ------------------
struct JointsFromRow {
	const double *p;
	__device__ __forceinline__ double operator()(int var) const { return __ldg(p + var); }
};
struct JointsLerp {  // RobotState.cpp:19-21: a (1 - t) + b t
	const double *a, *b;
	double t;
	bool moving;
	__device__ __forceinline__ double operator()(int var) const {
		return moving ? __ldg(a + var) * (1 - t) + __ldg(b + var) * t : __ldg(a + var);
	}
};
struct JointsGoal {  // genUprightState's draws (goal_sampling.cpp:13-49): caller-supplied, or a counter-based generator
	const double *draw_row, *limits;
	unsigned long long seed, ctr;
	__device__ __forceinline__ double operator()(int var) const {
		if (draw_row) return __ldg(draw_row + 1 + var);
		return limits[2 * var] + (limits[2 * var + 1] - limits[2 * var]) * u01_from_bits(mix64(seed ^ mix64(ctr + 1 + var)));
	}
};

template<class Joints>
__device__ __forceinline__ Pose fk_apply_j(const FkOp &op, const Pose &parent, const Joints &joints) {
	Pose p = then(parent, D3{op.pre_t[0], op.pre_t[1], op.pre_t[2]}, Q4{op.pre_q[0], op.pre_q[1], op.pre_q[2], op.pre_q[3]},
				  op.flags & FK_PRE_ROT_ID);
	if (op.var >= 0) {
		double s, c;
		sincos(joints(op.var) / 2.0, &s, &c);
		if (op.flags & FK_INVERT) s = -s;
		p.q = qmul(p.q, Q4{op.axis[0] * s, op.axis[1] * s, op.axis[2] * s, c});
	}
	return then(p, D3{op.post_t[0], op.post_t[1], op.post_t[2]}, Q4{op.post_q[0], op.post_q[1], op.post_q[2], op.post_q[3]},
				op.flags & FK_POST_ROT_ID);
}

/// Out-of-line copy for the fused kernel, whose instruction footprint matters more than a call.
template<class Joints>
__device__ __noinline__ Pose fk_apply_j_call(const FkOp &op, const Pose &parent, const Joints &joints) {
	return fk_apply_j(op, parent, joints);
}

/// The culls a box goes through before it is worth a queue item / ring slot: scene bounds, clearance grid, the root record's children.
__device__ __forceinline__ bool box_may_touch_scene(const SceneDev &sc, D3 c, Q4 q, const double *half, const BoxDev *spheres) {
	if (!sc.n_nodes) return false;  // empty scene
	const CullBox cb = make_cull_box(sc, c, q, half);
	if (!cull_overlap(cb, sc.root_c[0], sc.root_c[1], sc.root_c[2], sc.root_h[0], sc.root_h[1], sc.root_h[2])) return false;
	if (sc.clear) {
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
		if (free_of_scene) return false;
	} else if (sc.wide) {
		bool any = false;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			any |= __float_as_int(sc.top[2 * k].w) != WIDE_EMPTY &&
				   cull_overlap(cb, sc.top[2 * k].x, sc.top[2 * k].y, sc.top[2 * k].z, sc.top[2 * k + 1].x, sc.top[2 * k + 1].y, sc.top[2 * k + 1].z);
		if (!any) return false;
	}
	return true;
}

__device__ __noinline__ bool box_may_touch_scene_call(const SceneDev &sc, D3 c, Q4 q, const double *half, const BoxDev *spheres) {
	return box_may_touch_scene(sc, c, q, half, spheres);
}

/// FK for the links that carry collision geometry, one `emit` per box (check_robot_collision, collision_detection.cpp:57-80;
/// check_link_collision :16-55).  `Sink::emit(want, centre, orientation, owner, step, box)` takes the boxes: the fused
/// kernel's per-warp ring (pipeline.cuh) or the box queue in HBM (QueueSink).  All lanes of a warp may run the same loop
/// (the robot is the same for everyone), so that every emit is a convergent point for sinks that need one; lanes whose
/// candidate is dead (`alive` false) skip the arithmetic.  Only the boxes
/// [box_lo, box_hi) are emitted (robots with more boxes than a lane owns ring slots are expanded in several groups).
/// CHAIN: every collision op continues from the previous one, so the running pose stays in registers and no link table
/// exists at all (every procedural drone; RobotDev::chain).
template<bool CHAIN, bool OUTLINE, class Joints, class Sink>
__device__ __forceinline__ void expand_boxes(const SceneDev &sc, const RobotDev &rb, Sink &w, bool alive, const Pose &base, const Joints &joints,
											 unsigned owner, unsigned step, int box_lo, int box_hi) {
	auto boxes_of = [&](const Pose &link, int lo, int hi) {
		lo = max(lo, box_lo), hi = min(hi, box_hi);
		for (int bi = lo; bi < hi; ++bi) {
			bool want = false;
			Pose p = link;
			if (alive) {
				const BoxDev &bx = rb.boxes[bi];
				if (!bx.tf_identity) p = then(link, D3{bx.t[0], bx.t[1], bx.t[2]}, Q4{bx.q[0], bx.q[1], bx.q[2], bx.q[3]}, false);
				want = OUTLINE ? box_may_touch_scene_call(sc, p.t, p.q, bx.half, &bx) : box_may_touch_scene(sc, p.t, p.q, bx.half, &bx);
			}
			w.emit(want, p.t, p.q, owner, step, (unsigned) bi);
		}
	};
	boxes_of(base, rb.root_box_begin, rb.root_box_end);
	Pose cur = base;
	Pose link_tf[CHAIN ? 1 : MGODPL_MAX_LINKS];
	if (!CHAIN) link_tf[rb.root] = base;
	for (int i = 0; i < rb.n_ops; ++i) {
		const FkOp &op = rb.ops[i];
		if (!(op.flags & FK_FOR_COLLISION)) continue;
		if (op.box_begin >= box_hi && op.box_begin < op.box_end) break;  // boxes are sorted by op: nothing of this group is left
		if (alive) {
			const Pose &parent = (CHAIN || (op.flags & FK_PARENT_IS_PREV)) ? cur : link_tf[op.parent];
			cur = OUTLINE ? fk_apply_j_call(op, parent, joints) : fk_apply_j(op, parent, joints);
			if (!CHAIN && (op.flags & FK_STORE)) link_tf[op.child] = cur;
		}
		boxes_of(cur, op.box_begin, op.box_end);
	}
}

/// One interpolation step of one motion as a candidate: interpolate(a, b, t) (RobotState.cpp:14-24, Transform.h:69-78, Eigen
/// slerp rule of Quaternion.cpp:12-20 with the per-motion constants of stage A1) and the exact reach cull.  Returns false
/// when nothing of the robot can touch the scene at this step (or an earlier step of the motion already collides).
__device__ __forceinline__ bool make_motion_step(const SceneDev &sc, const RobotDev &rb, const double *__restrict__ A, const double *__restrict__ B,
												 unsigned n_steps, const SlerpConst *__restrict__ slerp_k, unsigned step, const unsigned *first_step_of_motion,
												 Pose &base, JointsLerp &joints, LocalCounters &lc) {
	// the "an earlier step of this motion already collides" probe travels with the endpoint rows: one round trip, not two
	const unsigned first_hit = first_step_of_motion ? __ldcg(first_step_of_motion) : NO_HIT;
	double a7[7], b7[7];
#pragma unroll
	for (int j = 0; j < 7; ++j) a7[j] = __ldg(A + j), b7[j] = __ldg(B + j);
	if (first_hit <= step) return false;
	lc.states++;
	// for n_steps == 0 the reference evaluates t = 0/0; here the start state is checked once.
	const double t = n_steps ? (double) step / (double) n_steps : 0.0;
	const D3 tr{(1.0 - t) * a7[0] + t * b7[0], (1.0 - t) * a7[1] + t * b7[1], (1.0 - t) * a7[2] + t * b7[2]};
	if (robot_out_of_reach(sc, rb, tr, false)) return false;
	base = Pose{tr, Q4{a7[3], a7[4], a7[5], a7[6]}};
	if (n_steps) {
		const SlerpConst sk = load_slerp(slerp_k);
		double w0, w1;
		if (sk.theta < 0.0) {
			w0 = 1.0 - t, w1 = t;
		} else {
			double s1, c1;
			sincos(t * sk.theta, &s1, &c1);
			w0 = c1 - sk.cot * s1;
			w1 = s1 * fabs(sk.inv_sin);
		}
		if (signbit(sk.inv_sin)) w1 = -w1;
		base.q = Q4{w0 * a7[3] + w1 * b7[3], w0 * a7[4] + w1 * b7[4], w0 * a7[5] + w1 * b7[5], w0 * a7[6] + w1 * b7[6]};
	}
	joints = JointsLerp{A + 7, B + 7, t, n_steps != 0};
	return isfinite(base.q.x + base.q.y + base.q.z + base.q.w);
}

/// One goal sample as a candidate: genGoalStateUniform (goal_sampling.cpp:51-79).  The upright state comes from
/// caller-supplied draws (yaw, joint values — goal_sampling.cpp:13-49) or from a counter-based generator; FK to the end
/// effector along the op path that leads to it; base shifted so the end effector lands on the target.  Returns false when the
/// robot cannot touch the scene there.  `index` = row * per_row + sample; `owner` receives the result bit index.
__device__ __forceinline__ bool make_goal_sample(const SceneDev &sc, const RobotDev &rb, const double *__restrict__ targets, size_t per_row,
												 const double *__restrict__ draws, unsigned long long seed, double distance_from_target,
												 double *__restrict__ states_out, unsigned long long index, Pose &base, JointsGoal &jg, unsigned &owner,
												 LocalCounters &lc) {
	const size_t row = (size_t) (index / per_row), s = (size_t) (index - row * per_row);
	const int nv = rb.n_vars;
	jg = JointsGoal{nullptr, rb.limits, seed, 0ull};
	double yaw;
	if (draws) {
		jg.draw_row = draws + s * (size_t) (1 + nv);
		yaw = __ldg(jg.draw_row);
	} else {
		jg.ctr = index * (unsigned long long) (1 + nv);
		yaw = -M_PI + 2.0 * M_PI * u01_from_bits(mix64(seed ^ mix64(jg.ctr)));
	}
	// genUprightState: base at the origin, rotation about z by yaw (Quaternion.h:53-62 with axis UnitZ)
	double sy, cy;
	sincos(yaw / 2.0, &sy, &cy);
	const Q4 yq{0 * sy, 0 * sy, 1 * sy, cy};
	Pose ee{{0, 0, 0}, yq};
	for (int k = 0; k < rb.n_ee_path; ++k) ee = fk_apply_j(rb.ops[rb.ee_path[k]], ee, jg);
	D3 delta = D3{__ldg(targets + 3 * row), __ldg(targets + 3 * row + 1), __ldg(targets + 3 * row + 2)} - ee.t;
	if (distance_from_target > 0.0) delta = delta + qrot(ee.q, D3{0.0, -distance_from_target, 0.0});
	base = Pose{D3{0, 0, 0} + delta, yq};
	if (states_out) {
		double *o = states_out + index * (size_t) (7 + nv);
		o[0] = base.t.x, o[1] = base.t.y, o[2] = base.t.z;
		o[3] = base.q.x, o[4] = base.q.y, o[5] = base.q.z, o[6] = base.q.w;
		for (int k = 0; k < nv; ++k) o[7 + k] = jg(k);
	}
	lc.states++;
	owner = (unsigned) (row * (((per_row + 31) / 32) * 32) + s);
	return !robot_out_of_reach(sc, rb, base.t);
}

/// Queue-overflow path of the sink below, out of line (rare; its cull box and traversal state stay out of the expansion kernels).
__device__ __noinline__ void traverse_box_now(const SceneDev &sc, const RobotDev &rb, const Results &res, D3 c, Q4 q, unsigned owner, unsigned step,
											  unsigned box, LocalCounters &lc) {
	if (already_decided(res, owner, step)) return;
	const BoxDev &bx = rb.boxes[box];
	const CullBox cb = make_cull_box(sc, c, q, bx.half);
	if (box_hits_scene(sc, cb, c, q, bx.half, lc, &bx, rb.shape_data)) report_hit(res, owner, step);
}

/// Sink of the two-kernel path: the sharded box queue in HBM (emit_box's append, without its culls).  A box that finds the
/// queue full is traversed on the spot — but not from inside the FK loop: it is parked in `deferred` (local memory, touched on
/// that path only) and the kernel calls flush() once the candidate is done, where nothing but loop counters is live.  (The call
/// sat inside emit() before; the callee tree of the exact tests then shaped the register allocation of the whole expansion.)
struct DeferredBox {
	D3 c;
	Q4 q;
	unsigned owner, step, box;
};
struct QueueSink {
	const SceneDev &sc;
	const QueueDev &Q;
	const Results &res;
	const RobotDev &rb;
	LocalCounters &lc;
	DeferredBox *deferred;  // MGODPL_MAX_BOXES entries
	int n_deferred;
	__device__ __forceinline__ void emit(bool want, D3 c, Q4 q, unsigned owner, unsigned step, unsigned box) {
		if (!want) return;
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
		} else {
			deferred[n_deferred++] = DeferredBox{c, q, owner, step, box};
		}
	}
	__device__ __forceinline__ void flush() {
		if (n_deferred == 0) return;
		for (int i = 0; i < n_deferred; ++i) traverse_box_now(sc, rb, res, deferred[i].c, deferred[i].q, deferred[i].owner, deferred[i].step, deferred[i].box, lc);
		n_deferred = 0;
	}
};

}  // namespace mgodpl_b200
