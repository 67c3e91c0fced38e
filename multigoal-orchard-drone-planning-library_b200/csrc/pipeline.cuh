// Stage A2 + B fused: one persistent kernel expands candidate robot states into link boxes AND walks the BVH with them.
//
// Included by query_kernels.cu (after the shared helpers).  Every warp is an independent worker with its own region
// of shared memory; nothing a warp produces ever travels through HBM:
//
//   expand   all 32 lanes take one candidate each (a surviving state, one interpolation step of a motion, a goal sample,
//            an explicit box): interpolate / FK in fp64, cull each link box (scene bounds, clearance grid, the root
//            record's children) and drop the survivors into the warp's RING of 64-byte box items in shared memory.
//            Free ring slots sit on a per-warp stack; every emit and every release happens at a point all 32 lanes reach
//            together, so slots are handed out and returned by ballot rank, without atomics.  A FIFO of slot numbers
//            ("pending") hands the boxes to whichever lanes are free.
//   descend  every lane that holds a box visits one BVH record per step (4-wide: up to four fp32 OBB-vs-AABB tests per
//            dependent fetch), keeps the nearest overlapping internal child, pushes the others on its stack (shared
//            memory, [depth][lane]; a small local array only beyond STACK_SMEM entries) and appends overlapping LEAF
//            children to the warp's leaf FIFO — a lane never stops at a leaf.
//   leaves   as soon as the FIFO holds 32 (box, leaf) pairs the whole warp tests 32 of them at once, one pair per lane,
//            whichever lane walked the box: rebuild the cull box from the item, fp32 pre-filter, exact fp64 test for
//            what it cannot decide.  A hit is reported with an atomic and marks the ring slot dead; the lane walking
//            that box drops it at its next step.
//   refill   lanes whose descent is over take the next pending box at once; a ring slot goes back to its owner when the
//            descent is over AND the last of its leaf pairs has been tested (whichever happens last releases it).
//
// The any-hit early exit of the one-box-per-lane traversal ("park at the first overlapping leaf, test, stop on a hit")
// becomes "keep walking until the next leaf batch": only ~2 % of the boxes of a PRM-style batch hit anything, so the
// extra visits are noise, while the three kinds of work each run with (nearly) full warps.
#pragma once

namespace mgodpl_b200 {

constexpr int BLOCK_P = 128;          // 4 warps per block, each an independent worker
constexpr int SPL = 4;                // boxes of one candidate expanded per round (robots with more boxes take several rounds)
constexpr int RING = 128;             // box items per warp
constexpr int STACK_SMEM = 24;        // per-lane stack entries kept in shared memory
constexpr int STACK_LOCAL = 48;       // beyond that: local memory (never touched by trees of sane depth)
constexpr int LEAF_FIFO = 256;        // power of two, >= 31 + 4 * 32
constexpr unsigned CLAIM_ROUNDS = 4;  // large batches: candidates a warp claims per cursor atomic = 32 * CLAIM_ROUNDS

struct __align__(16) WarpMem {
	Item ring[RING];
	int stack[STACK_SMEM][32];
	int leaf_ref[LEAF_FIFO];
	unsigned char leaf_slot[LEAF_FIFO];
	int pairs[RING];            // (box, leaf) pairs of the box in this slot still waiting in the leaf FIFO
	unsigned char pend[RING];
	unsigned char dead[RING];   // a leaf test found a hit for the box in this slot (or its owner is decided already)
	unsigned char done[RING];   // the descent of the box in this slot is over; the slot goes back once `pairs` reaches 0
	unsigned char free_slots[RING];  // stack of free ring slots
};

enum SourceKind : int { SRC_STATES = 0, SRC_MOTION_STEPS = 1, SRC_GOALS = 2, SRC_BOXES = 3, SRC_QUEUE = 4 };

/// Where the candidates come from (one struct for all kinds; unused members stay zero).
struct SourceArgs {
	// SRC_STATES / SRC_MOTION_STEPS: survivor lists of stage A1 (one for states; head then tail for motions)
	SurvivorsDev list[2];
	int n_lists;
	const double *states;  // SRC_STATES rows
	const double *a, *b;   // SRC_MOTION_STEPS endpoint rows
	size_t stride;
	const uint32_t *n_steps;
	const SlerpConst *slerp;
	// SRC_GOALS
	const double *targets, *draws;
	size_t rows, per_row;
	unsigned long long seed;
	double distance_from_target;
	double *states_out;
	// SRC_BOXES (and SRC_QUEUE with explicit boxes: the half extents live here)
	const double *boxes;
	size_t total;  // SRC_GOALS / SRC_BOXES: number of candidates
	// SRC_QUEUE: boxes already expanded and culled by a stage A2 kernel (the two-kernel path), sharded list in HBM
	QueueDev queue;
};

struct PipeTuning {
	int refill_min;    // free lanes that trigger a refill while others are still walking
	unsigned leaf_min; // (box, leaf) pairs that trigger a leaf batch (<= 32)
	int no_prefilter;  // MGODPL_NO_PREFILTER=1: every leaf triangle goes to the fp64 test
	unsigned claim;    // candidates a warp claims per cursor atomic (a multiple of 32)
};

/// Per-lane view of the warp's worker state.  Members marked (u) hold the same value in every lane.
struct Worker {
	WarpMem *wm;
	unsigned lane;
	unsigned n_free;               // (u) free ring slots (on wm->free_slots)
	unsigned pend_head, pend_tail; // (u) FIFO of slots waiting for a lane
	unsigned leaf_head, leaf_tail; // (u) FIFO of (slot, leaf) pairs waiting for a test

	/// Producer side, called by all 32 lanes together: lanes with `want` drop their box into a free slot.
	__device__ __forceinline__ void emit(bool want, D3 c, Q4 q, unsigned owner, unsigned step, unsigned box) {
		const unsigned m = __ballot_sync(0xffffffffu, want);
		if (!m) return;
		if (want) {
			const unsigned rank = __popc(m & ((1u << lane) - 1u));
			const unsigned slot = wm->free_slots[n_free - 1u - rank];
			double2 *dst = reinterpret_cast<double2 *>(wm->ring + slot);
			dst[0] = make_double2(c.x, c.y);
			dst[1] = make_double2(c.z, q.x);
			dst[2] = make_double2(q.y, q.z);
			dst[3] = make_double2(q.w, __hiloint2double((int) (step | box << 28), (int) owner));
			wm->dead[slot] = 0, wm->done[slot] = 0, wm->pairs[slot] = 0;
			wm->pend[(pend_tail + rank) & (RING - 1)] = (unsigned char) slot;
		}
		n_free -= __popc(m), pend_tail += __popc(m);
		__syncwarp();
	}
	/// Called by all 32 lanes together: lanes with `rel` give slot `s` back.
	__device__ __forceinline__ void release(bool rel, unsigned s) {
		const unsigned m = __ballot_sync(0xffffffffu, rel);
		if (!m) return;
		if (rel) wm->free_slots[n_free + __popc(m & ((1u << lane) - 1u))] = (unsigned char) s;
		n_free += __popc(m);
		__syncwarp();
	}
};

/// Exact phase of a leaf test with the pose passed by value (the pipeline's items live in shared memory).
__device__ __noinline__ bool exact_leaf_test_pose(const SceneDev &sc, D3 c, Q4 q, const double *half, int ref, unsigned unknown, unsigned *n_tests) {
	const unsigned info = (unsigned) (-1 - ref), first = info >> 4, cnt = (info & 15u) + 1u;
	const ExactBox eb = make_exact_box(c, q, half);
	for (unsigned k = 0; k < cnt; ++k) {
		if (!(unknown >> k & 1u)) continue;
		++*n_tests;
		if (box_triangle_hit(eb, sc.tris + 9ull * (first + k))) return true;
	}
	return false;
}

template<int KIND, bool CHAIN, bool WIDE>
__global__ void __launch_bounds__(BLOCK_P, 4)
k_pipeline(const __grid_constant__ SceneDev sc, const __grid_constant__ RobotDev rb, const __grid_constant__ SourceArgs src, unsigned *cursor,
		   Results res, PipeTuning tune) {
	constexpr unsigned FULL = 0xffffffffu;
	extern __shared__ __align__(16) unsigned char pipe_smem[];
	__shared__ unsigned s_counts[2][SHARDS];
	const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
	WarpMem *wm = reinterpret_cast<WarpMem *>(pipe_smem) + (threadIdx.x >> 5);
	LocalCounters lc;

	// ---- candidate index space ------------------------------------------------------------------------------------------------
	unsigned long long range0 = 0, total = 0;  // (u) flat slots of list 0; all candidates
	if (KIND == SRC_STATES || KIND == SRC_MOTION_STEPS) {
		for (int l = 0; l < src.n_lists; ++l)
			if (threadIdx.x < SHARDS) s_counts[l][threadIdx.x] = __ldcg(src.list[l].counts + threadIdx.x * COUNTER_STRIDE);
		__syncthreads();
		range0 = slot_range(s_counts[0], src.list[0].cap);
		total = range0 + (src.n_lists > 1 ? slot_range(s_counts[1], src.list[1].cap) : 0u);
	} else if (KIND == SRC_QUEUE) {
		if (threadIdx.x < SHARDS) s_counts[0][threadIdx.x] = __ldcg(src.queue.counts + threadIdx.x * COUNTER_STRIDE);
		__syncthreads();
		total = slot_range(s_counts[0], src.queue.cap);
	} else {
		total = src.total;
	}
	constexpr bool ONE_BOX = KIND == SRC_BOXES || KIND == SRC_QUEUE;  // a candidate is a single box
	const int n_groups = ONE_BOX ? 1 : max(1, (rb.n_boxes + SPL - 1) / SPL);  // box groups per candidate
	const int group_boxes = ONE_BOX ? 1 : min(rb.n_boxes, SPL);
	total *= (unsigned) n_groups;
	// few candidates per warp: claim one round at a time so that the work spreads over all the warps of the grid
	const unsigned claim = total > (unsigned long long) gridDim.x * (BLOCK_P / 32) * 32ull * 16ull ? tune.claim : 32u;

	Worker w{wm, lane, (unsigned) RING, 0u, 0u, 0u, 0u};
	for (int k = lane; k < RING; k += 32) wm->free_slots[k] = (unsigned char) k;
	unsigned long long chunk_lo = 0, chunk_hi = 0;  // (u) unclaimed part of this warp's chunk of candidates
	bool exhausted = total == 0;                     // (u)
	// lane state: the box being walked (cur == REF_END: none)
	CullBox cb{};
	int cur = REF_END, sp = 0;
	unsigned slot = 0;
	int lstack[STACK_LOCAL];
	unsigned box_nodes = 0, max_box_nodes = 0, stalls = 0;
	__syncwarp();

	auto push = [&](int v) {
		if (sp < STACK_SMEM) wm->stack[sp][lane] = v;
		else lstack[sp - STACK_SMEM] = v;
		++sp;
	};
	auto pop = [&]() {
		--sp;
		return sp < STACK_SMEM ? wm->stack[sp][lane] : lstack[sp - STACK_SMEM];
	};
	auto half_of = [&](unsigned owner, unsigned step_box, double *buf) -> const double * {
		if (KIND == SRC_BOXES || (KIND == SRC_QUEUE && src.boxes)) {
			const double *bx = src.boxes + (size_t) owner * 10;
			buf[0] = __ldg(bx + 7) * 0.5, buf[1] = __ldg(bx + 8) * 0.5, buf[2] = __ldg(bx + 9) * 0.5;
			return buf;
		}
		return rb.boxes[step_box >> 28].half;
	};
	/// One batch of leaf tests: pair `leaf_head + lane` for lane < batch, whichever lane walked the box.
	auto leaf_batch = [&](unsigned batch) {
		bool last_pair = false;
		unsigned s = 0;
		if (lane < batch) {
			const unsigned e = (w.leaf_head + lane) & (LEAF_FIFO - 1);
			s = wm->leaf_slot[e];
			const int ref = wm->leaf_ref[e];
			if (!wm->dead[s]) {
				const double2 *it = reinterpret_cast<const double2 *>(wm->ring + s);
				const double2 v0 = it[0], v1 = it[1], v2 = it[2], v3 = it[3];
				const unsigned owner = (unsigned) __double2loint(v3.y), step_box = (unsigned) __double2hiint(v3.y), step = step_box & 0x0FFFFFFFu;
				if (already_decided(res, owner, step)) {
					wm->dead[s] = 1;
				} else {
					double hb[3];
					const double *half = half_of(owner, step_box, hb);
					const D3 c{v0.x, v0.y, v1.x};
					const Q4 q{v1.y, v2.x, v2.y, v3.x};
					const CullBox tb = make_cull_box(sc, c, q, half);
					unsigned unknown = 0xffffu;
					bool hit = tune.no_prefilter ? false : leaf_prefilter(sc, tb, ref, &unknown);
					if (!hit && unknown) {
						unsigned n_exact = 0;
						hit = exact_leaf_test_pose(sc, c, q, half, ref, unknown, &n_exact);
						lc.leaf += n_exact;
					}
					if (hit) {
						report_hit(res, owner, step);
						wm->dead[s] = 1;
					}
				}
			}
			// last pair of a box whose descent is over: the slot goes back on the free stack
			last_pair = atomicSub(&wm->pairs[s], 1) == 1 && wm->done[s];
		}
		w.release(last_pair, s);
		w.leaf_head += batch;
	};

	for (;;) {
		__syncwarp();
		const unsigned n_leaf = w.leaf_tail - w.leaf_head, n_pend = w.pend_tail - w.pend_head;
		const unsigned walking = __ballot_sync(FULL, cur != REF_END), n_free = 32u - __popc(walking);
		// ---- leaves: 32 pairs waiting -> one full-warp batch ----------------------------------------------------------------
		if (n_leaf >= tune.leaf_min) {
			leaf_batch(min(n_leaf, 32u));
			continue;
		}
		// ---- refill: free lanes take pending boxes ------------------------------------------------------------------------------
		if (n_pend && n_free && ((int) n_free >= tune.refill_min || !walking || n_pend >= 64u)) {
			const unsigned rank = __popc(~walking & lt);
			bool drop = false;
			if (cur == REF_END && rank < n_pend) {
				slot = wm->pend[(w.pend_head + rank) & (RING - 1)];
				const double2 *it = reinterpret_cast<const double2 *>(wm->ring + slot);
				const double2 v0 = it[0], v1 = it[1], v2 = it[2], v3 = it[3];
				const unsigned owner = (unsigned) __double2loint(v3.y), step_box = (unsigned) __double2hiint(v3.y);
				unsigned first_hit = NO_HIT;
				if (res.mode != RESULT_BITS) first_hit = __ldcg(res.first_step + owner);  // issued first, consumed last
				double hb[3];
				const double *half = half_of(owner, step_box, hb);
				cb = make_cull_box(sc, D3{v0.x, v0.y, v1.x}, Q4{v1.y, v2.x, v2.y, v3.x}, half);
				// motions: a box of a step that can no longer be the first colliding one is dropped at once
				if (res.mode == RESULT_BITS || !(first_hit <= (step_box & 0x0FFFFFFFu))) {
					cur = 0, sp = 0;
					lc.boxes++;
					max_box_nodes = max(max_box_nodes, box_nodes), box_nodes = 0;
				} else {
					drop = true;
				}
			}
			w.release(drop, slot);
			w.pend_head += min(n_pend, n_free);
			stalls = 0;
			continue;
		}
		// ---- expand: keep at least a warp's worth of boxes pending while candidates remain ----------------------------------------
		// (a full round wants 32 x group_boxes free slots; narrower rounds only when the lanes would otherwise starve)
		const unsigned can_lanes = min(32u, w.n_free / (unsigned) max(group_boxes, 1));
		if (!exhausted && n_pend < 32u && (can_lanes == 32u || (can_lanes && !n_leaf && (!n_pend || !walking)))) {
			const bool can = lane < can_lanes;
			const unsigned can_mask = can_lanes == 32u ? FULL : (1u << can_lanes) - 1u;
			{
				stalls = 0;
				if (chunk_lo == chunk_hi) {
					unsigned long long lo = 0;
					if (lane == 0) lo = atomicAdd(cursor, claim);
					lo = __shfl_sync(FULL, lo, 0);
					chunk_lo = min(lo, total), chunk_hi = min(lo + claim, total);
					if (chunk_lo >= total) {
						exhausted = true;
						continue;
					}
				}
				const unsigned avail = (unsigned) min(chunk_hi - chunk_lo, (unsigned long long) 32u), rank = __popc(can_mask & lt);
				bool alive = can && rank < avail;
				const unsigned long long idx = chunk_lo + rank;
				chunk_lo += min(avail, (unsigned) __popc(can_mask));
				const unsigned long long cand = idx / (unsigned) n_groups;
				const int box_lo = (int) (idx % (unsigned) n_groups) * SPL, box_hi = box_lo + SPL;
				Pose base{};
				unsigned owner = 0, step = 0;
				if (KIND == SRC_STATES) {
					const double *row = nullptr;
					if (alive) {
						alive = slot_filled(s_counts[0], (unsigned) cand);
						if (alive) {
							owner = __ldcg(&src.list[0].entries[cand]).x;
							row = src.states + (size_t) owner * src.stride;
							double r7[7];
#pragma unroll
							for (int j = 0; j < 7; ++j) r7[j] = __ldg(row + j);
							alive = finite7(r7);
							base = Pose{D3{r7[0], r7[1], r7[2]}, Q4{r7[3], r7[4], r7[5], r7[6]}};
						}
					}
					expand_boxes<CHAIN, true>(sc, rb, w, alive, base, JointsFromRow{row + 7}, owner, 0u, box_lo, box_hi);  // (stage A1 counted the states)
				} else if (KIND == SRC_MOTION_STEPS) {
					JointsLerp jl{nullptr, nullptr, 0.0, false};
					if (alive) {
						const bool tail = cand >= range0;
						const unsigned k = (unsigned) (tail ? cand - range0 : cand);
						alive = slot_filled(s_counts[tail ? 1 : 0], k);
						if (alive) {
							const uint2 e = __ldcg(&src.list[tail ? 1 : 0].entries[k]);
							owner = e.x, step = e.y;
							LocalCounters once;
							alive = make_motion_step(sc, rb, src.a + (size_t) owner * src.stride, src.b + (size_t) owner * src.stride, __ldg(src.n_steps + owner),
													 src.slerp + owner, step, res.first_step + owner, base, jl, once);
							if (box_lo == 0) lc.states += once.states;
						}
					}
					expand_boxes<CHAIN, true>(sc, rb, w, alive, base, jl, owner, step, box_lo, box_hi);
				} else if (KIND == SRC_GOALS) {
					JointsGoal jg{nullptr, rb.limits, src.seed, 0ull};
					if (alive) {
						LocalCounters once;
						alive = make_goal_sample(sc, rb, src.targets, src.per_row, src.draws, src.seed, src.distance_from_target,
												 box_lo == 0 ? src.states_out : nullptr, cand, base, jg, owner, once);
						if (box_lo == 0) lc.states += once.states;
					}
					expand_boxes<CHAIN, true>(sc, rb, w, alive, base, jg, owner, 0u, box_lo, box_hi);
				} else if (KIND == SRC_QUEUE) {  // a box queued by a stage A2 kernel: copy it on chip
					bool want = alive && slot_filled(s_counts[0], (unsigned) cand);
					unsigned step_box = 0;
					if (want) {
						const double2 *it = reinterpret_cast<const double2 *>(src.queue.items + cand);
						const double2 v0 = __ldcg(it), v1 = __ldcg(it + 1), v2 = __ldcg(it + 2), v3 = __ldcg(it + 3);
						base = Pose{D3{v0.x, v0.y, v1.x}, Q4{v1.y, v2.x, v2.y, v3.x}};
						owner = (unsigned) __double2loint(v3.y), step_box = (unsigned) __double2hiint(v3.y);
					}
					w.emit(want, base.t, base.q, owner, step_box & 0x0FFFFFFFu, step_box >> 28);
				} else {  // SRC_BOXES: check_link_collision (collision_detection.cpp:16-55), n x (tf7, size3)
					bool want = false;
					if (alive) {
						double v[10];
#pragma unroll
						for (int k = 0; k < 10; ++k) v[k] = __ldg(src.boxes + cand * 10 + k);
						const double half[3] = {v[7] * 0.5, v[8] * 0.5, v[9] * 0.5};
						base = Pose{D3{v[0], v[1], v[2]}, Q4{v[3], v[4], v[5], v[6]}};
						want = finite7(v) && box_may_touch_scene_call(sc, base.t, base.q, half, nullptr);
						owner = (unsigned) cand;
					}
					w.emit(want, base.t, base.q, owner, 0u, 0u);
				}
				continue;
			}
		}
		// ---- descend: one record per walking lane, until something else becomes worthwhile -----------------------------------------
		if (walking) {
			stalls = 0;
			for (;;) {
				unsigned leaf_bits = 0;
				int leaf_refs[4] = {0, 0, 0, 0};
				const bool was_walking = cur != REF_END;
				if (was_walking) {
					if (wm->dead[slot]) {
						cur = REF_END;
					} else {
						lc.nodes++, box_nodes++;
						int next = REF_END;
						float next_d = __int_as_float(0x7f800000);
						if (WIDE) {
							// straight-line code for the four children: overlap, leaf / internal, distance; then the nearest internal child
							const float4 *rec = sc.wide + 8ull * (unsigned) cur;
							float4 a[4], b[4];
#pragma unroll
							for (int k = 0; k < 4; ++k) a[k] = __ldg(rec + 2 * k), b[k] = __ldg(rec + 2 * k + 1);
							float d[4];
#pragma unroll
							for (int k = 0; k < 4; ++k) {
								leaf_refs[k] = __float_as_int(a[k].w);
								const bool ov = (leaf_refs[k] != WIDE_EMPTY) & cull_overlap(cb, a[k].x, a[k].y, a[k].z, b[k].x, b[k].y, b[k].z);
								const bool leaf = ov & (leaf_refs[k] < 0);
								leaf_bits |= (leaf ? 1u : 0u) << k;
								const float dx = a[k].x - cb.cx, dy = a[k].y - cb.cy, dz = a[k].z - cb.cz;
								d[k] = (ov & !leaf) ? dx * dx + dy * dy + dz * dz : __int_as_float(0x7f800000);
							}
#pragma unroll
							for (int k = 0; k < 4; ++k) {
								const bool nearer = d[k] < next_d;
								next = nearer ? leaf_refs[k] : next, next_d = nearer ? d[k] : next_d;
							}
#pragma unroll
							for (int k = 0; k < 4; ++k)
								if (d[k] < __int_as_float(0x7f800000) && leaf_refs[k] != next) push(leaf_refs[k]);
						} else {
							const float4 *rec = sc.nodes + 4ull * (unsigned) cur;
							const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1), n2 = __ldg(rec + 2), n3 = __ldg(rec + 3);
							const bool ov_l = cull_overlap(cb, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y), ov_r = cull_overlap(cb, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
							const int lref = __float_as_int(n3.x), rref = __float_as_int(n3.y);
							if (ov_l) {
								if (lref < 0) leaf_bits |= 1u, leaf_refs[0] = lref;
								else next = lref;
							}
							if (ov_r) {
								if (rref < 0) leaf_bits |= 2u, leaf_refs[1] = rref;
								else if (next == REF_END) next = rref;
								else push(rref);
							}
						}
						if (next == REF_END && sp > 0) next = pop();
						cur = next;
					}
				}
				__syncwarp();
				{
					// append the overlapping leaf children to the FIFO: a warp-wide exclusive scan of the per-lane counts (0..4)
					// from three ballots, then every lane writes its own pairs
					const unsigned cnt = __popc(leaf_bits);
					const unsigned b0 = __ballot_sync(FULL, cnt & 1u), b1 = __ballot_sync(FULL, cnt & 2u), b2 = __ballot_sync(FULL, cnt & 4u);
					if (b0 | b1 | b2) {
						unsigned e = w.leaf_tail + __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
#pragma unroll
						for (int k = 0; k < (WIDE ? 4 : 2); ++k)
							if ((leaf_bits >> k) & 1u) {
								wm->leaf_ref[e & (LEAF_FIFO - 1)] = leaf_refs[k];
								wm->leaf_slot[e & (LEAF_FIFO - 1)] = (unsigned char) slot;
								++e;
							}
						w.leaf_tail += __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
					}
				}
				if (leaf_bits) wm->pairs[slot] += __popc(leaf_bits);  // only the walking lane adds; testers subtract in another phase
				// descent over (or cut short by a hit): hand the slot back, now or with its last pair
				const bool over = was_walking && cur == REF_END, pairs_left = over && wm->pairs[slot] != 0;
				if (pairs_left) wm->done[slot] = 1;
				w.release(over && !pairs_left, slot);
				const unsigned still = __ballot_sync(FULL, cur != REF_END);
				if (!still || w.leaf_tail - w.leaf_head >= tune.leaf_min) break;
				// enough lanes have run dry and there is (or can be) something to give them: go and refill
				if ((int) (32u - __popc(still)) >= tune.refill_min && (n_pend || !exhausted)) break;
			}
			continue;
		}
		// ---- nothing walking and nothing to hand out: test what is left in the leaf FIFO (that also frees ring slots) ----------------
		if (n_leaf) {
			leaf_batch(n_leaf);
			continue;
		}
		if (exhausted && !n_pend) break;
		if (++stalls > 64u) break;  // unreachable by construction (every slot is free here, so expansion can proceed); never spin
	}
	flush_counters(sc, lc);
	if (sc.counters) atomicMax(sc.counters + CNT_MAX_BOX_NODES, (unsigned long long) max(max_box_nodes, box_nodes));
}

}  // namespace mgodpl_b200
