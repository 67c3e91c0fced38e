// Device-resident data layout of the collision-query path (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mgodpl_b200.h"

namespace mgodpl_b200 {

// ---------------------------------------------------------------------------------------------------------------
// Scene BVH: binary tree over the triangles, one 64-byte record per internal node (four float4 = four 16-byte loads,
// one 64 B half-line), records in depth-first preorder so a left child sits right behind its parent.  A record holds
// BOTH children (boxes relative to SceneDev::origin, centre + half extent rounded outward, and references):
//   f4[0] = (L.cx, L.cy, L.cz, L.hx)   f4[1] = (L.hy, L.hz, R.cx, R.cy)   f4[2] = (R.cz, R.hx, R.hy, R.hz)
//   f4[3] = (as_float(L.ref), as_float(R.ref), -, -)
// ref >= 0 : record index of an internal child
// ref <  0 : leaf; info = -1 - ref, (info >> 4) = first triangle slot, (info & 15) + 1 = triangle count
// One visit therefore tests two boxes per dependent fetch, and leaves never cost a fetch of their own.
// Triangles are stored in leaf (preorder) order as 9 fp64 world coordinates, so a leaf's triangles are contiguous.
// ---------------------------------------------------------------------------------------------------------------
//
// Stage B walks a 4-wide form of the same tree: every second level is folded into its parent, so a 128-byte record
// (eight float4) holds up to four children, child k as
//   f4[2k] = (cx, cy, cz, as_float(ref))   f4[2k+1] = (hx, hy, hz, -)
// with the same reference encoding (ref >= 0: wide record index) and WIDE_EMPTY in unused slots.  Half as many dependent
// fetches per descent as the binary records, which remain for the one-thread overflow path.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WIDE_EMPTY = (int) 0x80000000;
struct SceneDev {
	const float4 *nodes;     // 4 * n_nodes float4
	const float4 *wide;      // 8 * n_wide float4, or nullptr (MGODPL_WIDE_BVH=0): stage B then walks the binary records
	uint32_t n_wide;
	float4 top[8];           // copy of wide record 0 (the root's up to four grandchildren): stage A culls against it before queueing
	const double *tris;      // 9 * n_tris doubles (v0 v1 v2), world frame, leaf order: the exact test reads these
	const float4 *tris32;    // 3 * n_tris float4 (v0 v1 v2 relative to origin, w unused): the fp32 pre-filter reads these
	uint32_t n_nodes;
	uint32_t n_tris;
	double origin[3];        // fp32 node coordinates are relative to this point (centre of the scene bounds)
	float root_c[3], root_h[3]; // scene bounds relative to origin (centre, half extent)
	float margin;            // conservative inflation applied to every fp32 cull test
	unsigned long long *counters; // mgodpl_counters as 6 x u64, or nullptr
	// Clearance grid (scene_build.cu, k_clearance_grid): a regular grid over the scene bounds grown by `clear_pad`; cell value v
	// (u8) says "no triangle lies within v * clear_q metres of ANY point of this cell" (distance from the cell centre to the
	// nearest triangle, minus the cell's half diagonal, rounded down).  One byte load answers "is this sphere free of the scene"
	// for most candidates long before FK or a BVH descent; it can only say "certainly free", never "hit".
	const unsigned char *clear;  // clear_n[0] * clear_n[1] * clear_n[2] cells, x fastest; nullptr = no grid
	int clear_n[3];
	float clear_lo[3];       // grid corner relative to origin
	float clear_inv;         // 1 / cell size
	float clear_q;           // metres per quantum
	float clear_pad;         // what a point outside the grid is known to clear (it is that far from the scene bounds)
};

/// Lower bound on the distance from origin-relative point (x, y, z) to the nearest triangle; 0 when nothing is known.
/// NaN coordinates land in the "outside the grid" branch.
__device__ __forceinline__ float scene_clearance(const SceneDev &sc, float x, float y, float z) {
	if (!sc.clear) return 0.f;
	const float gx = (x - sc.clear_lo[0]) * sc.clear_inv, gy = (y - sc.clear_lo[1]) * sc.clear_inv, gz = (z - sc.clear_lo[2]) * sc.clear_inv;
	// float -> int conversion saturates and maps NaN to 0; the unsigned compares reject negative cells and cells beyond the grid
	// (a NaN coordinate is caught by the explicit test: it must not alias cell 0)
	const unsigned ix = (unsigned) __float2int_rd(gx), iy = (unsigned) __float2int_rd(gy), iz = (unsigned) __float2int_rd(gz);
	if (ix >= (unsigned) sc.clear_n[0] || iy >= (unsigned) sc.clear_n[1] || iz >= (unsigned) sc.clear_n[2] || !(gx + gy + gz == gx + gy + gz))
		return sc.clear_pad;
	return (float) __ldg(sc.clear + ((iz * (unsigned) sc.clear_n[1] + iy) * (unsigned) sc.clear_n[0] + ix)) * sc.clear_q;
}

// ---------------------------------------------------------------------------------------------------------------
// Robot: the reference's explicit-stack DFS (RobotModel.cpp:17-90) only depends on the model, never on the state,
// so the host unrolls it once into a list of ops executed in pop order:
//     tf[child] = tf[parent] . pre . J(theta[var]) . post
// with pre = this-side attachment, post = inverse(other-side attachment), J = axis-angle rotation (inverse when the
// joint is crossed from linkB) or identity for fixed joints.
// ---------------------------------------------------------------------------------------------------------------
struct FkOp {
	int32_t parent, child;
	int32_t var;          // index into joint values, -1 = no variable (fixed joint)
	int32_t flags;        // FK_*
	int32_t box_begin, box_end;  // RobotDev::boxes[box_begin, box_end) hang off `child`
	double pre_t[3], pre_q[4];
	double post_t[3], post_q[4];
	double axis[3];
};
enum : int32_t {
	FK_INVERT = 1,        // joint crossed from linkB: use the inverse rotation
	FK_PRE_ROT_ID = 2,    // pre.q is exactly identity
	FK_POST_ROT_ID = 4,   // post.q is exactly identity
	FK_FOR_COLLISION = 8, // child (or something below it) carries collision geometry
	FK_PARENT_IS_PREV = 16, // among the collision ops, the parent is the previous op's child (or the root for the first):
	                        // the running pose stays in registers, no link table access
	FK_STORE = 32         // a later, non-adjacent collision op uses this child as its parent: keep it in the link table
};

enum ShapeKind : int { SHAPE_KIND_BOX = 0, SHAPE_KIND_CAPSULE = 1, SHAPE_KIND_CONVEX = 2 };
constexpr int BOX_SPHERES = 4;
struct BoxDev {
	int32_t link;
	int32_t tf_identity;  // shape transform is exactly identity
	double half[3];
	double t[3], q[4];    // shape transform relative to the link
	// BOX_SPHERES equal spheres (centres in the box frame, one radius) that together cover the box: the box is free of the
	// scene when every sphere is (clearance grid, SceneDev::clear)
	float sphere_c[BOX_SPHERES][3];
	float sphere_r;
	// ShapeKind (geometry.cuh).  A capsule or convex polytope travels as its bounding box (half, t, q: the tree descent and every
	// fp32 cull are a box link's) and differs in the exact leaf test only: capsule radius = half[0], segment half length =
	// half[2] - half[0]; a convex polytope's hull lives at RobotDev::shape_data + data_off (doubles).
	// (16-bit fields in the struct's old padding word: its size — and with it the kernels' register allocation — stays what it was)
	int16_t kind;
	int16_t data_off;
};

struct RobotDev {
	int32_t n_links, n_ops, n_boxes, n_vars;
	int32_t root, ee;
	int32_t root_box_begin, root_box_end;  // boxes of the root link (boxes are sorted by the op that reaches their link)
	int32_t chain;        // every collision op continues from the previous one: FK needs no link table
	int32_t all_boxes;    // every collision shape is a box (the reference's robots): the fp32 pre-filter may answer "certain hit"
	const double *shape_data;  // device copy of the convex hulls (per device, patched in by the C ABI), or nullptr
	int32_t n_ee_path;    // ops on the way from the root to the end effector (ee_path), in execution order
	double reach;         // radius around the root link origin that contains every box for every joint value
	int8_t ee_path[MGODPL_MAX_LINKS];
	double limits[2 * MGODPL_MAX_JOINT_VALUES];  // (min, max) per joint variable: device-side goal sampling draws inside them
	FkOp ops[MGODPL_MAX_LINKS];
	BoxDev boxes[MGODPL_MAX_BOXES];
};

struct Tf64 {
	double tx, ty, tz, qx, qy, qz, qw;
};

// Counter slots (mgodpl_counters order)
enum { CNT_STATES = 0, CNT_BOXES, CNT_NODES, CNT_LEAF, CNT_MOTIONS, CNT_UNCERTAIN, CNT_MAX_BOX_NODES, CNT_N };

}  // namespace mgodpl_b200
