// C ABI of the collision-query path (include/mgodpl_b200.h): handle management, robot flattening, host-buffer
// staging, and the host-side fix-up that keeps motion step counts bit-identical to the reference's libm.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"

using namespace mgodpl_b200;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
	g_last_error = msg;
	return code;
}
int fail_cuda(cudaError_t e, const char *what) {
	// sticky errors must not poison the next call's cudaGetLastError
	cudaGetLastError();
	return fail(e == cudaErrorMemoryAllocation ? MGODPL_E_OUT_OF_MEMORY : MGODPL_E_CUDA,
				std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(x)                                    \
	do {                                         \
		cudaError_t e_ = (x);                    \
		if (e_ != cudaSuccess) return fail_cuda(e_, #x); \
	} while (0)

int require_device() {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return fail(MGODPL_E_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
	}
	return MGODPL_OK;
}

/// Scene handles are bound to the device they were built on: every call on a scene runs there, whatever device is current
/// in the calling thread, and leaves the caller's current device as it found it.
struct DeviceGuard {
	int prev = -1;
	bool changed = false;
	explicit DeviceGuard(int want) {
		if (cudaGetDevice(&prev) == cudaSuccess && prev != want) changed = cudaSetDevice(want) == cudaSuccess;
	}
	~DeviceGuard() {
		if (changed) cudaSetDevice(prev);
	}
	DeviceGuard(const DeviceGuard &) = delete;
	DeviceGuard &operator=(const DeviceGuard &) = delete;
};

static constexpr int E2E_PRIO_DEFAULT = 0;
/// A growable device + pinned-host staging area, one per in-flight host-buffer call.
struct Workspace {
	char *dev = nullptr, *host = nullptr;
	size_t dev_cap = 0, host_cap = 0;
	// Streams of one host-buffer call (created on first use): every chunk's input copy goes through `copy_in`, back to back, so
	// the PCIe link is busy from the first microsecond to the last input byte; chunk k's kernels and result copies run on
	// side[k], which waits for "chunk k has arrived" (ev_in[k]) only.
	static constexpr size_t MAX_CHUNKS = 8;
	cudaStream_t copy_in = nullptr, side[MAX_CHUNKS] = {};
	cudaEvent_t ev_start = nullptr, ev_in[MAX_CHUNKS] = {};
	cudaEvent_t ev_chunk[MAX_CHUNKS] = {};  // "results of chunk k are in pinned memory"
	// A chunk's kernels + result copy as an instantiated CUDA graph, re-used while the call's shape stays what it was (a planner
	// validates batch after batch of one size): one graph launch instead of ~12 enqueues per chunk, and the pipeline records its
	// two rounds overlapped (launch_check_motions), which only pays off inside a graph.
	struct ChunkGraph {
		cudaGraphExec_t exec = nullptr;
		unsigned long long key[10] = {};
	};
	ChunkGraph chunk_graph[MAX_CHUNKS];
	MotionOverlap overlap[MAX_CHUNKS];
	bool graphs_failed = false;  // a capture went wrong once: plain launches from then on
	void drop_graphs() {
		for (auto &g: chunk_graph) {
			if (g.exec) cudaGraphExecDestroy(g.exec);
			g.exec = nullptr;
		}
	}

	cudaError_t side_streams() {
		if (copy_in) return cudaSuccess;
		cudaError_t e = cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming);
		// MGODPL_E2E_PRIO (experiment): 1 = earlier chunks at higher stream priority (they retire first and leave the SMs to the
		// chunks behind them), 2 = later chunks at higher priority (the chunk the call ends on is served first); 0 = all equal.
		const char *pe = getenv("MGODPL_E2E_PRIO");
		const int prio_mode = pe ? atoi(pe) : E2E_PRIO_DEFAULT;
		int least = 0, greatest = 0;
		if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&least, &greatest);  // numerically: greatest <= least
		for (size_t i = 0; i < MAX_CHUNKS && e == cudaSuccess; ++i) {
			const int levels = least - greatest, rank = (int) std::min<size_t>(i, (size_t) std::max(levels, 0));
			const int prio = prio_mode == 1 ? std::min(least, greatest + rank) : prio_mode == 2 ? std::max(greatest, least - rank) : least;
			e = cudaStreamCreateWithPriority(&side[i], cudaStreamNonBlocking, prio);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_chunk[i], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&overlap[i].side, cudaStreamNonBlocking, prio);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&overlap[i].ev_head, cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&overlap[i].ev_tail, cudaEventDisableTiming);
		}
		return e;
	}
	cudaError_t reserve(size_t dev_bytes, size_t host_bytes) {
		if (dev_bytes > dev_cap || host_bytes > host_cap) drop_graphs();  // they point into the buffers that are about to move
		if (dev_bytes > dev_cap) {
			if (dev) cudaFree(dev);
			dev = nullptr, dev_cap = 0;
			size_t cap = dev_bytes + dev_bytes / 4 + 4096;
			cudaError_t e = cudaMalloc((void **) &dev, cap);
			if (e != cudaSuccess) return e;
			dev_cap = cap;
		}
		if (host_bytes > host_cap) {
			if (host) cudaFreeHost(host);
			host = nullptr, host_cap = 0;
			size_t cap = host_bytes + host_bytes / 4 + 4096;
			cudaError_t e = cudaMallocHost((void **) &host, cap);
			if (e != cudaSuccess) return e;
			host_cap = cap;
		}
		return cudaSuccess;
	}
	~Workspace() {
		drop_graphs();
		for (auto &o: overlap) {
			if (o.side) cudaStreamDestroy(o.side);
			if (o.ev_head) cudaEventDestroy(o.ev_head);
			if (o.ev_tail) cudaEventDestroy(o.ev_tail);
		}
		if (dev) cudaFree(dev);
		if (host) cudaFreeHost(host);
		if (copy_in) cudaStreamDestroy(copy_in);
		if (ev_start) cudaEventDestroy(ev_start);
		for (size_t i = 0; i < MAX_CHUNKS; ++i) {
			if (side[i]) cudaStreamDestroy(side[i]);
			if (ev_in[i]) cudaEventDestroy(ev_in[i]);
			if (ev_chunk[i]) cudaEventDestroy(ev_chunk[i]);
		}
	}
};

struct WorkspacePool {
	std::mutex mu;
	std::vector<std::unique_ptr<Workspace>> free_list;
	std::unique_ptr<Workspace> acquire() {
		std::lock_guard<std::mutex> g(mu);
		if (free_list.empty()) return std::make_unique<Workspace>();
		auto w = std::move(free_list.back());
		free_list.pop_back();
		return w;
	}
	void release(std::unique_ptr<Workspace> w) {
		// keep a few workspaces of ordinary size around; a very large one (or a fifth idle one) goes back to the driver
		// instead of pinning device memory for the life of the handle
		std::unique_lock<std::mutex> g(mu);
		if (free_list.size() < 4 && w->dev_cap <= ((size_t) 4 << 30)) {
			free_list.push_back(std::move(w));
			return;
		}
		g.unlock();
		w.reset();
	}
};
struct Lease {
	WorkspacePool &pool;
	std::unique_ptr<Workspace> ws;
	bool settled = false;  // set once the call has synchronised with everything it enqueued on this workspace
	explicit Lease(WorkspacePool &p) : pool(p), ws(p.acquire()) {}
	~Lease() {
		// an error return may leave copies or kernels in flight on the workspace: wait before another call can reuse it
		if (!settled) cudaDeviceSynchronize(), cudaGetLastError();
		pool.release(std::move(ws));
	}
	Workspace *operator->() { return ws.get(); }
};

/// Carves aligned sub-buffers out of a workspace.
struct Carver {
	size_t off = 0;
	size_t take(size_t bytes) {
		size_t o = off;
		off += (bytes + 255) & ~(size_t) 255;
		return o;
	}
};

/// Scratch of the calls that take no scene handle (forward kinematics, k-NN, roadmap shortest paths, mesh components):
/// pooled like a scene's, so a planner calling them in a loop pays no cudaMalloc / cudaFree (each an implicit device sync).
/// Device-pointer calls without a scene handle only enqueue work, so their scratch must outlive the call: one workspace per
/// stream, used in stream order, its mutex held while a call enqueues (same scheme as mgodpl_scene::stream_slot).
struct GlobalStreamSlot {
	std::mutex mu;
	Workspace ws;
};
GlobalStreamSlot &global_stream_slot(cudaStream_t stream) {
	static std::mutex mu;
	static auto *slots = new std::map<cudaStream_t, std::unique_ptr<GlobalStreamSlot>>;  // never destroyed (see global_pool)
	std::lock_guard<std::mutex> g(mu);
	auto &s = (*slots)[stream];
	if (!s) s = std::make_unique<GlobalStreamSlot>();
	return *s;
}

WorkspacePool &global_pool() {
	static WorkspacePool *pool = new WorkspacePool;  // never destroyed: static destructors run after the CUDA runtime has shut down
	return *pool;
}

}  // namespace

static std::atomic<unsigned long long> g_handle_uid{1};

struct mgodpl_scene {
	const unsigned long long uid = g_handle_uid++;  // (cached graphs are keyed by it: a pointer can be re-used after destroy)
	int device = 0;  // the CUDA device that holds the BVH
	SceneDev dev{};
	BuiltScene built{};
	mgodpl_scene_info info{};
	unsigned long long *d_counters = nullptr;
	WorkspacePool pool;  // host-buffer calls: one exclusive workspace per call, held until the call has synchronised
	// Device-buffer calls only enqueue work, so their scratch must outlive the call: one workspace per stream, reused in
	// stream order (growing it frees the old block with cudaFree, which waits for the device first).
	// Two host threads may enqueue on the same stream (both passing NULL, say): the slot's mutex is held from the scratch
	// look-up to the last launch of a call, so calls on one stream never interleave their launches or resize under each other.
	struct StreamSlot {
		std::mutex mu;
		Workspace ws;
		MotionOverlap overlap;  // created on first use (motion batches): see launch_check_motions
		cudaError_t overlap_streams() {
			if (overlap.side) return cudaSuccess;
			cudaError_t e = cudaStreamCreateWithFlags(&overlap.side, cudaStreamNonBlocking);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&overlap.ev_head, cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&overlap.ev_tail, cudaEventDisableTiming);
			return e;
		}
		~StreamSlot() {
			if (overlap.side) cudaStreamDestroy(overlap.side);
			if (overlap.ev_head) cudaEventDestroy(overlap.ev_head);
			if (overlap.ev_tail) cudaEventDestroy(overlap.ev_tail);
		}
	};
	std::mutex stream_mu;
	std::map<cudaStream_t, std::unique_ptr<StreamSlot>> stream_ws;
	StreamSlot &stream_slot(cudaStream_t stream) {
		std::lock_guard<std::mutex> g(stream_mu);
		auto &s = stream_ws[stream];
		if (!s) s = std::make_unique<StreamSlot>();
		return *s;
	}
};

struct mgodpl_robot {
	const unsigned long long uid = g_handle_uid++;
	RobotDev dev{};
	int n_joints = 0;
	// Convex links only: the hulls (vertices, face planes, edge directions; convex_triangle_hit's layout) and their device copies,
	// uploaded on first use per device.  Robots made of boxes and capsules hold no device memory at all.
	std::vector<double> shape_data;
	mutable std::mutex mu;
	mutable std::map<int, double *> shape_data_on;
	~mgodpl_robot() {
		int cur = 0;
		cudaGetDevice(&cur);
		for (auto &kv: shape_data_on) {
			cudaSetDevice(kv.first);
			cudaFree(kv.second);
		}
		if (!shape_data_on.empty()) cudaSetDevice(cur);
	}
};

namespace {
/// The robot as the kernels of `device` (the current device) take it: robot->dev itself, or — for robots with convex links — a
/// copy that points at this device's upload of the hull data.
cudaError_t robot_for_device(const mgodpl_robot *robot, int device, RobotDev &patched, const RobotDev **out) {
	*out = &robot->dev;
	if (robot->shape_data.empty()) return cudaSuccess;
	std::lock_guard<std::mutex> g(robot->mu);
	double *&p = robot->shape_data_on[device];
	if (!p) {
		cudaError_t e = cudaMalloc((void **) &p, robot->shape_data.size() * sizeof(double));
		if (e == cudaSuccess) e = cudaMemcpy(p, robot->shape_data.data(), robot->shape_data.size() * sizeof(double), cudaMemcpyHostToDevice);
		if (e != cudaSuccess) {
			cudaFree(p);
			p = nullptr;
			return e;
		}
	}
	patched = robot->dev;
	patched.shape_data = p;
	*out = &patched;
	return cudaSuccess;
}
}  // namespace
/// `const RobotDev &RB` for the query entry points (after DeviceGuard has switched to the scene's device).
#define ROBOT_ON_SCENE_DEVICE(RB)      \
	RobotDev patched_robot_;           \
	const RobotDev *robot_ptr_ = nullptr; \
	CU(robot_for_device(robot, scene->device, patched_robot_, &robot_ptr_)); \
	const RobotDev &RB = *robot_ptr_

// =================================================================================================================
extern "C" {

const char *mgodpl_last_error(void) { return g_last_error.c_str(); }

int mgodpl_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int mgodpl_set_device(int device) {
	if (int rc = require_device()) return rc;
	CU(cudaSetDevice(device));
	return MGODPL_OK;
}
int mgodpl_scene_device(const mgodpl_scene *scene, int *device) {
	if (!scene || !device) return fail(MGODPL_E_INVALID_ARGUMENT, "null argument");
	*device = scene->device;
	return MGODPL_OK;
}

// ---- scene ---------------------------------------------------------------------------------------------------------
static int scene_create_impl(const double *verts, size_t nv, const uint32_t *tris32, const uint64_t *tris64, size_t nt,
							 const mgodpl_scene_opts *opts, mgodpl_scene **out) {
	if (!out) return fail(MGODPL_E_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if ((nv && !verts) || (nt && !tris32 && !tris64)) return fail(MGODPL_E_INVALID_ARGUMENT, "null mesh pointer");
	if (nt > (1u << 27)) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^27 triangles");
	// Validate indices and gather the bounds of the vertices the triangles actually use (host, one pass).
	std::vector<uint32_t> idx(nt * 3);
	double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
	for (size_t i = 0; i < nt * 3; ++i) {
		uint64_t v = tris32 ? tris32[i] : tris64[i];
		if (v >= nv) return fail(MGODPL_E_INVALID_ARGUMENT, "triangle index out of range");
		idx[i] = (uint32_t) v;
		for (int a = 0; a < 3; ++a) {
			double c = verts[3 * v + a];
			if (!std::isfinite(c)) return fail(MGODPL_E_INVALID_ARGUMENT, "non-finite vertex coordinate");
			lo[a] = std::min(lo[a], c), hi[a] = std::max(hi[a], c);
		}
	}
	if (!nt) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.0;
	if (int rc = require_device()) return rc;
	auto sc = std::make_unique<mgodpl_scene>();
	cudaGetDevice(&sc->device);
	double origin[3], max_abs = 0.0;
	for (int a = 0; a < 3; ++a) {
		origin[a] = 0.5 * (lo[a] + hi[a]);
		max_abs = std::max(max_abs, 0.5 * (hi[a] - lo[a]));
	}
	cudaStream_t stream = nullptr;
	double *d_verts = nullptr;
	uint32_t *d_tris = nullptr;
	unsigned leaf = opts && opts->max_leaf_size ? opts->max_leaf_size : 2;
	bool refine = !(opts && (opts->flags & MGODPL_SCENE_NO_REFINE));
	if (nt) {
		CU(cudaMalloc((void **) &d_verts, nv * 3 * sizeof(double)));
		cudaError_t e = cudaMalloc((void **) &d_tris, nt * 3 * sizeof(uint32_t));
		if (e != cudaSuccess) {
			cudaFree(d_verts);
			return fail_cuda(e, "cudaMalloc(tris)");
		}
		e = cudaMemcpyAsync(d_verts, verts, nv * 3 * sizeof(double), cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_tris, idx.data(), nt * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
		static const bool want_wide = !(getenv("MGODPL_WIDE_BVH") && atoi(getenv("MGODPL_WIDE_BVH")) == 0);
		static const bool env_grid = !(getenv("MGODPL_CLEARANCE_GRID") && atoi(getenv("MGODPL_CLEARANCE_GRID")) == 0);
		const bool want_grid = env_grid && !(opts && (opts->flags & MGODPL_SCENE_NO_CLEARANCE_GRID));
		if (e == cudaSuccess) e = build_scene(d_verts, nv, d_tris, nt, lo, hi, origin, leaf, refine, want_wide, want_grid, &sc->built, stream);
		cudaStreamSynchronize(stream);
		cudaFree(d_verts);
		cudaFree(d_tris);
		if (e != cudaSuccess) {
			cudaFree(sc->built.nodes), cudaFree(sc->built.wide), cudaFree(sc->built.tris), cudaFree(sc->built.tris32), cudaFree(sc->built.tri_ids);
			cudaFree(sc->built.clear);
			return fail_cuda(e, "build_scene");
		}
	}
	SceneDev &d = sc->dev;
	d.nodes = sc->built.nodes;
	d.tris = sc->built.tris;
	d.tris32 = sc->built.tris32;
	d.n_nodes = sc->built.n_nodes;
	d.n_tris = sc->built.n_tris;
	for (int a = 0; a < 3; ++a) d.origin[a] = origin[a];
	// fp32 cull tests carry ~1e-7 relative error on coordinates up to max_abs (plus robot reach): inflate every
	// box by a margin two orders of magnitude above that.  Exactness is restored by the fp64 leaf test.
	d.margin = (float) std::max(2e-5, 4e-6 * (max_abs + 2.0));
	for (int a = 0; a < 3; ++a) d.root_c[a] = sc->built.root_c[a], d.root_h[a] = sc->built.root_h[a];  // empty scene: half = -1
	if (sc->built.max_depth > 48) {
		mgodpl_scene_destroy(sc.release());
		return fail(MGODPL_E_INVALID_ARGUMENT, "BVH deeper than the traversal stack (pathological triangle distribution)");
	}
	if (sc->built.wide) {
		d.wide = sc->built.wide, d.n_wide = sc->built.n_wide;
		std::memcpy(d.top, sc->built.top, sizeof(d.top));
	}
	if (sc->built.clear) {
		d.clear = sc->built.clear;
		for (int a = 0; a < 3; ++a) d.clear_n[a] = sc->built.clear_n[a], d.clear_lo[a] = sc->built.clear_lo[a];
		d.clear_inv = 1.0f / sc->built.clear_cell, d.clear_q = sc->built.clear_quantum;
		d.clear_pad = sc->built.clear_max - 1e-3f;  // outside the grid = at least the pad away from the bounds, minus fp32 slack
	}
	if (opts && (opts->flags & MGODPL_SCENE_COUNTERS)) {
		cudaError_t e = cudaMalloc((void **) &sc->d_counters, CNT_N * sizeof(unsigned long long));
		if (e == cudaSuccess) e = cudaMemset(sc->d_counters, 0, CNT_N * sizeof(unsigned long long));
		if (e != cudaSuccess) {
			mgodpl_scene_destroy(sc.release());
			return fail_cuda(e, "scene counters");
		}
		d.counters = sc->d_counters;
	}
	mgodpl_scene_info &in = sc->info;
	in.n_triangles = nt, in.n_nodes = d.n_nodes, in.n_leaves = sc->built.n_leaves;
	in.device_bytes = sc->built.device_bytes;
	in.n_wide_nodes = d.n_wide;
	in.traversal_bytes = (d.wide ? (uint64_t) d.n_wide * 128 : (uint64_t) d.n_nodes * 64) + (uint64_t) nt * (72 + 48);
	for (int a = 0; a < 3; ++a) in.bounds_lo[a] = lo[a], in.bounds_hi[a] = hi[a];
	in.build_ms = sc->built.build_ms, in.sah_cost = sc->built.sah_cost;
	in.max_depth = sc->built.max_depth, in.max_leaf_size = sc->built.max_leaf;
	in.clearance_cells = (uint64_t) sc->built.clear_n[0] * sc->built.clear_n[1] * sc->built.clear_n[2];
	in.clearance_cell_size = sc->built.clear_cell, in.clearance_build_ms = sc->built.clear_build_ms;
	*out = sc.release();
	return MGODPL_OK;
}

int mgodpl_scene_create(const double *verts, size_t nv, const uint32_t *tris, size_t nt, const mgodpl_scene_opts *opts,
						mgodpl_scene **out) {
	return scene_create_impl(verts, nv, tris, nullptr, nt, opts, out);
}
int mgodpl_scene_create_u64(const double *verts, size_t nv, const uint64_t *tris, size_t nt, const mgodpl_scene_opts *opts,
							mgodpl_scene **out) {
	return scene_create_impl(verts, nv, nullptr, tris, nt, opts, out);
}
int mgodpl_scene_destroy(mgodpl_scene *s) {
	if (!s) return MGODPL_OK;
	DeviceGuard on_device(s->device);
	cudaFree(s->built.wide), cudaFree(s->built.clear);
	cudaFree(s->built.nodes), cudaFree(s->built.tris), cudaFree(s->built.tris32), cudaFree(s->built.tri_ids), cudaFree(s->d_counters);
	delete s;
	return MGODPL_OK;
}
int mgodpl_scene_get_info(const mgodpl_scene *s, mgodpl_scene_info *info) {
	if (!s || !info) return fail(MGODPL_E_INVALID_ARGUMENT, "null argument");
	*info = s->info;
	return MGODPL_OK;
}
int mgodpl_scene_get_counters(mgodpl_scene *s, mgodpl_counters *out, int reset) {
	if (!s || !out) return fail(MGODPL_E_INVALID_ARGUMENT, "null argument");
	memset(out, 0, sizeof(*out));
	if (!s->d_counters) return fail(MGODPL_E_INVALID_ARGUMENT, "scene was created without MGODPL_SCENE_COUNTERS");
	DeviceGuard on_device(s->device);
	unsigned long long v[CNT_N];
	CU(cudaMemcpy(v, s->d_counters, sizeof(v), cudaMemcpyDeviceToHost));
	out->states_checked = v[CNT_STATES], out->boxes_traversed = v[CNT_BOXES], out->nodes_visited = v[CNT_NODES];
	out->leaf_tests = v[CNT_LEAF], out->motions_checked = v[CNT_MOTIONS], out->uncertain_steps = v[CNT_UNCERTAIN];
	out->max_nodes_per_box = v[CNT_MAX_BOX_NODES];
	if (reset) CU(cudaMemset(s->d_counters, 0, sizeof(v)));
	return MGODPL_OK;
}

// ---- robot -----------------------------------------------------------------------------------------------------------
namespace {
struct HTf {
	double t[3], q[4];
};
HTf load_tf(const double *p) { return {{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}}; }
void hq_mul(const double *a, const double *b, double *o) {
	o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
	o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
	o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
	o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
/// Transformd::inverse (Transform.h:52-58): conjugate, then rotate the negated translation by it.
HTf inverse(const HTf &a) {
	HTf r;
	r.q[0] = -a.q[0], r.q[1] = -a.q[1], r.q[2] = -a.q[2], r.q[3] = a.q[3];
	double v[4] = {-a.t[0], -a.t[1], -a.t[2], 0.0}, c[4] = {a.q[0], a.q[1], a.q[2], a.q[3]}, tmp[4], res[4];
	hq_mul(r.q, v, tmp);
	hq_mul(tmp, c, res);
	r.t[0] = res[0], r.t[1] = res[1], r.t[2] = res[2];
	return r;
}
bool rot_is_identity(const double *q) { return q[0] == 0.0 && q[1] == 0.0 && q[2] == 0.0 && q[3] == 1.0; }
double vnorm(const double *t) { return std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]); }

/// Rotate v by the unit quaternion q (x, y, z, w).
void hq_rot(const double *q, const double *v, double *o) {
	const double tx = 2 * (q[1] * v[2] - q[2] * v[1]), ty = 2 * (q[2] * v[0] - q[0] * v[2]), tz = 2 * (q[0] * v[1] - q[1] * v[0]);
	o[0] = v[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
	o[1] = v[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
	o[2] = v[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
}

/// Hull of a small point set by exhaustive search (n <= MGODPL_MAX_CONVEX_VERTICES): every vertex triple whose plane has all
/// points on one side is a face; every vertex pair that lies on two different faces is an edge.  Appends convex_triangle_hit's
/// record to `out` (vertices relative to `centre`, the middle of their bounding box) and returns false for point sets without
/// volume (a flat or collinear "hull" has no face normals to separate with).
bool build_hull(const double *v, int n, const double *centre, std::vector<double> &out) {
	std::vector<std::array<double, 3>> p(n);
	double scale = 0.0;
	for (int i = 0; i < n; ++i)
		for (int a = 0; a < 3; ++a) p[i][a] = v[3 * i + a] - centre[a], scale = std::max(scale, std::fabs(p[i][a]));
	if (!(scale > 0.0)) return false;
	const double eps = 1e-10 * scale;
	struct Face {
		double n[3], d;
		std::vector<int> on;
	};
	std::vector<Face> faces;
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j)
			for (int k = j + 1; k < n; ++k) {
				const double e0[3] = {p[j][0] - p[i][0], p[j][1] - p[i][1], p[j][2] - p[i][2]}, e1[3] = {p[k][0] - p[i][0], p[k][1] - p[i][1], p[k][2] - p[i][2]};
				double nn[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
				const double len = vnorm(nn);
				if (!(len > 1e-12 * scale * scale)) continue;
				for (double &x: nn) x /= len;
				double lo = 0.0, hi = 0.0;
				for (int q = 0; q < n; ++q) {
					const double s = nn[0] * (p[q][0] - p[i][0]) + nn[1] * (p[q][1] - p[i][1]) + nn[2] * (p[q][2] - p[i][2]);
					lo = std::min(lo, s), hi = std::max(hi, s);
				}
				const bool above = hi > eps, below = lo < -eps;
				if (above && below) continue;      // points on both sides: not a supporting plane
				if (!above && !below) return false;  // every point lies in this plane: no volume
				if (above)
					for (double &x: nn) x = -x;  // everything lies on the positive side: the outward normal points the other way
				bool dup = false;
				for (auto &f: faces) dup |= f.n[0] * nn[0] + f.n[1] * nn[1] + f.n[2] * nn[2] > 1.0 - 1e-12;
				if (dup) continue;
				Face f;
				memcpy(f.n, nn, sizeof(nn));
				f.d = -1e300;
				for (int q = 0; q < n; ++q) f.d = std::max(f.d, nn[0] * p[q][0] + nn[1] * p[q][1] + nn[2] * p[q][2]);
				for (int q = 0; q < n; ++q)
					if (nn[0] * p[q][0] + nn[1] * p[q][1] + nn[2] * p[q][2] > f.d - eps) f.on.push_back(q);
				faces.push_back(f);
			}
	if (faces.size() < 4) return false;
	std::vector<std::array<double, 3>> edges;
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j) {
			int shared = 0;
			for (auto &f: faces)
				shared += std::find(f.on.begin(), f.on.end(), i) != f.on.end() && std::find(f.on.begin(), f.on.end(), j) != f.on.end();
			if (shared < 2) continue;
			double d[3] = {p[j][0] - p[i][0], p[j][1] - p[i][1], p[j][2] - p[i][2]};
			const double len = vnorm(d);
			if (!(len > eps)) continue;
			for (double &x: d) x /= len;
			bool dup = false;
			for (auto &e: edges) dup |= std::fabs(e[0] * d[0] + e[1] * d[1] + e[2] * d[2]) > 1.0 - 1e-12;
			if (!dup) edges.push_back({d[0], d[1], d[2]});
		}
	out.push_back((double) n), out.push_back((double) faces.size()), out.push_back((double) edges.size()), out.push_back(0.0);
	for (auto &q: p) out.insert(out.end(), q.begin(), q.end());
	for (auto &f: faces) out.insert(out.end(), {f.n[0], f.n[1], f.n[2], f.d});
	for (auto &e: edges) out.insert(out.end(), e.begin(), e.end());
	return true;
}
}  // namespace

int mgodpl_robot_create(int32_t n_links, const mgodpl_shape_desc *shapes, int32_t n_shapes, const mgodpl_joint_desc *joints,
						int32_t n_joints, int32_t root_link, int32_t ee_link, mgodpl_robot **out) {
	return mgodpl_robot_create_ex(n_links, shapes, n_shapes, joints, n_joints, root_link, ee_link, nullptr, nullptr, out);
}

int mgodpl_robot_create_ex(int32_t n_links, const mgodpl_shape_desc *shapes, int32_t n_shapes, const mgodpl_joint_desc *joints,
						   int32_t n_joints, int32_t root_link, int32_t ee_link, const double *mesh_vertices,
						   const int32_t *mesh_vertex_offsets, mgodpl_robot **out) {
	if (!out) return fail(MGODPL_E_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if (n_links <= 0 || n_shapes < 0 || n_joints < 0 || (n_shapes && !shapes) || (n_joints && !joints))
		return fail(MGODPL_E_INVALID_ARGUMENT, "bad robot description");
	if (root_link < 0 || root_link >= n_links) return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: root link index out of range");
	if (ee_link >= n_links) return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: end effector index out of range");
	if (n_links > MGODPL_MAX_LINKS) return fail(MGODPL_E_UNSUPPORTED_ROBOT, "too many links");
	if (n_shapes > MGODPL_MAX_BOXES) return fail(MGODPL_E_UNSUPPORTED_ROBOT, "too many collision shapes");
	for (int i = 0; i < n_shapes; ++i) {
		const int st = shapes[i].shape_type;
		// the reference's own variant is Box | Mesh and it collides boxes only (collision_detection.cpp:21-22,49-51): a general
		// Mesh link keeps failing the way it does there; capsules and convex polytopes are this library's additions
		if (st != MGODPL_SHAPE_BOX && st != MGODPL_SHAPE_CAPSULE && st != MGODPL_SHAPE_CONVEX_MESH)
			return fail(MGODPL_E_UNSUPPORTED_SHAPE, "Only boxes are implemented for collision geometry.");
		if (st == MGODPL_SHAPE_CAPSULE && !(shapes[i].size[0] > 0.0 && shapes[i].size[2] >= 0.0))
			return fail(MGODPL_E_INVALID_ARGUMENT, "capsule needs size[0] = radius > 0 and size[2] = length >= 0");
		if (st == MGODPL_SHAPE_CONVEX_MESH) {
			if (!mesh_vertices || !mesh_vertex_offsets) return fail(MGODPL_E_INVALID_ARGUMENT, "convex mesh shapes need mgodpl_robot_create_ex with their vertices");
			const int nv = mesh_vertex_offsets[i + 1] - mesh_vertex_offsets[i];
			if (mesh_vertex_offsets[i] < 0 || nv < 4 || nv > MGODPL_MAX_CONVEX_VERTICES)
				return fail(MGODPL_E_UNSUPPORTED_SHAPE, "a convex mesh link needs 4 .. MGODPL_MAX_CONVEX_VERTICES vertices");
		}
		if (shapes[i].link < 0 || shapes[i].link >= n_links) return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: shape link index");
	}
	for (int i = 0; i < n_joints; ++i) {
		if (joints[i].joint_type != MGODPL_JOINT_REVOLUTE && joints[i].joint_type != MGODPL_JOINT_FIXED)
			return fail(MGODPL_E_UNKNOWN_JOINT_TYPE, "Unknown joint type");
		if (joints[i].link_a < 0 || joints[i].link_a >= n_links || joints[i].link_b < 0 || joints[i].link_b >= n_links)
			return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: joint link index");
	}
	if (int rc = require_device()) return rc;

	auto rb = std::make_unique<mgodpl_robot>();
	RobotDev &d = rb->dev;
	d.n_links = n_links, d.root = root_link, d.ee = ee_link, d.n_ops = 0, d.n_boxes = n_shapes;
	rb->n_joints = n_joints;

	// Back-links in insertJoint order (RobotModel.cpp:103-112).
	std::vector<std::vector<int>> link_joints(n_links);
	for (int j = 0; j < n_joints; ++j) {
		link_joints[joints[j].link_a].push_back(j);
		link_joints[joints[j].link_b].push_back(j);
	}
	// Replay forwardKinematics' explicit-stack DFS (RobotModel.cpp:17-90) symbolically.  An entry remembers which
	// joint produced it and the variable cursor at expansion time; a link takes the transform of its first pop.
	struct Entry {
		int link, parent, joint, var;
		bool from_a;
	};
	std::vector<Entry> stack{{root_link, -1, -1, -1, true}};
	std::vector<char> visited(n_links, 0);
	int var_cursor = 0;
	std::vector<double> depth_reach(n_links, 0.0);  // bound on |link origin - root origin|
	while (!stack.empty()) {
		Entry e = stack.back();
		stack.pop_back();
		if (visited[e.link]) continue;
		visited[e.link] = 1;
		if (e.joint >= 0) {
			const mgodpl_joint_desc &J = joints[e.joint];
			FkOp &op = d.ops[d.n_ops++];
			op.parent = e.parent, op.child = e.link, op.var = e.var, op.flags = 0;
			HTf pre = load_tf(e.from_a ? J.attachment_a : J.attachment_b);
			HTf post = inverse(load_tf(e.from_a ? J.attachment_b : J.attachment_a));
			if (!e.from_a) op.flags |= FK_INVERT;  // joint crossed from linkB: variable transform inverted (RobotModel.cpp:66-67)
			memcpy(op.pre_t, pre.t, sizeof(pre.t)), memcpy(op.pre_q, pre.q, sizeof(pre.q));
			memcpy(op.post_t, post.t, sizeof(post.t)), memcpy(op.post_q, post.q, sizeof(post.q));
			memcpy(op.axis, J.axis, sizeof(J.axis));
			if (rot_is_identity(pre.q)) op.flags |= FK_PRE_ROT_ID;
			if (rot_is_identity(post.q)) op.flags |= FK_POST_ROT_ID;
			depth_reach[e.link] = depth_reach[e.parent] + vnorm(pre.t) + vnorm(post.t);
		}
		for (int ji: link_joints[e.link]) {
			const mgodpl_joint_desc &J = joints[ji];
			if (visited[J.link_b]) continue;  // the reference only looks at linkB here (RobotModel.cpp:56-57)
			bool from_a = J.link_a == e.link;
			int var = J.joint_type == MGODPL_JOINT_REVOLUTE ? var_cursor : -1;
			var_cursor += J.joint_type == MGODPL_JOINT_REVOLUTE ? 1 : 0;
			stack.push_back({from_a ? J.link_b : J.link_a, e.link, ji, var, from_a});
		}
	}
	int n_vars = 0;
	for (int j = 0; j < n_joints; ++j) n_vars += joints[j].joint_type == MGODPL_JOINT_REVOLUTE;
	d.n_vars = n_vars;
	if (n_vars > MGODPL_MAX_JOINT_VALUES) return fail(MGODPL_E_UNSUPPORTED_ROBOT, "too many joint variables");
	for (int k = 0; k < d.n_ops; ++k)  // joint graphs with cycles can push a joint twice and run the variable cursor past the end
		if (d.ops[k].var >= n_vars) return fail(MGODPL_E_UNSUPPORTED_ROBOT, "joint graph is not a tree: a joint variable index exceeds the variable count");
	if (ee_link >= 0 && !visited[ee_link]) return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: end effector unreachable from root");

	// Boxes (sorted by the op that reaches their link, root first), the links FK must reach for collision, and the
	// reach radius of the whole robot.
	std::vector<int> link_rank(n_links, 0);  // order in which FK produces each link
	for (int k = 0; k < d.n_ops; ++k) link_rank[d.ops[k].child] = k + 1;
	// links the DFS never reaches keep the root transform (RobotModel.cpp:23-24): their boxes ride on the root link
	std::vector<int> box_link(n_shapes);
	for (int i = 0; i < n_shapes; ++i) box_link[i] = visited[shapes[i].link] ? shapes[i].link : root_link;
	std::vector<int> order(n_shapes);
	for (int i = 0; i < n_shapes; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return link_rank[box_link[x]] < link_rank[box_link[y]]; });
	std::vector<char> needed(n_links, 0);
	double reach = 0.0;
	for (int i = 0; i < n_shapes; ++i) {
		const mgodpl_shape_desc &sh = shapes[order[i]];
		BoxDev &b = d.boxes[i];
		b.link = box_link[order[i]];
		HTf tf = load_tf(sh.transform);
		b.kind = SHAPE_KIND_BOX, b.data_off = 0;
		if (sh.shape_type == MGODPL_SHAPE_CAPSULE) {
			// fcl::Capsule's convention: the cylinder part of length size[2] runs along the shape's z axis, centred
			b.kind = SHAPE_KIND_CAPSULE;
			b.half[0] = b.half[1] = sh.size[0], b.half[2] = 0.5 * sh.size[2] + sh.size[0];
		} else if (sh.shape_type == MGODPL_SHAPE_CONVEX_MESH) {
			// travels as the bounding box of its vertices in the shape frame; the shape frame moves to that box's centre
			const double *v = mesh_vertices + 3 * (size_t) mesh_vertex_offsets[order[i]];
			const int nv = mesh_vertex_offsets[order[i] + 1] - mesh_vertex_offsets[order[i]];
			double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
			for (int k = 0; k < nv; ++k)
				for (int a = 0; a < 3; ++a) {
					if (!std::isfinite(v[3 * k + a])) return fail(MGODPL_E_INVALID_ARGUMENT, "non-finite convex mesh vertex");
					lo[a] = std::min(lo[a], v[3 * k + a]), hi[a] = std::max(hi[a], v[3 * k + a]);
				}
			double centre[3], shift[3];
			for (int a = 0; a < 3; ++a) centre[a] = 0.5 * (lo[a] + hi[a]), b.half[a] = 0.5 * (hi[a] - lo[a]);
			b.kind = SHAPE_KIND_CONVEX, b.data_off = (int16_t) rb->shape_data.size();
			if (!build_hull(v, nv, centre, rb->shape_data)) return fail(MGODPL_E_UNSUPPORTED_SHAPE, "convex mesh link has no volume (flat or collinear vertices)");
			hq_rot(tf.q, centre, shift);
			for (int a = 0; a < 3; ++a) tf.t[a] += shift[a];
		} else {
			for (int a = 0; a < 3; ++a) b.half[a] = 0.5 * sh.size[a];
		}
		memcpy(b.t, tf.t, sizeof(tf.t)), memcpy(b.q, tf.q, sizeof(tf.q));
		b.tf_identity = rot_is_identity(tf.q) && tf.t[0] == 0.0 && tf.t[1] == 0.0 && tf.t[2] == 0.0;
		{
			// cover the box with BOX_SPHERES equal spheres: the axis split (n0 x n1 x n2 <= BOX_SPHERES sub-boxes) whose
			// circumscribed spheres are smallest
			int best_n[3] = {1, 1, 1};
			double best_r = vnorm(b.half);
			for (int n0 = 1; n0 <= BOX_SPHERES; ++n0)
				for (int n1 = 1; n0 * n1 <= BOX_SPHERES; ++n1)
					for (int n2 = 1; n0 * n1 * n2 <= BOX_SPHERES; ++n2) {
						const double sub[3] = {b.half[0] / n0, b.half[1] / n1, b.half[2] / n2};
						if (vnorm(sub) < best_r) best_r = vnorm(sub), best_n[0] = n0, best_n[1] = n1, best_n[2] = n2;
					}
			int k = 0;
			for (int i0 = 0; i0 < best_n[0]; ++i0)
				for (int i1 = 0; i1 < best_n[1]; ++i1)
					for (int i2 = 0; i2 < best_n[2]; ++i2, ++k) {
						const int idx[3] = {i0, i1, i2};
						for (int a = 0; a < 3; ++a) b.sphere_c[k][a] = (float) (-b.half[a] + (2 * idx[a] + 1) * b.half[a] / best_n[a]);
					}
			for (; k < BOX_SPHERES; ++k)
				for (int a = 0; a < 3; ++a) b.sphere_c[k][a] = b.sphere_c[k - 1][a];  // fewer sub-boxes than slots: repeat the last
			b.sphere_r = (float) (best_r * (1.0 + 1e-6) + 1e-6);
		}
		needed[b.link] = 1;
		reach = std::max(reach, depth_reach[b.link] + vnorm(tf.t) + vnorm(b.half));
	}
	d.reach = reach * (1.0 + 1e-9) + 1e-9;
	d.all_boxes = 1, d.shape_data = nullptr;
	for (int i = 0; i < n_shapes; ++i) d.all_boxes &= d.boxes[i].kind == SHAPE_KIND_BOX;
	for (int k = d.n_ops - 1; k >= 0; --k) {
		if (needed[d.ops[k].child]) {
			d.ops[k].flags |= FK_FOR_COLLISION;
			needed[d.ops[k].parent] = 1;
		}
	}
	// Contiguous box range of every link (boxes are sorted by link).
	auto range_of = [&](int link, int32_t *lo, int32_t *hi) {
		*lo = n_shapes, *hi = 0;
		for (int i = 0; i < n_shapes; ++i)
			if (d.boxes[i].link == link) *lo = std::min(*lo, i), *hi = std::max(*hi, i + 1);
		if (*hi == 0) *lo = 0;
	};
	range_of(root_link, &d.root_box_begin, &d.root_box_end);
	for (int k = 0; k < d.n_ops; ++k) range_of(d.ops[k].child, &d.ops[k].box_begin, &d.ops[k].box_end);
	// Register-resident FK: mark ops whose parent is the previous collision op's child, and children that must be stored.
	{
		int prev_child = root_link;
		std::vector<int> executed;
		for (int k = 0; k < d.n_ops; ++k)
			if (d.ops[k].flags & FK_FOR_COLLISION) executed.push_back(k);
		for (size_t e = 0; e < executed.size(); ++e) {
			FkOp &op = d.ops[executed[e]];
			if (op.parent == prev_child) op.flags |= FK_PARENT_IS_PREV;
			prev_child = op.child;
		}
		for (size_t e = 0; e < executed.size(); ++e)
			for (size_t f = e + 1; f < executed.size(); ++f)
				if (d.ops[executed[f]].parent == d.ops[executed[e]].child && !(d.ops[executed[f]].flags & FK_PARENT_IS_PREV))
					d.ops[executed[e]].flags |= FK_STORE;
	}
	{
		// chain robots (every procedural drone): each collision op's parent is the previous collision op's child
		d.chain = 1;
		for (int k = 0; k < d.n_ops; ++k)
			if ((d.ops[k].flags & FK_FOR_COLLISION) && !(d.ops[k].flags & FK_PARENT_IS_PREV)) d.chain = 0;
		// ops from the root to the end effector, for goal sampling (genGoalStateUniform needs the end-effector pose only)
		d.n_ee_path = 0;
		if (ee_link >= 0) {
			std::vector<int> path;
			for (int link = ee_link; link != root_link;) {
				int k = link_rank[link] - 1;  // the op that produced this link
				if (k < 0) break;
				path.push_back(k);
				link = d.ops[k].parent;
			}
			for (auto it = path.rbegin(); it != path.rend(); ++it) d.ee_path[d.n_ee_path++] = (int8_t) *it;
		}
	}
	// Joint limits for device-side goal sampling: one (min, max) per variable in variable order.  They travel inside RobotDev
	// (kernel parameters), so a robot handle owns no device memory and works with scenes on any device.
	{
		int v = 0;
		for (int j = 0; j < n_joints; ++j)
			if (joints[j].joint_type == MGODPL_JOINT_REVOLUTE) {
				d.limits[2 * v] = joints[j].min_angle, d.limits[2 * v + 1] = joints[j].max_angle;
				++v;
			}
	}
	*out = rb.release();
	return MGODPL_OK;
}

int mgodpl_robot_destroy(mgodpl_robot *r) {
	if (!r) return MGODPL_OK;
	delete r;
	return MGODPL_OK;
}
int mgodpl_robot_n_variables(const mgodpl_robot *r) { return r ? r->dev.n_vars : -1; }

// ---- argument checks shared by the query entry points ----------------------------------------------------------------
static int check_state_args(const mgodpl_robot *robot, size_t stride, int32_t js) {
	if (!robot) return fail(MGODPL_E_INVALID_ARGUMENT, "robot is null");
	if (js < robot->dev.n_vars) return fail(MGODPL_E_INVALID_ARGUMENT, "fewer joint values per state than joint variables");
	if (js > MGODPL_MAX_JOINT_VALUES) return fail(MGODPL_E_UNSUPPORTED_ROBOT, "too many joint values per state");
	if (stride < (size_t) (7 + js)) return fail(MGODPL_E_INVALID_ARGUMENT, "stride smaller than 7 + n_joint_values");
	if (stride > 7 + MGODPL_MAX_JOINT_VALUES) return fail(MGODPL_E_INVALID_ARGUMENT, "stride larger than 7 + MGODPL_MAX_JOINT_VALUES");
	return MGODPL_OK;
}

// ---- forward kinematics ------------------------------------------------------------------------------------------------
int mgodpl_forward_kinematics(const mgodpl_robot *robot, const double *states, size_t n, size_t stride, int32_t js,
							  double *out, void *stream_) {
	if (int rc = check_state_args(robot, stride, js)) return rc;
	if (n && (!states || !out)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (int rc = require_device()) return rc;
	cudaStream_t stream = (cudaStream_t) stream_;
	if (!n) return MGODPL_OK;
	size_t in_b = n * stride * 8, out_b = n * robot->dev.n_links * 7 * 8;
	Lease ws(global_pool());
	Carver c;
	const size_t o_in = c.take(in_b), o_out = c.take(out_b);
	CU(ws->reserve(c.off, 0));
	double *d_in = (double *) (ws->dev + o_in), *d_out = (double *) (ws->dev + o_out);
	CU(cudaMemcpyAsync(d_in, states, in_b, cudaMemcpyHostToDevice, stream));
	CU(launch_forward_kinematics(robot->dev, d_in, n, stride, d_out, stream));
	CU(cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	ws.settled = true;
	return MGODPL_OK;
}

// ---- boxes ---------------------------------------------------------------------------------------------------------------
int mgodpl_check_boxes(mgodpl_scene *scene, const double *boxes, size_t n, uint32_t *hit_bits, void *stream_) {
	if (!scene) return fail(MGODPL_E_INVALID_ARGUMENT, "scene is null");
	if (n && (!boxes || !hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n > SURVIVOR_MAX_ENTRIES) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^26 queries in one call: split the batch");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!n) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Lease ws(scene->pool);
	Carver c;
	const size_t words = (n + 31) / 32, qb = query_scratch_bytes(n * 9 / 8 + 4096);
	size_t o_in = c.take(n * 80), o_out = c.take(words * 4), o_q = c.take(qb);
	CU(ws->reserve(c.off, words * 4));
	CU(cudaMemcpyAsync(ws->dev + o_in, boxes, n * 80, cudaMemcpyHostToDevice, stream));
	CU(launch_check_boxes(scene->dev, (const double *) (ws->dev + o_in), n, (uint32_t *) (ws->dev + o_out), ws->dev + o_q, qb, stream));
	CU(cudaMemcpyAsync(ws->host, ws->dev + o_out, words * 4, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	memcpy(hit_bits, ws->host, words * 4);
	ws.settled = true;
	return MGODPL_OK;
}

// ---- states --------------------------------------------------------------------------------------------------------------
int mgodpl_check_states_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_states, size_t n, size_t stride,
							   int32_t js, uint32_t *d_hit_bits, void *stream_) {
	if (!scene) return fail(MGODPL_E_INVALID_ARGUMENT, "scene is null");
	if (int rc = check_state_args(robot, stride, js)) return rc;
	if (n && (!d_states || !d_hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n > SURVIVOR_MAX_ENTRIES) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^26 queries in one call: split the batch");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!n) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	const size_t qb = query_scratch_bytes(n * (size_t) robot->dev.n_boxes + 4096, state_survivor_slots(n));
	auto &slot = scene->stream_slot(stream);
	std::lock_guard<std::mutex> enqueue(slot.mu);
	CU(slot.ws.reserve(qb, 0));
	ROBOT_ON_SCENE_DEVICE(rdev);
	CU(launch_check_states(scene->dev, rdev, d_states, n, stride, d_hit_bits, slot.ws.dev, qb, stream));
	return MGODPL_OK;
}

int mgodpl_check_states(mgodpl_scene *scene, const mgodpl_robot *robot, const double *states, size_t n, size_t stride,
						int32_t js, uint32_t *hit_bits, void *stream_) {
	if (!scene) return fail(MGODPL_E_INVALID_ARGUMENT, "scene is null");
	if (int rc = check_state_args(robot, stride, js)) return rc;
	if (n && (!states || !hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n > SURVIVOR_MAX_ENTRIES) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^26 queries in one call: split the batch");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!n) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Lease ws(scene->pool);
	Carver c;
	const size_t in_b = n * stride * 8, words = (n + 31) / 32, qb = query_scratch_bytes(n * (size_t) robot->dev.n_boxes + 4096, state_survivor_slots(n));
	size_t o_in = c.take(in_b), o_out = c.take(words * 4), o_q = c.take(qb);
	CU(ws->reserve(c.off, words * 4));
	// Caller memory goes straight to the device; only the (small) result is staged through pinned memory.
	CU(cudaMemcpyAsync(ws->dev + o_in, states, in_b, cudaMemcpyHostToDevice, stream));
	ROBOT_ON_SCENE_DEVICE(rdev);
	CU(launch_check_states(scene->dev, rdev, (const double *) (ws->dev + o_in), n, stride, (uint32_t *) (ws->dev + o_out),
						   ws->dev + o_q, qb, stream));
	CU(cudaMemcpyAsync(ws->host, ws->dev + o_out, words * 4, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	memcpy(hit_bits, ws->host, words * 4);
	ws.settled = true;
	return MGODPL_OK;
}

// ---- motions ---------------------------------------------------------------------------------------------------------------
/// Work-list sizing for motions: what survives the clearance-grid trimming is only known on the device, so assume up to 64
/// steps per motion within reach of the scene (the PRM sampler on a single tree averages ~7).  A wrong guess only moves work
/// to the overflow paths: a full survivor list expands inline, a full queue traverses inline.
static size_t motion_survivors(size_t m) { return m * 64; }
static size_t motion_scratch_bytes(size_t m, int n_boxes) {
	// head and tail list hold m * 32 entries each; two list grains (4096 slots) of slack so that a pass (a whole number of grains)
	// covers a whole list: no survivor-count read-back, the call never blocks the host
	return query_scratch_bytes((m * 32 + 4096) * (size_t) n_boxes, motion_survivors(m)) + motion_overflow_bytes(m);
}

int mgodpl_check_motions_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_a, const double *d_b, size_t m,
								size_t stride, int32_t js, double max_step, const uint32_t *d_override, uint32_t *d_hit_bits,
								double *d_toi, uint32_t *d_n_steps, uint32_t *d_uncertain, void *stream_) {
	if (!scene) return fail(MGODPL_E_INVALID_ARGUMENT, "scene is null");
	if (int rc = check_state_args(robot, stride, js)) return rc;
	if (!(max_step > 0.0)) return fail(MGODPL_E_INVALID_ARGUMENT, "max_step must be positive");
	if (m && (!d_a || !d_b || !d_hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (m >= 0xFFFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^32-2 queries in one call");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!m) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Carver c;
	const size_t qb = motion_scratch_bytes(m, robot->dev.n_boxes) + motion_tail_queue_bytes(m, robot->dev.n_boxes);
	size_t o_first = c.take(m * 4), o_steps = c.take(d_n_steps ? 0 : m * 4), o_slerp = c.take(m * 32), o_q = c.take(qb);
	auto &slot = scene->stream_slot(stream);
	std::lock_guard<std::mutex> enqueue(slot.mu);
	CU(slot.ws.reserve(c.off, 0));
	CU(slot.overlap_streams());
	char *scratch = slot.ws.dev;
	ROBOT_ON_SCENE_DEVICE(rdev);
	CU(launch_check_motions(scene->dev, rdev, d_a, d_b, m, stride, js, max_step, d_override, d_hit_bits, d_toi,
							d_n_steps ? d_n_steps : (uint32_t *) (scratch + o_steps), d_uncertain, (unsigned *) (scratch + o_first),
							scratch + o_slerp, scratch + o_q, qb, motion_survivors(m), stream, &slot.overlap));
	return MGODPL_OK;
}

namespace {
/// equal_weights_distance + ceil exactly as the reference computes them on the host (distance.cpp:17-25,
/// Quaternion.h:78-85, collision_detection.cpp:88-92), with the host libm.
uint32_t host_n_steps(const double *a, const double *b, int js, double max_step) {
	double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
	double d = std::sqrt(dx * dx + dy * dy + dz * dz);
	double c = a[6] * b[6] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
	c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
	d += std::acos(2.0 * c * c - 1.0);
	for (int i = 0; i < js; ++i) d += std::abs(a[7 + i] - b[7 + i]);
	double n = std::ceil(d / max_step);
	if (!(n >= 0.0 && n < 268435455.0)) return 0;
	return (uint32_t) n;
}
}  // namespace

/// Ratio of consecutive chunk sizes in the host-buffer motion call (see there); 1 = equal chunks.
static constexpr double E2E_TAPER_DEFAULT = 0.7;
int mgodpl_check_motions(mgodpl_scene *scene, const mgodpl_robot *robot, const double *a, const double *b, size_t m, size_t stride,
						 int32_t js, double max_step, uint32_t *hit_bits, double *toi, uint32_t *n_steps_out, void *stream_) {
	if (!scene) return fail(MGODPL_E_INVALID_ARGUMENT, "scene is null");
	if (int rc = check_state_args(robot, stride, js)) return rc;
	if (!(max_step > 0.0)) return fail(MGODPL_E_INVALID_ARGUMENT, "max_step must be positive");
	if (m && (!a || !b || !hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (m >= 0xFFFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^32-2 queries in one call");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!m) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	static const bool trace = getenv("MGODPL_TRACE_E2E") != nullptr;  // dev aid: where a host-buffer call spends its time
	auto now = [] { return std::chrono::steady_clock::now(); };
	const auto t_begin = now();
	static thread_local cudaEvent_t tev[Workspace::MAX_CHUNKS][5] = {};  // trace only: copy start / end, kernels start / end, results on the host
	ROBOT_ON_SCENE_DEVICE(rdev);
	Lease ws(scene->pool);
	Carver c;
	// Large batches are cut into chunks.  All input copies go through one copy stream, back to back (the PCIe link carries input
	// from the first microsecond to the last byte); chunk k's kernels and result copies run on a stream of their own that waits for
	// "chunk k has arrived" only, so the kernels of the early chunks hide behind the input copies of the later ones.  A chunk's
	// kernels have a latency floor of ~0.12-0.15 ms however small it is, and concurrent chunks share the SMs: for 100 k motions
	// equal chunks measured 0.70 ms (1 chunk), 0.59 (2), 0.58 (3), 0.58 (4), 0.7 (6); four chunks whose sizes fall off by 0.7
	// (below) are the default (0.522 against 0.543-0.554 for three equal ones in the same run; tools/gpu_e2e_taper.py).  Chunk
	// boundaries are multiples of 32: result words never straddle.
	static const size_t chunk_pref = [] {
		const char *e = getenv("MGODPL_E2E_CHUNKS");
		return e && atoi(e) > 0 ? std::min((size_t) atoi(e), Workspace::MAX_CHUNKS) : (size_t) 4;
	}();
	static const bool use_graphs = [] {
		const char *e = getenv("MGODPL_E2E_GRAPHS");
		return !(e && atoi(e) == 0);
	}();
	// Chunk sizes fall off geometrically (chunk k+1 = taper x chunk k): what the call waits for in the end is the LAST chunk's
	// kernels, which cannot start before the last input byte has arrived, so that chunk should be small — down to where a chunk's
	// latency floor makes a smaller one pointless.  MGODPL_E2E_TAPER=1: equal chunks.
	static const double chunk_taper = [] {
		const char *e = getenv("MGODPL_E2E_TAPER");
		const double r = e ? atof(e) : E2E_TAPER_DEFAULT;
		return r > 0.05 && r <= 1.0 ? r : 1.0;
	}();
	const bool use_side = m >= 32768;  // large batches run on the workspace's own non-blocking streams, in chunk_pref chunks
	size_t n_chunks = use_side ? chunk_pref : 1;
	size_t cb[Workspace::MAX_CHUNKS + 1] = {0};  // chunk ci = motions [cb[ci], cb[ci + 1]), boundaries multiples of 32
	{
		double total = 0.0, w = 1.0, acc = 0.0;
		for (size_t ci = 0; ci < n_chunks; ++ci, w *= chunk_taper) total += w;
		size_t made = 0;
		w = 1.0;
		for (size_t ci = 0; ci < n_chunks; ++ci, w *= chunk_taper) {
			acc += w;
			size_t hi = ci + 1 == n_chunks ? m : std::min(m, ((size_t) ((double) m * (acc / total)) + 31) / 32 * 32);
			if (hi > cb[made]) cb[++made] = hi;
		}
		n_chunks = made ? made : 1;
		cb[n_chunks] = m;
	}
	size_t chunk_m = 0;  // the largest chunk: scratch and result blocks are sized for it
	for (size_t ci = 0; ci < n_chunks; ++ci) chunk_m = std::max(chunk_m, cb[ci + 1] - cb[ci]);
	const size_t in_b = m * stride * 8, words = (m + 31) / 32;
	const size_t qb = motion_scratch_bytes(chunk_m, robot->dev.n_boxes) + (use_side ? motion_tail_queue_bytes(chunk_m, robot->dev.n_boxes) : 0);
	size_t o_a = c.take(in_b), o_b = c.take(in_b), o_first = c.take(m * 4), o_slerp = c.take(m * 32);
	size_t o_hit = c.take(words * 4), o_unc = c.take(words * 4), o_steps = c.take(m * 4), o_toi = c.take(toi ? m * 8 : 0);
	const size_t out_lo = o_hit, out_hi = c.off;
	size_t o_ovr = c.take(m * 4);
	const size_t host_hi = c.off;
	size_t o_q[Workspace::MAX_CHUNKS];  // chunks run concurrently: a work queue each
	for (size_t ci = 0; ci < n_chunks; ++ci) o_q[ci] = c.take(qb);
	const size_t qb_all = c.off - o_q[0];  // (the rare whole-batch re-run uses the chunks' queues as one)
	// a chunk's results — hit words, boundary-case words, step counts, toi — sit in one block, so that they come back in ONE copy
	auto up = [](size_t x) { return (x + 255) & ~(size_t) 255; };
	struct BlockLayout {
		size_t unc, steps, toi, size;
	};
	auto block_layout = [&](size_t mc) {  // of a chunk of mc motions: only what the chunk needs travels back
		const size_t cw = (mc + 31) / 32;
		BlockLayout l;
		l.unc = up(cw * 4), l.steps = l.unc + up(cw * 4), l.toi = l.steps + up(mc * 4), l.size = l.toi + (toi ? up(mc * 8) : 0);
		return l;
	};
	const size_t b_size = block_layout(chunk_m).size;  // blocks are spaced for the largest chunk
	size_t o_blk[Workspace::MAX_CHUNKS];
	for (size_t ci = 0; ci < n_chunks; ++ci) o_blk[ci] = c.take(b_size);
	CU(ws->reserve(c.off, (host_hi - out_lo) + n_chunks * b_size));
	char *D = ws->dev, *H = ws->host - out_lo;  // host staging mirrors the device layout from out_lo onwards (whole-batch re-run) ...
	char *HB = ws->host + (host_hi - out_lo);    // ... followed by the chunks' result blocks
	if (use_side) {
		CU(ws->side_streams());
		CU(cudaEventRecord(ws->ev_start, stream));
		CU(cudaStreamWaitEvent(ws->copy_in, ws->ev_start, 0));
		for (size_t ci = 0; ci < n_chunks; ++ci) {
			const size_t lo = cb[ci], hi = cb[ci + 1];
			if (lo >= hi) break;
			if (trace) {
				for (int k = 0; k < 5; ++k)
					if (!tev[ci][k]) CU(cudaEventCreate(&tev[ci][k]));
				CU(cudaEventRecord(tev[ci][0], ws->copy_in));
			}
			CU(cudaMemcpyAsync(D + o_a + lo * stride * 8, a + lo * stride, (hi - lo) * stride * 8, cudaMemcpyHostToDevice, ws->copy_in));
			CU(cudaMemcpyAsync(D + o_b + lo * stride * 8, b + lo * stride, (hi - lo) * stride * 8, cudaMemcpyHostToDevice, ws->copy_in));
			CU(cudaEventRecord(ws->ev_in[ci], ws->copy_in));
			if (trace) CU(cudaEventRecord(tev[ci][1], ws->copy_in));
		}
	}
	for (size_t ci = 0; ci < n_chunks; ++ci) {
		const size_t lo = cb[ci], hi = cb[ci + 1];
		if (lo >= hi) break;
		const size_t mc = hi - lo, w_lo = lo / 32, w_n = (mc + 31) / 32;
		cudaStream_t s = use_side ? ws->side[ci] : stream;
		if (use_side) {
			CU(cudaStreamWaitEvent(s, ws->ev_in[ci], 0));
		} else {
			if (trace) {
				for (int k = 0; k < 5; ++k)
					if (!tev[ci][k]) CU(cudaEventCreate(&tev[ci][k]));
				CU(cudaEventRecord(tev[ci][0], s));
			}
			CU(cudaMemcpyAsync(D + o_a, a, in_b, cudaMemcpyHostToDevice, s));
			CU(cudaMemcpyAsync(D + o_b, b, in_b, cudaMemcpyHostToDevice, s));
			if (trace) CU(cudaEventRecord(tev[ci][1], s));
		}
		if (trace) CU(cudaEventRecord(tev[ci][2], s));  // (kernels may start here: the chunk has arrived)
		char *blk = D + o_blk[ci];
		const BlockLayout L = block_layout(mc);
		auto enqueue_chunk = [&](const MotionOverlap *ov) -> cudaError_t {
			cudaError_t e = launch_check_motions(scene->dev, rdev, (const double *) (D + o_a) + lo * stride, (const double *) (D + o_b) + lo * stride, mc,
												 stride, js, max_step, nullptr, (uint32_t *) blk, toi ? (double *) (blk + L.toi) : nullptr,
												 (uint32_t *) (blk + L.steps), (uint32_t *) (blk + L.unc), (unsigned *) (D + o_first) + lo,
												 D + o_slerp + lo * 32, D + o_q[ci], qb, motion_survivors(mc), s, ov);
			if (e != cudaSuccess) return e;
			if (trace && !ov) cudaEventRecord(tev[ci][3], s);
			e = cudaMemcpyAsync(HB + ci * b_size, blk, L.size, cudaMemcpyDeviceToHost, s);
			if (trace && !ov && e == cudaSuccess) e = cudaEventRecord(tev[ci][4], s);
			return e;
		};
		// Chunked calls replay the chunk's kernels + result copy from a cached CUDA graph when nothing about the call's shape has
		// changed since it was recorded (MGODPL_E2E_GRAPHS=0: plain launches).
		bool replayed = false;
		if (use_side && use_graphs && !ws->graphs_failed && !trace && motion_batch_is_one_pass(mc, robot->dev.n_boxes)) {
			const unsigned long long key[10] = {scene->uid, robot->uid, (unsigned long long) mc, (unsigned long long) lo, (unsigned long long) stride,
											   (unsigned long long) js, (unsigned long long) __builtin_bit_cast(unsigned long long, max_step), toi ? 1ull : 0ull,
											   (unsigned long long) (uintptr_t) blk ^ ((unsigned long long) (uintptr_t) (D + o_q[ci]) << 1), (unsigned long long) qb ^ ((unsigned long long) m << 32)};
			Workspace::ChunkGraph &cg = ws->chunk_graph[ci];
			if (!cg.exec || memcmp(cg.key, key, sizeof(key)) != 0) {
				if (cg.exec) cudaGraphExecDestroy(cg.exec);
				cg.exec = nullptr;
				cudaGraph_t graph = nullptr;
				cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
				if (e == cudaSuccess) {
					e = enqueue_chunk(&ws->overlap[ci]);
					const cudaError_t e_end = cudaStreamEndCapture(s, &graph);
					if (e == cudaSuccess) e = e_end;
				}
				if (e == cudaSuccess) e = cudaGraphInstantiate(&cg.exec, graph, 0);
				if (graph) cudaGraphDestroy(graph);
				if (e != cudaSuccess) {
					cudaGetLastError();
					cg.exec = nullptr, ws->graphs_failed = true;
				} else {
					memcpy(cg.key, key, sizeof(key));
				}
			}
			if (cg.exec) {
				CU(cudaGraphLaunch(cg.exec, s));
				replayed = true;
			}
		}
		if (!replayed) CU(enqueue_chunk(nullptr));
		if (use_side) {
			CU(cudaEventRecord(ws->ev_chunk[ci], s));
			CU(cudaStreamWaitEvent(stream, ws->ev_chunk[ci], 0));  // the caller's stream continues after every chunk
		}
	}
	const auto t_enqueued = now();
	// Step-count fix-up: the device libm may differ from the host's by an ulp, which only matters when distance /
	// max_step sits on an integer boundary.  Re-derive those on the host and re-run if any count disagrees.
	uint32_t *h_steps = (uint32_t *) (H + o_steps), *h_ovr = (uint32_t *) (H + o_ovr);
	size_t n_fix = 0;
	// Results of chunk ci = motions [lo, hi) are in pinned memory: look for step counts to fix up and hand the range to the caller.
	auto deliver = [&](size_t ci, size_t lo, size_t hi) {
		const char *hb = HB + ci * b_size;
		const BlockLayout L = block_layout(hi - lo);
		const size_t b_unc = L.unc, b_steps = L.steps, b_toi = L.toi;
		const uint32_t *c_hit = (const uint32_t *) hb, *c_unc = (const uint32_t *) (hb + b_unc), *c_steps = (const uint32_t *) (hb + b_steps);
		const size_t w_lo = lo / 32, w_n = (hi - lo + 31) / 32;
		for (size_t w = 0; w < w_n; ++w) {
			uint32_t bits = c_unc[w];
			while (bits) {
				int k = __builtin_ctz(bits);
				bits &= bits - 1;
				size_t i = lo + w * 32 + k;
				uint32_t want = host_n_steps(a + i * stride, b + i * stride, js, max_step);
				if (want != c_steps[i - lo]) {
					if (!n_fix) memset(h_ovr, 0xff, m * 4);
					h_ovr[i] = want;
					++n_fix;
				}
			}
		}
		memcpy(hit_bits + w_lo, c_hit, w_n * 4);
		if (toi) memcpy(toi + lo, hb + b_toi, (hi - lo) * 8);
		if (n_steps_out) memcpy(n_steps_out + lo, c_steps, (hi - lo) * 4);
	};
	if (use_side) {
		// a chunk's results are handed over while the chunks behind it are still on the GPU
		for (size_t ci = 0; ci < n_chunks; ++ci) {
			CU(cudaEventSynchronize(ws->ev_chunk[ci]));
			deliver(ci, cb[ci], cb[ci + 1]);
		}
	}
	CU(cudaStreamSynchronize(stream));
	if (!use_side) deliver(0, 0, m);
	const auto t_synced = now();
	if (n_fix) {
		// Rare path (the boundary set is normally empty for sampled states): re-run with the corrected counts pinned.
		CU(cudaMemcpyAsync(D + o_ovr, h_ovr, m * 4, cudaMemcpyHostToDevice, stream));
		CU(launch_check_motions(scene->dev, rdev, (const double *) (D + o_a), (const double *) (D + o_b), m, stride, js,
								max_step, (const uint32_t *) (D + o_ovr), (uint32_t *) (D + o_hit),
								toi ? (double *) (D + o_toi) : nullptr, (uint32_t *) (D + o_steps), nullptr,
								(unsigned *) (D + o_first), D + o_slerp, D + o_q[0], qb_all, motion_survivors(m), stream));
		CU(cudaMemcpyAsync(H + out_lo, D + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, stream));
		CU(cudaStreamSynchronize(stream));
		memcpy(hit_bits, H + o_hit, words * 4);
		if (toi) memcpy(toi, H + o_toi, m * 8);
		if (n_steps_out) memcpy(n_steps_out, h_steps, m * 4);
	}
	if (trace) {
		auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
		fprintf(stderr, "[mgodpl] check_motions m=%zu: enqueue %.0f us, wait + fix-up scan + copy-out %.0f us, re-run %.0f us (%zu re-derived)\n", m,
				us(t_begin, t_enqueued), us(t_enqueued, t_synced), us(t_synced, now()), n_fix);
		for (size_t ci = 0; ci < n_chunks; ++ci) {  // device timeline, microseconds after the first input copy started
			float t[5] = {};
			for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&t[k], tev[0][0], tev[ci][k]);
			fprintf(stderr, "[mgodpl]   chunk %zu (%zu motions): input copy %.0f-%.0f us, kernels %.0f-%.0f us, results on the host %.0f us\n", ci,
					cb[ci + 1] - cb[ci], t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3);
		}
	}
	ws.settled = true;
	return MGODPL_OK;
}

int mgodpl_check_path(mgodpl_scene *scene, const mgodpl_robot *robot, const double *states, size_t w, size_t stride, int32_t js,
					  double max_step, int64_t *first, void *stream) {
	if (!first) return fail(MGODPL_E_INVALID_ARGUMENT, "first_colliding_segment is null");
	*first = -1;
	if (w < 2) return MGODPL_OK;
	size_t m = w - 1;
	std::vector<uint32_t> bits((m + 31) / 32);
	// consecutive waypoint pairs: a = states[0..w-2], b = states[1..w-1] (collision_detection.cpp:111-113)
	int rc = mgodpl_check_motions(scene, robot, states, states + stride, m, stride, js, max_step, bits.data(), nullptr, nullptr, stream);
	if (rc) return rc;
	for (size_t i = 0; i < bits.size(); ++i)
		if (bits[i]) {
			*first = (int64_t) (i * 32 + __builtin_ctz(bits[i]));
			break;
		}
	return MGODPL_OK;
}

// ---- goal sampling -------------------------------------------------------------------------------------------------------------
static int check_goal_args(mgodpl_scene *scene, const mgodpl_robot *robot, size_t f, size_t s) {
	if (!scene || !robot) return fail(MGODPL_E_INVALID_ARGUMENT, "null handle");
	if (robot->dev.ee < 0) return fail(MGODPL_E_LINK_NOT_FOUND, "Link not found: end_effector");
	if (f && s && f * (((s + 31) / 32) * 32) >= 0xFFFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "more than 2^32-2 samples in one call");
	return MGODPL_OK;
}

int mgodpl_sample_goals_device(mgodpl_scene *scene, const mgodpl_robot *robot, const double *d_targets, size_t f, const double *d_draws,
							   size_t s, uint64_t seed, double dist, uint32_t *d_hit_bits, uint32_t *d_collisions, double *d_states_out,
							   void *stream_) {
	if (int rc = check_goal_args(scene, robot, f, s)) return rc;
	if (f && s && (!d_targets || !d_hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!f || !s) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	const size_t qb = query_scratch_bytes(f * s * (size_t) robot->dev.n_boxes * 9 / 8 + 4096);
	auto &slot = scene->stream_slot(stream);
	std::lock_guard<std::mutex> enqueue(slot.mu);
	CU(slot.ws.reserve(qb, 0));
	ROBOT_ON_SCENE_DEVICE(rdev);
	CU(launch_sample_goals(scene->dev, rdev, d_targets, f, d_draws, s, seed, dist, d_hit_bits, d_collisions,
						   d_states_out, slot.ws.dev, qb, stream));
	return MGODPL_OK;
}

int mgodpl_sample_goals(mgodpl_scene *scene, const mgodpl_robot *robot, const double *targets, size_t f, const double *draws, size_t s,
						uint64_t seed, double dist, uint32_t *hit_bits, uint32_t *collisions, double *states_out, void *stream_) {
	if (int rc = check_goal_args(scene, robot, f, s)) return rc;
	if (f && s && (!targets || !hit_bits)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (int rc = require_device()) return rc;
	DeviceGuard on_device(scene->device);
	if (!f || !s) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	const int nv = robot->dev.n_vars;
	Lease ws(scene->pool);
	Carver c;
	const size_t wpr = (s + 31) / 32, qb = query_scratch_bytes(f * s * (size_t) robot->dev.n_boxes * 9 / 8 + 4096);
	size_t o_t = c.take(f * 24), o_d = c.take(draws ? s * (1 + nv) * 8 : 0);
	size_t o_hit = c.take(f * wpr * 4), o_cnt = c.take(f * 4);
	const size_t small_hi = c.off;
	size_t o_st = c.take(states_out ? f * s * (7 + nv) * 8 : 0), o_q = c.take(qb);
	CU(ws->reserve(c.off, small_hi - o_hit));
	char *D = ws->dev, *H = ws->host - o_hit;  // host staging mirrors the device layout of the two result arrays
	CU(cudaMemcpyAsync(D + o_t, targets, f * 24, cudaMemcpyHostToDevice, stream));
	if (draws) CU(cudaMemcpyAsync(D + o_d, draws, s * (1 + nv) * 8, cudaMemcpyHostToDevice, stream));
	ROBOT_ON_SCENE_DEVICE(rdev);
	CU(launch_sample_goals(scene->dev, rdev, (const double *) (D + o_t), f, draws ? (const double *) (D + o_d) : nullptr,
						   s, seed, dist, (uint32_t *) (D + o_hit), (uint32_t *) (D + o_cnt),
						   states_out ? (double *) (D + o_st) : nullptr, D + o_q, qb, stream));
	CU(cudaMemcpyAsync(H + o_hit, D + o_hit, small_hi - o_hit, cudaMemcpyDeviceToHost, stream));
	if (states_out) CU(cudaMemcpyAsync(states_out, D + o_st, f * s * (7 + nv) * 8, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	memcpy(hit_bits, H + o_hit, f * wpr * 4);
	if (collisions) memcpy(collisions, H + o_cnt, f * 4);
	ws.settled = true;
	return MGODPL_OK;
}

// ---- k nearest neighbours over equal_weights_distance (the GNAT query of tsp_over_prm.cpp:44-45, batched) -------------------
int mgodpl_knn_device(const double *d_states, size_t n, size_t stride, int32_t js, int32_t k, int32_t causal, int32_t *d_idx,
					  double *d_dist, void *stream_) {
	if (js < 0 || js > MGODPL_MAX_JOINT_VALUES || stride < (size_t) (7 + js)) return fail(MGODPL_E_INVALID_ARGUMENT, "bad state layout");
	if (k < 1 || k > 32) return fail(MGODPL_E_INVALID_ARGUMENT, "k must be in 1..32");
	if (n && (!d_states || !d_idx)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n >= 0x7FFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "too many states");
	if (int rc = require_device()) return rc;
	if (!n) return MGODPL_OK;
	auto &slot = global_stream_slot((cudaStream_t) stream_);
	std::lock_guard<std::mutex> enqueue(slot.mu);
	CU(slot.ws.reserve(knn_scratch_bytes(n, n, k), 0));
	CU(launch_knn(d_states, n, stride, js, k, causal, d_idx, d_dist, slot.ws.dev, (cudaStream_t) stream_));
	return MGODPL_OK;
}

int mgodpl_knn(const double *states, size_t n, size_t stride, int32_t js, int32_t k, int32_t causal, int32_t *idx, double *dist,
			   void *stream_) {
	if (js < 0 || js > MGODPL_MAX_JOINT_VALUES || stride < (size_t) (7 + js)) return fail(MGODPL_E_INVALID_ARGUMENT, "bad state layout");
	if (k < 1 || k > 32) return fail(MGODPL_E_INVALID_ARGUMENT, "k must be in 1..32");
	if (n && (!states || !idx)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n >= 0x7FFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "too many states");
	if (int rc = require_device()) return rc;
	if (!n) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Lease ws(global_pool());
	Carver c;
	const size_t o_s = c.take(n * stride * 8), o_i = c.take(n * k * 4), o_d = c.take(dist ? n * k * 8 : 0), o_p = c.take(knn_scratch_bytes(n, n, k));
	CU(ws->reserve(c.off, 0));
	double *d_s = (double *) (ws->dev + o_s), *d_d = dist ? (double *) (ws->dev + o_d) : nullptr;
	int32_t *d_i = (int32_t *) (ws->dev + o_i);
	CU(cudaMemcpyAsync(d_s, states, n * stride * 8, cudaMemcpyHostToDevice, stream));
	CU(launch_knn(d_s, n, stride, js, k, causal, d_i, d_d, ws->dev + o_p, stream));
	CU(cudaMemcpyAsync(idx, d_i, n * k * 4, cudaMemcpyDeviceToHost, stream));
	if (dist) CU(cudaMemcpyAsync(dist, d_d, n * k * 8, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	ws.settled = true;
	return MGODPL_OK;
}

int mgodpl_knn_query(const double *database, size_t n_db, const double *queries, size_t n_q, size_t stride, int32_t js, int32_t k,
					 int32_t *idx, double *dist, void *stream_) {
	if (js < 0 || js > MGODPL_MAX_JOINT_VALUES || stride < (size_t) (7 + js)) return fail(MGODPL_E_INVALID_ARGUMENT, "bad state layout");
	if (k < 1 || k > 32) return fail(MGODPL_E_INVALID_ARGUMENT, "k must be in 1..32");
	if ((n_db && !database) || (n_q && (!queries || !idx))) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n_db >= 0x7FFFFFFFull || n_q >= 0x7FFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "too many states");
	if (int rc = require_device()) return rc;
	if (!n_q) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Lease ws(global_pool());
	Carver c;
	const size_t o_db = c.take(n_db * stride * 8 + 8), o_q = c.take(n_q * stride * 8), o_i = c.take(n_q * k * 4), o_d = c.take(dist ? n_q * k * 8 : 0);
	const size_t o_p = c.take(knn_scratch_bytes(n_q, n_db, k));
	CU(ws->reserve(c.off, 0));
	double *d_db = (double *) (ws->dev + o_db), *d_q = (double *) (ws->dev + o_q), *d_d = dist ? (double *) (ws->dev + o_d) : nullptr;
	int32_t *d_i = (int32_t *) (ws->dev + o_i);
	if (n_db) CU(cudaMemcpyAsync(d_db, database, n_db * stride * 8, cudaMemcpyHostToDevice, stream));
	CU(cudaMemcpyAsync(d_q, queries, n_q * stride * 8, cudaMemcpyHostToDevice, stream));
	CU(launch_knn_query(d_db, n_db, d_q, n_q, stride, js, k, d_i, d_d, ws->dev + o_p, stream));
	CU(cudaMemcpyAsync(idx, d_i, n_q * k * 4, cudaMemcpyDeviceToHost, stream));
	if (dist) CU(cudaMemcpyAsync(dist, d_d, n_q * k * 8, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	ws.settled = true;
	return MGODPL_OK;
}

// ---- roadmap shortest paths (runDijkstra per goal sample, tsp_over_prm.cpp:80-97 and :497-517, batched) ---------------------
namespace {
unsigned *pinned_flag() {
	static thread_local unsigned *flag = nullptr;
	if (!flag && cudaMallocHost((void **) &flag, 64) != cudaSuccess) flag = nullptr, cudaGetLastError();
	return flag;
}
}  // namespace

int mgodpl_roadmap_shortest_paths_device(const uint32_t *d_row_offsets, const uint32_t *d_neighbours, const double *d_weights,
										 uint32_t n_vertices, const uint32_t *d_sources, size_t n_sources, double *d_dist, uint32_t *d_pred,
										 uint32_t *sweeps_out, void *stream_) {
	if (n_vertices && n_sources && (!d_row_offsets || !d_sources || !d_dist)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (int rc = require_device()) return rc;
	if (sweeps_out) *sweeps_out = 0;
	if (!n_vertices || !n_sources) return MGODPL_OK;
	unsigned *flag = pinned_flag();
	if (!flag) return fail(MGODPL_E_OUT_OF_MEMORY, "pinned host memory");
	Lease ws(global_pool());
	CU(ws->reserve(roadmap_scratch_bytes(n_vertices, n_sources), 0));
	CU(launch_roadmap_shortest_paths(d_row_offsets, d_neighbours, d_weights, n_vertices, d_sources, n_sources, d_dist, d_pred, ws->dev, flag,
									 sweeps_out, (cudaStream_t) stream_));
	CU(cudaStreamSynchronize((cudaStream_t) stream_));
	ws.settled = true;
	return MGODPL_OK;
}

int mgodpl_roadmap_repair_predecessors(uint32_t n_vertices, const uint32_t *row, const uint32_t *col, const double *w, const double *ds,
									   uint32_t *ps) {
	if (n_vertices && (!row || !ds || !ps)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (n_vertices && row[n_vertices] && (!col || !w)) return fail(MGODPL_E_INVALID_ARGUMENT, "null adjacency");
	for (uint32_t v = 0; v < n_vertices; ++v)
		if (ps[v] >= n_vertices) return fail(MGODPL_E_INVALID_ARGUMENT, "predecessor out of range");
	// members: vertices whose predecessor sits at the same distance (the edge between them vanished in the sum): decide those again
	std::vector<uint32_t> frontier, members;
	std::vector<char> settled(n_vertices, 1);
	for (uint32_t v = 0; v < n_vertices; ++v)
		if (ps[v] != v && ds[ps[v]] == ds[v]) settled[v] = 0, members.push_back(v);
	if (members.empty()) return MGODPL_OK;
	// a member with a strictly nearer neighbour on a shortest path takes that one (ties: nearer the source, then lower index)
	for (uint32_t v: members) {
		uint32_t best = v;
		for (uint32_t k = row[v]; k < row[v + 1]; ++k) {
			const uint32_t u = col[k];
			if (u != v && ds[u] < ds[v] && ds[u] + w[k] == ds[v] && (best == v || ds[u] < ds[best] || (ds[u] == ds[best] && u < best))) best = u;
		}
		if (best != v) ps[v] = best, settled[v] = 1, frontier.push_back(v);
	}
	// every settled vertex that reaches a still-undecided member over a vanishing edge starts the breadth-first search as well
	for (uint32_t v: members)
		if (!settled[v])
			for (uint32_t k = row[v]; k < row[v + 1]; ++k) {
				const uint32_t u = col[k];
				if (u != v && settled[u] && ds[u] == ds[v] && ds[u] + w[k] == ds[v]) frontier.push_back(u);
			}
	std::sort(frontier.begin(), frontier.end());
	frontier.erase(std::unique(frontier.begin(), frontier.end()), frontier.end());
	for (size_t head = 0; head < frontier.size(); ++head) {
		const uint32_t u = frontier[head];
		for (uint32_t k = row[u]; k < row[u + 1]; ++k) {
			const uint32_t v = col[k];
			if (!settled[v] && ds[v] == ds[u] && ds[u] + w[k] == ds[v]) ps[v] = u, settled[v] = 1, frontier.push_back(v);
		}
	}
	return MGODPL_OK;
}

int mgodpl_roadmap_shortest_paths(uint32_t n_vertices, const uint32_t *edges, const double *weights, size_t n_edges,
								  const uint32_t *sources, size_t n_sources, double *dist, uint32_t *pred, uint32_t *sweeps_out,
								  void *stream_) {
	if (n_edges && (!edges || !weights)) return fail(MGODPL_E_INVALID_ARGUMENT, "null edge list");
	if (n_sources && n_vertices && (!sources || !dist)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (2 * n_edges >= 0xFFFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "too many edges");
	for (size_t e = 0; e < n_edges; ++e) {
		if (edges[2 * e] >= n_vertices || edges[2 * e + 1] >= n_vertices) return fail(MGODPL_E_INVALID_ARGUMENT, "edge endpoint out of range");
		// boost::dijkstra_shortest_paths throws negative_edge; NaN weights would never compare
		if (!(weights[e] >= 0.0) || std::isinf(weights[e])) return fail(MGODPL_E_INVALID_ARGUMENT, "edge weights must be finite and non-negative");
	}
	for (size_t s = 0; s < n_sources; ++s)
		if (sources[s] >= n_vertices) return fail(MGODPL_E_INVALID_ARGUMENT, "source vertex out of range");
	if (int rc = require_device()) return rc;
	if (sweeps_out) *sweeps_out = 0;
	if (!n_vertices || !n_sources) return MGODPL_OK;
	// CSR with both directions of every undirected edge, neighbours in edge-list order (boost's out-edge order)
	std::vector<uint32_t> row(n_vertices + 1, 0), col(2 * n_edges);
	std::vector<double> w(2 * n_edges);
	for (size_t e = 0; e < n_edges; ++e) ++row[edges[2 * e] + 1], ++row[edges[2 * e + 1] + 1];
	for (uint32_t v = 0; v < n_vertices; ++v) row[v + 1] += row[v];
	{
		std::vector<uint32_t> cursor(row.begin(), row.end() - 1);
		for (size_t e = 0; e < n_edges; ++e) {
			const uint32_t a = edges[2 * e], b = edges[2 * e + 1];
			col[cursor[a]] = b, w[cursor[a]++] = weights[e];
			col[cursor[b]] = a, w[cursor[b]++] = weights[e];
		}
	}
	cudaStream_t stream = (cudaStream_t) stream_;
	const size_t cells = n_sources * (size_t) n_vertices;
	Lease ws(global_pool());
	Carver c;
	const size_t o_row = c.take(row.size() * 4), o_col = c.take(col.size() * 4 + 4), o_w = c.take(w.size() * 8 + 8), o_src = c.take(n_sources * 4);
	const size_t o_dist = c.take(cells * 8), o_pred = c.take(pred ? cells * 4 : 0);
	CU(ws->reserve(c.off, 0));
	uint32_t *d_row = (uint32_t *) (ws->dev + o_row), *d_col = (uint32_t *) (ws->dev + o_col), *d_src = (uint32_t *) (ws->dev + o_src);
	uint32_t *d_pred = pred ? (uint32_t *) (ws->dev + o_pred) : nullptr;
	double *d_w = (double *) (ws->dev + o_w), *d_dist = (double *) (ws->dev + o_dist);
	CU(cudaMemcpyAsync(d_row, row.data(), row.size() * 4, cudaMemcpyHostToDevice, stream));
	if (!col.empty()) CU(cudaMemcpyAsync(d_col, col.data(), col.size() * 4, cudaMemcpyHostToDevice, stream));
	if (!w.empty()) CU(cudaMemcpyAsync(d_w, w.data(), w.size() * 8, cudaMemcpyHostToDevice, stream));
	CU(cudaMemcpyAsync(d_src, sources, n_sources * 4, cudaMemcpyHostToDevice, stream));
	// (the device variant leases a second workspace for its label scratch and synchronises the stream)
	if (int rc = mgodpl_roadmap_shortest_paths_device(d_row, d_col, d_w, n_vertices, d_src, n_sources, d_dist, d_pred, sweeps_out, stream_)) return rc;
	CU(cudaMemcpyAsync(dist, d_dist, cells * 8, cudaMemcpyDeviceToHost, stream));
	if (pred) CU(cudaMemcpyAsync(pred, d_pred, cells * 4, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	ws.settled = true;
	// Edges so short that they vanish in the sum (zero weight, or below half an ulp of the distance) join vertices at ONE distance,
	// and the device, which picks a predecessor per vertex on its own, may let such vertices name each other.  The distances are
	// exact; mgodpl_roadmap_repair_predecessors makes the predecessors a tree again (what boost's Dijkstra always returns).
	// No distance exceeds the sum of all weights, so an edge can only vanish if w <= 2^-53 * that sum: graphs without such an
	// edge — every roadmap of distinct states — skip the pass.
	if (pred) {
		double total = 0.0;
		for (size_t e = 0; e < n_edges; ++e) total += weights[e];
		bool may_vanish = false;
		for (size_t e = 0; e < n_edges && !may_vanish; ++e) may_vanish = edges[2 * e] != edges[2 * e + 1] && weights[e] <= total * 0x1p-52;
		if (may_vanish)
			for (size_t s = 0; s < n_sources; ++s)
				mgodpl_roadmap_repair_predecessors(n_vertices, row.data(), col.data(), w.data(), dist + s * n_vertices, pred + s * n_vertices);
	}
	return MGODPL_OK;
}

// ---- connected components of a mesh (fruit separation, mesh_connected_components.cpp:25-74) ---------------------------------
}  // extern "C" (templates cannot have C linkage)
namespace {
template<class Index>
int mesh_components_impl(const Index *tris, size_t nt, size_t nv, uint32_t *labels, void *stream_) {
	if ((nt && !tris) || (nv && !labels)) return fail(MGODPL_E_INVALID_ARGUMENT, "null buffer");
	if (nv >= 0xFFFFFFFFull) return fail(MGODPL_E_INVALID_ARGUMENT, "too many vertices");
	for (size_t i = 0; i < 3 * nt; ++i)
		if ((size_t) tris[i] >= nv) return fail(MGODPL_E_INVALID_ARGUMENT, "triangle index out of range");
	if (int rc = require_device()) return rc;
	if (!nv) return MGODPL_OK;
	cudaStream_t stream = (cudaStream_t) stream_;
	Lease ws(global_pool());
	Carver c;
	const size_t o_t = c.take(3 * nt * sizeof(Index) + 8), o_l = c.take(nv * 4);
	CU(ws->reserve(c.off, 0));
	Index *d_t = (Index *) (ws->dev + o_t);
	unsigned *d_l = (unsigned *) (ws->dev + o_l);
	if (nt) CU(cudaMemcpyAsync(d_t, tris, 3 * nt * sizeof(Index), cudaMemcpyHostToDevice, stream));
	CU(launch_mesh_components<Index>(d_t, nt, (unsigned) nv, d_l, stream));
	CU(cudaMemcpyAsync(labels, d_l, nv * 4, cudaMemcpyDeviceToHost, stream));
	CU(cudaStreamSynchronize(stream));
	ws.settled = true;
	return MGODPL_OK;
}
}  // namespace
extern "C" {

int mgodpl_mesh_components(const uint32_t *tris, size_t nt, size_t nv, uint32_t *labels, void *stream) {
	return mesh_components_impl<uint32_t>(tris, nt, nv, labels, stream);
}
int mgodpl_mesh_components_u64(const uint64_t *tris, size_t nt, size_t nv, uint32_t *labels, void *stream) {
	return mesh_components_impl<uint64_t>(tris, nt, nv, labels, stream);
}

}  // extern "C"
