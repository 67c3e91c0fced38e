// Connected components of a triangle mesh's vertex graph — how the reference splits the combined "fruit" mesh of a tree
// model into individual fruit (connected_vertex_components, src/experiment_utils/mesh_connected_components.cpp:25-74,
// used by loadTreeMeshes, src/experiment_utils/TreeMeshes.cpp:28-66).  The reference merges vertex lists triangle by
// triangle on the host; here it is a lock-free union-find: every triangle hooks its two edges (the third follows by
// transitivity, as in the reference), always the larger root under the smaller, so a component's label is its
// smallest vertex index — a canonical answer independent of thread order.
#include "kernels.h"

namespace mgodpl_b200 {

namespace {

__device__ __forceinline__ unsigned cc_root(unsigned *parent, unsigned v) {
	for (;;) {
		const unsigned p = *(volatile unsigned *) (parent + v);
		if (p == v) return v;
		const unsigned gp = *(volatile unsigned *) (parent + p);
		if (gp != p) parent[v] = gp;  // path halving: racing writers only ever store an ancestor
		v = p;
	}
}

__device__ __forceinline__ void cc_union(unsigned *parent, unsigned a, unsigned b) {
	for (;;) {
		a = cc_root(parent, a), b = cc_root(parent, b);
		if (a == b) return;
		const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
		if (atomicCAS(parent + hi, hi, lo) == hi) return;  // hi was still a root: hooked; otherwise retry from the new roots
	}
}

__global__ void k_cc_init(unsigned *parent, unsigned nv) {
	const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < nv) parent[v] = v;
}

template<class Index>
__global__ void k_cc_hook(const Index *__restrict__ tris, size_t nt, unsigned *parent) {
	const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nt) return;
	const unsigned a = (unsigned) tris[3 * t], b = (unsigned) tris[3 * t + 1], c = (unsigned) tris[3 * t + 2];
	cc_union(parent, a, b);
	cc_union(parent, b, c);
}

/// After the last hook every root is final.  Each thread walks up read-only and writes only its own entry: path halving here
/// could overwrite an entry its owner has already set to the root with an older, non-root ancestor.
__global__ void k_cc_flatten(unsigned *parent, unsigned nv) {
	const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nv) return;
	unsigned r = v;
	for (unsigned p; (p = *(volatile unsigned *) (parent + r)) != r;) r = p;
	parent[v] = r;
}

}  // namespace

template<class Index>
cudaError_t launch_mesh_components(const Index *d_tris, size_t nt, unsigned nv, unsigned *d_labels, cudaStream_t stream) {
	if (!nv) return cudaSuccess;
	k_cc_init<<<(nv + 255) / 256, 256, 0, stream>>>(d_labels, nv);
	if (nt) k_cc_hook<Index><<<(unsigned) ((nt + 255) / 256), 256, 0, stream>>>(d_tris, nt, d_labels);
	k_cc_flatten<<<(nv + 255) / 256, 256, 0, stream>>>(d_labels, nv);
	return cudaGetLastError();
}
template cudaError_t launch_mesh_components<uint32_t>(const uint32_t *, size_t, unsigned, unsigned *, cudaStream_t);
template cudaError_t launch_mesh_components<uint64_t>(const uint64_t *, size_t, unsigned, unsigned *, cudaStream_t);

}  // namespace mgodpl_b200
