// rtb_lbvh.cu — world BVH built on the GPU (RTB_WORLD_BVH_GPU_LBVH).
//
// The reference builds its BVH on the host, recursively (BVH_Handle::Factory, main/src/rt_engine/geometry/BVH.cu:166-383:
// median split with a std::sort per level, or binned SAH, or an O(n^3) bottom-up merge).  This is the device-side
// alternative for scenes whose build time matters (large meshes, scenes rebuilt per frame): a linear BVH
//   1. 63-bit Morton code of every primitive's box centre, quantised to 2^21 cells per axis of the centre bounds,
//   2. one radix sort of (code, primitive) pairs (cub::DeviceRadixSort — library code, like cuBLAS for a plain GEMM),
//   3. the whole hierarchy in one pass, every inner node found independently from the sorted codes (the
//      longest-common-prefix construction: node i covers the maximal run of keys that share a longer prefix with key i
//      than key i shares with its other neighbour; equal codes are told apart by their sorted position),
//   4. boxes fitted bottom-up, the second thread to arrive at a node forming the union of its children.
// The result is handed back in the reference's node layout (rtb_bvh_node: box, left, right-or-primitive), so the
// flattener's conversion to the wide traversal layout, the depth check against the traversal stack and the
// rtb_scene_world_bvh parity hook are shared with the host builders.  Closest hits do not depend on the tree.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "rtb_scene.h"

namespace rtb {
namespace {

#define LBVH_TRY(expr)                                                                                              \
	do {                                                                                                           \
		cudaError_t e_ = (expr);                                                                                   \
		if (e_ != cudaSuccess) { rc = fail(RTB_ERR_CUDA, std::string("lbvh: ") + #expr + ": " + cudaGetErrorString(e_)); goto done; } \
	} while (0)

struct Frame { float lo[3], inv[3]; };   // centre bounds: cell = (c - lo) * inv, inv = 2^21 / extent (0 on a flat axis)

__device__ __forceinline__ unsigned long long spread21(unsigned int v) {   // bit k of v -> bit 3k
	unsigned long long x = v & 0x1fffffull;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

__global__ void lbvh_codes_kernel(const float* __restrict__ boxes, int n, Frame fr, unsigned long long* __restrict__ codes, int* __restrict__ ids) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* b = boxes + 6 * (size_t)i;
	unsigned int q[3];
	for (int k = 0; k < 3; ++k) {
		float c = 0.5f * b[k] + 0.5f * b[3 + k];
		float g = (c - fr.lo[k]) * fr.inv[k];
		g = fminf(fmaxf(g, 0.0f), 2097151.0f);
		q[k] = (unsigned int)g;
	}
	codes[i] = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
	ids[i] = i;
}

// Length of the common prefix of keys i and j (sorted positions), -1 outside the array.  Equal codes compare by position.
__device__ __forceinline__ int prefix(const unsigned long long* __restrict__ codes, int n, int i, int j) {
	if (j < 0 || j >= n) return -1;
	unsigned long long a = codes[i], b = codes[j];
	return a != b ? __clzll((long long)(a ^ b)) : 64 + __clz(i ^ j);
}

// Node numbering: inner nodes 0 .. n-2 (root 0), leaf of sorted position k = n-1+k.
__global__ void lbvh_hierarchy_kernel(const unsigned long long* __restrict__ codes, int n, rtb_bvh_node* __restrict__ nodes, int* __restrict__ parent) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	const int d = prefix(codes, n, i, i + 1) - prefix(codes, n, i, i - 1) > 0 ? 1 : -1;   // which way the node's run extends
	const int floor_prefix = prefix(codes, n, i, i - d);
	int reach = 2;
	while (prefix(codes, n, i, i + reach * d) > floor_prefix) reach <<= 1;
	int len = 0;
	for (int t = reach >> 1; t >= 1; t >>= 1)
		if (prefix(codes, n, i, i + (len + t) * d) > floor_prefix) len += t;
	const int j = i + len * d;                                                          // other end of the run
	const int node_prefix = prefix(codes, n, i, j);
	int s = 0, t = len;
	do {                                                                                // last position sharing more than node_prefix with i
		t = (t + 1) >> 1;
		if (prefix(codes, n, i, i + (s + t) * d) > node_prefix) s += t;
	} while (t > 1);
	const int split = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	const int left = lo == split ? n - 1 + split : split;
	const int right = hi == split + 1 ? n - 1 + split + 1 : split + 1;
	nodes[i].left_child_idx = left;
	nodes[i].right_child_hittable_idx = right;
	parent[left] = i; parent[right] = i;
	if (i == 0) parent[0] = -1;
}

__global__ void lbvh_fit_kernel(const float* __restrict__ boxes, const int* __restrict__ ids, int n, rtb_bvh_node* nodes, const int* __restrict__ parent, int* arrived) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	int node = n - 1 + k;
	{
		const float* b = boxes + 6 * (size_t)ids[k];
		rtb_bvh_node leaf;
		for (int a = 0; a < 3; ++a) { leaf.bmin[a] = b[a]; leaf.bmax[a] = b[3 + a]; }
		leaf.left_child_idx = -1; leaf.right_child_hittable_idx = k;
		nodes[node] = leaf;
	}
	int p = n > 1 ? parent[node] : -1;
	while (p >= 0) {
		__threadfence();                                       // the box written above is visible before the arrival is counted
		if (atomicAdd(&arrived[p], 1) == 0) return;            // first child to arrive: the sibling's thread finishes this node
		__threadfence();
		const volatile rtb_bvh_node* l = nodes + nodes[p].left_child_idx;
		const volatile rtb_bvh_node* r = nodes + nodes[p].right_child_hittable_idx;
		for (int a = 0; a < 3; ++a) {
			float l0 = l->bmin[a], r0 = r->bmin[a], l1 = l->bmax[a], r1 = r->bmax[a];
			nodes[p].bmin[a] = r0 < l0 ? r0 : l0;              // aabb::operator+= (aabb.cuh:24): component-wise min / max
			nodes[p].bmax[a] = l1 < r1 ? r1 : l1;
		}
		p = parent[p];
	}
}

}  // namespace

void free_gpu_scratch(GpuScratch& scratch) {
	cudaFree(scratch.device); cudaFreeHost(scratch.pinned);
	scratch = GpuScratch{};
}

int build_bvh_lbvh_gpu(const std::vector<Box3>& boxes, std::vector<rtb_bvh_node>& nodes, std::vector<int>& order, int& root, const GpuBuildContext& gpu) {
	const int n = (int)boxes.size();
	if (n == 0) return fail(RTB_ERR_INVALID, "build_bvh_lbvh_gpu: no primitives");
	if (!gpu.scratch) return fail(RTB_ERR_INVALID, "build_bvh_lbvh_gpu: no scratch");
	cudaStream_t st = static_cast<cudaStream_t>(gpu.stream);
	const int device = gpu.device;
	GpuScratch& scratch = *gpu.scratch;
	int rc = RTB_OK;
	float* d_boxes = nullptr; unsigned long long *d_codes = nullptr, *d_codes_sorted = nullptr; int *d_ids = nullptr, *d_ids_sorted = nullptr, *d_parent = nullptr, *d_arrived = nullptr;
	rtb_bvh_node* d_nodes = nullptr; void* d_temp = nullptr; size_t temp_bytes = 0; uint8_t* d_arena = nullptr;
	const int n_nodes = 2 * n - 1, threads = 256, blocks = (n + threads - 1) / threads;

	// Bounds of the box centres (one host pass over data that is already in cache from the flattener).
	Frame fr;
	{
		float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
		for (const Box3& b : boxes)
			for (int k = 0; k < 3; ++k) {
				float c = 0.5f * b.mn[k] + 0.5f * b.mx[k];
				if (c < lo[k]) lo[k] = c;
				if (c > hi[k]) hi[k] = c;
			}
		for (int k = 0; k < 3; ++k) {
			float ext = hi[k] - lo[k];
			fr.lo[k] = lo[k];
			fr.inv[k] = (ext > 0.0f && ext < INFINITY) ? 2097152.0f / ext : 0.0f;
			if (!(fr.lo[k] > -INFINITY && fr.lo[k] < INFINITY)) { fr.lo[k] = 0.0f; fr.inv[k] = 0.0f; }
		}
	}

	const bool trace = getenv("RTB_LBVH_TRACE") != nullptr;
	auto t0 = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!trace) return;
		cudaStreamSynchronize(st);
		auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[lbvh] %-12s %8.3f ms\n", what, std::chrono::duration<float, std::milli>(t1 - t0).count());
		t0 = t1;
	};
	lap("bounds");
	LBVH_TRY(cudaSetDevice(device));
	// one scratch arena kept by the renderer between builds, carved up (every section 256-byte aligned)
	LBVH_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_codes, d_codes_sorted, d_ids, d_ids_sorted, n, 0, 63, st));
	{
		size_t off = 0;
		auto take = [&off](size_t bytes) { size_t at = off; off = (off + bytes + 255) & ~(size_t)255; return at; };
		const size_t o_boxes = take(sizeof(Box3) * (size_t)n), o_codes = take(8 * (size_t)n), o_codes2 = take(8 * (size_t)n), o_ids = take(4 * (size_t)n),
		             o_ids2 = take(4 * (size_t)n), o_parent = take(4 * (size_t)n_nodes), o_arrived = take(4 * (size_t)n),
		             o_nodes = take(sizeof(rtb_bvh_node) * (size_t)n_nodes), o_temp = take(temp_bytes ? temp_bytes : 1);
		if (off > scratch.device_bytes) {
			cudaFree(scratch.device); scratch.device = nullptr; scratch.device_bytes = 0;
			LBVH_TRY(cudaMalloc(&scratch.device, off + off / 4));
			scratch.device_bytes = off + off / 4;
		}
		const size_t host_bytes = sizeof(rtb_bvh_node) * (size_t)n_nodes + 4 * (size_t)n;
		if (host_bytes > scratch.pinned_bytes) {
			cudaFreeHost(scratch.pinned); scratch.pinned = nullptr; scratch.pinned_bytes = 0;
			LBVH_TRY(cudaMallocHost(&scratch.pinned, host_bytes + host_bytes / 4));
			scratch.pinned_bytes = host_bytes + host_bytes / 4;
		}
		d_arena = static_cast<uint8_t*>(scratch.device);
		d_boxes = reinterpret_cast<float*>(d_arena + o_boxes);
		d_codes = reinterpret_cast<unsigned long long*>(d_arena + o_codes); d_codes_sorted = reinterpret_cast<unsigned long long*>(d_arena + o_codes2);
		d_ids = reinterpret_cast<int*>(d_arena + o_ids); d_ids_sorted = reinterpret_cast<int*>(d_arena + o_ids2);
		d_parent = reinterpret_cast<int*>(d_arena + o_parent); d_arrived = reinterpret_cast<int*>(d_arena + o_arrived);
		d_nodes = reinterpret_cast<rtb_bvh_node*>(d_arena + o_nodes); d_temp = d_arena + o_temp;
	}
	lap("malloc");
	LBVH_TRY(cudaMemcpyAsync(d_boxes, boxes.data(), sizeof(Box3) * (size_t)n, cudaMemcpyHostToDevice, st));
	LBVH_TRY(cudaMemsetAsync(d_arrived, 0, 4 * (size_t)n, st));
	lap("h2d");
	lbvh_codes_kernel<<<blocks, threads, 0, st>>>(d_boxes, n, fr, d_codes, d_ids);
	lap("codes");
	LBVH_TRY(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_codes, d_codes_sorted, d_ids, d_ids_sorted, n, 0, 63, st));
	lap("sort");
	if (n > 1) lbvh_hierarchy_kernel<<<(n - 1 + threads - 1) / threads, threads, 0, st>>>(d_codes_sorted, n, d_nodes, d_parent);
	lap("hierarchy");
	lbvh_fit_kernel<<<blocks, threads, 0, st>>>(d_boxes, d_ids_sorted, n, d_nodes, d_parent, d_arrived);
	LBVH_TRY(cudaGetLastError());
	lap("fit");

	{
		rtb_bvh_node* h_nodes = static_cast<rtb_bvh_node*>(scratch.pinned);
		int* h_order = reinterpret_cast<int*>(h_nodes + n_nodes);
		LBVH_TRY(cudaMemcpyAsync(h_nodes, d_nodes, sizeof(rtb_bvh_node) * (size_t)n_nodes, cudaMemcpyDeviceToHost, st));
		LBVH_TRY(cudaMemcpyAsync(h_order, d_ids_sorted, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
		LBVH_TRY(cudaStreamSynchronize(st));
		nodes.assign(h_nodes, h_nodes + n_nodes); order.assign(h_order, h_order + n);
	}
	root = 0;
	lap("d2h");
done:
	return rc;
}

}  // namespace rtb
