// oracle/ref_bvh_harness.cu — TEST INFRASTRUCTURE.  Drives the REFERENCE's own BVH builder
// (BVH_Handle::Factory in /root/reference/main/src/rt_engine/geometry/BVH.cu, compiled from where it
// lies, unmodified) on boxes read from a file and dumps the node array + primitive order, so the
// restated builders (oracle.cpp, csrc/rtb_scene.cpp) can be checked bit for bit.  Host only: the
// builders never touch the GPU (MakeHandle, which uploads, is not called).
//
//   ref_bvh <in.bin> <out.bin> <builder: 0 top-down median | 2 bottom-up>
//   in : int32 n, then n x 6 float (min.xyz, max.xyz)
//   out: int32 n_nodes, int32 root, n_nodes x 32-byte BVH::Node, n x int32 order
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <tuple>
#include <vector>
#include <cuda_runtime.h>
#include <glm/glm.hpp>

#include "rt_engine/geometry/BVH.cuh"

// The factory's results are private members without accessors (BVH.cuh:71-73).  This mirror has the
// same leading members in the same order, so the harness (and only the harness) can read them.
struct FactoryView {
	std::vector<BVH::Node> bvh_nodes;
	std::vector<const Hittable*> hittables;
	int root_idx;
};

int main(int argc, char** argv) {
	if (argc != 4) { fprintf(stderr, "usage: ref_bvh in.bin out.bin builder\n"); return 2; }
	FILE* f = fopen(argv[1], "rb"); if (!f) { perror("open in"); return 1; }
	int32_t n = 0; if (fread(&n, 4, 1, f) != 1 || n <= 0) return 1;
	std::vector<float> boxes(6 * (size_t)n);
	if (fread(boxes.data(), 4, boxes.size(), f) != boxes.size()) return 1;
	fclose(f);
	std::vector<std::tuple<aabb, const Hittable*>> arr;
	for (int i = 0; i < n; ++i) {
		aabb b(glm::vec3(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]), glm::vec3(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]));
		arr.push_back(std::make_tuple(b, reinterpret_cast<const Hittable*>((uintptr_t)(i + 1))));
	}
	BVH_Handle::Factory factory(arr);
	int builder = atoi(argv[3]);
	if (builder == 0) factory.BuildBVH_TopDown(); else factory.BuildBVH_BottomUp();
	static_assert(sizeof(BVH::Node) == 32, "BVH::Node layout");
	const FactoryView& view = *reinterpret_cast<const FactoryView*>(&factory);
	FILE* o = fopen(argv[2], "wb"); if (!o) { perror("open out"); return 1; }
	int32_t nn = (int32_t)view.bvh_nodes.size(), root = view.root_idx;
	fwrite(&nn, 4, 1, o); fwrite(&root, 4, 1, o);
	fwrite(view.bvh_nodes.data(), sizeof(BVH::Node), nn, o);
	for (int i = 0; i < n; ++i) { int32_t id = (int32_t)((uintptr_t)view.hittables[i]) - 1; fwrite(&id, 4, 1, o); }
	fclose(o);
	return 0;
}
