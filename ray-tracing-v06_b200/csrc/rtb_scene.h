// rtb_scene.h — host scene graph behind the C ABI (include/rtb.h) and its flattener.
#ifndef RTB_SCENE_H
#define RTB_SCENE_H

#include <string>
#include <vector>

#include "../../include/rtb.h"
#include "../../include/rtb_scene_format.h"
#include "rtb_types.h"

struct rtb_scene {
	std::vector<rtbs_texture> textures;
	std::vector<rtbs_material> materials;
	std::vector<rtbs_object> objects;
	std::vector<int32_t> children;
	std::vector<uint8_t> blob;
	int32_t root = -1;
	int32_t background_mode = RTB_BG_SKY_GRADIENT;
	int32_t world_bvh_mode = RTB_WORLD_BVH_QUALITY;
	uint64_t uid = 0;       // unique per rtb_scene_create
	uint64_t version = 0;   // bumped by every mutating call: lets a renderer skip re-flattening an unchanged scene
	float background[3] = {0, 0, 0};

	// Filled by flatten(): the world BVH in the reference's node layout (parity hook).
	std::vector<rtb_bvh_node> world_nodes;
	int32_t world_root = -1;
};

namespace rtb {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

struct Box3 { float mn[3], mx[3]; };

// BVH_Handle::Factory restated (BVH.cu:166-383). order[i] = input index at slot i.
int build_bvh(const std::vector<Box3>& boxes, int builder, std::vector<rtb_bvh_node>& nodes,
              std::vector<int>& order, int& root);
// Quality builder used for the world BVH when the scene root is not a reference BVH.
int build_bvh_world_sah(const std::vector<Box3>& boxes, std::vector<rtb_bvh_node>& nodes,
                        std::vector<int>& order, int& root);
int bvh_depth(const std::vector<rtb_bvh_node>& nodes, int root);

// Where RTB_WORLD_BVH_GPU_LBVH builds: the renderer's device and stream, and two grow-only scratch buffers the
// renderer keeps between builds (device memory; pinned host memory for the download), so that a rebuild allocates nothing.
struct GpuScratch { void* device = nullptr; size_t device_bytes = 0; void* pinned = nullptr; size_t pinned_bytes = 0; };
struct GpuBuildContext { int device; void* stream; GpuScratch* scratch; };

// GPU linear BVH (rtb_lbvh.cu): same outputs as build_bvh; inner nodes 0..n-2 (root 0), leaf of sorted position k = n-1+k.
int build_bvh_lbvh_gpu(const std::vector<Box3>& boxes, std::vector<rtb_bvh_node>& nodes, std::vector<int>& order,
                       int& root, const GpuBuildContext& gpu);
void free_gpu_scratch(GpuScratch& scratch);

// Flattens the object graph to world-space primitives + one world BVH.
int flatten(rtb_scene& s, FlatScene& out, const GpuBuildContext* gpu = nullptr);

}  // namespace rtb
#endif
