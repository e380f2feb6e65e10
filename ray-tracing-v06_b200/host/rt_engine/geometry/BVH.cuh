// Host mirror of main/src/rt_engine/geometry/BVH.{cuh,cu}: BVH::Node (same 32-byte layout,
// BVH.cuh:16-25), BVH_Handle and BVH_Handle::Factory with BuildBVH_TopDown / BuildBVH_BottomUp /
// MakeHandle (BVH.cuh:69-102).  The builders themselves are restated behind rtb_bvh_build
// (csrc/rtb_scene.cpp) and reproduce the reference's node and primitive order bit for bit; the
// traversal (BVH.cu:54-106) runs in the traverse kernel.
#pragma once
#include <tuple>
#include <vector>

#include "../../rtb_context.h"
#include "aabb.cuh"
#include "hittable.cuh"

class BVH : public Hittable {
public:
#define _IS_LEAF_CODE (-1)
	struct Node {
		aabb bounds;
		int left_child_idx;
		int right_child_hittable_idx;
		bool isLeaf() const { return left_child_idx == _IS_LEAF_CODE; }
	};
	explicit BVH(int id) : Hittable(id) {}
};
static_assert(sizeof(BVH::Node) == sizeof(rtb_bvh_node), "BVH::Node must keep the reference's 32-byte layout");

class BVH_Handle {
	BVH* d_bvh{};
	aabb bounds;
	BVH_Handle(aabb bounds, int object) : d_bvh(new BVH(object)), bounds(bounds) {}
	BVH_Handle(const BVH_Handle&) = delete;
	BVH_Handle& operator=(const BVH_Handle&) = delete;

public:
	class Factory;
	~BVH_Handle() { delete d_bvh; }
	BVH_Handle(BVH_Handle&& o) noexcept : d_bvh(o.d_bvh), bounds(o.bounds) { o.d_bvh = nullptr; }
	BVH_Handle& operator=(BVH_Handle&& o) noexcept { if (this != &o) { delete d_bvh; d_bvh = o.d_bvh; bounds = o.bounds; o.d_bvh = nullptr; } return *this; }
	const BVH* getBVHPtr() const { return d_bvh; }
	aabb getBounds() const { return bounds; }
};

class BVH_Handle::Factory {
	std::vector<BVH::Node> bvh_nodes;
	std::vector<const Hittable*> hittables;
	int root_idx = -1;
	int builder = RTB_BVH_TOPDOWN_MEDIAN;
	std::vector<std::tuple<aabb, const Hittable*>>& arr;
	std::vector<int> input_ids;   // children in the caller's order, before the builder permutes `arr`

	void _build(int which) {
		builder = which;
		const int n = (int)arr.size();
		std::vector<float> boxes(6 * (size_t)n);
		for (int i = 0; i < n; ++i) {
			glm::vec3 mn = std::get<0>(arr[i]).getMin(), mx = std::get<0>(arr[i]).getMax();
			boxes[6 * i] = mn.x; boxes[6 * i + 1] = mn.y; boxes[6 * i + 2] = mn.z; boxes[6 * i + 3] = mx.x; boxes[6 * i + 4] = mx.y; boxes[6 * i + 5] = mx.z;
		}
		std::vector<rtb_bvh_node> nodes(2 * (size_t)n);
		std::vector<int> order(n);
		int count = rtb_host::check(rtb_bvh_build(boxes.data(), n, which, nodes.data(), order.data(), &root_idx), "BVH_Handle::Factory");
		bvh_nodes.resize(count);
		for (int i = 0; i < count; ++i) {
			bvh_nodes[i].bounds = aabb(glm::vec3(nodes[i].bmin[0], nodes[i].bmin[1], nodes[i].bmin[2]), glm::vec3(nodes[i].bmax[0], nodes[i].bmax[1], nodes[i].bmax[2]));
			bvh_nodes[i].left_child_idx = nodes[i].left_child_idx;
			bvh_nodes[i].right_child_hittable_idx = nodes[i].right_child_hittable_idx;
		}
		// the reference sorts `arr` in place (BVH.cu:195) and then copies it out (BVH.cu:174-177)
		std::vector<std::tuple<aabb, const Hittable*>> sorted(n);
		for (int i = 0; i < n; ++i) sorted[i] = arr[order[i]];
		arr = sorted;
		hittables.clear();
		for (int i = 0; i < n; ++i) hittables.push_back(std::get<1>(arr[i]));
	}
	Factory(Factory&) = delete;
	Factory& operator=(Factory&) = delete;

public:
	Factory(std::vector<std::tuple<aabb, const Hittable*>>& arr) : arr(arr) {
		for (auto& t : arr) input_ids.push_back(std::get<1>(t)->rtb_object);
	}
	void BuildBVH_TopDown() { _build(RTB_BVH_TOPDOWN_MEDIAN); }
	void BuildBVH_TopDownSAH() { _build(RTB_BVH_TOPDOWN_SAH); }   // _build_bvh_rec2, selectable here instead of by #if (BVH.cu:168)
	void BuildBVH_BottomUp() { _build(RTB_BVH_BOTTOMUP); }
	BVH_Handle* MakeHandle() {
		int id = rtb_host::check(rtb_add_bvh(rtb_host::scene(), input_ids.data(), (int)input_ids.size(), builder), "BVH_Handle::Factory::MakeHandle");
		return new BVH_Handle(bvh_nodes[root_idx].bounds, id);
	}
	// read-only views for tests (the reference keeps these private, BVH.cuh:71-73)
	const std::vector<BVH::Node>& nodes() const { return bvh_nodes; }
	const std::vector<const Hittable*>& sorted_hittables() const { return hittables; }
	int root() const { return root_idx; }
};
