// Host mirror of main/src/rt_engine/geometry/bvh_node.cuh:9-24: the pointer-tree node (bounds test, then
// both children).  For a closest-hit query that is a two-element list, which is how it is registered.
#pragma once
#include "../../rtb_context.h"
#include "aabb.cuh"
#include "hittable.cuh"

class bvh_node : public Hittable {
	aabb bounds;

public:
	bvh_node(const Hittable* left, const Hittable* right, const aabb& bounds) : bounds(bounds) {
		int ids[2] = {left->rtb_object, right->rtb_object};
		rtb_object = rtb_host::check(rtb_add_list(rtb_host::scene(), ids, 2), "bvh_node");
	}
	aabb getBounds() const { return bounds; }
};
