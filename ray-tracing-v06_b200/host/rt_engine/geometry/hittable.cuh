// Host mirror of main/src/rt_engine/geometry/hittable.cuh:9-31.  A Hittable* handed around by scene
// code is a host proxy carrying the rtb object id; the device never sees it.
#pragma once
#include <concepts>

#include "../ray_data.cuh"

class Hittable {
protected:
	Hittable() = default;
	Hittable(const Hittable&) = default;
	Hittable& operator=(const Hittable&) = default;
	explicit Hittable(int id) : rtb_object(id) {}

public:
	virtual ~Hittable() = default;
	int rtb_object = -1;
};

class Geometry {};
template <typename T>
concept Geometry_t = std::derived_from<T, Geometry>;
