// Host mirror of main/src/rt_engine/ray_data.cuh:8-17 (the payload structs are device-side detail
// of the reference's megakernel and have no counterpart here).
#pragma once
#include <glm/glm.hpp>

struct Ray {
	glm::vec3 o{0, 0, 0}, d{0, 0, 1};
	float time{0.0f};
	Ray() = default;
	Ray(glm::vec3 origin, glm::vec3 direction, float time = 0.0f) : o(origin), d(direction), time(time) {}
	glm::vec3 at(float t) const { return o + d * t; }
};
#define _MISS_DIST 3.402823466e+38F
