// Host mirror of main/src/rt_engine/geometry/aabb.cuh:11-68 (bounds bookkeeping only; the slab test
// lives in the traverse kernel).
#pragma once
#include <glm/glm.hpp>

class aabb {
	glm::vec3 min, max;

public:
	aabb() : min(1e9f), max(-1e9f) {}
	aabb(glm::vec3 min, glm::vec3 max) : min(min), max(max) {}
	aabb(const aabb& a, const aabb& b) : min(glm::min(a.min, b.min)), max(glm::max(a.max, b.max)) {}
	glm::vec3 getMin() const { return min; }
	glm::vec3 getMax() const { return max; }
	aabb& operator+=(const aabb& b) { min = glm::min(min, b.min); max = glm::max(max, b.max); return *this; }
	int longest_axis() const {
		glm::vec3 span = glm::abs(max - min);
		if (span.x > span.y) return span.x > span.z ? 0 : 2;
		return span.y > span.z ? 1 : 2;
	}
	glm::vec3 centeroid() const { return (max + min) * 0.5f; }
};
