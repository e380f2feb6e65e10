// Host mirror of main/src/rt_engine/geometry/SphereHittable.{cuh,cu}: Sphere / MovingSphere PODs,
// their bounds helpers and the move-only SphereHandle factory (SphereHittable.cuh:35-158).  The
// intersection routine (_sphere_closest_intersection, :15-33) runs in the traverse kernel.
#pragma once
#include <glm/glm.hpp>

#include "../../rtb_context.h"
#include "../shaders/material.cuh"
#include "aabb.cuh"
#include "hittable.cuh"

class Sphere : public Geometry {
public:
	glm::vec3 center;
	float radius;
	Sphere() = default;
	Sphere(glm::vec3 center, float radius) : center(center), radius(radius) {}
};
inline aabb getSphereBounds(const Sphere& sp) { return aabb(sp.center - glm::vec3(sp.radius), sp.center + glm::vec3(sp.radius)); }

class MovingSphere : public Geometry {
public:
	glm::vec3 center0, center1;
	float radius;
	MovingSphere() = default;
	MovingSphere(glm::vec3 center0, glm::vec3 center1, float radius) : center0(center0), center1(center1), radius(radius) {}
};
inline aabb getMovingSphereBounds(const MovingSphere& sp) {
	aabb t0(sp.center0 - glm::vec3(sp.radius), sp.center0 + glm::vec3(sp.radius));
	aabb t1(sp.center1 - glm::vec3(sp.radius), sp.center1 + glm::vec3(sp.radius));
	return aabb(t0, t1);
}

class SphereHittable : public Hittable {
public:
	explicit SphereHittable(int id) : Hittable(id) {}
};
class MovingSphereHittable : public Hittable {
public:
	explicit MovingSphereHittable(int id) : Hittable(id) {}
};

class SphereHandle {
	aabb bounds;
	Material* material_ptr{};
	Hittable* hittable_ptr{};
	SphereHandle() = default;
	void _delete() { delete material_ptr; delete hittable_ptr; material_ptr = nullptr; hittable_ptr = nullptr; }
	SphereHandle(const SphereHandle&) = delete;
	SphereHandle& operator=(const SphereHandle&) = delete;

public:
	SphereHandle(SphereHandle&& sp) noexcept : bounds(sp.bounds), material_ptr(sp.material_ptr), hittable_ptr(sp.hittable_ptr) { sp.material_ptr = nullptr; sp.hittable_ptr = nullptr; }
	SphereHandle& operator=(SphereHandle&& sp) noexcept {
		if (this != &sp) { _delete(); bounds = sp.bounds; material_ptr = sp.material_ptr; hittable_ptr = sp.hittable_ptr; sp.material_ptr = nullptr; sp.hittable_ptr = nullptr; }
		return *this;
	}
	~SphereHandle() { _delete(); }

	// The handle takes ownership of the material, as in the reference (SphereHittable.cu:106-116);
	// pass take_ownership = false for a material shared by several objects.
	template <typename MatType> requires GeoAcceptableMat<Sphere, MatType>
	static SphereHandle MakeSphere(const Sphere& sphere, MatType* mat_ptr, bool take_ownership = true) {
		SphereHandle sp{};
		sp.bounds = getSphereBounds(sphere);
		sp.material_ptr = take_ownership ? mat_ptr : nullptr;
		int id = rtb_host::check(rtb_add_sphere(rtb_host::scene(), &sphere.center.x, sphere.radius, mat_ptr->rtb_material), "MakeSphere");
		sp.hittable_ptr = new SphereHittable(id);
		return sp;
	}
	template <typename MatType> requires GeoAcceptableMat<MovingSphere, MatType>
	static SphereHandle MakeMovingSphere(const MovingSphere& sphere, MatType* mat_ptr, bool take_ownership = true) {
		SphereHandle sp{};
		sp.bounds = getMovingSphereBounds(sphere);
		sp.material_ptr = take_ownership ? mat_ptr : nullptr;
		int id = rtb_host::check(rtb_add_moving_sphere(rtb_host::scene(), &sphere.center0.x, &sphere.center1.x, sphere.radius, mat_ptr->rtb_material), "MakeMovingSphere");
		sp.hittable_ptr = new MovingSphereHittable(id);
		return sp;
	}
	const Hittable* getHittablePtr() const { return hittable_ptr; }
	aabb getBounds() const { return bounds; }
};
