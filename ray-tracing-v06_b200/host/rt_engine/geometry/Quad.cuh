// Quads, triangles and boxes in the reference's Geometry / *Handle idiom (SphereHittable.cuh:35-158
// is the model).  Not present in the reference; semantics follow "Ray Tracing: The Next Week"
// quad / box.
#pragma once
#include <glm/glm.hpp>

#include "../../rtb_context.h"
#include "../shaders/material.cuh"
#include "aabb.cuh"
#include "hittable.cuh"

class Quad : public Geometry {
public:
	glm::vec3 Q, u, v;
	Quad() = default;
	Quad(glm::vec3 Q, glm::vec3 u, glm::vec3 v) : Q(Q), u(u), v(v) {}
};
class Triangle : public Geometry {
public:
	glm::vec3 Q, u, v;
	Triangle() = default;
	Triangle(glm::vec3 Q, glm::vec3 u, glm::vec3 v) : Q(Q), u(u), v(v) {}
};
class Box : public Geometry {
public:
	glm::vec3 a, b;
	Box() = default;
	Box(glm::vec3 a, glm::vec3 b) : a(a), b(b) {}
};

class GeoHittable : public Hittable {
public:
	explicit GeoHittable(int id) : Hittable(id) {}
};

// Owns the material (like SphereHandle) and the host proxy of the registered object.
class GeoHandle {
	aabb bounds;
	Material* material_ptr{};
	Hittable* hittable_ptr{};
	GeoHandle() = default;
	GeoHandle(const GeoHandle&) = delete;
	GeoHandle& operator=(const GeoHandle&) = delete;
	static GeoHandle _finish(int id, Material* mat, bool own) {
		GeoHandle h{};
		float b[6]; rtb_host::check(rtb_object_bounds(rtb_host::scene(), id, b), "rtb_object_bounds");
		h.bounds = aabb(glm::vec3(b[0], b[1], b[2]), glm::vec3(b[3], b[4], b[5]));
		h.material_ptr = own ? mat : nullptr;
		h.hittable_ptr = new GeoHittable(id);
		return h;
	}

public:
	GeoHandle(GeoHandle&& o) noexcept : bounds(o.bounds), material_ptr(o.material_ptr), hittable_ptr(o.hittable_ptr) { o.material_ptr = nullptr; o.hittable_ptr = nullptr; }
	~GeoHandle() { delete material_ptr; delete hittable_ptr; }

	template <typename MatType> requires GeoAcceptableMat<Quad, MatType>
	static GeoHandle MakeQuad(const Quad& q, MatType* mat, bool take_ownership = false) {
		return _finish(rtb_host::check(rtb_add_quad(rtb_host::scene(), &q.Q.x, &q.u.x, &q.v.x, mat->rtb_material), "MakeQuad"), mat, take_ownership);
	}
	template <typename MatType> requires GeoAcceptableMat<Triangle, MatType>
	static GeoHandle MakeTriangle(const Triangle& q, MatType* mat, bool take_ownership = false) {
		return _finish(rtb_host::check(rtb_add_triangle(rtb_host::scene(), &q.Q.x, &q.u.x, &q.v.x, mat->rtb_material), "MakeTriangle"), mat, take_ownership);
	}
	template <typename MatType> requires GeoAcceptableMat<Box, MatType>
	static GeoHandle MakeBox(const Box& bx, MatType* mat, bool take_ownership = false) {
		return _finish(rtb_host::check(rtb_add_box(rtb_host::scene(), &bx.a.x, &bx.b.x, mat->rtb_material), "MakeBox"), mat, take_ownership);
	}
	// translate / rotate_y / constant_medium wrap an existing hittable (book instances)
	static GeoHandle MakeTranslate(const Hittable* child, glm::vec3 offset) {
		return _finish(rtb_host::check(rtb_add_translate(rtb_host::scene(), child->rtb_object, &offset.x), "MakeTranslate"), nullptr, false);
	}
	static GeoHandle MakeRotateY(const Hittable* child, float degrees) {
		return _finish(rtb_host::check(rtb_add_rotate_y(rtb_host::scene(), child->rtb_object, degrees), "MakeRotateY"), nullptr, false);
	}
	template <typename MatType> requires GeoIndependantMat<MatType>
	static GeoHandle MakeConstantMedium(const Hittable* boundary, float density, MatType* phase, bool take_ownership = false) {
		return _finish(rtb_host::check(rtb_add_constant_medium(rtb_host::scene(), boundary->rtb_object, density, phase->rtb_material), "MakeConstantMedium"), phase, take_ownership);
	}
	const Hittable* getHittablePtr() const { return hittable_ptr; }
	aabb getBounds() const { return bounds; }
};
