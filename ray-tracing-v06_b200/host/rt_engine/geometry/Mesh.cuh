// Triangle meshes in the reference's *Handle idiom.  The reference's RT engine has no mesh type; its GL
// demo loads Wavefront OBJ files through tiny_obj (main/src/gl_engine/gl_mesh.cpp:124-218).  This loader
// reads the same subset that demo uses — `v x y z` and `f a b c ...` records (1-based or negative indices,
// `i/j/k` corner syntax, polygons fanned into triangles) — and registers one Triangle per face under a
// BVH group, so a mesh is an ordinary Hittable that can be listed, instanced or put in a BVH.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <glm/glm.hpp>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../rtb_context.h"
#include "../shaders/material.cuh"
#include "Quad.cuh"
#include "aabb.cuh"
#include "hittable.cuh"

class Mesh : public Geometry {
public:
	std::vector<glm::vec3> vertices;
	std::vector<int> indices;   // 3 per triangle, 0-based
};

class MeshHandle {
	aabb bounds;
	Hittable* hittable_ptr{};
	int triangle_count = 0;
	MeshHandle() = default;
	MeshHandle(const MeshHandle&) = delete;
	MeshHandle& operator=(const MeshHandle&) = delete;

public:
	MeshHandle(MeshHandle&& o) noexcept : bounds(o.bounds), hittable_ptr(o.hittable_ptr), triangle_count(o.triangle_count) { o.hittable_ptr = nullptr; }
	~MeshHandle() { delete hittable_ptr; }

	// Reads the `v` / `f` records of a Wavefront OBJ file.
	static Mesh LoadObj(const std::string& path) {
		FILE* f = fopen(path.c_str(), "r");
		if (!f) throw std::runtime_error("MeshHandle::LoadObj: cannot open " + path);
		Mesh m; char line[1024];
		while (fgets(line, sizeof line, f)) {
			if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
				float x, y, z;
				if (sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3) m.vertices.push_back(glm::vec3(x, y, z));
			} else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
				std::vector<int> corner;
				for (char* tok = strtok(line + 2, " \t\r\n"); tok; tok = strtok(nullptr, " \t\r\n")) {
					int i = atoi(tok);                                  // "i", "i/j", "i//k", "i/j/k": the vertex index comes first
					if (i < 0) i = (int)m.vertices.size() + i + 1;      // negative: relative to the vertices read so far
					if (i < 1 || i > (int)m.vertices.size()) { fclose(f); throw std::runtime_error("MeshHandle::LoadObj: bad face index in " + path); }
					corner.push_back(i - 1);
				}
				for (size_t k = 2; k < corner.size(); ++k) { m.indices.push_back(corner[0]); m.indices.push_back(corner[k - 1]); m.indices.push_back(corner[k]); }
			}
		}
		fclose(f);
		if (m.indices.empty()) throw std::runtime_error("MeshHandle::LoadObj: no faces in " + path);
		return m;
	}

	template <typename MatType> requires GeoAcceptableMat<Triangle, MatType>
	static MeshHandle MakeMesh(const Mesh& mesh, MatType* mat) {
		MeshHandle h{};
		static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "vertices are handed over as packed floats");
		const int group = rtb_host::check(rtb_add_mesh(rtb_host::scene(), &mesh.vertices[0].x, (int)mesh.vertices.size(), mesh.indices.data(),
		                                               (int)(mesh.indices.size() / 3), mat->rtb_material), "MakeMesh");
		const int faces = rtb_host::check(rtb_scene_num_children(rtb_host::scene(), group), "MakeMesh");
		float bb[6]; rtb_host::check(rtb_object_bounds(rtb_host::scene(), group, bb), "rtb_object_bounds");
		h.bounds = aabb(glm::vec3(bb[0], bb[1], bb[2]), glm::vec3(bb[3], bb[4], bb[5]));
		h.hittable_ptr = new GeoHittable(group);
		h.triangle_count = faces;
		return h;
	}
	const Hittable* getHittablePtr() const { return hittable_ptr; }
	aabb getBounds() const { return bounds; }
	int triangleCount() const { return triangle_count; }
};
